# SPHtoGridCUDA.jl — the reference-side binding of libsphtogrid_cuda.so.
#
# This is the file a SPHtoGrid.jl maintainer would `include` (after the existing includes in src/SPHtoGrid.jl) to
# swap the bodies of the deposit hot path for calls into the B200 library.  It re-defines, with IDENTICAL signatures,
#
#   cic_mapping_2D            (src/cic_interpolation/cic_2D.jl:103-111)
#   cic_mapping_3D            (src/cic_interpolation/cic_3D.jl:110-115)
#   reduce_image_2D / _3D     (src/cic_interpolation/reduce_image.jl:8-10, :39-40)
#   the particle loop of healpix_map (src/healpix_interpolation/main.jl:143-213) as `healpix_deposit!`
#
# so that `sphMapping`, `map_it`, `healpix_map`, `distributed_cic_map` ... keep working unchanged on top of them.
# Julia is not installed in the build environment of this repository, so this file has not been executed here;
# it is a mechanical transcription of the ctypes binding in ../_lib.py, which IS exercised by the test-suite.
#
# Memory layout notes: Julia `Matrix{T}(3,N)` positions, `Vector{T}(N)` fields and `Matrix{Float64}(Npix, Nimg+1)`
# images are passed as-is (column-major = what the C ABI documents); nothing is copied on the Julia side.

module SPHtoGridCUDA

using SPHKernels

const LIB = get(ENV, "SPHTOGRID_CUDA_LIB", "libsphtogrid_cuda")

const S2G_F32 = Int32(0)
const S2G_F64 = Int32(1)

struct S2GStats
    n_in::Int64; n_mapped::Int64; footprint_pixels::Int64; touched_pixels::Int64; n_fallback::Int64
    n_pairs::Int64; n_scatter::Int64; n_gather::Int64; n_launches::Int64
    ms_h2d::Float64; ms_compute::Float64; ms_d2h::Float64; ms_total::Float64
    ms_prep::Float64; ms_sort::Float64; ms_norm::Float64; ms_deposit::Float64; ms_epilogue::Float64
end

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:s2g_init, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        ctx = new(h[])
        finalizer(c -> ccall((:s2g_shutdown, LIB), Cint, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

const _default_ctx = Ref{Union{Nothing,Context}}(nothing)
default_context() = something(_default_ctx[], (_default_ctx[] = Context(0)))

last_error() = unsafe_string(ccall((:s2g_last_error, LIB), Cstring, ()))
check(rc::Integer) = rc == 0 ? nothing : error("libsphtogrid_cuda ($rc): " * last_error())

kernel_id(::Cubic) = Int32(0)
kernel_id(::Quintic) = Int32(1)
kernel_id(::WendlandC2) = Int32(2)
kernel_id(::WendlandC4) = Int32(3)
kernel_id(::WendlandC6) = Int32(4)
kernel_id(::WendlandC8) = Int32(5)

# all six particle arrays must share one element type across the ABI; promotion to Float64 is exact
function _uniform(Pos, HSML, M, Rho, Bin_Q, Weights)
    T = (eltype(Pos) == Float32 && all(a -> eltype(a) == Float32, (HSML, M, Rho, Bin_Q, Weights))) ? Float32 : Float64
    conv(a) = eltype(a) == T ? a : convert(Array{T}, a)
    return T, conv(Pos), conv(HSML), conv(M), conv(Rho), conv(Bin_Q), conv(Weights)
end

"""
    _uniform_fused!(Pos, HSML, M, Rho, Bin_Q, Weights, par, ctx) -> (T, pos, hsml, m, rho, bq, w, shift, periodic, boxsize, writeback)

Element type rule of the fused calls: `Float32` only if EVERY array is `Float32`, otherwise everything is widened to
`Float64` — never narrowed (the reference promotes: `bin_q = Float64(Bin_Q[p])`, cic_2D.jl:186-199; `get_quantities_2D`
multiplies by the Float64 `len2pix`).  With Float32 positions and Float64 fields the positions are first recentred in
Float32 (center_particles works in the precision of `Pos`, filter_shift.jl:15 — quirk Q2), written back into `Pos`,
and the widened copy is then mapped with a zero shift.
"""
function _uniform_fused!(Pos::Matrix, HSML, M, Rho, Bin_Q, Weights, par, ctx)
    T = (eltype(Pos) == Float32 && all(a -> eltype(a) == Float32, (HSML, M, Rho, Bin_Q, Weights))) ? Float32 : Float64
    conv(a) = eltype(a) == T ? a : convert(Array{T}, a)
    shift = Float64.(par.center); periodic = par.periodic; boxsize = Float64(par.boxsize)
    if eltype(Pos) == T
        return T, Pos, conv(HSML), conv(M), conv(Rho), conv(Bin_Q), conv(Weights), shift, periodic, boxsize, true
    end
    # mixed: recentre in the precision of Pos on the device (mask not needed), then widen
    N = size(Pos, 2)
    mask = Vector{UInt8}(undef, N); pos_c = similar(Pos); zero3 = zeros(3); big3 = fill(Inf, 3)
    GC.@preserve Pos pos_c mask shift zero3 big3 begin
        check(ccall((:s2g_center_filter, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Int32, Float64, Ptr{Float64}, Ptr{Float64},
                     Ptr{Cvoid}, Ptr{UInt8}),
                    ctx.handle, Pos, N, eltype(Pos) == Float32 ? S2G_F32 : S2G_F64, shift, periodic, boxsize,
                    zero3, big3, pos_c, mask))
    end
    Pos .= pos_c
    return T, convert(Matrix{T}, Pos), conv(HSML), conv(M), conv(Rho), conv(Bin_Q), conv(Weights), zeros(3), false,
           -1.0, false
end

"""
    cic_mapping_2D(Pos, HSML, M, Rho, Bin_Q, Weights, RM=nothing; param, kernel, show_progress, calc_mean, stokes)

Drop-in for src/cic_interpolation/cic_2D.jl:103-244.  Returns `Matrix{Float64}(Nx*Ny, N_images+1)`.
"""
function cic_mapping_2D(Pos, HSML, M, Rho, Bin_Q, Weights, RM=nothing;
                        param, kernel::AbstractSPHKernel, show_progress::Bool=false,
                        calc_mean::Bool=true, stokes::Bool=false, ctx::Context=default_context())
    T, pos, hsml, m, rho, bq, w = _uniform(Pos, HSML, M, Rho, Bin_Q, Weights)
    N = length(m)
    n_images = ndims(bq) == 1 ? 1 : size(bq, 1)
    nx, ny = param.Npixels[1], param.Npixels[2]
    image = Matrix{Float64}(undef, nx * ny, n_images + 1)
    if !isnothing(RM)
        # Faraday-rotation branch (cic_2D.jl:201-217, cic_shared.jl:129-159): ordered compositing on the device;
        # stokes=false leaves RM inert exactly like faraday_rotate_pixel! does
        rm = RM::Vector{Float64}
        GC.@preserve pos hsml m rho bq w rm image begin
            check(ccall((:s2g_deposit_2d_rm, LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                         Ptr{Float64}, Int64, Int32, Int32, Float64, Int64, Int64, Int32, Int32, Int32, Ptr{Float64},
                         Ptr{Cvoid}),
                        ctx.handle, pos, hsml, m, rho, bq, w, rm, N, n_images, T == Float32 ? S2G_F32 : S2G_F64,
                        Float64(param.len2pix), nx, ny, kernel_id(kernel), calc_mean, stokes, image, C_NULL))
        end
        return image
    end
    GC.@preserve pos hsml m rho bq w image begin
        check(ccall((:s2g_deposit_2d, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                     Int64, Int32, Int32, Float64, Int64, Int64, Int32, Int32, Ptr{Float64}, Ptr{Cvoid}),
                    ctx.handle, pos, hsml, m, rho, bq, w, N, n_images, T == Float32 ? S2G_F32 : S2G_F64,
                    Float64(param.len2pix), nx, ny, kernel_id(kernel), calc_mean, image, C_NULL))
    end
    return image
end

"""
    cic_mapping_3D(Pos, HSML, M, Rho, Bin_Q, Weights; param, kernel, show_progress, calc_mean)

Drop-in for src/cic_interpolation/cic_3D.jl:110-209.  Returns `Matrix{Float64}(N^3, 2)`.
"""
function cic_mapping_3D(Pos, HSML, M, Rho, Bin_Q, Weights;
                        param, kernel::AbstractSPHKernel, show_progress::Bool=false, calc_mean=false,
                        ctx::Context=default_context())
    T, pos, hsml, m, rho, bq, w = _uniform(Pos, HSML, M, Rho, Bin_Q, Weights)
    N = length(m)
    n = param.Npixels[1]
    image = Matrix{Float64}(undef, n^3, 2)
    GC.@preserve pos hsml m rho bq w image begin
        check(ccall((:s2g_deposit_3d, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                     Int64, Int32, Float64, Int64, Int32, Int32, Ptr{Float64}, Ptr{Cvoid}),
                    ctx.handle, pos, hsml, m, rho, bq, w, N, T == Float32 ? S2G_F32 : S2G_F64,
                    Float64(param.len2pix), n, kernel_id(kernel), calc_mean, image, C_NULL))
    end
    return image
end

"reduce_image.jl:8-31"
function reduce_image_2D(image::Matrix{<:Real}, x_pixels::Int64, y_pixels::Int64, reduce_image::Bool;
                         ctx::Context=default_context())
    img = convert(Matrix{Float64}, image)
    n_images = size(img, 2) - 1
    out = Array{Float64,3}(undef, y_pixels, x_pixels, n_images)
    GC.@preserve img out check(ccall((:s2g_reduce_image_2d, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int32, Int32, Ptr{Float64}),
        ctx.handle, img, x_pixels, y_pixels, n_images, reduce_image, out))
    return out
end

"reduce_image.jl:39-55 (the caller has already set image[:,2] .= 1 when !reduce_image, so reduce_image=true here)"
function reduce_image_3D(image::Matrix{<:Real}, x_pixels::Int64, y_pixels::Int64, z_pixels::Int64;
                         ctx::Context=default_context())
    img = convert(Matrix{Float64}, image)
    out = Array{Float64,3}(undef, z_pixels, y_pixels, x_pixels)
    GC.@preserve img out check(ccall((:s2g_reduce_image_3d, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Float64}),
        ctx.handle, img, x_pixels, Int32(1), out))
    return out
end

"""
    healpix_deposit!(allsky_map, weight_map, pos, hsml, m, rho, bin_q, weights, Nside, kernel, calc_mean)

Replaces the `for ipart ∈ 1:Npart` loop of healpix_map (src/healpix_interpolation/main.jl:143-213); the two
`HealpixMap{Float64,RingOrder}` are filled through their `.pixels` vectors.
"""
function healpix_deposit!(allsky_map, weight_map, pos::Matrix{Float64}, hsml::Vector{Float64}, m::Vector{Float64},
                          rho::Vector{Float64}, bin_q::Vector{Float64}, weights::Vector{Float64},
                          Nside::Integer, kernel::AbstractSPHKernel, calc_mean::Bool;
                          ctx::Context=default_context())
    a, w = allsky_map.pixels, weight_map.pixels
    GC.@preserve pos hsml m rho bin_q weights a w begin
        check(ccall((:s2g_healpix_deposit, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                     Int64, Int32, Int64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
                    ctx.handle, pos, hsml, m, rho, bin_q, weights, length(hsml), S2G_F64, Nside,
                    kernel_id(kernel), calc_mean, a, w, C_NULL))
    end
    return allsky_map, weight_map
end

"""
    healpix_map_fused!(allsky_map, weight_map, Pos, Hsml, M, Rho, Bin_q, Weights; center, radius_limits, Nside,
                       kernel, calc_mean)

The whole body of `healpix_map` after the map allocation (src/healpix_interpolation/main.jl:123-213) in ONE call:
`Pos .-= center` (Pos is mutated like `filter_sort_particles` does), shell filter, the far-to-near selection
`sorted[sel]` and the particle loop, all on the device (s2g_healpix_map).  The `calc_mean == false` BoundsError of
filter_particles.jl:28-30 must be raised by the caller before (it depends only on `Bin_q` and the shell mask).
"""
function healpix_map_fused!(allsky_map, weight_map, Pos::Matrix{Float64}, Hsml::Vector{Float64}, M::Vector{Float64},
                            Rho::Vector{Float64}, Bin_q::Vector{Float64}, Weights::Vector{Float64};
                            center::Vector{<:Real}, radius_limits::Vector{<:Real}, Nside::Integer,
                            kernel::AbstractSPHKernel, calc_mean::Bool=true, ctx::Context=default_context())
    a, w = allsky_map.pixels, weight_map.pixels
    cen = Float64.(center); rl = Float64.(radius_limits)
    pos_out = similar(Pos)
    GC.@preserve Pos Hsml M Rho Bin_q Weights a w cen rl pos_out begin
        check(ccall((:s2g_healpix_map, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64,
                     Ptr{Float64}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
                    ctx.handle, Pos, Hsml, M, Rho, Bin_q, Weights, length(Hsml), cen, rl, Nside, kernel_id(kernel),
                    calc_mean, pos_out, a, w, C_NULL))
    end
    Pos .= pos_out
    return allsky_map, weight_map
end

"""
    sphmap_fused(Pos, HSML, M, Rho, Bin_Q, Weights; param, par_centred, kernel, dimensions, calc_mean,
                 reduce_image, return_both_maps)

The whole body of `sphMapping` (centre -> filter -> deposit -> reduce) in ONE call (s2g_sphmap); `Pos` is
recentred in place exactly like center_particles does (src/cic_interpolation/filter_shift.jl:15).
"""
function sphmap_fused(Pos::Matrix{T}, HSML, M, Rho, Bin_Q, Weights; param, par_centred, kernel, dimensions::Int=2,
                      calc_mean::Bool=false, reduce_image::Bool=true, return_both_maps::Bool=false,
                      ctx::Context=default_context()) where {T<:Union{Float32,Float64}}
    U, pos, hsml, m, rho, bq, w, shift, periodic, boxsize, writeback =
        _uniform_fused!(Pos, HSML, M, Rho, Bin_Q, Weights, param, ctx)
    N = length(m)
    n_images = ndims(bq) == 1 ? 1 : size(bq, 1)
    n = par_centred.Npixels[1]
    out = dimensions == 2 ? (return_both_maps ? Matrix{Float64}(undef, n * n, n_images + 1) :
                                                Array{Float64,3}(undef, n, n, n_images)) :
                            Array{Float64,3}(undef, n, n, n)
    half = Float64.(par_centred.halfsize)
    pos_out = writeback ? similar(pos) : pos
    GC.@preserve pos hsml m rho bq w out shift half pos_out begin
        check(ccall((:s2g_sphmap, LIB), Cint,
                    (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                     Int64, Int32, Int32, Ptr{Float64}, Int32, Float64, Ptr{Float64}, Float64, Int64, Int32, Int32,
                     Int32, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}),
                    ctx.handle, dimensions, pos, hsml, m, rho, bq, w, N, n_images, U == Float32 ? S2G_F32 : S2G_F64,
                    shift, periodic, boxsize, half, Float64(par_centred.len2pix), n,
                    kernel_id(kernel), calc_mean, reduce_image, return_both_maps,
                    writeback ? pointer(pos_out) : C_NULL, out, C_NULL))
    end
    writeback && (Pos .= pos_out)
    return out
end

"""
    sphmap_projected(pos_in, HSML, M, Rho, Bin_Q, Weights; projection, param, kernel, ...)

`map_it`'s projection pre-step (cic_interpolation.jl:331-345) fused into the deposit: `projection` is "xy", "xz",
"yz" (axis permutation, exact) or a vector of three Euler angles in degrees (rotate_3D).  `pos_in` is only read —
no `copy(pos_in)`, no rotated array.  `param` is the UNROTATED map; the rotated parameters are built here with the
reference's own rotate_to_xz_plane(par) / rotate_to_yz_plane(par).
"""
function sphmap_projected(pos_in::Matrix{T}, HSML, M, Rho, Bin_Q, Weights; projection="xy", param, kernel,
                          dimensions::Int=2, calc_mean::Bool=true, reduce_image::Bool=true,
                          ctx::Context=default_context()) where {T<:Union{Float32,Float64}}
    perm = C_NULL; rot = C_NULL; par = param
    if projection == "xz"
        perm = Int32[0, 2, 1]; par = rotate_to_xz_plane(param)
    elseif projection == "yz"
        perm = Int32[1, 2, 0]; par = rotate_to_yz_plane(param)
    elseif projection isa AbstractVector
        R = RotXYZ(deg2rad.(projection)...)
        rot = Float64[R[1, 1], R[1, 2], R[1, 3], R[2, 1], R[2, 2], R[2, 3], R[3, 1], R[3, 2], R[3, 3]]  # row-major
    elseif projection != "xy"
        error("projection must be either along in 'xy', 'xz', or 'yz' plane of defined by a vector of Euler angles!")
    end
    _, par_centred = center_particles(Matrix{T}(undef, 3, 0), par)   # only the recentred parameters are needed
    if !(T == Float32 && all(a -> eltype(a) == Float32, (HSML, M, Rho, Bin_Q, Weights))) && T != Float64
        # Float32 positions with Float64 fields: never narrow the fields.  Rotate/permute + recentre a COPY in Float32
        # like map_it does (cic_interpolation.jl:327-345, filter_shift.jl:15), then run the unprojected fused call.
        pos = copy(pos_in)
        if perm !== C_NULL
            pos = pos[perm .+ 1, :]
        elseif rot !== C_NULL
            pos = Matrix{T}(reshape(rot, 3, 3)' * pos)
        end
        return sphmap_fused(pos, HSML, M, Rho, Bin_Q, Weights; param=par, par_centred=par_centred, kernel=kernel,
                            dimensions=dimensions, calc_mean=calc_mean, reduce_image=reduce_image, ctx=ctx)
    end
    conv(a) = eltype(a) == T ? a : convert(Array{T}, a)   # T == Float64 here unless every array is Float32: widening only
    hsml, m, rho, bq, w = conv(HSML), conv(M), conv(Rho), conv(Bin_Q), conv(Weights)
    N = length(m)
    n_images = ndims(bq) == 1 ? 1 : size(bq, 1)
    n = par_centred.Npixels[1]
    out = dimensions == 2 ? Array{Float64,3}(undef, n, n, n_images) : Array{Float64,3}(undef, n, n, n)
    shift = Float64.(par.center); half = Float64.(par_centred.halfsize)
    GC.@preserve pos_in hsml m rho bq w out shift half perm rot begin
        check(ccall((:s2g_sphmap_projected, LIB), Cint,
                    (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                     Int64, Int32, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ptr{Float64},
                     Float64, Int64, Int32, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}),
                    ctx.handle, dimensions, pos_in, hsml, m, rho, bq, w, N, n_images,
                    T == Float32 ? S2G_F32 : S2G_F64, perm, rot, shift, par.periodic, Float64(par.boxsize), half,
                    Float64(par_centred.len2pix), n, kernel_id(kernel), calc_mean, reduce_image, false, C_NULL, out,
                    C_NULL))
    end
    return out
end

# ------------------------------------------------------------------------------------------------------------------
# parallel=true: several GPUs, ONE Julia process (no Distributed workers, no NCCL.jl)
# ------------------------------------------------------------------------------------------------------------------
"""
    DeviceGroup(devices = 0:device_count()-1)

One library context per listed GPU, driven by host threads inside the library (s2g_group_init).  Replaces the
`addprocs` + `@spawnat` worker pool of the `parallel=true` branch (src/cic_interpolation/cic_interpolation.jl:171-215,
236-271): `sphmap_parallel` shards the particles with `domain_decomposition` (src/parallel/domain_decomp.jl:7-17),
every GPU deposits its slice, and `sum(fetch.(futures))` + `reduce_image` happen on the devices over NVLink peer memory.
"""
mutable struct DeviceGroup
    handle::Ptr{Cvoid}
    devices::Vector{Int32}
    function DeviceGroup(devices=0:(ccall((:s2g_device_count, LIB), Cint, ()) - 1))
        devs = Int32.(collect(devices))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:s2g_group_init, LIB), Cint, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), devs, length(devs), h))
        grp = new(h[], devs)
        finalizer(g -> ccall((:s2g_group_shutdown, LIB), Cint, (Ptr{Cvoid},), g.handle), grp)
        return grp
    end
end

Base.length(g::DeviceGroup) = length(g.devices)

const _default_group = Ref{Union{Nothing,DeviceGroup}}(nothing)
default_group() = something(_default_group[], (_default_group[] = DeviceGroup()))

"""
    sphmap_parallel(Pos, HSML, M, Rho, Bin_Q, Weights; param, par_centred, kernel, dimensions, calc_mean,
                    reduce_image, return_both_maps, group)

Body of `sphMapping(...; parallel=true)` (cic_interpolation.jl:171-215 for 2D, :236-271 for 3D) on all GPUs of `group`
(s2g_group_sphmap).  Same arguments and result as `sphmap_fused`; `Pos` is recentred in place.
"""
function sphmap_parallel(Pos::Matrix{T}, HSML, M, Rho, Bin_Q, Weights; param, par_centred, kernel, dimensions::Int=2,
                         calc_mean::Bool=false, reduce_image::Bool=true, return_both_maps::Bool=false,
                         group::DeviceGroup=default_group(),
                         ctx::Context=default_context()) where {T<:Union{Float32,Float64}}
    U, pos, hsml, m, rho, bq, w, shift, periodic, boxsize, writeback =
        _uniform_fused!(Pos, HSML, M, Rho, Bin_Q, Weights, param, ctx)
    N = length(m)
    n_images = ndims(bq) == 1 ? 1 : size(bq, 1)
    n = par_centred.Npixels[1]
    out = dimensions == 2 ? (return_both_maps ? Matrix{Float64}(undef, n * n, n_images + 1) :
                                                Array{Float64,3}(undef, n, n, n_images)) :
                            Array{Float64,3}(undef, n, n, n)
    half = Float64.(par_centred.halfsize)
    pos_out = writeback ? similar(pos) : pos
    GC.@preserve pos hsml m rho bq w out shift half pos_out begin
        check(ccall((:s2g_group_sphmap, LIB), Cint,
                    (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                     Int64, Int32, Int32, Ptr{Float64}, Int32, Float64, Ptr{Float64}, Float64, Int64, Int32, Int32,
                     Int32, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}),
                    group.handle, dimensions, pos, hsml, m, rho, bq, w, N, n_images,
                    U == Float32 ? S2G_F32 : S2G_F64, shift, periodic, boxsize, half,
                    Float64(par_centred.len2pix), n, kernel_id(kernel), calc_mean, reduce_image, return_both_maps,
                    writeback ? pointer(pos_out) : C_NULL, out, C_NULL))
    end
    writeback && (Pos .= pos_out)
    return out
end

"""
    healpix_map_parallel!(allsky_map, weight_map, Pos, Hsml, M, Rho, Bin_q, Weights; center, radius_limits, Nside,
                          kernel, calc_mean, group)

`healpix_map_fused!` over all GPUs of `group` (s2g_group_healpix_map); the `sorted[sel]` selection of
filter_sort_particles (src/healpix_interpolation/filter_particles.jl:33-41) is made over ALL particles, so the maps
are those of the single-device call.
"""
function healpix_map_parallel!(allsky_map, weight_map, Pos::Matrix{Float64}, Hsml::Vector{Float64},
                               M::Vector{Float64}, Rho::Vector{Float64}, Bin_q::Vector{Float64},
                               Weights::Vector{Float64}; center::Vector{<:Real}, radius_limits::Vector{<:Real},
                               Nside::Integer, kernel::AbstractSPHKernel, calc_mean::Bool=true,
                               group::DeviceGroup=default_group())
    a, w = allsky_map.pixels, weight_map.pixels
    cen = Float64.(center); rl = Float64.(radius_limits)
    pos_out = similar(Pos)
    GC.@preserve Pos Hsml M Rho Bin_q Weights a w cen rl pos_out begin
        check(ccall((:s2g_group_healpix_map, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64,
                     Ptr{Float64}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
                    group.handle, Pos, Hsml, M, Rho, Bin_q, Weights, length(Hsml), cen, rl, Nside, kernel_id(kernel),
                    calc_mean, pos_out, a, w, C_NULL))
    end
    Pos .= pos_out
    return allsky_map, weight_map
end

# ---------------------------------------------------------------------------------------------------------------
# context knobs, housekeeping
# ---------------------------------------------------------------------------------------------------------------
version() = unsafe_string(ccall((:s2g_version, LIB), Cstring, ()))
device_count() = Int(ccall((:s2g_device_count, LIB), Cint, ()))
sync(ctx::Context=default_context()) = check(ccall((:s2g_sync, LIB), Cint, (Ptr{Cvoid},), ctx.handle))

"strategy: :auto (per-particle choice by footprint size), :scatter (warp per particle, red.add), :gather (tile-owning CTAs)"
set_strategy!(s::Symbol; ctx::Context=default_context()) =
    check(ccall((:s2g_set_strategy, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.handle,
                Dict(:auto => 0, :scatter => 1, :gather => 2)[s]))

"force the numerical pass-A sum for every particle (default: closed form for unclipped, well resolved footprints)"
set_exact_norm!(on::Bool; ctx::Context=default_context()) =
    check(ccall((:s2g_set_exact_norm, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.handle, on))

"accumulate mode: :f64 (default, 1e-10 against the CPU path) or :f32 (optional, 1e-5; 2D tile-gather kernel only)"
set_accumulate_mode!(m::Symbol; ctx::Context=default_context()) =
    check(ccall((:s2g_set_accumulate_mode, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.handle, Dict(:f64 => 0, :f32 => 1)[m]))

"""
    domain_decomposition(N, N_workers)

src/parallel/domain_decomp.jl:7-17 through the library (needs no device): vector of 1-based `UnitRange`s.
"""
function domain_decomposition(N::Integer, N_workers::Integer)
    starts = Vector{Int64}(undef, N_workers); counts = Vector{Int64}(undef, N_workers)
    check(ccall((:s2g_domain_decomposition, LIB), Cint, (Int64, Int32, Ptr{Int64}, Ptr{Int64}), N, N_workers, starts,
                counts))
    return [(starts[i] + 1):(starts[i] + counts[i]) for i in 1:N_workers]
end

"""
    center_and_filter!(Pos, par, par_centred)

`center_particles` (src/cic_interpolation/filter_shift.jl:6-32: in place, in the precision of `Pos`, periodic wrap by
`boxsize/2`) followed by `filter_particles_in_image` (:40-58) on the device; returns the `BitVector`-like mask.
"""
function center_and_filter!(Pos::Matrix{T}, par, par_centred; ctx::Context=default_context()) where {T<:Union{Float32,Float64}}
    N = size(Pos, 2)
    mask = Vector{UInt8}(undef, N)
    pos_out = similar(Pos)
    shift = Float64.(par.center); zero3 = zeros(3); hs = Float64.(par_centred.halfsize)
    GC.@preserve Pos pos_out mask shift zero3 hs begin
        check(ccall((:s2g_center_filter, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Int32, Float64, Ptr{Float64}, Ptr{Float64},
                     Ptr{Cvoid}, Ptr{UInt8}),
                    ctx.handle, Pos, N, T == Float32 ? S2G_F32 : S2G_F64, shift, par.periodic, Float64(par.boxsize),
                    zero3, hs, pos_out, mask))
    end
    Pos .= pos_out
    return mask .!= 0x00
end

"""
    cic_deposit(Pos, Bin_Q; param, dimensions=3, average=true, periodic=false)
    tsc_deposit(Pos, Bin_Q; param, dimensions=3, average=true, periodic=false)

Cloud-in-cell / triangular-shaped-cloud grid assignment (the reference's src/tsc_interpolation/tsc_interpolation.jl is
commented out; semantics in DESIGN.md §6).  `Pos` relative to the image centre.  Returns the `(field, weight)` planes
as `Matrix{Float64}(N^dims, 2)`, or the averaged grid when `average`.
"""
function _stencil(order::Integer, Pos::Matrix{T}, Bin_Q::Vector{T}; param, dimensions::Int=3, average::Bool=true,
                  periodic::Bool=false, ctx::Context=default_context()) where {T<:Union{Float32,Float64}}
    n = param.Npixels[1]
    image = Matrix{Float64}(undef, n^dimensions, 2)
    GC.@preserve Pos Bin_Q image begin
        check(ccall((:s2g_stencil_deposit, LIB), Cint,
                    (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Float64, Int64, Int32,
                     Ptr{Float64}, Ptr{Cvoid}),
                    ctx.handle, order, dimensions, Pos, Bin_Q, length(Bin_Q), T == Float32 ? S2G_F32 : S2G_F64,
                    Float64(param.len2pix), n, periodic, image, C_NULL))
    end
    average || return image
    field = @views image[:, 1]; wgt = @views image[:, 2]
    out = [wgt[i] > 0 ? field[i] / wgt[i] : field[i] for i in eachindex(field)]
    return reshape(out, ntuple(_ -> n, dimensions))
end
cic_deposit(Pos, Bin_Q; kw...) = _stencil(2, Pos, Bin_Q; kw...)
tsc_deposit(Pos, Bin_Q; kw...) = _stencil(3, Pos, Bin_Q; kw...)

# The *_dev entry points (s2g_deposit_2d_dev, s2g_sphmap_dev, s2g_healpix_map_dev, s2g_reduce_image_*_dev,
# s2g_stencil_deposit_dev, s2g_accumulate_finite_dev) take DEVICE pointers: they are for callers that already hold
# their particles in GPU memory (CUDA.jl `CuArray`s: pass `pointer(x)`), with s2g_set_stream to order them after the
# caller's own kernels.  They mirror the host-buffer calls above argument for argument (include/sphtogrid_cuda.h).

end # module
