#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): sphMapping(parallel=True) — particles sharded into slices of equal summed
footprint, partial flat images reduce-scattered over NCCL, reduce_image division on every rank's pixel slice, gather,
transposition — must equal the single-GPU map and the CPU oracle.  Also: Float32 positions with Float64 fields,
several quantities at once, `return_both_maps`."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from util import assert_parity, random_particles

s2g = ge.load_package()
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = s2g.Context(lr)
pos, hsml, m, rho, q, w = random_particles(99, 20001, box=6.5, hmax=0.5, center=3.0)
kw = dict(center=[3.0, 3.0, 3.0], x_size=5.4, y_size=5.4, z_size=5.4, Npixels=256, boxsize=6.0)
par = s2g.mappingParameters(**kw)
for dims, kern in ((2, s2g.WendlandC6(2)), (3, s2g.Cubic(3))):
    if dims == 3:
        par = s2g.mappingParameters(center=[3.0, 3.0, 3.0], x_size=5.4, y_size=5.4, z_size=5.4, Npixels=48, boxsize=6.0)
    p1, p2 = pos.copy(), pos.copy()
    a = s2g.sphMapping(p1, hsml, m, rho, q, w, param=par, kernel=kern, calc_mean=True, dimensions=dims, parallel=True,
                       show_progress=False, ctx=ctx)
    b = s2g.sphMapping(p2, hsml, m, rho, q, w, param=par, kernel=kern, calc_mean=True, dimensions=dims,
                       parallel=False, show_progress=False, ctx=ctx)
    assert np.array_equal(p1, p2)
    e = assert_parity(a, b, rtol=1e-12, what=f"parallel vs serial dims={dims}")
    if rank == 0:
        from oracle import oracle as orc
        opar = orc.mapping_parameters(center=[3.0, 3.0, 3.0], x_size=5.4, y_size=5.4, z_size=5.4,
                                      Npixels=int(par.Npixels[0]), boxsize=6.0)
        ref = orc.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=opar, kernel=kern.name, calc_mean=True,
                              dimensions=dims)
        e2 = assert_parity(a, ref, what="parallel vs oracle")
        print(f"world={world} dims={dims}: parallel==serial (max rel {e:.1e}), vs oracle {e2:.1e}  OK", flush=True)
# Float32 positions + Float64 fields, two quantities, both reduce_image settings, return_both_maps
par = s2g.mappingParameters(**kw)
pos32 = pos.astype(np.float32)
Q = np.stack([q, np.sqrt(q + 1.0)], axis=1)
for kwargs in (dict(reduce_image=True), dict(reduce_image=False), dict(return_both_maps=True)):
    p1, p2 = pos32.copy(), pos32.copy()
    a = s2g.sphMapping(p1, hsml, m, rho, Q, w, param=par, kernel=s2g.WendlandC4(2), calc_mean=True, parallel=True,
                       show_progress=False, ctx=ctx, **kwargs)
    b = s2g.sphMapping(p2, hsml, m, rho, Q, w, param=par, kernel=s2g.WendlandC4(2), calc_mean=True, parallel=False,
                       show_progress=False, ctx=ctx, **kwargs)
    assert np.array_equal(p1, p2) and p1.dtype == np.float32 and a.shape == b.shape
    e = assert_parity(a, b, rtol=1e-12, what=f"mixed dtype, {kwargs}")
    if rank == 0:
        print(f"world={world} Float32 Pos + Float64 fields, 2 quantities, {kwargs}: parallel==serial (max rel {e:.1e})  OK",
              flush=True)
dist.barrier()
dist.destroy_process_group()
