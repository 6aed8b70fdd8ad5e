"""sphtogrid.jl_b200 — B200-native particle deposition behind SPHtoGrid.jl's API.

Host mirror (same names as the Julia package) of the ONE hot path this project accelerates; every compute call goes
through the C ABI of libsphtogrid_cuda.so (include/sphtogrid_cuda.h).  There is no CPU fallback.
"""
from ._lib import (Context, DeviceGroup, S2GError, build, default_context, lib, LIB_PATH,  # noqa: F401
                   EXPORTED_SYMBOLS)
from .kernels import (AbstractSPHKernel, Cubic, Quintic, WendlandC2, WendlandC4, WendlandC6,  # noqa: F401
                      WendlandC8)
from .parameters import mappingParameters, recentred_parameters  # noqa: F401
from .mapping import (sphMapping, map_it, cic_mapping_2D, cic_mapping_3D, reduce_image_2D,  # noqa: F401
                      reduce_image_3D, center_particles, filter_particles_in_image, domain_decomposition,
                      part_weight_one, part_weight_physical, part_weight_emission, part_weight_spectroscopic)
from .healpix import healpix_map, healpix_deposit, filter_sort_particles, find_in_shell  # noqa: F401
from .stencils import cic_deposit, tsc_deposit  # noqa: F401
from .rotate import (rotate_3D, rotate_3D_, rotate_to_xz_plane, rotate_to_yz_plane,  # noqa: F401
                     project_along_axis, euler_matrix)
from . import distributed, gadget, io  # noqa: F401
from .distributed import distributed_cic_map, distributed_allsky_map  # noqa: F401
from .io import (write_fits_image, read_fits_image, read_allsky_fits_image, save_healpix_fits,  # noqa: F401
                 read_healpix_fits, write_vtk_image, get_map_grid_3D)

__all__ = ["sphMapping", "map_it", "mappingParameters", "healpix_map", "Cubic", "Quintic", "WendlandC2", "WendlandC4",
           "WendlandC6", "WendlandC8", "cic_deposit", "tsc_deposit", "Context", "DeviceGroup"]
