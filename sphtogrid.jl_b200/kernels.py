"""SPHKernels.jl kernel types (names and constructor signature as in SPHKernels.jl v2: `Cubic()`, `WendlandC6(2)`, ...).

The deposit only needs the kernel *shape*: norm(dim) * h_inv^dim cancels between `kernel_norm` and `weight_per_pix`
(src/cic_interpolation/cic_2D.jl:187-199), which is also why the reference's tests can hand a 3-D `Cubic()` to a 2-D
map (test/runtests.jl:720,729).  `dim` is kept for API fidelity.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class AbstractSPHKernel:
    dim: int = 3
    kernel_id: int = -1
    name: str = ""


def _mk(name, kid):
    def ctor(*args):
        # SPHKernels: Kernel(dim::Integer=3) or Kernel(T::DataType=Float64, dim::Integer=3)
        dim = 3
        for a in args:
            if isinstance(a, int):
                dim = a
        if dim not in (1, 2, 3):
            raise ValueError(f"{name} not defined for {dim} dimensions!")
        return AbstractSPHKernel(dim=dim, kernel_id=kid, name=name)
    ctor.__name__ = name
    return ctor


Cubic = _mk("Cubic", 0)
Quintic = _mk("Quintic", 1)
WendlandC2 = _mk("WendlandC2", 2)
WendlandC4 = _mk("WendlandC4", 3)
WendlandC6 = _mk("WendlandC6", 4)
WendlandC8 = _mk("WendlandC8", 5)
