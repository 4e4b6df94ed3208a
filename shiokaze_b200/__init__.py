"""shiokaze_b200 — B200-native drop-in for Shiokaze's 3D pressure projection (macpressuresolver3 path).

The product is the CUDA library `_build/libshkz_b200.so` behind the C-ABI of include/shkz_b200.h and the
Shiokaze module in plugin/. This Python package is the thin host mirror used by tests and bench.py:
`capi` binds the C-ABI with ctypes, `solver.MacPressureSolver3` mirrors the reference module's
interface, `linsolver.B200CG` the `LinSolver` plug-in point for assembled systems, `advection.MacAdvection3` the `Advection` plug-in point, `scenes` generates the synthetic inputs, `dist` wires z-slabs over torch.distributed.
There is no CPU fallback anywhere in this package.
"""
from . import capi, dist, scenes  # noqa: F401
from .advection import MacAdvection3  # noqa: F401
from .linsolver import B200CG, LinResult  # noqa: F401
from .solver import MacPressureSolver3, ProjectionResult  # noqa: F401
