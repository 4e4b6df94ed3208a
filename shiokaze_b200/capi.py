"""ctypes binding of include/shkz_b200.h (the C-ABI of libshkz_b200.so). Fails loudly when the library
is missing or cannot be loaded; nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libshkz_b200.so")
HOOKS_LIB_PATH = os.path.join(HERE, "_build", "libshkz_b200_testhooks.so")  # same sources + -DSHKZ_B200_TEST_HOOKS (tests only)
ABI_VERSION = 4

OK, ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_COMM, ERR_STATE = range(6)
PRECOND_NONE, PRECOND_MG = 0, 1
PREC_FP64, PREC_MIXED, PREC_FP32 = 0, 1, 2
REAL_F32, REAL_F64 = 0, 1
IPC_BYTES = 128

EXPORTS = [
    "shkz_b200_abi_version", "shkz_b200_last_error", "shkz_b200_default_params", "shkz_b200_device_count",
    "shkz_b200_create", "shkz_b200_create_slab", "shkz_b200_destroy", "shkz_b200_project_host",
    "shkz_b200_project_device", "shkz_b200_prepare", "shkz_b200_extrapolate_constrain_device", "shkz_b200_extrapolate_constrain_host", "shkz_b200_resolve", "shkz_b200_host_alloc", "shkz_b200_host_free", "shkz_b200_slab_export",
    "shkz_b200_slab_connect", "shkz_b200_slab_connect_local", "shkz_b200_debug_fetch", "shkz_b200_profile_enable", "shkz_b200_profile_count",
    "shkz_b200_profile_get", "shkz_b200_debug_vcycle",
    "shkz_b200_csr_last_error", "shkz_b200_csr_default_params", "shkz_b200_csr_create", "shkz_b200_csr_destroy", "shkz_b200_csr_solve_host",
    "shkz_b200_advect_last_error", "shkz_b200_advect_default_params", "shkz_b200_advect_create", "shkz_b200_advect_destroy",
    "shkz_b200_advect_vector_host", "shkz_b200_advect_vector_device", "shkz_b200_advect_scalar_host", "shkz_b200_advect_scalar_device",
]
CSR_PRECOND_NONE, CSR_PRECOND_JACOBI = 0, 1


class Params(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("second_order_fluid", C.c_int32), ("second_order_solid", C.c_int32),
                ("apply_rhs_correct", C.c_int32), ("eps_fluid", C.c_double), ("eps_solid", C.c_double),
                ("surface_tension", C.c_double), ("rhs_correct", C.c_double), ("residual", C.c_double),
                ("max_iterations", C.c_uint32), ("precond", C.c_int32), ("precision", C.c_int32),
                ("mg_pre_sweeps", C.c_int32), ("mg_post_sweeps", C.c_int32), ("mg_coarse_sweeps", C.c_int32),
                ("mg_min_size", C.c_int32), ("check_every", C.c_int32), ("mg_coarse_scale", C.c_double),
                ("mg_gamma", C.c_int32), ("warm_start", C.c_int32), ("mg_omega", C.c_double),
                ("extrapolate_width", C.c_int32), ("velocity_masked", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("n_rows_global", C.c_uint64), ("iterations", C.c_uint32), ("converged", C.c_int32),
                ("reresid", C.c_double), ("rhs_absmax", C.c_double), ("has_dirichlet", C.c_int32), ("mg_levels", C.c_int32),
                ("kernel_launches", C.c_uint64), ("ms_h2d", C.c_float), ("ms_assemble", C.c_float), ("ms_setup", C.c_float),
                ("ms_solve", C.c_float), ("ms_update", C.c_float), ("ms_d2h", C.c_float), ("ms_total", C.c_float),
                ("ms_surftension", C.c_float), ("active_tiles", C.c_uint32), ("total_tiles", C.c_uint32),
                ("mg_mid_level", C.c_int32), ("mg_tail_level", C.c_int32), ("tile_depth", C.c_uint32), ("host_copies", C.c_uint32),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CsrParams(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("max_iterations", C.c_uint32), ("residual", C.c_double), ("precond", C.c_int32),
                ("check_every", C.c_int32)]


class CsrStats(C.Structure):
    _fields_ = [("iterations", C.c_uint32), ("converged", C.c_int32), ("reresid", C.c_double), ("rhs_absmax", C.c_double),
                ("ell_width", C.c_int32), ("reserved", C.c_int32), ("kernel_launches", C.c_uint64), ("ms_h2d", C.c_float),
                ("ms_solve", C.c_float), ("ms_d2h", C.c_float), ("reserved2", C.c_float)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


class AdvectParams(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("maccormack", C.c_int32), ("weno", C.c_int32), ("trim_narrowband", C.c_uint32),
                ("scalar_background", C.c_double)]


class AdvectStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("ms_h2d", C.c_float),
                ("ms_advect", C.c_float), ("ms_d2h", C.c_float), ("host_copies", C.c_uint32)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


class ShkzError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libshkz_b200 error {code}: {message}")
        self.code = code


_libs = {}


def lib(test_hooks: bool = False):
    """Load libshkz_b200.so (built in-tree by `make -C shiokaze_b200/csrc` / __graft_entry__.build()).
    test_hooks=True: the -DSHKZ_B200_TEST_HOOKS build of the same sources (adds shkz_b200_debug_vcycle + the validation kernels); tests only."""
    if test_hooks in _libs:
        return _libs[test_hooks]
    path = HOOKS_LIB_PATH if test_hooks else LIB_PATH
    if not os.path.isfile(path):
        raise ImportError(f"{path} is missing: build it with `make -C shiokaze_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u8p = C.c_void_p, C.POINTER(C.c_uint8)
    L.shkz_b200_abi_version.restype = C.c_int
    L.shkz_b200_last_error.restype = C.c_char_p
    L.shkz_b200_default_params.argtypes = [C.POINTER(Params)]
    L.shkz_b200_default_params.restype = None
    L.shkz_b200_device_count.restype = C.c_int
    L.shkz_b200_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(vp)]
    L.shkz_b200_create_slab.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(vp)]
    L.shkz_b200_destroy.argtypes = [vp]
    L.shkz_b200_destroy.restype = None
    proj = [vp, C.c_double, C.POINTER(vp), C.POINTER(vp), vp, vp, C.c_int, C.POINTER(Params), vp, vp, C.POINTER(Stats)]
    L.shkz_b200_project_host.argtypes = proj
    L.shkz_b200_project_device.argtypes = proj + [vp]
    L.shkz_b200_prepare.argtypes = [vp, C.POINTER(Params), C.c_int, C.c_int]
    L.shkz_b200_extrapolate_constrain_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, C.c_int, vp]
    L.shkz_b200_extrapolate_constrain_host.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, C.c_int]
    L.shkz_b200_resolve.argtypes = [vp, C.POINTER(Params), C.POINTER(Stats), vp]
    L.shkz_b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.shkz_b200_host_free.argtypes = [vp]
    L.shkz_b200_host_free.restype = None
    L.shkz_b200_slab_export.argtypes = [vp, u8p]
    L.shkz_b200_slab_connect.argtypes = [vp, C.c_int, C.c_int, u8p]
    L.shkz_b200_slab_connect_local.argtypes = [C.POINTER(vp), C.c_int]
    L.shkz_b200_debug_fetch.argtypes = [vp, C.c_char_p, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.shkz_b200_profile_enable.argtypes = [vp, C.c_int]
    L.shkz_b200_profile_count.argtypes = [vp]
    L.shkz_b200_profile_get.argtypes = [vp, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    L.shkz_b200_debug_vcycle.argtypes = [vp, C.POINTER(Params), C.c_int]
    L.shkz_b200_csr_last_error.restype = C.c_char_p
    L.shkz_b200_csr_default_params.argtypes = [C.POINTER(CsrParams)]
    L.shkz_b200_csr_default_params.restype = None
    L.shkz_b200_csr_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.shkz_b200_csr_destroy.argtypes = [vp]
    L.shkz_b200_csr_destroy.restype = None
    L.shkz_b200_csr_solve_host.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, vp, C.POINTER(CsrParams), C.POINTER(CsrStats)]
    L.shkz_b200_advect_last_error.restype = C.c_char_p
    L.shkz_b200_advect_default_params.argtypes = [C.POINTER(AdvectParams)]
    L.shkz_b200_advect_default_params.restype = None
    L.shkz_b200_advect_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(vp)]
    L.shkz_b200_advect_destroy.argtypes = [vp]
    L.shkz_b200_advect_destroy.restype = None
    adv_v = [vp, C.c_double, C.POINTER(vp), C.POINTER(vp), vp, C.POINTER(AdvectParams), C.POINTER(AdvectStats)]
    adv_s = [vp, C.c_double, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, C.POINTER(AdvectParams), C.POINTER(AdvectStats)]
    L.shkz_b200_advect_vector_host.argtypes = adv_v
    L.shkz_b200_advect_vector_device.argtypes = adv_v + [vp]
    L.shkz_b200_advect_scalar_host.argtypes = adv_s
    L.shkz_b200_advect_scalar_device.argtypes = adv_s + [vp]
    _libs[test_hooks] = L
    return L


def check(code, L=None):
    """Raise on a non-zero status; L: the library the failing call went through (default: the product library)."""
    if code != OK:
        raise ShkzError(code, (L or lib()).shkz_b200_last_error().decode(errors="replace"))


def check_csr(code):
    if code != OK:
        raise ShkzError(code, lib().shkz_b200_csr_last_error().decode(errors="replace"))


def check_advect(code):
    if code != OK:
        raise ShkzError(code, lib().shkz_b200_advect_last_error().decode(errors="replace"))


def default_params() -> Params:
    p = Params()
    lib().shkz_b200_default_params(C.byref(p))
    return p
