"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/shkz_b200.h declares,
and refuses to compute without a CUDA device (no CPU fallback). No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from shiokaze_b200 import capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "shkz_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shkz_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(L, name), name
    assert sorted(capi.EXPORTS) == names
    assert L.shkz_b200_abi_version() == capi.ABI_VERSION


def test_validation_kernels_ship_only_in_the_test_hook_library():
    """kernels_legacy.cuh + shkz_b200_debug_vcycle are test code: the product library carries neither the kernels nor a working hook."""
    prod = open(capi.LIB_PATH, "rb").read()
    hooks = open(capi.HOOKS_LIB_PATH, "rb").read()
    assert b"k_legacy_rbgs" not in prod and b"k_legacy_residual" not in prod
    assert b"k_legacy_rbgs" in hooks
    assert capi.lib().shkz_b200_debug_vcycle(None, None, 0) == capi.ERR_STATE          # compiled out
    assert b"testhooks" in capi.lib().shkz_b200_last_error()
    H = capi.lib(test_hooks=True)
    assert H.shkz_b200_abi_version() == capi.ABI_VERSION
    assert H.shkz_b200_debug_vcycle(None, None, 0) == capi.ERR_ARG                     # present: complains about the NULL solver


def test_default_params_are_the_reference_defaults():
    p = capi.default_params()
    assert p.struct_size == C.sizeof(capi.Params)
    assert (p.second_order_fluid, p.second_order_solid) == (1, 1)          # macpressuresolver3.cpp:300-301
    assert (p.eps_fluid, p.eps_solid) == (1e-2, 1e-2)                      # macutility3.cpp:419-420
    assert p.residual == 1e-4 and p.max_iterations == 30000                # pcg.cpp:76-77
    assert p.precond == capi.PRECOND_MG and p.precision == capi.PREC_MIXED
    assert p.warm_start == 0                                               # macpressuresolver3.cpp:304


def test_argument_errors_do_not_need_a_device():
    L = capi.lib()
    h = C.c_void_p()
    assert L.shkz_b200_create(0, 8, 8, 0.1, capi.REAL_F32, 0, C.byref(h)) == capi.ERR_ARG
    assert L.shkz_b200_create(8, 8, 8, -1.0, capi.REAL_F32, 0, C.byref(h)) == capi.ERR_ARG
    assert L.shkz_b200_create_slab(8, 8, 8, 4, 2, 0.1, capi.REAL_F32, 0, C.byref(h)) == capi.ERR_ARG
    assert L.shkz_b200_create(8, 8, 8, 0.1, 7, 0, C.byref(h)) == capi.ERR_ARG
    assert b"real type" in L.shkz_b200_last_error()
    assert L.shkz_b200_resolve(None, None, None, None) == capi.ERR_ARG


def test_no_cpu_fallback():
    L = capi.lib()
    if L.shkz_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert L.shkz_b200_create(8, 8, 8, 0.125, capi.REAL_F32, 0, C.byref(h)) == capi.ERR_NO_DEVICE
    assert not h.value
    assert b"no CPU fallback" in L.shkz_b200_last_error()
    from shiokaze_b200 import MacPressureSolver3
    with pytest.raises(capi.ShkzError):
        MacPressureSolver3((8, 8, 8), 0.125)


def test_product_never_imports_the_oracle():
    """The shipped package must not reach into oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "shiokaze_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(base, f), errors="replace").read()
                assert "dense_oracle" not in src and "oracle." not in src.replace("oracle/", ""), os.path.join(base, f)


def test_csr_entry_points_refuse_without_a_device_and_check_arguments():
    L = capi.lib()
    p = capi.CsrParams()
    L.shkz_b200_csr_default_params(C.byref(p))
    assert p.struct_size == C.sizeof(capi.CsrParams)
    assert p.residual == 1e-4 and p.max_iterations == 30000 and p.precond == capi.CSR_PRECOND_NONE      # pcg.cpp:76-77
    assert L.shkz_b200_csr_create(0, None) == capi.ERR_ARG
    assert L.shkz_b200_csr_solve_host(None, 0, None, None, None, None, None, None, None) == capi.ERR_ARG
    if L.shkz_b200_device_count() == 0:
        h = C.c_void_p()
        assert L.shkz_b200_csr_create(0, C.byref(h)) == capi.ERR_NO_DEVICE and not h.value
        assert b"no CPU fallback" in L.shkz_b200_csr_last_error()
        from shiokaze_b200 import B200CG
        with pytest.raises(capi.ShkzError):
            B200CG()


def test_advect_entry_points_refuse_without_a_device_and_check_arguments():
    L = capi.lib()
    p = capi.AdvectParams()
    L.shkz_b200_advect_default_params(C.byref(p))
    assert p.struct_size == C.sizeof(capi.AdvectParams)
    assert p.maccormack == 1 and p.weno == 0 and p.trim_narrowband == 1 and p.scalar_background == 0.0      # macadvection3.cpp:285-289
    h = C.c_void_p()
    assert L.shkz_b200_advect_create(8, 8, 8, 0.125, capi.REAL_F32, 0, None) == capi.ERR_ARG
    assert L.shkz_b200_advect_create(8, 1, 8, 0.125, capi.REAL_F32, 0, C.byref(h)) == capi.ERR_ARG
    assert L.shkz_b200_advect_create(8, 8, 8, -1.0, capi.REAL_F32, 0, C.byref(h)) == capi.ERR_ARG
    assert L.shkz_b200_advect_vector_host(None, 0.1, None, None, None, None, None) == capi.ERR_ARG
    assert L.shkz_b200_advect_scalar_host(None, 0.1, None, None, None, None, None, None, None) == capi.ERR_ARG
    if L.shkz_b200_device_count() == 0:
        assert L.shkz_b200_advect_create(8, 8, 8, 0.125, capi.REAL_F32, 0, C.byref(h)) == capi.ERR_NO_DEVICE and not h.value
        assert b"no CPU fallback" in L.shkz_b200_advect_last_error()
        from shiokaze_b200 import MacAdvection3
        with pytest.raises(capi.ShkzError):
            MacAdvection3((8, 8, 8), 0.125)


def test_host_build_of_the_advection_source_is_test_infrastructure_only():
    """csrc/advect.cu carries a host-loop harness behind -DSHKZ_B200_ADVECT_HOSTCHECK (tests/test_advect_cpu.py builds it into oracle/_build): the product
    library must not export it."""
    L = capi.lib()
    for sym in ("shkz_b200_hostcheck_advect_vector", "shkz_b200_hostcheck_advect_scalar"):
        assert not hasattr(L, sym), sym


@pytest.mark.parametrize("flags,module", [({"Projection": "b200pressure3"}, "b200pressure3.so"), ({"LinSolver": "b200cg"}, "b200cg.so"),
                                          ({"Advection": "b200advection3", "RefAdvect": "vector"}, "b200advection3.so")])
def test_shiokaze_modules_load_under_the_reference_loader_and_refuse_without_a_device(flags, module):
    """The drop-in boundary without a GPU: the reference's own host (oracle/ref_driver) dlopens our module by name, casts it to the
    interface, configures it — and the module then stops the run with the library's no-device error instead of computing anything."""
    from oracle import refio
    from shiokaze_b200 import scenes
    if capi.lib().shkz_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    if not (refio.ref_available("f32") and os.path.isfile(os.path.join(refio.ref_dir("f32"), "libshiokaze_" + module))):
        pytest.skip("oracle/_ref (reference build + modules) is not built here")
    projection = flags.pop("Projection", None)
    with pytest.raises(RuntimeError) as e:
        refio.run_reference(scenes.dambreak(12), "f32", flags=flags, projection=projection, timeout=120)
    text = str(e.value)
    assert f'Loaded "{module}"' in text
    assert "no CUDA device available" in text and "no CPU fallback" in text
