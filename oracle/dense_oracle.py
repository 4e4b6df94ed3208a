"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/dense_oracle.c (the dense CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libdense_oracle.so")


class Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("second_order_fluid", C.c_int32), ("second_order_solid", C.c_int32),
                ("real_is_double", C.c_int32), ("fluid_levelset_exist", C.c_int32), ("have_solid", C.c_int32),
                ("max_iterations", C.c_uint32), ("pad", C.c_int32),
                ("dx", C.c_double), ("dt", C.c_double), ("eps_fluid", C.c_double), ("eps_solid", C.c_double),
                ("surface_tension", C.c_double), ("rhs_correct", C.c_double), ("residual", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("nnz", C.c_uint64), ("iterations", C.c_uint32), ("converged", C.c_int32),
                ("reresid", C.c_double), ("rhs_absmax", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "dense_oracle.c")
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_volume_correction.restype = C.c_double
        _lib.oracle_volume_correction.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
    return _lib


@dataclass
class OracleResult:
    vel: list
    vel_active: list
    pressure: np.ndarray
    in_rows: np.ndarray
    areas: list
    rhos: list
    rhs: np.ndarray
    diag: np.ndarray
    dirichlet: np.ndarray
    iterations: int
    reresid: float
    n_rows: int
    nnz: int
    converged: bool
    rhs_absmax: float


def _dptr3(arrs):
    return (C.POINTER(C.c_double) * 3)(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs])


def _bptr3(arrs):
    return (C.POINTER(C.c_uint8) * 3)(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in arrs])


def project(scene, residual=1e-4, max_iterations=30000, eps_fluid=1e-2, eps_solid=1e-2,
            second_order_fluid=True, second_order_solid=True, real_is_double=False,
            rhs_correct=0.0, surface_tension=None, warm_state=None) -> OracleResult:
    """Dense restatement of macpressuresolver3::project() on a shiokaze_b200.scenes.Scene (whole grid).
    warm_state: WarmStart=Yes — a dict kept by the caller between calls (the module's m_prev_pressure, by row number)."""
    nx, ny, nz = scene.nx, scene.ny, scene.nz
    assert scene.zrange == (0, nz)
    P = Params(nx, ny, nz, int(second_order_fluid), int(second_order_solid), int(real_is_double),
               int(scene.fluid_levelset), int(scene.solid is not None), max_iterations, 0,
               scene.dx, scene.dt, eps_fluid, eps_solid,
               scene.surface_tension if surface_tension is None else surface_tension, rhs_correct, residual)
    vel = [np.ascontiguousarray(v, dtype=np.float64).copy() for v in scene.vel]
    act = [np.ascontiguousarray(a, dtype=np.uint8).copy() for a in scene.vel_active]
    fluid = np.ascontiguousarray(scene.fluid, dtype=np.float64)
    solid = np.ascontiguousarray(scene.solid, dtype=np.float64) if scene.solid is not None else None
    pressure = np.zeros((nz, ny, nx), dtype=np.float64)
    in_rows = np.zeros((nz, ny, nx), dtype=np.uint8)
    areas = [np.zeros(v.shape, dtype=np.float64) for v in vel]
    rhos = [np.zeros(v.shape, dtype=np.float64) for v in vel]
    rhs = np.zeros((nz, ny, nx), dtype=np.float64)
    diag = np.zeros((nz, ny, nx), dtype=np.float64)
    dirichlet = np.zeros((nz, ny, nx), dtype=np.float64)
    st = Stats()
    extra = []
    fn = lib().oracle_project
    if warm_state is not None:
        if "prev" not in warm_state:
            warm_state["prev"] = np.zeros(nx * ny * nz, dtype=np.float64)
            warm_state["rows"] = C.c_uint64(0)
        fn = lib().oracle_project_warm
        extra = [warm_state["prev"].ctypes.data_as(C.POINTER(C.c_double)), C.byref(warm_state["rows"])]
    rc = fn(C.byref(P), _dptr3(vel), _bptr3(act),
                              solid.ctypes.data_as(C.POINTER(C.c_double)) if solid is not None else None,
                              fluid.ctypes.data_as(C.POINTER(C.c_double)),
                              pressure.ctypes.data_as(C.POINTER(C.c_double)), in_rows.ctypes.data_as(C.POINTER(C.c_uint8)),
                              _dptr3(areas), _dptr3(rhos), rhs.ctypes.data_as(C.POINTER(C.c_double)),
                              diag.ctypes.data_as(C.POINTER(C.c_double)),
                              dirichlet.ctypes.data_as(C.POINTER(C.c_double)), C.byref(st), *extra)
    assert rc == 0
    return OracleResult(vel, act, pressure, in_rows, areas, rhos, rhs, diag, dirichlet, int(st.iterations), float(st.reresid),
                        int(st.n_rows), int(st.nnz), bool(st.converged), float(st.rhs_absmax))


def volume_correction(gain, current_volume, target_volume, dt, y_prev):
    y = C.c_double(y_prev)
    v = lib().oracle_volume_correction(gain, current_volume, target_volume, dt, C.byref(y))
    return float(v), float(y.value)
