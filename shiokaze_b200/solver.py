"""Python host mirror of the reference module `macpressuresolver3` (src/projection/macpressuresolver3.cpp)
on top of the C-ABI. Flag names and defaults are the reference's (macpressuresolver3.cpp:274-280,296-305;
macutility3.cpp:408-421; pcg.cpp:39-44,75-80); the `set_target_volume` / volume-correction PI state lives
here on the host exactly as it does in the reference module (:204-217,316-320).

Device memory is PyTorch's (plumbing only); all arithmetic happens in libshkz_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import capi

_PRECOND = {"none": capi.PRECOND_NONE, "cg": capi.PRECOND_NONE, "mg": capi.PRECOND_MG}
_PRECISION = {"fp64": capi.PREC_FP64, "mixed": capi.PREC_MIXED, "fp32": capi.PREC_FP32}


@dataclass
class ProjectionResult:
    iterations: int
    reresid: float
    converged: bool
    n_rows: int
    stats: dict


class MacPressureSolver3:
    """project(dt, velocity, solid, fluid, surface_tension) with the reference's configurable parameters.

    Reference flags: SecondOrderAccurateFluid, SecondOrderAccurateSolid, Gain, WarmStart, EpsFluid, EpsSolid,
    Residual, MaxIterations. Additive flags: Precond ("mg"|"none"),
    Precision ("mixed"|"fp64"|"fp32"), MGPreSweeps, MGPostSweeps, MGCoarseSweeps, MGMinSize, CheckEvery.
    """

    def __init__(self, shape: Sequence[int], dx: float, real: str = "f32", device: int = 0, zrange=None, test_hooks: bool = False, **flags):
        self._L = capi.lib(test_hooks)   # test_hooks: the -DSHKZ_B200_TEST_HOOKS build (debug_vcycle); never used by the product path
        self.nx, self.ny, self.nz = (int(v) for v in shape)
        self.dx = float(dx)
        self.real = real
        self.np_real = np.float32 if real == "f32" else np.float64
        self.device = device
        self.zrange = (0, self.nz) if zrange is None else (int(zrange[0]), int(zrange[1]))
        self.nzl = self.zrange[1] - self.zrange[0]
        self.params = capi.default_params()
        self.gain = 1.0
        self._target_volume = self._current_volume = self._y_prev = 0.0
        self.configure(**flags)
        h = C.c_void_p()
        self._ck(self._L.shkz_b200_create_slab(self.nx, self.ny, self.nz, self.zrange[0], self.zrange[1], self.dx,
                                                     capi.REAL_F32 if real == "f32" else capi.REAL_F64, device, C.byref(h)))
        self._h = h
        self.last_rhs_correct = 0.0

    def _ck(self, code):
        capi.check(code, self._L)

    # -- reference interface ----------------------------------------------------------------------
    def configure(self, **flags):
        p = self.params
        for key, value in flags.items():
            if key == "SecondOrderAccurateFluid": p.second_order_fluid = int(bool(value))
            elif key == "SecondOrderAccurateSolid": p.second_order_solid = int(bool(value))
            elif key == "Gain": self.gain = float(value)
            elif key == "WarmStart": p.warm_start = int(bool(value))
            elif key == "EpsFluid": p.eps_fluid = float(value)
            elif key == "EpsSolid": p.eps_solid = float(value)
            elif key == "Residual": p.residual = float(value)
            elif key == "MaxIterations": p.max_iterations = int(value)
            elif key == "Precond": p.precond = _PRECOND[str(value).lower()]
            elif key == "Precision": p.precision = _PRECISION[str(value).lower()]
            elif key == "MGPreSweeps": p.mg_pre_sweeps = int(value)
            elif key == "MGPostSweeps": p.mg_post_sweeps = int(value)
            elif key == "MGCoarseSweeps": p.mg_coarse_sweeps = int(value)
            elif key == "MGMinSize": p.mg_min_size = int(value)
            elif key == "MGCoarseScale": p.mg_coarse_scale = float(value)
            elif key == "CheckEvery": p.check_every = int(value)
            elif key == "MGGamma": p.mg_gamma = int(value)
            elif key == "MGOmega": p.mg_omega = float(value)
            elif key == "ExtrapolateWidth": p.extrapolate_width = int(value)
            elif key == "VelocityMasked": p.velocity_masked = int(bool(value))   # (library-level: the entries of inactive faces are unspecified, include/shkz_b200.h)
            else:
                raise KeyError(f"unknown flag {key}")

    def set_target_volume(self, current_volume: float, target_volume: float):
        self._current_volume, self._target_volume = float(current_volume), float(target_volume)

    def _volume_correction(self, dt: float):
        """PI controller of macpressuresolver3.cpp:204-214 (host-side scalar state)."""
        p = self.params
        p.apply_rhs_correct = 0
        p.rhs_correct = 0.0
        if self.gain and self._target_volume:
            x = (self._current_volume - self._target_volume) / self._target_volume
            y = self._y_prev + x * dt
            self._y_prev = y
            kp = self.gain * 2.3 / (25.0 * 0.01)
            ki = kp * kp / 16.0
            p.rhs_correct = -(kp * x + ki * y) / (x + 1.0)
            p.apply_rhs_correct = 1
        self.last_rhs_correct = p.rhs_correct

    # -- shapes -----------------------------------------------------------------------------------
    def face_shapes(self):
        nx, ny, nzl = self.nx, self.ny, self.nzl
        return [(nzl, ny, nx + 1), (nzl, ny + 1, nx), (nzl + 1, ny, nx)]

    def _finish(self, st: capi.Stats) -> ProjectionResult:
        return ProjectionResult(int(st.iterations), float(st.reresid), bool(st.converged), int(st.n_rows), st.asdict())

    # -- host buffers (numpy): the call the Shiokaze plugin makes ------------------------------------
    def project(self, dt, velocity, velocity_active, solid, fluid, fluid_levelset: bool, surface_tension: float = 0.0,
                pressure_out=None, pressure_active_out=None):
        """In place on numpy arrays (velocity: 3 face arrays, velocity_active: 3 uint8). Returns
        (pressure, pressure_active, ProjectionResult). pressure_out / pressure_active_out: optional caller-owned
        result buffers (page-locked ones make the device-to-host copy run at PCIe speed)."""
        rt = self.np_real
        for v, a, shp in zip(velocity, velocity_active, self.face_shapes()):
            assert v.dtype == rt and v.flags.c_contiguous and v.shape == shp, (v.dtype, v.shape, shp)
            assert a.dtype == np.uint8 and a.flags.c_contiguous and a.shape == shp
        assert fluid.dtype == rt and fluid.flags.c_contiguous and fluid.shape == (self.nzl, self.ny, self.nx)
        if solid is not None:
            assert solid.dtype == rt and solid.flags.c_contiguous and solid.shape == (self.nzl + 1, self.ny + 1, self.nx + 1)
        self._volume_correction(dt)
        self.params.surface_tension = float(surface_tension)
        shp = (self.nzl, self.ny, self.nx)
        pressure = np.zeros(shp, dtype=rt) if pressure_out is None else pressure_out
        pact = np.zeros(shp, dtype=np.uint8) if pressure_active_out is None else pressure_active_out
        assert pressure.dtype == rt and pressure.flags.c_contiguous and pressure.shape == shp
        assert pact.dtype == np.uint8 and pact.flags.c_contiguous and pact.shape == shp
        vp = (C.c_void_p * 3)(*[v.ctypes.data for v in velocity])
        ap = (C.c_void_p * 3)(*[a.ctypes.data for a in velocity_active])
        st = capi.Stats()
        self._ck(self._L.shkz_b200_project_host(self._h, float(dt), vp, ap, solid.ctypes.data if solid is not None else None,
                                                     fluid.ctypes.data, int(bool(fluid_levelset)), C.byref(self.params),
                                                     pressure.ctypes.data, pact.ctypes.data, C.byref(st)))
        return pressure, pact, self._finish(st)

    def prepare(self, host_buffers: bool = True, have_solid: bool = True):
        """shkz_b200_prepare: allocate now what the next project() with the current flags would allocate on first use."""
        self._ck(self._L.shkz_b200_prepare(self._h, C.byref(self.params), int(host_buffers), int(have_solid)))

    def project_scene(self, scene, surface_tension: Optional[float] = None):
        """Convenience for tests: run a scenes.Scene through project() on copies; returns dict of outputs."""
        rt = self.np_real
        vel = [np.ascontiguousarray(v, dtype=rt).copy() for v in scene.vel]
        act = [np.ascontiguousarray(a, dtype=np.uint8).copy() for a in scene.vel_active]
        solid = np.ascontiguousarray(scene.solid, dtype=rt) if scene.solid is not None else None
        fluid = np.ascontiguousarray(scene.fluid, dtype=rt)
        p, pa, res = self.project(scene.dt, vel, act, solid, fluid, scene.fluid_levelset,
                                  scene.surface_tension if surface_tension is None else surface_tension)
        return dict(vel=vel, vel_active=act, pressure=p, pressure_active=pa, result=res)

    # -- device buffers (torch tensors on this solver's GPU) ----------------------------------------
    def project_device(self, dt, velocity, velocity_active, solid, fluid, fluid_levelset: bool, pressure=None,
                       pressure_active=None, surface_tension: float = 0.0, stream: int = 0):
        self._volume_correction(dt)
        self.params.surface_tension = float(surface_tension)
        vp = (C.c_void_p * 3)(*[v.data_ptr() for v in velocity])
        ap = (C.c_void_p * 3)(*[a.data_ptr() for a in velocity_active])
        st = capi.Stats()
        self._ck(self._L.shkz_b200_project_device(self._h, float(dt), vp, ap, solid.data_ptr() if solid is not None else None,
                                                       fluid.data_ptr(), int(bool(fluid_levelset)), C.byref(self.params),
                                                       pressure.data_ptr() if pressure is not None else None,
                                                       pressure_active.data_ptr() if pressure_active is not None else None,
                                                       C.byref(st), stream or None))
        return self._finish(st)

    def extrapolate_and_constrain_velocity(self, solid, velocity, velocity_active, width: int):
        """macutility3::extrapolate_and_constrain_velocity (src/utility/macutility3.cpp:89-93) in place on numpy arrays; solid: nodal level set
        or None (the host's levelset_exist(solid) is false)."""
        rt = self.np_real
        for v, a, shp in zip(velocity, velocity_active, self.face_shapes()):
            assert v.dtype == rt and v.flags.c_contiguous and v.shape == shp and a.dtype == np.uint8 and a.flags.c_contiguous and a.shape == shp
        if solid is not None:
            assert solid.dtype == rt and solid.flags.c_contiguous and solid.shape == (self.nzl + 1, self.ny + 1, self.nx + 1)
        vp = (C.c_void_p * 3)(*[v.ctypes.data for v in velocity])
        ap = (C.c_void_p * 3)(*[a.ctypes.data for a in velocity_active])
        self._ck(self._L.shkz_b200_extrapolate_constrain_host(self._h, vp, ap, solid.ctypes.data if solid is not None else None, int(width)))

    def resolve(self, stream: int = 0) -> ProjectionResult:
        """Repeat only the linear solve of the last project() (same matrix and right-hand side)."""
        st = capi.Stats()
        self._ck(self._L.shkz_b200_resolve(self._h, C.byref(self.params), C.byref(st), stream or None))
        return self._finish(st)

    # -- per-kernel timing ---------------------------------------------------------------------------
    def profile(self, on: bool = True):
        self._ck(self._L.shkz_b200_profile_enable(self._h, int(on)))

    def profile_table(self) -> dict:
        """name -> (launches, total ms) accumulated since profile(True)."""
        out = {}
        L = self._L
        for i in range(L.shkz_b200_profile_count(self._h)):
            name = C.create_string_buffer(64)
            n, ms = C.c_uint64(), C.c_double()
            capi.check(L.shkz_b200_profile_get(self._h, i, name, 64, C.byref(n), C.byref(ms)))
            out[name.value.decode()] = (int(n.value), float(ms.value))
        return out

    # -- test hook ----------------------------------------------------------------------------------
    def debug_fetch(self, name: str) -> np.ndarray:
        need = C.c_size_t()
        self._ck(self._L.shkz_b200_debug_fetch(self._h, name.encode(), None, 0, C.byref(need)))
        buf = np.empty(need.value, dtype=np.uint8)
        self._ck(self._L.shkz_b200_debug_fetch(self._h, name.encode(), buf.ctypes.data, buf.nbytes, None))
        return buf

    def debug_vcycle(self, legacy=0) -> np.ndarray:
        """One V-cycle applied to the last right-hand side: 0 product kernels, 1 unfused validation kernels,
        2 product path with the scalar sweep kernel forced."""
        self._ck(self._L.shkz_b200_debug_vcycle(self._h, C.byref(self.params), int(legacy)))
        return self.debug_fetch("vcycle").view(np.float32).reshape(self.nzl, self.ny, self.nx).copy()

    def close(self):
        if getattr(self, "_h", None):
            self._L.shkz_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
