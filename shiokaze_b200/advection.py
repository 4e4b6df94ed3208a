"""Host mirror of the reference's `Advection` plug-in point (macadvection3_interface::advect_vector / advect_scalar,
include/shiokaze/advection/macadvection3_interface.h:52-71; module src/advection/macadvection3.cpp):
`MacAdvection3(shape, dx, MacCormack=..., WENO=..., TrimNarrowBand=...)` with the reference's flag names (macadvection3.cpp:57-62).
It only marshals dense numpy grids into the C-ABI (shkz_b200_advect_*): no compute here, no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def _yes(value) -> int:
    if isinstance(value, str):
        return 1 if value.lower() in ("yes", "true", "1") else 0
    return 1 if value else 0


class MacAdvection3:
    """`Advection=b200advection3`. Grids are dense numpy arrays shaped (nz, ny, nx) (+1 along a face grid's own axis), x fastest;
    activity masks are uint8 of the same shapes. Only active entries are rewritten; the masks never change."""

    def __init__(self, shape, dx: float, real: str = "f32", device: int = 0, **flags):
        self.nx, self.ny, self.nz = (int(v) for v in shape)
        self.dx = float(dx)
        self.dtype = np.float64 if real == "f64" else np.float32
        self.params = capi.AdvectParams()
        capi.lib().shkz_b200_advect_default_params(C.byref(self.params))
        self.configure(**flags)
        self._h = C.c_void_p()
        capi.check_advect(capi.lib().shkz_b200_advect_create(self.nx, self.ny, self.nz, self.dx, capi.REAL_F64 if real == "f64" else capi.REAL_F32,
                                                           int(flags.get("GPU", device)), C.byref(self._h)))
        self.last_stats: dict = {}

    def configure(self, **flags):
        p = self.params
        for key, value in flags.items():
            if key == "MacCormack": p.maccormack = _yes(value)
            elif key == "WENO": p.weno = _yes(value)
            elif key == "TrimNarrowBand": p.trim_narrowband = int(value)
            elif key == "GPU": pass
            else:
                raise ValueError(f"unknown flag {key}")

    def face_shape(self, dim):
        return (self.nz + (dim == 2), self.ny + (dim == 1), self.nx + (dim == 0))

    def _faces(self, vel, act):
        v = [np.ascontiguousarray(vel[d], dtype=self.dtype) for d in range(3)]
        a = [np.ascontiguousarray(act[d], dtype=np.uint8) for d in range(3)]
        for d in range(3):
            if v[d].shape != self.face_shape(d) or a[d].shape != self.face_shape(d):
                raise ValueError(f"face grid {d}: shape {v[d].shape} / {a[d].shape}, expected {self.face_shape(d)}")
        return v, a, (C.c_void_p * 3)(*[x.ctypes.data for x in v]), (C.c_void_p * 3)(*[x.ctypes.data for x in a])

    def _cells(self, grid, dtype):
        g = np.ascontiguousarray(grid, dtype=dtype)
        if g.shape != (self.nz, self.ny, self.nx):
            raise ValueError(f"cell grid: shape {g.shape}, expected {(self.nz, self.ny, self.nx)}")
        return g

    def advect_vector(self, u, u_active, fluid, dt: float):
        """macadvection3_interface::advect_vector(u, velocity, fluid, dt) — the reference traces u with itself, its `velocity` argument is unused
        (macadvection3.cpp:79). Returns the three advected face grids (new arrays)."""
        v, a, pv, pa = self._faces(u, u_active)
        v = [x.copy() for x in v]
        pv = (C.c_void_p * 3)(*[x.ctypes.data for x in v])
        fl = None if fluid is None else self._cells(fluid, self.dtype)
        st = capi.AdvectStats()
        capi.check_advect(capi.lib().shkz_b200_advect_vector_host(self._h, float(dt), pv, pa, None if fl is None else fl.ctypes.data, C.byref(self.params), C.byref(st)))
        self.last_stats = st.asdict()
        return v

    def advect_vector_inplace(self, u, u_active, fluid, dt: float):
        """The same call on the caller's own buffers (contiguous arrays of this object's dtype, e.g. numpy views of page-locked memory: the library then moves
        only the values of active faces, include/shkz_b200.h). u is overwritten on its active faces."""
        for d in range(3):
            if not (u[d].flags.c_contiguous and u[d].dtype == self.dtype and u[d].shape == self.face_shape(d)):
                raise ValueError(f"face grid {d}: a contiguous {self.dtype} array of shape {self.face_shape(d)} is needed")
            if not (u_active[d].flags.c_contiguous and u_active[d].dtype == np.uint8 and u_active[d].shape == self.face_shape(d)):
                raise ValueError(f"face mask {d}: a contiguous uint8 array of shape {self.face_shape(d)} is needed")
        fl = None if fluid is None else self._cells(fluid, self.dtype)
        st = capi.AdvectStats()
        capi.check_advect(capi.lib().shkz_b200_advect_vector_host(self._h, float(dt), (C.c_void_p * 3)(*[x.ctypes.data for x in u]),
                                                                 (C.c_void_p * 3)(*[x.ctypes.data for x in u_active]), None if fl is None else fl.ctypes.data,
                                                                 C.byref(self.params), C.byref(st)))
        self.last_stats = st.asdict()
        return self.last_stats

    def advect_scalar(self, q, q_active, vel, vel_active, fluid, dt: float, background: float = 0.0):
        """macadvection3_interface::advect_scalar(scalar, velocity, fluid, dt). background = scalar.get_background_value()
        (what the MacCormack forward result reads off the active set). Returns the advected cell grid (a new array)."""
        v, a, pv, pa = self._faces(vel, vel_active)
        qq = self._cells(q, self.dtype).copy()
        qa = self._cells(q_active, np.uint8)
        fl = None if fluid is None else self._cells(fluid, self.dtype)
        self.params.scalar_background = float(background)
        st = capi.AdvectStats()
        capi.check_advect(capi.lib().shkz_b200_advect_scalar_host(self._h, float(dt), qq.ctypes.data, qa.ctypes.data, pv, pa, None if fl is None else fl.ctypes.data,
                                                                 C.byref(self.params), C.byref(st)))
        self.last_stats = st.asdict()
        return qq

    def advect_vector_device(self, u_ptrs, act_ptrs, fluid_ptr, dt: float, stream=None):
        """shkz_b200_advect_vector_device on raw device pointers (three face grids in/out, their masks, the level set or None) on this object's GPU."""
        st = capi.AdvectStats()
        capi.check_advect(capi.lib().shkz_b200_advect_vector_device(self._h, float(dt), (C.c_void_p * 3)(*u_ptrs), (C.c_void_p * 3)(*act_ptrs), fluid_ptr,
                                                                   C.byref(self.params), C.byref(st), stream))
        self.last_stats = st.asdict()
        return self.last_stats

    def advect_scalar_device(self, q_ptr, qact_ptr, vel_ptrs, vact_ptrs, fluid_ptr, dt: float, background: float = 0.0, stream=None):
        self.params.scalar_background = float(background)
        st = capi.AdvectStats()
        capi.check_advect(capi.lib().shkz_b200_advect_scalar_device(self._h, float(dt), q_ptr, qact_ptr, (C.c_void_p * 3)(*vel_ptrs), (C.c_void_p * 3)(*vact_ptrs),
                                                                   fluid_ptr, C.byref(self.params), C.byref(st), stream))
        self.last_stats = st.asdict()
        return self.last_stats

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().shkz_b200_advect_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
