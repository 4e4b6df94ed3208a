"""shkz_b200_project_host with page-locked buffers on liquid scenes moves only what the projection touches (csrc/kernels_xfer.cuh): the caller's buffers
must end up BYTE FOR BYTE as with whole-array copies (SHKZ_B200_HOST_COPIES=dense), call after call on the same buffers, while far fewer bytes cross PCIe."""
import ctypes as C
import os

import numpy as np
import pytest

from shiokaze_b200 import MacPressureSolver3, capi, scenes

pytestmark = pytest.mark.gpu


class Pinned:
    """numpy views over shkz_b200_host_alloc memory (what Array=b200array3 hands to the module)."""

    def __init__(self):
        self.L = capi.lib()
        self.ptrs = []

    def empty(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        capi.check(self.L.shkz_b200_host_alloc(n, C.byref(p)))
        self.ptrs.append(p)
        return np.frombuffer((C.c_ubyte * n).from_address(p.value), dtype=dtype).reshape(shape)

    def like(self, a, dtype=None):
        out = self.empty(a.shape, dtype or a.dtype)
        out[...] = a
        return out

    def free(self):
        for p in self.ptrs:
            self.L.shkz_b200_host_free(p)
        self.ptrs = []


def run(S, sc, bufs, rt):
    for d in range(3):
        bufs["vel"][d][...] = sc.vel[d].astype(rt)
        bufs["act"][d][...] = sc.vel_active[d]
    bufs["fluid"][...] = sc.fluid.astype(rt)
    if sc.solid is not None:
        bufs["solid"][...] = sc.solid.astype(rt)
    _, _, res = S.project(sc.dt, bufs["vel"], bufs["act"], bufs["solid"] if sc.solid is not None else None, bufs["fluid"], sc.fluid_levelset,
                          pressure_out=bufs["pressure"], pressure_active_out=bufs["pact"])
    snap = dict(vel=[v.copy() for v in bufs["vel"]], act=[a.copy() for a in bufs["act"]], pressure=bufs["pressure"].copy(), pact=bufs["pact"].copy())
    return snap, res


def alloc(pin, n3, rt):
    nx, ny, nz = n3
    shapes = [(nz, ny, nx + 1), (nz, ny + 1, nx), (nz + 1, ny, nx)]
    return dict(vel=[pin.empty(s, rt) for s in shapes], act=[pin.empty(s, np.uint8) for s in shapes], fluid=pin.empty((nz, ny, nx), rt),
                solid=pin.empty((nz + 1, ny + 1, nx + 1), rt), pressure=pin.empty((nz, ny, nx), rt), pact=pin.empty((nz, ny, nx), np.uint8))


def same(a, b):
    for d in range(3):
        assert np.array_equal(a["act"][d], b["act"][d]), d
        assert a["vel"][d].tobytes() == b["vel"][d].tobytes(), (d, float(np.abs(a["vel"][d] - b["vel"][d]).max()))
    assert np.array_equal(a["pact"], b["pact"])
    assert a["pressure"].tobytes() == b["pressure"].tobytes()


SEQUENCES = {
    # one solver, one set of buffers, several projections in a row: the wet region moves, shrinks, and comes back
    "dambreak72": [lambda: scenes.dambreak(72, True), lambda: scenes.flip_splash(72), lambda: scenes.dambreak(72, True)],
    "flip96": [lambda: scenes.flip_splash(96), lambda: scenes.dambreak(96)],
    "blobs": [lambda: scenes.random_blobs(70, 37, 29, seed=11), lambda: scenes.random_blobs(70, 37, 29, seed=12, with_solid=False)],
}


@pytest.mark.parametrize("name", list(SEQUENCES))
@pytest.mark.parametrize("real", ["f32", "f64"])
def test_sparse_host_copies_equal_whole_array_copies(cuda_device, name, real, monkeypatch):
    rt = np.float32 if real == "f32" else np.float64
    seq = [f() for f in SEQUENCES[name]]
    n3 = (seq[0].nx, seq[0].ny, seq[0].nz)
    pin = Pinned()
    outs = {}
    try:
        for mode in ("dense", "sparse"):
            if mode == "dense":
                monkeypatch.setenv("SHKZ_B200_HOST_COPIES", "dense")
            else:
                monkeypatch.delenv("SHKZ_B200_HOST_COPIES", raising=False)
            S = MacPressureSolver3(n3, seq[0].dx, real=real, Precision="fp64" if real == "f64" else "mixed")
            bufs = alloc(pin, n3, rt)
            for b in (bufs["pressure"], bufs["pact"]):
                b[...] = 77  # garbage the first call has to clear
            outs[mode] = [run(S, sc, bufs, rt) for sc in seq]
            S.close()
        for (a, ra), (b, rb) in zip(outs["dense"], outs["sparse"]):
            same(a, b)
            assert ra.iterations == rb.iterations and ra.n_rows == rb.n_rows
            assert ra.stats["host_copies"] == 0 and rb.stats["host_copies"] == 1, (ra.stats["host_copies"], rb.stats["host_copies"])
        # fewer bytes over PCIe (on the toy grid of "blobs" the 64 x 16 x 8 transfer blocks cover most of the grid: nothing to save there)
        a, b = outs["dense"][-1][1].stats, outs["sparse"][-1][1].stats
        if name != "blobs":
            assert b["h2d_bytes"] < a["h2d_bytes"] and b["d2h_bytes"] < a["d2h_bytes"], (a["h2d_bytes"], a["d2h_bytes"], b["h2d_bytes"], b["d2h_bytes"])
    finally:
        pin.free()


def test_pageable_buffers_and_all_fluid_scenes_keep_whole_array_copies(cuda_device):
    sc = scenes.dambreak(40, True)
    S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx)
    out = S.project_scene(sc)  # numpy (pageable) buffers
    assert out["result"].stats["host_copies"] == 0
    S.close()
    sm = scenes.smoke_plume(40)
    pin = Pinned()
    try:
        S = MacPressureSolver3((sm.nx, sm.ny, sm.nz), sm.dx)
        bufs = alloc(pin, (sm.nx, sm.ny, sm.nz), np.float32)
        _, res = run(S, sm, bufs, np.float32)
        assert res.stats["host_copies"] == 0 and res.converged
        S.close()
    finally:
        pin.free()


def test_a_device_call_between_two_host_calls_does_not_leave_stale_pressure(cuda_device):
    """The sparse path only rewrites the tiles of this and the previous projection; a projection through another entry point in between
    moves the solver's tile lists, so the next host call must clear the caller's grids again."""
    import torch
    a, b = scenes.dambreak(72, True), scenes.flip_splash(72)
    pin = Pinned()
    try:
        S = MacPressureSolver3((72, 72, 72), a.dx)
        bufs = alloc(pin, (72, 72, 72), np.float32)
        first, _ = run(S, a, bufs, np.float32)
        dev = torch.device("cuda", 0)
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        S.project_device(b.dt, [t(v) for v in b.vel], [t(m) for m in b.vel_active], t(b.solid), t(b.fluid), b.fluid_levelset)
        again, res = run(S, b, bufs, np.float32)
        assert res.stats["host_copies"] == 1
        S.close()
        os.environ["SHKZ_B200_HOST_COPIES"] = "dense"
        try:
            S = MacPressureSolver3((72, 72, 72), a.dx)
            ref, _ = run(S, b, bufs, np.float32)
            S.close()
        finally:
            del os.environ["SHKZ_B200_HOST_COPIES"]
        same(again, ref)
    finally:
        pin.free()


@pytest.mark.parametrize("name", ["dambreak_solid", "flip", "smoke"])
def test_velocity_masked_reads_inactive_faces_as_zero(cuda_device, name, monkeypatch):
    """params.velocity_masked (what the Shiokaze module sets when it hands over velocity grids of the dense array core without rewriting their inactive entries):
    junk on the inactive faces changes nothing — sparse host copies and whole-array host copies give the results of the clean input, bit for bit. (Whether junk
    WITHOUT the flag is visible depends on the scene: the assembly reads the faces of wet cells only, and on a dam-break all of those are active.)"""
    sc = {"dambreak_solid": lambda: scenes.dambreak(40, True), "flip": lambda: scenes.flip_splash(48), "smoke": lambda: scenes.smoke_plume(24)}[name]()
    rng = np.random.default_rng(11)
    junk = [np.where(a != 0, v, 1e3 * rng.standard_normal(v.shape)).astype(np.float32) for v, a in zip(sc.vel, sc.vel_active)]

    def run(vel, pinned, **flags):
        S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, **flags)
        wrap = pinned.like if pinned else (lambda x: x.copy())
        v, a = [wrap(x) for x in vel], [wrap(x) for x in sc.vel_active]
        p, pa, res = S.project(sc.dt, v, a, None if sc.solid is None else wrap(sc.solid), wrap(sc.fluid), sc.fluid_levelset,
                               pressure_out=wrap(np.zeros(sc.fluid.shape, np.float32)), pressure_active_out=wrap(np.zeros(sc.fluid.shape, np.uint8)))
        S.close()
        return [x.copy() for x in v], [x.copy() for x in a], p.copy(), res

    clean_v, clean_a, clean_p, clean = run(sc.vel, None)
    P = Pinned()
    try:
        for mode in ("sparse", "dense"):
            if mode == "dense":
                monkeypatch.setenv("SHKZ_B200_HOST_COPIES", "dense")
            v, a, p, res = run(junk, P, VelocityMasked=1)
            assert res.iterations == clean.iterations, mode
            if mode == "dense" or not sc.fluid_levelset:
                assert res.stats["host_copies"] == 0
            assert np.array_equal(p, clean_p), mode
            for d in range(3):
                on = sc.vel_active[d] != 0
                assert np.array_equal(a[d], clean_a[d]), (mode, d)
                assert np.array_equal(v[d][on], clean_v[d][on]), (mode, d)
        monkeypatch.delenv("SHKZ_B200_HOST_COPIES", raising=False)
    finally:
        P.free()
