"""z-slab decomposition of one projection over the GPUs of an NVLink domain.

The grid is cut along z — the slowest index of the reference's layout (include/shiokaze/math/shape.h:883-888) — into
`world` equal slabs; rank r owns cell planes [r*nz/world, (r+1)*nz/world) of every cell field, the z faces / nodes
of its planes plus the shared plane above. Inside the library the slabs talk through peer memory (halo planes and
CG scalars are stored / loaded by the solve kernels themselves, csrc/slab_comm.cuh); this module only does the
bootstrap, which needs nothing faster than what the host already has:

  * one process per GPU (bench.py under torchrun): `connect(solver, rank, world)` all-gathers the CUDA IPC blobs of
    the slab arenas with torch.distributed (NCCL or gloo — it is 128 bytes per rank, once);
  * one process, several GPUs (tests, a Shiokaze host): `connect_local(solvers)` + `project_local(...)`, which runs
    the per-slab calls on one host thread each (every call blocks on its own GPU).

There is no CPU fallback here either: the functions below only move handles and call the C-ABI.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Sequence, Tuple

import numpy as np

from . import capi, scenes


def slab_range(nz: int, rank: int, world: int) -> Tuple[int, int]:
    """Planes [k0,k1) of rank `rank`; slabs are equal by contract (include/shkz_b200.h)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    if nz % world:
        raise ValueError(f"nz={nz} is not divisible by the number of slabs {world}")
    n = nz // world
    return rank * n, (rank + 1) * n


def slab_scene(workload: str, n: int, nzg: int, zr: Tuple[int, int]) -> "scenes.Scene":
    """Input of one rank. nzg == n: the slab [zr) of the n^3 scene (strong scaling). nzg == n*world with slabs of n
    planes: the weak-scaling workload, an n x n x nzg box in which every slab holds one copy of the n^3 scene (the
    bench scenes are mirror-symmetric in z and carry no z velocity, so the shared planes agree between neighbours)."""
    make = scenes.BENCH_SCENES[workload]
    if nzg == n:
        return make(n, zrange=zr)
    if zr[1] - zr[0] != n or nzg % n:
        raise ValueError("weak-scaling slabs must hold n planes each")
    if workload not in ("smoke_plume", "dambreak", "dambreak_solid"):
        raise ValueError(f"{workload} is not z-symmetric: it cannot be stacked")
    sc = make(n)
    sc.nz = nzg
    sc.zrange = (int(zr[0]), int(zr[1]))
    sc.name = f"{sc.name}_stack{nzg // n}"
    return sc


def split_dense(sc: "scenes.Scene", world: int) -> List["scenes.Scene"]:
    """Cut a whole-grid scene into `world` slab scenes (views are copied; shared z-face / node planes duplicated)."""
    import copy
    out = []
    for r in range(world):
        k0, k1 = slab_range(sc.nz, r, world)
        s = copy.copy(sc)
        s.zrange = (k0, k1)
        s.vel = [np.ascontiguousarray(sc.vel[0][k0:k1]), np.ascontiguousarray(sc.vel[1][k0:k1]), np.ascontiguousarray(sc.vel[2][k0:k1 + 1])]
        s.vel_active = [np.ascontiguousarray(sc.vel_active[0][k0:k1]), np.ascontiguousarray(sc.vel_active[1][k0:k1]),
                        np.ascontiguousarray(sc.vel_active[2][k0:k1 + 1])]
        s.fluid = np.ascontiguousarray(sc.fluid[k0:k1])
        s.solid = np.ascontiguousarray(sc.solid[k0:k1 + 1]) if sc.solid is not None else None
        s.fluid_raw = s.solid_raw = None
        out.append(s)
    return out


def join_dense(parts: Sequence[dict], world: int) -> dict:
    """Inverse of split_dense for project_scene() outputs (the shared z-face planes must agree between neighbours)."""
    vel = [np.concatenate([p["vel"][d] for p in parts], axis=0) for d in range(2)]
    act = [np.concatenate([p["vel_active"][d] for p in parts], axis=0) for d in range(2)]
    for a, b in zip(parts[:-1], parts[1:]):
        if not (np.array_equal(a["vel"][2][-1], b["vel"][2][0]) and np.array_equal(a["vel_active"][2][-1], b["vel_active"][2][0])):
            raise AssertionError("neighbouring slabs disagree on their shared z-face plane")
    vel.append(np.concatenate([p["vel"][2][:-1] for p in parts[:-1]] + [parts[-1]["vel"][2]], axis=0))
    act.append(np.concatenate([p["vel_active"][2][:-1] for p in parts[:-1]] + [parts[-1]["vel_active"][2]], axis=0))
    return dict(vel=vel, vel_active=act, pressure=np.concatenate([p["pressure"] for p in parts], axis=0),
                pressure_active=np.concatenate([p["pressure_active"] for p in parts], axis=0), result=parts[0]["result"])


# ---- one process per GPU ---------------------------------------------------------------------------------
def export_blob(solver) -> bytes:
    buf = (C.c_uint8 * capi.IPC_BYTES)()
    capi.check(solver._L.shkz_b200_slab_export(solver._h, buf), solver._L)
    return bytes(buf)


def connect_blobs(solver, rank: int, world: int, blobs: Sequence[bytes]):
    if len(blobs) != world or any(len(b) != capi.IPC_BYTES for b in blobs):
        raise ValueError("need one IPC blob of capi.IPC_BYTES bytes per rank")
    flat = (C.c_uint8 * (capi.IPC_BYTES * world)).from_buffer_copy(b"".join(blobs))
    capi.check(solver._L.shkz_b200_slab_connect(solver._h, int(rank), int(world), flat), solver._L)


def gather_blobs(blob: bytes, world: int) -> List[bytes]:
    """All-gather of the blobs over the default torch.distributed group (any backend)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return [bytes(p.cpu().numpy().tobytes()) for p in parts]


def connect(solver, rank: int, world: int):
    """Wire `solver` (a slab MacPressureSolver3) to the slabs of the other ranks of the default process group."""
    connect_blobs(solver, rank, world, gather_blobs(export_blob(solver), world))


# ---- one process, several GPUs ---------------------------------------------------------------------------------
def connect_local(solvers: Sequence):
    arr = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
    capi.check(solvers[0]._L.shkz_b200_slab_connect_local(arr, len(solvers)), solvers[0]._L)


def run_per_slab(calls):
    """Run one callable per slab concurrently (each blocks on its own GPU and the slabs wait for each other on the
    device, so they must not be issued one after the other from a single thread)."""
    out, err = [None] * len(calls), [None] * len(calls)

    def work(n):
        try:
            out[n] = calls[n]()
        except BaseException as e:  # noqa: BLE001
            err[n] = e
    threads = [threading.Thread(target=work, args=(n,)) for n in range(len(calls))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


def project_local(solvers: Sequence, slab_scenes: Sequence["scenes.Scene"], **kw) -> List[dict]:
    # one process, several devices: everything a first call allocates is allocated before any slab's kernels start waiting for its neighbours (shkz_b200_prepare)
    for s, sc in zip(solvers, slab_scenes):
        s.prepare(host_buffers=True, have_solid=sc.solid is not None)
    return run_per_slab([(lambda s=s, sc=sc: s.project_scene(sc, **kw)) for s, sc in zip(solvers, slab_scenes)])
