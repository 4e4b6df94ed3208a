"""Host-side logic of the z-slab decomposition (no GPU): partition, slab scenes, split / join, and the bootstrap
all-gather of the arena blobs over torch.distributed with the gloo backend, world_size 2."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from shiokaze_b200 import capi, dist, scenes


def test_slab_range_partitions_the_grid():
    for nz, world in ((64, 1), (64, 2), (512, 4), (1024, 8)):
        r = [dist.slab_range(nz, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == nz
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        assert len({b - a for a, b in r}) == 1
    with pytest.raises(ValueError):
        dist.slab_range(30, 0, 4)
    with pytest.raises(ValueError):
        dist.slab_range(32, 4, 4)


def test_split_and_join_round_trip():
    sc = scenes.flip_splash(16)
    for world in (2, 4):
        parts = dist.split_dense(sc, world)
        assert [p.zrange for p in parts] == [dist.slab_range(16, r, world) for r in range(world)]
        for p in parts:
            nzl = p.zrange[1] - p.zrange[0]
            assert p.vel[2].shape == (nzl + 1, 16, 16) and p.solid.shape == (nzl + 1, 17, 17) and p.fluid.shape == (nzl, 16, 16)
        fake = [dict(vel=p.vel, vel_active=p.vel_active, pressure=p.fluid, pressure_active=(p.fluid < 0).astype(np.uint8), result=None) for p in parts]
        whole = dist.join_dense(fake, world)
        for d in range(3):
            assert np.array_equal(whole["vel"][d], sc.vel[d]) and np.array_equal(whole["vel_active"][d], sc.vel_active[d])
        assert np.array_equal(whole["pressure"], sc.fluid)
    bad = [dict(vel=[v.copy() for v in p.vel], vel_active=p.vel_active, pressure=p.fluid, pressure_active=p.fluid, result=None) for p in dist.split_dense(sc, 2)]
    bad[1]["vel"][2][0] += 1.0
    with pytest.raises(AssertionError):
        dist.join_dense(bad, 2)


def test_slab_scene_strong_and_weak():
    whole = scenes.dambreak(16, True)
    part = dist.slab_scene("dambreak_solid", 16, 16, (8, 16))       # strong scaling: a cut of the n^3 scene
    assert np.array_equal(part.fluid, whole.fluid[8:16]) and np.array_equal(part.vel[2], whole.vel[2][8:17])
    for name in ("smoke_plume", "dambreak_solid"):                    # weak scaling: stacked copies, shared planes must agree
        a = dist.slab_scene(name, 16, 32, (0, 16))
        b = dist.slab_scene(name, 16, 32, (16, 32))
        assert a.nz == b.nz == 32 and b.zrange == (16, 32)
        assert np.array_equal(a.vel[2][-1], b.vel[2][0]) and np.array_equal(a.vel_active[2][-1], b.vel_active[2][0])
        if a.solid is not None:
            assert np.array_equal(a.solid[-1], b.solid[0])
    with pytest.raises(ValueError):
        dist.slab_scene("flip_splash", 16, 32, (0, 16))


def test_connect_argument_checks():
    with pytest.raises(ValueError):
        dist.connect_blobs(None, 0, 2, [b"x" * capi.IPC_BYTES])
    with pytest.raises(ValueError):
        dist.connect_blobs(None, 0, 1, [b"short"])


def test_blob_all_gather_over_gloo_world_size_2(tmp_path):
    worker = tmp_path / "worker.py"
    worker.write_text(
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch.distributed as td\n"
        "from shiokaze_b200 import capi, dist\n"
        "td.init_process_group('gloo')\n"
        "rank, world = td.get_rank(), td.get_world_size()\n"
        "blob = bytes([rank + 1]) * capi.IPC_BYTES\n"
        "got = dist.gather_blobs(blob, world)\n"
        "assert got == [bytes([r + 1]) * capi.IPC_BYTES for r in range(world)], got\n"
        "assert dist.slab_range(64, rank, world) == (32 * rank, 32 * rank + 32)\n"
        f"open(os.path.join({str(tmp_path)!r}, f'ok{{rank}}'), 'w').write('ok')\n"
        "td.destroy_process_group()\n")
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(worker)], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()   # (stdout of the two ranks may interleave)
