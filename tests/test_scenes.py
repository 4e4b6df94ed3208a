"""CPU tests of the synthetic scene generators (host logic)."""
import numpy as np

from shiokaze_b200 import scenes


def test_hash_noise_is_counter_based():
    a = scenes.hash_noise(7, 1, (6, 5, 4), 0)
    b = scenes.hash_noise(7, 1, (3, 5, 4), 3)
    assert np.array_equal(a[3:], b)
    assert 0.0 <= a.min() and a.max() < 1.0
    assert not np.array_equal(a, scenes.hash_noise(8, 1, (6, 5, 4), 0))
    # known answer: splitmix64(0) of the reference implementation
    assert int(scenes.splitmix64(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF


def test_slab_generation_matches_whole_grid():
    for make in (scenes.flip_splash, scenes.liquid_box, scenes.smoke_plume, lambda n, **kw: scenes.dambreak(n, True, **kw)):
        whole = make(16)
        part = make(16, zrange=(4, 12))
        assert part.nzl == 8
        assert np.array_equal(part.fluid, whole.fluid[4:12])
        for d in range(3):
            hi = 13 if d == 2 else 12
            assert np.array_equal(part.vel[d], whole.vel[d][4:hi])
            assert np.array_equal(part.vel_active[d], whole.vel_active[d][4:hi])
        if whole.solid is not None:
            assert np.array_equal(part.solid_raw, whole.solid_raw[4:13])


def test_levelset_clamp_matches_narrow_band_read():
    raw = np.array([-1.0, -0.2, -0.05, 0.0, 0.05, 0.2, 1.0], dtype=np.float32)
    out = scenes.clamp_levelset(raw, 0.1)
    assert np.allclose(out, [-0.1, -0.1, -0.05, 0.0, 0.05, 0.1, 0.1])


def test_scene_shapes():
    s = scenes.random_blobs(9, 7, 5)
    assert s.fluid.shape == (5, 7, 9) and s.solid.shape == (6, 8, 10)
    assert [v.shape for v in s.vel] == [(5, 7, 10), (5, 8, 9), (6, 7, 9)]
    assert all(v.dtype == np.float32 for v in s.vel) and all(a.dtype == np.uint8 for a in s.vel_active)
    # inactive faces read as the background 0
    for v, a in zip(s.vel, s.vel_active):
        assert float(np.abs(v[a == 0]).max(initial=0.0)) == 0.0
