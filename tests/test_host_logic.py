import numpy as np

from zstdsharp_b200 import datagen as dg
from zstdsharp_b200.sharding import shard_bounds, shard_bounds_native


def test_datagen_is_deterministic_and_shaped():
    for name, fn in dg.WORKLOADS.items():
        a, b = fn(4 * dg.FRAME), fn(4 * dg.FRAME)
        assert a.dtype == np.uint8 and a.size == 4 * dg.FRAME, name
        assert np.array_equal(a, b), name
    assert np.array_equal(dg.byte_ramp(600)[250:260], np.array([250, 251, 252, 253, 254, 255, 0, 1, 2, 3], dtype=np.uint8))
    t = dg.tile_to(dg.text_like(2 * dg.FRAME), 5 * dg.FRAME)
    assert t.size == 5 * dg.FRAME and np.array_equal(t[: dg.FRAME], t[2 * dg.FRAME: 3 * dg.FRAME])


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 8, 100, 8192):
        w = rng.integers(22, 131085, size=n)
        for world in (1, 2, 3, 4, 8):
            b = shard_bounds(w, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(lo <= hi for lo, hi in b)
            assert shard_bounds_native(w.tolist(), world) == b          # the C++ scheduler (ZSTDB200_*BatchMulti) cuts identically
            if n >= 8 * world:
                sums = [int(w[lo:hi].sum()) for lo, hi in b]
                assert max(sums) - min(sums) <= 2 * 131085
    assert shard_bounds([5] * 8, 8) == [(i, i + 1) for i in range(8)]
