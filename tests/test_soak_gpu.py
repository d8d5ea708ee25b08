"""Randomised differential soak as a judged GPU test: the fixed parity set missed two decoder defects at a
rate of about 1 in 30 000 inputs; every run covers three regression seeds and one FRESH seed (ZSTDB200_SOAK_SEED, else the
clock; printed, so a failure can be replayed with scripts/soak_gpu.py <n> <seed>)."""
import os
import time

import pytest

from _soak import run_soak

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    from zstdsharp_b200 import Compressor, Decompressor
    c, d = Compressor(1), Decompressor()
    yield c, d
    c.Dispose(); d.Dispose()


@pytest.mark.parametrize("seed", [777001, 777002, 777005])
def test_soak_regression_seeds(codec, seed):
    assert run_soak(1500, seed, *codec) == 0


def test_soak_fresh_seed(codec):
    seed = int(os.environ.get("ZSTDB200_SOAK_SEED", 0)) or int(time.time()) % 1_000_000_000
    print(f"ZSTDB200_SOAK_SEED={seed}")
    assert run_soak(5000, seed, *codec, levels=(-5, -1, 1, 2, 3)) == 0, f"replay with ZSTDB200_SOAK_SEED={seed}"
