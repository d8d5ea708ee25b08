"""Longer randomised differential soak (developer probe; the judged version is tests/test_soak_gpu.py).
Usage: soak_gpu.py [n_inputs] [seed] [levels, e.g. -5,1,3]     -- run after every decoder / encoder change with a FRESH seed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from zstdsharp_b200 import api
from _soak import run_soak

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 1
LEVELS = tuple(int(x) for x in sys.argv[3].split(',')) if len(sys.argv) > 3 else (1, 2, 3)
bad = run_soak(N, SEED, api.Compressor(1), api.Decompressor(), log=lambda m: print(m, flush=True), levels=LEVELS)
sys.exit(0 if bad == 0 else 1)
