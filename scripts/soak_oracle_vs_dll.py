"""CPU-only differential soak of the ORACLE against the reference's own libzstd.dll (oracle/_ref): the soak generator's inputs compressed
without a dictionary at levels -131072, -5, -1, 1, 2, 3, 4 and with random raw-content / trained dictionaries at levels -7, 1, 2, 3
(Compressor.LoadDictionary + Wrap), byte for byte.  Test infrastructure (it pins the checker, not the product).
Usage: soak_oracle_vs_dll.py [seed] [n_inputs]          -- needs /root/reference at build time (oracle/_ref/libzstdref.so)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _oracle import oracle, refdll, libzstd
from _soak import gen_one, _pool

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
o, r, z = oracle(), refdll(), libzstd()
rng = np.random.default_rng(seed)
inputs = [gen_one(rng) for _ in range(n)]
bad = 0
for level in (-131072, -5, -1, 1, 2, 3, 4):
    nb = skipped = 0; t0 = time.time()
    for a in inputs:
        rv, out = o.compress_raw(a, level)
        if o.lib.zo_isError(rv):
            skipped += 1; continue                      # level 4 where it is ZSTD_greedy: not restated
        nb += out[:rv].tobytes() != r.compress(a, level)
    bad += nb
    print(f"no dictionary, level {level}: {nb} mismatches of {n - skipped} ({time.time() - t0:.1f}s)", flush=True)
pool = _pool()
for rnd in range(4):
    if rnd % 2 == 0:
        src = pool[list(pool)[int(rng.integers(0, len(pool)))]]
        L = int(rng.integers(8, 250_000)); off = int(rng.integers(0, src.size - L)); d = src[off:off + L].tobytes(); name = f"raw content, {L} bytes"
    else:
        d = z.train_dictionary([pool["text"][i * 3000:(i + 1) * 3000].tobytes() for i in range(200)], int(rng.integers(1000, 110000))); name = f"trained, {len(d)} bytes"
    for level in (-7, 1, 2, 3):
        nb = 0; t0 = time.time()
        for a in inputs:
            nb += r.compress_loaded_dict(a, level, d) != o.compress_loaded_dict(a, level, d)
        bad += nb
        print(f"dictionary ({name}), level {level}: {nb} mismatches of {n} ({time.time() - t0:.1f}s)", flush=True)
print("ORACLE == DLL" if bad == 0 else f"MISMATCHES: {bad}")
sys.exit(0 if bad == 0 else 1)
