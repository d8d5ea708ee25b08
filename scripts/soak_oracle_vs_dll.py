"""CPU-only differential soak of the ORACLE against the reference's own libzstd.dll (oracle/_ref): the soak generator's inputs compressed
without a dictionary at levels -131072, -5, -1, 1, 2, 3, 4 and with random raw-content / trained dictionaries at levels -7, 1, 2, 3
(Compressor.LoadDictionary + Wrap), byte for byte; then a damaged copy of a frame of every input decoded by both (same error code or same bytes).
Test infrastructure (it pins the checker, not the product).
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
# damaged frames: the same verdict (error code) or the same bytes
nb = nerr = 0; t0 = time.time()
for k, a in enumerate(inputs):
    f = z.compress(a, int(4 + (a.size * 7919) % 16)) if k % 3 else o.compress(a, 1)
    b = bytearray(f)
    if len(b) > 16:
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(14, len(b)))] ^= 1 << int(rng.integers(0, 8))
    m = bytes(b)
    bound = o.decompress_bound(m); cap = 0 if bound >= 2 ** 62 else bound
    rv, out = o.decompress_raw(m, cap); rr, outr = r.decompress_raw(m, cap)
    oe = bool(o.lib.zo_isError(rv)); nerr += oe
    nb += r.error_code(rr) != o.error_code(rv) or (not oe and outr[:rr].tobytes() != out[:rv].tobytes())
bad += nb
print(f"damaged frames: {nb} disagreements of {n} ({nerr} rejected, {time.time() - t0:.1f}s)", flush=True)
print("ORACLE == DLL" if bad == 0 else f"MISMATCHES: {bad}")
sys.exit(0 if bad == 0 else 1)
