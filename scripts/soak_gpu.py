"""Randomised differential soak of the GPU codec against the oracle and libzstd 1.5.5 (developer probe, not the judged bench).

Every input is compressed on the GPU at levels 1..3 and must be byte-identical to the oracle's frame; every GPU frame and a
libzstd frame of a random higher level must decode on the GPU to the input; a mutated copy of every frame must give the
oracle's answer (same bytes or an error on both sides, never a crash).  Usage: soak_gpu.py [n_inputs] [seed]
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from concurrent.futures import ThreadPoolExecutor
from zstdsharp_b200 import datagen as dg, api
from _oracle import oracle, libzstd

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(SEED)
FRAME = dg.FRAME
EDGES = [0, 1, 2, 3, 6, 7, 8, 9, 15, 16, 17, 63, 64, 65, 255, 256, 257, 1023, 1024, 1025, 4095, 4096, 16383, 16384, 16385,
         65535, 65536, 65791, 65792, 65793, FRAME - 1, FRAME, FRAME + 1, FRAME + 6, FRAME + 7, 2 * FRAME - 1, 2 * FRAME, 2 * FRAME + 1, 3 * FRAME + 5]
pool = {k: f(6 * FRAME) for k, f in dg.WORKLOADS.items()}


def gen_one():
    kind = rng.integers(0, 10)
    r = rng.random()
    size = int(EDGES[rng.integers(0, len(EDGES))]) if r < 0.35 else int(rng.integers(0, 3 * FRAME)) if r < 0.5 else int(rng.integers(0, FRAME + 1))
    if size == 0:
        return np.zeros(0, dtype=np.uint8)
    if kind <= 3:                                    # slice of a workload at a random offset
        src = pool[list(pool)[rng.integers(0, len(pool))]]
        o = int(rng.integers(0, src.size - min(size, src.size) + 1))
        a = src[o:o + size].copy()
        if a.size < size:
            a = np.resize(a, size)
        return a
    if kind == 4:                                    # periodic with a random period, a few mutations
        period = int(rng.integers(1, 70000))
        base = rng.integers(0, 256, size=period, dtype=np.uint8)
        a = np.resize(base, size)
        for _ in range(int(rng.integers(0, 20))):
            a[rng.integers(0, size)] ^= np.uint8(rng.integers(1, 256))
        return a
    if kind == 5:                                    # small alphabet (literal heavy), skewed
        k = int(rng.integers(1, 17))
        pz = rng.random(k) ** 3 + 1e-3
        return rng.choice(np.arange(k, dtype=np.uint8) * np.uint8(rng.integers(1, 15)), size=size, p=pz / pz.sum()).astype(np.uint8)
    if kind == 6:                                    # runs of random lengths (RLE-ish blocks, long matches)
        out = np.empty(size, dtype=np.uint8); pos = 0
        while pos < size:
            L = int(min(size - pos, rng.integers(1, 1 << int(rng.integers(1, 18)))))
            out[pos:pos + L] = rng.integers(0, 256); pos += L
        return out
    if kind == 7:                                    # random bytes with copied segments (matches at random offsets)
        a = rng.integers(0, 256, size=size, dtype=np.uint8)
        for _ in range(int(rng.integers(0, 200))):
            L = int(rng.integers(3, 300)); s = int(rng.integers(0, max(1, size - L))); d = int(rng.integers(0, max(1, size - L)))
            a[d:d + L] = a[s:s + L].copy()
        return a
    if kind == 8:                                    # concatenation of different regimes (block type changes inside a frame)
        parts = []; left = size
        while left > 0:
            L = int(min(left, rng.integers(1, FRAME)))
            m = rng.integers(0, 4)
            parts.append(np.zeros(L, np.uint8) if m == 0 else rng.integers(0, 256, size=L, dtype=np.uint8) if m == 1 else pool["text"][:L] if m == 2 else pool["literal_heavy"][:L])
            left -= L
        return np.concatenate(parts)
    return (np.arange(size) * int(rng.integers(1, 7)) >> int(rng.integers(0, 4))).astype(np.uint8)   # ramps


t0 = time.time()
inputs = [gen_one() for _ in range(N)]
total = sum(a.size for a in inputs)
print(f"{N} inputs, {total / 1e6:.1f} MB, generated in {time.time() - t0:.1f}s", flush=True)
o, z = oracle(), libzstd()
DRY = bool(os.environ.get('SOAK_DRY'))
comp, dec = (None, None) if DRY else (api.Compressor(1), api.Decompressor())
if DRY:
    class _C:
        Level = 1
        def WrapBatch(self, xs): return [o.compress(a, self.Level) for a in xs]
    class _D:
        def UnwrapBatch(self, fs, raise_on_error=True):
            r = []
            for f in fs:
                b = o.decompress_bound(f); rv, out = o.decompress_raw(f, 0 if b >= 2 ** 62 else b)
                r.append(Exception('err') if o.lib.zo_isError(rv) else out[:rv].tobytes())
            return r
    comp, dec = _C(), _D()
bad = 0
allFrames = []
_dumped = [0]


def _dump(tag, frame, expect, got=None):              # failing cases go to gpurun_out/ so that they can be replayed on the CPU
    if _dumped[0] >= 8:
        return
    _dumped[0] += 1
    d = os.path.join(ROOT, "gpurun_out"); os.makedirs(d, exist_ok=True)
    open(os.path.join(d, f"soakfail_{SEED}_{tag}.frame"), "wb").write(bytes(frame))
    open(os.path.join(d, f"soakfail_{SEED}_{tag}.expect"), "wb").write(bytes(expect))
    if got is not None:
        open(os.path.join(d, f"soakfail_{SEED}_{tag}.got"), "wb").write(bytes(got))


for level in (1, 2, 3):
    comp.Level = level
    t0 = time.time(); frames = comp.WrapBatch(inputs); tg = time.time() - t0
    t0 = time.time()
    with ThreadPoolExecutor(16) as ex:
        want = list(ex.map(lambda a: o.compress(a, level), inputs))
    to = time.time() - t0
    nb = sum(f != w for f, w in zip(frames, want))
    bad += nb
    for i, (f, w) in enumerate(zip(frames, want)):
        if f != w:
            print(f"  MISMATCH level {level} input {i} size {inputs[i].size}: gpu {len(f)} oracle {len(w)}"); break
    print(f"level {level}: {nb} mismatches of {N} (gpu {tg:.1f}s, oracle {to:.1f}s), ratio {total / max(1, sum(map(len, frames))):.3f}", flush=True)
    allFrames.append(frames)
# decode: GPU frames of all levels + libzstd frames of random levels
with ThreadPoolExecutor(16) as ex:
    zl = list(ex.map(lambda a: z.compress(a, int(4 + (a.size * 7919) % 16)), inputs))
for name, frames in (("gpu L1", allFrames[0]), ("gpu L3", allFrames[2]), ("libzstd L4..19", zl)):
    outs = dec.UnwrapBatch(frames, raise_on_error=False)
    nb = sum(x != a.tobytes() for x, a in zip(outs, inputs))
    for i, (x, a) in enumerate(zip(outs, inputs)):
        if x != a.tobytes():
            print(f"  DECODE FAILURE {name} input {i} size {a.size}: {x if not isinstance(x, (bytes, bytearray)) else len(x)}")
            _dump(f"decode_{name.replace(' ', '_').replace('.', '')}_{i}", frames[i], a.tobytes())
    bad += nb
    print(f"decode {name}: {nb} mismatches of {N}", flush=True)
# corrupted frames: same answer as the oracle (bytes, or an error on both sides)
mut = []
for f in zl:
    b = bytearray(f)
    if len(b) > 16:                                  # behind the frame header: a flipped content-size bit would only change the buffer sizes
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(14, len(b)))] ^= 1 << int(rng.integers(0, 8))
    mut.append(bytes(b))
res = dec.UnwrapBatch(mut, raise_on_error=False)
nb = 0; nerr = 0
for i, (m, r) in enumerate(zip(mut, res)):
    bound = o.decompress_bound(m)
    rv, out = o.decompress_raw(m, 0 if bound >= 2 ** 62 else bound)             # the capacity UnwrapBatch gives the GPU
    oerr = bool(o.lib.zo_isError(rv))
    gerr = not isinstance(r, (bytes, bytearray))
    nerr += oerr
    if oerr != gerr or (not oerr and out[:rv].tobytes() != r):
        nb += 1
        _dump(f"corrupt_{i}", m, b"" if oerr else out[:rv].tobytes(), b"" if gerr else r)
        if nb <= 3: print(f"  corrupted frame {i}: oracle {'error' if oerr else rv} gpu {'error ' + str(r) if gerr else len(r)}")
bad += nb
print(f"corrupted frames: {nb} disagreements of {N} ({nerr} rejected by the oracle)", flush=True)
print("SOAK", "OK" if bad == 0 else f"FAILED ({bad})")
sys.exit(0 if bad == 0 else 1)
