"""Randomised differential soak of the GPU codec against the oracle, libzstd 1.5.5 and (when present) the reference's own
libzstd.dll -- shared by tests/test_soak_gpu.py (-m gpu) and scripts/soak_gpu.py (longer runs, fresh seeds).

Every input is compressed on the GPU at levels 1..3 and must be byte-identical to the oracle's frame; every GPU frame and a
libzstd frame of a random higher level must decode on the GPU to the input; a quarter of the inputs is compressed and decoded again
with a raw-content and with a trained dictionary; a mutated copy of every frame must give the oracle's answer: the same bytes,
or the same error code.  Failing cases are written to gpurun_out/soakfail_* so that they replay
on the CPU.  The fixed test set missed two decoder defects with a rate of about 1 in 30 000 inputs in round 1; this is the net.
"""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from zstdsharp_b200 import datagen as dg

from _oracle import oracle, libzstd, refdll, refdll_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRAME = dg.FRAME
EDGES = [0, 1, 2, 3, 6, 7, 8, 9, 15, 16, 17, 63, 64, 65, 255, 256, 257, 1023, 1024, 1025, 4095, 4096, 16383, 16384, 16385,
         65535, 65536, 65791, 65792, 65793, FRAME - 1, FRAME, FRAME + 1, FRAME + 6, FRAME + 7, 2 * FRAME - 1, 2 * FRAME, 2 * FRAME + 1, 3 * FRAME + 5]
_POOL = {}


def _pool():
    if not _POOL:
        _POOL.update({k: f(6 * FRAME) for k, f in dg.WORKLOADS.items()})
    return _POOL


def gen_one(rng):
    pool = _pool()
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




def run_soak(n_inputs, seed, comp, dec, log=print, levels=(1, 2, 3), dict_phase=True):
    """Returns the number of disagreements (0 = clean)."""
    rng = np.random.default_rng(seed)
    t0 = time.time()
    inputs = [gen_one(rng) for _ in range(n_inputs)]
    total = sum(a.size for a in inputs)
    log(f"soak seed {seed}: {n_inputs} inputs, {total / 1e6:.1f} MB, generated in {time.time() - t0:.1f}s")
    o, z = oracle(), libzstd()
    r = refdll() if refdll_available() else None
    bad = 0
    dumped = [0]

    def dump(tag, frame, expect, got=None):
        if dumped[0] >= 8:
            return
        dumped[0] += 1
        d = os.path.join(ROOT, "gpurun_out"); os.makedirs(d, exist_ok=True)
        open(os.path.join(d, f"soakfail_{seed}_{tag}.frame"), "wb").write(bytes(frame))
        open(os.path.join(d, f"soakfail_{seed}_{tag}.expect"), "wb").write(bytes(expect))
        if got is not None:
            open(os.path.join(d, f"soakfail_{seed}_{tag}.got"), "wb").write(bytes(got))

    all_frames = {}
    for level in levels:
        comp.Level = level
        t0 = time.time(); frames = comp.WrapBatch(inputs); tg = time.time() - t0
        t0 = time.time()
        with ThreadPoolExecutor(16) as ex:
            want = list(ex.map(lambda a: o.compress(a, level), inputs))
        to = time.time() - t0
        nb = 0
        for i, (f, w) in enumerate(zip(frames, want)):
            if f != w:
                nb += 1
                if nb <= 3:
                    log(f"  MISMATCH level {level} input {i} size {inputs[i].size}: gpu {len(f)} oracle {len(w)}")
                    dump(f"enc_L{level}_{i}", inputs[i].tobytes(), w, f)
        if r is not None:                               # the oracle itself against the reference binary, on a sample of this very run
            for i in range(0, n_inputs, max(1, n_inputs // 300)):
                if want[i] != r.compress(inputs[i], level):
                    nb += 1
                    log(f"  ORACLE != reference DLL, level {level} input {i} size {inputs[i].size}")
        bad += nb
        log(f"level {level}: {nb} mismatches of {n_inputs} (gpu {tg:.1f}s, oracle {to:.1f}s), ratio {total / max(1, sum(map(len, frames))):.3f}")
        all_frames[level] = frames
    with ThreadPoolExecutor(16) as ex:
        zl = list(ex.map(lambda a: z.compress(a, int(4 + (a.size * 7919) % 16)), inputs))
    for name, frames in [(f"gpu L{lv}", all_frames[lv]) for lv in (levels[0], levels[-1])] + [("libzstd L4..19", zl)]:
        outs = dec.UnwrapBatch(frames, raise_on_error=False)
        nb = 0
        for i, (x, a) in enumerate(zip(outs, inputs)):
            if x != a.tobytes():
                nb += 1
                if nb <= 3:
                    log(f"  DECODE FAILURE {name} input {i} size {a.size}: {x if not isinstance(x, (bytes, bytearray)) else len(x)}")
                dump(f"decode_{name.replace(' ', '_').replace('.', '')}_{i}", frames[i], a.tobytes())
        bad += nb
        log(f"decode {name}: {nb} mismatches of {n_inputs}")
    # dictionaries (Compressor.LoadDictionary / Decompressor.LoadDictionary): a raw-content dictionary cut from this run's own data at
    # a random size and a trained zstd-format one, a quarter of the inputs each, one ZSTD_fast and one ZSTD_dfast level
    if dict_phase:
        pool = _pool()
        raw_src = pool[list(pool)[int(rng.integers(0, len(pool)))]]
        raw_len = int(rng.integers(8, 200_000)); raw_off = int(rng.integers(0, raw_src.size - raw_len))
        samples = [pool["text"][i * 3000:(i + 1) * 3000].tobytes() for i in range(200)]
        dicts = [("raw", raw_src[raw_off:raw_off + raw_len].tobytes()), ("zdict", z.train_dictionary(samples, int(rng.integers(2000, 60000))))]
        sub = inputs[::4]
        for dname, d in dicts:
            for level in (levels[0], levels[-1]):
                comp.Level = level
                comp.LoadDictionary(d)
                try:
                    frames = comp.WrapBatch(sub)
                finally:
                    comp.LoadDictionary(None)
                with ThreadPoolExecutor(16) as ex:
                    want = list(ex.map(lambda a: o.compress_loaded_dict(a, level, d), sub))
                nb = 0
                for i, (f, w) in enumerate(zip(frames, want)):
                    if f != w:
                        nb += 1
                        if nb <= 3:
                            log(f"  MISMATCH dictionary {dname} ({len(d)} bytes) level {level} input {4 * i} size {sub[i].size}: gpu {len(f)} oracle {len(w)}")
                            dump(f"dict_{dname}_L{level}_{4 * i}", sub[i].tobytes(), w, f)
                if r is not None:
                    for i in range(0, len(sub), max(1, len(sub) // 60)):
                        if want[i] != r.compress_loaded_dict(sub[i], level, d):
                            nb += 1
                            log(f"  ORACLE != reference DLL, dictionary {dname} level {level} input {4 * i} size {sub[i].size}")
                dec.LoadDictionary(d)
                try:
                    outs = dec.UnwrapBatch(frames, raise_on_error=False)
                finally:
                    dec.LoadDictionary(None)
                for i, (x, a) in enumerate(zip(outs, sub)):
                    if x != a.tobytes():
                        nb += 1
                        if nb <= 3:
                            log(f"  DECODE FAILURE with dictionary {dname} level {level} input {4 * i} size {a.size}")
                bad += nb
                log(f"dictionary {dname} ({len(d)} bytes) level {level}: {nb} mismatches of {len(sub)}")
    # damaged frames: the oracle's answer, error code included (the oracle is held to the reference binary on the same kind of
    # damage by tests/test_reference_pin.py; a sample of this run is checked against it here too)
    mut = []
    for k, f in enumerate(zl):
        src = f if k % 3 else all_frames[levels[0]][k]      # one third: level-1 frames of the GPU encoder (Huffman-heavy blocks)
        b = bytearray(src)
        if len(b) > 16:                                  # behind the frame header: a flipped content-size bit would only change the buffer sizes
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(14, len(b)))] ^= 1 << int(rng.integers(0, 8))
        mut.append(bytes(b))
    res = dec.UnwrapBatch(mut, raise_on_error=False)
    nb = nerr = 0
    for i, (m, g) in enumerate(zip(mut, res)):
        bound = o.decompress_bound(m)
        cap = 0 if bound >= 2 ** 62 else bound
        rv, out = o.decompress_raw(m, cap)             # the capacity UnwrapBatch gives the GPU
        oerr = bool(o.lib.zo_isError(rv))
        gerr = not isinstance(g, (bytes, bytearray))
        nerr += oerr
        wrong = oerr != gerr or (not oerr and out[:rv].tobytes() != g) or (oerr and int(g.Code) != o.error_code(rv))
        if not wrong and r is not None and i % 16 == 0:
            rr, outr = r.decompress_raw(m, cap)
            if r.error_code(rr) != o.error_code(rv) or (not oerr and outr[:rr].tobytes() != out[:rv].tobytes()):
                wrong = True
                log(f"  ORACLE != reference DLL on damaged frame {i}: {o.error_code(rv)} vs {r.error_code(rr)}")
        if wrong:
            nb += 1
            dump(f"corrupt_{i}", m, b"" if oerr else out[:rv].tobytes(), b"" if gerr else g)
            if nb <= 3:
                log(f"  damaged frame {i}: oracle {'error ' + str(o.error_code(rv)) if oerr else rv} gpu {'error ' + str(g) if gerr else len(g)}")
    bad += nb
    log(f"damaged frames: {nb} disagreements of {n_inputs} ({nerr} rejected by the oracle)")
    log(f"SOAK seed {seed} " + ("OK" if bad == 0 else f"FAILED ({bad})"))
    return bad
