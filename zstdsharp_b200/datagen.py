"""Seeded synthetic corpora for the hot path.

All generators use ``numpy.random.Generator(PCG64(seed))`` and return ``numpy.uint8`` arrays whose length is a
multiple of the 131072-byte frame size (the last partial chunk is dropped).  They replace the reference's
fixtures, which are not available here (``dickens`` is absent from the mount: /root/reference/.MISSING_LARGE_BLOBS:1;
ZstdNetTests.GenerateSample depends on .NET's seeded PRNG: src/ZstdSharp.Test/ZstdNetTests.cs:607-615).

* ``text_like``       -- "dickens-like": Zipf(1.1) draws from a 4096-word vocabulary, sentences, ~70-column lines.
* ``silesia_mix``     -- per-chunk class mix: text / XML-ish / binary structs / x86-like / 16-bit sensor / random.
* ``incompressible``  -- uniform random bytes (raw blocks).
* ``literal_heavy``   -- i.i.d. order-0 skewed bytes (~5 bit/B): Huffman 4-stream literals, almost no sequences.
* ``literal_mix``     -- 50 % literal_heavy, 25 % incompressible, 25 % constant-byte chunks (config 4).
* ``byte_ramp``       -- ``(byte)i`` data of the reference's size sweep (ZstdNetTests.cs:617-622).
"""
from __future__ import annotations

import numpy as np

FRAME = 131072

SEED_TEXT = 0xD1C3
SEED_SILESIA = 0x51E5
SEED_INCOMPRESSIBLE = 0xBAD5EED
SEED_LITERAL = 0x11735

_ENGLISH = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_ENGLISH_P = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0,
                       1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])
_ENGLISH_P = _ENGLISH_P / _ENGLISH_P.sum()


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _vocab(rng: np.random.Generator, nwords: int = 4096):
    lens = rng.integers(2, 11, size=nwords)
    mat = np.full((nwords, 12), 0, dtype=np.uint8)
    letters = rng.choice(_ENGLISH, size=(nwords, 10), p=_ENGLISH_P)
    mat[:, :10] = letters
    return mat, lens


def _text_bytes(rng: np.random.Generator, nbytes: int, mat: np.ndarray, lens: np.ndarray, cdf: np.ndarray) -> np.ndarray:
    """Vectorised word stream of at least nbytes bytes."""
    out = []
    have = 0
    col = np.arange(12)[None, :]
    while have < nbytes:
        nw = max(1024, (nbytes - have) // 5)
        idx = np.searchsorted(cdf, rng.random(nw))
        idx = np.minimum(idx, len(lens) - 1)
        rows = mat[idx].copy()
        wl = lens[idx]
        # sentence ends every 8..20 words
        gaps = rng.integers(8, 21, size=nw // 8 + 2)
        ends = np.cumsum(gaps)
        ends = ends[ends < nw]
        eos = np.zeros(nw, dtype=bool)
        eos[ends] = True
        # capitalise the word after a sentence end
        cap = np.zeros(nw, dtype=bool)
        cap[np.minimum(ends + 1, nw - 1)] = True
        cap[0] = True
        rows[cap, 0] -= 32
        # punctuation + separator
        rows[np.arange(nw), wl] = np.where(eos, ord("."), ord(" "))
        rows[eos, wl[eos] + 1] = ord(" ")
        tot = wl + 1 + eos
        # line breaks roughly every 70 columns: turn the separator crossing a multiple of 70 into '\n'
        endpos = np.cumsum(tot)
        brk = (endpos // 70) != ((endpos - tot) // 70)
        rows[np.arange(nw)[brk], (tot - 1)[brk]] = ord("\n")
        mask = col < tot[:, None]
        chunk = rows[mask]
        out.append(chunk)
        have += chunk.size
    return np.concatenate(out)[:nbytes]


def text_like(nbytes: int, seed: int = SEED_TEXT) -> np.ndarray:
    nbytes = (nbytes // FRAME) * FRAME if nbytes >= FRAME else nbytes
    rng = _rng(seed)
    mat, lens = _vocab(rng)
    p = 1.0 / np.arange(1, len(lens) + 1) ** 1.1
    cdf = np.cumsum(p / p.sum())
    return _text_bytes(rng, nbytes, mat, lens, cdf)


def incompressible(nbytes: int, seed: int = SEED_INCOMPRESSIBLE) -> np.ndarray:
    nbytes = (nbytes // FRAME) * FRAME if nbytes >= FRAME else nbytes
    return _rng(seed).integers(0, 256, size=nbytes, dtype=np.uint8)


def _skewed(rng: np.random.Generator, n: int, ratio: float = 0.955) -> np.ndarray:
    p = ratio ** np.arange(256)
    cdf = np.cumsum(p / p.sum())
    sym = np.searchsorted(cdf, rng.random(n)).astype(np.uint8)
    perm = rng.permutation(256).astype(np.uint8)
    return perm[sym]


def literal_heavy(nbytes: int, seed: int = SEED_LITERAL) -> np.ndarray:
    nbytes = (nbytes // FRAME) * FRAME if nbytes >= FRAME else nbytes
    return _skewed(_rng(seed), nbytes)


def literal_mix(nbytes: int, seed: int = SEED_LITERAL) -> np.ndarray:
    """Config 4: HUF 4-stream literals + raw blocks + RLE-literal paths."""
    nchunks = max(1, nbytes // FRAME)
    rng = _rng(seed)
    kinds = rng.choice(3, size=nchunks, p=[0.5, 0.25, 0.25])
    out = np.empty(nchunks * FRAME, dtype=np.uint8)
    for i, k in enumerate(kinds):
        sl = slice(i * FRAME, (i + 1) * FRAME)
        if k == 0:
            out[sl] = _skewed(rng, FRAME)
        elif k == 1:
            out[sl] = rng.integers(0, 256, size=FRAME, dtype=np.uint8)
        else:
            out[sl] = rng.integers(0, 256)
    return out


def _xmlish(rng: np.random.Generator, n: int) -> np.ndarray:
    tags = [b"record", b"id", b"name", b"value", b"timestamp", b"status", b"item", b"price"]
    parts = []
    size = 0
    i = int(rng.integers(0, 100000))
    while size < n:
        t = tags[int(rng.integers(1, len(tags)))]
        v = str(int(rng.integers(0, 10 ** int(rng.integers(1, 7))))).encode()
        s = b"<record id=\"%d\"><%s>%s</%s></record>\n" % (i, t, v, t)
        parts.append(s)
        size += len(s)
        i += 1
    return np.frombuffer(b"".join(parts)[:n], dtype=np.uint8).copy()


def _structs(rng: np.random.Generator, n: int) -> np.ndarray:
    nrec = n // 32 + 1
    rec = np.zeros((nrec, 32), dtype=np.uint8)
    ctr = (np.arange(nrec, dtype=np.uint32) + rng.integers(0, 1 << 20)).astype("<u4")
    rec[:, 0:4] = ctr.view(np.uint8).reshape(nrec, 4)
    walk = np.cumsum(rng.integers(-3, 4, size=nrec)).astype("<i4")
    rec[:, 4:8] = walk.view(np.uint8).reshape(nrec, 4)
    rec[:, 8:12] = np.frombuffer(b"\x00\x01\x00\x00", dtype=np.uint8)
    rec[:, 12] = rng.integers(0, 4, size=nrec)
    rec[:, 16:20] = (rng.integers(0, 1000, size=nrec).astype("<u4")).view(np.uint8).reshape(nrec, 4)
    rec[:, 24:28] = np.frombuffer(b"ABCD", dtype=np.uint8)
    return rec.reshape(-1)[:n].copy()


def _x86ish(rng: np.random.Generator, n: int) -> np.ndarray:
    base = _skewed(rng, n, 0.97)
    # short repeats: copy 4..16 byte snippets from up to 4 KB back
    nrep = n // 24
    pos = np.sort(rng.integers(4096, n - 32, size=nrep))
    ln = rng.integers(4, 17, size=nrep)
    back = rng.integers(8, 4096, size=nrep)
    for p, l, b in zip(pos.tolist(), ln.tolist(), back.tolist()):
        base[p:p + l] = base[p - b:p - b + l]
    return base


def _sensor(rng: np.random.Generator, n: int) -> np.ndarray:
    ns = n // 2 + 1
    t = np.arange(ns)
    sig = 2000 * np.sin(t / 97.0) + 500 * np.sin(t / 13.0) + np.cumsum(rng.normal(0, 3, size=ns))
    return sig.astype("<i2").view(np.uint8)[:n].copy()


def silesia_mix(nbytes: int, seed: int = SEED_SILESIA) -> np.ndarray:
    nchunks = max(1, nbytes // FRAME)
    rng = _rng(seed)
    mat, lens = _vocab(rng)
    p = 1.0 / np.arange(1, len(lens) + 1) ** 1.1
    cdf = np.cumsum(p / p.sum())
    kinds = rng.choice(6, size=nchunks, p=[0.35, 0.15, 0.20, 0.15, 0.10, 0.05])
    out = np.empty(nchunks * FRAME, dtype=np.uint8)
    for i, k in enumerate(kinds):
        sl = slice(i * FRAME, (i + 1) * FRAME)
        if k == 0:
            out[sl] = _text_bytes(rng, FRAME, mat, lens, cdf)
        elif k == 1:
            out[sl] = _xmlish(rng, FRAME)
        elif k == 2:
            out[sl] = _structs(rng, FRAME)
        elif k == 3:
            out[sl] = _x86ish(rng, FRAME)
        elif k == 4:
            out[sl] = _sensor(rng, FRAME)
        else:
            out[sl] = rng.integers(0, 256, size=FRAME, dtype=np.uint8)
    return out


def byte_ramp(n: int) -> np.ndarray:
    """GenerateBuffer of the reference tests: bytes i % 256 (ZstdNetTests.cs:617-622)."""
    return (np.arange(n) % 256).astype(np.uint8)


def tile_to(data: np.ndarray, nbytes: int) -> np.ndarray:
    """Repeat a generated corpus chunk-wise up to nbytes (used to reach 1 GiB quickly: every frame stays a genuine
    sample of the generator; frames are independent, so tiling changes neither per-frame work nor ratios)."""
    reps = -(-nbytes // data.size)
    return np.tile(data, reps)[:nbytes]


WORKLOADS = {
    "text": text_like,
    "silesia": silesia_mix,
    "incompressible": incompressible,
    "literal_heavy": literal_heavy,
    "literal_mix": literal_mix,
}
