"""Developer probe for compute-sanitizer: a small pass through every kernel (decode, encode levels 1 and 3, multi-block
frames, checksum), checked against the oracle.  Run as: compute-sanitizer --tool memcheck python scripts/dev_memcheck.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from zstdsharp_b200 import datagen as dg, api
from _oracle import oracle

o = oracle()
F = dg.FRAME
text, sil, lit = dg.text_like(6 * F), dg.silesia_mix(6 * F), dg.literal_heavy(F)
chunks = [text[i * F:(i + 1) * F] for i in range(4)] + [sil[i * F:(i + 1) * F] for i in range(6)] + [lit, text[:100], text[:5], text[:0], np.zeros(70000, np.uint8), dg.incompressible(F)]
multi = [text[:F + 6], text[:300_000], np.concatenate([text[:F], np.zeros(F, np.uint8), text[:40]]), sil[:2 * F + 1]]
comp, dec = api.Compressor(1), api.Decompressor()
bad = 0
for level in (1, 3):
    comp.Level = level
    for cs in (0, 1):
        comp.SetParameter(201, cs)
        frames = comp.WrapBatch(chunks + multi)
        for c, f in zip(chunks + multi, frames):
            bad += f != o.compress(c, level, checksum=cs)
        out = dec.UnwrapBatch(frames)
        bad += sum(a != c.tobytes() for a, c in zip(out, chunks + multi))
comp.SetParameter(201, 0)
# frames of other levels and concatenated frames through the decoder
from _oracle import libzstd
z = libzstd()
fr = [z.compress(text[:3 * F], lvl) for lvl in (5, 9, 19)] + [z.compress(text[:F], 1) + z.compress(sil[:F], 3)]
want = [text[:3 * F].tobytes()] * 3 + [text[:F].tobytes() + sil[:F].tobytes()]
bad += sum(a != b for a, b in zip(dec.UnwrapBatch(fr), want))
print("memcheck probe: mismatches =", bad)
sys.exit(1 if bad else 0)
