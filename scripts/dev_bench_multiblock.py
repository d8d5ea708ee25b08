"""Developer probe: host-pointer Wrap of inputs above 128 KiB (multi-block frames vs independent chunks)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from zstdsharp_b200 import datagen as dg, api

comp = api.Compressor(1)
text = dg.text_like(80 * dg.FRAME)
def run(label, chunks, reps=3):
    total = sum(c.size for c in chunks)
    best = 1e9
    for _ in range(reps):
        t0 = time.time(); fr = comp.WrapBatch(chunks); best = min(best, time.time() - t0)
    print(f"{label}: {len(chunks)} items, {total/1e6:.1f} MB in {best*1e3:.1f} ms -> {total/best/1e9:.2f} GB/s, ratio {total/sum(len(f) for f in fr):.3f}, launches {comp.launch_count()}")
comp.WrapBatch([text[:dg.FRAME]])
one = text[:10_000_000]
run("one 10 MB frame (exact, 77 blocks)", [one])
comp.SetParameter(10001, 1); run("one 10 MB input, independent chunks", [one]); comp.SetParameter(10001, 0)
run("1024 x 1 MiB frames (exact, 8 blocks each)", [text[(i % 9) * dg.FRAME:(i % 9) * dg.FRAME + (1 << 20)] for i in range(1024)], reps=2)
run("4096 x 256 KiB frames (exact, 2 blocks each)", [text[(i % 70) * dg.FRAME:(i % 70) * dg.FRAME + (1 << 18)] for i in range(4096)], reps=2)
run("8192 x 128 KiB frames", [text[(i % 80) * dg.FRAME:(i % 80 + 1) * dg.FRAME] for i in range(8192)], reps=2)
