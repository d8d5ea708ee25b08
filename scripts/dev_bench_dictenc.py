"""Developer probe for ncu: one batch of 128 KiB text-like frames compressed with a 32 KiB raw-content dictionary at levels 1 and 3
(device path of WrapBatch).  Usage: dev_bench_dictenc.py [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zstdsharp_b200 import Compressor, datagen as dg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
text = dg.text_like((n + 1) * dg.FRAME)
d = np.ascontiguousarray(text[:32768])
chunks = [text[(i + 1) * dg.FRAME:(i + 2) * dg.FRAME] for i in range(n)]
for level in (1, 3):
    c = Compressor(level); c.LoadDictionary(d)
    for _ in range(2):
        frames = c.WrapBatch(chunks)
    t = c.timings()
    print(f"level {level}: match {t[8]:.2f} ms entropy {t[9]:.2f} ms ratio {n * dg.FRAME / sum(map(len, frames)):.3f}")
    c.Dispose()
