"""Developer probe: level-3 GPU frames vs oracle on a few chunks (run under `timeout`)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from zstdsharp_b200 import datagen as dg, api
from _oracle import oracle
o = oracle()
wl = sys.argv[1] if len(sys.argv) > 1 else "text"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
size = int(sys.argv[3]) if len(sys.argv) > 3 else dg.FRAME
data = dg.WORKLOADS[wl](n * dg.FRAME)
chunks = [data[i * dg.FRAME:i * dg.FRAME + size] for i in range(n)]
c = api.Compressor(3)
print("compressing", flush=True)
frames = c.WrapBatch(chunks)
print("done", [len(f) for f in frames], flush=True)
for i, (ch, f) in enumerate(zip(chunks, frames)):
    want = o.compress(ch, 3)
    if f != want:
        k = next((j for j in range(min(len(f), len(want))) if f[j] != want[j]), -1)
        print("chunk", i, "differs at", k, "sizes", len(f), len(want))
    else:
        print("chunk", i, "identical", len(f))
