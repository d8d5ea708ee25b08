"""Developer probe: end-to-end (pinned host -> host) batch decode through ZSTDB200_decompressBatch."""
import ctypes, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from zstdsharp_b200 import datagen as dg, api, _native
from _oracle import libzstd
from concurrent.futures import ThreadPoolExecutor
workload = sys.argv[1] if len(sys.argv) > 1 else "text"
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
uniq = min(nframes, 512)
z = libzstd()
data = dg.WORKLOADS[workload](uniq * dg.FRAME)
chunks = [data[i * dg.FRAME:(i + 1) * dg.FRAME] for i in range(uniq)]
with ThreadPoolExecutor(16) as ex:
    frames = list(ex.map(lambda c: z.compress(c, 1), chunks))
frames = [frames[i % uniq] for i in range(nframes)]
sizes = np.array([len(f) for f in frames], dtype=np.int64)
offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
h_src = torch.from_numpy(blob.copy()).pin_memory()
h_dst = torch.empty(nframes * dg.FRAME, dtype=torch.uint8).pin_memory()
n = nframes
vp = ctypes.c_void_p
sp = (vp * n)(*[h_src.data_ptr() + int(o) for o in offs]); ss = (ctypes.c_size_t * n)(*sizes.tolist())
dp = (vp * n)(*[h_dst.data_ptr() + i * dg.FRAME for i in range(n)]); dc = (ctypes.c_size_t * n)(*([dg.FRAME] * n))
res = (ctypes.c_size_t * n)()
dec = api.Decompressor()
lib = _native.lib
best = 1e9
for it in range(6):
    t0 = time.perf_counter()
    rc = lib.ZSTDB200_decompressBatch(dec.handle, n, sp, ss, dp, dc, res)
    dt = time.perf_counter() - t0
    assert rc == 0, lib.ZSTDB200_lastErrorString()
    t = dec.timings()
    if it: best = min(best, dt)
    print(f"iter {it}: wall {1e3*dt:.2f} ms -> {n*dg.FRAME/dt/1e9:.1f} GB/s | h2d span {t[0]:.2f} kernels(sum) {t[1]:.2f} d2h span {t[2]:.2f}")
assert all(r == dg.FRAME for r in res)
out = h_dst.numpy()
ok = all(np.array_equal(out[i * dg.FRAME:(i + 1) * dg.FRAME], chunks[i % uniq]) for i in range(0, n, max(1, n // 64)))
print(f"PIPE={os.environ.get('ZSTDB200_PIPE')} ITEMS={os.environ.get('ZSTDB200_PIPE_ITEMS')} best {1e3*best:.2f} ms {n*dg.FRAME/best/1e9:.1f} GB/s bit-exact {ok}")
