"""Developer probe: end-to-end (pinned host -> host) batch compress through ZSTDB200_compressBatch."""
import ctypes, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from zstdsharp_b200 import datagen as dg, api, _native
workload = sys.argv[1] if len(sys.argv) > 1 else "silesia"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
uniq = min(n, 512)
data = dg.WORKLOADS[workload](uniq * dg.FRAME)
h_in = torch.from_numpy(np.tile(data, n // uniq)).pin_memory()
comp = api.Compressor(1)
bound = comp.GetCompressBound(dg.FRAME); slot = (bound + 15) & ~15
h_out = torch.empty(n * slot, dtype=torch.uint8).pin_memory()
vp = ctypes.c_void_p
sp = (vp * n)(*[h_in.data_ptr() + i * dg.FRAME for i in range(n)]); ss = (ctypes.c_size_t * n)(*([dg.FRAME] * n))
dp = (vp * n)(*[h_out.data_ptr() + i * slot for i in range(n)]); dc = (ctypes.c_size_t * n)(*([bound] * n))
res = (ctypes.c_size_t * n)()
lib = _native.lib
for it in range(4):
    t0 = time.perf_counter()
    rc = lib.ZSTDB200_compressBatch(comp.handle, n, 1, sp, ss, dp, dc, res)
    dt = time.perf_counter() - t0
    assert rc == 0, lib.ZSTDB200_lastErrorString()
    t = comp.timings()
    print(f"iter {it}: wall {1e3*dt:.2f} ms -> {n*dg.FRAME/dt/1e9:.1f} GB/s | h2d span {t[0]:.2f} kernels(sum) {t[1]:.2f} match(sum) {t[8]:.2f} entropy(sum) {t[9]:.2f} d2h span {t[2]:.2f}")
