"""Developer probe: device-resident compression of many small records (ZSTDB200_compressBatchDevice), wall time per call against kernel time.
Usage: dev_bench_records.py [record_bytes] [records] [level]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from zstdsharp_b200 import api, _native, datagen as dg
lib = _native.lib
rec = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
lvl = int(sys.argv[3]) if len(sys.argv) > 3 else 1
text = dg.text_like(n * rec + 64)
d_src = torch.from_numpy(text).cuda()
slot = (int(lib.ZSTD_compressBound(rec)) + 15) & ~15
d_dst = torch.empty(n * slot + 64, dtype=torch.uint8, device="cuda")
u64, st = ctypes.c_uint64, ctypes.c_size_t
so = (u64 * n)(*[i * rec for i in range(n)]); ss = (st * n)(*([rec] * n)); do = (u64 * n)(*[i * slot for i in range(n)]); dc = (st * n)(*([slot] * n)); res = (st * n)()
c = api.Compressor(lvl)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.ZSTDB200_compressBatchDevice(c.handle, n, lvl, d_src.data_ptr(), so, ss, d_dst.data_ptr(), do, dc, res)
    dt = (time.perf_counter() - t0) * 1e3
    t = c.timings()
    print(f"call {it}: rc {rc} wall {dt:.2f} ms, kernels {t[1]:.2f} ms (match {t[8]:.2f} entropy {t[9]:.2f})", flush=True)
