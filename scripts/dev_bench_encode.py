"""Developer probe: kernel-level timing of the batch encoder on device-resident chunks (not the judged bench)."""
import ctypes, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from zstdsharp_b200 import datagen as dg, api, _native
from _oracle import oracle

workload = sys.argv[1] if len(sys.argv) > 1 else "silesia"
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
level = int(sys.argv[3]) if len(sys.argv) > 3 else 1
uniq = min(nframes, 512)
data = dg.WORKLOADS[workload](uniq * dg.FRAME)
d_u = torch.from_numpy(data).cuda()
ids = torch.arange(nframes, device="cuda") % uniq
d_src = d_u.view(uniq, dg.FRAME)[ids].contiguous().view(-1)
comp = api.Compressor(level)
bound = comp.GetCompressBound(dg.FRAME); slot = (bound + 15) & ~15
d_dst = torch.empty(nframes * slot, dtype=torch.uint8, device="cuda")
n = nframes
so = (ctypes.c_uint64 * n)(*[i * dg.FRAME for i in range(n)]); ss = (ctypes.c_size_t * n)(*([dg.FRAME] * n))
do = (ctypes.c_uint64 * n)(*[i * slot for i in range(n)]); dc = (ctypes.c_size_t * n)(*([bound] * n))
res = (ctypes.c_size_t * n)()
lib = _native.lib
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    rc = lib.ZSTDB200_compressBatchDevice(comp.handle, n, level, d_src.data_ptr(), so, ss, d_dst.data_ptr(), do, dc, res)
    t1 = time.time()
    assert rc == 0, lib.ZSTDB200_lastErrorString()
    t = comp.timings()
    print(f"iter {it}: wall {1e3*(t1-t0):.2f} ms kernels {t[1]:.3f} ms -> {n*dg.FRAME/t[1]/1e6:.1f} GB/s | match {t[8]:.3f} entropy {t[9]:.3f} launches {comp.launch_count()}")
sizes = np.array(list(res), dtype=np.int64)
print("ratio %.3f" % (n * dg.FRAME / sizes.sum()), "errors", int((sizes > bound).sum()))
o = oracle(); host = d_dst.cpu().numpy(); bad = 0
for i in range(0, min(n, uniq), max(1, uniq // 32)):
    want = o.compress(data[i * dg.FRAME:(i + 1) * dg.FRAME], level)
    got = host[i * slot:i * slot + sizes[i]].tobytes()
    bad += got != want
print("byte-identical spot check: bad =", bad)
