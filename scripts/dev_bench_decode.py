"""Developer probe: kernel-level timing of the batch decoder on device-resident frames (not the judged bench)."""
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
level = int(sys.argv[3]) if len(sys.argv) > 3 else 1
uniq = min(nframes, 512)
z = libzstd()
data = dg.WORKLOADS[workload](uniq * dg.FRAME)
chunks = [data[i * dg.FRAME:(i + 1) * dg.FRAME] for i in range(uniq)]
with ThreadPoolExecutor(16) as ex:
    frames = list(ex.map(lambda c: z.compress(c, level), chunks))
frames = [frames[i % uniq] for i in range(nframes)]
sizes = np.array([len(f) for f in frames], dtype=np.uint64)
offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
print(f"{workload}: {nframes} frames, compressed {blob.size/1e6:.1f} MB, ratio {nframes*dg.FRAME/blob.size:.2f}")
d_src = torch.from_numpy(blob.copy()).cuda()
d_dst = torch.empty(nframes * dg.FRAME, dtype=torch.uint8, device="cuda")
dec = api.Decompressor()
n = nframes
so = (ctypes.c_uint64 * n)(*offs.tolist()); ss = (ctypes.c_size_t * n)(*sizes.tolist())
do = (ctypes.c_uint64 * n)(*[i * dg.FRAME for i in range(n)]); dc = (ctypes.c_size_t * n)(*([dg.FRAME] * n))
res = (ctypes.c_size_t * n)()
lib = _native.lib
for it in range(5):
    torch.cuda.synchronize()
    t0 = time.time()
    rc = lib.ZSTDB200_decompressBatchDevice(dec.handle, n, d_src.data_ptr(), so, ss, d_dst.data_ptr(), do, dc, res)
    t1 = time.time()
    assert rc == 0, lib.ZSTDB200_lastErrorString()
    t = dec.timings()
    print(f"iter {it}: wall {1e3*(t1-t0):.2f} ms  kernels {t[1]:.3f} ms  -> {n*dg.FRAME/t[1]/1e6:.1f} GB/s | scan {t[3]:.3f} setup {t[4]:.3f} huf {t[5]:.3f} seq {t[6]:.3f} exec {t[7]:.3f}  launches {dec.launch_count()}")
bad = [i for i in range(n) if res[i] != dg.FRAME]
print("bad results:", len(bad), bad[:5], [hex(res[i]) for i in bad[:3]])
out = d_dst.cpu().numpy()
ok = all(np.array_equal(out[i * dg.FRAME:(i + 1) * dg.FRAME], chunks[i % uniq]) for i in range(0, n, max(1, n // 64)))
print("spot check bit-exact:", ok)
