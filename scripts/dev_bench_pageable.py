"""Developer probe: host-to-host decode of 8192 x 128 KiB frames from PAGEABLE, separately allocated buffers vs pinned contiguous ones."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from zstdsharp_b200 import api, _native, datagen as dg
lib = _native.lib
FRAME = dg.FRAME
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
uniq = 256
text = dg.text_like(uniq * FRAME)
comp, dec = api.Compressor(1), api.Decompressor()
frames_u = comp.WrapBatch([text[i * FRAME:(i + 1) * FRAME] for i in range(uniq)])
vp, st = ctypes.c_void_p, ctypes.c_size_t
src = [np.frombuffer(frames_u[i % uniq], dtype=np.uint8).copy() for i in range(n)]
dst = [np.empty(FRAME, dtype=np.uint8) for _ in range(n)]
for a in dst: a[::4096] = 0
sp = (vp * n)(*[a.ctypes.data for a in src]); ss = (st * n)(*[a.size for a in src])
dp = (vp * n)(*[a.ctypes.data for a in dst]); dc = (st * n)(*([FRAME] * n)); res = (st * n)()
def run(tag, sp, dp, reps=4):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = lib.ZSTDB200_decompressBatch(dec.handle, n, sp, ss, dp, dc, res)
        ts.append(time.perf_counter() - t0)
        assert rc == 0 and all(r == FRAME for r in res)
    t = dec.timings()
    print(f"{tag}: best {min(ts[1:]) * 1e3:.2f} ms = {n * FRAME / min(ts[1:]) / 1e9:.1f} GB/s   (h2d {t[0]:.1f} kernels {t[1]:.1f} d2h {t[2]:.1f})", flush=True)
run("pageable scattered", sp, dp)
assert all(np.array_equal(dst[i], text[(i % uniq) * FRAME:(i % uniq + 1) * FRAME]) for i in range(0, n, 37))
tot = sum(a.size for a in src)
h_c = torch.empty(tot + 64, dtype=torch.uint8).pin_memory(); h_o = torch.empty(n * FRAME, dtype=torch.uint8).pin_memory()
off = 0; offs = []
for a in src:
    h_c.numpy()[off:off + a.size] = a; offs.append(off); off += a.size
sp2 = (vp * n)(*[h_c.data_ptr() + o for o in offs]); dp2 = (vp * n)(*[h_o.data_ptr() + i * FRAME for i in range(n)])
run("pinned contiguous", sp2, dp2)
# raw host memcpy rate of this box: 16 threads, 128 KiB pieces
big = np.empty(n * FRAME, dtype=np.uint8); big[::4096] = 1
from concurrent.futures import ThreadPoolExecutor
def cp(t):
    for i in range(t, n, 16): dst[i][:] = big[i * FRAME:(i + 1) * FRAME]
with ThreadPoolExecutor(16) as ex:
    list(ex.map(cp, range(16)))
    t0 = time.perf_counter(); list(ex.map(cp, range(16))); dt = time.perf_counter() - t0
print(f"numpy 16-thread copy of {n} x 128 KiB: {n * FRAME / dt / 1e9:.1f} GB/s")
