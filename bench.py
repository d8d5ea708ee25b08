#!/usr/bin/env python
"""Benchmark of the hot path: batch decompress (and level-1 compress) of independent 128 KiB zstd frames.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's own native zstd 1.5.1 (oracle/_ref: its libzstd.dll
                                                           # through the PE mapper), else the oracle port, on all host cores

One JSON line on stdout (rank 0).  `value` = decompress GB/s of uncompressed bytes, whole job, frames resident in HBM
(configs[1] of BASELINE.json: 1 GiB of 128 KiB level-1 frames per GPU, weak scaling); `e2e` = the same through the
host-pointer C ABI (ZSTDB200_decompressBatch) with pinned HOST buffers, H2D and D2H inside the timed region;
`compress_l1` carries the level-1 compress numbers of configs[2] (1 GiB Silesia-mix-like in 128 KiB chunks).
A step = one pass of the pipeline over the rank's whole batch.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from zstdsharp_b200 import datagen as dg  # noqa: E402
from zstdsharp_b200.sharding import shard_bounds  # noqa: E402

FRAME = dg.FRAME
METRIC = "decompress & level-1 compress GB/s (128KB frames) at 1/2/4/8 B200 vs CPU"
UNIQUE_FRAMES = 4096         # 512 MiB of generated corpus (8 differently seeded 64 MiB blocks), tiled 2x up to the batch size


def make_corpus(fn, seed0: int, uniq: int) -> np.ndarray:
    """`uniq` frames of workload `fn`: blocks of 512 frames, block k generated with seed0 + 512 k (block 0 = the workload's default seed)."""
    parts = [fn(min(512, uniq - k) * FRAME, seed0 + k) if k else fn(min(512, uniq) * FRAME) for k in range(0, uniq, 512)]
    return parts[0] if len(parts) == 1 else np.concatenate(parts)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
#  CPU side (oracle port of the reference algorithm): --impl reference and the cpu_baseline leg
# ----------------------------------------------------------------------------------------------------------------
def cpu_frames(level: int = 1, nframes: int = 512, workload: str = "text"):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import oracle
    o = oracle()
    data = make_corpus(dg.WORKLOADS[workload], dg.SEED_TEXT if workload == "text" else dg.SEED_SILESIA, nframes) if workload in ("text", "silesia") else dg.WORKLOADS[workload](nframes * FRAME)
    chunks = [data[i * FRAME:(i + 1) * FRAME] for i in range(nframes)]
    threads = os.cpu_count() or 1
    with ThreadPoolExecutor(threads) as ex:
        frames = list(ex.map(lambda c: o.compress(c, level), chunks))
    return o, chunks, frames, threads


def cpu_decode_pass(o, frames, threads, seconds_min=0.0):
    """Decompresses every frame once with `threads` workers (one context per call, as ZstdNetTests.cs:498-522).
    Returns (bytes, seconds)."""
    outs = [np.empty(FRAME, dtype=np.uint8) for _ in range(threads)]
    bufs = [np.frombuffer(f, dtype=np.uint8) for f in frames]
    lib = o.lib

    ctxs = [lib.zo_createDCtx() for _ in range(threads)]     # one Decompressor per thread

    def work(tid):
        out = outs[tid]
        n = 0
        for i in range(tid, len(bufs), threads):
            r = lib.zo_decompressDCtx(ctxs[tid], out.ctypes.data, FRAME, bufs[i].ctypes.data, bufs[i].size)
            assert r == FRAME
            n += r
        return n
    total = 0
    with ThreadPoolExecutor(threads) as ex:
        t0 = time.perf_counter()
        while True:
            total += sum(ex.map(work, range(threads)))
            if time.perf_counter() - t0 >= seconds_min:
                break
        dt = time.perf_counter() - t0
    for c in ctxs:
        lib.zo_freeDCtx(c)
    return total, dt


def cpu_compress_pass(o, chunks, threads, level=1):
    lib = o.lib
    cap = lib.zo_compressBound(FRAME)
    outs = [np.empty(cap, dtype=np.uint8) for _ in range(threads)]

    ctxs = [lib.zo_createCCtx() for _ in range(threads)]     # one Compressor per thread

    def work(tid):
        n = 0
        for i in range(tid, len(chunks), threads):
            c = chunks[i]
            r = lib.zo_compressCCtx(ctxs[tid], outs[tid].ctypes.data, cap, c.ctypes.data, c.size, level, 0)
            assert r < (1 << 62)
            n += c.size
        return n
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))                   # warm-up pass
        t0 = time.perf_counter()
        total = sum(ex.map(work, range(threads)))
        dt = time.perf_counter() - t0
    for c in ctxs:
        lib.zo_freeCCtx(c)
    return total, dt


def native_libzstd_decode_rate(frames, threads):
    """Context only: upstream libzstd 1.5.5 (faster than the managed reference, README.md:44-58) on the same frames."""
    try:
        from _oracle import libzstd
        z = libzstd().lib
    except Exception:
        return None
    bufs = [np.frombuffer(f, dtype=np.uint8) for f in frames]
    outs = [np.empty(FRAME, dtype=np.uint8) for _ in range(threads)]

    def work(tid):
        n = 0
        for _ in range(4):
            for i in range(tid, len(bufs), threads):
                n += z.ZSTD_decompress(outs[tid].ctypes.data, FRAME, bufs[i].ctypes.data, bufs[i].size)
        return n
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        total = sum(ex.map(work, range(threads)))
    return total / (time.perf_counter() - t0) / 1e9


def ref_dll():
    """The reference's own native library (src/Zstd.Extern/libzstd.dll, zstd 1.5.1) through oracle/_ref, or None.  ZstdSharp's
    managed code is a translation of exactly this code; the reference's README.md:44-58 measures it 1.41x slower than this
    binary, so timing the binary is the conservative stand-in for "ZstdSharp on .NET" (which this image cannot run)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        from _oracle import refdll, refdll_available
        if refdll_available():
            return refdll()
    except Exception as e:                                   # noqa: BLE001
        log("bench.py: oracle/_ref unavailable:", e)
    return None


class DllBatch:
    """n frames decoded (or chunks compressed) by the DLL on `threads` host threads, one context per thread
    (ZREF_decompressBatchMT / ZREF_compressBatchMT in oracle/ref_pe/peload.c; contexts as in ZstdNetTests.cs:498-522)."""

    def __init__(self, r, srcs, dst_cap, threads):
        self.r, self.n, self.threads = r, len(srcs), threads
        self.srcs = srcs
        vp, st = ctypes.c_void_p, ctypes.c_size_t
        self.out = np.empty(self.n * dst_cap, dtype=np.uint8)
        self.out[::4096] = 0                                  # touch the pages outside the timed region
        self.sp = (vp * self.n)(*[s.ctypes.data for s in srcs])
        self.ss = (st * self.n)(*[s.size for s in srcs])
        self.dp = (vp * self.n)(*[self.out.ctypes.data + i * dst_cap for i in range(self.n)])
        self.dc = (st * self.n)(*([dst_cap] * self.n))
        self.res = (st * self.n)()
        self.cap = dst_cap

    def decompress(self):
        self.r.lib.ZREF_decompressBatchMT(self.n, self.sp, self.ss, self.dp, self.dc, self.res, self.threads)

    def compress(self, level):
        self.r.lib.ZREF_compressBatchMT(self.n, self.sp, self.ss, self.dp, self.dc, self.res, level, self.threads)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    r = ref_dll()
    nframes = args.frames
    uniq = min(args.unique, nframes)
    o, chunks, frames, _ = cpu_frames(1, uniq, "text")      # frames: the oracle's = the DLL's bytes (tests/test_reference_pin.py)
    port_b, port_t = cpu_decode_pass(o, frames[:256], threads, seconds_min=1.0)
    port = port_b / port_t / 1e9
    if r is None:                                            # no oracle/_ref on this box: the port is the arm
        for _ in range(args.warmup):
            cpu_decode_pass(o, frames, threads)
        t0 = time.perf_counter(); total = 0
        for _ in range(args.steps):
            b, _ = cpu_decode_pass(o, frames, threads); total += b
        dt = time.perf_counter() - t0
        val, kind = total / dt / 1e9, "port"
        sample = f"{uniq} text-like 128 KiB level-1 frames ({uniq * FRAME >> 20} MiB) per step, {threads} threads, oracle port of the C#"
        comp_val = None
    else:
        srcs = [np.frombuffer(frames[i % uniq], dtype=np.uint8) for i in range(nframes)]
        job = DllBatch(r, srcs, FRAME, threads)
        for _ in range(max(1, args.warmup)):
            job.decompress()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            job.decompress()
        dt = time.perf_counter() - t0
        assert all(x == FRAME for x in job.res), "the reference DLL failed to decode a frame"
        want = np.concatenate(chunks).reshape(uniq, FRAME)
        assert np.array_equal(job.out.reshape(nframes, FRAME)[:uniq], want), "reference DLL output differs from the corpus"
        val, kind = nframes * FRAME * args.steps / dt / 1e9, "reference"
        sample = (f"the whole configs[1] batch every step: {nframes} text-like 128 KiB level-1 frames ({nframes * FRAME >> 20} MiB out) decoded by the "
                  f"reference's own libzstd.dll (zstd 1.5.1, oracle/_ref) on {threads} host threads, one ZSTD_DCtx per thread")
        del job
        sil = make_corpus(dg.silesia_mix, dg.SEED_SILESIA, uniq).reshape(uniq, FRAME)
        cj = DllBatch(r, [sil[i % uniq] for i in range(nframes)], r.lib.ZREF_compressBound(FRAME), threads)
        cj.compress(1)
        t1 = time.perf_counter(); cj.compress(1); ct = time.perf_counter() - t1
        comp_val = nframes * FRAME / ct / 1e9
        del cj
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "batch decompress 1 GiB of 128 KiB level-1 frames, bit-exact (BASELINE.json configs[1])", "frames_per_gpu": nframes,
                   "frame_bytes": FRAME, "corpus": f"text_like seed 0x{dg.SEED_TEXT:X} (+512 per 64 MiB block): {uniq} unique frames tiled to {nframes}",
                   "note": "reference arm = the reference's own native zstd 1.5.1 binary (src/Zstd.Extern/libzstd.dll through oracle/ref_pe); ZstdSharp's "
                           "managed translation of this code cannot run here (no .NET) and is 1.41x slower by the reference's README.md:44-58"},
        "cpu_baseline": {"value": round(val, 4), "unit": "GB/s", "cores": threads, "kind": kind, "sample": sample},
        "oracle_port_decompress_GBps": round(port, 4),
        "compress_l1": None if comp_val is None else {"value": round(comp_val, 4), "unit": "GB/s", "sample": f"{nframes} Silesia-mix-like 128 KiB chunks, level 1, same DLL and threads"},
        "native_libzstd_1_5_5_decompress_GBps": native_libzstd_decode_rate(frames[:256], threads),
        "e2e": {"value": round(val, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
#  GPU side
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    # stdout carries exactly one JSON line: libraries (NCCL prints its version banner there) are sent to stderr meanwhile
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    import torch
    import torch.distributed as dist
    from zstdsharp_b200 import api, _native
    lib = _native.lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if lib.ZSTDB200_deviceCount() == 0:
        raise SystemExit("bench.py: no CUDA device; zstdsharp_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # one process per GPU: keep this rank's threads and the pinned buffers it allocates on the CPUs next to its GPU's PCIe root
    numa_rc = lib.ZSTDB200_bindThreadToDevice(local) if os.environ.get("ZSTDB200_NUMA_BIND", "1") != "0" else None
    # The ranks exchange nothing on the data path (frames share no state): the process group only carries the barrier and three
    # scalars (max step time, byte totals), so it is a CPU group (gloo over loopback); NCCL only if gloo cannot come up here.
    coll_device = "cpu"
    if world > 1:
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
        try:
            dist.init_process_group("gloo")
        except Exception as e:                                   # noqa: BLE001
            log("bench.py: gloo group failed (%r), using nccl for the bookkeeping" % (e,))
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            coll_device = "cuda"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=coll_device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=coll_device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nframes = args.frames                                    # per GPU (weak scaling)
    # host scatter: the global frame list (world * nframes frames, index g -> unique frame g % UNIQUE) is cut into
    # contiguous ranges balanced by byte weight; this rank materialises its own range only.
    uniq = min(args.unique, nframes)
    lo, hi = shard_bounds([FRAME] * (world * nframes), world)[rank]
    my_ids = np.arange(lo, hi) % uniq
    n = len(my_ids)

    stream = torch.cuda.Stream()
    comp, dec = api.Compressor(1), api.Decompressor()
    # contexts allocate lazily: give them the torch stream so that torch events bracket their work
    def use_stream(ctx):
        r = lib.ZSTDB200_setStream(ctx.handle, ctypes.c_void_p(stream.cuda_stream))
        assert r == 0, lib.ZSTDB200_lastErrorString()
    use_stream(comp); use_stream(dec)

    def dev_call(fn, ctx, nitems, *a):
        rc = fn(ctx.handle, nitems, *a)
        if rc != 0:
            raise SystemExit(f"bench.py: batch call failed: {lib.ZSTD_getErrorName(rc)} / {lib.ZSTDB200_lastErrorString()}")

    u64, st = ctypes.c_uint64, ctypes.c_size_t
    results = {}
    sampler = ClockSampler(local)

    def timed_steps(step_fn, warmup, steps):
        for _ in range(warmup):
            step_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches, slots = 0, np.zeros(_native.TIMING_SLOTS)
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                l, s = step_fn()
                launches += l; slots += s
            e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        return max_over_ranks(dev_ms), max_over_ranks(wall * 1e3), launches, slots / steps

    # ------------------------------------------------------------------ corpus + compressed frames (made by the GPU encoder)
    t_prep = time.perf_counter()
    text = make_corpus(dg.text_like, dg.SEED_TEXT, uniq)
    sil = make_corpus(dg.silesia_mix, dg.SEED_SILESIA, uniq)
    d_text_u = torch.from_numpy(text).cuda()
    d_sil_u = torch.from_numpy(sil).cuda()
    bound = comp.GetCompressBound(FRAME)
    slot = (bound + 15) & ~15

    def compress_unique(d_u, level=1):
        m = d_u.numel() // FRAME
        d_out = torch.empty(m * slot, dtype=torch.uint8, device="cuda")
        so = (u64 * m)(*[i * FRAME for i in range(m)]); ss = (st * m)(*([FRAME] * m))
        do = (u64 * m)(*[i * slot for i in range(m)]); dc = (st * m)(*([bound] * m)); res = (st * m)()
        dev_call(lib.ZSTDB200_compressBatchDevice, comp, m, level, d_u.data_ptr(), so, ss, d_out.data_ptr(), do, dc, res)
        sizes = np.array(list(res), dtype=np.int64)
        assert (sizes > 0).all() and (sizes <= bound).all(), "GPU compressor reported an error"
        host = d_out.cpu().numpy()
        return [host[i * slot:i * slot + sizes[i]].copy() for i in range(m)]

    frames_u = compress_unique(d_text_u)
    # parity gate outside every timed region (rank 0): the GPU encoder's frames of the bench corpus are the oracle's frames --
    # byte for byte, levels 1 and 3, both corpora (the oracle is held to the reference's libzstd.dll by tests/test_reference_pin.py)
    parity = None
    if rank == 0 and not args.skip_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _oracle import oracle as _oracle_fn
        _o = _oracle_fn()
        m_chk = min(uniq, 512)
        with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
            for name, host_u, d_u in (("text", text, d_text_u), ("silesia", sil, d_sil_u)):
                for level in (1, 3):
                    got = frames_u if (name == "text" and level == 1) else compress_unique(d_u[:m_chk * FRAME], level)
                    want = list(ex.map(lambda i: _o.compress(host_u[i * FRAME:(i + 1) * FRAME], level), range(m_chk)))
                    bad = [i for i in range(m_chk) if got[i].tobytes() != want[i]]
                    assert not bad, f"GPU level-{level} frames of the {name} corpus differ from the oracle at {bad[:5]}"
        parity = {"frames_checked": 4 * m_chk, "against": "oracle (C restatement of the reference, pinned to its libzstd.dll 1.5.1)", "levels": [1, 3], "mismatches": 0}
    csz_u = np.array([f.size for f in frames_u], dtype=np.int64)
    # decode batch for this rank: compressed frames back to back (contiguous => one DMA in the e2e path)
    csz = csz_u[my_ids]
    coff = np.concatenate([[0], np.cumsum(csz)[:-1]]).astype(np.int64)
    ctotal = int(csz.sum())
    h_comp = torch.empty(ctotal + 64, dtype=torch.uint8).pin_memory()
    hc = h_comp.numpy()
    uoff = np.concatenate([[0], np.cumsum(csz_u)[:-1]])
    ublob = np.concatenate(frames_u)
    for k in range(0, n, uniq):                               # tile the unique blob
        m = min(uniq, n - k)
        if np.array_equal(my_ids[k:k + m], np.arange(m)):
            e = int(uoff[m - 1] + csz_u[m - 1]); hc[int(coff[k]):int(coff[k]) + e] = ublob[:e]
        else:
            for j in range(m):
                hc[int(coff[k + j]):int(coff[k + j]) + int(csz[k + j])] = frames_u[int(my_ids[k + j])]
    d_comp = h_comp.cuda()
    d_out = torch.empty(n * FRAME, dtype=torch.uint8, device="cuda")
    log(f"[rank {rank}] prepared {n} frames, {ctotal / 1e6:.1f} MB compressed (ratio {n * FRAME / ctotal:.2f}) in {time.perf_counter() - t_prep:.1f}s")

    so = (u64 * n)(*coff.tolist()); ss = (st * n)(*csz.tolist())
    do = (u64 * n)(*[i * FRAME for i in range(n)]); dc = (st * n)(*([FRAME] * n)); res = (st * n)()

    def dec_step():
        dev_call(lib.ZSTDB200_decompressBatchDevice, dec, n, d_comp.data_ptr(), so, ss, d_out.data_ptr(), do, dc, res)
        return dec.launch_count(), np.array(dec.timings())

    sampler.start()
    dev_ms, wall_ms, launches, slots = timed_steps(dec_step, args.warmup, args.steps)
    clocks = sampler.stop()
    assert all(r == FRAME for r in res), "a frame failed to decode"
    # size-independent property at full size: every output frame equals the chunk it was made from
    ref = d_text_u.view(uniq, FRAME)[torch.from_numpy(my_ids).cuda()]
    assert torch.equal(d_out.view(n, FRAME), ref), "decoded batch differs from the source corpus"
    del ref
    total_u = sum_over_ranks(float(n * FRAME))
    total_c = sum_over_ranks(float(ctotal))
    dec_gbps = total_u * args.steps / (dev_ms * 1e-3) / 1e9

    # dominant kernel and its roofline (algorithmic bytes per launch = sum(U_i + C_i))
    names = {3: "dec_scan", 4: "dec_setup", 5: "dec_huf", 6: "dec_seq", 7: "dec_exec"}
    kern_ms = {names[k]: float(slots[k]) for k in names}
    top = max(kern_ms, key=kern_ms.get)
    peak, peak_src = measured_hbm_peak()
    algo_bytes = n * FRAME + ctotal
    achieved = algo_bytes / (kern_ms[top] * 1e-3) / 1e9 if kern_ms[top] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(top)
        except Exception:
            traffic = None
    roofline = {"kernel": top, "bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": {k: round(v, 4) for k, v in kern_ms.items()},
                "pipeline_frac": round((algo_bytes * world / (dev_ms * 1e-3 / args.steps) / 1e9) / (peak * world), 5)}

    # ------------------------------------------------------------------ e2e decompress: host-pointer C ABI, pinned buffers
    h_out = torch.empty(n * FRAME, dtype=torch.uint8).pin_memory()
    vp = ctypes.c_void_p
    base_c, base_o = h_comp.data_ptr(), h_out.data_ptr()
    sp = (vp * n)(*[base_c + int(o) for o in coff]); dp = (vp * n)(*[base_o + i * FRAME for i in range(n)])

    def e2e_dec_step():
        dev_call(lib.ZSTDB200_decompressBatch, dec, n, sp, ss, dp, dc, res)
        return dec.launch_count(), np.array(dec.timings())
    e2e_steps = max(1, min(args.steps, 5))
    _, e2e_wall_ms, _, e2e_slots = timed_steps(e2e_dec_step, min(args.warmup, 2), e2e_steps)
    assert all(r == FRAME for r in res)
    want_out = text.reshape(uniq, FRAME)[my_ids]
    assert np.array_equal(h_out.numpy().reshape(n, FRAME), want_out), "e2e output differs from the source corpus"      # the whole batch
    e2e_gbps = total_u * e2e_steps / (e2e_wall_ms * 1e-3) / 1e9
    e2e = {"value": round(e2e_gbps, 3), "unit": "GB/s", "h2d_bytes_per_step": int(total_c), "d2h_bytes_per_step": int(total_u),
           "ms_per_step": round(e2e_wall_ms / e2e_steps, 3), "api": "ZSTDB200_decompressBatch (host pointers, pinned, contiguous)",
           "phase_ms": {"h2d": round(float(e2e_slots[0]), 3), "kernels": round(float(e2e_slots[1]), 3), "d2h": round(float(e2e_slots[2]), 3)}}
    # ---- what the copies alone cost (untimed probe, all ranks at once): the same pinned buffers, H2D of the compressed batch and
    # D2H of the regenerated bytes issued together on two streams (PCIe is full duplex), no kernels.  e2e cannot beat this.
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_h2d):
            d_comp.copy_(h_comp, non_blocking=True)
        with torch.cuda.stream(s_d2h):
            h_out.copy_(d_out, non_blocking=True)
        s_h2d.synchronize(); s_d2h.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) * 1e3)
        best = dt if best is None else min(best, dt)
    one_way = None
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_d2h):
            h_out.copy_(d_out, non_blocking=True)
        s_d2h.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) * 1e3)
        one_way = dt if one_way is None else min(one_way, dt)
    e2e["copy_ceiling_gbs"] = round(total_u / (best * 1e-3) / 1e9, 3)
    e2e["copy_ceiling_ms"] = round(best, 3)
    e2e["d2h_only_ms"] = round(one_way, 3)
    e2e["frac_of_ceiling"] = round(e2e_gbps / (total_u / (best * 1e-3) / 1e9), 4)
    e2e["numa_bind_rc"] = numa_rc
    del h_out

    # ---- the drop-in callers' memory: pageable, one separately allocated array per frame (Decompressor.Unwrap: `fixed` over a
    # managed byte[] is GC pinning, not page locking; every Unwrap returns a `new byte[]`)
    pg_src = [np.array(frames_u[int(i)]) for i in my_ids]
    pg_dst = [np.empty(FRAME, dtype=np.uint8) for _ in range(n)]
    for a in pg_dst:
        a[::4096] = 0
    psp = (vp * n)(*[a.ctypes.data for a in pg_src]); pdp = (vp * n)(*[a.ctypes.data for a in pg_dst])

    def e2e_pageable_step():
        dev_call(lib.ZSTDB200_decompressBatch, dec, n, psp, ss, pdp, dc, res)
        return dec.launch_count(), np.array(dec.timings())
    _, pg_wall_ms, _, _ = timed_steps(e2e_pageable_step, 1, 2)
    assert all(r == FRAME for r in res)
    for i in range(0, n, max(1, n // 61)):
        assert np.array_equal(pg_dst[i], want_out[i]), "pageable e2e output differs"
    e2e_pageable = {"value": round(total_u * 2 / (pg_wall_ms * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(pg_wall_ms / 2, 3),
                    "buffers": f"{n} separately allocated pageable arrays in, {n} out (numpy malloc), staged through the library's pinned ring by host threads",
                    "frac_of_pinned": round((total_u * 2 / (pg_wall_ms * 1e-3) / 1e9) / e2e_gbps, 4)}
    del pg_dst, pg_src

    # ---- single-call latency of the zstd-named entry points on ONE 128 KiB frame (median), pageable buffers
    single = None
    if rank == 0:
        f0 = np.array(frames_u[0]); o0 = np.empty(FRAME, dtype=np.uint8); c0 = np.empty(bound, dtype=np.uint8)
        chunk0 = np.ascontiguousarray(text[:FRAME])
        lat_d, lat_c = [], []
        for k in range(60):
            t0 = time.perf_counter()
            r1 = lib.ZSTD_decompressDCtx(dec.handle, o0.ctypes.data, FRAME, f0.ctypes.data, f0.size)
            lat_d.append(time.perf_counter() - t0)
            assert r1 == FRAME
        for k in range(30):
            t0 = time.perf_counter()
            r2 = lib.ZSTD_compressCCtx(comp.handle, c0.ctypes.data, bound, chunk0.ctypes.data, FRAME, 1)
            lat_c.append(time.perf_counter() - t0)
            assert r2 == f0.size
        assert np.array_equal(o0, chunk0) and np.array_equal(c0[:r2], f0)
        single = {"ZSTD_decompressDCtx_ms": round(1e3 * float(np.median(lat_d[5:])), 3), "ZSTD_compressCCtx_level1_ms": round(1e3 * float(np.median(lat_c[3:])), 3),
                  "note": "one 128 KiB text-like frame per call, host to host; a single frame is a serial chain on the GPU (one lane per bit stream)"}

    # ------------------------------------------------------------------ level-1 compress (configs[2])
    compress = None
    if not args.skip_compress:
        d_sil = d_sil_u.view(uniq, FRAME)[torch.from_numpy(my_ids).cuda()].contiguous().view(-1)
        d_cout = torch.empty(n * slot, dtype=torch.uint8, device="cuda")
        cso = (u64 * n)(*[i * FRAME for i in range(n)]); css = (st * n)(*([FRAME] * n))
        cdo = (u64 * n)(*[i * slot for i in range(n)]); cdc = (st * n)(*([bound] * n)); cres = (st * n)()

        def comp_step():
            dev_call(lib.ZSTDB200_compressBatchDevice, comp, n, 1, d_sil.data_ptr(), cso, css, d_cout.data_ptr(), cdo, cdc, cres)
            return comp.launch_count(), np.array(comp.timings())
        csteps = max(1, min(args.steps, 3))
        c_dev_ms, _, c_launch, c_slots = timed_steps(comp_step, min(args.warmup, 1), csteps)
        csizes = np.array(list(cres), dtype=np.int64)
        assert (csizes > 0).all() and (csizes <= bound).all()
        # round trip of the compressed batch through the GPU decoder (encode -> decode property at full size)
        rdo = (u64 * n)(*[i * FRAME for i in range(n)])
        dev_call(lib.ZSTDB200_decompressBatchDevice, dec, n, d_cout.data_ptr(), cdo, (st * n)(*csizes.tolist()), d_out.data_ptr(), rdo, dc, res)
        assert all(r == FRAME for r in res) and torch.equal(d_out, d_sil), "compress -> decompress round trip failed"
        c_total_c = sum_over_ranks(float(csizes.sum()))
        c_gbps = total_u * csteps / (c_dev_ms * 1e-3) / 1e9
        c_algo = n * FRAME + int(csizes.sum())
        compress = {"value": round(c_gbps, 3), "unit": "GB/s", "ms_per_step": round(c_dev_ms / csteps, 3), "steps": csteps,
                    "workload": "batch compress 1 GiB Silesia-mix-like in 128 KiB chunks, level 1 (configs[2]), byte-identical per chunk (tests/test_encode_gpu.py)",
                    "ratio": round(n * FRAME * world / c_total_c, 4),
                    "kernel_ms": {"enc_match": round(float(c_slots[8]), 3), "enc_entropy": round(float(c_slots[9]), 3)},
                    "roofline_frac_top_kernel": round(c_algo / (max(float(c_slots[8]), float(c_slots[9])) * 1e-3) / 1e9 / peak, 5),
                    "gpu_launches": int(c_launch)}
        # e2e compress through the host-pointer ABI
        h_in = torch.from_numpy(np.ascontiguousarray(sil)).pin_memory() if n <= uniq else d_sil.cpu().pin_memory()
        h_cout = torch.empty(n * slot, dtype=torch.uint8).pin_memory()
        csp = (vp * n)(*[h_in.data_ptr() + i * FRAME for i in range(n)]); cdp = (vp * n)(*[h_cout.data_ptr() + i * slot for i in range(n)])

        def e2e_comp_step():
            dev_call(lib.ZSTDB200_compressBatch, comp, n, 1, csp, css, cdp, cdc, cres)
            return comp.launch_count(), np.array(comp.timings())
        _, ce_wall_ms, _, _ = timed_steps(e2e_comp_step, 1, 2)
        # the host-pointer path must hand back exactly the frames of the device-resident pass
        e2e_sizes = np.array(list(cres), dtype=np.int64)
        assert np.array_equal(e2e_sizes, csizes), "host-pointer compress sizes differ from the device-resident pass"
        dcheck = d_cout.cpu().numpy()
        hcheck = h_cout.numpy()
        for i in range(0, n, max(1, n // 97)):
            assert np.array_equal(hcheck[i * slot:i * slot + csizes[i]], dcheck[i * slot:i * slot + csizes[i]]), "host-pointer compress bytes differ"
        compress["e2e"] = {"value": round(total_u * 2 / (ce_wall_ms * 1e-3) / 1e9, 3), "unit": "GB/s",
                           "h2d_bytes_per_step": int(total_u), "d2h_bytes_per_step": int(c_total_c)}
        del h_in, h_cout

        # -------------------------------------------------------------- level 3 (ZSTD_dfast) compress + decompress (configs[4])
        # 1 GiB of the same Silesia-mix-like chunks per GPU (8 GiB at 8 GPUs), device-resident; the frames are byte-identical to the
        # oracle's level-3 frames (tests/test_encode_gpu.py), so the ratio IS the reference's ratio.
        def comp3_step():
            dev_call(lib.ZSTDB200_compressBatchDevice, comp, n, 3, d_sil.data_ptr(), cso, css, d_cout.data_ptr(), cdo, cdc, cres)
            return comp.launch_count(), np.array(comp.timings())
        c3_dev_ms, _, c3_launch, c3_slots = timed_steps(comp3_step, 1, 2)
        c3sizes = np.array(list(cres), dtype=np.int64)
        assert (c3sizes > 0).all() and (c3sizes <= bound).all()
        c3ss = (st * n)(*c3sizes.tolist())

        def dec3_step():
            dev_call(lib.ZSTDB200_decompressBatchDevice, dec, n, d_cout.data_ptr(), cdo, c3ss, d_out.data_ptr(), rdo, dc, res)
            return dec.launch_count(), np.array(dec.timings())
        d3_dev_ms, _, _, _ = timed_steps(dec3_step, 1, 3)
        assert all(r == FRAME for r in res) and torch.equal(d_out, d_sil), "level-3 compress -> decompress round trip failed"
        c3_total_c = sum_over_ranks(float(c3sizes.sum()))
        compress["level3"] = {
            "workload": "level 3 (ZSTD_dfast) compress + decompress of 1 GiB Silesia-mix-like per GPU in 128 KiB chunks (configs[4]: 8 GiB at 8 GPUs), device-resident",
            "compress": {"value": round(total_u * 2 / (c3_dev_ms * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(c3_dev_ms / 2, 3),
                         "kernel_ms": {"enc_match": round(float(c3_slots[8]), 3), "enc_entropy": round(float(c3_slots[9]), 3)}, "gpu_launches": int(c3_launch)},
            "decompress": {"value": round(total_u * 3 / (d3_dev_ms * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(d3_dev_ms / 3, 3)},
            "ratio": round(n * FRAME * world / c3_total_c, 4)}

    # ------------------------------------------------------------------ other decode workloads of BASELINE.json (device-resident)
    workloads = None
    if not args.skip_workloads:
        workloads = {}
        wu = min(uniq, 512)
        for name in ("silesia", "literal_mix", "literal_heavy", "incompressible"):
            host_u = sil[:wu * FRAME] if name == "silesia" else dg.WORKLOADS[name](wu * FRAME)
            fr = compress_unique(torch.from_numpy(host_u).cuda(), 1)
            wsz_u = np.array([f.size for f in fr], dtype=np.int64)
            ids = np.arange(n) % wu
            wsz = wsz_u[ids]; woff = np.concatenate([[0], np.cumsum(wsz)[:-1]]).astype(np.int64)
            ub = np.concatenate(fr)
            blob = np.concatenate([ub] * (n // wu) + ([np.concatenate(fr[:n % wu])] if n % wu else []))
            d_w = torch.from_numpy(np.concatenate([blob, np.zeros(64, dtype=np.uint8)])).cuda()
            wso = (u64 * n)(*woff.tolist()); wss = (st * n)(*wsz.tolist())

            def w_step():
                dev_call(lib.ZSTDB200_decompressBatchDevice, dec, n, d_w.data_ptr(), wso, wss, d_out.data_ptr(), do, dc, res)
                return dec.launch_count(), np.array(dec.timings())
            w_ms, _, _, w_slots = timed_steps(w_step, 1, 3)
            assert all(r == FRAME for r in res)
            refw = torch.from_numpy(host_u).cuda().view(wu, FRAME)
            assert torch.equal(d_out.view(n, FRAME)[:wu], refw) and torch.equal(d_out.view(n, FRAME)[n - wu:], refw[torch.from_numpy((np.arange(n - wu, n) % wu)).cuda()])
            wc = float(wsz.sum())
            workloads[name] = {"decompress_GBps": round(total_u * 3 / (w_ms * 1e-3) / 1e9, 2), "ms_per_step": round(w_ms / 3, 3), "ratio": round(n * FRAME / wc, 3),
                               "hbm_algorithmic_GBps_per_gpu": round((n * FRAME + wc) * 3 / (w_ms * 1e-3) / 1e9 / world, 1),
                               "kernel_ms": {names[k]: round(float(w_slots[k]), 3) for k in names}}
            del d_w, refw
        # ---- dictionary records: what Compressor.LoadDictionary is for -- many small records sharing one dictionary ----
        # 65536 text-like records of 4 KiB, a 32 KiB raw-content dictionary cut from the same generator, device-resident compression at
        # levels 1 and 3 with and without the dictionary; the frames with the dictionary are decoded again through the host API
        # (dictionary decoding lays its output out itself).  The frames are the reference's bytes (tests/test_encode_gpu.py, soak).
        try:
            rec, nrec = 4096, 65536
            dtext = dg.text_like((nrec + 64) * rec)
            dictionary = np.ascontiguousarray(dtext[:32768])
            d_rec = torch.from_numpy(np.concatenate([dtext[64 * rec:(64 + nrec) * rec], np.zeros(64, dtype=np.uint8)])).cuda()
            rslot = (int(lib.ZSTD_compressBound(rec)) + 15) & ~15
            d_rout = torch.empty(nrec * rslot + 64, dtype=torch.uint8, device="cuda")
            rso = (u64 * nrec)(*[i * rec for i in range(nrec)]); rss = (st * nrec)(*([rec] * nrec))
            rdo2 = (u64 * nrec)(*[i * rslot for i in range(nrec)]); rdc2 = (st * nrec)(*([rslot] * nrec)); rres = (st * nrec)()
            drec = {"workload": f"{nrec} text-like records of {rec} B per GPU, 32 KiB raw-content dictionary (ZSTD_CCtx_loadDictionary), device-resident"}
            for lvl in (1, 3):
                for use_dict in (False, True):
                    cd = api.Compressor(lvl)
                    if use_dict:
                        cd.LoadDictionary(dictionary)

                    def r_step():
                        dev_call(lib.ZSTDB200_compressBatchDevice, cd, nrec, lvl, d_rec.data_ptr(), rso, rss, d_rout.data_ptr(), rdo2, rdc2, rres)
                        return cd.launch_count(), np.array(cd.timings())
                    r_ms, _, _, r_slots = timed_steps(r_step, 1, 2)
                    rsz = np.array(list(rres), dtype=np.int64)
                    assert (rsz > 0).all() and (rsz <= rslot).all()
                    drec[f"level{lvl}_{'dict' if use_dict else 'nodict'}"] = {
                        "compress_GBps": round(nrec * rec * world * 2 / (r_ms * 1e-3) / 1e9, 3), "ms_per_step": round(r_ms / 2, 3),
                        "ratio": round(nrec * rec / float(rsz.sum()), 3),
                        "kernel_ms": {"enc_match": round(float(r_slots[8]), 3), "enc_entropy": round(float(r_slots[9]), 3)}}
                    if use_dict and lvl == 1:              # round trip of a sample through the decoder with the same dictionary
                        hb = d_rout.cpu().numpy()
                        fr = [hb[i * rslot:i * rslot + rsz[i]].tobytes() for i in range(0, nrec, 64)]
                        dd = api.Decompressor(); dd.LoadDictionary(dictionary)
                        back = dd.UnwrapBatch(fr); dd.Dispose()
                        hsrc = d_rec.cpu().numpy()
                        assert all(b == hsrc[i * rec:(i + 1) * rec].tobytes() for b, i in zip(back, range(0, nrec, 64))), "dictionary round trip failed"
                    cd.Dispose()
            workloads["dictionary_records"] = drec
            del d_rec, d_rout
        except Exception as e:                                   # noqa: BLE001  (a probe must not take the headline down with it)
            workloads["dictionary_records"] = {"error": repr(e)}
        workloads["note"] = "8192 x 128 KiB level-1 frames per GPU made by the GPU encoder from zstdsharp_b200.datagen workloads; incompressible = raw blocks (copy: 2 x bytes of HBM traffic)"

    # ------------------------------------------------------------------ CPU baseline (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        o, chunks, frames, threads = cpu_frames(1, 256, "text")
        cpu_decode_pass(o, frames, threads)
        b, t = cpu_decode_pass(o, frames, threads, seconds_min=1.0)
        port_rate = b / t / 1e9
        r = ref_dll()
        if r is not None:
            srcs = [np.frombuffer(frames[i % 256], dtype=np.uint8) for i in range(2048)]
            job = DllBatch(r, srcs, FRAME, threads)
            job.decompress()
            t0 = time.perf_counter(); reps = 0
            while time.perf_counter() - t0 < 2.0:
                job.decompress(); reps += 1
            t = time.perf_counter() - t0
            assert all(x == FRAME for x in job.res)
            cpu = {"value": round(2048 * FRAME * reps / t / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": "reference",
                   "sample": f"2048 text-like 128 KiB level-1 frames (256 MiB out) decoded {reps}x in {t:.1f} s by the reference's own libzstd.dll (zstd 1.5.1, oracle/_ref) on {threads} threads",
                   "oracle_port_GBps": round(port_rate, 4)}
            sil_c = dg.silesia_mix(256 * FRAME).reshape(256, FRAME)
            cj = DllBatch(r, [sil_c[i % 256] for i in range(1024)], r.lib.ZREF_compressBound(FRAME), threads)
            cj.compress(1); t1 = time.perf_counter(); cj.compress(1)
            cpu["compress_l1_GBps"] = round(1024 * FRAME / (time.perf_counter() - t1) / 1e9, 4)
            del job, cj
        else:
            cpu = {"value": round(port_rate, 4), "unit": "GB/s", "cores": threads, "kind": "port",
                   "sample": f"256 text-like 128 KiB level-1 frames decoded repeatedly for {t:.1f} s on {threads} threads (oracle port of the reference's C# code)"}
            cb, ct = cpu_compress_pass(o, [c for c in dg.silesia_mix(128 * FRAME).reshape(-1, FRAME)], threads, 1)
            cpu["compress_l1_GBps"] = round(cb / ct / 1e9, 4)
        cpu["native_libzstd_1_5_5_GBps"] = native_libzstd_decode_rate(frames, threads)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(dec_gbps, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dev_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "batch decompress 1 GiB of 128 KiB level-1 frames, bit-exact (BASELINE.json configs[1]); value = decompress, compress_l1 = configs[2]",
                       "frames_per_gpu": nframes, "frame_bytes": FRAME, "corpus": f"text_like seed 0x{dg.SEED_TEXT:X} (+512 per 64 MiB block): {uniq} unique frames tiled to {nframes}",
                       "compressed_bytes_per_gpu": ctotal, "parallelism": f"host scatter, {world} rank(s), no collective",
                       "l2": "inputs larger than L2 (compressed batch %.0f MB + 1 GiB output per step)" % (ctotal / 1e6)},
            "wall_ms_per_step": round(wall_ms / args.steps, 4),
            "clocks": clocks, "e2e": e2e, "e2e_pageable": e2e_pageable, "single_call": single, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "compress_l1": compress, "workloads": workloads, "parity_gate": parity,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=8192, help="frames per GPU (8192 x 128 KiB = 1 GiB)")
    ap.add_argument("--skip-compress", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-workloads", action="store_true")
    ap.add_argument("--unique", type=int, default=UNIQUE_FRAMES, help="unique frames in the corpus (tiled up to --frames)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
