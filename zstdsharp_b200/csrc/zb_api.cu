// zb_api.cu -- C ABI (include/zstd_b200.h) + batch frame scheduler of libzstdb200.so (product code).
//
// Host side of the drop-in boundary: contexts own a CUDA stream and a grow-only device arena; the batch entry
// points move frames host<->device with as few DMA operations as the caller's layout allows and run the decode
// (zb_decode.cu) / encode (zb_encode.cu) kernel pipelines.  There is no CPU fallback anywhere in this file:
// without a device every call returns ZSTD_error_GENERIC.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <cctype>
#include <cmath>
#include <sched.h>

#include "../../include/zstd_b200.h"
#include "zb_decode.cuh"
#include "zb_encode.cuh"

namespace zb {

void host_copy(void* dst, const void* src, size_t n);      // zb_hostcopy.cpp: staging copy with non-temporal stores

static thread_local std::string t_lastError;
static void set_error(const std::string& s) { t_lastError = s; }

#define ZB_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                             \
            fprintf(stderr, "[zstdb200] CUDA failure: %s: %s\n", #call, cudaGetErrorString(e_));      \
            return false;                                                                              \
        }                                                                                              \
    } while (0)

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    bool ensure(size_t n) {
        if (n <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { want = n + 256; e = cudaMalloc(&p, want); }
        if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); p = nullptr; return false; }
        cap = want; return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return (T*)p; }
};
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    bool ensure(size_t n) {
        if (n <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t const want = n + n / 4 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) { set_error(std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); p = nullptr; return false; }
        cap = want; return true;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return (T*)p; }
};

constexpr size_t kMaxItemsPerPass = 8192;
constexpr int kEvents = 24;
constexpr int kPipeMax = 8;                    // upper bound of sub-batches in flight in the host-pointer path (one scratch arena + compute stream each)
static int env_int(const char* name, int def, int lo, int hi) { const char* e = getenv(name); if (!e) return def; int v = atoi(e); return v < lo ? lo : (v > hi ? hi : v); }
// sub-batches in flight / target sub-batch size (items); tunable for experiments, defaults chosen on B200 (profiles/r01_notes.md)
static int pipe_depth() { static int v = env_int("ZSTDB200_PIPE", 4, 1, kPipeMax); return v; }      // 3 -> 4 with first sub-batch n/32: 24.5 -> 23.8 ms per GiB host to host (profiles/r02_notes.md)
static size_t pipe_items() { static int v = env_int("ZSTDB200_PIPE_ITEMS", 4096, 16, 8192); return (size_t)v; }

// scratch of one decode pass (DecPass layout, zb_decode.cuh)
struct DecArena {
    DevBuf dItems, dInit, dHuf, dFse, dLit, dSeq, dHufList, dSeqList, dRawList, dCounters, dResults;
    PinBuf hInit, hCounters, hResults;
    bool ensure(size_t m) {
        return dItems.ensure(m * sizeof(DecItem)) && dInit.ensure(m * sizeof(DecItemInit)) && hInit.ensure(m * sizeof(DecItemInit)) &&
               dHuf.ensure(m * kHufTableEntries * 2) && dFse.ensure(m * kFseTableEntries * 4) && dLit.ensure(m * (size_t)kLitStride) &&
               dSeq.ensure(m * (size_t)kSeqCap * 8) && dHufList.ensure(m * 4) && dSeqList.ensure(m * 4) && dRawList.ensure(m * 4) &&
               dCounters.ensure(64) && hCounters.ensure(64) && dResults.ensure(m * 8) && hResults.ensure(m * 8);
    }
    void release() {
        DevBuf* d[] = {&dItems, &dInit, &dHuf, &dFse, &dLit, &dSeq, &dHufList, &dSeqList, &dRawList, &dCounters, &dResults};
        for (auto* b : d) b->release();
        PinBuf* h[] = {&hInit, &hCounters, &hResults};
        for (auto* b : h) b->release();
    }
};

struct Engine {
    int device = -1;
    int wantDevice = -1;                // >= 0: the device this context must live on (multi-device scheduler); else ZSTDB200_DEVICE / the caller's current device
    bool ready = false;
    cudaStream_t stream = nullptr;      // stream all work of this context is issued on
    cudaStream_t ownStream = nullptr;   // created by the context; replaced by ZSTDB200_setStream
    cudaEvent_t ev[kEvents] = {};
    // decode arenas ([0] also serves the device-pointer API) and the copy/compute streams of the host-pointer pipeline
    DecArena dec[kPipeMax];
    DevBuf dDefaultFse;
    bool defaultTablesBuilt = false;
    cudaStream_t sIn = nullptr, sOut = nullptr, sComp[kPipeMax] = {};
    cudaStream_t sEnc[kPipeMax] = {};   // compress pipeline: sub-batch k on sEnc[k], earlier sub-batches at higher priority (their entropy stage goes first)
    cudaEvent_t evStart = nullptr, evIn[2] = {}, evOut[2] = {};
    std::vector<cudaEvent_t> evPool;
    // host<->device staging for the host-pointer API
    DevBuf dSrc, dDst; PinBuf hStage, hStageOut;
    // encode arena
    EncArena enc;                       // device-pointer API and small host batches
    EncArena encPipe[kPipeMax];         // one per sub-batch in flight in the pipelined host path
    DevBuf dEncInit; PinBuf hEncInit;
    // dictionary of a decode context (ZSTD_DCtx_loadDictionary): raw bytes, parsed tables, info words (zb_decode.cuh)
    DevBuf dDict, dDictHuf, dDictFse, dDictInfo;
    uint32_t dictInfo[kDictInfoWords] = {};
    bool dictLoaded = false;
    size_t dict_headroom() const { return dictLoaded ? (((size_t)dictInfo[11] + 127) & ~(size_t)127) : 0; }
    // instrumentation
    float timings[ZSTDB200_TIMING_SLOTS] = {};
    unsigned launches = 0;

    bool init() {
        if (ready) return true;
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) { set_error(std::string("no CUDA device: ") + cudaGetErrorString(e)); return false; }
        int dev = 0;
        if (wantDevice >= 0) dev = wantDevice;
        else if (const char* env = getenv("ZSTDB200_DEVICE")) dev = atoi(env);
        else if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        if (dev < 0 || dev >= count) { set_error("ZSTDB200_DEVICE out of range"); return false; }
        device = dev;
        ZB_CUDA(cudaSetDevice(device));
        ZB_CUDA(cudaStreamCreateWithFlags(&ownStream, cudaStreamNonBlocking));
        if (!stream) stream = ownStream;
        for (int i = 0; i < kEvents; i++) ZB_CUDA(cudaEventCreate(&ev[i]));
        ZB_CUDA(cudaStreamCreateWithFlags(&sIn, cudaStreamNonBlocking));
        ZB_CUDA(cudaStreamCreateWithFlags(&sOut, cudaStreamNonBlocking));
        for (int i = 0; i < kPipeMax; i++) ZB_CUDA(cudaStreamCreateWithFlags(&sComp[i], cudaStreamNonBlocking));
        {
            int least = 0, greatest = 0;
            ZB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));          // numerically lower = more urgent
            for (int i = 0; i < kPipeMax; i++) ZB_CUDA(cudaStreamCreateWithPriority(&sEnc[i], cudaStreamNonBlocking, std::min(least, greatest + i)));
        }
        ZB_CUDA(cudaEventCreate(&evStart));
        for (int i = 0; i < 2; i++) { ZB_CUDA(cudaEventCreate(&evIn[i])); ZB_CUDA(cudaEventCreate(&evOut[i])); }
        ready = true;
        return true;
    }
    void destroy() {
        if (device >= 0) cudaSetDevice(device);
        DevBuf* d[] = {&dDefaultFse, &dSrc, &dDst, &dEncInit};
        for (auto* b : d) b->release();
        for (auto& a : dec) a.release();
        PinBuf* h[] = {&hStage, &hStageOut, &hEncInit};
        for (auto* b : h) b->release();
        enc.release();
        for (auto& a : encPipe) a.release();
        if (ready) {
            for (int i = 0; i < kEvents; i++) cudaEventDestroy(ev[i]);
            for (auto e : evPool) cudaEventDestroy(e);
            evPool.clear();
            cudaEventDestroy(evStart);
            for (int i = 0; i < 2; i++) { cudaEventDestroy(evIn[i]); cudaEventDestroy(evOut[i]); }
            cudaStreamDestroy(sIn); cudaStreamDestroy(sOut);
            for (int i = 0; i < kPipeMax; i++) { cudaStreamDestroy(sComp[i]); cudaStreamDestroy(sEnc[i]); }
            cudaStreamDestroy(ownStream);
        }
        ready = false;
    }
    bool bind() { ZB_CUDA(cudaSetDevice(device)); return true; }
    bool need_events(size_t n) { while (evPool.size() < n) { cudaEvent_t e; ZB_CUDA(cudaEventCreate(&e)); evPool.push_back(e); } return true; }
};

// ---------------------------------------------------------------------------------------------------------------
//  Decode pass over device-resident frames
// ---------------------------------------------------------------------------------------------------------------
// Enqueues one pass (m <= kMaxItemsPerPass items) on `stream` using arena A.  nWaves == 0: the block count is read back
// from the scan kernel (one host sync); otherwise the caller already knows it (host-pointer path) and nothing blocks.
// The per-item results land in A.hResults once the stream has drained.
static bool decode_enqueue(Engine& E, DecArena& A, cudaStream_t stream, size_t m, const uint8_t* d_src, const uint64_t* srcOff, const size_t* srcSize,
                           uint8_t* d_dst, const uint64_t* dstOff, const size_t* dstCap, uint32_t nWaves, cudaEvent_t* timeEv /* [7] or null */)
{
    if (!A.ensure(m) || !E.dDefaultFse.ensure(kFseTableEntries * 4)) return false;
    if (!E.defaultTablesBuilt) {
        dec_build_default_tables(E.dDefaultFse.as<uint32_t>(), E.stream); E.launches++; E.defaultTablesBuilt = true;
        ZB_CUDA(cudaStreamSynchronize(E.stream));       // built once per context; every stream may use it afterwards
    }
    DecItemInit* hi = A.hInit.as<DecItemInit>();
    // oversized items cannot be addressed by the 32-bit cursors: they are reported per item by decode_collect
    for (size_t i = 0; i < m; i++) {
        hi[i].srcOff = srcOff[i]; hi[i].dstOff = dstOff[i];
        hi[i].srcSize = (uint32_t)std::min<size_t>(srcSize[i], 0xFFFFFFF0u);
        hi[i].dstCap = (uint32_t)std::min<size_t>(dstCap[i], 0xFFFFFFF0u);
    }
    // The item descriptors and the per-item results stay in pinned host memory and are read / written by the kernels
    // directly (zero copy): a tiny cudaMemcpyAsync would queue behind the bulk transfers of the other sub-batches on
    // the copy engines and stall this sub-batch's kernels for milliseconds.
    ZB_CUDA(cudaMemsetAsync(A.dCounters.p, 0, 64, stream));
    DecPass p;
    p.items = A.dItems.as<DecItem>(); p.nItems = (uint32_t)m; p.src = d_src; p.dst = d_dst;
    p.hufTable = A.dHuf.as<uint16_t>(); p.fseTable = A.dFse.as<uint32_t>(); p.litBuf = A.dLit.as<uint8_t>();
    p.seq = A.dSeq.as<uint2>();
    p.defaultFse = E.dDefaultFse.as<uint32_t>(); p.hufList = A.dHufList.as<uint32_t>(); p.seqList = A.dSeqList.as<uint32_t>(); p.rawList = A.dRawList.as<uint32_t>();
    p.counters = A.dCounters.as<uint32_t>(); p.results = A.hResults.as<uint64_t>();
    p.dictHuf = nullptr; p.dictFse = nullptr; p.dictBytes = nullptr; p.dictInfo = nullptr;
    if (E.dictLoaded) { p.dictHuf = E.dDictHuf.as<uint16_t>(); p.dictFse = E.dDictFse.as<uint32_t>(); p.dictBytes = E.dDict.as<uint8_t>(); p.dictInfo = E.dDictInfo.as<uint32_t>(); }
    if (timeEv) ZB_CUDA(cudaEventRecord(timeEv[0], stream));
    dec_launch_scan_init(p, hi, stream); E.launches++;
    if (E.dictLoaded && E.dictInfo[11]) { dec_launch_dict_prefill(p, E.dictInfo[10], E.dictInfo[11], stream); E.launches++; }
    if (nWaves == 0) {
        ZB_CUDA(cudaMemcpyAsync(A.hCounters.p, A.dCounters.p, 16, cudaMemcpyDeviceToHost, stream));
        ZB_CUDA(cudaStreamSynchronize(stream));
        nWaves = A.hCounters.as<uint32_t>()[3];
    }
    for (uint32_t w = 0; w < nWaves; w++) {
        if (timeEv && w == 0) dec_launch_wave_timed(p, stream, &timeEv[2]);   // records timeEv[2..6] between the kernels of the first wave
        else dec_launch_wave(p, stream);
        E.launches += 6;
    }
    dec_launch_finish(p, stream); E.launches++;
    if (timeEv) ZB_CUDA(cudaEventRecord(timeEv[1], stream));
    return true;
}

static void decode_collect(const DecArena& A, size_t m, const size_t* srcSize, size_t* result)
{
    const uint64_t* hr = A.hResults.as<uint64_t>();
    for (size_t i = 0; i < m; i++) result[i] = srcSize[i] > 0xFFFFFFF0u ? (size_t)make_error(kSrcSizeWrong) : (size_t)hr[i];
}

static bool decode_device(Engine& E, size_t n, const uint8_t* d_src, const uint64_t* srcOff, const size_t* srcSize,
                          uint8_t* d_dst, const uint64_t* dstOff, const size_t* dstCap, size_t* result, bool timeKernels)
{
    if (!E.bind()) return false;
    float kernelMs = 0; float slotMs[5] = {0, 0, 0, 0, 0};
    DecArena& A = E.dec[0];
    for (size_t base = 0; base < n; base += kMaxItemsPerPass) {
        size_t const m = std::min(kMaxItemsPerPass, n - base);
        if (!decode_enqueue(E, A, E.stream, m, d_src, srcOff + base, srcSize + base, d_dst, dstOff + base, dstCap + base, 0, &E.ev[0])) return false;
        ZB_CUDA(cudaStreamSynchronize(E.stream));
        ZB_CUDA(cudaGetLastError());
        float ms = 0; cudaEventElapsedTime(&ms, E.ev[0], E.ev[1]); kernelMs += ms;
        if (timeKernels) {
            float t;
            if (cudaEventElapsedTime(&t, E.ev[0], E.ev[2]) == cudaSuccess) {
                slotMs[0] += t;              // scan (+ sync gap)
                for (int k = 0; k < 4; k++) { cudaEventElapsedTime(&t, E.ev[2 + k], E.ev[3 + k]); slotMs[1 + k] += t; }
            }
        }
        decode_collect(A, m, srcSize + base, result + base);
    }
    (void)cudaGetLastError();
    E.timings[1] = kernelMs;
    for (int k = 0; k < 5; k++) E.timings[3 + k] = slotMs[k];
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
//  Host-pointer staging: contiguous runs go with one DMA each; scattered buffers are gathered through pinned memory
// ---------------------------------------------------------------------------------------------------------------
struct Run { size_t first, count; size_t bytes; };

static void find_runs(std::vector<Run>& runs, size_t n, const void* const* ptr, const size_t* size)
{
    runs.clear();
    size_t i = 0;
    while (i < n) {
        size_t j = i + 1; size_t bytes = size[i];
        while (j < n && (const uint8_t*)ptr[j] == (const uint8_t*)ptr[j - 1] + size[j - 1]) { bytes += size[j]; j++; }
        runs.push_back({i, j - i, bytes});
        i = j;
    }
}

// true when `p` is ordinary pageable host memory (malloc / a managed byte[] that is only GC-pinned): a cudaMemcpyAsync from or to
// it is staged by the driver in small synchronous pieces.  Such buffers go through the library's own pinned staging ring
// instead, filled / drained by host threads while the DMA of the neighbouring sub-batch is in flight.
static bool is_pageable(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
static int host_copy_threads() { static int v = env_int("ZSTDB200_HOST_THREADS", 16, 1, 64); return v; }

static void parallel_for(size_t n, size_t grain, const std::function<void(size_t, size_t)>& fn)
{
    unsigned hw = std::thread::hardware_concurrency(); if (hw == 0) hw = 4;
    size_t const nt = std::min<size_t>(std::min<unsigned>(hw, (unsigned)host_copy_threads()), (n + grain - 1) / grain);
    if (nt <= 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    size_t const per = (n + nt - 1) / nt;
    for (size_t t = 0; t < nt; t++) { size_t a = t * per, b = std::min(n, a + per); if (a < b) th.emplace_back(fn, a, b); }
    for (auto& t : th) t.join();
}

static bool upload_items(Engine& E, size_t n, const void* const* src, const size_t* srcSize, std::vector<uint64_t>& off, size_t* totalOut)
{
    std::vector<Run> runs;
    find_runs(runs, n, src, srcSize);
    off.resize(n);
    // layout: every run starts 16-byte aligned; items inside a run keep their host adjacency
    size_t total = 16;
    for (auto& r : runs) {
        size_t o = total;
        for (size_t k = 0; k < r.count; k++) { off[r.first + k] = o; o += srcSize[r.first + k]; }
        total = (o + 15) & ~(size_t)15;
    }
    total += 16;
    if (!E.dSrc.ensure(total)) return false;
    *totalOut = total;
    if (runs.size() <= 512) {
        for (auto& r : runs)
            if (r.bytes) ZB_CUDA(cudaMemcpyAsync(E.dSrc.as<uint8_t>() + off[r.first], src[r.first], r.bytes, cudaMemcpyHostToDevice, E.stream));
    } else {
        if (!E.hStage.ensure(total)) return false;
        uint8_t* st = E.hStage.as<uint8_t>();
        parallel_for(n, 256, [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) memcpy(st + off[i], src[i], srcSize[i]); });
        ZB_CUDA(cudaMemcpyAsync(E.dSrc.p, st, total, cudaMemcpyHostToDevice, E.stream));
    }
    return true;
}

static size_t layout_dst(size_t n, void* const* dst, const size_t* dstCap, std::vector<uint64_t>& off, size_t* total, size_t headroom)
{
    if (headroom) {         // dictionary decoding: every item gets `headroom` bytes in front of its slot (the dictionary content goes there)
        off.resize(n);
        size_t t = 128;
        for (size_t i = 0; i < n; i++) { off[i] = t + headroom; t = (off[i] + dstCap[i] + 127) & ~(size_t)127; }
        *total = t + 128;
        return n;
    }
    std::vector<Run> runs;
    find_runs(runs, n, (const void* const*)dst, dstCap);
    off.resize(n);
    // every run starts on its own 128-byte line: items of different runs never share a cache line on the device
    size_t t = 128;
    for (auto& r : runs) {
        size_t o = t;
        for (size_t k = 0; k < r.count; k++) { off[r.first + k] = o; o += dstCap[r.first + k]; }
        t = (o + 127) & ~(size_t)127;
    }
    *total = t + 128;
    return runs.size();
}

// copies result[i] bytes of items [first, first+count) back to the caller on stream `st`
static bool download_range(Engine& E, size_t first, size_t count, void* const* dst, const size_t* dstCap, const std::vector<uint64_t>& off,
                           const size_t* result, cudaStream_t st)
{
    auto ok = [&](size_t i) { return !is_error(result[i]); };
    if (E.dictLoaded) {     // items are not adjacent on the device (dictionary headroom between them): one DMA per item
        for (size_t i = first; i < first + count; i++)
            if (ok(i) && result[i]) ZB_CUDA(cudaMemcpyAsync(dst[i], E.dDst.as<uint8_t>() + off[i], result[i], cudaMemcpyDeviceToHost, st));
        return true;
    }
    size_t k = first; size_t const end = first + count;
    while (k < end) {
        // longest run of host-adjacent items that are completely filled, plus one final partial item: one DMA
        size_t j = k; size_t bytes = 0;
        while (j < end && ok(j) && result[j] == dstCap[j] && (j == k || (const uint8_t*)dst[j] == (const uint8_t*)dst[j - 1] + dstCap[j - 1])) { bytes += dstCap[j]; j++; }
        if (j < end && ok(j) && (j == k || (const uint8_t*)dst[j] == (const uint8_t*)dst[j - 1] + dstCap[j - 1])) { bytes += result[j]; j++; }
        else if (j == k) { j++; }   // failed item: nothing to copy
        if (bytes) ZB_CUDA(cudaMemcpyAsync(dst[k], E.dDst.as<uint8_t>() + off[k], bytes, cudaMemcpyDeviceToHost, st));
        k = j;
    }
    return true;
}

// Host-pointer batch decode.  The batch is cut into sub-batches that flow through three stages on separate streams:
// H2D of the compressed bytes (sIn), the kernel pipeline (sComp[k % kPipe], one scratch arena each), D2H of the
// regenerated bytes (sOut).  The number of waves of a sub-batch comes from a host-side walk over the block headers, so
// no stage waits for a read-back; the host only waits for a sub-batch's results before it issues that sub-batch's D2H.
static size_t decompress_batch_host(Engine& E, size_t n, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap, size_t* result)
{
    if (!E.init() || !E.bind()) return (size_t)make_error(kGeneric);
    E.launches = 0; memset(E.timings, 0, sizeof(E.timings));
    if (n == 0) return 0;
    // ---- layout of the packed source / destination buffers (runs of host-adjacent items stay adjacent) ----
    std::vector<uint64_t> sOff(n), dOff; size_t sTotal = 16, dTotal = 0;
    std::vector<Run> runs;
    find_runs(runs, n, src, srcSize);
    for (auto& r : runs) {
        size_t o = sTotal;
        for (size_t k = 0; k < r.count; k++) { sOff[r.first + k] = o; o += srcSize[r.first + k]; }
        sTotal = (o + 15) & ~(size_t)15;
    }
    sTotal += 64;
    size_t const dstRuns = layout_dst(n, dst, dstCap, dOff, &dTotal, E.dict_headroom());
    if (!E.dSrc.ensure(sTotal) || !E.dDst.ensure(dTotal)) return (size_t)make_error(kMemoryAllocation);
    // many scattered buffers, or pageable memory (what Unwrap's `fixed (byte* ...)` over a managed array hands over,
    // Decompressor.cs:62-88): stage through pinned memory with host threads
    static int const forceStage = env_int("ZSTDB200_FORCE_STAGING", 0, 0, 1);
    bool const gather = runs.size() > 512 || forceStage || (n >= 64 && is_pageable(src[0]));
    bool const scatter = dstRuns > 512 || forceStage || (n >= 64 && is_pageable(dst[0]));
    if (gather && !E.hStage.ensure(sTotal)) return (size_t)make_error(kMemoryAllocation);
    if (scatter && !E.hStageOut.ensure(dTotal)) return (size_t)make_error(kMemoryAllocation);
    // ---- sub-batches: sizes grow geometrically (n/16, n/8, n/4, ...) so that the first D2H starts early and the D2H
    // stream -- the slowest stage -- then never runs dry; no sub-batch exceeds the configured target ----
    // staged output: the LAST sub-batch's copy-out by host threads is exposed, so sub-batches stay small there
    static int const stageItems = env_int("ZSTDB200_STAGE_ITEMS", 1024, 16, 8192);
    // pinned, contiguous buffers: 4 sub-batches in flight, first one n/32; staged (pageable / scattered) buffers keep 3 and n/16, the setting their
    // host-thread copies were tuned with (an explicit ZSTDB200_PIPE / ZSTDB200_PIPE_FIRST_DIV applies to both)
    bool const staged = scatter || gather;
    size_t const kPipe = (staged && !getenv("ZSTDB200_PIPE")) ? 3 : (size_t)pipe_depth(), kPipeItems = scatter ? std::min<size_t>(pipe_items(), (size_t)stageItems) : pipe_items();
    std::vector<size_t> sub;                     // sub-batch k = items [sub[k], sub[k+1])
    {
        static int const firstDiv = env_int("ZSTDB200_PIPE_FIRST_DIV", 32, 2, 256);
        size_t const div = (staged && !getenv("ZSTDB200_PIPE_FIRST_DIV")) ? 16 : (size_t)firstDiv;
        size_t pos = 0, sz = std::max<size_t>(64, std::min(kPipeItems, n / div));
        while (pos < n) {
            size_t take = std::min(sz, n - pos);
            if (n - pos - take < take / 2) take = n - pos;      // fold a small remainder into the last sub-batch
            sub.push_back(pos); pos += take;
            sz = std::min(kPipeItems, sz * 2);
        }
        sub.push_back(n);
    }
    size_t const nSub = sub.size() - 1;
    if (!E.need_events(4 * nSub)) return (size_t)make_error(kGeneric);
    cudaEvent_t* const evH2D = E.evPool.data(); cudaEvent_t* const evDone = evH2D + nSub; cudaEvent_t* const evBegin = evDone + nSub; cudaEvent_t* const evD2H = evBegin + nSub;
    auto fail = [&](ErrorCode c) { cudaDeviceSynchronize(); (void)cudaGetLastError(); return (size_t)make_error(c); };
    if (cudaEventRecord(E.evStart, E.stream) != cudaSuccess) return fail(kGeneric);   // order after the caller's stream
    cudaStreamWaitEvent(E.sIn, E.evStart, 0);
    cudaEventRecord(E.evIn[0], E.sIn);
    // stage 1: all uploads are queued up front.  Staged input: a helper thread fills the pinned ring sub-batch by sub-batch (host
    // threads) and queues each DMA, while this thread already launches kernels and drains results; `uploaded` counts the
    // sub-batches whose H2D event has been recorded (a stream wait on an event that is not recorded yet would be a no-op).
    std::atomic<size_t> uploaded{0}; std::atomic<bool> uploadFailed{false};
    std::thread uploader;
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joinUploader{uploader};
    if (gather) {
        int const dev = E.device;
        uploader = std::thread([&, dev]() {
            cudaSetDevice(dev);
            uint8_t* st = E.hStage.as<uint8_t>();
            for (size_t k = 0; k < nSub; k++) {
                size_t const a = sub[k], b = sub[k + 1];
                parallel_for(b - a, 32, [&](size_t x, size_t y) { for (size_t i = a + x; i < a + y; i++) host_copy(st + sOff[i], src[i], srcSize[i]); });
                size_t const lo = sOff[a], hi = sOff[b - 1] + srcSize[b - 1];
                if (hi > lo && cudaMemcpyAsync(E.dSrc.as<uint8_t>() + lo, st + lo, hi - lo, cudaMemcpyHostToDevice, E.sIn) != cudaSuccess) uploadFailed = true;
                cudaEventRecord(evH2D[k], E.sIn);
                if (k + 1 == nSub) cudaEventRecord(E.evIn[1], E.sIn);
                uploaded.store(k + 1, std::memory_order_release);
            }
        });
    }
    size_t runIdx = 0;
    for (size_t k = 0; k < nSub && !gather; k++) {
        size_t const a = sub[k], b = sub[k + 1];
        {
            // the pieces of the runs that intersect [a, b): one DMA each
            size_t i = a;
            while (i < b) {
                while (runs[runIdx].first + runs[runIdx].count <= i) runIdx++;
                size_t const e = std::min(b, runs[runIdx].first + runs[runIdx].count);
                size_t const bytes = sOff[e - 1] + srcSize[e - 1] - sOff[i];
                if (bytes && cudaMemcpyAsync(E.dSrc.as<uint8_t>() + sOff[i], src[i], bytes, cudaMemcpyHostToDevice, E.sIn) != cudaSuccess) return fail(kGeneric);
                i = e;
            }
        }
        cudaEventRecord(evH2D[k], E.sIn);
    }
    if (!gather) { cudaEventRecord(E.evIn[1], E.sIn); uploaded.store(nSub); }
    // stage 2 + 3
    bool firstOut = true;
    // scattered destinations: the sub-batch's device region comes back with one DMA into pinned staging and is copied
    // out by host threads one sub-batch later (while the next DMA is in flight)
    // (a helper thread does the copy-out, so that this thread keeps launching kernels and queueing DMAs meanwhile)
    std::atomic<size_t> d2hQueued{0}; std::atomic<bool> drainFailed{false}, drainStop{false};
    std::thread drainer;
    Joiner joinDrainer{drainer};
    if (scatter) {
        int const dev = E.device;
        drainer = std::thread([&, dev]() {
            cudaSetDevice(dev);
            const uint8_t* st = E.hStageOut.as<uint8_t>();
            for (size_t k = 0; k < nSub; k++) {
                while (d2hQueued.load(std::memory_order_acquire) <= k) { if (drainStop.load()) return; std::this_thread::yield(); }
                if (cudaEventSynchronize(evD2H[k]) != cudaSuccess) { drainFailed = true; return; }
                size_t const a = sub[k], b = sub[k + 1];
                parallel_for(b - a, 32, [&](size_t x, size_t y) { for (size_t i = a + x; i < a + y; i++) if (!is_error(result[i])) host_copy(dst[i], st + dOff[i], result[i]); });
            }
        });
    }
    struct Stopper { std::atomic<bool>& f; ~Stopper() { f = true; } } stopDrainer{drainStop};     // runs before joinDrainer (reverse order of construction)
    auto finish = [&](size_t k) -> bool {
        size_t const a = sub[k], b = sub[k + 1];
        if (cudaEventSynchronize(evDone[k]) != cudaSuccess) return false;
        decode_collect(E.dec[k % kPipe], b - a, srcSize + a, result + a);
        if (firstOut) { cudaEventRecord(E.evOut[0], E.sOut); firstOut = false; }
        if (!scatter) return download_range(E, a, b - a, dst, dstCap, dOff, result, E.sOut);
        size_t const lo = dOff[a], hi = dOff[b - 1] + dstCap[b - 1];
        if (hi > lo && cudaMemcpyAsync(E.hStageOut.as<uint8_t>() + lo, E.dDst.as<uint8_t>() + lo, hi - lo, cudaMemcpyDeviceToHost, E.sOut) != cudaSuccess) return false;
        cudaEventRecord(evD2H[k], E.sOut);
        d2hQueued.store(k + 1, std::memory_order_release);
        return true;
    };
    for (size_t k = 0; k < nSub; k++) {
        size_t const a = sub[k], b = sub[k + 1];
        if (k >= kPipe && !finish(k - kPipe)) return fail(kGeneric);
        uint32_t waves = 1;
        for (size_t i = a; i < b; i++) waves = std::max(waves, count_item_blocks((const uint8_t*)src[i], (uint32_t)std::min<size_t>(srcSize[i], 0xFFFFFFF0u)));
        cudaStream_t const cs = E.sComp[k % kPipe];
        while (uploaded.load(std::memory_order_acquire) <= k) std::this_thread::yield();
        if (uploadFailed) return fail(kGeneric);
        cudaStreamWaitEvent(cs, evH2D[k], 0);
        cudaEventRecord(evBegin[k], cs);
        if (!decode_enqueue(E, E.dec[k % kPipe], cs, b - a, E.dSrc.as<uint8_t>(), sOff.data() + a, srcSize + a, E.dDst.as<uint8_t>(), dOff.data() + a, dstCap + a, waves, nullptr))
            return fail(kMemoryAllocation);
        cudaEventRecord(evDone[k], cs);
    }
    for (size_t k = nSub > kPipe ? nSub - kPipe : 0; k < nSub; k++) if (!finish(k)) return fail(kGeneric);
    if (scatter) { drainer.join(); if (drainFailed) return fail(kGeneric); }
    cudaEventRecord(E.evOut[1], E.sOut);
    if (cudaStreamSynchronize(E.sOut) != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail(kGeneric);
    // the stages overlap: [0] / [2] are the spans of the copy streams, [1] is the sum of the sub-batch kernel spans
    cudaEventElapsedTime(&E.timings[0], E.evIn[0], E.evIn[1]);
    cudaEventElapsedTime(&E.timings[2], E.evOut[0], E.evOut[1]);
    for (size_t k = 0; k < nSub; k++) { float t = 0; cudaEventElapsedTime(&t, evBegin[k], evDone[k]); E.timings[1] += t; }
    if (getenv("ZSTDB200_TRACE")) {               // developer aid: stage times of every sub-batch relative to the first H2D
        for (size_t k = 0; k < nSub; k++) {
            float h = 0, b0 = 0, b1 = 0; cudaEventElapsedTime(&h, E.evIn[0], evH2D[k]); cudaEventElapsedTime(&b0, E.evIn[0], evBegin[k]); cudaEventElapsedTime(&b1, E.evIn[0], evDone[k]);
            fprintf(stderr, "[zstdb200] sub %zu items %zu: h2d done %.2f  kernels %.2f .. %.2f ms\n", k, sub[k + 1] - sub[k], h, b0, b1);
        }
        float o0 = 0, o1 = 0; cudaEventElapsedTime(&o0, E.evIn[0], E.evOut[0]); cudaEventElapsedTime(&o1, E.evIn[0], E.evOut[1]);
        fprintf(stderr, "[zstdb200] d2h %.2f .. %.2f ms\n", o0, o1);
    }
    return 0;
}

// Host-pointer batch compression.  Every item becomes ONE frame, byte-identical to the reference's Wrap: a single block up
// to 128 KiB, a multi-block frame above (window, repcodes and Huffman table carried from block to block; the blocks of a
// frame are a serial chain on the GPU, so this pays off for batches of frames, not for one huge input).
// With ZSTDB200_c_independentChunks = 1 (or for items of 2 GiB and more) a larger item is instead cut into 128 KiB pieces that
// are compressed as independent frames and written back to back: a valid zstd stream for any decoder
// (ZSTD_decompressMultiFrame, ZstdDecompress.cs:1216) whose ZSTD_decompressBound is the item size, all pieces in
// parallel, but NOT the reference's bytes -- see DESIGN.md, deviations.
static size_t compress_batch_host(Engine& E, size_t n, int level, int checksum, int chunked, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap, size_t* result,
                                  const EncDict* dict = nullptr)
{
    auto cut = [&](size_t ss) { return ss > kBlockSizeMax && (chunked || ss > enc_max_frame_bytes()); };
    if (!E.init() || !E.bind()) return (size_t)make_error(kGeneric);
    E.launches = 0; memset(E.timings, 0, sizeof(E.timings));
    if (n == 0) return 0;
    // pieces: item i owns pieces [first[i], first[i+1])
    std::vector<size_t> first(n + 1);
    size_t np = 0;
    for (size_t i = 0; i < n; i++) { first[i] = np; np += !cut(srcSize[i]) ? 1 : (srcSize[i] + kBlockSizeMax - 1) / kBlockSizeMax; }
    first[n] = np;
    // A pipelined pass takes at most kPipeMax sub-batches of 8192 pieces (one arena each): larger batches are run as consecutive slices
    // of whole items (a single item with more pieces than that takes the non-pipelined path below, which loops over 8192-piece passes).
    static int const encPipeCap = env_int("ZSTDB200_ENC_PIPE", 4, 1, kPipeMax);
    size_t const sliceCap = (size_t)encPipeCap * kMaxItemsPerPass;
    if (np > sliceCap && n > 1) {
        size_t a = 0;
        while (a < n) {
            size_t b = a + 1;
            while (b < n && first[b + 1] - first[a] <= sliceCap) b++;
            size_t const rc = compress_batch_host(E, b - a, level, checksum, chunked, src + a, srcSize + a, dst + a, dstCap + a, result + a, dict);
            if (is_error(rc)) return rc;
            a = b;
        }
        return 0;
    }
    std::vector<Run> runs;
    find_runs(runs, n, src, srcSize);
    // One contiguous host buffer and enough pieces: sub-batches flow through H2D (sIn) -> kernels (sComp[k], one arena each)
    // -> compaction + D2H (sOut).  The match finder is latency bound (26 ms for 1024 chunks, 43 ms for 8192), so the
    // sub-batches are NOT run one after the other: their kernels overlap each other and the uploads that are still in flight.
    static int const encPipe = env_int("ZSTDB200_ENC_PIPE", 4, 1, kPipeMax);
    static int const encSub = env_int("ZSTDB200_ENC_SUB", 2048, 256, 8192);       // smallest sub-batch (pieces)
    size_t const nSub = (runs.size() == 1 && np >= 4096 && np <= sliceCap) ? std::min<size_t>((size_t)encPipe, np / (size_t)encSub) : 1;
    std::vector<uint64_t> sOff(n); size_t sTotal = 16;
    if (nSub > 1) {
        size_t o = sTotal;
        for (size_t i = 0; i < n; i++) { sOff[i] = o; o += srcSize[i]; }
        sTotal = ((o + 15) & ~(size_t)15) + 64;
        if (!E.dSrc.ensure(sTotal)) return (size_t)make_error(kMemoryAllocation);
    } else {
        cudaEventRecord(E.ev[10], E.stream);
        if (!upload_items(E, n, src, srcSize, sOff, &sTotal)) return (size_t)make_error(kMemoryAllocation);
        cudaEventRecord(E.ev[11], E.stream);
    }
    // device output: one slot of compressBound(pieceSize) per piece (the reference's Wrap contract, Compressor.cs:80)
    std::vector<uint64_t> pSrcOff(np), pDstOff(np); std::vector<size_t> pSize(np), slotCap(np), r(np);
    size_t dTotal = 16;
    for (size_t i = 0; i < n; i++)
        for (size_t k = first[i]; k < first[i + 1]; k++) {
            size_t const o = (k - first[i]) * (size_t)kBlockSizeMax;
            pSrcOff[k] = sOff[i] + o; pSize[k] = cut(srcSize[i]) ? std::min<size_t>(srcSize[i] - o, kBlockSizeMax) : srcSize[i];
            pDstOff[k] = dTotal; slotCap[k] = enc_compress_bound(pSize[k]); dTotal += (slotCap[k] + 15) & ~(size_t)15;
        }
    dTotal += 16;
    if (!E.dDst.ensure(dTotal)) return (size_t)make_error(kMemoryAllocation);
    std::vector<uint64_t> hOff(np);              // where piece k ends up in the pinned staging buffer (pipelined / compacted paths)
    auto fail = [&](ErrorCode c) { cudaDeviceSynchronize(); (void)cudaGetLastError(); return (size_t)make_error(c); };
    // per item: total size, first error of its pieces, dstSize_tooSmall when the caller's buffer cannot take it (ZstdCompress.cs:4690-4800)
    auto item_result = [&](size_t i) {
        size_t tot = 0;
        for (size_t k = first[i]; k < first[i + 1]; k++) { if (is_error(r[k])) { tot = r[k]; break; } tot += r[k]; }
        if (!is_error(tot) && tot > dstCap[i]) tot = (size_t)make_error(kDstSizeTooSmall);
        result[i] = tot;
    };
    enc_set_overlap_mode(nSub > 1);
    if (nSub > 1) {
        if (!E.hStage.ensure(dTotal) || !E.need_events(4 * nSub)) return (size_t)make_error(kMemoryAllocation);
        cudaEvent_t* const evK = E.evPool.data();                  // [3k..3k+2]: before match / after match / after entropy
        cudaEvent_t* const evOutK = evK + 3 * nSub;
        size_t const per = (np + nSub - 1) / nSub;
        if (cudaEventRecord(E.evStart, E.stream) != cudaSuccess) return fail(kGeneric);   // order after the caller's stream
        cudaStreamWaitEvent(E.sIn, E.evStart, 0);
        cudaEventRecord(E.evIn[0], E.sIn);
        const uint8_t* const hostBase = (const uint8_t*)src[0];
        for (size_t k = 0; k < nSub; k++) {
            size_t const a = k * per, b = std::min(np, a + per);
            size_t const lo = pSrcOff[a], hi = pSrcOff[b - 1] + pSize[b - 1];
            if (hi > lo && cudaMemcpyAsync(E.dSrc.as<uint8_t>() + lo, hostBase + (lo - sOff[0]), hi - lo, cudaMemcpyHostToDevice, E.sIn) != cudaSuccess) return fail(kGeneric);
            if (!enc_enqueue(E.encPipe[k], E.sEnc[k], E.sIn, b - a, level, checksum, E.dSrc.as<uint8_t>(), pSrcOff.data() + a, pSize.data() + a,
                             E.dDst.as<uint8_t>(), pDstOff.data() + a, slotCap.data() + a, evK + 3 * k, &E.launches, dict)) { set_error(enc_last_error()); return fail(kGeneric); }
        }
        cudaEventRecord(E.evIn[1], E.sIn);
        size_t stageBase = 0, itemNext = 0;
        std::vector<size_t> itemEnd(nSub);           // items [itemEnd[k-1], itemEnd[k]) are complete once sub-batch k is in host memory
        double scatterMs = 0;
        // frames of finished items: pinned staging -> the caller's buffers, by host threads, while the next sub-batch is still on its way
        auto scatter_items = [&](size_t iLo, size_t iHi) {
            auto const tc0 = std::chrono::steady_clock::now();
            const uint8_t* st = E.hStage.as<uint8_t>();
            parallel_for(iHi - iLo, 256, [&](size_t a, size_t b) {
                for (size_t i = iLo + a; i < iLo + b; i++) {
                    if (is_error(result[i])) continue;
                    size_t o = 0;
                    for (size_t q = first[i]; q < first[i + 1]; q++) { memcpy((uint8_t*)dst[i] + o, st + hOff[q], r[q]); o += r[q]; }
                }
            });
            scatterMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count();
        };
        for (size_t k = 0; k < nSub; k++) {
            size_t const a = k * per, b = std::min(np, a + per);
            if (cudaEventSynchronize(evK[3 * k + 2]) != cudaSuccess) return fail(kGeneric);
            const uint64_t* hr = enc_results(E.encPipe[k]);
            std::vector<uint64_t> cOff(b - a); std::vector<size_t> cSize(b - a); size_t cTotal = 0;
            for (size_t q = a; q < b; q++) { r[q] = (size_t)hr[q - a]; cOff[q - a] = cTotal; cSize[q - a] = is_error(r[q]) ? 0 : r[q]; hOff[q] = stageBase + cTotal; cTotal += cSize[q - a]; }
            if (k == 0) cudaEventRecord(E.evOut[0], E.sOut);
            if (!enc_compact_device(E.encPipe[k], E.sOut, b - a, E.dDst.as<uint8_t>(), pDstOff.data() + a, cSize.data(), cOff.data(), cTotal, &E.launches)) return fail(kGeneric);
            if (cTotal && cudaMemcpyAsync(E.hStage.as<uint8_t>() + stageBase, E.encPipe[k].compactBuf(), cTotal, cudaMemcpyDeviceToHost, E.sOut) != cudaSuccess) return fail(kGeneric);
            cudaEventRecord(evOutK[k], E.sOut);
            stageBase += cTotal;
            while (itemNext < n && first[itemNext + 1] <= b) { item_result(itemNext); itemNext++; }
            itemEnd[k] = itemNext;
            if (k > 0) {
                if (cudaEventSynchronize(evOutK[k - 1]) != cudaSuccess) return fail(kGeneric);
                scatter_items(k > 1 ? itemEnd[k - 2] : 0, itemEnd[k - 1]);
            }
        }
        cudaEventRecord(E.evOut[1], E.sOut);
        if (cudaStreamSynchronize(E.sOut) != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail(kGeneric);
        scatter_items(nSub > 1 ? itemEnd[nSub - 2] : 0, n);
        cudaEventElapsedTime(&E.timings[0], E.evIn[0], E.evIn[1]);
        cudaEventElapsedTime(&E.timings[2], E.evOut[0], E.evOut[1]);
        for (size_t k = 0; k < nSub; k++) {
            float t = 0;
            cudaEventElapsedTime(&t, evK[3 * k], evK[3 * k + 2]); E.timings[1] += t;
            cudaEventElapsedTime(&t, evK[3 * k], evK[3 * k + 1]); E.timings[8] += t;
            cudaEventElapsedTime(&t, evK[3 * k + 1], evK[3 * k + 2]); E.timings[9] += t;
        }
        if (getenv("ZSTDB200_TRACE")) {           // developer aid: stage times of every sub-batch relative to the first H2D
            for (size_t k = 0; k < nSub; k++) {
                float m0 = 0, m1 = 0, e1 = 0, o1 = 0;
                cudaEventElapsedTime(&m0, E.evIn[0], evK[3 * k]); cudaEventElapsedTime(&m1, E.evIn[0], evK[3 * k + 1]);
                cudaEventElapsedTime(&e1, E.evIn[0], evK[3 * k + 2]); cudaEventElapsedTime(&o1, E.evIn[0], evOutK[k]);
                fprintf(stderr, "[zstdb200] enc sub %zu: match %.2f .. %.2f  entropy .. %.2f  compact+d2h .. %.2f ms\n", k, m0, m1, e1, o1);
            }
            float h1 = 0; cudaEventElapsedTime(&h1, E.evIn[0], E.evIn[1]);
            fprintf(stderr, "[zstdb200] enc h2d span .. %.2f ms, host scatter %.2f ms in total\n", h1, scatterMs);
        }
        return 0;
    } else {
        if (!enc_compress_device(E.enc, E.stream, E.ev, np, level, checksum, E.dSrc.as<uint8_t>(), pSrcOff.data(), pSize.data(), E.dDst.as<uint8_t>(), pDstOff.data(), slotCap.data(), r.data(), E.timings, &E.launches, dict, &E.encPipe[0]))
            { set_error(enc_last_error()); return (size_t)make_error(kGeneric); }
        cudaEventRecord(E.ev[12], E.stream);
    }
    for (size_t i = 0; i < n; i++) item_result(i);
    if (np <= 512) {
        for (size_t i = 0; i < n; i++) {
            if (is_error(result[i])) continue;
            size_t o = 0;
            for (size_t k = first[i]; k < first[i + 1]; k++) {
                if (r[k] && cudaMemcpyAsync((uint8_t*)dst[i] + o, E.dDst.as<uint8_t>() + pDstOff[k], r[k], cudaMemcpyDeviceToHost, E.stream) != cudaSuccess) return (size_t)make_error(kGeneric);
                o += r[k];
            }
        }
        if (cudaStreamSynchronize(E.stream) != cudaSuccess) return (size_t)make_error(kGeneric);
    } else {
        // compact on the device first so that only compressed bytes cross PCIe; the pieces of an item end up adjacent
        std::vector<uint64_t> cOff(np); std::vector<size_t> cSize(np); size_t cTotal = 0;
        for (size_t i = 0; i < n; i++)
            for (size_t k = first[i]; k < first[i + 1]; k++) { cOff[k] = cTotal; cSize[k] = is_error(result[i]) ? 0 : r[k]; cTotal += cSize[k]; }
        if (!E.hStage.ensure(cTotal + 16)) return (size_t)make_error(kMemoryAllocation);
        if (!enc_compact_device(E.enc, E.stream, np, E.dDst.as<uint8_t>(), pDstOff.data(), cSize.data(), cOff.data(), cTotal, &E.launches)) return (size_t)make_error(kGeneric);
        if (cudaMemcpyAsync(E.hStage.p, E.enc.compactBuf(), cTotal, cudaMemcpyDeviceToHost, E.stream) != cudaSuccess) return (size_t)make_error(kGeneric);
        if (cudaStreamSynchronize(E.stream) != cudaSuccess) return (size_t)make_error(kGeneric);
        const uint8_t* st = E.hStage.as<uint8_t>();
        parallel_for(n, 256, [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) if (!is_error(result[i])) memcpy(dst[i], st + cOff[first[i]], result[i]); });
    }
    cudaEventRecord(E.ev[13], E.stream);
    cudaEventSynchronize(E.ev[13]);
    cudaEventElapsedTime(&E.timings[0], E.ev[10], E.ev[11]);
    cudaEventElapsedTime(&E.timings[2], E.ev[12], E.ev[13]);
    return 0;
}

// ZSTD_decompressBound on host memory: header walking only (ZstdDecompress.cs:877-995). Host logic, not compute.
static const uint8_t* rd(const void* p) { return (const uint8_t*)p; }
static uint32_t h_le16(const uint8_t* p) { return p[0] | (p[1] << 8); }
static uint32_t h_le24(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16); }
static uint32_t h_le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t h_le64(const uint8_t* p) { return (uint64_t)h_le32(p) | ((uint64_t)h_le32(p + 4) << 32); }

// ZSTD_findFrameSizeInfo (ZstdDecompress.cs:877): header walking only, host logic.  Returns the compressed size of the
// frame at src (or an error code in the size_t convention) and its decompressed bound.
static size_t frame_size_info_host(const uint8_t* src, size_t srcSize, unsigned long long* dBoundOut)
{
    *dBoundOut = 0ULL - 2;
    if (srcSize >= 8 && (h_le32(src) & kMagicSkippableMask) == kMagicSkippableStart) {      // readSkippableFrameSize :674
        uint32_t const sz = h_le32(src + 4);
        if ((uint32_t)(sz + 8) < sz) return (size_t)make_error(kFrameParameterUnsupported);
        if ((size_t)sz + 8 > srcSize) return (size_t)make_error(kSrcSizeWrong);
        *dBoundOut = 0;
        return (size_t)sz + 8;
    }
    if (srcSize < 5) return (size_t)make_error(kSrcSizeWrong);
    if (h_le32(src) != kMagic) return (size_t)make_error(kPrefixUnknown);
    uint32_t const fhd = src[4];
    uint32_t const dictID = fhd & 3, single = (fhd >> 5) & 1, fcsId = fhd >> 6;
    static const uint32_t did[4] = {0, 1, 2, 4}, fcsz[4] = {0, 2, 4, 8};
    size_t const hs = 5 + !single + did[dictID] + fcsz[fcsId] + (single && !fcsId);
    if (srcSize < hs) return (size_t)make_error(kSrcSizeWrong);
    if (fhd & 8) return (size_t)make_error(kFrameParameterUnsupported);
    size_t pos = 5; unsigned long long windowSize = 0, fcs = 0ULL - 1;
    if (!single) { uint32_t const wl = (src[pos] >> 3) + 10; if (wl > 31) return (size_t)make_error(kWindowTooLarge); windowSize = 1ULL << wl; windowSize += (windowSize >> 3) * (src[pos] & 7); pos++; }
    pos += did[dictID];
    if (fcsId == 0) { if (single) fcs = src[pos]; } else if (fcsId == 1) fcs = h_le16(src + pos) + 256; else if (fcsId == 2) fcs = h_le32(src + pos); else fcs = h_le64(src + pos);
    if (single) windowSize = fcs;
    unsigned long long const blockSizeMax = windowSize < kBlockSizeMax ? windowSize : kBlockSizeMax;
    size_t ip = hs, remaining = srcSize - hs, nbBlocks = 0;
    for (;;) {
        if (remaining < 3) return (size_t)make_error(kSrcSizeWrong);
        uint32_t const h = h_le24(src + ip); uint32_t const type = (h >> 1) & 3, cs = h >> 3;
        if (type == 3) return (size_t)make_error(kCorruptionDetected);
        size_t const csz = type == 1 ? 1 : cs;
        if (3 + csz > remaining) return (size_t)make_error(kSrcSizeWrong);
        ip += 3 + csz; remaining -= 3 + csz; nbBlocks++;
        if (h & 1) break;
    }
    if (fhd & 4) { if (remaining < 4) return (size_t)make_error(kSrcSizeWrong); ip += 4; }
    *dBoundOut = (fcs != 0ULL - 1) ? fcs : (unsigned long long)nbBlocks * blockSizeMax;
    return ip;
}

// ZSTD_decompressBound (ZstdDecompress.cs:971)
static unsigned long long decompress_bound_host(const void* srcV, size_t srcSize)
{
    const uint8_t* src = rd(srcV);
    unsigned long long bound = 0;
    while (srcSize > 0) {
        unsigned long long dBound;
        size_t const compressedSize = frame_size_info_host(src, srcSize, &dBound);
        if (is_error(compressedSize) || dBound == 0ULL - 2) return 0ULL - 2;
        src += compressedSize; srcSize -= compressedSize; bound += dBound;
    }
    return bound;
}

// ---------------------------------------------------------------------------------------------------------------
//  Multi-device scheduler: the host scatter of BASELINE.json's north_star ("batches of independent frames are partitioned
//  across the 8 GPUs of one box by a host scatter; no NCCL, frames share no state"; SURVEY.md section 8e: one host thread + one
//  CUDA context / stream set per GPU, pinned-host scatter in, pinned-host gather out).
// ---------------------------------------------------------------------------------------------------------------
// Contiguous ranges balanced by byte weight (greedy on the prefix sum; the Python mirror is zstdsharp_b200/sharding.py).
static void shard_bounds(size_t n, const size_t* weight, int parts, size_t* bounds /* [parts + 1] */)
{
    std::vector<unsigned long long> csum(n);
    unsigned long long total = 0;
    for (size_t i = 0; i < n; i++) { total += weight[i]; csum[i] = total; }
    bounds[0] = 0;
    for (int r = 1; r < parts; r++) {
        double const target = (double)total * r / parts;
        size_t cut = (size_t)(std::lower_bound(csum.begin(), csum.end(), target, [](unsigned long long c, double t) { return (double)c < t; }) - csum.begin()) + 1;
        cut = std::min(std::max(cut, bounds[r - 1]), n);
        if (cut >= 2 && cut - 1 > bounds[r - 1] && std::abs((double)csum[cut - 2] - target) <= std::abs((double)csum[cut - 1] - target)) cut--;
        bounds[r] = cut;
    }
    bounds[parts] = n;
}

// Pins the calling thread to the CPUs that are local to `device` (its PCIe root's NUMA node, /sys/bus/pci/devices/<id>/local_cpulist):
// the staging memcpy threads and the pinned allocations made afterwards then live next to the GPU.  Best effort: 0 = done.
static int bind_thread_near_device(int device)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    std::string const path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -2;
    char line[4096] = {0};
    char* const got = fgets(line, sizeof(line), f);
    fclose(f);
    if (!got) return -3;
    cpu_set_t set; CPU_ZERO(&set);
    int count = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        int const k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k < 1) continue;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); count++; }
    }
    if (count == 0) return -4;
    return sched_setaffinity(0, sizeof(set), &set) == 0 ? 0 : -5;
}

// ZSTD_fast / ZSTD_dfast levels: the negative levels (ZSTD_fast with an acceleration factor), 0 (= 3), 1..3, and 4 for the input sizes where
// it is still ZSTD_dfast (16 KiB < n <= 128 KiB and n > 256 KiB, Clevels.cs row 4; other sizes answer parameter_unsupported per item)
static bool level_supported(int level) { return level >= -(1 << 17) && level <= 4; }

// The entry points select the context's device (cudaSetDevice is per host thread); the caller's current device is put back on return.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; (void)cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static const char* error_name(uint32_t code)   // ErrorPrivate.cs:34-184
{
    switch (code) {
    case kNoError: return "No error detected";
    case kGeneric: return "Error (generic)";
    case kPrefixUnknown: return "Unknown frame descriptor";
    case kVersionUnsupported: return "Version not supported";
    case kFrameParameterUnsupported: return "Unsupported frame parameter";
    case kWindowTooLarge: return "Frame requires too much memory for decoding";
    case kCorruptionDetected: return "Corrupted block detected";
    case kChecksumWrong: return "Restored data doesn't match checksum";
    case kParameterUnsupported: return "Unsupported parameter";
    case kParameterOutOfBound: return "Parameter is out of bound";
    case kInitMissing: return "Context should be init first";
    case kMemoryAllocation: return "Allocation error : not enough memory";
    case kWorkSpaceTooSmall: return "workSpace buffer is not large enough";
    case kStageWrong: return "Operation not authorized at current processing stage";
    case kTableLogTooLarge: return "tableLog requires too much memory : unsupported";
    case kMaxSymbolValueTooLarge: return "Unsupported max Symbol Value : too large";
    case kMaxSymbolValueTooSmall: return "Specified maxSymbolValue is too small";
    case kDictionaryCorrupted: return "Dictionary is corrupted";
    case kDictionaryWrong: return "Dictionary mismatch";
    case 34: return "Cannot create Dictionary from provided samples";
    case kDstSizeTooSmall: return "Destination buffer is too small";
    case kSrcSizeWrong: return "Src size is incorrect";
    case kDstBufferNull: return "Operation on NULL destination buffer";
    case 100: return "Frame index is too large";
    case 102: return "An I/O error occurred when reading/seeking";
    case 104: return "Destination buffer is wrong";
    case 105: return "Source buffer is wrong";
    default: return "Unspecified error code";
    }
}

}  // namespace zb

// =================================================================================================================
//  extern "C" surface
// =================================================================================================================
struct ZSTD_CCtx_s {
    zb::Engine E; int level = 3; int checksum = 0; int chunked = 0;
    // Compressor.LoadDictionary (Compressor.cs:43-56): the bytes ZSTD_CCtx_loadDictionary stores, and their digest for the level of the first
    // compression that follows (ZSTD_initLocalDict, ZstdCompress.cs:1581: built once, kept until the dictionary is replaced)
    std::vector<uint8_t> dictBytes; zb::EncDict encDict;
};
// The digest of the loaded dictionary (built on first use); *err receives the zstd error code when it cannot be built.
static const zb::EncDict* cctx_dict(ZSTD_CCtx_s* c, int level, size_t* err)
{
    *err = 0;
    if (c->dictBytes.empty()) return nullptr;
    if (!c->encDict.ready) {
        if (!c->E.init() || !c->E.bind()) { *err = (size_t)zb::make_error(zb::kGeneric); return nullptr; }
        size_t const rc = zb::enc_dict_digest(c->encDict, c->E.stream, c->dictBytes.data(), c->dictBytes.size(), level);
        if (zb::is_error(rc)) { *err = rc; return nullptr; }
    }
    return &c->encDict;
}
struct ZSTD_DCtx_s { zb::Engine E; int windowLogMax = 27; };

using zb::make_error;

extern "C" {

ZSTD_CCtx* ZSTD_createCCtx(void) { return new (std::nothrow) ZSTD_CCtx_s(); }
size_t ZSTD_freeCCtx(ZSTD_CCtx* c) { if (c) { zb::DeviceGuard guard; if (c->E.device >= 0) cudaSetDevice(c->E.device); c->encDict.release(); c->E.destroy(); delete c; } return 0; }
ZSTD_DCtx* ZSTD_createDCtx(void) { return new (std::nothrow) ZSTD_DCtx_s(); }
size_t ZSTD_freeDCtx(ZSTD_DCtx* d) { if (d) { d->E.destroy(); delete d; } return 0; }

size_t ZSTD_compressBound(size_t srcSize) { return zb::enc_compress_bound(srcSize); }

size_t ZSTD_CCtx_setParameter(ZSTD_CCtx* cctx, int param, int value)
{
    if (!cctx) return (size_t)make_error(zb::kGeneric);
    if (param == 100) {   // ZSTD_c_compressionLevel (ZSTD_cParameter.cs), bounds per ZSTD_cParam_getBounds (ZstdCompress.cs:444)
        if (!zb::level_supported(value)) return (size_t)make_error(zb::kParameterUnsupported);   // only fast/dfast levels are implemented on the GPU
        cctx->level = value; return 0;
    }
    if (param == 201) {   // ZSTD_c_checksumFlag
        if (value != 0 && value != 1) return (size_t)make_error(zb::kParameterOutOfBound);
        cctx->checksum = value; return 0;
    }
    if (param == 200) { if (value != 1) return (size_t)make_error(zb::kParameterUnsupported); return 0; }   // ZSTD_c_contentSizeFlag: always on
    if (param == ZSTDB200_c_independentChunks) {
        if (value != 0 && value != 1) return (size_t)make_error(zb::kParameterOutOfBound);
        cctx->chunked = value; return 0;
    }
    return (size_t)make_error(zb::kParameterUnsupported);
}

// Compressor.GetParameter (Compressor.cs:35-41 -> U/ZstdCompress.cs:1289-1299): the parameters ZSTD_CCtx_setParameter accepts.
size_t ZSTD_CCtx_getParameter(const ZSTD_CCtx* cctx, int param, int* value)
{
    if (!cctx || !value) return (size_t)make_error(zb::kGeneric);
    if (param == 100) { *value = cctx->level == 0 ? 3 : cctx->level; return 0; }   // level 0 is stored as ZSTD_CLEVEL_DEFAULT (U/ZstdCompress.cs:874-881)
    if (param == 200) { *value = 1; return 0; }
    if (param == 201) { *value = cctx->checksum; return 0; }
    if (param == ZSTDB200_c_independentChunks) { *value = cctx->chunked; return 0; }
    return (size_t)make_error(zb::kParameterUnsupported);
}

// Decompressor.SetParameter / GetParameter (Decompressor.cs:22-34 -> U/ZstdDecompress.cs:2532-2559, 2477-2486).  ZSTD_d_windowLogMax (100)
// is kept with the reference's bounds and default (10..31, 0 -> 27); as in the reference it limits the STREAMING decoder only
// (U/ZstdDecompress.cs:2349-2357), the one-shot path of ZSTD_decompressDCtx never reads it.  The experimental parameters
// (1000..1003) are not implemented: parameter_unsupported.
size_t ZSTD_DCtx_setParameter(ZSTD_DCtx* dctx, int param, int value)
{
    if (!dctx) return (size_t)make_error(zb::kGeneric);
    if (param == 100) {
        if (value == 0) value = 27;
        if (value < 10 || value > 31) return (size_t)make_error(zb::kParameterOutOfBound);
        dctx->windowLogMax = value; return 0;
    }
    return (size_t)make_error(zb::kParameterUnsupported);
}
size_t ZSTD_DCtx_getParameter(const ZSTD_DCtx* dctx, int param, int* value)
{
    if (!dctx || !value) return (size_t)make_error(zb::kGeneric);
    if (param == 100) { *value = dctx->windowLogMax; return 0; }
    return (size_t)make_error(zb::kParameterUnsupported);
}

size_t ZSTDB200_compressBatch(ZSTD_CCtx* cctx, size_t n, int level, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap, size_t* result)
{
    if (!cctx) return (size_t)make_error(zb::kGeneric);
    zb::DeviceGuard guard;
    if (!zb::level_supported(level)) { for (size_t i = 0; i < n; i++) result[i] = (size_t)make_error(zb::kParameterUnsupported); return 0; }
    size_t derr = 0;
    const zb::EncDict* const dict = cctx_dict(cctx, level, &derr);
    if (derr) { for (size_t i = 0; i < n; i++) result[i] = derr; return 0; }
    return zb::compress_batch_host(cctx->E, n, level, cctx->checksum, cctx->chunked, src, srcSize, dst, dstCap, result, dict);
}

// Compressor.LoadDictionary (Compressor.cs:43-56) -> ZSTD_CCtx_loadDictionary (U/ZstdCompress.cs:1683): the bytes are copied and kept; NULL / 0
// returns the context to no-dictionary mode.  The dictionary is digested by the first compression that follows.
size_t ZSTD_CCtx_loadDictionary(ZSTD_CCtx* cctx, const void* dict, size_t dictSize)
{
    if (!cctx) return (size_t)make_error(zb::kGeneric);
    zb::DeviceGuard guard;
    if (cctx->E.device >= 0) cudaSetDevice(cctx->E.device);
    cctx->encDict.release();
    cctx->dictBytes.clear();
    if (dict && dictSize) cctx->dictBytes.assign((const uint8_t*)dict, (const uint8_t*)dict + dictSize);
    return 0;
}

// ZSTD_compressCCtx (U/ZstdCompress.cs:5772): parameters derived from the level alone (contentSize 1, checksum 0), whatever was set
// on the context with ZSTD_CCtx_setParameter; ZSTD_compress2 (:7138) is the entry point that honours the context's parameters.
size_t ZSTD_compressCCtx(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level)
{
    if (!cctx) return (size_t)make_error(zb::kGeneric);
    if (!zb::level_supported(level)) return (size_t)make_error(zb::kParameterUnsupported);
    zb::DeviceGuard guard;
    size_t r = 0; const void* s = src; void* d = dst;
    size_t const rc = zb::compress_batch_host(cctx->E, 1, level, 0, 0, &s, &srcSize, &d, &dstCapacity, &r);
    return zb::is_error(rc) ? rc : r;
}
size_t ZSTD_compress2(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize)
{
    if (!cctx) return (size_t)make_error(zb::kGeneric);
    size_t r = 0; const void* s = src; void* d = dst;
    size_t const rc = ZSTDB200_compressBatch(cctx, 1, cctx->level, &s, &srcSize, &d, &dstCapacity, &r);
    return zb::is_error(rc) ? rc : r;
}

size_t ZSTDB200_decompressBatch(ZSTD_DCtx* dctx, size_t n, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap, size_t* result)
{
    if (!dctx) return (size_t)make_error(zb::kGeneric);
    zb::DeviceGuard guard;
    return zb::decompress_batch_host(dctx->E, n, src, srcSize, dst, dstCap, result);
}
size_t ZSTD_decompressDCtx(ZSTD_DCtx* dctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize)
{
    size_t r = 0; const void* s = src; void* d = dst;
    size_t const rc = ZSTDB200_decompressBatch(dctx, 1, &s, &srcSize, &d, &dstCapacity, &r);
    return zb::is_error(rc) ? rc : r;
}

size_t ZSTD_DCtx_loadDictionary(ZSTD_DCtx* dctx, const void* dict, size_t dictSize)
{
    if (!dctx) return (size_t)make_error(zb::kGeneric);
    zb::DeviceGuard guard;
    zb::Engine& E = dctx->E;
    if (!dict || dictSize == 0) { E.dictLoaded = false; return 0; }           // ZSTD_DCtx_loadDictionary(NULL, 0) clears (:2255)
    if (dictSize > 0x7FFFFFF0u) return (size_t)make_error(zb::kMemoryAllocation);
    if (!E.init() || !E.bind()) return (size_t)make_error(zb::kGeneric);
    E.dictLoaded = false;
    if (!E.dDict.ensure(dictSize + 16) || !E.dDictHuf.ensure(zb::kHufTableEntries * 2) || !E.dDictFse.ensure(zb::kFseTableEntries * 4) || !E.dDictInfo.ensure(zb::kDictInfoWords * 4))
        return (size_t)make_error(zb::kMemoryAllocation);
    if (cudaMemcpyAsync(E.dDict.p, dict, dictSize, cudaMemcpyHostToDevice, E.stream) != cudaSuccess) return (size_t)make_error(zb::kGeneric);
    zb::dec_launch_dict_setup(E.dDict.as<uint8_t>(), (uint32_t)dictSize, E.dDictHuf.as<uint16_t>(), E.dDictFse.as<uint32_t>(), E.dDictInfo.as<uint32_t>(), E.stream);
    if (cudaMemcpyAsync(E.dictInfo, E.dDictInfo.p, sizeof(E.dictInfo), cudaMemcpyDeviceToHost, E.stream) != cudaSuccess ||
        cudaStreamSynchronize(E.stream) != cudaSuccess) { (void)cudaGetLastError(); return (size_t)make_error(zb::kGeneric); }
    if (E.dictInfo[0] != 1) return (size_t)make_error(zb::kDictionaryCorrupted);      // the reference reports it at the first Unwrap (ZstdDecompress.cs:1897)
    E.dictLoaded = true;
    return 0;
}

size_t ZSTDB200_decompressBatchDevice(ZSTD_DCtx* dctx, size_t n, const void* d_src, const uint64_t* srcOffset, const size_t* srcSize,
                                      void* d_dst, const uint64_t* dstOffset, const size_t* dstCapacity, size_t* result)
{
    if (!dctx) return (size_t)make_error(zb::kGeneric);
    zb::DeviceGuard guard;
    zb::Engine& E = dctx->E;
    if (!E.init()) return (size_t)make_error(zb::kGeneric);
    if (E.dictLoaded) { zb::set_error("a dictionary is loaded: dictionary decoding lays the output out itself, use ZSTDB200_decompressBatch"); return (size_t)make_error(zb::kGeneric); }
    E.launches = 0; memset(E.timings, 0, sizeof(E.timings));
    if (n == 0) return 0;
    if (!zb::decode_device(E, n, (const uint8_t*)d_src, srcOffset, srcSize, (uint8_t*)d_dst, dstOffset, dstCapacity, result, true))
        return (size_t)make_error(zb::kGeneric);
    return 0;
}

size_t ZSTDB200_compressBatchDevice(ZSTD_CCtx* cctx, size_t n, int level, const void* d_src, const uint64_t* srcOffset, const size_t* srcSize,
                                    void* d_dst, const uint64_t* dstOffset, const size_t* dstCapacity, size_t* result)
{
    if (!cctx) return (size_t)make_error(zb::kGeneric);
    zb::DeviceGuard guard;
    zb::Engine& E = cctx->E;
    if (!E.init() || !E.bind()) return (size_t)make_error(zb::kGeneric);
    E.launches = 0; memset(E.timings, 0, sizeof(E.timings));
    if (n == 0) return 0;
    if (!zb::level_supported(level)) { for (size_t i = 0; i < n; i++) result[i] = (size_t)make_error(zb::kParameterUnsupported); return 0; }
    size_t derr = 0;
    const zb::EncDict* const dict = cctx_dict(cctx, level, &derr);
    if (derr) { for (size_t i = 0; i < n; i++) result[i] = derr; return 0; }
    if (!zb::enc_compress_device(E.enc, E.stream, E.ev, n, level, cctx->checksum, (const uint8_t*)d_src, srcOffset, srcSize, (uint8_t*)d_dst, dstOffset, dstCapacity, result, E.timings, &E.launches, dict, &E.encPipe[0]))
        { zb::set_error(zb::enc_last_error()); return (size_t)make_error(zb::kGeneric); }
    return 0;
}

unsigned long long ZSTD_decompressBound(const void* src, size_t srcSize) { return zb::decompress_bound_host(src, srcSize); }
size_t ZSTD_findFrameCompressedSize(const void* src, size_t srcSize) { unsigned long long b; return zb::frame_size_info_host((const uint8_t*)src, srcSize, &b); }
unsigned ZSTD_isError(size_t code) { return zb::is_error(code); }
const char* ZSTD_getErrorName(size_t code) { return zb::error_name(zb::is_error(code) ? (uint32_t)(0 - code) : 0); }
unsigned ZSTD_versionNumber(void) { return 1 * 100 * 100 + 5 * 100 + 1; }     /* format/behaviour of zstd 1.5.1 (ZstdCommon.cs:11-20) */
const char* ZSTD_versionString(void) { return "1.5.1"; }

void ZSTDB200_getLastTimings(const void* ctx, float* msOut)
{
    // both context types start with the Engine
    const zb::Engine* E = (const zb::Engine*)ctx;
    for (int i = 0; i < ZSTDB200_TIMING_SLOTS; i++) msOut[i] = E ? E->timings[i] : 0.f;
}
unsigned ZSTDB200_getLastLaunchCount(const void* ctx) { const zb::Engine* E = (const zb::Engine*)ctx; return E ? E->launches : 0; }
const char* ZSTDB200_lastErrorString(void) { return zb::t_lastError.c_str(); }
size_t ZSTDB200_setStream(void* ctx, void* stream)
{
    zb::Engine* E = (zb::Engine*)ctx;
    if (!E) return (size_t)make_error(zb::kGeneric);
    if (!E->init()) return (size_t)make_error(zb::kGeneric);
    E->stream = stream ? (cudaStream_t)stream : E->ownStream;
    return 0;
}
int ZSTDB200_deviceCount(void) { int c = 0; if (cudaGetDeviceCount(&c) != cudaSuccess) return 0; return c; }

// ---- multi-device scheduler (include/zstd_b200.h) ----
struct ZSTDB200_Multi_s {
    std::vector<ZSTD_DCtx_s*> d; std::vector<ZSTD_CCtx_s*> c; std::vector<int> dev; int bindNuma = 1;
};

ZSTDB200_Multi* ZSTDB200_createMulti(int nDevices)
{
    int have = ZSTDB200_deviceCount();
    if (have <= 0) { zb::set_error("no CUDA device"); return nullptr; }
    if (nDevices <= 0 || nDevices > have) nDevices = have;
    ZSTDB200_Multi_s* m = new (std::nothrow) ZSTDB200_Multi_s();
    if (!m) return nullptr;
    for (int k = 0; k < nDevices; k++) {
        ZSTD_DCtx_s* d = new (std::nothrow) ZSTD_DCtx_s(); ZSTD_CCtx_s* c = new (std::nothrow) ZSTD_CCtx_s();
        if (!d || !c) { delete d; delete c; ZSTDB200_freeMulti(m); return nullptr; }
        d->E.wantDevice = k; c->E.wantDevice = k;
        m->d.push_back(d); m->c.push_back(c); m->dev.push_back(k);
    }
    if (const char* e = getenv("ZSTDB200_NUMA_BIND")) m->bindNuma = atoi(e);
    return m;
}
size_t ZSTDB200_freeMulti(ZSTDB200_Multi* m)
{
    if (!m) return 0;
    for (auto* d : m->d) ZSTD_freeDCtx(d);
    for (auto* c : m->c) ZSTD_freeCCtx(c);
    delete m;
    return 0;
}
int ZSTDB200_multiDeviceCount(const ZSTDB200_Multi* m) { return m ? (int)m->dev.size() : 0; }
size_t ZSTDB200_multiSetParameter(ZSTDB200_Multi* m, int param, int value)
{
    if (!m) return (size_t)make_error(zb::kGeneric);
    size_t r = 0;
    for (auto* c : m->c) { size_t const x = ZSTD_CCtx_setParameter(c, param, value); if (zb::is_error(x)) r = x; }
    return r;
}
void ZSTDB200_shardBounds(size_t n, const size_t* weight, int parts, size_t* bounds) { if (parts > 0) zb::shard_bounds(n, weight, parts, bounds); }
int ZSTDB200_bindThreadToDevice(int device) { return zb::bind_thread_near_device(device); }

}  // extern "C"
// One host thread per device: thread k owns the contiguous item range [bounds[k], bounds[k+1]) (balanced by srcSize), runs the
// single-device batch call on its context and writes result[] in the caller's order.  No data crosses between devices.
template <typename Fn>
static size_t multi_run(ZSTDB200_Multi* m, size_t n, const size_t* srcSize, size_t* bounds, Fn&& perDevice)
{
    int const parts = (int)m->dev.size();
    zb::shard_bounds(n, srcSize, parts, bounds);
    std::vector<size_t> rc((size_t)parts, 0);
    std::vector<std::thread> th;
    for (int k = 0; k < parts; k++) {
        if (bounds[k] == bounds[k + 1]) continue;
        th.emplace_back([&, k]() {
            if (m->bindNuma) (void)zb::bind_thread_near_device(m->dev[(size_t)k]);
            rc[(size_t)k] = perDevice(k, bounds[k], bounds[k + 1] - bounds[k]);
        });
    }
    for (auto& t : th) t.join();
    for (int k = 0; k < parts; k++) if (zb::is_error(rc[(size_t)k])) return rc[(size_t)k];
    return 0;
}
extern "C" {

size_t ZSTDB200_decompressBatchMulti(ZSTDB200_Multi* m, size_t n, const void* const* src, const size_t* srcSize,
                                     void* const* dst, const size_t* dstCap, size_t* result)
{
    if (!m || m->dev.empty()) return (size_t)make_error(zb::kGeneric);
    if (n == 0) return 0;
    std::vector<size_t> bounds(m->dev.size() + 1);
    // decode work follows the regenerated bytes more closely than the compressed ones
    return multi_run(m, n, dstCap, bounds.data(), [&](int k, size_t a, size_t cnt) {
        return ZSTDB200_decompressBatch(m->d[(size_t)k], cnt, src + a, srcSize + a, dst + a, dstCap + a, result + a);
    });
}
size_t ZSTDB200_compressBatchMulti(ZSTDB200_Multi* m, size_t n, int level, const void* const* src, const size_t* srcSize,
                                   void* const* dst, const size_t* dstCap, size_t* result)
{
    if (!m || m->dev.empty()) return (size_t)make_error(zb::kGeneric);
    if (n == 0) return 0;
    std::vector<size_t> bounds(m->dev.size() + 1);
    return multi_run(m, n, srcSize, bounds.data(), [&](int k, size_t a, size_t cnt) {
        return ZSTDB200_compressBatch(m->c[(size_t)k], cnt, level, src + a, srcSize + a, dst + a, dstCap + a, result + a);
    });
}
size_t ZSTDB200_multiLoadDictionary(ZSTDB200_Multi* m, const void* dict, size_t dictSize)
{
    if (!m) return (size_t)make_error(zb::kGeneric);
    size_t r = 0;
    for (auto* d : m->d) { size_t const x = ZSTD_DCtx_loadDictionary(d, dict, dictSize); if (zb::is_error(x)) r = x; }
    for (auto* c : m->c) { size_t const x = ZSTD_CCtx_loadDictionary(c, dict, dictSize); if (zb::is_error(x)) r = x; }
    return r;
}

}  // extern "C"
