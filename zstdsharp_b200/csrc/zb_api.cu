// zb_api.cu -- C ABI (include/zstd_b200.h) + batch frame scheduler of libzstdb200.so (product code).
//
// Host side of the drop-in boundary: contexts own a CUDA stream and a grow-only device arena; the batch entry
// points move frames host<->device with as few DMA operations as the caller's layout allows and run the decode
// (zb_decode.cu) / encode (zb_encode.cu) kernel pipelines.  There is no CPU fallback anywhere in this file:
// without a device every call returns ZSTD_error_GENERIC.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zstd_b200.h"
#include "zb_decode.cuh"
#include "zb_encode.cuh"

namespace zb {

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

struct Engine {
    int device = -1;
    bool ready = false;
    cudaStream_t stream = nullptr;      // stream all work of this context is issued on
    cudaStream_t ownStream = nullptr;   // created by the context; replaced by ZSTDB200_setStream
    cudaEvent_t ev[kEvents] = {};
    // decode arena
    DevBuf dItems, dInit, dHuf, dFse, dLit, dSeq, dDefaultFse, dHufList, dSeqList, dCounters, dResults;
    PinBuf hInit, hCounters, hResults;
    bool defaultTablesBuilt = false;
    // host<->device staging for the host-pointer API
    DevBuf dSrc, dDst; PinBuf hStage;
    // encode arena
    EncArena enc;
    DevBuf dEncInit; PinBuf hEncInit;
    // instrumentation
    float timings[ZSTDB200_TIMING_SLOTS] = {};
    unsigned launches = 0;

    bool init() {
        if (ready) return true;
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) { set_error(std::string("no CUDA device: ") + cudaGetErrorString(e)); return false; }
        int dev = 0;
        if (const char* env = getenv("ZSTDB200_DEVICE")) dev = atoi(env);
        else if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        if (dev < 0 || dev >= count) { set_error("ZSTDB200_DEVICE out of range"); return false; }
        device = dev;
        ZB_CUDA(cudaSetDevice(device));
        ZB_CUDA(cudaStreamCreateWithFlags(&ownStream, cudaStreamNonBlocking));
        if (!stream) stream = ownStream;
        for (int i = 0; i < kEvents; i++) ZB_CUDA(cudaEventCreate(&ev[i]));
        ready = true;
        return true;
    }
    void destroy() {
        if (device >= 0) cudaSetDevice(device);
        DevBuf* d[] = {&dItems, &dInit, &dHuf, &dFse, &dLit, &dSeq, &dDefaultFse, &dHufList, &dSeqList, &dCounters, &dResults, &dSrc, &dDst, &dEncInit};
        for (auto* b : d) b->release();
        PinBuf* h[] = {&hInit, &hCounters, &hResults, &hStage, &hEncInit};
        for (auto* b : h) b->release();
        enc.release();
        if (ready) { for (int i = 0; i < kEvents; i++) cudaEventDestroy(ev[i]); cudaStreamDestroy(ownStream); }
        ready = false;
    }
    bool bind() { ZB_CUDA(cudaSetDevice(device)); return true; }
};

// ---------------------------------------------------------------------------------------------------------------
//  Decode pass over device-resident frames
// ---------------------------------------------------------------------------------------------------------------
static bool decode_device(Engine& E, size_t n, const uint8_t* d_src, const uint64_t* srcOff, const size_t* srcSize,
                          uint8_t* d_dst, const uint64_t* dstOff, const size_t* dstCap, size_t* result, bool timeKernels)
{
    if (!E.bind()) return false;
    float kernelMs = 0; float slotMs[5] = {0, 0, 0, 0, 0};
    for (size_t base = 0; base < n; base += kMaxItemsPerPass) {
        size_t const m = std::min(kMaxItemsPerPass, n - base);
        if (!E.dItems.ensure(m * sizeof(DecItem)) || !E.dInit.ensure(m * sizeof(DecItemInit)) || !E.hInit.ensure(m * sizeof(DecItemInit)) ||
            !E.dHuf.ensure(m * kHufTableEntries * 2) || !E.dFse.ensure(m * kFseTableEntries * 4) || !E.dLit.ensure(m * (size_t)kLitStride) ||
            !E.dSeq.ensure(m * (size_t)kSeqCap * 8) ||
            !E.dDefaultFse.ensure(kFseTableEntries * 4) || !E.dHufList.ensure(m * 4) || !E.dSeqList.ensure(m * 4) ||
            !E.dCounters.ensure(64) || !E.hCounters.ensure(64) || !E.dResults.ensure(m * 8) || !E.hResults.ensure(m * 8))
            return false;
        if (!E.defaultTablesBuilt) { dec_build_default_tables(E.dDefaultFse.as<uint32_t>(), E.stream); E.launches++; E.defaultTablesBuilt = true; }
        DecItemInit* hi = E.hInit.as<DecItemInit>();
        // oversized items cannot be addressed by the 32-bit cursors: they are reported per item below
        for (size_t i = 0; i < m; i++) {
            size_t const ss = srcSize[base + i], dc = dstCap[base + i];
            hi[i].srcOff = srcOff[base + i]; hi[i].dstOff = dstOff[base + i];
            hi[i].srcSize = (uint32_t)std::min<size_t>(ss, 0xFFFFFFF0u);
            hi[i].dstCap = (uint32_t)std::min<size_t>(dc, 0xFFFFFFF0u);
        }
        ZB_CUDA(cudaMemcpyAsync(E.dInit.p, hi, m * sizeof(DecItemInit), cudaMemcpyHostToDevice, E.stream));
        ZB_CUDA(cudaMemsetAsync(E.dCounters.p, 0, 64, E.stream));
        DecPass p;
        p.items = E.dItems.as<DecItem>(); p.nItems = (uint32_t)m; p.src = d_src; p.dst = d_dst;
        p.hufTable = E.dHuf.as<uint16_t>(); p.fseTable = E.dFse.as<uint32_t>(); p.litBuf = E.dLit.as<uint8_t>();
        p.seq = E.dSeq.as<uint2>();
        p.defaultFse = E.dDefaultFse.as<uint32_t>(); p.hufList = E.dHufList.as<uint32_t>(); p.seqList = E.dSeqList.as<uint32_t>();
        p.counters = E.dCounters.as<uint32_t>(); p.results = E.dResults.as<uint64_t>();
        ZB_CUDA(cudaEventRecord(E.ev[0], E.stream));
        dec_launch_scan_init(p, E.dInit.p, E.stream); E.launches++;
        ZB_CUDA(cudaMemcpyAsync(E.hCounters.p, E.dCounters.p, 16, cudaMemcpyDeviceToHost, E.stream));
        ZB_CUDA(cudaStreamSynchronize(E.stream));
        uint32_t const nWaves = E.hCounters.as<uint32_t>()[3];
        for (uint32_t w = 0; w < nWaves; w++) {
            if (timeKernels && w == 0) {
                dec_launch_wave_timed(p, E.stream, &E.ev[2]);   // records ev[2..6] between the kernels of the first wave
            } else dec_launch_wave(p, E.stream);
            E.launches += 5;
        }
        dec_launch_finish(p, E.stream); E.launches++;
        ZB_CUDA(cudaEventRecord(E.ev[1], E.stream));
        ZB_CUDA(cudaMemcpyAsync(E.hResults.p, E.dResults.p, m * 8, cudaMemcpyDeviceToHost, E.stream));
        ZB_CUDA(cudaStreamSynchronize(E.stream));
        ZB_CUDA(cudaGetLastError());
        float ms = 0; cudaEventElapsedTime(&ms, E.ev[0], E.ev[1]); kernelMs += ms;
        if (timeKernels && nWaves > 0) {
            float t;
            cudaEventElapsedTime(&t, E.ev[0], E.ev[2]); slotMs[0] += t;              // scan (+ sync gap)
            for (int k = 0; k < 4; k++) { cudaEventElapsedTime(&t, E.ev[2 + k], E.ev[3 + k]); slotMs[1 + k] += t; }
        }
        const uint64_t* hr = E.hResults.as<uint64_t>();
        for (size_t i = 0; i < m; i++) {
            if (srcSize[base + i] > 0xFFFFFFF0u) result[base + i] = (size_t)make_error(kSrcSizeWrong);
            else result[base + i] = (size_t)hr[i];
        }
    }
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

static void parallel_for(size_t n, size_t grain, const std::function<void(size_t, size_t)>& fn)
{
    unsigned hw = std::thread::hardware_concurrency(); if (hw == 0) hw = 4;
    size_t const nt = std::min<size_t>(std::min<unsigned>(hw, 16), (n + grain - 1) / grain);
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

// copies result[i] bytes of every item back to the caller
static bool download_items(Engine& E, size_t n, void* const* dst, const size_t* dstCap, const std::vector<uint64_t>& off, const size_t* result, size_t total)
{
    std::vector<Run> runs;
    find_runs(runs, n, (const void* const*)dst, dstCap);
    auto ok = [&](size_t i) { return !is_error(result[i]); };
    if (runs.size() <= 512) {
        for (auto& r : runs) {
            // one DMA while items are completely filled; partial / failed items are copied individually
            size_t k = 0;
            while (k < r.count) {
                size_t i = r.first + k; size_t j = k; size_t bytes = 0;
                while (j < r.count && ok(r.first + j) && result[r.first + j] == dstCap[r.first + j]) { bytes += dstCap[r.first + j]; j++; }
                if (j < r.count && ok(r.first + j)) { bytes += result[r.first + j]; j++; }
                else if (j == k) { j++; }   // failed item: nothing to copy
                if (bytes) ZB_CUDA(cudaMemcpyAsync(dst[i], E.dDst.as<uint8_t>() + off[i], bytes, cudaMemcpyDeviceToHost, E.stream));
                k = j;
            }
        }
        ZB_CUDA(cudaStreamSynchronize(E.stream));
    } else {
        if (!E.hStage.ensure(total)) return false;
        ZB_CUDA(cudaMemcpyAsync(E.hStage.p, E.dDst.p, total, cudaMemcpyDeviceToHost, E.stream));
        ZB_CUDA(cudaStreamSynchronize(E.stream));
        const uint8_t* st = E.hStage.as<uint8_t>();
        parallel_for(n, 256, [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) if (ok(i)) memcpy(dst[i], st + off[i], result[i]); });
    }
    return true;
}


static void layout_dst(size_t n, void* const* dst, const size_t* dstCap, std::vector<uint64_t>& off, size_t* total)
{
    std::vector<Run> runs;
    find_runs(runs, n, (const void* const*)dst, dstCap);
    off.resize(n);
    size_t t = 16;
    for (auto& r : runs) {
        size_t o = t;
        for (size_t k = 0; k < r.count; k++) { off[r.first + k] = o; o += dstCap[r.first + k]; }
        t = (o + 15) & ~(size_t)15;
    }
    *total = t + 16;
}

static size_t decompress_batch_host(Engine& E, size_t n, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap, size_t* result)
{
    if (!E.init() || !E.bind()) return (size_t)make_error(kGeneric);
    E.launches = 0; memset(E.timings, 0, sizeof(E.timings));
    if (n == 0) return 0;
    std::vector<uint64_t> sOff, dOff; size_t sTotal = 0, dTotal = 0;
    cudaEventRecord(E.ev[10], E.stream);
    if (!upload_items(E, n, src, srcSize, sOff, &sTotal)) return (size_t)make_error(kMemoryAllocation);
    cudaEventRecord(E.ev[11], E.stream);
    layout_dst(n, dst, dstCap, dOff, &dTotal);
    if (!E.dDst.ensure(dTotal)) return (size_t)make_error(kMemoryAllocation);
    if (!decode_device(E, n, E.dSrc.as<uint8_t>(), sOff.data(), srcSize, E.dDst.as<uint8_t>(), dOff.data(), dstCap, result, false))
        return (size_t)make_error(kGeneric);
    cudaEventRecord(E.ev[12], E.stream);
    if (!download_items(E, n, dst, dstCap, dOff, result, dTotal)) return (size_t)make_error(kGeneric);
    cudaEventRecord(E.ev[13], E.stream);
    cudaEventSynchronize(E.ev[13]);
    cudaEventElapsedTime(&E.timings[0], E.ev[10], E.ev[11]);
    cudaEventElapsedTime(&E.timings[2], E.ev[12], E.ev[13]);
    return 0;
}

static size_t compress_batch_host(Engine& E, size_t n, int level, int checksum, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap, size_t* result)
{
    if (!E.init() || !E.bind()) return (size_t)make_error(kGeneric);
    E.launches = 0; memset(E.timings, 0, sizeof(E.timings));
    if (n == 0) return 0;
    if (checksum) { for (size_t i = 0; i < n; i++) result[i] = (size_t)make_error(kParameterUnsupported); return 0; }
    std::vector<uint64_t> sOff, dOff; size_t sTotal = 0;
    cudaEventRecord(E.ev[10], E.stream);
    if (!upload_items(E, n, src, srcSize, sOff, &sTotal)) return (size_t)make_error(kMemoryAllocation);
    cudaEventRecord(E.ev[11], E.stream);
    // device output: one slot of compressBound(srcSize) per item (the reference's Wrap contract, Compressor.cs:80)
    dOff.resize(n); std::vector<size_t> slotCap(n); size_t dTotal = 16;
    for (size_t i = 0; i < n; i++) { dOff[i] = dTotal; slotCap[i] = enc_compress_bound(srcSize[i]); dTotal += (slotCap[i] + 15) & ~(size_t)15; }
    dTotal += 16;
    if (!E.dDst.ensure(dTotal)) return (size_t)make_error(kMemoryAllocation);
    std::vector<size_t> r(n);
    if (!enc_compress_device(E.enc, E.stream, E.ev, n, level, E.dSrc.as<uint8_t>(), sOff.data(), srcSize, E.dDst.as<uint8_t>(), dOff.data(), slotCap.data(), r.data(), E.timings, &E.launches))
        { set_error(enc_last_error()); return (size_t)make_error(kGeneric); }
    cudaEventRecord(E.ev[12], E.stream);
    // results that do not fit the caller's capacity become dstSize_tooSmall (ZstdCompress.cs:4690-4800 error paths)
    for (size_t i = 0; i < n; i++) {
        if (!is_error(r[i]) && r[i] > dstCap[i]) r[i] = (size_t)make_error(kDstSizeTooSmall);
        result[i] = r[i];
    }
    // copy back: gather compacted frames. Contiguous caller buffers with exact sizes are rare here (sizes are data dependent),
    // so each frame is one DMA when n is small, otherwise the whole slot area is staged through pinned memory.
    if (n <= 512) {
        for (size_t i = 0; i < n; i++)
            if (!is_error(result[i]) && result[i])
                if (cudaMemcpyAsync(dst[i], E.dDst.as<uint8_t>() + dOff[i], result[i], cudaMemcpyDeviceToHost, E.stream) != cudaSuccess) return (size_t)make_error(kGeneric);
        if (cudaStreamSynchronize(E.stream) != cudaSuccess) return (size_t)make_error(kGeneric);
    } else {
        // compact on the device first so that only compressed bytes cross PCIe
        std::vector<uint64_t> cOff(n); size_t cTotal = 0;
        for (size_t i = 0; i < n; i++) { cOff[i] = cTotal; cTotal += is_error(result[i]) ? 0 : result[i]; }
        if (!E.hStage.ensure(cTotal + 16)) return (size_t)make_error(kMemoryAllocation);
        if (!enc_compact_device(E.enc, E.stream, n, E.dDst.as<uint8_t>(), dOff.data(), result, cOff.data(), cTotal, &E.launches)) return (size_t)make_error(kGeneric);
        if (cudaMemcpyAsync(E.hStage.p, E.enc.compactBuf(), cTotal, cudaMemcpyDeviceToHost, E.stream) != cudaSuccess) return (size_t)make_error(kGeneric);
        if (cudaStreamSynchronize(E.stream) != cudaSuccess) return (size_t)make_error(kGeneric);
        const uint8_t* st = E.hStage.as<uint8_t>();
        parallel_for(n, 256, [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) if (!is_error(result[i])) memcpy(dst[i], st + cOff[i], result[i]); });
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

static unsigned long long decompress_bound_host(const void* srcV, size_t srcSize)
{
    const uint8_t* src = rd(srcV);
    unsigned long long const kErr = 0ULL - 2; unsigned long long bound = 0;
    while (srcSize > 0) {
        size_t compressedSize; unsigned long long dBound;
        if (srcSize >= 8 && (h_le32(src) & kMagicSkippableMask) == kMagicSkippableStart) {
            uint32_t const sz = h_le32(src + 4);
            if ((uint32_t)(sz + 8) < sz) return kErr;
            if ((size_t)sz + 8 > srcSize) return kErr;
            compressedSize = (size_t)sz + 8; dBound = 0;
        } else {
            if (srcSize < 5) return kErr;
            if (h_le32(src) != kMagic) return kErr;
            uint32_t const fhd = src[4];
            uint32_t const dictID = fhd & 3, single = (fhd >> 5) & 1, fcsId = fhd >> 6;
            static const uint32_t did[4] = {0, 1, 2, 4}, fcsz[4] = {0, 2, 4, 8};
            size_t const hs = 5 + !single + did[dictID] + fcsz[fcsId] + (single && !fcsId);
            if (srcSize < hs) return kErr;
            if (fhd & 8) return kErr;
            size_t pos = 5; unsigned long long windowSize = 0, fcs = 0ULL - 1;
            if (!single) { uint32_t const wl = (src[pos] >> 3) + 10; if (wl > 31) return kErr; windowSize = 1ULL << wl; windowSize += (windowSize >> 3) * (src[pos] & 7); pos++; }
            pos += did[dictID];
            if (fcsId == 0) { if (single) fcs = src[pos]; } else if (fcsId == 1) fcs = h_le16(src + pos) + 256; else if (fcsId == 2) fcs = h_le32(src + pos); else fcs = h_le64(src + pos);
            if (single) windowSize = fcs;
            unsigned long long const blockSizeMax = windowSize < kBlockSizeMax ? windowSize : kBlockSizeMax;
            size_t ip = hs, remaining = srcSize - hs, nbBlocks = 0;
            for (;;) {
                if (remaining < 3) return kErr;
                uint32_t const h = h_le24(src + ip); uint32_t const type = (h >> 1) & 3, cs = h >> 3;
                if (type == 3) return kErr;
                size_t const csz = type == 1 ? 1 : cs;
                if (3 + csz > remaining) return kErr;
                ip += 3 + csz; remaining -= 3 + csz; nbBlocks++;
                if (h & 1) break;
            }
            if (fhd & 4) { if (remaining < 4) return kErr; ip += 4; }
            compressedSize = ip;
            dBound = (fcs != 0ULL - 1) ? fcs : (unsigned long long)nbBlocks * blockSizeMax;
        }
        src += compressedSize; srcSize -= compressedSize; bound += dBound;
    }
    return bound;
}

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
struct ZSTD_CCtx_s { zb::Engine E; int level = 3; int checksum = 0; };
struct ZSTD_DCtx_s { zb::Engine E; };

using zb::make_error;

extern "C" {

ZSTD_CCtx* ZSTD_createCCtx(void) { return new (std::nothrow) ZSTD_CCtx_s(); }
size_t ZSTD_freeCCtx(ZSTD_CCtx* c) { if (c) { c->E.destroy(); delete c; } return 0; }
ZSTD_DCtx* ZSTD_createDCtx(void) { return new (std::nothrow) ZSTD_DCtx_s(); }
size_t ZSTD_freeDCtx(ZSTD_DCtx* d) { if (d) { d->E.destroy(); delete d; } return 0; }

size_t ZSTD_compressBound(size_t srcSize) { return zb::enc_compress_bound(srcSize); }

size_t ZSTD_CCtx_setParameter(ZSTD_CCtx* cctx, int param, int value)
{
    if (!cctx) return (size_t)make_error(zb::kGeneric);
    if (param == 100) {   // ZSTD_c_compressionLevel (ZSTD_cParameter.cs), bounds per ZSTD_cParam_getBounds (ZstdCompress.cs:444)
        if (value < 0 || value > 3) return (size_t)make_error(zb::kParameterUnsupported);   // only fast/dfast levels are implemented on the GPU
        cctx->level = value; return 0;
    }
    if (param == 201) {   // ZSTD_c_checksumFlag
        if (value != 0) return (size_t)make_error(zb::kParameterUnsupported);
        cctx->checksum = 0; return 0;
    }
    if (param == 200) { if (value != 1) return (size_t)make_error(zb::kParameterUnsupported); return 0; }   // ZSTD_c_contentSizeFlag: always on
    return (size_t)make_error(zb::kParameterUnsupported);
}

size_t ZSTDB200_compressBatch(ZSTD_CCtx* cctx, size_t n, int level, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap, size_t* result)
{
    if (!cctx) return (size_t)make_error(zb::kGeneric);
    if (level < 0 || level > 3) { for (size_t i = 0; i < n; i++) result[i] = (size_t)make_error(zb::kParameterUnsupported); return 0; }
    return zb::compress_batch_host(cctx->E, n, level, cctx->checksum, src, srcSize, dst, dstCap, result);
}

size_t ZSTD_compressCCtx(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level)
{
    size_t r = 0; const void* s = src; void* d = dst;
    size_t const rc = ZSTDB200_compressBatch(cctx, 1, level, &s, &srcSize, &d, &dstCapacity, &r);
    return zb::is_error(rc) ? rc : r;
}
size_t ZSTD_compress2(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize)
{
    if (!cctx) return (size_t)make_error(zb::kGeneric);
    return ZSTD_compressCCtx(cctx, dst, dstCapacity, src, srcSize, cctx->level);
}

size_t ZSTDB200_decompressBatch(ZSTD_DCtx* dctx, size_t n, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap, size_t* result)
{
    if (!dctx) return (size_t)make_error(zb::kGeneric);
    return zb::decompress_batch_host(dctx->E, n, src, srcSize, dst, dstCap, result);
}
size_t ZSTD_decompressDCtx(ZSTD_DCtx* dctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize)
{
    size_t r = 0; const void* s = src; void* d = dst;
    size_t const rc = ZSTDB200_decompressBatch(dctx, 1, &s, &srcSize, &d, &dstCapacity, &r);
    return zb::is_error(rc) ? rc : r;
}

size_t ZSTDB200_decompressBatchDevice(ZSTD_DCtx* dctx, size_t n, const void* d_src, const uint64_t* srcOffset, const size_t* srcSize,
                                      void* d_dst, const uint64_t* dstOffset, const size_t* dstCapacity, size_t* result)
{
    if (!dctx) return (size_t)make_error(zb::kGeneric);
    zb::Engine& E = dctx->E;
    if (!E.init()) return (size_t)make_error(zb::kGeneric);
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
    zb::Engine& E = cctx->E;
    if (!E.init() || !E.bind()) return (size_t)make_error(zb::kGeneric);
    E.launches = 0; memset(E.timings, 0, sizeof(E.timings));
    if (n == 0) return 0;
    if (level < 0 || level > 3) { for (size_t i = 0; i < n; i++) result[i] = (size_t)make_error(zb::kParameterUnsupported); return 0; }
    if (!zb::enc_compress_device(E.enc, E.stream, E.ev, n, level, (const uint8_t*)d_src, srcOffset, srcSize, (uint8_t*)d_dst, dstOffset, dstCapacity, result, E.timings, &E.launches))
        { zb::set_error(zb::enc_last_error()); return (size_t)make_error(zb::kGeneric); }
    return 0;
}

unsigned long long ZSTD_decompressBound(const void* src, size_t srcSize) { return zb::decompress_bound_host(src, srcSize); }
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

}  // extern "C"
