// zb_encode.cuh -- host-visible interface of the batch encoder (product code).
#pragma once
#include "zb_common.cuh"

namespace zb {

struct EncArenaImpl;
struct EncArena {
    EncArenaImpl* impl = nullptr;
    void release();
    const uint8_t* compactBuf() const;
};

// A loaded compression dictionary, digested on the device for one level (Compressor.LoadDictionary, Compressor.cs:43-56: ZSTD_CCtx_loadDictionary
// stores the bytes, the first ZSTD_compress2 builds the CDict with the level the context has then: ZstdCompress.cs:1581).
struct EncDictImpl;
struct EncDict {
    EncDictImpl* impl = nullptr; int level = 0; bool ready = false;
    void release();
};
// Returns 0 or a zstd error code (memory_allocation for a dictionary the reference cannot digest either, parameter_unsupported for a level outside ZSTD_fast / ZSTD_dfast).
size_t enc_dict_digest(EncDict& D, cudaStream_t stream, const void* dict, size_t dictSize, int level);

// ZSTD_compressBound, ZstdCompress.cs:19-22
inline size_t enc_compress_bound(size_t srcSize)
{ return srcSize + (srcSize >> 8) + ((srcSize < (128u << 10)) ? (((128u << 10) - srcSize) >> 11) : 0); }

// Largest frame the encoder takes (positions inside a frame are 31-bit integers); above: ZSTD_error_srcSize_wrong.
size_t enc_max_frame_bytes();

// Compresses n device-resident chunks (each one frame; frames above 128 KiB are multi-block frames, see enc_enqueue) at `level` (0..3). result[i] = frame size or error code.
// timings: [1] all kernels, [8] match finder, [9] entropy stage.
bool enc_compress_device(EncArena& A, cudaStream_t stream, cudaEvent_t* ev, size_t n, int level, int checksumFlag,
                         const uint8_t* d_src, const uint64_t* srcOff, const size_t* srcSize,
                         uint8_t* d_dst, const uint64_t* dstOff, const size_t* dstCap, size_t* result,
                         float* timings, unsigned* launches, const EncDict* dict = nullptr, EncArena* second = nullptr);
// Queues one pass (m <= 8192 chunks) on `stream` without waiting; descriptors are copied on `copyStream`; results land in
// pinned host memory (enc_results) once `stream` has drained.  ev3 (optional): events before/after match, after entropy.
bool enc_enqueue(EncArena& A, cudaStream_t stream, cudaStream_t copyStream, size_t m, int level, int checksumFlag,
                 const uint8_t* d_src, const uint64_t* srcOff, const size_t* srcSize,
                 uint8_t* d_dst, const uint64_t* dstOff, const size_t* dstCap, cudaEvent_t* ev3, unsigned* launches, const EncDict* dict = nullptr);
const uint64_t* enc_results(const EncArena& A);
// Frames one enc_enqueue pass can hold when its largest block has maxBlockBytes: 8192 at 128 KiB, more for smaller records (up to 65536).
size_t enc_pass_capacity(size_t maxBlockBytes);
// Gathers the n variable-size frames into one dense device buffer (A.compactBuf()) at offsets cOff.
bool enc_compact_device(EncArena& A, cudaStream_t stream, size_t n, const uint8_t* d_dst, const uint64_t* dstOff,
                        const size_t* sizes, const uint64_t* cOff, size_t total, unsigned* launches);
// Lets the encoder's kernels of different streams share SMs (pipelined host path) or restores each kernel's own L1 / shared-memory split.
void enc_set_overlap_mode(bool overlap);
const char* enc_last_error();

}  // namespace zb
