/*
 * zstd_b200.h -- C ABI of libzstdb200.so, the B200-native drop-in for the ZstdSharp hot path
 * (many independent zstd frames: decompress any level, compress levels 1..3).
 *
 * The first block mirrors, name for name and argument for argument, the native zstd entry points the
 * reference already binds through P/Invoke (reference: src/Zstd.Extern/ExternMethods.cs:8-37,
 * DllImport("libzstd", CallingConvention.Cdecl)); binding ZstdSharp to this library is a DllName change
 * (see INTEGRATION.md).  The functions replace the managed implementations
 *     ZSTD_compress2        src/ZstdSharp/Unsafe/ZstdCompress.cs:7138   (called by Compressor.Wrap, Compressor.cs:78-96)
 *     ZSTD_compressCCtx     src/ZstdSharp/Unsafe/ZstdCompress.cs:5772   (called by Benchmark.cs:62-70)
 *     ZSTD_decompressDCtx   src/ZstdSharp/Unsafe/ZstdDecompress.cs:1365 (called by Decompressor.Unwrap, Decompressor.cs:62-88)
 *     ZSTD_decompressBound  src/ZstdSharp/Unsafe/ZstdDecompress.cs:971  (called by Decompressor.GetDecompressedSize, :50-54)
 *     ZSTD_compressBound    src/ZstdSharp/Unsafe/ZstdCompress.cs:19
 *     ZSTD_CCtx_setParameter src/ZstdSharp/Unsafe/ZstdCompress.cs:784   (Compressor.Level setter, Compressor.cs:16-33)
 *     ZSTD_isError / ZSTD_getErrorName  src/ZstdSharp/Unsafe/ZstdCommon.cs:26,33 (ThrowHelper.cs:10-16)
 * Conventions kept from the reference: size_t results, error <=> result > (size_t)-120 with code = 0 - result
 * (Unsafe/ErrorPrivate.cs:10-13, Unsafe/ZSTD_ErrorCode.cs:5-35); caller owns src/dst and the library never
 * retains them after a call; a context is single-threaded, distinct contexts may run concurrently
 * (ZstdNetTests.cs:498-522).
 *
 * All compute runs on the GPU (CUDA, sm_100a).  There is no CPU fallback: without a usable device every
 * compress/decompress entry point returns ZSTD_error_GENERIC and ZSTDB200_lastErrorString() says why.
 *
 * The second block (ZSTDB200_*) is new: the batch frame scheduler the reference does not have.
 */
#ifndef ZSTD_B200_H
#define ZSTD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#  define ZSTDB200_API __attribute__((visibility("default")))
#else
#  define ZSTDB200_API
#endif

typedef struct ZSTD_CCtx_s ZSTD_CCtx;
typedef struct ZSTD_DCtx_s ZSTD_DCtx;

/* ---- zstd-named surface (ExternMethods.cs:10-37) ---- */
ZSTDB200_API ZSTD_CCtx* ZSTD_createCCtx(void);                                                   /* ExternMethods.cs:11 */
ZSTDB200_API size_t     ZSTD_freeCCtx(ZSTD_CCtx* cctx);                                           /* :14 */
ZSTDB200_API size_t     ZSTD_compressCCtx(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity,
                                          const void* src, size_t srcSize, int compressionLevel);  /* :17 */
ZSTDB200_API size_t     ZSTD_compress2(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity,
                                       const void* src, size_t srcSize);                           /* :21 */
ZSTDB200_API ZSTD_DCtx* ZSTD_createDCtx(void);                                                   /* :24 */
ZSTDB200_API size_t     ZSTD_freeDCtx(ZSTD_DCtx* dctx);                                           /* :27 */
ZSTDB200_API size_t     ZSTD_decompressDCtx(ZSTD_DCtx* dctx, void* dst, size_t dstCapacity,
                                            const void* src, size_t srcSize);                      /* :30 */
ZSTDB200_API size_t     ZSTD_compressBound(size_t srcSize);                                        /* :34 */
/* param: ZSTD_c_compressionLevel = 100 (levels 0..3; 0 means 3), ZSTD_c_checksumFlag = 201 (0/1: XXH64 trailer, U/ZstdCompress.cs:5641-5652),
 * ZSTD_c_contentSizeFlag = 200 (1 only), ZSTDB200_c_independentChunks = 10001, see below;
 * anything else -> ZSTD_error_parameter_unsupported. */
/* New, host-pointer API only.  0 (default): every item becomes one frame with the reference's bytes; items above 128 KiB are
 * multi-block frames (U/ZstdCompress.cs:4690 ZSTD_compress_frameChunk), whose blocks the GPU must take one after the other.
 * 1: items above 128 KiB are cut into independent 128 KiB frames written back to back (all pieces in parallel; valid for every
 * zstd decoder, ZSTD_decompressBound = item size, a few percent larger, NOT the reference's bytes). */
#define ZSTDB200_c_independentChunks 10001
ZSTDB200_API size_t     ZSTD_CCtx_setParameter(ZSTD_CCtx* cctx, int param, int value);             /* :37 */
/* Compressor.GetParameter (Compressor.cs:35-41 -> U/ZstdCompress.cs:1289): reads back the parameters listed above; the level
 * reads 3 after 0 was set (ZSTD_CLEVEL_DEFAULT), as in the reference. */
ZSTDB200_API size_t     ZSTD_CCtx_getParameter(const ZSTD_CCtx* cctx, int param, int* value);
/* Compressor.LoadDictionary (Compressor.cs:43-56 -> U/ZstdCompress.cs:1683 ZSTD_CCtx_loadDictionary): the dictionary is copied and
 * applies to every ZSTD_compress2 / ZSTDB200_compressBatch[Device] call of the context (ZSTD_compressCCtx ignores it, as in the
 * reference); NULL / 0 removes it.  A zstd-format dictionary (magic 0xEC30A437: entropy tables, repcodes, content, dictionary id
 * written into every frame header) or raw content.  As in the reference the dictionary is digested by the FIRST compression that
 * follows, at that call's level (ZSTD_initLocalDict, U/ZstdCompress.cs:1581), and the digest is kept until the dictionary is
 * replaced; a dictionary that cannot be digested makes those compressions fail with ZSTD_error_memory_allocation, the code the
 * reference reports (:1604).  Frames are byte-identical to the reference's (levels as for ZSTD_compress2). */
ZSTDB200_API size_t     ZSTD_CCtx_loadDictionary(ZSTD_CCtx* cctx, const void* dict, size_t dictSize);
/* Decompressor.SetParameter / GetParameter (Decompressor.cs:22-34 -> U/ZstdDecompress.cs:2532, 2477).  ZSTD_d_windowLogMax = 100:
 * bounds 10..31, 0 = default 27, else ZSTD_error_parameter_outOfBound; it limits the streaming decoder only, exactly as in the
 * reference (the one-shot ZSTD_decompressDCtx path never reads it).  Other parameters -> ZSTD_error_parameter_unsupported. */
ZSTDB200_API size_t     ZSTD_DCtx_setParameter(ZSTD_DCtx* dctx, int param, int value);
ZSTDB200_API size_t     ZSTD_DCtx_getParameter(const ZSTD_DCtx* dctx, int param, int* value);
/* Decompressor.LoadDictionary (Decompressor.cs:43-56 -> U/ZstdDecompress.cs:2239 ZSTD_DCtx_loadDictionary): the dictionary is
 * copied; a zstd-format dictionary (magic 0xEC30A437: entropy tables, repcodes, content, U/ZstdDecompress.cs:1770-1931) or raw
 * content.  It then applies to every ZSTD_decompressDCtx / ZSTDB200_decompressBatch call of the context; NULL / 0 removes it.
 * A corrupted dictionary is reported here (ZSTD_error_dictionary_corrupted) rather than at the first decompression.
 * With a dictionary, an item holds ONE frame (further frames in the same buffer: ZSTD_error_frameParameter_unsupported). */
ZSTDB200_API size_t     ZSTD_DCtx_loadDictionary(ZSTD_DCtx* dctx, const void* dict, size_t dictSize);
/* needed by the safe wrappers in addition (Decompressor.cs:53, ThrowHelper.cs:12-13) */
ZSTDB200_API unsigned long long ZSTD_decompressBound(const void* src, size_t srcSize);
/* compressed size of the first frame (regular or skippable) at src; used by the stream adapter to cut a concatenated
 * stream into independent items (U/ZstdDecompress.cs:958) */
ZSTDB200_API size_t     ZSTD_findFrameCompressedSize(const void* src, size_t srcSize);
ZSTDB200_API unsigned    ZSTD_isError(size_t code);
ZSTDB200_API const char* ZSTD_getErrorName(size_t code);
ZSTDB200_API unsigned    ZSTD_versionNumber(void);
ZSTDB200_API const char* ZSTD_versionString(void);

/* ---- batch frame scheduler (new; no reference counterpart) ----
 * n independent items; item i is decompressed/compressed exactly as the single-call API would, and result[i]
 * receives what that call would have returned (size or error code): a bad item never poisons the batch.
 * The function's own return value is 0, or an error code when the batch as a whole could not run
 * (no device, out of memory).  src/dst are HOST pointers; adjacent buffers (src[i]+srcSize[i]==src[i+1]) are
 * moved with one DMA per run, so callers that keep a batch contiguous in pinned memory get full PCIe rate. */
ZSTDB200_API size_t ZSTDB200_decompressBatch(ZSTD_DCtx* dctx, size_t n,
                                             const void* const* src, const size_t* srcSize,
                                             void* const* dst, const size_t* dstCapacity, size_t* result);
ZSTDB200_API size_t ZSTDB200_compressBatch(ZSTD_CCtx* cctx, size_t n, int compressionLevel,
                                           const void* const* src, const size_t* srcSize,
                                           void* const* dst, const size_t* dstCapacity, size_t* result);

/* Device-resident variants: d_src / d_dst are DEVICE pointers on the context's GPU; offsets, sizes and results are
 * host arrays.  Used when frames already live in HBM (and by bench.py's kernel-only timing). */
ZSTDB200_API size_t ZSTDB200_decompressBatchDevice(ZSTD_DCtx* dctx, size_t n,
                                                   const void* d_src, const uint64_t* srcOffset, const size_t* srcSize,
                                                   void* d_dst, const uint64_t* dstOffset, const size_t* dstCapacity,
                                                   size_t* result);
ZSTDB200_API size_t ZSTDB200_compressBatchDevice(ZSTD_CCtx* cctx, size_t n, int compressionLevel,
                                                 const void* d_src, const uint64_t* srcOffset, const size_t* srcSize,
                                                 void* d_dst, const uint64_t* dstOffset, const size_t* dstCapacity,
                                                 size_t* result);

/* ---- multi-device scheduler (new): the host scatter of the batch across the GPUs of one box ----
 * BASELINE.json north_star: "batches of independent frames are partitioned across the 8 GPUs of one box by a host scatter; no
 * NCCL is used, because frames share no state".  The reference's closest notion is many contexts used at once
 * (src/ZstdSharp.Test/ZstdNetTests.cs:498-522).  A Multi owns one decompression and one compression context per device;
 * a batch call cuts the item list into contiguous ranges balanced by byte weight (ZSTDB200_shardBounds), runs every range on
 * its device from its own host thread (bound to the CPUs local to that GPU's PCIe root unless ZSTDB200_NUMA_BIND=0), and
 * writes result[] in the caller's order.  nDevices <= 0: every visible device. */
typedef struct ZSTDB200_Multi_s ZSTDB200_Multi;
ZSTDB200_API ZSTDB200_Multi* ZSTDB200_createMulti(int nDevices);
ZSTDB200_API size_t ZSTDB200_freeMulti(ZSTDB200_Multi* m);
ZSTDB200_API int    ZSTDB200_multiDeviceCount(const ZSTDB200_Multi* m);
ZSTDB200_API size_t ZSTDB200_multiSetParameter(ZSTDB200_Multi* m, int param, int value);     /* ZSTD_CCtx_setParameter on every device's context */
ZSTDB200_API size_t ZSTDB200_multiLoadDictionary(ZSTDB200_Multi* m, const void* dict, size_t dictSize);   /* ZSTD_DCtx_loadDictionary and ZSTD_CCtx_loadDictionary on every device */
ZSTDB200_API size_t ZSTDB200_decompressBatchMulti(ZSTDB200_Multi* m, size_t n,
                                                  const void* const* src, const size_t* srcSize,
                                                  void* const* dst, const size_t* dstCapacity, size_t* result);
ZSTDB200_API size_t ZSTDB200_compressBatchMulti(ZSTDB200_Multi* m, size_t n, int compressionLevel,
                                                const void* const* src, const size_t* srcSize,
                                                void* const* dst, const size_t* dstCapacity, size_t* result);
/* bounds[0..parts]: item range [bounds[k], bounds[k+1]) goes to part k; contiguous, complete, balanced by weight */
ZSTDB200_API void   ZSTDB200_shardBounds(size_t n, const size_t* weight, int parts, size_t* bounds);
/* Pins the CALLING thread to the CPUs local to `device` (sysfs local_cpulist of its PCI function); pinned allocations made by the
 * thread afterwards are then first-touched next to the GPU.  For one-process-per-GPU callers (bench.py); 0 = done, < 0 = not possible. */
ZSTDB200_API int    ZSTDB200_bindThreadToDevice(int device);

/* ---- instrumentation ---- */
/* Milliseconds (CUDA events on the context's stream) of the last batch call: [0] host->device, [1] all kernels,
 * [2] device->host, [3..] per-kernel slots (decode: scan, setup, huf, seq, exec; encode: match, entropy). */
#define ZSTDB200_TIMING_SLOTS 12
ZSTDB200_API void     ZSTDB200_getLastTimings(const void* ctx, float* msOut /* [ZSTDB200_TIMING_SLOTS] */);
ZSTDB200_API unsigned ZSTDB200_getLastLaunchCount(const void* ctx);   /* kernels launched by the last batch call */
ZSTDB200_API const char* ZSTDB200_lastErrorString(void);              /* thread-local description of the last library-level failure */
ZSTDB200_API int      ZSTDB200_deviceCount(void);
/* Issue all work of this context on a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream), so that
 * the caller's events bracket it; NULL restores the context's own stream. ctx is a ZSTD_CCtx* or ZSTD_DCtx*. */
ZSTDB200_API size_t   ZSTDB200_setStream(void* ctx, void* cudaStream);

#ifdef __cplusplus
}
#endif
#endif /* ZSTD_B200_H */
