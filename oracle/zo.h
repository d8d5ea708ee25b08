/*
 * zo.h -- public surface of the CPU ORACLE (test infrastructure only; see zo_common.h header).
 *
 * zo_compress / zo_decompress restate Compressor.Wrap / Decompressor.Unwrap of the reference
 * (src/ZstdSharp/Compressor.cs:78-96 -> Unsafe/ZstdCompress.cs:7138 ZSTD_compress2;
 *  src/ZstdSharp/Decompressor.cs:62-88 -> Unsafe/ZstdDecompress.cs:1365 ZSTD_decompressDCtx)
 * for the no-dictionary, default-parameter path.  Return values follow the zstd size_t convention
 * (error <=> value > (size_t)-120, Unsafe/ErrorPrivate.cs:10-13).
 */
#ifndef ZO_H
#define ZO_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

unsigned     zo_isError(size_t code);
int          zo_getErrorCode(size_t code);
const char*  zo_getErrorName(size_t code);

/* Unsafe/ZstdCompress.cs:19-22 */
size_t zo_compressBound(size_t srcSize);

/* level->cParams resolution for a one-shot frame of srcSize bytes
 * (Unsafe/ZstdCompress.cs:7891 ZSTD_getCParams_internal + :2023 ZSTD_adjustCParams_internal).
 * out[7] = {windowLog, chainLog, hashLog, searchLog, minMatch, targetLength, strategy}. */
void zo_getCParams(int level, size_t srcSize, unsigned out[7]);

/* One-shot frame compression, levels 1..3 (0 -> 3), contentSizeFlag=1, no checksum unless
 * checksumFlag!=0, no dictionary.  Multi-block frames (srcSize > 128 KiB) are supported. */
size_t zo_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level);
size_t zo_compress_advanced(void* dst, size_t dstCapacity, const void* src, size_t srcSize,
                            int level, int checksumFlag);

/* Context-reusing variants (one context per thread, as the reference's tests do: ZstdNetTests.cs:498-522). */
/* Compressor.LoadDictionary(dict) + Wrap(src): one frame by the CDict the first ZSTD_compress2 builds from the loaded bytes */
size_t zo_compress_usingLoadedDict(void* dst, size_t dstCapacity, const void* src, size_t srcSize,
                                   const void* dict, size_t dictSize, int level, int checksumFlag);
void*  zo_createCCtx(void);
void   zo_freeCCtx(void* cctx);
size_t zo_compressCCtx(void* cctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level, int checksumFlag);
void*  zo_createDCtx(void);
void   zo_freeDCtx(void* dctx);
size_t zo_decompressDCtx(void* dctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize);

/* Multi-frame decompression incl. skippable frames and checksum verification. */
size_t             zo_decompress(void* dst, size_t dstCapacity, const void* src, size_t srcSize);
/* Decompressor.LoadDictionary + Unwrap (Decompressor.cs:43-56 -> Unsafe/ZstdDecompress.cs:1770-1931): raw-content or
 * zstd-format dictionary (magic 0xEC30A437: entropy tables + repcodes + content), applied to every frame of src. */
size_t             zo_decompress_usingDict(void* dst, size_t dstCapacity, const void* src, size_t srcSize, const void* dict, size_t dictSize);
unsigned long long zo_decompressBound(const void* src, size_t srcSize);
size_t             zo_findFrameCompressedSize(const void* src, size_t srcSize);

/* ---- stage-level taps (used by tests to localise GPU/oracle differences) ---- */
typedef struct { uint32_t offset; uint16_t litLength; uint16_t matchLength; } zo_seqDef;  /* Unsafe/seqDef_s.cs */

/* Runs the level's match finder on ONE block (srcSize <= 128 KiB, first block of a frame).
 * seqs: capacity >= srcSize/3+1; lits: capacity >= srcSize.  longLength[0]=type(0 none,1 lit,2 match),
 * longLength[1]=position.  repOut = rep codes after the block. Returns number of sequences. */
size_t zo_matchfinder_block(int level, const void* src, size_t srcSize,
                            zo_seqDef* seqs, uint8_t* lits, size_t* litSize,
                            uint32_t longLength[2], uint32_t repOut[3]);

/* Decodes the FIRST block of the FIRST frame in src down to its literals and (ll, ml, offset) triples.
 * triples: 3*u32 per sequence, capacity in sequences = seqCapacity. Returns nbSeq or an error code. */
size_t zo_decode_first_block_stages(const void* src, size_t srcSize,
                                    uint8_t* lits, size_t litCapacity, size_t* litSize,
                                    uint32_t* triples, size_t seqCapacity);

#ifdef __cplusplus
}
#endif
#endif
