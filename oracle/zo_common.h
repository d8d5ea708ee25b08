/*
 * zo_common.h -- shared helpers for the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a plain-C restatement of the reference's (CHeavyarms/ZstdSharp, a C# translation of
 * zstd 1.5.1) algorithm for the hot path named in BASELINE.json: one-shot, no-dictionary frame
 * compression at levels 1..3 (ZSTD_fast / ZSTD_dfast) and full-format frame decompression.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, link or call anything under oracle/.  The product (zstdsharp_b200/csrc) never does.
 *
 * Parity pin: the reference itself cannot run in this image (no .NET).  The oracle is pinned against
 *   (a) the reference tests' known answers (frame-header descriptor bytes, error codes, size sweeps:
 *       src/ZstdSharp.Test/ZstdNetTests.cs:179-258, 456-496) and
 *   (b) system libzstd 1.5.5 -- the upstream C the reference is a mechanical translation of (at 1.5.1);
 *       byte-identical frames at levels 1..3 and identical decode results (tests/test_oracle_*.py).
 * See DESIGN.md "Oracle".
 *
 * All file:line citations are relative to /root/reference/src/ZstdSharp/Unsafe/.
 */
#ifndef ZO_COMMON_H
#define ZO_COMMON_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef uint8_t  BYTE;
typedef uint16_t U16;
typedef int16_t  S16;
typedef uint32_t U32;
typedef int32_t  S32;
typedef uint64_t U64;

/* ---- error convention: ErrorPrivate.cs:10-13, ZSTD_ErrorCode.cs:5-35 ---- */
enum {
    ZO_error_no_error = 0, ZO_error_GENERIC = 1, ZO_error_prefix_unknown = 10,
    ZO_error_version_unsupported = 12, ZO_error_frameParameter_unsupported = 14,
    ZO_error_frameParameter_windowTooLarge = 16, ZO_error_corruption_detected = 20,
    ZO_error_checksum_wrong = 22, ZO_error_dictionary_corrupted = 30, ZO_error_dictionary_wrong = 32,
    ZO_error_dictionaryCreation_failed = 34, ZO_error_parameter_unsupported = 40,
    ZO_error_parameter_outOfBound = 42, ZO_error_tableLog_tooLarge = 44,
    ZO_error_maxSymbolValue_tooLarge = 46, ZO_error_maxSymbolValue_tooSmall = 48,
    ZO_error_stage_wrong = 60, ZO_error_init_missing = 62, ZO_error_memory_allocation = 64,
    ZO_error_workSpace_tooSmall = 66, ZO_error_dstSize_tooSmall = 70, ZO_error_srcSize_wrong = 72,
    ZO_error_dstBuffer_null = 74, ZO_error_frameIndex_tooLarge = 100, ZO_error_seekableIO = 102,
    ZO_error_dstBuffer_wrong = 104, ZO_error_srcBuffer_wrong = 105, ZO_error_maxCode = 120
};
#define ERROR(name) ((size_t)-(ptrdiff_t)ZO_error_##name)
static inline unsigned ERR_isError(size_t code) { return code > ERROR(maxCode); }
#define CHECK_F(f) do { size_t const e_ = (f); if (ERR_isError(e_)) return e_; } while (0)

/* ---- little-endian unaligned memory access: Mem.cs ---- */
static inline U16 MEM_read16(const void* p) { U16 v; memcpy(&v, p, 2); return v; }
static inline U32 MEM_read32(const void* p) { U32 v; memcpy(&v, p, 4); return v; }
static inline U64 MEM_read64(const void* p) { U64 v; memcpy(&v, p, 8); return v; }
static inline void MEM_write16(void* p, U16 v) { memcpy(p, &v, 2); }
static inline void MEM_write32(void* p, U32 v) { memcpy(p, &v, 4); }
static inline void MEM_write64(void* p, U64 v) { memcpy(p, &v, 8); }
static inline U32 MEM_readLE24(const void* p) { const BYTE* b = (const BYTE*)p; return b[0] | ((U32)b[1] << 8) | ((U32)b[2] << 16); }
static inline void MEM_writeLE24(void* p, U32 v) { BYTE* b = (BYTE*)p; b[0] = (BYTE)v; b[1] = (BYTE)(v >> 8); b[2] = (BYTE)(v >> 16); }
static inline size_t MEM_readST(const void* p) { return (size_t)MEM_read64(p); }

static inline U32 BIT_highbit32(U32 v) { return 31 - (U32)__builtin_clz(v); }   /* Bitstream.cs:15-21 */

/* ---- format constants: ZstdInternal.cs:13-269, ZstdDecompressInternal.cs:9-160, Arrays.cs:8-221 ---- */
#define ZSTD_MAGICNUMBER            0xFD2FB528U
#define ZSTD_MAGIC_SKIPPABLE_START  0x184D2A50U
#define ZSTD_MAGIC_SKIPPABLE_MASK   0xFFFFFFF0U
#define ZSTD_BLOCKSIZE_MAX          (1 << 17)
#define ZSTD_blockHeaderSize        3
#define ZSTD_CONTENTSIZE_UNKNOWN    (0ULL - 1)
#define ZSTD_CONTENTSIZE_ERROR      (0ULL - 2)
#define MINMATCH   3
#define MaxLL  35
#define MaxML  52
#define MaxOff 31
#define MaxSeq 52
#define LLFSELog  9
#define MLFSELog  9
#define OffFSELog 8
#define LL_DEFAULTNORMLOG 6
#define ML_DEFAULTNORMLOG 6
#define OF_DEFAULTNORMLOG 5
#define DefaultMaxOff 28
#define HUF_TABLELOG_MAX 12
#define HUF_TABLELOG_DEFAULT 11
#define HUF_SYMBOLVALUE_MAX 255
#define LONGNBSEQ 0x7F00
#define MIN_CBLOCK_SIZE 3

typedef enum { bt_raw = 0, bt_rle = 1, bt_compressed = 2, bt_reserved = 3 } blockType_e;
typedef enum { set_basic = 0, set_rle = 1, set_compressed = 2, set_repeat = 3 } symbolEncodingType_e;
typedef enum { ZSTD_fast = 1, ZSTD_dfast = 2, ZSTD_greedy = 3, ZSTD_lazy = 4 } ZSTD_strategy;

static const U32 repStartValue[3] = { 1, 4, 8 };

static const BYTE LL_bits[MaxLL + 1] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16 };
static const S16 LL_defaultNorm[MaxLL + 1] = {
    4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1,
    2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1 };
static const BYTE ML_bits[MaxML + 1] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16 };
static const S16 ML_defaultNorm[MaxML + 1] = {
    1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1 };
static const S16 OF_defaultNorm[DefaultMaxOff + 1] = {
    1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1 };
static const U32 LL_base[MaxLL + 1] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 0x80, 0x100, 0x200, 0x400, 0x800, 0x1000,
    0x2000, 0x4000, 0x8000, 0x10000 };
static const U32 OF_base[MaxOff + 1] = {
    0, 1, 1, 5, 0xD, 0x1D, 0x3D, 0x7D, 0xFD, 0x1FD, 0x3FD, 0x7FD, 0xFFD, 0x1FFD, 0x3FFD, 0x7FFD,
    0xFFFD, 0x1FFFD, 0x3FFFD, 0x7FFFD, 0xFFFFD, 0x1FFFFD, 0x3FFFFD, 0x7FFFFD,
    0xFFFFFD, 0x1FFFFFD, 0x3FFFFFD, 0x7FFFFFD, 0xFFFFFFD, 0x1FFFFFFD, 0x3FFFFFFD, 0x7FFFFFFD };
static const BYTE OF_bits[MaxOff + 1] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31 };
static const U32 ML_base[MaxML + 1] = {
    3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18,
    19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
    35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 0x83, 0x103, 0x203, 0x403, 0x803,
    0x1003, 0x2003, 0x4003, 0x8003, 0x10003 };
static const BYTE LL_Code[64] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21,
    22, 22, 22, 22, 22, 22, 22, 22, 23, 23, 23, 23, 23, 23, 23, 23,
    24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24 };
static const BYTE ML_Code[128] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31,
    32, 32, 33, 33, 34, 34, 35, 35, 36, 36, 36, 36, 37, 37, 37, 37,
    38, 38, 38, 38, 38, 38, 38, 38, 39, 39, 39, 39, 39, 39, 39, 39,
    40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40,
    41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41,
    42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42,
    42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42 };
static const U32 rtbTable[8] = { 0, 473195, 504333, 520860, 550000, 700000, 750000, 830000 };

/* ---- XXH64 (Xxhash.cs) : only used for the optional frame checksum ---- */
static inline U64 zo_rotl64(U64 x, int r) { return (x << r) | (x >> (64 - r)); }
#define XXP1 0x9E3779B185EBCA87ULL
#define XXP2 0xC2B2AE3D27D4EB4FULL
#define XXP3 0x165667B19E3779F9ULL
#define XXP4 0x85EBCA77C2B2AE63ULL
#define XXP5 0x27D4EB2F165667C5ULL
static inline U64 zo_xxround(U64 acc, U64 in) { acc += in * XXP2; acc = zo_rotl64(acc, 31); return acc * XXP1; }
static inline U64 zo_xxmerge(U64 acc, U64 v) { v = zo_xxround(0, v); acc ^= v; return acc * XXP1 + XXP4; }
static inline U64 zo_xxh64(const void* src, size_t len, U64 seed)
{
    const BYTE* p = (const BYTE*)src; const BYTE* const end = p + len; U64 h;
    if (len >= 32) {
        const BYTE* const limit = end - 32;
        U64 v1 = seed + XXP1 + XXP2, v2 = seed + XXP2, v3 = seed, v4 = seed - XXP1;
        do { v1 = zo_xxround(v1, MEM_read64(p)); v2 = zo_xxround(v2, MEM_read64(p + 8));
             v3 = zo_xxround(v3, MEM_read64(p + 16)); v4 = zo_xxround(v4, MEM_read64(p + 24)); p += 32; } while (p <= limit);
        h = zo_rotl64(v1, 1) + zo_rotl64(v2, 7) + zo_rotl64(v3, 12) + zo_rotl64(v4, 18);
        h = zo_xxmerge(h, v1); h = zo_xxmerge(h, v2); h = zo_xxmerge(h, v3); h = zo_xxmerge(h, v4);
    } else h = seed + XXP5;
    h += (U64)len;
    while (p + 8 <= end) { h ^= zo_xxround(0, MEM_read64(p)); h = zo_rotl64(h, 27) * XXP1 + XXP4; p += 8; }
    if (p + 4 <= end) { h ^= (U64)MEM_read32(p) * XXP1; h = zo_rotl64(h, 23) * XXP2 + XXP3; p += 4; }
    while (p < end) { h ^= (*p) * XXP5; h = zo_rotl64(h, 11) * XXP1; p++; }
    h ^= h >> 33; h *= XXP2; h ^= h >> 29; h *= XXP3; h ^= h >> 32;
    return h;
}

/* zo_decode.c: entropy-header readers also used by the encode side's dictionary loader */
size_t zo_FSE_readNCount(S16* normalizedCounter, unsigned* maxSVPtr, unsigned* tableLogPtr, const void* headerBuffer, size_t hbSize);
size_t zo_HUF_readStats(BYTE* huffWeight, size_t hwSize, U32* rankStats, U32* nbSymbolsPtr, U32* tableLogPtr, const void* src, size_t srcSize);

#endif /* ZO_COMMON_H */
