// zb_common.cuh -- shared definitions for the B200-native Zstandard batch codec (product code).
//
// Hot path (BASELINE.json north_star): many independent zstd frames, decode (any level) and
// encode (level 1 byte-identical, levels 2-3 fast/dfast).  Everything here is integer/byte work.
// Reference file:line citations are relative to /root/reference/src/ZstdSharp/Unsafe/.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cuda_runtime.h>

// Debug build (make -C zstdsharp_b200/csrc debug -> _build/libzstdb200_dbg.so, selected with ZSTDB200_LIB): index assertions on the
// shared-memory tiles, staging blocks and per-item slots of every kernel.  compute-sanitizer is closed on this GPU pool, so this build,
// run over the test suite and the randomised soak, is the memory-safety evidence (profiles/r02_notes.md).  Release builds compile it out.
#ifndef ZB_DEBUG_MASK
#define ZB_DEBUG_MASK 0xFFFF
#endif
#ifdef ZB_DEBUG_ASSERTS
#define ZB_ASSERTK(k, cond) do { if (((ZB_DEBUG_MASK) >> (k)) & 1) ZB_ASSERT(cond); } while (0)
#define ZB_ASSERT(cond) do { if (!(cond)) { printf("ZB_ASSERT failed %s:%d: %s (block %d thread %d)\n", __FILE__, __LINE__, #cond, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define ZB_ASSERT(cond) do { } while (0)
#define ZB_ASSERTK(k, cond) do { } while (0)
#endif

namespace zb {

// ---- error convention: ErrorPrivate.cs:10-13, ZSTD_ErrorCode.cs:5-35 ----
enum ErrorCode : uint32_t {
    kNoError = 0, kGeneric = 1, kPrefixUnknown = 10, kVersionUnsupported = 12, kFrameParameterUnsupported = 14,
    kWindowTooLarge = 16, kCorruptionDetected = 20, kChecksumWrong = 22, kDictionaryCorrupted = 30,
    kDictionaryWrong = 32, kParameterUnsupported = 40, kParameterOutOfBound = 42, kTableLogTooLarge = 44,
    kMaxSymbolValueTooLarge = 46, kMaxSymbolValueTooSmall = 48, kStageWrong = 60, kInitMissing = 62,
    kMemoryAllocation = 64, kWorkSpaceTooSmall = 66, kDstSizeTooSmall = 70, kSrcSizeWrong = 72,
    kDstBufferNull = 74, kMaxCode = 120
};
__host__ __device__ inline uint64_t make_error(uint32_t code) { return (uint64_t)0 - (uint64_t)code; }
__host__ __device__ inline bool is_error(uint64_t v) { return v > make_error(kMaxCode); }

// ---- format constants: ZstdInternal.cs:13-269, ZstdDecompressInternal.cs:9-160 ----
constexpr uint32_t kMagic = 0xFD2FB528u;
constexpr uint32_t kMagicSkippableStart = 0x184D2A50u;
constexpr uint32_t kMagicSkippableMask = 0xFFFFFFF0u;
constexpr uint32_t kBlockSizeMax = 1u << 17;
constexpr int kMaxLL = 35, kMaxML = 52, kMaxOff = 31;
constexpr int kLLFSELog = 9, kMLFSELog = 9, kOffFSELog = 8;
constexpr int kLLDefaultNormLog = 6, kMLDefaultNormLog = 6, kOFDefaultNormLog = 5;
constexpr int kDefaultMaxOff = 28;
constexpr int kHufTableLogMax = 12;
constexpr uint32_t kLongNbSeq = 0x7F00;

// Maximum number of sequences a conformant block can hold: every match is >= 3 bytes and a block
// regenerates <= 128 KiB.  Blocks declaring more are rejected with corruption_detected (deviations).
constexpr uint32_t kSeqCap = kBlockSizeMax / 3 + 1;

__device__ __constant__ uint8_t c_LL_bits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
__device__ __constant__ uint8_t c_ML_bits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
__device__ __constant__ uint32_t c_LL_base[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 0x80, 0x100, 0x200, 0x400, 0x800, 0x1000,
    0x2000, 0x4000, 0x8000, 0x10000};
__device__ __constant__ uint32_t c_ML_base[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18,
    19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
    35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 0x83, 0x103, 0x203, 0x403, 0x803,
    0x1003, 0x2003, 0x4003, 0x8003, 0x10003};
__device__ __constant__ int16_t c_LL_defaultNorm[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1,
    2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
__device__ __constant__ int16_t c_ML_defaultNorm[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
__device__ __constant__ int16_t c_OF_defaultNorm[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
__device__ __constant__ uint8_t c_LL_Code[64] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21,
    22, 22, 22, 22, 22, 22, 22, 22, 23, 23, 23, 23, 23, 23, 23, 23,
    24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24};
__device__ __constant__ uint8_t c_ML_Code[128] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31,
    32, 32, 33, 33, 34, 34, 35, 35, 36, 36, 36, 36, 37, 37, 37, 37,
    38, 38, 38, 38, 38, 38, 38, 38, 39, 39, 39, 39, 39, 39, 39, 39,
    40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40,
    41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41,
    42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42,
    42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42};
__device__ __constant__ uint32_t c_rtbTable[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};

// ---- small device helpers ----
__device__ __forceinline__ uint32_t highbit32(uint32_t v) { return 31u - (uint32_t)__clz((int)v); }
__device__ __forceinline__ uint32_t ld_u8(const uint8_t* p) { return *p; }
__device__ __forceinline__ uint32_t ld_le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
__device__ __forceinline__ uint32_t ld_le24(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }
__device__ __forceinline__ uint32_t ld_le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
__device__ __forceinline__ uint64_t ld_le64(const uint8_t* p) { return (uint64_t)ld_le32(p) | ((uint64_t)ld_le32(p + 4) << 32); }

// ---- XXH64 (Xxhash.cs) for the optional frame checksum: ZSTD_c_checksumFlag, low 32 bits stored after the last block
// (ZstdCompress.cs:5641-5652; verified in ZstdDecompress.cs:1186-1207).  One warp hashes one buffer: all lanes fetch
// 8 bytes each (256 coalesced bytes = 8 stripes per round, next round prefetched), lanes 0..3 own the four accumulators
// and pick their words up with shuffles.  Every lane returns the hash.  `p` may have any alignment.
__device__ __forceinline__ uint64_t xxh_rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xxh_round(uint64_t acc, uint64_t in)
{ acc += in * 0xC2B2AE3D27D4EB4FULL; acc = xxh_rotl(acc, 31); return acc * 0x9E3779B185EBCA87ULL; }
__device__ __forceinline__ uint64_t xxh_merge(uint64_t acc, uint64_t v)
{ v = xxh_round(0, v); acc ^= v; return acc * 0x9E3779B185EBCA87ULL + 0x85EBCA77C2B2AE63ULL; }
__device__ __forceinline__ uint64_t xxh_ld64(const uint8_t* p)       // unaligned 8-byte read, L2-coherent (the buffer may have just been written)
{
    uintptr_t const a = (uintptr_t)p; const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3); uint32_t const sh = (uint32_t)(a & 3) * 8;
    uint32_t const w0 = __ldcg(w), w1 = __ldcg(w + 1), w2 = sh ? __ldcg(w + 2) : 0u;
    return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
}
__device__ inline uint64_t xxh64_warp(const uint8_t* p, uint32_t len, uint32_t lane)
{
    uint64_t const P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    uint32_t const FULL = 0xFFFFFFFFu;
    uint64_t h;
    uint32_t pos = 0;
    if (len >= 32) {
        uint64_t v = lane == 0 ? P1 + P2 : (lane == 1 ? P2 : (lane == 2 ? 0ull : 0ull - P1));   // seed 0
        uint32_t const nStripes = len >> 5;
        uint64_t cur = (8 * lane + 8 <= nStripes * 32) ? xxh_ld64(p + 8 * lane) : 0ull;
        for (uint32_t s0 = 0; s0 < nStripes; s0 += 8) {
            uint32_t const nextOff = (s0 + 8) * 32 + 8 * lane;
            uint64_t const nxt = (nextOff + 8 <= nStripes * 32) ? xxh_ld64(p + nextOff) : 0ull;
            uint32_t const m = min(8u, nStripes - s0);
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                uint64_t const in = __shfl_sync(FULL, cur, (4 * k + lane) & 31);       // lane a < 4 takes word 4k + a of the round
                if (k < m) v = xxh_round(v, in);
            }
            cur = nxt;
        }
        uint64_t const v1 = __shfl_sync(FULL, v, 0), v2 = __shfl_sync(FULL, v, 1), v3 = __shfl_sync(FULL, v, 2), v4 = __shfl_sync(FULL, v, 3);
        h = xxh_rotl(v1, 1) + xxh_rotl(v2, 7) + xxh_rotl(v3, 12) + xxh_rotl(v4, 18);
        h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
        pos = nStripes * 32;
    } else h = P5;
    h += (uint64_t)len;
    // tail (< 32 bytes): every lane computes it redundantly
    while (pos + 8 <= len) { h ^= xxh_round(0, xxh_ld64(p + pos)); h = xxh_rotl(h, 27) * P1 + P4; pos += 8; }
    if (pos + 4 <= len) {
        uint32_t const w = (uint32_t)__ldcg(p + pos) | ((uint32_t)__ldcg(p + pos + 1) << 8) | ((uint32_t)__ldcg(p + pos + 2) << 16) | ((uint32_t)__ldcg(p + pos + 3) << 24);
        h ^= (uint64_t)w * P1; h = xxh_rotl(h, 23) * P2 + P3; pos += 4;
    }
    while (pos < len) { h ^= (uint64_t)__ldcg(p + pos) * P5; h = xxh_rotl(h, 11) * P1; pos++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

}  // namespace zb
