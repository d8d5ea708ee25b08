// zb_common.cuh -- shared definitions for the B200-native Zstandard batch codec (product code).
//
// Hot path (BASELINE.json north_star): many independent zstd frames, decode (any level) and
// encode (level 1 byte-identical, levels 2-3 fast/dfast).  Everything here is integer/byte work.
// Reference file:line citations are relative to /root/reference/src/ZstdSharp/Unsafe/.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

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

}  // namespace zb
