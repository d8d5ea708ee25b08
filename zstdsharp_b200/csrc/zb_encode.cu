// zb_encode.cu -- batch Zstandard frame encoder for sm_100a: ZSTD_fast / ZSTD_dfast levels (-131072..-1, 0..3, 4 where it is dfast), with or without a loaded dictionary (product code; no CPU fallback).
//
// Replaces, for batches of independent inputs (one frame each: one block up to 128 KiB, a multi-block frame above), the reference's
//   ZSTD_compress2 -> ZSTD_compressEnd -> ZSTD_compress_frameChunk -> ZSTD_compressBlock_internal   (ZstdCompress.cs:7138,5665,4690,4528)
//     -> ZSTD_buildSeqStore -> ZSTD_compressBlock_fast / ZSTD_compressBlock_doubleFast               (:3432, ZstdFast.cs:96, ZstdDoubleFast.cs:51)
//     -> ZSTD_entropyCompressSeqStore: ZSTD_compressLiterals (HIST_count, HUF_buildCTable, HUF_writeCTable,
//        HUF_compress{1,4}X_usingCTable), ZSTD_buildSequencesStatistics (FSE_normalizeCount, FSE_writeNCount,
//        FSE_buildCTable), ZSTD_encodeSequences                                                     (:3357; HufCompress.cs; FseCompress.cs; ZstdCompressSequences.cs)
// The output of every input is byte-identical to what the reference's Compressor.Wrap produces for it.
// Kernels (MB = false: every frame is one block; MB = true: wave b = block b of every frame that has one, with the window,
// repcodes and Huffman table the blocks before it left behind):
//   enc_match_group_kernel<16, MB>        16 lanes / block   exact ZSTD_fast parse (levels 1-2), speculative window of 8 reference iterations
//   enc_match_dfast_group_kernel<16, MB>  16 lanes / block   exact ZSTD_dfast parse (level 3), 15 probe positions + 1 look-ahead per window
//   enc_entropy_kernel<MB, 0>             CTA  / block       literal gather + histograms, Huffman table (new / repeated), parallel bit scatter of
//                                                            the 4 Huffman streams, sequence codes + histograms, the three FSE tables and their descriptions
//   enc_fse_chain_kernel<MB>              lane / FSE chain   the three FSE state chains of every block (3 x blocks lanes)
//   enc_entropy_kernel<MB, 1>             CTA  / block       parallel bit scatter of the sequence bitstream, block/frame assembly with the
//                                                            reference's accept/reject gates (raw / RLE / compressed), state confirmation, XXH64 trailer
//   enc_compact_kernel                    CTA  / frame       packs the frames densely before D2H
// With a loaded dictionary (Compressor.LoadDictionary; ZstdCompress.cs:1581-1700, 2725-2900, 5126-5500, 5826-6010):
//   enc_dict_build_kernel                 thread             the CDict on the device, once per dictionary and level: compression forms of the entropy tables, repeat
//                                                            modes, match-finder tables filled over the content (dtlm_full)
//   enc_dict_init_kernel                  grid over frames   per pass: the dictionary's Huffman table into every frame's `prev` slot, the CDict's tables into frames that copy it
//   enc_match_dict_fast_group_kernel<16>  16 lanes / frame   ZSTD_fast with a dictionary (dictMatchState / extDict), speculative window over the classic loop; block bookkeeping
//   enc_match_dict_dfast_group_kernel<16> 16 lanes / frame   ZSTD_dfast with a dictionary, two tables, long-table look-up at ip + 1 by the winning position
//   enc_match_dict_kernel                 warp / frame       serial restatement of the six variants: blocks behind an invalidated dictionary; cross-check of the two above
// The entropy kernels start such frames from the dictionary's state (Huffman valid / check, set_repeat FSE tables, repcodes, dictionary id in the header).
// All hash tables live in HBM/L2 (zeroed per frame) and every block of a wave is in flight at once: the parse is a chain of
// dependent memory round trips per sequence, and only concurrency across frames hides it (profiles/r01_notes.md).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "zb_encode.cuh"
#include "zb_decode.cuh"      // the dictionary header is parsed by the decode side's dec_dict_kernel

namespace zb {

// ------------------------------------------------------------------------------------------------------------
//  Parameters: Clevels.cs:8 rows 0..3 of the four size tables, ZSTD_getCParams_internal (ZstdCompress.cs:7891),
//  ZSTD_adjustCParams_internal (:2023).  Host-side integer logic, evaluated once per chunk.
// ------------------------------------------------------------------------------------------------------------
struct CParams { uint32_t windowLog, chainLog, hashLog, searchLog, minMatch, targetLength, strategy; };
static const CParams kDefaultCParams[4][5] = {     // rows 0..4 (row 0 = base of the negative levels); strategy 0 = not a fast/dfast row (level 4 is ZSTD_greedy there)
    {{19, 12, 13, 1, 6, 1, 1}, {19, 13, 14, 1, 7, 0, 1}, {20, 15, 16, 1, 6, 0, 1}, {21, 16, 17, 1, 5, 0, 2}, {21, 18, 18, 1, 5, 0, 2}},
    {{18, 12, 13, 1, 5, 1, 1}, {18, 13, 14, 1, 6, 0, 1}, {18, 14, 14, 1, 5, 0, 2}, {18, 16, 16, 1, 4, 0, 2}, {18, 16, 17, 3, 5, 2, 0}},
    {{17, 12, 12, 1, 5, 1, 1}, {17, 12, 13, 1, 6, 0, 1}, {17, 13, 15, 1, 5, 0, 1}, {17, 15, 16, 2, 5, 0, 2}, {17, 17, 17, 2, 4, 0, 2}},
    {{14, 12, 13, 1, 5, 1, 1}, {14, 14, 15, 1, 5, 0, 1}, {14, 14, 15, 1, 4, 0, 1}, {14, 14, 15, 2, 4, 0, 2}, {14, 14, 14, 4, 4, 2, 0}},
};
static uint32_t h_highbit(uint32_t v) { return 31 - (uint32_t)__builtin_clz(v); }
static CParams adjust_cparams(CParams c, uint64_t srcSize)
{
    if (srcSize < (1ull << 30)) {
        uint32_t const t = (uint32_t)srcSize;
        uint32_t const srcLog = t < 64 ? 6 : h_highbit(t - 1) + 1;
        if (c.windowLog > srcLog) c.windowLog = srcLog;
    }
    if (c.hashLog > c.windowLog + 1) c.hashLog = c.windowLog + 1;
    if (c.chainLog > c.windowLog) c.chainLog -= (c.chainLog - c.windowLog);
    if (c.windowLog < 10) c.windowLog = 10;
    return c;
}
static CParams get_cparams(int level, uint64_t srcSize)
{
    uint32_t const tableID = (srcSize <= 256 * 1024) + (srcSize <= 128 * 1024) + (srcSize <= 16 * 1024);
    int const row = level == 0 ? 3 : (level < 0 ? 0 : level);      // negative levels: row 0 + acceleration (ZstdCompress.cs:7901-7923)
    CParams c = kDefaultCParams[tableID][row];
    if (c.strategy == 0) return c;                                  // caller reports parameter_unsupported
    if (level < 0) c.targetLength = (uint32_t)(-(level < -(1 << 17) ? -(1 << 17) : level));     // ZSTD_minCLevel() = -(1 << 17)
    return adjust_cparams(adjust_cparams(c, srcSize), srcSize);
}

// With a dictionary (ZSTD_getCParamsFromCCtxParams :2156 = ZSTD_getCParams_internal :7891 + ZSTD_adjustCParams_internal :2023 twice;
// row size :7852, ZSTD_dictAndWindowLog :1985).  mode: 0 noAttachDict, 1 attachDict, 2 createCDict.
constexpr uint64_t kSrcSizeUnknown = ~0ull;
static uint32_t dict_and_window_log(uint32_t windowLog, uint64_t srcSize, uint64_t dictSize)
{
    if (dictSize == 0) return windowLog;
    uint64_t const windowSize = 1ull << windowLog, dictAndWindowSize = dictSize + windowSize;
    if (windowSize >= dictSize + srcSize) return windowLog;
    if (dictAndWindowSize >= (1ull << 31)) return 31;
    return h_highbit((uint32_t)dictAndWindowSize - 1) + 1;
}
static CParams adjust_cparams_dict(CParams c, uint64_t srcSize, uint64_t dictSize, int mode)
{
    if (mode == 2) { if (dictSize && srcSize == kSrcSizeUnknown) srcSize = 513; }
    else if (mode == 1) dictSize = 0;
    if (srcSize < (1ull << 30) && dictSize < (1ull << 30)) {
        uint32_t const t = (uint32_t)(srcSize + dictSize);
        uint32_t const srcLog = t < 64 ? 6 : h_highbit(t - 1) + 1;
        if (c.windowLog > srcLog) c.windowLog = srcLog;
    }
    if (srcSize != kSrcSizeUnknown) {
        uint32_t const dawl = dict_and_window_log(c.windowLog, srcSize, dictSize);
        if (c.hashLog > dawl + 1) c.hashLog = dawl + 1;
        if (c.chainLog > dawl) c.chainLog -= (c.chainLog - dawl);
    }
    if (c.windowLog < 10) c.windowLog = 10;
    return c;
}
static CParams get_cparams_dict(int level, uint64_t srcSizeHint, uint64_t dictSize, int mode)
{
    uint64_t const rowDict = mode == 1 ? 0 : dictSize;
    bool const unknown = srcSizeHint == kSrcSizeUnknown;
    uint64_t const rSize = (unknown && rowDict == 0) ? kSrcSizeUnknown : srcSizeHint + rowDict + ((unknown && rowDict > 0) ? 500 : 0);   // wraps for `unknown` with a dictionary, as in the reference
    uint32_t const tableID = (rSize <= 256 * 1024) + (rSize <= 128 * 1024) + (rSize <= 16 * 1024);
    int const row = level == 0 ? 3 : (level < 0 ? 0 : level);
    CParams c = kDefaultCParams[tableID][row];
    if (c.strategy == 0) return c;
    if (level < 0) c.targetLength = (uint32_t)(-(level < -(1 << 17) ? -(1 << 17) : level));
    return adjust_cparams_dict(adjust_cparams_dict(c, srcSizeHint, dictSize, mode), srcSizeHint, dictSize, mode);
}

// ------------------------------------------------------------------------------------------------------------
//  Device-side layout
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t kEncSeqCap = kBlockSizeMax / 4 + 1;      // maxNbSeq = blockSize / 4 (minMatch != 3), ZstdCompress.cs:2570
constexpr uint32_t kEncLitStride = kBlockSizeMax + 64;
constexpr size_t kEncMaxFrameBytes = 0x7FFF0000u;           // positions inside a frame are 31-bit integers in the match kernels
// Warps per CTA of the group match kernels.  The warps are independent; four per CTA keep the kernel (one warp per two blocks,
// ~28 warps per SM for 8192 blocks) from holding 28 of an SM's 32 CTA slots while the entropy kernels of other sub-batches
// share the SMs with it in the pipelined host path (enc_set_overlap_mode).  No effect on the kernel's own time (42.5 ms either way).
constexpr uint32_t kMatchWarps = 4;

struct __align__(16) EncItem {
    uint64_t srcOff, dstOff;
    uint32_t srcSize, dstCap;      // whole frame
    uint32_t windowLog, hashLog, chainLog, minMatch, strategy;
    uint32_t tableOff;      // offset (in u32 entries) of this frame's hash table(s) inside the table arena
    uint32_t nbSeq, lastLL; // current block
    // ---- state carried from block to block of a multi-block frame (ZSTD_compressedBlockState_t, ZstdCompressInternal.cs) ----
    uint32_t rep[2];        // prevCBlock->rep[0..1]: repcodes confirmed by the last compressed block (rep[2] stays 8: fast/dfast never touch it)
    uint32_t repNext[2];    // nextCBlock->rep[0..1]: repcodes after the current block's parse, confirmed by the entropy stage
    uint32_t outPos;        // bytes of the frame written so far
    uint32_t hufRepeat;     // prevCBlock->entropy.huf.repeatMode: 0 none, 1 check (valid needs a dictionary)
    uint32_t hufCur;        // which of the frame's two Huffman table slots holds prevCBlock's table
    uint32_t stepSize;      // ZSTD_fast: distance between probe pairs (2, or targetLength + 1 for the negative levels)
    uint32_t rawLits;       // literal compression disabled (negative levels)
    // ---- frames compressed with a loaded dictionary (enc_match_dict_kernel); all 0 otherwise ----
    uint32_t dMode;         // 0 no dictionary, 1 the CDict is attached (ZSTD_resetCCtx_byAttachingCDict :2746), 2 copied (:2803)
    uint32_t dPrefix;       // the reference's index of src[0]: 2 + dictionary content length
    uint32_t dStep;         // stepSize of the dictionary variants: targetLength + !targetLength (ZstdFast.cs:395, :588)
    uint32_t wLow, wDictLimit, loadedDictEnd, dms;   // ms->window.lowLimit / .dictLimit, ms->loadedDictEnd, ms->dictMatchState != NULL: advanced block by block
    uint32_t fseValid;      // prevCBlock->entropy.fse.*_repeatMode == FSE_repeat_valid: 1 LL, 2 OF, 4 ML (only a dictionary makes them valid)
    uint32_t fseValidNext;  // nextCBlock's, confirmed with the block
    uint32_t dBlkMode, dBlkLow, dBlkPrefix;   // the current block's variant (0 none, 1 dictMatchState, 2 extDict; | kDictSerial) and its dictStartIndex / prefixStartIndex (dict_block_mode)
    uint32_t _pad[3];
};
static_assert(sizeof(EncItem) % 16 == 0, "EncItem is copied and indexed as 16-byte units");

// Geometry of block `wave` of a frame (ZSTD_compress_frameChunk :4690 + ZSTD_window_enforceMaxDist, ZstdCompressInternal.cs:630;
// ZSTD_getLowestPrefixIndex :802).  Positions are byte offsets from the frame start; the reference's index of a position is +2.
struct BlkGeom { int B, end, lowPos, ip0; uint32_t maxRep; };
__device__ __forceinline__ BlkGeom blk_geom(uint32_t frameSize, uint32_t windowLog, uint32_t wave)
{
    BlkGeom g;
    uint32_t const B = wave * kBlockSizeMax, size = min(kBlockSizeMax, frameSize - B), end = B + size;
    uint32_t const maxDist = 1u << windowLog;
    uint32_t const startIdx = B + 2, endIdx = end + 2;
    uint32_t const dictLimit = startIdx > maxDist ? max(2u, startIdx - maxDist) : 2u;      // lowLimit == dictLimit without a dictionary
    uint32_t const prefixStartIndex = (endIdx - dictLimit > maxDist) ? endIdx - maxDist : dictLimit;
    g.B = (int)B; g.end = (int)end; g.lowPos = (int)prefixStartIndex - 2;
    g.ip0 = (int)B + (g.B == g.lowPos);                                                     // ip0 += (ip0 == prefixStart)
    uint32_t const curr = (uint32_t)g.ip0 + 2;
    uint32_t const windowLow = (curr - dictLimit > maxDist) ? curr - maxDist : dictLimit;
    g.maxRep = curr - windowLow;
    return g;
}

struct FseGTable; struct EntCarry;
// The digested dictionary of a pass (ZSTD_CDict_s: content, match-finder tables filled with dtlm_full, entropy tables); content == nullptr: none
struct EncDictDev {
    const uint8_t* content; uint32_t contentLen; uint32_t dictID;
    const uint32_t* tables;         // ZSTD_fast: hashTable[1 << hashLog]; ZSTD_dfast: hashLong[1 << hashLog] | hashSmall[1 << chainLog]
    uint32_t hashLog, chainLog;
    const uint8_t* huf;             // {u8 nbBits[256]; u16 value[256]}: the dictionary's Huffman CTable (HUF_readCTable)
    const FseGTable* fse;           // [3] LL, OF, ML
};
struct EncPass {
    EncItem* items; uint32_t nItems;
    const uint8_t* src; uint8_t* dst;
    uint32_t* tables;       // zero-initialised hash tables (one region per chunk)
    uint32_t* seqLL; uint32_t* seqML; uint32_t* seqOF;   // litLength, matchLength-3, offCode+1 (seqDef_s.cs)
    uint8_t* litBuf;        // gathered literals
    uint64_t* stateBits;    // per sequence: FSE state bits of OF | ML | LL (13 bits each: 9 value + 4 count)
    uint64_t* results;
    FseGTable* fseTabs;     // [item][3]: LL, OF, ML compression tables of the current block
    EntCarry* carry;        // [item]
    uint8_t* hufState;      // multi-block frames: 2 slots of {u8 nbBits[256]; u16 value[256]} per frame (prev / next Huffman CTable)
    uint32_t checksumFlag;  // ZSTD_c_checksumFlag: append the low 32 bits of XXH64(src) to every frame
    uint32_t seqStride;     // entries per frame in seqLL / seqML / seqOF / stateBits: largest block of the pass / 4 + 1 (maxNbSeq, ZstdCompress.cs:2570), at most kEncSeqCap
    uint32_t litStride;     // bytes per frame in litBuf: largest block of the pass + 64
    EncDictDev dict;
};
constexpr uint32_t kHufStateSlot = 768;

__device__ __forceinline__ uint32_t rd32(const uint8_t* p)
{
    uintptr_t const a = (uintptr_t)p; const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3); uint32_t const sh = (uint32_t)(a & 3) * 8;
    uint32_t const lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}
__device__ __forceinline__ uint64_t rd64(const uint8_t* p)
{
    uintptr_t const a = (uintptr_t)p; const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3); uint32_t const sh = (uint32_t)(a & 3) * 8;
    uint32_t const w0 = w[0], w1 = w[1];
    if (sh == 0) return (uint64_t)w0 | ((uint64_t)w1 << 32);
    uint32_t const w2 = w[2];
    return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
}

// ZSTD_hashPtr on 8 loaded bytes (ZstdCompressInternal.cs:340-437)
__device__ __forceinline__ uint32_t hash_val(uint64_t x, uint32_t hBits, uint32_t mls)
{
    switch (mls) {
    default:
    case 4: return ((uint32_t)x * 2654435761u) >> (32 - hBits);
    case 5: return (uint32_t)(((x << 24) * 889523592379ull) >> (64 - hBits));
    case 6: return (uint32_t)(((x << 16) * 227718039650203ull) >> (64 - hBits));
    case 7: return (uint32_t)(((x << 8) * 58295818150454627ull) >> (64 - hBits));
    case 8: return (uint32_t)((x * 0xCF1BBCDCB7A56463ull) >> (64 - hBits));
    }
}

// ------------------------------------------------------------------------------------------------------------
//  Group-per-chunk, exact ZSTD_fast parse: GS lanes (8 / 16 / 32) own one chunk, a warp carries 32/GS chunks.
//  The reference loop (ZstdFast.cs:147-230) visits positions ip0, ip0+1, ip0+d, ip0+d+1, ... on a schedule that depends only
//  on (ip0, step, nextStep) until a match is found, so a window of GS/2 iterations is evaluated speculatively: lane 2k / 2k+1
//  own the two probes of iteration k, a probe sees earlier lanes of the window through __match_any_sync and older positions
//  through the table, the first event in the reference's order (repcode@ip2, hash@ip0, hash@ip1 of the lowest iteration) wins,
//  and table writes are committed up to that event only -- the table always equals the serial algorithm's table.
//  The kernel is a warp-uniform state machine: every group walks through the same phases in
//  every round and a group that has nothing to do in a phase is predicated off, so the groups of a warp share one
//  instruction stream.  A 32-lane window wastes ~28 of its 32 probes on text (the first event sits within the first few
//  probes): ncu showed 33 G warp instructions and 120 GB of DRAM reads per GiB; 8 lanes per chunk cut both.
// ------------------------------------------------------------------------------------------------------------
// Single-block frames (MB = false) never see a position above 2^17, so a table entry has room for a 14-bit tag of the four bytes
// that sit at the position: entry = (position + 2) | tag << 18.  A probe whose tag differs cannot match (ZstdFast.cs:179-192 compares
// MEM_read32(base + idx) with MEM_read32(ip)), so its candidate bytes are never fetched: 15 of the 16 candidate sectors of a window
// stay in DRAM, and only likely hits pay the second dependent round trip.  The table still holds exactly the reference's positions.
constexpr uint32_t kPosBits = 18, kPosMask = (1u << kPosBits) - 1u;
__device__ __forceinline__ uint32_t tag4(uint32_t v) { return (v * 2654435761u) >> kPosBits; }
template <bool TAG> __device__ __forceinline__ uint32_t tab_entry(uint32_t posPlus2, uint32_t first4) { return TAG ? (posPlus2 | (tag4(first4) << kPosBits)) : posPlus2; }

template <int GS, bool MB>
__global__ void __launch_bounds__(32 * kMatchWarps) enc_match_group_kernel(EncPass p, const uint32_t* __restrict__ workList, uint32_t nWork, uint32_t wave)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr int NG = 32 / GS, NIT = GS / 2;
    constexpr uint32_t LOW = GS == 32 ? FULL : ((1u << (GS & 31)) - 1u);
    uint32_t const lane = threadIdx.x & 31, g = lane / GS, l = lane % GS, gbase = g * GS;
    uint32_t const gmask = LOW << gbase;
    uint32_t const wi = (blockIdx.x * kMatchWarps + (threadIdx.x >> 5)) * NG + g;
    bool active = wi < nWork;
    uint32_t const item = workList[active ? wi : 0];
    EncItem& it = p.items[item];
    uint32_t const hlog = it.hashLog, mls = it.minMatch;
    const uint8_t* const src = p.src + it.srcOff;    // frame start; every position below is an offset from it
    uint32_t* const T = p.tables + it.tableOff;      // zero-initialised per frame, HBM/L2-resident; entries hold position + 2 (the reference's index)
    uint32_t* const oLL = p.seqLL + (size_t)item * p.seqStride;
    uint32_t* const oML = p.seqML + (size_t)item * p.seqStride;
    uint32_t* const oOF = p.seqOF + (size_t)item * p.seqStride;
    // block geometry: first block of a frame unless MB (then block `wave`, with the window and repcodes of the blocks before it)
    int srcSize = active ? (int)it.srcSize : 64;     // end of the block
    int ip0 = 1, anchor = 0, lowPos = 0;             // first position of a frame is skipped (:129)
    uint32_t rep1 = 1, rep2 = 0, offsetSaved = 4;    // rep2 = 4 exceeds the history at frame start (:131-145)
    if (MB && active) {
        BlkGeom const bg = blk_geom(it.srcSize, it.windowLog, wave);
        srcSize = bg.end; ip0 = bg.ip0; anchor = bg.B; lowPos = bg.lowPos;
        rep1 = it.rep[0]; rep2 = it.rep[1]; offsetSaved = 0;
        if (rep2 > bg.maxRep) { offsetSaved = rep2; rep2 = 0; }
        if (rep1 > bg.maxRep) { offsetSaved = rep1; rep1 = 0; }
    }
    int const ilimit = srcSize - 8;
    uint32_t nseq = 0;
    int const stepSize = active ? (int)it.stepSize : 2;
    int step = stepSize, nextStep = ip0 + 128, d = stepSize;       // _start
    bool afterMatch = false;                         // the greedy rep2 loop (:264-285) is still open at ip0
    uint32_t const k = l >> 1, odd = l & 1;
    auto gballot = [&](bool pr) -> uint32_t { return (__ballot_sync(FULL, pr) >> gbase) & LOW; };
    while (__any_sync(FULL, active)) {
        // ---- positions of this lane's probe: iteration k of the window (ZstdFast.cs:147-230 schedule) ----
        int P = ip0, D = d, S = step, N = nextStep, pk = 0, dk = 0;
        // right after a match (and for the first 100 bytes without one) the schedule is simply consecutive pairs: no step change inside the window
        bool const simple = !active || (d == 2 && step == 2 && ip0 + 2 * NIT + 2 < nextStep);
        if (__all_sync(FULL, simple)) { pk = ip0 + 2 * (int)k; dk = 2; P = ip0 + 2 * NIT; }
        else {
#pragma unroll
            for (int j = 0; j < NIT; j++) { if (j == (int)k) { pk = P; dk = D; } P += D; int const ip2n = P + S; D = S; if (ip2n >= N) { S++; N += 128; } }
        }
        bool const vk = active && (pk + dk + 1 < ilimit);          // loop condition ip3 < ilimit for this iteration
        uint32_t const validMask = gballot(vk);
        // ---- open rep2 loop first (:264-285): same answer in every lane of the group ----
        bool const r2 = active && afterMatch && ip0 <= ilimit && rep2 > 0;
        bool const r2hit = r2 && rd32(src + ip0) == rd32(src + ip0 - (int)rep2);
        // ---- probes ----
        int const q = pk + (int)odd;
        uint64_t const x = vk ? rd64(src + q) : 0ull;
        uint32_t const cur4 = (uint32_t)x;
        bool repHit = false;
        if (vk && !odd && rep1) { int const r = pk + dk; repHit = rd32(src + r) == rd32(src + r - (int)rep1); }
        uint32_t const h = hash_val(x, hlog, mls);
        uint32_t const tv = vk ? __ldcg(T + h) : 0u;
        constexpr bool TAG = !MB;
        uint32_t const peers = (__match_any_sync(FULL, vk ? (h | (g << 24)) : (0x80000000u | lane)) >> gbase) & LOW;
        uint32_t const lower = peers & ((1u << l) - 1u);
        int const cl = lower ? 31 - __clz((int)lower) : (int)l;
        int const cq = __shfl_sync(FULL, q, gbase + cl);
        uint32_t const cq4 = __shfl_sync(FULL, cur4, gbase + cl);  // a candidate forwarded from a lower lane of the window: its bytes are in that lane's registers
        int const cand = lower ? cq : (int)(TAG ? (tv & kPosMask) : tv) - 2;                  // table stores position + 2, 0 = empty
        bool const maybe = vk && !lower && cand >= lowPos && (!TAG || (tv >> kPosBits) == tag4(cur4));   // idx >= prefixStartIndex
        uint32_t const c4 = maybe ? rd32(src + cand) : 0u;
        if (maybe) asm volatile("prefetch.global.L1 [%0];" :: "l"(src + cand + 32));     // the match extension reads on from here (39.8 -> 39.2 ms; a fault-free hint)
        bool const hit = lower ? (vk && cq4 == cur4) : (maybe && c4 == cur4);
        uint32_t key = 0xFFFFFFFFu;                                 // 0: rep2 at ip0; 1 + 3k: repcode at ip2; 2 + 3k / 3 + 3k: hash hit at ip0 / ip1
        if (hit) key = 3 * k + 2 + odd;
        if (repHit) key = 3 * k + 1;
        if (r2hit) key = 0;
        uint32_t const best = __reduce_min_sync(gmask, key);
        bool const ev = active && best != 0xFFFFFFFFu;
        int const type = !ev ? -1 : (best == 0 ? 3 : (int)((best - 1) % 3));
        uint32_t const ke = (ev && best) ? (best - 1) / 3 : 0u;
        if (active && type != 3) afterMatch = false;
        // ---- table writes that precede the event; the latest position of a bucket wins ----
        {
            uint32_t const lastLane = type < 0 ? (uint32_t)GS - 1 : 2 * ke + 1;
            uint32_t const peersC = peers & (lastLane >= 31 ? FULL : ((2u << lastLane) - 1u));
            if (vk && type != 3 && l <= lastLane && ((peersC >> l) >> 1) == 0) T[h] = tab_entry<TAG>((uint32_t)q + 2, cur4);
        }
        if (type == 3 && l == 0) { uint64_t const xi = rd64(src + ip0); T[hash_val(xi, hlog, mls)] = tab_entry<TAG>((uint32_t)ip0 + 2, (uint32_t)xi); }     // :278
        // ---- match geometry (group-uniform) ----
        int const pke = __shfl_sync(FULL, pk, gbase + 2 * ke), dke = __shfl_sync(FULL, dk, gbase + 2 * ke);
        uint32_t const evLane = 2 * ke + (type == 2 ? 1u : 0u);
        int const qe = __shfl_sync(FULL, q, gbase + evLane), ce = __shfl_sync(FULL, cand, gbase + evLane);
        int mpos = 0, msrc = 0, mlen = 0, current0 = 0; uint32_t offcode = 0;
        if (type == 3) { mpos = ip0; msrc = ip0 - (int)rep2; mlen = 4; uint32_t const t = rep2; rep2 = rep1; rep1 = t; }
        else if (type == 0) {
            mpos = pke + dke; msrc = mpos - (int)rep1;
            int const back = src[mpos - 1] == src[msrc - 1];        // one byte, no anchor test (:171)
            mpos -= back; msrc -= back; mlen = 4 + back; current0 = pke;
        } else if (type > 0) { mpos = qe; msrc = ce; rep2 = rep1; rep1 = (uint32_t)(qe - ce); offcode = rep1 + 2; mlen = 4; current0 = qe; }
        // backward extension of hash matches (:236-247): lane l looks l+1 bytes back, GS bytes per round
        {
            bool ext = type == 1 || type == 2;
            while (__any_sync(FULL, ext)) {
                int const a = mpos - 1 - (int)l, b = msrc - 1 - (int)l;
                bool const ok = ext && a >= anchor && b >= lowPos && src[a] == src[b];     // ip0 > anchor and match0 > prefixStart before every step
                uint32_t const okm = gballot(ok);
                uint32_t const n = okm == LOW ? (uint32_t)GS : (uint32_t)__ffs((int)~okm) - 1u;
                if (ext) { mpos -= (int)n; msrc -= (int)n; mlen += (int)n; ext = n == (uint32_t)GS; }
            }
        }
        // forward extension (ZSTD_count :264): lane l compares 4 bytes, 4*GS bytes per round
        {
            bool cnt = ev;
            while (__any_sync(FULL, cnt)) {
                int const pa = mpos + mlen + 4 * (int)l, pb = msrc + mlen + 4 * (int)l;
                int const rem = srcSize - pa;
                uint32_t n = 4;
                if (cnt) {
                    if (rem >= 4) { uint32_t const diff = rd32(src + pa) ^ rd32(src + pb); n = diff ? (uint32_t)(__ffs((int)diff) - 1) >> 3 : 4u; }
                    else { n = 0; for (int j = 0; j < rem; j++) { if (src[pa + j] == src[pb + j]) n++; else break; } }
                }
                uint32_t const notFull = gballot(cnt && n != 4);
                uint32_t const f = notFull ? (uint32_t)__ffs((int)notFull) - 1u : 0u;
                uint32_t const nf = __shfl_sync(FULL, n, gbase + f);
                if (cnt) { if (notFull) { mlen += 4 * (int)f + (int)nf; cnt = false; } else mlen += 4 * GS; }
            }
        }
        // ---- sequence, post-match inserts (:251-263), next state ----
        if (ev) {
            ZB_ASSERT(nseq < p.seqStride && mpos >= anchor && msrc >= lowPos && msrc < mpos && mpos + mlen <= srcSize);
            if (l == 0) { oLL[nseq] = (uint32_t)(mpos - anchor); oOF[nseq] = offcode + 1; oML[nseq] = (uint32_t)mlen - 3; }
            nseq++;
            int const mend = mpos + mlen;
            if (l == 0) {
                if (type == 2 && pke + dke < mend) { int const pp = pke + dke; uint64_t const xp = rd64(src + pp); T[hash_val(xp, hlog, mls)] = tab_entry<TAG>((uint32_t)pp + 2, (uint32_t)xp); }   // `if (ip1 < ip0) hashTable[hash1] = ip1`
                if (type != 3 && mend <= ilimit) {
                    uint64_t const xa = rd64(src + current0 + 2), xb = rd64(src + mend - 2);
                    T[hash_val(xa, hlog, mls)] = tab_entry<TAG>((uint32_t)current0 + 2 + 2, (uint32_t)xa);
                    T[hash_val(xb, hlog, mls)] = tab_entry<TAG>((uint32_t)(mend - 2) + 2, (uint32_t)xb);
                }
            }
            ip0 = mend; anchor = mend;
            afterMatch = true;
            step = stepSize; nextStep = ip0 + 128; d = stepSize;    // _start
        } else if (active) {
            if (validMask != LOW) active = false;     // the loop condition failed inside the window: _cleanup
            else { ip0 = P; d = D; step = S; nextStep = N; }
        }
        __syncwarp();
    }
    if (wi < nWork && l == 0) {
        it.nbSeq = nseq; it.lastLL = (uint32_t)(srcSize - anchor);
        if (MB) { it.repNext[0] = rep1 ? rep1 : offsetSaved; it.repNext[1] = rep2 ? rep2 : offsetSaved; }     // :232-233
    }
}

// ------------------------------------------------------------------------------------------------------------
//  Group-per-chunk, exact ZSTD_dfast parse (level 3; ZstdDoubleFast.cs:51-248), same warp-uniform construction.
//  Lane j < GS-1 owns probe position ip_j of the window (the schedule ip_{j+1} = ip_j + step, step + 1 every 256 bytes,
//  depends only on the state until a match is found); lane GS-1 only supplies the long-table probe of the position
//  after the last one ("_search_next_long", :167-205).  Both tables are read before they are written at every position,
//  in position order: a probe sees earlier probes of its window through __match_any_sync and older positions
//  through the tables, and writes are committed up to the first event only.
// ------------------------------------------------------------------------------------------------------------
template <int GS, bool MB>
__global__ void __launch_bounds__(32 * kMatchWarps) enc_match_dfast_group_kernel(EncPass p, const uint32_t* __restrict__ workList, uint32_t nWork, uint32_t wave)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr int NG = 32 / GS;
    constexpr uint32_t LOW = GS == 32 ? FULL : ((1u << (GS & 31)) - 1u);
    uint32_t const lane = threadIdx.x & 31, g = lane / GS, l = lane % GS, gbase = g * GS;
    uint32_t const gmask = LOW << gbase;
    uint32_t const wi = (blockIdx.x * kMatchWarps + (threadIdx.x >> 5)) * NG + g;
    bool active = wi < nWork;
    uint32_t const item = workList[active ? wi : 0];
    EncItem& it = p.items[item];
    uint32_t const hBitsL = it.hashLog, hBitsS = it.chainLog, mls = it.minMatch;
    const uint8_t* const src = p.src + it.srcOff;                 // frame start
    uint32_t* const TL = p.tables + it.tableOff;                  // long table (hash8), then short table (hash mls)
    uint32_t* const TS = TL + (1u << hBitsL);
    uint32_t* const oLL = p.seqLL + (size_t)item * p.seqStride;
    uint32_t* const oML = p.seqML + (size_t)item * p.seqStride;
    uint32_t* const oOF = p.seqOF + (size_t)item * p.seqStride;
    int srcSize = active ? (int)it.srcSize : 64;     // end of the block
    int ip = 1, anchor = 0, lowPos = 0;              // first position of a frame is skipped (:84)
    uint32_t off1 = 1, off2 = 0, offsetSaved = 4;    // offset_2 = 4 exceeds the history at frame start
    if (MB && active) {
        BlkGeom const bg = blk_geom(it.srcSize, it.windowLog, wave);
        srcSize = bg.end; ip = bg.ip0; anchor = bg.B; lowPos = bg.lowPos;
        off1 = it.rep[0]; off2 = it.rep[1]; offsetSaved = 0;
        if (off2 > bg.maxRep) { offsetSaved = off2; off2 = 0; }
        if (off1 > bg.maxRep) { offsetSaved = off1; off1 = 0; }
    }
    int const ilimit = srcSize - 8;
    uint32_t nseq = 0;
    int step = 1, nextStep = ip + 256;
    bool afterMatch = false;
    constexpr bool TAG = !MB;      // single-block frames: entries carry a 14-bit tag of the four bytes at the position (see enc_match_group_kernel)
    auto gballot = [&](bool pr) -> uint32_t { return (__ballot_sync(FULL, pr) >> gbase) & LOW; };
    auto hashL = [&](uint64_t x) { return (uint32_t)((x * 0xCF1BBCDCB7A56463ull) >> (64 - hBitsL)); };
    // common prefix of src[a..) and src[b..) by the group, lane l compares 4 bytes per round (ZSTD_count :264); valid in every lane
    auto gcount = [&](bool on, int a, int b) -> int {
        int total = 0; bool cnt = on;
        while (__any_sync(FULL, cnt)) {
            int const pa = a + total + 4 * (int)l, pb = b + total + 4 * (int)l;
            int const rem = srcSize - pa;
            uint32_t n = 4;
            if (cnt) {
                if (rem >= 4) { uint32_t const diff = rd32(src + pa) ^ rd32(src + pb); n = diff ? (uint32_t)(__ffs((int)diff) - 1) >> 3 : 4u; }
                else { n = 0; for (int j = 0; j < rem; j++) { if (src[pa + j] == src[pb + j]) n++; else break; } }
            }
            uint32_t const notFull = gballot(cnt && n != 4);
            uint32_t const f = notFull ? (uint32_t)__ffs((int)notFull) - 1u : 0u;
            uint32_t const nf = __shfl_sync(FULL, n, gbase + f);
            if (cnt) { if (notFull) { total += 4 * (int)f + (int)nf; cnt = false; } else total += 4 * GS; }
        }
        return total;
    };
    while (__any_sync(FULL, active)) {
        // ---- window schedule: position of lane l, the step in force there, state after GS-1 positions ----
        int P = ip, S = step, N = nextStep, pj = 0, sj = 1;
        int Pn = 0, Sn = 0, Nn = 0;
#pragma unroll
        for (int j = 0; j < GS; j++) {
            if (j == (int)l) { pj = P; sj = S; }
            if (j == GS - 1) { Pn = P; Sn = S; Nn = N; }           // state if positions 0..GS-2 all fail
            int const p1 = P + S;
            if (p1 >= N) { S++; N += 256; }
            P = p1;
        }
        bool const vj = active && (pj + sj <= ilimit);             // loop condition ip1 <= ilimit at this position
        bool const probe = vj && l < GS - 1;                        // lane GS-1 only looks ahead
        uint32_t const validMask = gballot(vj);
        // ---- open offset_2 loop (:231-243) ----
        bool const r2 = active && afterMatch && ip <= ilimit && off2 > 0;
        bool const r2hit = r2 && rd32(src + ip) == rd32(src + ip - (int)off2);
        // ---- probes ----
        bool const rdok = active && pj <= ilimit;                   // 8 readable bytes (the look-ahead lane may sit beyond the last position)
        uint64_t const x = rdok ? rd64(src + pj) : 0ull;
        uint32_t const hl = hashL(x), hs = hash_val(x, hBitsS, mls);
        bool const repHit = probe && off1 > 0 && rd32(src + pj + 1) == rd32(src + pj + 1 - (int)off1);
        uint32_t const tl = rdok ? __ldcg(TL + hl) : 0u, ts = probe ? __ldcg(TS + hs) : 0u;
        uint32_t const peersL = (__match_any_sync(FULL, rdok ? (hl | (g << 24)) : (0x80000000u | lane)) >> gbase) & LOW;
        uint32_t const peersS = (__match_any_sync(FULL, probe ? (hs | (g << 24)) : (0x80000000u | lane)) >> gbase) & LOW;
        uint32_t const lowL = peersL & ((1u << l) - 1u), lowS = peersS & ((1u << l) - 1u);
        int const fl = lowL ? 31 - __clz((int)lowL) : (int)l, fs = lowS ? 31 - __clz((int)lowS) : (int)l;
        int const pfl = __shfl_sync(FULL, pj, gbase + fl), pfs = __shfl_sync(FULL, pj, gbase + fs);
        int const candL = lowL ? pfl : (int)(TAG ? (tl & kPosMask) : tl) - 2, candS = lowS ? pfs : (int)(TAG ? (ts & kPosMask) : ts) - 2;     // tables store position + 2; valid iff index > 2
        // candidates forwarded from a lower lane of the window are compared in registers; table candidates are fetched only when the tag agrees
        uint64_t const xfl = __shfl_sync(FULL, x, gbase + fl); uint32_t const xfs = (uint32_t)__shfl_sync(FULL, x, gbase + fs);
        uint32_t const myTag = tag4((uint32_t)x);
        bool const mayL = rdok && !lowL && candL > lowPos && (!TAG || (tl >> kPosBits) == myTag);          // idx > prefixLowestIndex
        bool const mayS = probe && !lowS && candS > lowPos && (!TAG || (ts >> kPosBits) == myTag);
        uint64_t const cL8 = mayL ? rd64(src + candL) : 0ull; uint32_t const cS4 = mayS ? rd32(src + candS) : 0u;
        bool const Lhit = lowL ? (rdok && candL > lowPos && xfl == x) : (mayL && cL8 == x);
        bool const Shit = lowS ? (probe && candS > lowPos && xfs == (uint32_t)x) : (mayS && cS4 == (uint32_t)x);
        uint32_t key = 0xFFFFFFFFu;                                 // 0: offset_2 repeat at ip; 1+3j: repcode at ip+1; 2+3j: long at ip; 3+3j: short at ip
        if (probe) { if (Shit) key = 3 * l + 3; if (Lhit) key = 3 * l + 2; if (repHit) key = 3 * l + 1; }
        if (r2hit) key = 0;
        uint32_t const best = __reduce_min_sync(gmask, key);
        bool const ev = active && best != 0xFFFFFFFFu;
        int const type = !ev ? -1 : (best == 0 ? 3 : (int)((best - 1) % 3));         // 0 repcode, 1 long, 2 short (-> next long?), 3 offset_2 repeat
        uint32_t const je = (ev && best) ? (best - 1) / 3 : 0u;
        if (active && type != 3) afterMatch = false;
        // ---- both tables are written at every position up to the event (:122); the latest position of a bucket wins ----
        {
            uint32_t const lastLane = type < 0 ? (uint32_t)GS - 2 : je;
            // only positions that are visited write: a lane whose loop condition failed but which still had 8 readable bytes takes part in
            // peersL (its look-up may be the `ip1` of the lane below) and must not shadow a lower lane's write (soak seed 993001: the last
            // visited position of an incompressible block and the unvisited one behind it held the same 8 bytes)
            uint32_t const keep = ((2u << lastLane) - 1u) & gballot(probe);
            if (probe && type != 3 && l <= lastLane) {
                if ((((peersL & keep) >> l) >> 1) == 0) TL[hl] = tab_entry<TAG>((uint32_t)pj + 2, (uint32_t)x);
                if ((((peersS & keep) >> l) >> 1) == 0) TS[hs] = tab_entry<TAG>((uint32_t)pj + 2, (uint32_t)x);
            }
        }
        if (type == 3 && l == 0) { uint64_t const xi = rd64(src + ip); TS[hash_val(xi, hBitsS, mls)] = tab_entry<TAG>((uint32_t)ip + 2, (uint32_t)xi); TL[hashL(xi)] = tab_entry<TAG>((uint32_t)ip + 2, (uint32_t)xi); }   // :237-238
        // ---- event data (group-uniform) ----
        int const pe = __shfl_sync(FULL, pj, gbase + je), se = __shfl_sync(FULL, sj, gbase + je);
        int const cLe = __shfl_sync(FULL, candL, gbase + je), cSe = __shfl_sync(FULL, candS, gbase + je);
        uint32_t const nx = je + 1 < (uint32_t)GS ? je + 1 : je;
        bool const nextLong = __shfl_sync(FULL, (int)Lhit, gbase + nx) != 0;       // long match at ip1 (:167-177); Lhit implies 8 readable bytes there
        int const cLn = __shfl_sync(FULL, candL, gbase + nx);
        uint32_t const hln = __shfl_sync(FULL, hl, gbase + nx);
        uint32_t const x4n = (uint32_t)__shfl_sync(FULL, x, gbase + nx);
        int const p1e = pe + se;                                    // ip1 of the event position
        int mpos = 0, msrc = 0, mlen = 0, curr = pe; uint32_t offcode = 0; int base8 = 4;
        bool bext = false;
        if (type == 3) { mpos = ip; msrc = ip - (int)off2; curr = 0; uint32_t const t = off2; off2 = off1; off1 = t; }
        else if (type == 0) { mpos = pe + 1; msrc = mpos - (int)off1; }
        else if (type == 1) { mpos = pe; msrc = cLe; base8 = 8; bext = true; }
        else if (type == 2) {
            if (nextLong && p1e <= ilimit) { mpos = p1e; msrc = cLn; base8 = 8; } else { mpos = pe; msrc = cSe; }
            bext = true;
        }
        mlen = base8 + gcount(ev, mpos + base8, msrc + base8);
        // backward extension (:141-146, :178-205): lane l looks l+1 bytes back
        {
            bool ext = bext;
            while (__any_sync(FULL, ext)) {
                int const a = mpos - 1 - (int)l, b = msrc - 1 - (int)l;
                bool const ok = ext && a >= anchor && b >= lowPos && src[a] == src[b];     // ip > anchor and match > prefixLowest before every step
                uint32_t const okm = gballot(ok);
                uint32_t const n = okm == LOW ? (uint32_t)GS : (uint32_t)__ffs((int)~okm) - 1u;
                if (ext) { mpos -= (int)n; msrc -= (int)n; mlen += (int)n; ext = n == (uint32_t)GS; }
            }
        }
        if (ev) {
            if (type == 1 || type == 2) {
                uint32_t const offset = (uint32_t)(mpos - msrc);
                off2 = off1; off1 = offset; offcode = offset + 2;
                if (l == 0 && se < 4 && p1e <= ilimit) TL[hln] = tab_entry<TAG>((uint32_t)p1e + 2, x4n);   // complementary insertion (:210-213): hashLong[hl1] = ip1
            }
            ZB_ASSERT(nseq < p.seqStride && mpos >= anchor && msrc >= lowPos && msrc < mpos && mpos + mlen <= srcSize);
            if (l == 0) { oLL[nseq] = (uint32_t)(mpos - anchor); oOF[nseq] = offcode + 1; oML[nseq] = (uint32_t)mlen - 3; }
            nseq++;
            int const mend = mpos + mlen;
            if (l == 0 && type != 3 && mend <= ilimit) {             // :221-229
                int const ins = curr + 2;
                uint64_t const xa = rd64(src + ins), xb = rd64(src + mend - 2), xc = rd64(src + mend - 1);
                TL[hashL(xa)] = tab_entry<TAG>((uint32_t)ins + 2, (uint32_t)xa);
                TL[hashL(xb)] = tab_entry<TAG>((uint32_t)(mend - 2) + 2, (uint32_t)xb);
                TS[hash_val(xa, hBitsS, mls)] = tab_entry<TAG>((uint32_t)ins + 2, (uint32_t)xa);
                TS[hash_val(xc, hBitsS, mls)] = tab_entry<TAG>((uint32_t)(mend - 1) + 2, (uint32_t)xc);
            }
            ip = mend; anchor = mend;
            afterMatch = true;
            step = 1; nextStep = ip + 256;
        } else if (active) {
            if ((validMask & (LOW >> 1)) != (LOW >> 1)) active = false;     // a position of the window failed the loop condition: _cleanup
            else { ip = Pn; step = Sn; nextStep = Nn; }
        }
        __syncwarp();
    }
    if (wi < nWork && l == 0) {
        it.nbSeq = nseq; it.lastLL = (uint32_t)(srcSize - anchor);
        if (MB) { it.repNext[0] = off1 ? off1 : offsetSaved; it.repNext[1] = off2 ? off2 : offsetSaved; }     // :215-216
    }
}

// ------------------------------------------------------------------------------------------------------------
//  Entropy stage: one CTA per chunk
// ------------------------------------------------------------------------------------------------------------
constexpr int kEntThreads = 128;

struct NodeElt { uint32_t count; uint16_t parent; uint8_t byte; uint8_t nbBits; };   // nodeElt_s.cs
struct SymbolTT { int32_t deltaFindState; uint32_t deltaNbBits; };                   // FSE_symbolCompressionTransform.cs
struct FseCTable { uint32_t tableLog; uint16_t stateTable[512]; SymbolTT tt[53]; };
// the same table in HBM, handed from the front half of the entropy stage to the state-chain kernel (tt first: 8-byte aligned)
struct __align__(8) FseGTable { SymbolTT tt[53]; uint32_t tableLog; uint32_t _pad; uint16_t stateTable[512]; };
// per block: what the back half of the entropy stage needs from the front half and from the state chains
struct EntCarry { uint32_t op, lastCountSize, flags /* 1: has sequences (chains run), 2: new Huffman table */, finalState[3], _pad[2]; };

struct EntShared {
    // The literals section is finished before the sequences section starts, so their scratch shares storage:
    // 13.5 KB per CTA instead of 21 KB -> 15 CTAs of 128 threads per SM.  The kernel's time is dominated by serial
    // sections (one thread builds the Huffman tree, three threads walk the FSE state chains), so what counts is how
    // many chunks are resident at once.
    union {
        struct {    // literals
            uint32_t hist[4][256];          // per-segment literal histograms
            uint32_t count[256];            // total literal histogram
            uint8_t hufNbBits[256]; uint16_t hufValue[256];
            NodeElt huffNode[513];
            uint16_t rankBase[192], rankCurr[192];
            uint8_t huffWeight[256];
            FseCTable wct;                  // weights table (log <= 6)
            int16_t norm[64];
            uint16_t cumul[64];
            uint8_t tableSymbol[512];
            uint8_t hdr[192];               // Huffman tree description staging
        };
        struct {    // sequences: the three tables (0 LL, 1 OF, 2 ML) are built concurrently by three threads, private scratch each
            FseCTable ct[3];
            uint32_t count3[3][64]; int16_t norm3[3][64]; uint16_t cumul3[3][64]; uint8_t tableSymbol3[3][512]; uint8_t hdr3[3][128];
            uint32_t res3[3][2];            // [k] = {countSize, type}
        };
    };
    uint32_t scanA[kEntThreads / 32], scanB[kEntThreads / 32];
    uint32_t streamBits[4];
    // scalars
    uint32_t litSize, hufLog, maxSym, hSize, litMode /*0 raw 1 rle 2 huf*/, singleStream, cLitSize, litSectionSize;
    uint32_t seqHdrSize, lastCountSize, llType, ofType, mlType, bitstreamBits;
    uint32_t op;                    // write cursor inside the dst slot
};


__device__ __forceinline__ uint32_t ent_scan_excl(uint32_t v, uint32_t* warpSums, uint32_t* total)
{
    uint32_t const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t const o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kEntThreads / 32; k++) { uint32_t const s = warpSums[k]; if ((uint32_t)k < warp) base += s; tot += s; }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

// ---- serial helpers (thread 0) ----
__device__ uint32_t fse_min_table_log(uint32_t srcSize, uint32_t maxSymbolValue)      // FseCompress.cs:384
{ uint32_t const a = highbit32(srcSize) + 1, b = highbit32(maxSymbolValue) + 2; return a < b ? a : b; }
__device__ uint32_t fse_optimal_table_log(uint32_t maxTableLog, uint32_t srcSize, uint32_t maxSymbolValue, uint32_t minus)   // :397
{
    uint32_t const maxBitsSrc = highbit32(srcSize - 1) - minus;
    uint32_t tableLog = maxTableLog;
    uint32_t const minBits = fse_min_table_log(srcSize, maxSymbolValue);
    if (tableLog == 0) tableLog = 11;
    if (maxBitsSrc < tableLog) tableLog = maxBitsSrc;
    if (minBits > tableLog) tableLog = minBits;
    if (tableLog < 5) tableLog = 5;
    if (tableLog > 12) tableLog = 12;
    return tableLog;
}

__device__ bool fse_normalize_m2(int16_t* norm, uint32_t tableLog, const uint32_t* count, uint32_t total, uint32_t maxSymbolValue, int16_t lowProbCount)   // :443
{
    int16_t const NOT_YET_ASSIGNED = -2;
    uint32_t s, distributed = 0, ToDistribute;
    uint32_t const lowThreshold = total >> tableLog;
    uint32_t lowOne = (uint32_t)(((uint64_t)total * 3) >> (tableLog + 1));
    for (s = 0; s <= maxSymbolValue; s++) {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = lowProbCount; distributed++; total -= count[s]; continue; }
        if (count[s] <= lowOne) { norm[s] = 1; distributed++; total -= count[s]; continue; }
        norm[s] = NOT_YET_ASSIGNED;
    }
    ToDistribute = (1u << tableLog) - distributed;
    if (ToDistribute == 0) return true;
    if ((total / ToDistribute) > lowOne) {
        lowOne = (uint32_t)(((uint64_t)total * 3) / (ToDistribute * 2));
        for (s = 0; s <= maxSymbolValue; s++)
            if ((norm[s] == NOT_YET_ASSIGNED) && (count[s] <= lowOne)) { norm[s] = 1; distributed++; total -= count[s]; }
        ToDistribute = (1u << tableLog) - distributed;
    }
    if (distributed == maxSymbolValue + 1) {
        uint32_t maxV = 0, maxC = 0;
        for (s = 0; s <= maxSymbolValue; s++) if (count[s] > maxC) { maxV = s; maxC = count[s]; }
        norm[maxV] += (int16_t)ToDistribute;
        return true;
    }
    if (total == 0) {
        for (s = 0; ToDistribute > 0; s = (s + 1) % (maxSymbolValue + 1)) if (norm[s] > 0) { ToDistribute--; norm[s]++; }
        return true;
    }
    {   uint64_t const vStepLog = 62 - tableLog;
        uint64_t const mid = (1ull << (vStepLog - 1)) - 1;
        uint64_t const rStep = ((((uint64_t)1 << vStepLog) * ToDistribute) + mid) / total;
        uint64_t tmpTotal = mid;
        for (s = 0; s <= maxSymbolValue; s++) {
            if (norm[s] == NOT_YET_ASSIGNED) {
                uint64_t const end = tmpTotal + ((uint64_t)count[s] * rStep);
                uint32_t const sStart = (uint32_t)(tmpTotal >> vStepLog), sEnd = (uint32_t)(end >> vStepLog);
                uint32_t const weight = sEnd - sStart;
                if (weight < 1) return false;
                norm[s] = (int16_t)weight;
                tmpTotal = end;
            }
        }
    }
    return true;
}

// FSE_normalizeCount (:574). Returns 0 on error / rle special case, else tableLog.
__device__ uint32_t fse_normalize_count(int16_t* norm, uint32_t tableLog, const uint32_t* count, uint32_t total, uint32_t maxSymbolValue, uint32_t useLowProbCount)
{
    if (tableLog < 5 || tableLog > 12) return 0;
    if (tableLog < fse_min_table_log(total, maxSymbolValue)) return 0;
    int16_t const lowProbCount = useLowProbCount ? -1 : 1;
    uint64_t const scale = 62 - tableLog;
    uint64_t const step = ((uint64_t)1 << 62) / total;
    uint64_t const vStep = 1ull << (scale - 20);
    int stillToDistribute = 1 << tableLog;
    uint32_t s, largest = 0; int16_t largestP = 0;
    uint32_t const lowThreshold = total >> tableLog;
    for (s = 0; s <= maxSymbolValue; s++) {
        if (count[s] == total) return 0;
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = lowProbCount; stillToDistribute--; }
        else {
            int16_t proba = (int16_t)(((uint64_t)count[s] * step) >> scale);
            if (proba < 8) {
                uint64_t const restToBeat = vStep * c_rtbTable[proba];
                proba += ((uint64_t)count[s] * step) - ((uint64_t)proba << scale) > restToBeat;
            }
            if (proba > largestP) { largestP = proba; largest = s; }
            norm[s] = proba;
            stillToDistribute -= proba;
        }
    }
    if (-stillToDistribute >= (norm[largest] >> 1)) { if (!fse_normalize_m2(norm, tableLog, count, total, maxSymbolValue, lowProbCount)) return 0; }
    else norm[largest] += (int16_t)stillToDistribute;
    return tableLog;
}

// FSE_writeNCount_generic (:203), writeIsSafe: the staging buffer is always large enough. Returns size or 0 on error.
__device__ uint32_t fse_write_ncount(uint8_t* out0, const int16_t* norm, uint32_t maxSymbolValue, uint32_t tableLog)
{
    uint8_t* out = out0;
    int nbBits; int const tableSize = 1 << tableLog; int remaining, threshold;
    uint32_t bitStream = 0; int bitCount = 0; uint32_t symbol = 0; uint32_t const alphabetSize = maxSymbolValue + 1; int previousIs0 = 0;
    bitStream += (tableLog - 5) << bitCount; bitCount += 4;
    remaining = tableSize + 1; threshold = tableSize; nbBits = (int)tableLog + 1;
    while ((symbol < alphabetSize) && (remaining > 1)) {
        if (previousIs0) {
            uint32_t start = symbol;
            while ((symbol < alphabetSize) && !norm[symbol]) symbol++;
            if (symbol == alphabetSize) break;
            while (symbol >= start + 24) { start += 24; bitStream += 0xFFFFu << bitCount; out[0] = (uint8_t)bitStream; out[1] = (uint8_t)(bitStream >> 8); out += 2; bitStream >>= 16; }
            while (symbol >= start + 3) { start += 3; bitStream += 3u << bitCount; bitCount += 2; }
            bitStream += (symbol - start) << bitCount; bitCount += 2;
            if (bitCount > 16) { out[0] = (uint8_t)bitStream; out[1] = (uint8_t)(bitStream >> 8); out += 2; bitStream >>= 16; bitCount -= 16; }
        }
        {   int count = norm[symbol++];
            int const max = (2 * threshold - 1) - remaining;
            remaining -= count < 0 ? -count : count;
            count++;
            if (count >= threshold) count += max;
            bitStream += (uint32_t)count << bitCount;
            bitCount += nbBits;
            bitCount -= (count < max);
            previousIs0 = (count == 1);
            if (remaining < 1) return 0;
            while (remaining < threshold) { nbBits--; threshold >>= 1; }
        }
        if (bitCount > 16) { out[0] = (uint8_t)bitStream; out[1] = (uint8_t)(bitStream >> 8); out += 2; bitStream >>= 16; bitCount -= 16; }
    }
    if (remaining != 1) return 0;
    out[0] = (uint8_t)bitStream; out[1] = (uint8_t)(bitStream >> 8);
    out += (bitCount + 7) / 8;
    return (uint32_t)(out - out0);
}

// FSE_buildCTable_wksp (:13)
__device__ void fse_build_ctable(FseCTable& ct, uint16_t* cumul, uint8_t* tableSymbol, const int16_t* norm, uint32_t maxSymbolValue, uint32_t tableLog)
{
    uint32_t const tableSize = 1u << tableLog, tableMask = tableSize - 1;
    uint32_t const step = (tableSize >> 1) + (tableSize >> 3) + 3;
    uint32_t const maxSV1 = maxSymbolValue + 1;
    uint32_t highThreshold = tableSize - 1;
    ct.tableLog = tableLog;
    cumul[0] = 0;
    for (uint32_t u = 1; u <= maxSV1; u++) {
        if (norm[u - 1] == -1) { cumul[u] = (uint16_t)(cumul[u - 1] + 1); tableSymbol[highThreshold--] = (uint8_t)(u - 1); }
        else cumul[u] = (uint16_t)(cumul[u - 1] + (uint16_t)norm[u - 1]);
    }
    cumul[maxSV1] = (uint16_t)(tableSize + 1);
    {   uint32_t position = 0;
        for (uint32_t symbol = 0; symbol < maxSV1; symbol++) {
            int const freq = norm[symbol];
            for (int n = 0; n < freq; n++) {
                tableSymbol[position] = (uint8_t)symbol;
                position = (position + step) & tableMask;
                while (position > highThreshold) position = (position + step) & tableMask;
            }
        }
    }
    for (uint32_t u = 0; u < tableSize; u++) { uint8_t const s = tableSymbol[u]; ct.stateTable[cumul[s]++] = (uint16_t)(tableSize + u); }
    {   uint32_t total = 0;
        for (uint32_t s = 0; s <= maxSymbolValue; s++) {
            int const nc = norm[s];
            if (nc == 0) { ct.tt[s].deltaNbBits = ((tableLog + 1) << 16) - (1u << tableLog); ct.tt[s].deltaFindState = 0; }
            else if (nc == -1 || nc == 1) { ct.tt[s].deltaNbBits = (tableLog << 16) - (1u << tableLog); ct.tt[s].deltaFindState = (int32_t)(total - 1); total++; }
            else {
                uint32_t const maxBitsOut = tableLog - highbit32((uint32_t)nc - 1);
                uint32_t const minStatePlus = (uint32_t)nc << maxBitsOut;
                ct.tt[s].deltaNbBits = (maxBitsOut << 16) - minStatePlus;
                ct.tt[s].deltaFindState = (int32_t)(total - (uint32_t)nc);
                total += (uint32_t)nc;
            }
        }
    }
}

// FSE state helpers (Fse.cs:10-96)
__device__ __forceinline__ uint32_t fse_init_state(const FseCTable& ct, uint32_t symbol)
{
    SymbolTT const tt = ct.tt[symbol];
    uint32_t const nbBitsOut = (tt.deltaNbBits + (1u << 15)) >> 16;
    uint32_t const value = (nbBitsOut << 16) - tt.deltaNbBits;
    return ct.stateTable[(int32_t)(value >> nbBitsOut) + tt.deltaFindState];
}

// simple serial LSB-first bit writer into shared/global bytes (used only for the <= 128-byte weight stream)
struct TinyBW { uint64_t acc; uint32_t n; uint8_t* p; uint8_t* start;
    __device__ void add(uint32_t v, uint32_t nb) { if (!nb) return; acc |= (uint64_t)(v & ((1u << nb) - 1)) << n; n += nb; while (n >= 8) { *p++ = (uint8_t)acc; acc >>= 8; n -= 8; } }
    __device__ uint32_t close() { add(1, 1); if (n) { *p++ = (uint8_t)acc; } return (uint32_t)(p - start); } };

// HUF_compressWeights (HufCompress.cs:40). Returns compressed size, 0 = not compressible, 1 = rle; 0xFFFFFFFF on error.
__device__ uint32_t huf_compress_weights(uint8_t* dst, uint32_t dstSize, EntShared& S, uint32_t wtSize)
{
    uint32_t maxSymbolValue = 12, tableLog = 6;
    uint32_t count[13];
    if (wtSize <= 1) return 0;
    for (int i = 0; i < 13; i++) count[i] = 0;
    for (uint32_t i = 0; i < wtSize; i++) count[S.huffWeight[i]]++;
    while (!count[maxSymbolValue]) maxSymbolValue--;
    uint32_t maxCount = 0; for (uint32_t s = 0; s <= maxSymbolValue; s++) if (count[s] > maxCount) maxCount = count[s];
    if (maxCount == wtSize) return 1;
    if (maxCount == 1) return 0;
    tableLog = fse_optimal_table_log(tableLog, wtSize, maxSymbolValue, 2);
    if (!fse_normalize_count(S.norm, tableLog, count, wtSize, maxSymbolValue, 0)) return 0xFFFFFFFFu;
    uint32_t const hSize = fse_write_ncount(dst, S.norm, maxSymbolValue, tableLog);
    if (!hSize) return 0xFFFFFFFFu;
    fse_build_ctable(S.wct, S.cumul, S.tableSymbol, S.norm, maxSymbolValue, tableLog);
    // FSE_compress_usingCTable_generic (FseCompress.cs:722), 64-bit variant
    if (wtSize <= 2) return 0;
    if (dstSize - hSize <= 8) return 0;
    TinyBW bw{0, 0, dst + hSize, dst + hSize};
    const uint8_t* ip = S.huffWeight + wtSize;
    uint32_t s1, s2; uint32_t n = wtSize;
    auto enc = [&](uint32_t& st, uint32_t sym) { SymbolTT const tt = S.wct.tt[sym]; uint32_t const nb = (st + tt.deltaNbBits) >> 16; bw.add(st, nb); st = S.wct.stateTable[(int32_t)(st >> nb) + tt.deltaFindState]; };
    if (n & 1) { s1 = fse_init_state(S.wct, *--ip); s2 = fse_init_state(S.wct, *--ip); enc(s1, *--ip); }
    else { s2 = fse_init_state(S.wct, *--ip); s1 = fse_init_state(S.wct, *--ip); }
    n -= 2;
    if (n & 2) { enc(s2, *--ip); enc(s1, *--ip); }
    while (ip > S.huffWeight) { enc(s2, *--ip); enc(s1, *--ip); enc(s2, *--ip); enc(s1, *--ip); }
    bw.add(s2, tableLog); bw.add(s1, tableLog);
    uint32_t const cSize = bw.close();
    return hSize + cSize;
}

// HUF_sort helpers (HufCompress.cs:520-687)
__device__ __forceinline__ uint32_t huf_get_index(uint32_t count) { return count < 165 ? count : highbit32(count) + 158; }
__device__ void huf_insertion_sort(NodeElt* a, int low, int high)
{
    int const size = high - low + 1; a += low;
    for (int i = 1; i < size; ++i) { NodeElt const key = a[i]; int j = i - 1; while (j >= 0 && a[j].count < key.count) { a[j + 1] = a[j]; j--; } a[j + 1] = key; }
}
__device__ int huf_partition(NodeElt* arr, int low, int high)
{
    uint32_t const pivot = arr[high].count; int i = low - 1;
    for (int j = low; j < high; j++) if (arr[j].count > pivot) { i++; NodeElt t = arr[i]; arr[i] = arr[j]; arr[j] = t; }
    NodeElt t = arr[i + 1]; arr[i + 1] = arr[high]; arr[high] = t;
    return i + 1;
}
// HUF_simpleQuickSort (:607-633) with an explicit stack.  The reference checks the insertion-sort threshold only on
// entry to a (recursive) call; the part it keeps iterating on is partitioned down to single elements.  Both modes
// are reproduced (mode 0 = call entry, mode 1 = loop continuation); sub-ranges are disjoint, so their order is free.
__device__ void huf_quick_sort(NodeElt* arr, int low0, int high0)
{
    int stackLo[48], stackHi[48]; uint8_t stackMode[48]; int sp = 0;
    stackLo[0] = low0; stackHi[0] = high0; stackMode[0] = 0; sp = 1;
    while (sp) {
        sp--; int const low = stackLo[sp], high = stackHi[sp]; int const mode = stackMode[sp];
        if (mode == 0 && high - low < 8) { huf_insertion_sort(arr, low, high); continue; }
        if (!(low < high)) continue;
        int const idx = huf_partition(arr, low, high);
        if (idx - low < high - idx) {
            stackLo[sp] = idx + 1; stackHi[sp] = high; stackMode[sp] = 1; sp++;
            stackLo[sp] = low; stackHi[sp] = idx - 1; stackMode[sp] = 0; sp++;
        } else {
            stackLo[sp] = low; stackHi[sp] = idx - 1; stackMode[sp] = 1; sp++;
            stackLo[sp] = idx + 1; stackHi[sp] = high; stackMode[sp] = 0; sp++;
        }
    }
}

// HUF_buildCTable_wksp (HufCompress.cs:790): sort, tree, depth limit, canonical codes. Returns maxNbBits.
__device__ uint32_t huf_build_ctable(EntShared& S, uint32_t maxSymbolValue, uint32_t maxNbBits)
{
    NodeElt* const huffNode0 = S.huffNode; NodeElt* const huffNode = huffNode0 + 1;
    for (int i = 0; i < 513; i++) { huffNode0[i].count = 0; huffNode0[i].parent = 0; huffNode0[i].byte = 0; huffNode0[i].nbBits = 0; }
    // HUF_sort :635
    uint32_t const maxSymbolValue1 = maxSymbolValue + 1;
    for (int i = 0; i < 192; i++) { S.rankBase[i] = 0; S.rankCurr[i] = 0; }
    for (uint32_t n = 0; n < maxSymbolValue1; ++n) S.rankBase[huf_get_index(S.count[n])]++;
    for (uint32_t n = 191; n > 0; --n) { S.rankBase[n - 1] += S.rankBase[n]; S.rankCurr[n - 1] = S.rankBase[n - 1]; }
    for (uint32_t n = 0; n < maxSymbolValue1; ++n) {
        uint32_t const c = S.count[n]; uint32_t const r = huf_get_index(c) + 1; uint32_t const pos = S.rankCurr[r]++;
        huffNode[pos].count = c; huffNode[pos].byte = (uint8_t)n;
    }
    for (uint32_t n = 165; n < 191; ++n) {
        uint32_t const bucketSize = S.rankCurr[n] - S.rankBase[n]; uint32_t const bucketStartIdx = S.rankBase[n];
        if (bucketSize > 1) huf_quick_sort(huffNode + bucketStartIdx, 0, (int)bucketSize - 1);
    }
    // HUF_buildTree :689
    int nonNullRank = (int)maxSymbolValue;
    while (huffNode[nonNullRank].count == 0) nonNullRank--;
    {
        int lowS = nonNullRank, nodeNb = 256; int const nodeRoot = nodeNb + lowS - 1; int lowN = nodeNb;
        huffNode[nodeNb].count = huffNode[lowS].count + huffNode[lowS - 1].count;
        huffNode[lowS].parent = huffNode[lowS - 1].parent = (uint16_t)nodeNb;
        nodeNb++; lowS -= 2;
        for (int n = nodeNb; n <= nodeRoot; n++) huffNode[n].count = 1u << 30;
        huffNode0[0].count = 1u << 31;
        while (nodeNb <= nodeRoot) {
            int const n1 = (huffNode[lowS].count < huffNode[lowN].count) ? lowS-- : lowN++;
            int const n2 = (huffNode[lowS].count < huffNode[lowN].count) ? lowS-- : lowN++;
            huffNode[nodeNb].count = huffNode[n1].count + huffNode[n2].count;
            huffNode[n1].parent = huffNode[n2].parent = (uint16_t)nodeNb;
            nodeNb++;
        }
        huffNode[nodeRoot].nbBits = 0;
        for (int n = nodeRoot - 1; n >= 256; n--) huffNode[n].nbBits = (uint8_t)(huffNode[huffNode[n].parent].nbBits + 1);
        for (int n = 0; n <= nonNullRank; n++) huffNode[n].nbBits = (uint8_t)(huffNode[huffNode[n].parent].nbBits + 1);
    }
    // HUF_setMaxHeight :377
    {
        uint32_t const largestBits = huffNode[nonNullRank].nbBits;
        if (largestBits > maxNbBits) {
            int totalCost = 0; uint32_t const baseCost = 1u << (largestBits - maxNbBits); int n = nonNullRank;
            while (huffNode[n].nbBits > maxNbBits) { totalCost += (int)(baseCost - (1u << (largestBits - huffNode[n].nbBits))); huffNode[n].nbBits = (uint8_t)maxNbBits; n--; }
            while (huffNode[n].nbBits == maxNbBits) --n;
            totalCost >>= (largestBits - maxNbBits);
            uint32_t const noSymbol = 0xF0F0F0F0u; uint32_t rankLast[14];
            for (int i = 0; i < 14; i++) rankLast[i] = noSymbol;
            {   uint32_t currentNbBits = maxNbBits;
                for (int pos = n; pos >= 0; pos--) { if (huffNode[pos].nbBits >= currentNbBits) continue; currentNbBits = huffNode[pos].nbBits; rankLast[maxNbBits - currentNbBits] = (uint32_t)pos; } }
            while (totalCost > 0) {
                uint32_t nBitsToDecrease = highbit32((uint32_t)totalCost) + 1;
                for (; nBitsToDecrease > 1; nBitsToDecrease--) {
                    uint32_t const highPos = rankLast[nBitsToDecrease], lowPos = rankLast[nBitsToDecrease - 1];
                    if (highPos == noSymbol) continue;
                    if (lowPos == noSymbol) break;
                    if (huffNode[highPos].count <= 2 * huffNode[lowPos].count) break;
                }
                while ((nBitsToDecrease <= 12) && (rankLast[nBitsToDecrease] == noSymbol)) nBitsToDecrease++;
                totalCost -= 1 << (nBitsToDecrease - 1);
                huffNode[rankLast[nBitsToDecrease]].nbBits++;
                if (rankLast[nBitsToDecrease - 1] == noSymbol) rankLast[nBitsToDecrease - 1] = rankLast[nBitsToDecrease];
                if (rankLast[nBitsToDecrease] == 0) rankLast[nBitsToDecrease] = noSymbol;
                else { rankLast[nBitsToDecrease]--; if (huffNode[rankLast[nBitsToDecrease]].nbBits != maxNbBits - nBitsToDecrease) rankLast[nBitsToDecrease] = noSymbol; }
            }
            while (totalCost < 0) {
                if (rankLast[1] == noSymbol) { while (huffNode[n].nbBits == maxNbBits) n--; huffNode[n + 1].nbBits--; rankLast[1] = (uint32_t)(n + 1); totalCost++; continue; }
                huffNode[rankLast[1] + 1].nbBits--; rankLast[1]++; totalCost++;
            }
        } else maxNbBits = largestBits;
    }
    // HUF_buildCTableFromTree :750
    {
        uint16_t nbPerRank[13], valPerRank[13];
        for (int i = 0; i < 13; i++) { nbPerRank[i] = 0; valPerRank[i] = 0; }
        for (int n = 0; n <= nonNullRank; n++) nbPerRank[huffNode[n].nbBits]++;
        {   uint16_t min = 0; for (int n = (int)maxNbBits; n > 0; n--) { valPerRank[n] = min; min += nbPerRank[n]; min >>= 1; } }
        for (int i = 0; i < 256; i++) { S.hufNbBits[i] = 0; S.hufValue[i] = 0; }
        for (uint32_t n = 0; n < maxSymbolValue1; n++) S.hufNbBits[huffNode[n].byte] = huffNode[n].nbBits;
        for (uint32_t n = 0; n < maxSymbolValue1; n++) { uint32_t const nb = S.hufNbBits[n]; uint16_t const v = valPerRank[nb]++; S.hufValue[n] = nb ? v : 0; }
    }
    return maxNbBits;
}

// HUF_writeCTable_wksp (HufCompress.cs:168) into S.hdr. Returns header size, 0 on error.
__device__ uint32_t huf_write_ctable(EntShared& S, uint32_t maxSymbolValue, uint32_t huffLog)
{
    uint8_t bitsToWeight[13];
    bitsToWeight[0] = 0;
    for (uint32_t n = 1; n < huffLog + 1; n++) bitsToWeight[n] = (uint8_t)(huffLog + 1 - n);
    for (uint32_t n = 0; n < maxSymbolValue; n++) S.huffWeight[n] = bitsToWeight[S.hufNbBits[n]];
    uint32_t const hSize = huf_compress_weights(S.hdr + 1, sizeof(S.hdr) - 1, S, maxSymbolValue);
    if (hSize == 0xFFFFFFFFu) return 0;
    if ((hSize > 1) && (hSize < maxSymbolValue / 2)) { S.hdr[0] = (uint8_t)hSize; return hSize + 1; }
    if (maxSymbolValue > 128) return 0;
    S.hdr[0] = (uint8_t)(128 + (maxSymbolValue - 1));
    S.huffWeight[maxSymbolValue] = 0;
    for (uint32_t n = 0; n < maxSymbolValue; n += 2) S.hdr[(n / 2) + 1] = (uint8_t)((S.huffWeight[n] << 4) + S.huffWeight[n + 1]);
    return ((maxSymbolValue + 1) / 2) + 1;
}

// ZSTD_selectEncodingType (ZstdCompressSequences.cs:400), strategy < ZSTD_lazy.  `valid`: the previous table's repeat mode is
// FSE_repeat_valid (only a dictionary's tables are); every outcome but set_repeat (3) leaves the mode `none` or `check`.
__device__ uint32_t select_encoding_type(uint32_t mostFrequent, uint32_t nbSeq, uint32_t defaultNormLog, bool isDefaultAllowed, uint32_t strategy, bool valid = false)
{
    if (mostFrequent == nbSeq) { if (isDefaultAllowed && nbSeq <= 2) return 0; return 1; }
    if (isDefaultAllowed) {
        uint32_t const mult = 10 - strategy;
        uint32_t const dynamicFse_nbSeq_min = ((1u << defaultNormLog) * mult) >> 3;
        if (valid && nbSeq < 1000) return 3;                                              // staticFse_nbSeq_max (:420)
        if ((nbSeq < dynamicFse_nbSeq_min) || (mostFrequent < (nbSeq >> (defaultNormLog - 1)))) return 0;
    }
    return 2;
}

__device__ __forceinline__ uint32_t ll_code(uint32_t ll) { return ll > 63 ? highbit32(ll) + 19 : c_LL_Code[ll]; }     // ZstdCompressInternal.cs:20
__device__ __forceinline__ uint32_t ml_code(uint32_t ml) { return ml > 127 ? highbit32(ml) + 36 : c_ML_Code[ml]; }    // :32

// scatter `nb` (<= 32) bits of v at absolute bit position `bit` of the zeroed word array w
__device__ __forceinline__ void put_bits(uint32_t* w, uint64_t bit, uint32_t v, uint32_t nb)
{
    if (!nb) return;
    uint32_t const sh = (uint32_t)(bit & 31); uint64_t const x = (uint64_t)(nb < 32 ? (v & ((1u << nb) - 1)) : v) << sh;
    atomicOr(&w[bit >> 5], (uint32_t)x);
    if (sh + nb > 32) atomicOr(&w[(bit >> 5) + 1], (uint32_t)(x >> 32));
}

// ------------------------------------------------------------------------------------------------------------
//  Dictionary compression (Compressor.LoadDictionary + Wrap): serial match finders, one THREAD per frame.
//  ZstdFast.cs:390 (dictMatchState), :583 (extDict), ZstdDoubleFast.cs:250, :590, and the no-dictionary forms (ZstdFast.cs:96,
//  ZstdDoubleFast.cs:51) for the blocks behind an invalidated dictionary.  The dictionary content is one shared device buffer;
//  a frame's bytes are its source: `dictBase + index` / `base + index` are the reference's two segments, and a match may run
//  from the first into the second (ZSTD_count_2segments, ZstdCompressInternal.cs:283).  This is the first, correct form of the
//  path: 32 independent parses per warp, no speculation (the batch is the parallelism); the group kernels above remain the
//  no-dictionary path.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_ptr(const uint8_t* p, uint32_t hBits, uint32_t mls) { return hash_val(rd64(p), hBits, mls); }
// ZSTD_count, ZstdCompressInternal.cs:264
__device__ uint32_t count_match(const uint8_t* pIn, const uint8_t* pMatch, const uint8_t* const pInLimit)
{
    const uint8_t* const pStart = pIn;
    while (pIn + 4 <= pInLimit) {
        uint32_t const diff = rd32(pIn) ^ rd32(pMatch);
        if (diff) return (uint32_t)(pIn - pStart) + (__ffs((int)diff) - 1) / 8;
        pIn += 4; pMatch += 4;
    }
    while (pIn < pInLimit && *pMatch == *pIn) { pIn++; pMatch++; }
    return (uint32_t)(pIn - pStart);
}
// ZSTD_count_2segments, ZstdCompressInternal.cs:283
__device__ uint32_t count_2segments(const uint8_t* ip, const uint8_t* match, const uint8_t* iEnd, const uint8_t* mEnd, const uint8_t* iStart)
{
    const uint8_t* const vEnd = (ip + (mEnd - match)) < iEnd ? ip + (mEnd - match) : iEnd;
    uint32_t const matchLength = count_match(ip, match, vEnd);
    if (match + matchLength != mEnd) return matchLength;
    return matchLength + count_match(ip + matchLength, iStart, iEnd);
}
struct SeqWriter {
    uint32_t* ll; uint32_t* ml; uint32_t* of; uint32_t n; uint32_t cap;
    __device__ __forceinline__ void store(uint32_t litLength, uint32_t offCode, uint32_t mlBase)   // ZSTD_storeSeq :204
    { ZB_ASSERT(n < cap); ll[n] = litLength; of[n] = offCode + 1; ml[n] = mlBase; n++; }
};
// what the serial parsers share: the two segments and the block
struct DictBlk {
    const uint8_t* base; const uint8_t* dictBase;     // index -> byte, in the prefix (>= prefixStartIndex) / in the dictionary segment
    const uint8_t* istart; const uint8_t* iend;
    uint32_t prefixStartIndex, dictStartIndex;        // extDict / dictMatchState: dictionary segment = [dictStartIndex, prefixStartIndex)
    uint32_t mls, dStep;
};

// ZstdFast.cs:390 ZSTD_compressBlock_fast_dictMatchState_generic
__device__ uint32_t dict_fast_dms(uint32_t* hashTable, uint32_t hlog, const uint32_t* dictHashTable, uint32_t dictHLog, const DictBlk& b, uint32_t rep[2], SeqWriter& sw)
{
    const uint8_t* const base = b.base; const uint8_t* const dictBase = b.dictBase; uint32_t const mls = b.mls, stepSize = b.dStep;
    uint32_t const prefixStartIndex = b.prefixStartIndex, dictStartIndex = b.dictStartIndex;
    const uint8_t* const prefixStart = base + prefixStartIndex; const uint8_t* const dictStart = dictBase + dictStartIndex; const uint8_t* const dictEnd = dictBase + prefixStartIndex;
    const uint8_t* ip = b.istart; const uint8_t* anchor = b.istart; const uint8_t* const iend = b.iend; const uint8_t* const ilimit = iend - 8;
    uint32_t offset_1 = rep[0], offset_2 = rep[1];
    uint32_t const dictAndPrefixLength = (uint32_t)(ip - prefixStart) + (prefixStartIndex - dictStartIndex);
    ip += (dictAndPrefixLength == 0);
    while (ip < ilimit) {
        uint32_t mLength; uint32_t const h = hash_ptr(ip, hlog, mls);
        uint32_t const curr = (uint32_t)(ip - base), matchIndex = hashTable[h]; const uint8_t* match = base + matchIndex;
        uint32_t const repIndex = curr + 1 - offset_1; const uint8_t* const repMatch = repIndex < prefixStartIndex ? dictBase + repIndex : base + repIndex;
        hashTable[h] = curr;
        if (((uint32_t)((prefixStartIndex - 1) - repIndex) >= 3) && (rd32(repMatch) == rd32(ip + 1))) {
            const uint8_t* const repMatchEnd = repIndex < prefixStartIndex ? dictEnd : iend;
            mLength = count_2segments(ip + 1 + 4, repMatch + 4, iend, repMatchEnd, prefixStart) + 4;
            ip++;
            sw.store((uint32_t)(ip - anchor), 0, mLength - 3);
        } else if (matchIndex <= prefixStartIndex) {
            uint32_t const dictMatchIndex = dictHashTable[hash_ptr(ip, dictHLog, mls)]; const uint8_t* dictMatch = dictBase + dictMatchIndex;
            if (dictMatchIndex <= dictStartIndex || rd32(dictMatch) != rd32(ip)) { ip += ((ip - anchor) >> 8) + stepSize; continue; }
            uint32_t const offset = curr - dictMatchIndex;
            mLength = count_2segments(ip + 4, dictMatch + 4, iend, dictEnd, prefixStart) + 4;
            while (((ip > anchor) & (dictMatch > dictStart)) && (ip[-1] == dictMatch[-1])) { ip--; dictMatch--; mLength++; }
            offset_2 = offset_1; offset_1 = offset;
            sw.store((uint32_t)(ip - anchor), offset + 2, mLength - 3);
        } else if (rd32(match) != rd32(ip)) { ip += ((ip - anchor) >> 8) + stepSize; continue; }
        else {
            uint32_t const offset = (uint32_t)(ip - match);
            mLength = count_match(ip + 4, match + 4, iend) + 4;
            while (((ip > anchor) & (match > prefixStart)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
            offset_2 = offset_1; offset_1 = offset;
            sw.store((uint32_t)(ip - anchor), offset + 2, mLength - 3);
        }
        ip += mLength; anchor = ip;
        if (ip <= ilimit) {
            hashTable[hash_ptr(base + curr + 2, hlog, mls)] = curr + 2;
            hashTable[hash_ptr(ip - 2, hlog, mls)] = (uint32_t)(ip - 2 - base);
            while (ip <= ilimit) {
                uint32_t const current2 = (uint32_t)(ip - base), repIndex2 = current2 - offset_2;
                const uint8_t* const repMatch2 = repIndex2 < prefixStartIndex ? dictBase + repIndex2 : base + repIndex2;
                if (((uint32_t)((prefixStartIndex - 1) - repIndex2) >= 3) && (rd32(repMatch2) == rd32(ip))) {
                    const uint8_t* const repEnd2 = repIndex2 < prefixStartIndex ? dictEnd : iend;
                    uint32_t const repLength2 = count_2segments(ip + 4, repMatch2 + 4, iend, repEnd2, prefixStart) + 4;
                    uint32_t const t = offset_2; offset_2 = offset_1; offset_1 = t;
                    sw.store(0, 0, repLength2 - 3);
                    hashTable[hash_ptr(ip, hlog, mls)] = current2;
                    ip += repLength2; anchor = ip;
                    continue;
                }
                break;
    }   }   }
    rep[0] = offset_1; rep[1] = offset_2;         // offsetSaved is 0 and the offsets never are (:520-521)
    return (uint32_t)(iend - anchor);
}

// ZstdFast.cs:583 ZSTD_compressBlock_fast_extDict_generic (the caller has already ruled out prefixStartIndex == dictStartIndex)
__device__ uint32_t dict_fast_ext(uint32_t* hashTable, uint32_t hlog, const DictBlk& b, uint32_t rep[2], SeqWriter& sw)
{
    const uint8_t* const base = b.base; const uint8_t* const dictBase = b.dictBase; uint32_t const mls = b.mls, stepSize = b.dStep;
    uint32_t const prefixStartIndex = b.prefixStartIndex, dictStartIndex = b.dictStartIndex;
    const uint8_t* const prefixStart = base + prefixStartIndex; const uint8_t* const dictStart = dictBase + dictStartIndex; const uint8_t* const dictEnd = dictBase + prefixStartIndex;
    const uint8_t* ip = b.istart; const uint8_t* anchor = b.istart; const uint8_t* const iend = b.iend; const uint8_t* const ilimit = iend - 8;
    uint32_t offset_1 = rep[0], offset_2 = rep[1];
    while (ip < ilimit) {
        uint32_t const h = hash_ptr(ip, hlog, mls);
        uint32_t const matchIndex = hashTable[h]; const uint8_t* match = (matchIndex < prefixStartIndex ? dictBase : base) + matchIndex;
        uint32_t const curr = (uint32_t)(ip - base), repIndex = curr + 1 - offset_1;
        const uint8_t* const repMatch = (repIndex < prefixStartIndex ? dictBase : base) + repIndex;
        hashTable[h] = curr;
        if ((((uint32_t)((prefixStartIndex - 1) - repIndex) >= 3) & (offset_1 <= curr + 1 - dictStartIndex)) && (rd32(repMatch) == rd32(ip + 1))) {
            const uint8_t* const repMatchEnd = repIndex < prefixStartIndex ? dictEnd : iend;
            uint32_t const rLength = count_2segments(ip + 1 + 4, repMatch + 4, iend, repMatchEnd, prefixStart) + 4;
            ip++;
            sw.store((uint32_t)(ip - anchor), 0, rLength - 3);
            ip += rLength; anchor = ip;
        } else {
            if ((matchIndex < dictStartIndex) || (rd32(match) != rd32(ip))) { ip += ((ip - anchor) >> 8) + stepSize; continue; }
            const uint8_t* const matchEnd = matchIndex < prefixStartIndex ? dictEnd : iend;
            const uint8_t* const lowMatchPtr = matchIndex < prefixStartIndex ? dictStart : prefixStart;
            uint32_t const offset = curr - matchIndex;
            uint32_t mLength = count_2segments(ip + 4, match + 4, iend, matchEnd, prefixStart) + 4;
            while (((ip > anchor) & (match > lowMatchPtr)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
            offset_2 = offset_1; offset_1 = offset;
            sw.store((uint32_t)(ip - anchor), offset + 2, mLength - 3);
            ip += mLength; anchor = ip;
        }
        if (ip <= ilimit) {
            hashTable[hash_ptr(base + curr + 2, hlog, mls)] = curr + 2;
            hashTable[hash_ptr(ip - 2, hlog, mls)] = (uint32_t)(ip - 2 - base);
            while (ip <= ilimit) {
                uint32_t const current2 = (uint32_t)(ip - base), repIndex2 = current2 - offset_2;
                const uint8_t* const repMatch2 = (repIndex2 < prefixStartIndex ? dictBase : base) + repIndex2;
                if ((((uint32_t)((prefixStartIndex - 1) - repIndex2) >= 3) & (offset_2 <= curr - dictStartIndex)) && (rd32(repMatch2) == rd32(ip))) {     // `curr`, as the reference has it (:676)
                    const uint8_t* const repEnd2 = repIndex2 < prefixStartIndex ? dictEnd : iend;
                    uint32_t const repLength2 = count_2segments(ip + 4, repMatch2 + 4, iend, repEnd2, prefixStart) + 4;
                    uint32_t const t = offset_2; offset_2 = offset_1; offset_1 = t;
                    sw.store(0, 0, repLength2 - 3);
                    hashTable[hash_ptr(ip, hlog, mls)] = current2;
                    ip += repLength2; anchor = ip;
                    continue;
                }
                break;
    }   }   }
    rep[0] = offset_1; rep[1] = offset_2;
    return (uint32_t)(iend - anchor);
}

// ZstdDoubleFast.cs:250 ZSTD_compressBlock_doubleFast_dictMatchState_generic
__device__ uint32_t dict_dfast_dms(uint32_t* hashLong, uint32_t hBitsL, uint32_t* hashSmall, uint32_t hBitsS,
                                   const uint32_t* dictHashLong, uint32_t dictHBitsL, const uint32_t* dictHashSmall, uint32_t dictHBitsS,
                                   const DictBlk& b, uint32_t rep[2], SeqWriter& sw)
{
    const uint8_t* const base = b.base; const uint8_t* const dictBase = b.dictBase; uint32_t const mls = b.mls;
    uint32_t const prefixLowestIndex = b.prefixStartIndex, dictStartIndex = b.dictStartIndex;
    const uint8_t* const prefixLowest = base + prefixLowestIndex; const uint8_t* const dictStart = dictBase + dictStartIndex; const uint8_t* const dictEnd = dictBase + prefixLowestIndex;
    const uint8_t* ip = b.istart; const uint8_t* anchor = b.istart; const uint8_t* const iend = b.iend; const uint8_t* const ilimit = iend - 8;
    uint32_t offset_1 = rep[0], offset_2 = rep[1];
    uint32_t const dictAndPrefixLength = (uint32_t)(ip - prefixLowest) + (prefixLowestIndex - dictStartIndex);
    ip += (dictAndPrefixLength == 0);
    while (ip < ilimit) {
        uint32_t mLength, offset;
        uint64_t const x = rd64(ip);
        uint32_t const h2 = hash_val(x, hBitsL, 8), h = hash_val(x, hBitsS, mls), dictHL = hash_val(x, dictHBitsL, 8), dictHS = hash_val(x, dictHBitsS, mls);
        uint32_t const curr = (uint32_t)(ip - base);
        uint32_t const matchIndexL = hashLong[h2]; uint32_t matchIndexS = hashSmall[h];
        const uint8_t* matchLong = base + matchIndexL; const uint8_t* match = base + matchIndexS;
        uint32_t const repIndex = curr + 1 - offset_1; const uint8_t* const repMatch = repIndex < prefixLowestIndex ? dictBase + repIndex : base + repIndex;
        hashLong[h2] = hashSmall[h] = curr;
        if (((uint32_t)((prefixLowestIndex - 1) - repIndex) >= 3) && (rd32(repMatch) == rd32(ip + 1))) {
            const uint8_t* const repMatchEnd = repIndex < prefixLowestIndex ? dictEnd : iend;
            mLength = count_2segments(ip + 1 + 4, repMatch + 4, iend, repMatchEnd, prefixLowest) + 4;
            ip++;
            sw.store((uint32_t)(ip - anchor), 0, mLength - 3);
            goto _match_stored;
        }
        if (matchIndexL > prefixLowestIndex) {
            if (rd64(matchLong) == x) {
                mLength = count_match(ip + 8, matchLong + 8, iend) + 8;
                offset = (uint32_t)(ip - matchLong);
                while (((ip > anchor) & (matchLong > prefixLowest)) && (ip[-1] == matchLong[-1])) { ip--; matchLong--; mLength++; }
                goto _match_found;
            }
        } else {
            uint32_t const dictMatchIndexL = dictHashLong[dictHL]; const uint8_t* dictMatchL = dictBase + dictMatchIndexL;
            if (dictMatchL > dictStart && rd64(dictMatchL) == x) {
                mLength = count_2segments(ip + 8, dictMatchL + 8, iend, dictEnd, prefixLowest) + 8;
                offset = curr - dictMatchIndexL;
                while (((ip > anchor) & (dictMatchL > dictStart)) && (ip[-1] == dictMatchL[-1])) { ip--; dictMatchL--; mLength++; }
                goto _match_found;
        }   }
        if (matchIndexS > prefixLowestIndex) {
            if (rd32(match) == (uint32_t)x) goto _search_next_long;
        } else {
            uint32_t const dictMatchIndexS = dictHashSmall[dictHS];
            match = dictBase + dictMatchIndexS; matchIndexS = dictMatchIndexS;
            if (match > dictStart && rd32(match) == (uint32_t)x) goto _search_next_long;
        }
        ip += ((ip - anchor) >> 8) + 1;
        continue;
_search_next_long:
        {   uint64_t const x1 = rd64(ip + 1);
            uint32_t const hl3 = hash_val(x1, hBitsL, 8), dictHLNext = hash_val(x1, dictHBitsL, 8);
            uint32_t const matchIndexL3 = hashLong[hl3]; const uint8_t* matchL3 = base + matchIndexL3;
            hashLong[hl3] = curr + 1;
            if (matchIndexL3 > prefixLowestIndex) {
                if (rd64(matchL3) == x1) {
                    mLength = count_match(ip + 9, matchL3 + 8, iend) + 8;
                    ip++;
                    offset = (uint32_t)(ip - matchL3);
                    while (((ip > anchor) & (matchL3 > prefixLowest)) && (ip[-1] == matchL3[-1])) { ip--; matchL3--; mLength++; }
                    goto _match_found;
                }
            } else {
                uint32_t const dictMatchIndexL3 = dictHashLong[dictHLNext]; const uint8_t* dictMatchL3 = dictBase + dictMatchIndexL3;
                if (dictMatchL3 > dictStart && rd64(dictMatchL3) == x1) {
                    mLength = count_2segments(ip + 1 + 8, dictMatchL3 + 8, iend, dictEnd, prefixLowest) + 8;
                    ip++;
                    offset = curr + 1 - dictMatchIndexL3;
                    while (((ip > anchor) & (dictMatchL3 > dictStart)) && (ip[-1] == dictMatchL3[-1])) { ip--; dictMatchL3--; mLength++; }
                    goto _match_found;
        }   }   }
        if (matchIndexS < prefixLowestIndex) {
            mLength = count_2segments(ip + 4, match + 4, iend, dictEnd, prefixLowest) + 4;
            offset = curr - matchIndexS;
            while (((ip > anchor) & (match > dictStart)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
        } else {
            mLength = count_match(ip + 4, match + 4, iend) + 4;
            offset = (uint32_t)(ip - match);
            while (((ip > anchor) & (match > prefixLowest)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
        }
_match_found:
        offset_2 = offset_1; offset_1 = offset;
        sw.store((uint32_t)(ip - anchor), offset + 2, mLength - 3);
_match_stored:
        ip += mLength; anchor = ip;
        if (ip <= ilimit) {
            {   uint32_t const indexToInsert = curr + 2;
                hashLong[hash_ptr(base + indexToInsert, hBitsL, 8)] = indexToInsert;
                hashLong[hash_ptr(ip - 2, hBitsL, 8)] = (uint32_t)(ip - 2 - base);
                hashSmall[hash_ptr(base + indexToInsert, hBitsS, mls)] = indexToInsert;
                hashSmall[hash_ptr(ip - 1, hBitsS, mls)] = (uint32_t)(ip - 1 - base);
            }
            while (ip <= ilimit) {
                uint32_t const current2 = (uint32_t)(ip - base), repIndex2 = current2 - offset_2;
                const uint8_t* const repMatch2 = repIndex2 < prefixLowestIndex ? dictBase + repIndex2 : base + repIndex2;
                if (((uint32_t)((prefixLowestIndex - 1) - repIndex2) >= 3) && (rd32(repMatch2) == rd32(ip))) {
                    const uint8_t* const repEnd2 = repIndex2 < prefixLowestIndex ? dictEnd : iend;
                    uint32_t const repLength2 = count_2segments(ip + 4, repMatch2 + 4, iend, repEnd2, prefixLowest) + 4;
                    uint32_t const t = offset_2; offset_2 = offset_1; offset_1 = t;
                    sw.store(0, 0, repLength2 - 3);
                    hashSmall[hash_ptr(ip, hBitsS, mls)] = current2;
                    hashLong[hash_ptr(ip, hBitsL, 8)] = current2;
                    ip += repLength2; anchor = ip;
                    continue;
                }
                break;
    }   }   }
    rep[0] = offset_1; rep[1] = offset_2;
    return (uint32_t)(iend - anchor);
}

// ZstdDoubleFast.cs:590 ZSTD_compressBlock_doubleFast_extDict_generic
__device__ uint32_t dict_dfast_ext(uint32_t* hashLong, uint32_t hBitsL, uint32_t* hashSmall, uint32_t hBitsS, const DictBlk& b, uint32_t rep[2], SeqWriter& sw)
{
    const uint8_t* const base = b.base; const uint8_t* const dictBase = b.dictBase; uint32_t const mls = b.mls;
    uint32_t const prefixStartIndex = b.prefixStartIndex, dictStartIndex = b.dictStartIndex;
    const uint8_t* const prefixStart = base + prefixStartIndex; const uint8_t* const dictStart = dictBase + dictStartIndex; const uint8_t* const dictEnd = dictBase + prefixStartIndex;
    const uint8_t* ip = b.istart; const uint8_t* anchor = b.istart; const uint8_t* const iend = b.iend; const uint8_t* const ilimit = iend - 8;
    uint32_t offset_1 = rep[0], offset_2 = rep[1];
    auto seg = [&](uint32_t idx) { return (idx < prefixStartIndex ? dictBase : base) + idx; };
    while (ip < ilimit) {
        uint64_t const x = rd64(ip);
        uint32_t const hSmall = hash_val(x, hBitsS, mls); uint32_t const matchIndex = hashSmall[hSmall]; const uint8_t* match = seg(matchIndex);
        uint32_t const hLong = hash_val(x, hBitsL, 8); uint32_t const matchLongIndex = hashLong[hLong]; const uint8_t* matchLong = seg(matchLongIndex);
        uint32_t const curr = (uint32_t)(ip - base), repIndex = curr + 1 - offset_1; const uint8_t* const repMatch = seg(repIndex);
        uint32_t mLength;
        hashSmall[hSmall] = hashLong[hLong] = curr;
        if ((((uint32_t)((prefixStartIndex - 1) - repIndex) >= 3) & (offset_1 <= curr + 1 - dictStartIndex)) && (rd32(repMatch) == rd32(ip + 1))) {
            const uint8_t* const repMatchEnd = repIndex < prefixStartIndex ? dictEnd : iend;
            mLength = count_2segments(ip + 1 + 4, repMatch + 4, iend, repMatchEnd, prefixStart) + 4;
            ip++;
            sw.store((uint32_t)(ip - anchor), 0, mLength - 3);
        } else {
            if ((matchLongIndex > dictStartIndex) && (rd64(matchLong) == x)) {
                const uint8_t* const matchEnd = matchLongIndex < prefixStartIndex ? dictEnd : iend;
                const uint8_t* const lowMatchPtr = matchLongIndex < prefixStartIndex ? dictStart : prefixStart;
                mLength = count_2segments(ip + 8, matchLong + 8, iend, matchEnd, prefixStart) + 8;
                uint32_t const offset = curr - matchLongIndex;
                while (((ip > anchor) & (matchLong > lowMatchPtr)) && (ip[-1] == matchLong[-1])) { ip--; matchLong--; mLength++; }
                offset_2 = offset_1; offset_1 = offset;
                sw.store((uint32_t)(ip - anchor), offset + 2, mLength - 3);
            } else if ((matchIndex > dictStartIndex) && (rd32(match) == (uint32_t)x)) {
                uint64_t const x1 = rd64(ip + 1);
                uint32_t const h3 = hash_val(x1, hBitsL, 8); uint32_t const matchIndex3 = hashLong[h3]; const uint8_t* match3 = seg(matchIndex3);
                uint32_t offset;
                hashLong[h3] = curr + 1;
                if ((matchIndex3 > dictStartIndex) && (rd64(match3) == x1)) {
                    const uint8_t* const matchEnd = matchIndex3 < prefixStartIndex ? dictEnd : iend;
                    const uint8_t* const lowMatchPtr = matchIndex3 < prefixStartIndex ? dictStart : prefixStart;
                    mLength = count_2segments(ip + 9, match3 + 8, iend, matchEnd, prefixStart) + 8;
                    ip++;
                    offset = curr + 1 - matchIndex3;
                    while (((ip > anchor) & (match3 > lowMatchPtr)) && (ip[-1] == match3[-1])) { ip--; match3--; mLength++; }
                } else {
                    const uint8_t* const matchEnd = matchIndex < prefixStartIndex ? dictEnd : iend;
                    const uint8_t* const lowMatchPtr = matchIndex < prefixStartIndex ? dictStart : prefixStart;
                    mLength = count_2segments(ip + 4, match + 4, iend, matchEnd, prefixStart) + 4;
                    offset = curr - matchIndex;
                    while (((ip > anchor) & (match > lowMatchPtr)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
                }
                offset_2 = offset_1; offset_1 = offset;
                sw.store((uint32_t)(ip - anchor), offset + 2, mLength - 3);
            } else { ip += ((ip - anchor) >> 8) + 1; continue; }
        }
        ip += mLength; anchor = ip;
        if (ip <= ilimit) {
            {   uint32_t const indexToInsert = curr + 2;
                hashLong[hash_ptr(base + indexToInsert, hBitsL, 8)] = indexToInsert;
                hashLong[hash_ptr(ip - 2, hBitsL, 8)] = (uint32_t)(ip - 2 - base);
                hashSmall[hash_ptr(base + indexToInsert, hBitsS, mls)] = indexToInsert;
                hashSmall[hash_ptr(ip - 1, hBitsS, mls)] = (uint32_t)(ip - 1 - base);
            }
            while (ip <= ilimit) {
                uint32_t const current2 = (uint32_t)(ip - base), repIndex2 = current2 - offset_2; const uint8_t* const repMatch2 = seg(repIndex2);
                if ((((uint32_t)((prefixStartIndex - 1) - repIndex2) >= 3) & (offset_2 <= current2 - dictStartIndex)) && (rd32(repMatch2) == rd32(ip))) {
                    const uint8_t* const repEnd2 = repIndex2 < prefixStartIndex ? dictEnd : iend;
                    uint32_t const repLength2 = count_2segments(ip + 4, repMatch2 + 4, iend, repEnd2, prefixStart) + 4;
                    uint32_t const t = offset_2; offset_2 = offset_1; offset_1 = t;
                    sw.store(0, 0, repLength2 - 3);
                    hashSmall[hash_ptr(ip, hBitsS, mls)] = current2;
                    hashLong[hash_ptr(ip, hBitsL, 8)] = current2;
                    ip += repLength2; anchor = ip;
                    continue;
                }
                break;
    }   }   }
    rep[0] = offset_1; rep[1] = offset_2;
    return (uint32_t)(iend - anchor);
}

// ZstdFast.cs:96 ZSTD_compressBlock_fast_noDict_generic, serial (blocks of a dictionary frame that no longer see the dictionary)
__device__ uint32_t nodict_fast(uint32_t* hashTable, uint32_t hlog, const DictBlk& b, uint32_t stepSize, uint32_t maxRep, uint32_t rep[2], SeqWriter& sw)
{
    const uint8_t* const base = b.base; uint32_t const mls = b.mls; uint32_t const prefixStartIndex = b.prefixStartIndex;
    const uint8_t* const prefixStart = base + prefixStartIndex;
    const uint8_t* const iend = b.iend; const uint8_t* const ilimit = iend - 8;
    const uint8_t* anchor = b.istart; const uint8_t* ip0 = b.istart; const uint8_t* ip1; const uint8_t* ip2; const uint8_t* ip3;
    uint32_t current0 = 0; uint32_t rep1 = rep[0], rep2 = rep[1], offsetSaved = 0;
    uint32_t hash0, hash1, idx, mval, offcode; const uint8_t* match0; uint32_t mLength; uint32_t step; const uint8_t* nextStep;
    ip0 += (ip0 == prefixStart);
    if (rep2 > maxRep) { offsetSaved = rep2; rep2 = 0; }          // :131-145
    if (rep1 > maxRep) { offsetSaved = rep1; rep1 = 0; }
_start:
    step = stepSize; nextStep = ip0 + 128;
    ip1 = ip0 + 1; ip2 = ip0 + step; ip3 = ip2 + 1;
    if (ip3 >= ilimit) goto _cleanup;
    hash0 = hash_ptr(ip0, hlog, mls); hash1 = hash_ptr(ip1, hlog, mls);
    idx = hashTable[hash0];
    do {
        uint32_t const rval = rep1 ? rd32(ip2 - rep1) : 0;
        current0 = (uint32_t)(ip0 - base);
        hashTable[hash0] = current0;
        if ((rep1 > 0) && (rd32(ip2) == rval)) {
            ip0 = ip2; match0 = ip0 - rep1;
            mLength = ip0[-1] == match0[-1];
            ip0 -= mLength; match0 -= mLength;
            offcode = 0; mLength += 4;
            goto _match;
        }
        mval = idx >= prefixStartIndex ? rd32(base + idx) : (rd32(ip0) ^ 1);
        if (rd32(ip0) == mval) goto _offset;
        idx = hashTable[hash1];
        hash0 = hash1; hash1 = hash_ptr(ip2, hlog, mls);
        ip0 = ip1; ip1 = ip2; ip2 = ip3;
        current0 = (uint32_t)(ip0 - base);
        hashTable[hash0] = current0;
        mval = idx >= prefixStartIndex ? rd32(base + idx) : (rd32(ip0) ^ 1);
        if (rd32(ip0) == mval) goto _offset;
        idx = hashTable[hash1];
        hash0 = hash1; hash1 = hash_ptr(ip2, hlog, mls);
        ip0 = ip1; ip1 = ip2; ip2 = ip0 + step; ip3 = ip1 + step;
        if (ip2 >= nextStep) { step++; nextStep += 128; }
    } while (ip3 < ilimit);
_cleanup:
    rep[0] = rep1 ? rep1 : offsetSaved;                 // :232-233
    rep[1] = rep2 ? rep2 : offsetSaved;
    return (uint32_t)(iend - anchor);
_offset:
    match0 = base + idx;
    rep2 = rep1; rep1 = (uint32_t)(ip0 - match0);
    offcode = rep1 + 2;
    mLength = 4;
    while (((ip0 > anchor) & (match0 > prefixStart)) && (ip0[-1] == match0[-1])) { ip0--; match0--; mLength++; }
_match:
    mLength += count_match(ip0 + mLength, match0 + mLength, iend);
    sw.store((uint32_t)(ip0 - anchor), offcode, mLength - 3);
    ip0 += mLength; anchor = ip0;
    if (ip1 < ip0) hashTable[hash1] = (uint32_t)(ip1 - base);
    if (ip0 <= ilimit) {
        hashTable[hash_ptr(base + current0 + 2, hlog, mls)] = current0 + 2;
        hashTable[hash_ptr(ip0 - 2, hlog, mls)] = (uint32_t)(ip0 - 2 - base);
        if (rep2 > 0) {
            while ((ip0 <= ilimit) && (rd32(ip0) == rd32(ip0 - rep2))) {
                uint32_t const rLength = count_match(ip0 + 4, ip0 + 4 - rep2, iend) + 4;
                { uint32_t const t = rep2; rep2 = rep1; rep1 = t; }
                hashTable[hash_ptr(ip0, hlog, mls)] = (uint32_t)(ip0 - base);
                ip0 += rLength;
                sw.store(0, 0, rLength - 3);
                anchor = ip0;
            }
        }
    }
    goto _start;
}

// ZstdDoubleFast.cs:51 ZSTD_compressBlock_doubleFast_noDict_generic, serial
__device__ uint32_t nodict_dfast(uint32_t* hashLong, uint32_t hBitsL, uint32_t* hashSmall, uint32_t hBitsS, const DictBlk& b, uint32_t maxRep, uint32_t rep[2], SeqWriter& sw)
{
    const uint8_t* const base = b.base; uint32_t const mls = b.mls; uint32_t const prefixLowestIndex = b.prefixStartIndex;
    const uint8_t* const prefixLowest = base + prefixLowestIndex;
    const uint8_t* const istart = b.istart; const uint8_t* const iend = b.iend; const uint8_t* const ilimit = iend - 8;
    const uint8_t* anchor = istart;
    uint32_t offset_1 = rep[0], offset_2 = rep[1], offsetSaved = 0;
    uint32_t mLength, offset, curr = 0;
    const uint8_t* nextStep; uint32_t step; uint32_t hl0, hl1 = 0; uint32_t idxl0, idxl1 = 0;
    const uint8_t* matchl0; const uint8_t* matchs0; const uint8_t* matchl1 = istart;
    const uint8_t* ip = istart; const uint8_t* ip1;
    ip += ((ip - prefixLowest) == 0);
    if (offset_2 > maxRep) { offsetSaved = offset_2; offset_2 = 0; }
    if (offset_1 > maxRep) { offsetSaved = offset_1; offset_1 = 0; }
    while (1) {
        step = 1; nextStep = ip + 256; ip1 = ip + step;
        if (ip1 > ilimit) goto _cleanup;
        hl0 = hash_ptr(ip, hBitsL, 8);
        idxl0 = hashLong[hl0]; matchl0 = base + idxl0;
        do {
            uint32_t const hs0 = hash_ptr(ip, hBitsS, mls);
            uint32_t const idxs0 = hashSmall[hs0];
            curr = (uint32_t)(ip - base);
            matchs0 = base + idxs0;
            hashLong[hl0] = hashSmall[hs0] = curr;
            if ((offset_1 > 0) && (rd32(ip + 1 - offset_1) == rd32(ip + 1))) {
                mLength = count_match(ip + 1 + 4, ip + 1 + 4 - offset_1, iend) + 4;
                ip++;
                sw.store((uint32_t)(ip - anchor), 0, mLength - 3);
                goto _match_stored;
            }
            hl1 = hash_ptr(ip1, hBitsL, 8);
            if (idxl0 > prefixLowestIndex) {
                if (rd64(matchl0) == rd64(ip)) {
                    mLength = count_match(ip + 8, matchl0 + 8, iend) + 8;
                    offset = (uint32_t)(ip - matchl0);
                    while (((ip > anchor) & (matchl0 > prefixLowest)) && (ip[-1] == matchl0[-1])) { ip--; matchl0--; mLength++; }
                    goto _match_found;
                }
            }
            idxl1 = hashLong[hl1]; matchl1 = base + idxl1;
            if (idxs0 > prefixLowestIndex) {
                if (rd32(matchs0) == rd32(ip)) goto _search_next_long;
            }
            if (ip1 >= nextStep) { step++; nextStep += 256; }
            ip = ip1; ip1 += step;
            hl0 = hl1; idxl0 = idxl1; matchl0 = matchl1;
        } while (ip1 <= ilimit);
_cleanup:
        rep[0] = offset_1 ? offset_1 : offsetSaved;     // :215-216
        rep[1] = offset_2 ? offset_2 : offsetSaved;
        return (uint32_t)(iend - anchor);
_search_next_long:
        if (idxl1 > prefixLowestIndex) {
            if (rd64(matchl1) == rd64(ip1)) {
                ip = ip1;
                mLength = count_match(ip + 8, matchl1 + 8, iend) + 8;
                offset = (uint32_t)(ip - matchl1);
                while (((ip > anchor) & (matchl1 > prefixLowest)) && (ip[-1] == matchl1[-1])) { ip--; matchl1--; mLength++; }
                goto _match_found;
            }
        }
        mLength = count_match(ip + 4, matchs0 + 4, iend) + 4;
        offset = (uint32_t)(ip - matchs0);
        while (((ip > anchor) & (matchs0 > prefixLowest)) && (ip[-1] == matchs0[-1])) { ip--; matchs0--; mLength++; }
_match_found:
        offset_2 = offset_1; offset_1 = offset;
        if (step < 4) hashLong[hl1] = (uint32_t)(ip1 - base);
        sw.store((uint32_t)(ip - anchor), offset + 2, mLength - 3);
_match_stored:
        ip += mLength; anchor = ip;
        if (ip <= ilimit) {
            {   uint32_t const indexToInsert = curr + 2;
                hashLong[hash_ptr(base + indexToInsert, hBitsL, 8)] = indexToInsert;
                hashLong[hash_ptr(ip - 2, hBitsL, 8)] = (uint32_t)(ip - 2 - base);
                hashSmall[hash_ptr(base + indexToInsert, hBitsS, mls)] = indexToInsert;
                hashSmall[hash_ptr(ip - 1, hBitsS, mls)] = (uint32_t)(ip - 1 - base);
            }
            while ((ip <= ilimit) && ((offset_2 > 0) && (rd32(ip) == rd32(ip - offset_2)))) {
                uint32_t const rLength = count_match(ip + 4, ip + 4 - offset_2, iend) + 4;
                uint32_t const t = offset_2; offset_2 = offset_1; offset_1 = t;
                hashSmall[hash_ptr(ip, hBitsS, mls)] = (uint32_t)(ip - base);
                hashLong[hash_ptr(ip, hBitsL, 8)] = (uint32_t)(ip - base);
                sw.store(0, 0, rLength - 3);
                ip += rLength; anchor = ip;
            }
        }
    }
}

// Window state of block `wave` of a dictionary frame: ZSTD_checkDictValidity (ZstdCompressInternal.cs:697) on the block's end, ZSTD_window_enforceMaxDist
// (:659) on its start, then the variant ZSTD_matchState_dictMode / ZSTD_selectBlockCompressor pick (:576, ZstdCompress.cs:3398) and its two limits.
struct DictBlkMode { uint32_t lowLimit, dictLimit, loadedDictEnd, dms, mode /* 0 no dictionary, 1 dictMatchState, 2 extDict */, dictStartIndex, prefixStartIndex; };
__device__ __forceinline__ DictBlkMode dict_block_mode(const EncItem& it, uint32_t wave)
{
    DictBlkMode m;
    uint32_t const blkStart = wave * kBlockSizeMax, blkSize = min(kBlockSizeMax, it.srcSize - blkStart), maxDist = 1u << it.windowLog;
    uint32_t const startIdx = it.dPrefix + blkStart, endIdx = startIdx + blkSize;
    m.lowLimit = it.wLow; m.dictLimit = it.wDictLimit; m.loadedDictEnd = it.loadedDictEnd; m.dms = it.dms;
    if (endIdx > m.loadedDictEnd + maxDist) { m.loadedDictEnd = 0; m.dms = 0; }
    if (startIdx > maxDist + m.loadedDictEnd) {
        uint32_t const newLowLimit = startIdx - maxDist;
        if (m.lowLimit < newLowLimit) m.lowLimit = newLowLimit;
        if (m.dictLimit < m.lowLimit) m.dictLimit = m.lowLimit;
        m.loadedDictEnd = 0; m.dms = 0;
    }
    auto lowest = [&](uint32_t lowestValid, uint32_t curr) { return m.loadedDictEnd != 0 ? lowestValid : ((curr - lowestValid > maxDist) ? curr - maxDist : lowestValid); };   // ZSTD_getLowestMatchIndex / ..PrefixIndex (:787, :802)
    bool const extDict = m.lowLimit < m.dictLimit;
    uint32_t const extLow = lowest(m.lowLimit, endIdx), extPrefix = m.dictLimit < extLow ? extLow : m.dictLimit;
    if (extDict && extPrefix != extLow) { m.mode = 2; m.dictStartIndex = extLow; m.prefixStartIndex = extPrefix; }
    else if (!extDict && m.dms) { m.mode = 1; m.dictStartIndex = 2; m.prefixStartIndex = m.dictLimit; }     // loadedDictEnd != 0 here: ZSTD_getLowestPrefixIndex is dictLimit
    else { m.mode = 0; m.prefixStartIndex = lowest(m.dictLimit, endIdx); m.dictStartIndex = m.prefixStartIndex; }
    return m;
}
constexpr uint32_t kDictSerial = 8;     // EncItem::dBlkMode bit: this block is left to a later kernel (ZSTD_dfast group kernel, serial kernel)
constexpr uint32_t kDictForced = 16;    // developer knob ZSTDB200_DICT_SERIAL: the serial kernel takes everything

// ------------------------------------------------------------------------------------------------------------
//  Group-per-frame ZSTD_fast parse with a dictionary (ZstdFast.cs:390 dictMatchState, :583 extDict): the same construction as
//  enc_match_group_kernel, over the classic loop of the dictionary variants.  That loop visits one position per step,
//  ip += ((ip - anchor) >> 8) + stepSize, until the repcode at ip + 1 or the table candidate at ip matches, so lane l of a group
//  evaluates the l-th position of the schedule speculatively: it sees the window's earlier positions through __match_any_sync and
//  older ones through the table, the first event in the reference's order wins and only the writes up to it are committed.
//  Positions are the reference's indices: index < prefixStartIndex lives in the dictionary content (`dictBase + index`), the rest
//  in the frame (`base + index`), and a match runs from the first segment into the second as if they were adjacent
//  (ZSTD_count_2segments, ZstdCompressInternal.cs:283).  This kernel also does the window bookkeeping of the block for every dictionary
//  frame of the pass; ZSTD_dfast frames go on to enc_match_dict_dfast_group_kernel, blocks that no longer see the dictionary to the
//  serial kernel (enc_match_dict_kernel), through EncItem::dBlkMode.
// ------------------------------------------------------------------------------------------------------------
template <int GS>
__global__ void __launch_bounds__(32 * kMatchWarps) enc_match_dict_fast_group_kernel(EncPass p, const uint32_t* __restrict__ workList, uint32_t nWork, uint32_t wave, uint32_t forceSerial)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr int NG = 32 / GS;
    constexpr uint32_t LOW = GS == 32 ? FULL : ((1u << (GS & 31)) - 1u);
    uint32_t const lane = threadIdx.x & 31, g = lane / GS, l = lane % GS, gbase = g * GS;
    uint32_t const gmask = LOW << gbase;
    uint32_t const wi = (blockIdx.x * kMatchWarps + (threadIdx.x >> 5)) * NG + g;
    bool active = wi < nWork;
    uint32_t const item = workList[active ? wi : 0];
    EncItem& it = p.items[item];
    DictBlkMode const bm = dict_block_mode(it, wave);
    bool const serial = forceSerial || bm.mode == 0 || it.strategy != 1;
    __syncwarp();                                      // every lane has read the window state before lane 0 of its group advances it
    if (active && l == 0) {
        it.wLow = bm.lowLimit; it.wDictLimit = bm.dictLimit; it.loadedDictEnd = bm.loadedDictEnd; it.dms = bm.dms;
        it.dBlkMode = bm.mode | (serial ? kDictSerial : 0u) | (forceSerial ? kDictForced : 0u); it.dBlkLow = bm.dictStartIndex; it.dBlkPrefix = bm.prefixStartIndex;
    }
    active = active && !serial;
    bool const isDms = bm.mode == 1;
    uint32_t const hlog = it.hashLog, mls = it.minMatch, dictHLog = p.dict.hashLog;
    const uint8_t* const src = p.src + it.srcOff;
    uint32_t const dP = it.dPrefix;                    // index of src[0]
    const uint8_t* const base = src - dP; const uint8_t* const dictBase = p.dict.content - 2;
    uint32_t const prefixStartIndex = bm.prefixStartIndex, dictStartIndex = bm.dictStartIndex;
    uint32_t* const T = p.tables + it.tableOff; const uint32_t* const DT = p.dict.tables;
    uint32_t* const oLL = p.seqLL + (size_t)item * p.seqStride;
    uint32_t* const oML = p.seqML + (size_t)item * p.seqStride;
    uint32_t* const oOF = p.seqOF + (size_t)item * p.seqStride;
    int const blkStart = (int)(wave * kBlockSizeMax), blkEnd = active ? blkStart + (int)min(kBlockSizeMax, it.srcSize - (uint32_t)blkStart) : blkStart + 64;
    int const ilimit = blkEnd - 8;
    int const stepSize = active ? (int)it.dStep : 1;
    int ip = blkStart, anchor = blkStart;
    if (isDms && ((uint32_t)blkStart + dP - prefixStartIndex) + (prefixStartIndex - dictStartIndex) == 0) ip++;      // ip += (dictAndPrefixLength == 0)
    uint32_t offset_1 = it.rep[0], offset_2 = it.rep[1], nseq = 0, currPrev = 0;
    bool afterMatch = false;                           // the repcode loop behind a match (:499-520, :670-690) is still open at ip
    auto gballot = [&](bool pr) -> uint32_t { return (__ballot_sync(FULL, pr) >> gbase) & LOW; };
    auto ptrOf = [&](uint32_t idx) -> const uint8_t* { return idx < prefixStartIndex ? dictBase + idx : base + idx; };
    auto rd32x = [&](uint32_t idx) -> uint32_t {       // four bytes at an index, also across the segment boundary
        if (idx + 4 <= prefixStartIndex || idx >= prefixStartIndex) return rd32(ptrOf(idx));
        uint32_t v = 0;
        for (int j = 0; j < 4; j++) v |= (uint32_t)*ptrOf(idx + j) << (8 * j);
        return v;
    };
    // the guard in front of both repcode tests: not inside the last three bytes of the dictionary segment; extDict: inside the window
    auto repOk = [&](uint32_t repIndex, uint32_t off, uint32_t span) { return ((uint32_t)((prefixStartIndex - 1) - repIndex) >= 3) && (isDms || off <= span); };
    while (__any_sync(FULL, active)) {
        // ---- this lane's position: the l-th of the schedule ----
        int pl, P;
        {
            int const s0 = ((ip - anchor) >> 8) + stepSize;
            bool const simple = !active || (((ip + GS * s0 - anchor) >> 8) == ((ip - anchor) >> 8));
            if (__all_sync(FULL, simple)) { pl = ip + (int)l * s0; P = ip + GS * s0; }
            else {
                P = ip; pl = ip;
#pragma unroll
                for (int j = 0; j < GS; j++) { if (j == (int)l) pl = P; P += ((P - anchor) >> 8) + stepSize; }
            }
        }
        bool const vk = active && pl < ilimit;
        uint32_t const validMask = gballot(vk);
        // ---- the open repcode loop at ip (same answer in every lane of the group) ----
        bool r2hit = false; uint32_t repIndex2 = 0;
        if (active && afterMatch && ip <= ilimit) {
            uint32_t const current2 = (uint32_t)ip + dP;
            repIndex2 = current2 - offset_2;
            r2hit = repOk(repIndex2, offset_2, currPrev - dictStartIndex) && rd32(ptrOf(repIndex2)) == rd32(src + ip);       // `curr` of the last search, as the reference has it (:676)
        }
        // ---- probes ----
        uint64_t const x = vk ? rd64(src + pl) : 0ull;
        uint32_t const cur4 = (uint32_t)x, next4 = (uint32_t)(x >> 8);
        uint32_t const curr = (uint32_t)pl + dP;
        uint32_t const repIndex = curr + 1 - offset_1;
        bool const repHit = vk && repOk(repIndex, offset_1, curr + 1 - dictStartIndex) && rd32(ptrOf(repIndex)) == next4;
        uint32_t const h = hash_val(x, hlog, mls);
        uint32_t const tv = vk ? __ldcg(T + h) : 0u;
        uint32_t const peers = (__match_any_sync(FULL, vk ? (h | (g << 24)) : (0x80000000u | lane)) >> gbase) & LOW;
        uint32_t const lower = peers & ((1u << l) - 1u);
        int const cl = lower ? 31 - __clz((int)lower) : (int)l;
        uint32_t const cqIdx = __shfl_sync(FULL, curr, gbase + cl);
        uint32_t const cq4 = __shfl_sync(FULL, cur4, gbase + cl);      // a candidate forwarded from a lower lane of the window: its bytes are in that lane's registers
        uint32_t cand = lower ? cqIdx : tv;
        bool hit;
        // a forwarded candidate is a position of this block, inside the prefix -- but dictMatchState treats the prefix's FIRST position like an
        // empty cell (`matchIndex <= prefixStartIndex`, :441) and asks the dictionary instead (found by the soak: inputs that repeat their first bytes)
        if (lower && !(isDms && cand <= prefixStartIndex)) hit = vk && cq4 == cur4;
        else if (isDms) {
            if (cand <= prefixStartIndex) {                            // nothing usable in the frame's own table: the dictionary's (:441-449)
                cand = vk ? __ldg(DT + hash_val(x, dictHLog, mls)) : 0u;
                hit = vk && cand > dictStartIndex && rd32(dictBase + cand) == cur4;
            } else hit = vk && rd32(base + cand) == cur4;
        } else hit = vk && cand >= dictStartIndex && rd32(ptrOf(cand)) == cur4;
        uint32_t key = 0xFFFFFFFFu;                                    // 0: repcode loop at ip; 1 + 2l: repcode at position l + 1; 2 + 2l: table candidate at position l
        if (hit) key = 2 * l + 2;
        if (repHit) key = 2 * l + 1;
        if (r2hit) key = 0;
        uint32_t const best = __reduce_min_sync(gmask, key);
        bool const ev = active && best != 0xFFFFFFFFu;
        int const type = !ev ? -1 : (best == 0 ? 3 : (int)((best - 1) & 1u));       // 0 repcode, 1 table candidate, 3 repcode loop
        uint32_t const le = (ev && best) ? (best - 1) >> 1 : 0u;
        if (active && type != 3) afterMatch = false;
        // ---- table writes of the positions visited up to the event (`hashTable[h] = curr` precedes both tests); the latest position of a bucket wins ----
        {
            uint32_t const lastLane = type < 0 ? (uint32_t)GS - 1 : le;
            uint32_t const peersC = peers & (lastLane >= 31 ? FULL : ((2u << lastLane) - 1u));
            if (vk && type != 3 && l <= lastLane && ((peersC >> l) >> 1) == 0) T[h] = curr;
        }
        if (type == 3 && l == 0) T[hash_val(rd64(src + ip), hlog, mls)] = (uint32_t)ip + dP;
        // ---- match geometry (group-uniform): position in the frame, index of its source ----
        int const pe = __shfl_sync(FULL, pl, gbase + le);
        uint32_t const ce = __shfl_sync(FULL, cand, gbase + le), re = __shfl_sync(FULL, repIndex, gbase + le);
        int mpos = 0, mlen = 0; uint32_t msrc = 0, offcode = 0;
        if (type == 3) { mpos = ip; msrc = repIndex2; mlen = 4; uint32_t const t = offset_2; offset_2 = offset_1; offset_1 = t; }
        else if (type == 0) { mpos = pe + 1; msrc = re; mlen = 4; currPrev = (uint32_t)pe + dP; }
        else if (type == 1) { mpos = pe; msrc = ce; mlen = 4; currPrev = (uint32_t)pe + dP; offset_2 = offset_1; offset_1 = currPrev - ce; offcode = offset_1 + 2; }
        // backward extension of table matches (:462-467, :648-653): not below the anchor, not below the start of the source's segment
        {
            bool ext = type == 1;
            uint32_t const lowIdx = msrc < prefixStartIndex ? dictStartIndex : prefixStartIndex;
            while (__any_sync(FULL, ext)) {
                int const a = mpos - 1 - (int)l; uint32_t const b = msrc - 1 - l;
                bool const ok = ext && a >= anchor && msrc >= lowIdx + 1 + l && src[a] == *ptrOf(b);
                uint32_t const okm = gballot(ok);
                uint32_t const n = okm == LOW ? (uint32_t)GS : (uint32_t)__ffs((int)~okm) - 1u;
                if (ext) { mpos -= (int)n; msrc -= n; mlen += (int)n; ext = n == (uint32_t)GS; }
            }
        }
        // forward extension (ZSTD_count_2segments): lane l compares 4 bytes, 4*GS bytes per round; the source is read by index
        {
            bool cnt = ev;
            while (__any_sync(FULL, cnt)) {
                int const pa = mpos + mlen + 4 * (int)l; uint32_t const ib = msrc + (uint32_t)mlen + 4 * l;
                int const rem = blkEnd - pa;
                uint32_t n = 4;
                if (cnt) {
                    if (rem >= 4) { uint32_t const diff = rd32(src + pa) ^ rd32x(ib); n = diff ? (uint32_t)(__ffs((int)diff) - 1) >> 3 : 4u; }
                    else { n = 0; for (int j = 0; j < rem; j++) { if (src[pa + j] == *ptrOf(ib + j)) n++; else break; } }
                }
                uint32_t const notFull = gballot(cnt && n != 4);
                uint32_t const f = notFull ? (uint32_t)__ffs((int)notFull) - 1u : 0u;
                uint32_t const nf = __shfl_sync(FULL, n, gbase + f);
                if (cnt) { if (notFull) { mlen += 4 * (int)f + (int)nf; cnt = false; } else mlen += 4 * GS; }
            }
        }
        // ---- sequence, post-match inserts (:490-498, :662-669), next state ----
        if (ev) {
            ZB_ASSERT(nseq < p.seqStride && mpos >= anchor && msrc < (uint32_t)mpos + dP && mpos + mlen <= blkEnd);
            if (l == 0) { oLL[nseq] = (uint32_t)(mpos - anchor); oOF[nseq] = offcode + 1; oML[nseq] = (uint32_t)mlen - 3; }
            nseq++;
            int const mend = mpos + mlen;
            if (l == 0 && type != 3 && mend <= ilimit) {
                int const c2 = (int)(currPrev - dP) + 2;
                T[hash_val(rd64(src + c2), hlog, mls)] = currPrev + 2;
                T[hash_val(rd64(src + mend - 2), hlog, mls)] = (uint32_t)(mend - 2) + dP;
            }
            ip = mend; anchor = mend;
            afterMatch = mend <= ilimit;               // the repcode loop only opens `if (ip <= ilimit)`
        } else if (active) {
            if (validMask != LOW) active = false;      // the loop condition failed inside the window
            else ip = P;
        }
        __syncwarp();
    }
    if (wi < nWork && !serial && l == 0) { it.nbSeq = nseq; it.lastLL = (uint32_t)(blkEnd - anchor); it.repNext[0] = offset_1; it.repNext[1] = offset_2; }
}

// ------------------------------------------------------------------------------------------------------------
//  Group-per-frame ZSTD_dfast parse with a dictionary (ZstdDoubleFast.cs:250 dictMatchState, :590 extDict): as above, with the two
//  tables.  Order of the tests at a position: repcode at ip + 1, long table (8 bytes) at ip, short table (4 bytes) at ip -- and a short
//  hit first looks the long table up at ip + 1 (`_search_next_long`), which only the winning position does, after the window's
//  writes up to it have been committed.
// ------------------------------------------------------------------------------------------------------------
template <int GS>
__global__ void __launch_bounds__(32 * kMatchWarps) enc_match_dict_dfast_group_kernel(EncPass p, const uint32_t* __restrict__ workList, uint32_t nWork, uint32_t wave)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr int NG = 32 / GS;
    constexpr uint32_t LOW = GS == 32 ? FULL : ((1u << (GS & 31)) - 1u);
    uint32_t const lane = threadIdx.x & 31, g = lane / GS, l = lane % GS, gbase = g * GS;
    uint32_t const gmask = LOW << gbase;
    uint32_t const wi = (blockIdx.x * kMatchWarps + (threadIdx.x >> 5)) * NG + g;
    bool active = wi < nWork;
    uint32_t const item = workList[active ? wi : 0];
    EncItem& it = p.items[item];
    // the window bookkeeping of the block was done by enc_match_dict_fast_group_kernel; this kernel takes the ZSTD_dfast blocks that still see the dictionary
    uint32_t const mode = it.dBlkMode & 3u;
    bool const mine = active && it.strategy == 2 && mode != 0 && (it.dBlkMode & kDictSerial) && !(it.dBlkMode & kDictForced);
    active = mine;
    __syncwarp();
    if (mine && l == 0) it.dBlkMode &= ~kDictSerial;                 // taken here: the serial kernel skips it
    bool const isDms = mode == 1;
    uint32_t const hBitsL = it.hashLog, hBitsS = it.chainLog, mls = it.minMatch, dictHBitsL = p.dict.hashLog, dictHBitsS = p.dict.chainLog;
    const uint8_t* const src = p.src + it.srcOff;
    uint32_t const dP = it.dPrefix;
    const uint8_t* const base = src - dP; const uint8_t* const dictBase = p.dict.content - 2;
    uint32_t const prefixStartIndex = it.dBlkPrefix, dictStartIndex = it.dBlkLow;
    uint32_t* const TL = p.tables + it.tableOff; uint32_t* const TS = TL + ((size_t)1 << hBitsL);
    const uint32_t* const DTL = p.dict.tables; const uint32_t* const DTS = DTL + ((size_t)1 << dictHBitsL);
    uint32_t* const oLL = p.seqLL + (size_t)item * p.seqStride;
    uint32_t* const oML = p.seqML + (size_t)item * p.seqStride;
    uint32_t* const oOF = p.seqOF + (size_t)item * p.seqStride;
    int const blkStart = (int)(wave * kBlockSizeMax), blkEnd = active ? blkStart + (int)min(kBlockSizeMax, it.srcSize - (uint32_t)blkStart) : blkStart + 64;
    int const ilimit = blkEnd - 8;
    int ip = blkStart, anchor = blkStart;
    if (isDms && ((uint32_t)blkStart + dP - prefixStartIndex) + (prefixStartIndex - dictStartIndex) == 0) ip++;
    uint32_t offset_1 = it.rep[0], offset_2 = it.rep[1], nseq = 0, currPrev = 0;
    bool afterMatch = false;
    auto gballot = [&](bool pr) -> uint32_t { return (__ballot_sync(FULL, pr) >> gbase) & LOW; };
    auto ptrOf = [&](uint32_t idx) -> const uint8_t* { return idx < prefixStartIndex ? dictBase + idx : base + idx; };
    auto rd32x = [&](uint32_t idx) -> uint32_t {
        if (idx + 4 <= prefixStartIndex || idx >= prefixStartIndex) return rd32(ptrOf(idx));
        uint32_t v = 0;
        for (int j = 0; j < 4; j++) v |= (uint32_t)*ptrOf(idx + j) << (8 * j);
        return v;
    };
    auto repOk = [&](uint32_t repIndex, uint32_t off, uint32_t span) { return ((uint32_t)((prefixStartIndex - 1) - repIndex) >= 3) && (isDms || off <= span); };
    auto hashL = [&](uint64_t v, uint32_t bits) { return (uint32_t)((v * 0xCF1BBCDCB7A56463ull) >> (64 - bits)); };
    // a long-table candidate for the 8 bytes v: (dms) outside the frame's own range the dictionary's table answers (:296-330, :367-399)
    auto longHit = [&](uint32_t& cand, uint64_t v) -> bool {
        if (isDms) {
            if (cand > prefixStartIndex) return rd64(base + cand) == v;
            cand = __ldg(DTL + hashL(v, dictHBitsL));
            return cand > dictStartIndex && rd64(dictBase + cand) == v;
        }
        return cand > dictStartIndex && rd64(ptrOf(cand)) == v;
    };
    while (__any_sync(FULL, active)) {
        int pl, P;
        {
            int const s0 = ((ip - anchor) >> 8) + 1;
            bool const simple = !active || (((ip + GS * s0 - anchor) >> 8) == ((ip - anchor) >> 8));
            if (__all_sync(FULL, simple)) { pl = ip + (int)l * s0; P = ip + GS * s0; }
            else {
                P = ip; pl = ip;
#pragma unroll
                for (int j = 0; j < GS; j++) { if (j == (int)l) pl = P; P += ((P - anchor) >> 8) + 1; }
            }
        }
        bool const vk = active && pl < ilimit;
        uint32_t const validMask = gballot(vk);
        bool r2hit = false; uint32_t repIndex2 = 0;
        if (active && afterMatch && ip <= ilimit) {
            uint32_t const current2 = (uint32_t)ip + dP;
            repIndex2 = current2 - offset_2;
            r2hit = repOk(repIndex2, offset_2, current2 - dictStartIndex) && rd32(ptrOf(repIndex2)) == rd32(src + ip);      // current2 here (:739), unlike ZSTD_fast
        }
        // ---- probes: both tables, each with the window's earlier positions forwarded ----
        uint64_t const x = vk ? rd64(src + pl) : 0ull;
        uint32_t const cur4 = (uint32_t)x, next4 = (uint32_t)(x >> 8);
        uint32_t const curr = (uint32_t)pl + dP;
        uint32_t const repIndex = curr + 1 - offset_1;
        bool const repHit = vk && repOk(repIndex, offset_1, curr + 1 - dictStartIndex) && rd32(ptrOf(repIndex)) == next4;
        uint32_t const hL = hashL(x, hBitsL), hS = hash_val(x, hBitsS, mls);
        uint32_t const tL = vk ? __ldcg(TL + hL) : 0u, tS = vk ? __ldcg(TS + hS) : 0u;
        uint32_t const peersL = (__match_any_sync(FULL, vk ? (hL | (g << 24)) : (0x80000000u | lane)) >> gbase) & LOW;
        uint32_t const peersS = (__match_any_sync(FULL, vk ? (hS | (g << 24)) : (0x80000000u | lane)) >> gbase) & LOW;
        uint32_t const lowerL = peersL & ((1u << l) - 1u), lowerS = peersS & ((1u << l) - 1u);
        int const clL = lowerL ? 31 - __clz((int)lowerL) : (int)l, clS = lowerS ? 31 - __clz((int)lowerS) : (int)l;
        uint32_t const fL = __shfl_sync(FULL, curr, gbase + clL), fS = __shfl_sync(FULL, curr, gbase + clS);
        uint32_t const fxLo = __shfl_sync(FULL, cur4, gbase + clL), fxHi = __shfl_sync(FULL, (uint32_t)(x >> 32), gbase + clL);
        uint32_t const fS4 = __shfl_sync(FULL, cur4, gbase + clS);
        uint32_t candL = lowerL ? fL : tL, candS = lowerS ? fS : tS;
        bool hitL, hitS;
        // (dictMatchState: a candidate at the prefix's first position counts as an empty cell, `matchIndex > prefixLowestIndex` :296, :333)
        if (lowerL && !(isDms && candL <= prefixStartIndex)) hitL = vk && fxLo == cur4 && fxHi == (uint32_t)(x >> 32);
        else hitL = vk && longHit(candL, x);
        if (lowerS && !(isDms && candS <= prefixStartIndex)) hitS = vk && fS4 == cur4;
        else if (isDms) {
            if (candS > prefixStartIndex) hitS = vk && rd32(base + candS) == cur4;
            else { candS = vk ? __ldg(DTS + hash_val(x, dictHBitsS, mls)) : 0u; hitS = vk && candS > dictStartIndex && rd32(dictBase + candS) == cur4; }
        } else hitS = vk && candS > dictStartIndex && rd32(ptrOf(candS)) == cur4;
        uint32_t key = 0xFFFFFFFFu;                                    // 0: repcode loop; 1 + 3l repcode at l + 1; 2 + 3l long at l; 3 + 3l short at l
        if (hitS) key = 3 * l + 3;
        if (hitL) key = 3 * l + 2;
        if (repHit) key = 3 * l + 1;
        if (r2hit) key = 0;
        uint32_t const best = __reduce_min_sync(gmask, key);
        bool const ev = active && best != 0xFFFFFFFFu;
        int type = !ev ? -1 : (best == 0 ? 3 : (int)((best - 1) % 3));   // 0 repcode, 1 long, 2 short, 3 repcode loop
        uint32_t const le = (ev && best) ? (best - 1) / 3 : 0u;
        if (active && type != 3) afterMatch = false;
        {
            uint32_t const lastLane = type < 0 ? (uint32_t)GS - 1 : le;
            uint32_t const upTo = lastLane >= 31 ? FULL : ((2u << lastLane) - 1u);
            if (vk && type != 3 && l <= lastLane) {
                if ((((peersL & upTo) >> l) >> 1) == 0) TL[hL] = curr;
                if ((((peersS & upTo) >> l) >> 1) == 0) TS[hS] = curr;
            }
        }
        if (type == 3 && l == 0) { uint64_t const xi = rd64(src + ip); TS[hash_val(xi, hBitsS, mls)] = (uint32_t)ip + dP; TL[hashL(xi, hBitsL)] = (uint32_t)ip + dP; }
        __syncwarp();
        int const pe = __shfl_sync(FULL, pl, gbase + le);
        uint32_t const cLe = __shfl_sync(FULL, candL, gbase + le), cSe = __shfl_sync(FULL, candS, gbase + le), re = __shfl_sync(FULL, repIndex, gbase + le);
        int mpos = 0, mlen = 0; uint32_t msrc = 0, offcode = 0;
        bool back = false;
        if (type == 3) { mpos = ip; msrc = repIndex2; mlen = 4; uint32_t const t = offset_2; offset_2 = offset_1; offset_1 = t; }
        else if (type == 0) { mpos = pe + 1; msrc = re; mlen = 4; currPrev = (uint32_t)pe + dP; }
        else if (type == 1) { mpos = pe; msrc = cLe; mlen = 8; currPrev = (uint32_t)pe + dP; back = true; }
        else if (type == 2) {
            // _search_next_long: the long table at ip + 1, read after the writes of the positions up to ip, then written (:340-399, :642-660)
            currPrev = (uint32_t)pe + dP;
            uint64_t const x1 = rd64(src + pe + 1);
            uint32_t const h3 = hashL(x1, hBitsL);
            uint32_t m3 = __ldcg(TL + h3);
            if (l == 0) TL[h3] = currPrev + 1;
            if (longHit(m3, x1)) { mpos = pe + 1; msrc = m3; mlen = 8; }
            else { mpos = pe; msrc = cSe; mlen = 4; }
            back = true;
        }
        if (back) { offset_2 = offset_1; offset_1 = ((uint32_t)mpos + dP) - msrc; offcode = offset_1 + 2; }
        {
            bool ext = back;
            uint32_t const lowIdx = msrc < prefixStartIndex ? dictStartIndex : prefixStartIndex;
            while (__any_sync(FULL, ext)) {
                int const a = mpos - 1 - (int)l; uint32_t const b = msrc - 1 - l;
                bool const ok = ext && a >= anchor && msrc >= lowIdx + 1 + l && src[a] == *ptrOf(b);
                uint32_t const okm = gballot(ok);
                uint32_t const n = okm == LOW ? (uint32_t)GS : (uint32_t)__ffs((int)~okm) - 1u;
                if (ext) { mpos -= (int)n; msrc -= n; mlen += (int)n; ext = n == (uint32_t)GS; }
            }
        }
        {
            bool cnt = ev;
            while (__any_sync(FULL, cnt)) {
                int const pa = mpos + mlen + 4 * (int)l; uint32_t const ib = msrc + (uint32_t)mlen + 4 * l;
                int const rem = blkEnd - pa;
                uint32_t n = 4;
                if (cnt) {
                    if (rem >= 4) { uint32_t const diff = rd32(src + pa) ^ rd32x(ib); n = diff ? (uint32_t)(__ffs((int)diff) - 1) >> 3 : 4u; }
                    else { n = 0; for (int j = 0; j < rem; j++) { if (src[pa + j] == *ptrOf(ib + j)) n++; else break; } }
                }
                uint32_t const notFull = gballot(cnt && n != 4);
                uint32_t const f = notFull ? (uint32_t)__ffs((int)notFull) - 1u : 0u;
                uint32_t const nf = __shfl_sync(FULL, n, gbase + f);
                if (cnt) { if (notFull) { mlen += 4 * (int)f + (int)nf; cnt = false; } else mlen += 4 * GS; }
            }
        }
        if (ev) {
            ZB_ASSERT(nseq < p.seqStride && mpos >= anchor && msrc < (uint32_t)mpos + dP && mpos + mlen <= blkEnd);
            if (l == 0) { oLL[nseq] = (uint32_t)(mpos - anchor); oOF[nseq] = offcode + 1; oML[nseq] = (uint32_t)mlen - 3; }
            nseq++;
            int const mend = mpos + mlen;
            if (l == 0 && type != 3 && mend <= ilimit) {
                int const c2 = (int)(currPrev - dP) + 2;
                uint64_t const xa = rd64(src + c2), xb = rd64(src + mend - 2), xc = rd64(src + mend - 1);
                TL[hashL(xa, hBitsL)] = currPrev + 2;
                TL[hashL(xb, hBitsL)] = (uint32_t)(mend - 2) + dP;
                TS[hash_val(xa, hBitsS, mls)] = currPrev + 2;
                TS[hash_val(xc, hBitsS, mls)] = (uint32_t)(mend - 1) + dP;
            }
            ip = mend; anchor = mend;
            afterMatch = mend <= ilimit;
        } else if (active) {
            if (validMask != LOW) active = false;
            else ip = P;
        }
        __syncwarp();
    }
    if (mine && l == 0) { it.nbSeq = nseq; it.lastLL = (uint32_t)(blkEnd - anchor); it.repNext[0] = offset_1; it.repNext[1] = offset_2; }
}

// One thread per frame: block `wave` of a frame compressed with a loaded dictionary, for the blocks enc_match_dict_fast_group_kernel
// (which runs first and does the window bookkeeping of the block: dict_block_mode) leaves to it: ZSTD_dfast frames and blocks that no
// longer see the dictionary.
// `shift`: a frame gets 1 << shift consecutive threads, of which the first one works.  Few large frames: one frame per warp (shift 5: every
// parse has its own instruction stream, nothing waits for a divergent neighbour); many small records: one per thread (shift 0).
__global__ void __launch_bounds__(128) enc_match_dict_kernel(EncPass p, const uint32_t* __restrict__ workList, uint32_t nWork, uint32_t wave, uint32_t shift)
{
    uint32_t const gt = blockIdx.x * blockDim.x + threadIdx.x, wi = gt >> shift;
    if (wi >= nWork || (gt & ((1u << shift) - 1u)) != 0) return;
    uint32_t const item = workList[wi];
    EncItem& it = p.items[item];
    if (!(it.dBlkMode & kDictSerial)) return;
    uint32_t const mode = it.dBlkMode & 3u;
    const uint8_t* const src = p.src + it.srcOff;
    uint32_t const blkStart = wave * kBlockSizeMax, blkSize = min(kBlockSizeMax, it.srcSize - blkStart);
    uint32_t const srcIdx0 = it.dPrefix;                             // index of src[0]: 2 + dictionary content length (2 when nothing was attached)
    const uint8_t* const base = src - srcIdx0;
    uint32_t const maxDist = 1u << it.windowLog, dictLimit = it.wDictLimit, loadedDictEnd = it.loadedDictEnd;
    SeqWriter sw{p.seqLL + (size_t)item * p.seqStride, p.seqML + (size_t)item * p.seqStride, p.seqOF + (size_t)item * p.seqStride, 0, p.seqStride};
    uint32_t rep[2] = {it.rep[0], it.rep[1]};
    uint32_t* const T = p.tables + it.tableOff;
    uint32_t* const TS = T + ((size_t)1 << it.hashLog);              // ZSTD_dfast: the small table (chainTable) follows the long one
    bool const fast = it.strategy == 1;
    DictBlk b;
    b.base = base; b.dictBase = p.dict.content - 2; b.istart = src + blkStart; b.iend = src + blkStart + blkSize; b.mls = it.minMatch; b.dStep = it.dStep;
    b.dictStartIndex = it.dBlkLow; b.prefixStartIndex = it.dBlkPrefix;
    uint32_t lastLL;
    if (mode == 2) lastLL = fast ? dict_fast_ext(T, it.hashLog, b, rep, sw) : dict_dfast_ext(T, it.hashLog, TS, it.chainLog, b, rep, sw);
    else if (mode == 1)
        lastLL = fast ? dict_fast_dms(T, it.hashLog, p.dict.tables, p.dict.hashLog, b, rep, sw)
                      : dict_dfast_dms(T, it.hashLog, TS, it.chainLog, p.dict.tables, p.dict.hashLog, p.dict.tables + ((size_t)1 << p.dict.hashLog), p.dict.chainLog, b, rep, sw);
    else {
        auto lowest = [&](uint32_t lowestValid, uint32_t curr) { return loadedDictEnd != 0 ? lowestValid : ((curr - lowestValid > maxDist) ? curr - maxDist : lowestValid); };
        uint32_t const curr0 = srcIdx0 + blkStart + ((srcIdx0 + blkStart) == b.prefixStartIndex);
        uint32_t const maxRep = curr0 - lowest(dictLimit, curr0);
        lastLL = fast ? nodict_fast(T, it.hashLog, b, it.stepSize, maxRep, rep, sw) : nodict_dfast(T, it.hashLog, TS, it.chainLog, b, maxRep, rep, sw);
    }
    it.nbSeq = sw.n; it.lastLL = lastLL; it.repNext[0] = rep[0]; it.repNext[1] = rep[1];
}

// Start state of a frame that uses the CDict: prevCBlock's Huffman table is the dictionary's (both :2746 and :2803 copy cdict->cBlockState),
// and a frame that COPIES the CDict (ZSTD_resetCCtx_byCopyingCDict :2803) starts with the CDict's match-finder tables as its own.
__global__ void enc_dict_init_kernel(EncPass p, uint32_t nItems, uint32_t entries)
{
    uint32_t const item = blockIdx.x, part = blockIdx.y, parts = gridDim.y;      // frames along x: a pass of small records has more than 65535 of them
    if (item >= nItems) return;
    EncItem const& it = p.items[item];
    if (it.dMode == 0) return;
    if (part == 0 && it.hufRepeat) {
        uint32_t* const d = (uint32_t*)(p.hufState + (size_t)item * 2 * kHufStateSlot); const uint32_t* const s = (const uint32_t*)p.dict.huf;
        for (uint32_t k = threadIdx.x; k < kHufStateSlot / 4; k += blockDim.x) d[k] = s[k];
    }
    if (it.dMode != 2 || it.srcSize < 7) return;                     // attached frames search the CDict's tables in place
    uint4* const d = (uint4*)(p.tables + it.tableOff); const uint4* const s = (const uint4*)p.dict.tables;
    for (uint32_t k = part * blockDim.x + threadIdx.x; k < entries / 4; k += parts * blockDim.x) d[k] = s[k];
}

// ZSTD_initCDict_internal (:5826) on the device, one thread, once per (dictionary, level): ZSTD_loadCEntropy (:5264: HUF_readCTable
// HufCompress.cs:249, three FSE_buildCTable_wksp, ZSTD_dictNCountRepeat :5239) from the statistics dec_dict_kernel has already
// parsed and validated, then ZSTD_fillHashTable / ZSTD_fillDoubleHashTable with dtlm_full (ZstdFast.cs:9, ZstdDoubleFast.cs:9) over
// the content.  out[]: [0] 1 ok / 2 corrupted | [1] Huffman repeat mode (0 none, 1 check, 2 valid) | [2] FSE `valid` bits (1 LL, 2 OF, 4 ML).
__global__ void enc_dict_build_kernel(const uint32_t* __restrict__ stats, const uint32_t* __restrict__ info, const uint8_t* content, uint32_t contentLen,
                                      uint32_t strategy, uint32_t hashLog, uint32_t chainLog, uint32_t mls,
                                      uint32_t* tables, uint8_t* hufSlot, FseGTable* fse, uint32_t* out)
{
    __shared__ FseCTable ct; __shared__ uint16_t cumul[64]; __shared__ uint8_t tableSymbol[512]; __shared__ int16_t norm[64];
    if (threadIdx.x != 0) return;
    out[0] = 1; out[1] = 0; out[2] = 0;
    if (info[2]) {                                                   // a zstd-format dictionary: entropy tables
        uint32_t const tableLog = stats[256], nbSymbols = stats[257];
        if (nbSymbols < 256) { out[0] = 2; return; }                 // maxSymbolValue < 255: dictionary_corrupted (:5283)
        uint16_t* const val = (uint16_t*)(hufSlot + 256);
        uint32_t nbPerRank[14], valPerRank[14]; bool hasZero = false;
        for (int r = 0; r < 14; r++) nbPerRank[r] = valPerRank[r] = 0;
        for (uint32_t n = 0; n < 256; n++) { uint32_t const w = stats[n]; uint32_t const nb = w ? tableLog + 1 - w : 0; hufSlot[n] = (uint8_t)nb; nbPerRank[nb]++; hasZero |= (w == 0); }
        {   uint32_t mn = 0; for (uint32_t n = tableLog; n > 0; n--) { valPerRank[n] = mn; mn += nbPerRank[n]; mn >>= 1; } }
        for (uint32_t n = 0; n < 256; n++) val[n] = (uint16_t)(valPerRank[hufSlot[n]]++);
        out[1] = hasZero ? 1 : 2;
        uint32_t validBits = 0;
        for (int q = 0; q < 3; q++) {                                // stats order: OF, ML, LL; fse[] order: LL, OF, ML
            const uint32_t* const s = stats + 258 + q * 66;
            uint32_t const maxSV = s[64], tlog = s[65];
            for (int u = 0; u < 64; u++) norm[u] = (int16_t)(int32_t)s[u];
            uint32_t const full = q == 0 ? kMaxOff : (q == 1 ? kMaxML : kMaxLL);
            fse_build_ctable(ct, cumul, tableSymbol, norm, q == 0 ? kMaxOff : maxSV, tlog);      // the offset table is built over all 32 codes (:5297)
            FseGTable& g = fse[q == 0 ? 1 : (q == 1 ? 2 : 0)];
            for (uint32_t u = 0; u < 53; u++) g.tt[u] = ct.tt[u];
            g.tableLog = tlog; g._pad = 0;
            for (uint32_t u = 0; u < (1u << tlog); u++) g.stateTable[u] = ct.stateTable[u];
            uint32_t need = full;
            if (q == 0) { uint32_t const maxOffset = contentLen + 128 * 1024; uint32_t const offcodeMax = contentLen <= 0xFFFFFFFFu - 128 * 1024 ? highbit32(maxOffset) : kMaxOff; need = offcodeMax < kMaxOff ? offcodeMax : kMaxOff; }
            bool valid = maxSV >= need;                               // ZSTD_dictNCountRepeat
            for (uint32_t u = 0; valid && u <= need; u++) if (norm[u] == 0) valid = false;
            if (valid) validBits |= q == 0 ? 2u : (q == 1 ? 4u : 1u);
        }
        out[2] = validBits;
    }
    if (contentLen > 8) {
        const uint8_t* const base = content - 2; const uint8_t* ip = content; const uint8_t* const iend = content + contentLen - 8;
        if (strategy == 1) {
            for (; ip + 3 < iend + 2; ip += 3) {
                uint32_t const curr = (uint32_t)(ip - base);
                tables[hash_ptr(ip, hashLog, mls)] = curr;
                for (uint32_t q = 1; q < 3; ++q) { uint32_t const h = hash_ptr(ip + q, hashLog, mls); if (tables[h] == 0) tables[h] = curr + q; }
            }
        } else {
            uint32_t* const hashLarge = tables; uint32_t* const hashSmall = tables + ((size_t)1 << hashLog);
            for (; ip + 3 - 1 <= iend; ip += 3) {
                uint32_t const curr = (uint32_t)(ip - base);
                for (uint32_t i = 0; i < 3; ++i) {
                    uint64_t const x = rd64(ip + i);
                    uint32_t const smHash = hash_val(x, chainLog, mls), lgHash = hash_val(x, hashLog, 8);
                    if (i == 0) hashSmall[smHash] = curr + i;
                    if (i == 0 || hashLarge[lgHash] == 0) hashLarge[lgHash] = curr + i;
                }
            }
        }
    }
}

// MB = false: every frame is one block (srcSize <= 128 KiB), one CTA per frame, no state.  MB = true: the CTA handles block
// `wave` of frame workList[blockIdx.x]; the frame's write cursor, confirmed repcodes and previous Huffman table live in
// EncItem / hufState and are advanced here (ZSTD_compress_frameChunk :4690, ZSTD_compressBlock_internal :4528,
// ZSTD_blockState_confirmRepcodesAndEntropyTables).  Without a dictionary and below ZSTD_lazy the FSE tables never repeat
// (ZSTD_selectEncodingType returns set_repeat only for FSE_repeat_valid), so the Huffman table is the only entropy state.
// The stage runs as three launches: PHASE 0 (this kernel: literals section, sequence statistics, the three FSE tables and their
// descriptions), enc_fse_chain_kernel (one lane per FSE state chain), PHASE 1 (this kernel: bitstream scatter, block and frame
// assembly).  Inside one kernel the three chains of a chunk kept three single-lane warps busy for 12 K dependent steps each:
// 70 % of the stage's warp instructions ran with one active lane (profiles/r01_notes.md).
template <bool MB, int PHASE>
__global__ void __launch_bounds__(kEntThreads, MB ? 8 : 15) enc_entropy_kernel(EncPass p, const uint32_t* __restrict__ workList, uint32_t wave)
{
    __shared__ EntShared S;
    uint32_t const item = MB ? workList[blockIdx.x] : blockIdx.x, tid = threadIdx.x;
    EncItem& it = p.items[item];
    uint32_t const frameSize = it.srcSize;
    uint32_t const blkStart = MB ? wave * kBlockSizeMax : 0u;
    const uint8_t* const src = p.src + it.srcOff + blkStart;   // the block
    uint8_t* const dst = p.dst + it.dstOff;                    // 16-byte aligned slot of compressBound(frameSize) bytes
    uint32_t* const dstW = (uint32_t*)dst;
    uint32_t const srcSize = MB ? min(kBlockSizeMax, frameSize - blkStart) : frameSize;     // block size
    bool const firstBlock = !MB || wave == 0, lastBlock = !MB || blkStart + srcSize == frameSize;
    if (!MB && srcSize > kBlockSizeMax) { if (tid == 0) { p.results[item] = make_error(kSrcSizeWrong); p.carry[item].flags = 0; } return; }
    // ---- frame header: ZSTD_writeFrameHeader (ZstdCompress.cs:4817), contentSizeFlag = 1, the dictionary's id when one is loaded ----
    uint32_t fhSize = 0;
    if (firstBlock) {
        uint32_t const fcsCode = (frameSize >= 256) + (frameSize >= 65536 + 256);
        bool const singleSegment = !MB || ((1u << it.windowLog) >= frameSize);     // windowSize >= pledgedSrcSize (:4823); always for one-block frames
        uint32_t const dictID = (MB && it.dMode) ? p.dict.dictID : 0u;
        uint32_t const dictIDSizeCode = (dictID > 0) + (dictID >= 256) + (dictID >= 65536);
        fhSize = 4 + 1 + (singleSegment ? 0 : 1) + (dictIDSizeCode == 3 ? 4 : dictIDSizeCode) + (fcsCode == 0 ? (singleSegment ? 1 : 0) : (fcsCode == 1 ? 2 : 4));
        if (tid == 0 && PHASE == 0) {
            dst[0] = 0x28; dst[1] = 0xB5; dst[2] = 0x2F; dst[3] = 0xFD;
            dst[4] = (uint8_t)(dictIDSizeCode + (p.checksumFlag ? 4u : 0u) + ((singleSegment ? 1u : 0u) << 5) + (fcsCode << 6));   // FHD: dictID size bits 0-1, checksum bit 2, singleSegment bit 5, fcsID bits 6-7 (:4823-4829)
            uint32_t pos = 5;
            if (!singleSegment) dst[pos++] = (uint8_t)((it.windowLog - 10) << 3);
            if (MB) { uint32_t const nDid = dictIDSizeCode == 3 ? 4u : dictIDSizeCode; for (uint32_t q = 0; q < nDid; q++) dst[pos++] = (uint8_t)(dictID >> (8 * q)); }
            if (fcsCode == 0) { if (singleSegment) dst[pos++] = (uint8_t)frameSize; }
            else if (fcsCode == 1) { uint32_t const v = frameSize - 256; dst[pos] = (uint8_t)v; dst[pos + 1] = (uint8_t)(v >> 8); }
            else { dst[pos] = (uint8_t)frameSize; dst[pos + 1] = (uint8_t)(frameSize >> 8); dst[pos + 2] = (uint8_t)(frameSize >> 16); dst[pos + 3] = (uint8_t)(frameSize >> 24); }
        }
    }
    uint32_t const blkPos = firstBlock ? fhSize : it.outPos;    // block header goes here
    uint8_t* const blk = dst + blkPos;
    uint32_t const payload = blkPos + 3;                        // byte offset of the block content
    bool raw = false, rle = false, newHuf = false;
    uint32_t cSize = 0;
    EntCarry& cy = p.carry[item];
    FseGTable* const gt = p.fseTabs + (size_t)item * 3;
    if (frameSize == 0) {                                       // ZSTD_writeEpilogue: one empty last raw block (:5621-5631)
        if (tid == 0 && PHASE == 0) {
            cy.flags = 0;
            blk[0] = 1; blk[1] = 0; blk[2] = 0;
            uint32_t total = fhSize + 3;
            if (p.checksumFlag) { uint32_t const c = 0x51D8E999u; /* low 32 bits of XXH64("", seed 0) = 0xEF46DB3751D8E999 */ dst[total] = (uint8_t)c; dst[total + 1] = (uint8_t)(c >> 8); dst[total + 2] = (uint8_t)(c >> 16); dst[total + 3] = (uint8_t)(c >> 24); total += 4; }
            p.results[item] = total;
        }
        return;
    }
    uint32_t const nbSeq = it.nbSeq;
    if (srcSize < 7) { raw = true; if (PHASE == 0) { if (tid == 0) cy.flags = 0; return; } }
    if (!raw) {
        const uint32_t* const aLL = p.seqLL + (size_t)item * p.seqStride;
        const uint32_t* const aML = p.seqML + (size_t)item * p.seqStride;
        const uint32_t* const aOF = p.seqOF + (size_t)item * p.seqStride;
        uint64_t* const sb = p.stateBits + (size_t)item * p.seqStride;
        uint32_t op = 0, lastCountSize = 0;
      if (PHASE == 0) {
        uint8_t* const lit = p.litBuf + (size_t)item * p.litStride;
        // ---- 1. gather literals (ZSTD_storeSeq copies + ZSTD_storeLastLiterals) ----
        uint32_t litSize;
        {
            uint32_t srcPos = 0, litPos = 0;
            for (uint32_t base = 0; base < nbSeq; base += kEntThreads) {
                uint32_t const n = base + tid;
                uint32_t ll = 0, ml = 0;
                if (n < nbSeq) { ll = aLL[n]; ml = aML[n] + 3; }
                uint32_t totS, totL;
                uint32_t const sOff = ent_scan_excl(ll + ml, S.scanA, &totS);
                uint32_t const lOff = ent_scan_excl(ll, S.scanB, &totL);
                const uint8_t* s = src + srcPos + sOff; uint8_t* d = lit + litPos + lOff;
                for (uint32_t k = 0; k < ll; k++) d[k] = s[k];
                srcPos += totS; litPos += totL;
            }
            uint32_t const lastLL = it.lastLL;
            for (uint32_t k = tid; k < lastLL; k += kEntThreads) lit[litPos + k] = src[srcSize - lastLL + k];
            litSize = litPos + lastLL;
            ZB_ASSERT(litSize <= kBlockSizeMax);
        }
        for (uint32_t k = tid; k < 1024; k += kEntThreads) (&S.hist[0][0])[k] = 0;
        __syncthreads();
        // ---- 2. literals section: ZSTD_compressLiterals (ZstdCompressLiterals.cs:86) ----
        uint32_t const minGainLit = (litSize >> 6) + 2;
        uint32_t const lhSize = 3 + (litSize >= 1024) + (litSize >= 16384);
        // prevHuf->repeatMode == HUF_repeat_valid (a dictionary's table that codes every byte value): literals from 7 bytes up are
        // compressed, below 1 KiB as one stream (ZstdCompressLiterals.cs:103-120), and up to 1 KiB the old table is used without
        // looking at the literals at all (preferRepeat, HufCompress.cs:1404-1410)
        bool const hufValid = MB && it.hufRepeat == 2u;
        bool const direct = hufValid && litSize <= 1024;
        uint32_t const singleStream = litSize < 256 || (hufValid && lhSize == 3);
        uint32_t const seg = (litSize + 3) / 4;
        uint32_t litMode = 0;   // 0 raw, 1 rle, 2 huffman
        if (litSize > (hufValid ? 6u : 63u) && !it.rawLits) {       // rawLits: ZSTD_noCompressLiterals right away (ZstdCompressLiterals.cs:100-101)
            bool const suspect = (nbSeq == 0) || (litSize / nbSeq >= 20);       // ZstdCompress.cs:3262
            bool skip = false;
            if (!direct && suspect && litSize >= 40960) {                        // HufCompress.cs:1412-1446
                for (uint32_t k = tid; k < 4096; k += kEntThreads) { atomicAdd(&S.hist[0][lit[k]], 1u); atomicAdd(&S.hist[1][lit[litSize - 4096 + k]], 1u); }
                __syncthreads();
                uint32_t m0 = 0, m1 = 0;
                for (uint32_t q = tid; q < 256; q += kEntThreads) { m0 = max(m0, S.hist[0][q]); m1 = max(m1, S.hist[1][q]); }
                for (int d = 16; d; d >>= 1) { m0 = max(m0, __shfl_xor_sync(0xFFFFFFFFu, m0, d)); m1 = max(m1, __shfl_xor_sync(0xFFFFFFFFu, m1, d)); }
                if ((tid & 31) == 0) { S.scanA[tid >> 5] = m0; S.scanB[tid >> 5] = m1; }
                __syncthreads();
                uint32_t lb = 0, le = 0;
                for (int k = 0; k < kEntThreads / 32; k++) { lb = max(lb, S.scanA[k]); le = max(le, S.scanB[k]); }
                skip = (lb + le) <= ((2 * 4096) >> 7) + 4;
                __syncthreads();
                for (uint32_t k = tid; k < 1024; k += kEntThreads) (&S.hist[0][0])[k] = 0;
                __syncthreads();
            }
            if (!skip) {
                // per-segment histograms (HIST_count_wksp, Hist.cs:196; the 4 stream totals need per-segment counts)
                for (uint32_t k = tid; k < litSize; k += kEntThreads) atomicAdd(&S.hist[singleStream ? 0 : k / seg][lit[k]], 1u);
                __syncthreads();
                {
                    uint32_t m = 0, ms = 0;
                    for (uint32_t q = tid; q < 256; q += kEntThreads) {
                        uint32_t const c = S.hist[0][q] + S.hist[1][q] + S.hist[2][q] + S.hist[3][q];
                        S.count[q] = c;
                        m = max(m, c); if (c) ms = max(ms, q);
                    }
                    for (int d = 16; d; d >>= 1) { m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, d)); ms = max(ms, __shfl_xor_sync(0xFFFFFFFFu, ms, d)); }
                    if ((tid & 31) == 0) { S.scanA[tid >> 5] = m; S.scanB[tid >> 5] = ms; }
                }
                __syncthreads();
                uint32_t largest = 0, maxSym = 0;
                for (int k = 0; k < kEntThreads / 32; k++) { largest = max(largest, S.scanA[k]); maxSym = max(maxSym, S.scanB[k]); }
                __syncthreads();
                if (!direct && largest == litSize) litMode = 1;                   // all same byte -> rle (HufCompress.cs:1458)
                else if (!direct && largest <= (litSize >> 7) + 4) litMode = 0;   // not compressible enough (:1463)
                else {
                    // HUF_repeat (HufCompress.cs:1475-1530): the previous block's table may be reused (literals header type set_repeat, no
                    // table description).  repeatMode is `check` after any block that shipped a new table; `valid` needs a dictionary.
                    bool useOld = false;
                    uint32_t repeat = MB ? it.hufRepeat : 0u;
                    const uint8_t* const oldNb = p.hufState + ((size_t)item * 2 + (MB ? it.hufCur : 0u)) * kHufStateSlot;
                    if (MB && repeat) {
                        int bad = 0;                                               // HUF_validateCTable (only in `check` mode; a valid table codes every symbol)
                        for (uint32_t q = tid; q <= maxSym; q += kEntThreads) bad |= (S.count[q] != 0) & (oldNb[q] == 0);
                        if (__syncthreads_or(bad)) repeat = 0;
                        if (repeat && litSize <= 1024) useOld = true;              // preferRepeat (ZstdCompressLiterals.cs:121)
                    }
                    if (!useOld) {
                        if (tid == 0) {
                            uint32_t huffLog = fse_optimal_table_log(11, litSize, maxSym, 1);       // HUF_optimalTableLog :12
                            huffLog = huf_build_ctable(S, maxSym, huffLog);
                            S.hufLog = huffLog;
                            S.hSize = huffLog > 12 ? 0 : huf_write_ctable(S, maxSym, huffLog);
                        }
                        __syncthreads();
                        if (MB && repeat && S.hSize != 0) {                        // old table vs header + new table (:1511-1520)
                            uint32_t o = 0, nw = 0, totO, totN;
                            for (uint32_t q = tid; q <= maxSym; q += kEntThreads) { uint32_t const c = S.count[q]; o += oldNb[q] * c; nw += S.hufNbBits[q] * c; }
                            ent_scan_excl(o, S.scanA, &totO); ent_scan_excl(nw, S.scanB, &totN);
                            if ((totO >> 3) <= S.hSize + (totN >> 3) || S.hSize + 12 >= litSize) useOld = true;
                        }
                    }
                    if (MB && useOld) {
                        const uint16_t* const oldVal = (const uint16_t*)(oldNb + 256);
                        __syncthreads();
                        for (uint32_t q = tid; q < 256; q += kEntThreads) { S.hufNbBits[q] = oldNb[q]; S.hufValue[q] = oldVal[q]; }
                        __syncthreads();
                    }
                    uint32_t const hSize = useOld ? 0u : S.hSize;
                    if (!useOld && (hSize == 0 || hSize + 12 >= litSize)) litMode = 0;          // (:1525-1528)
                    else {
                        // stream sizes from the per-segment histograms
                        uint32_t const nStreams = singleStream ? 1 : 4;
                        if (tid < 4) S.streamBits[tid] = 0;
                        __syncthreads();
                        for (uint32_t k = 0; k < nStreams; k++) {
                            uint32_t v = 0;
                            for (uint32_t q = tid; q < 256; q += kEntThreads) v += S.hist[k][q] * S.hufNbBits[q];
                            for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
                            if ((tid & 31) == 0) atomicAdd(&S.streamBits[k], v);
                        }
                        __syncthreads();
                        uint32_t total = hSize + (singleStream ? 0 : 6); bool bad = false;
                        uint32_t sz[4] = {0, 0, 0, 0};
                        for (uint32_t k = 0; k < nStreams; k++) { sz[k] = (S.streamBits[k] >> 3) + 1; if (sz[k] > 65535 && !singleStream) bad = true; total += sz[k]; }
                        if (!singleStream && litSize < 12) bad = true;
                        if (bad || total >= litSize - 1) litMode = 0;              // HUF_compressCTable_internal :1333-1356
                        else if (total >= litSize - minGainLit) litMode = 0;       // ZstdCompressLiterals.cs:133
                        else if (total == 1) litMode = 1;                          // cLitSize == 1 -> ZSTD_compressRleLiteralsBlock (:139): one byte value, 7 times a 1-bit code
                        else {
                            litMode = 2;
                            // zero the stream area, then scatter every symbol's code (symbols are appended last-to-first)
                            uint32_t const strBase = payload + lhSize + hSize + (singleStream ? 0 : 6);
                            for (uint32_t k = tid; k < total - hSize - (singleStream ? 0 : 6); k += kEntThreads) dst[strBase + k] = 0;
                            __syncthreads();
                            uint32_t strOff = strBase;
                            for (uint32_t k = 0; k < nStreams; k++) {
                                uint32_t const a = k * seg, b = singleStream ? litSize : min(litSize, (k + 1) * seg);
                                uint64_t const bit0 = (uint64_t)strOff * 8;
                                uint32_t carry = 0;           // bits already placed (from the end of the segment)
                                // tiles from the end of the segment backwards; thread t takes 8 symbols
                                for (uint32_t hi = b; hi > a; ) {
                                    uint32_t const tile = min(hi - a, (uint32_t)kEntThreads * 8);
                                    uint32_t const lo = hi - tile;
                                    // thread t covers symbols [hi-8(t+1), hi-8t) clipped to >= lo, processed in descending order
                                    int64_t const myHi = (int64_t)hi - 8 * (int64_t)tid, myLo = max((int64_t)lo, myHi - 8);
                                    uint32_t bitsMine = 0;
                                    for (int64_t q = myHi - 1; q >= myLo; q--) bitsMine += S.hufNbBits[lit[q]];
                                    uint32_t tot;
                                    uint32_t const off = ent_scan_excl(bitsMine, S.scanA, &tot);
                                    uint64_t pos = bit0 + carry + off;
                                    for (int64_t q = myHi - 1; q >= myLo; q--) { uint32_t const sym = lit[q]; uint32_t const nb = S.hufNbBits[sym]; put_bits(dstW, pos, S.hufValue[sym], nb); pos += nb; }
                                    carry += tot; hi = lo;
                                }
                                if (tid == 0) put_bits(dstW, bit0 + carry, 1, 1);      // HUF_closeCStream end mark (:964)
                                strOff += sz[k];
                            }
                            __syncthreads();
                            if (MB && !useOld) {            // nextCBlock's table: becomes prevCBlock's if this block is confirmed
                                uint8_t* const nNb = p.hufState + ((size_t)item * 2 + (it.hufCur ^ 1u)) * kHufStateSlot; uint16_t* const nVal = (uint16_t*)(nNb + 256);
                                for (uint32_t q = tid; q < 256; q += kEntThreads) { nNb[q] = S.hufNbBits[q]; nVal[q] = S.hufValue[q]; }
                                newHuf = true;
                            }
                            if (tid == 0) {
                                uint8_t* o = dst + payload;
                                uint32_t const cLitSize = total;
                                uint32_t const hType = useOld ? 3u : 2u;           // set_repeat / set_compressed
                                if (lhSize == 3) { uint32_t const lhc = hType + ((!singleStream) << 2) + (litSize << 4) + (cLitSize << 14); o[0] = (uint8_t)lhc; o[1] = (uint8_t)(lhc >> 8); o[2] = (uint8_t)(lhc >> 16); }
                                else if (lhSize == 4) { uint32_t const lhc = hType + (2 << 2) + (litSize << 4) + (cLitSize << 18); o[0] = (uint8_t)lhc; o[1] = (uint8_t)(lhc >> 8); o[2] = (uint8_t)(lhc >> 16); o[3] = (uint8_t)(lhc >> 24); }
                                else { uint32_t const lhc = hType + (3 << 2) + (litSize << 4) + (cLitSize << 22); o[0] = (uint8_t)lhc; o[1] = (uint8_t)(lhc >> 8); o[2] = (uint8_t)(lhc >> 16); o[3] = (uint8_t)(lhc >> 24); o[4] = (uint8_t)(cLitSize >> 10); }
                                for (uint32_t k = 0; k < hSize; k++) o[lhSize + k] = S.hdr[k];
                                if (!singleStream) { uint8_t* j = o + lhSize + hSize; j[0] = (uint8_t)sz[0]; j[1] = (uint8_t)(sz[0] >> 8); j[2] = (uint8_t)sz[1]; j[3] = (uint8_t)(sz[1] >> 8); j[4] = (uint8_t)sz[2]; j[5] = (uint8_t)(sz[2] >> 8); }
                                S.litSectionSize = lhSize + cLitSize;
                            }
                        }
                    }
                }
            }
        }
        if (litMode == 0) {             // ZSTD_noCompressLiterals (:8)
            uint32_t const flSize = 1 + (litSize > 31) + (litSize > 4095);
            uint8_t* o = dst + payload;
            if (tid == 0) {
                if (flSize == 1) o[0] = (uint8_t)(0 + (litSize << 3));
                else if (flSize == 2) { uint32_t const v = 0 + (1 << 2) + (litSize << 4); o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); }
                else { uint32_t const v = 0 + (3 << 2) + (litSize << 4); o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16); }
                S.litSectionSize = flSize + litSize;
            }
            for (uint32_t k = tid; k < litSize; k += kEntThreads) o[flSize + k] = lit[k];
        } else if (litMode == 1) {      // ZSTD_compressRleLiteralsBlock (:49)
            if (tid == 0) {
                uint32_t const flSize = 1 + (litSize > 31) + (litSize > 4095);
                uint8_t* o = dst + payload;
                if (flSize == 1) o[0] = (uint8_t)(1 + (litSize << 3));
                else if (flSize == 2) { uint32_t const v = 1 + (1 << 2) + (litSize << 4); o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); }
                else { uint32_t const v = 1 + (3 << 2) + (litSize << 4); o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16); }
                o[flSize] = lit[0];
                S.litSectionSize = flSize + 1;
            }
        }
        __syncthreads();
        op = payload + S.litSectionSize;
        // ---- 3. sequences section (ZstdCompress.cs:3285-3352) ----
        if (tid == 0) {
            uint8_t* o = dst + op;
            if (nbSeq < 128) { o[0] = (uint8_t)nbSeq; S.seqHdrSize = 1; }
            else if (nbSeq < kLongNbSeq) { o[0] = (uint8_t)((nbSeq >> 8) + 0x80); o[1] = (uint8_t)nbSeq; S.seqHdrSize = 2; }
            else { o[0] = 0xFF; uint32_t const v = nbSeq - kLongNbSeq; o[1] = (uint8_t)v; o[2] = (uint8_t)(v >> 8); S.seqHdrSize = 3; }
        }
        __syncthreads();
        op += S.seqHdrSize;
        if (nbSeq > 0) {
            uint32_t const strategy = it.strategy;
            uint32_t const seqHead = op; op += 1;
            uint32_t types[3];
            // ZSTD_buildSequencesStatistics (:3127): one pass writes the three symbol codes of every sequence (into the slots
            // the state chains will overwrite) and counts them; then three threads build the LL / OF / ML tables
            // concurrently, each with private scratch; thread 0 finally lays the table descriptions out in order.
            for (uint32_t q = tid; q < 192; q += kEntThreads) (&S.count3[0][0])[q] = 0;
            __syncthreads();
            for (uint32_t n = tid; n < nbSeq; n += kEntThreads) {
                uint32_t const llc = ll_code(aLL[n]), ofc = highbit32(aOF[n]), mlc = ml_code(aML[n]);
                sb[n] = (uint64_t)llc | ((uint64_t)ofc << 16) | ((uint64_t)mlc << 32);      // u16 slots: [0] LL [1] OF [2] ML
                atomicAdd(&S.count3[0][llc], 1u); atomicAdd(&S.count3[1][ofc], 1u); atomicAdd(&S.count3[2][mlc], 1u);
            }
            __syncthreads();
            if (tid < 96 && (tid & 31) == 0) {
                int const k = tid >> 5;
                uint32_t* const count = S.count3[k]; int16_t* const norm = S.norm3[k];
                const uint16_t* const codes = (const uint16_t*)sb + k;                   // 4 x u16 per sequence
                uint32_t const maxCode = k == 0 ? kMaxLL : (k == 1 ? kMaxOff : kMaxML);
                uint32_t max = maxCode; while (!count[max]) max--;
                uint32_t mostFrequent = 0; for (uint32_t q = 0; q <= max; q++) if (count[q] > mostFrequent) mostFrequent = count[q];
                uint32_t const defLog = k == 1 ? kOFDefaultNormLog : kLLDefaultNormLog;
                bool const defAllowed = k == 1 ? (max <= (uint32_t)kDefaultMaxOff) : true;
                bool const valid = MB && ((it.fseValid >> (k == 0 ? 0 : (k == 1 ? 1 : 2))) & 1u);      // bits: 1 LL, 2 OF, 4 ML
                uint32_t const type = select_encoding_type(mostFrequent, nbSeq, defLog, defAllowed, strategy, valid);
                uint32_t const lastCode = codes[(size_t)(nbSeq - 1) * 4];
                uint32_t countSize = 0;
                FseCTable& ct = S.ct[k];
                if (type == 1) {            // set_rle: FSE_buildCTable_rle (FseCompress.cs:706)
                    ct.tableLog = 0; ct.stateTable[0] = 0; ct.stateTable[1] = 0; ct.tt[max].deltaNbBits = 0; ct.tt[max].deltaFindState = 0;
                    S.hdr3[k][0] = (uint8_t)codes[0];   // the reference writes codeTable[0], the code of the FIRST sequence (all codes are equal here)
                    countSize = 1;
                } else if (type == 3) {     // set_repeat: the dictionary's table, no description (ZstdCompressSequences.cs:494)
                    const FseGTable& dg = p.dict.fse[k];
                    ct.tableLog = dg.tableLog;
                    for (uint32_t q = 0; q < 53; q++) ct.tt[q] = dg.tt[q];
                    for (uint32_t q = 0; q < (1u << dg.tableLog); q++) ct.stateTable[q] = dg.stateTable[q];
                } else if (type == 0) {     // set_basic
                    uint32_t const dmax = k == 0 ? kMaxLL : (k == 1 ? kDefaultMaxOff : kMaxML);
                    for (uint32_t q = 0; q <= dmax; q++) norm[q] = k == 0 ? c_LL_defaultNorm[q] : (k == 1 ? c_OF_defaultNorm[q] : c_ML_defaultNorm[q]);
                    fse_build_ctable(ct, S.cumul3[k], S.tableSymbol3[k], norm, dmax, defLog);
                } else {                    // set_compressed: ZSTD_buildCTable (ZstdCompressSequences.cs:471)
                    uint32_t const FSELog = k == 1 ? kOffFSELog : kLLFSELog;
                    uint32_t nbSeq_1 = nbSeq;
                    uint32_t const tableLog = fse_optimal_table_log(FSELog, nbSeq, max, 2);
                    if (count[lastCode] > 1) { count[lastCode]--; nbSeq_1--; }
                    fse_normalize_count(norm, tableLog, count, nbSeq_1, max, nbSeq_1 >= 2048);
                    countSize = fse_write_ncount(S.hdr3[k], norm, max, tableLog);
                    fse_build_ctable(ct, S.cumul3[k], S.tableSymbol3[k], norm, max, tableLog);
                }
                S.res3[k][0] = countSize; S.res3[k][1] = type;
            }
            __syncthreads();
            for (int k = 0; k < 3; k++) {
                uint32_t const cs = S.res3[k][0];
                types[k] = S.res3[k][1];
                for (uint32_t q = tid; q < cs; q += kEntThreads) dst[op + q] = S.hdr3[k][q];
                op += cs;
                if (types[k] == 2) lastCountSize = cs;
            }
            __syncthreads();
            if (tid == 0) dst[seqHead] = (uint8_t)((types[0] << 6) + (types[1] << 4) + (types[2] << 2));
            if (MB && tid == 0) it.fseValidNext = (types[0] == 3 ? 1u : 0u) | (types[1] == 3 ? 2u : 0u) | (types[2] == 3 ? 4u : 0u);    // any other type leaves `none` or `check`
            // ---- hand the three tables to the state-chain kernel ----
            for (int k = 0; k < 3; k++) {
                const FseCTable& ct = S.ct[k]; FseGTable& g = gt[k];
                uint32_t const tlog = ct.tableLog;
                const uint32_t* const ttw = (const uint32_t*)ct.tt; uint32_t* const gtw = (uint32_t*)g.tt;
                for (uint32_t q = tid; q < 106; q += kEntThreads) gtw[q] = ttw[q];
                for (uint32_t q = tid; q < max(2u, 1u << tlog); q += kEntThreads) g.stateTable[q] = ct.stateTable[q];
                if (tid == 0) g.tableLog = tlog;
            }
        }
        if (tid == 0) { cy.op = op; cy.lastCountSize = lastCountSize; cy.flags = (nbSeq > 0 ? 1u : 0u) | (newHuf ? 2u : 0u); }
        if (MB && tid == 0 && nbSeq == 0) it.fseValidNext = it.fseValid;        // no sequences: the FSE state is carried over as it is (:3283)
        return;
      }
        // ================= PHASE 1: after the state chains =================
        op = cy.op; lastCountSize = cy.lastCountSize; newHuf = (cy.flags & 2u) != 0;
        if (nbSeq > 0) {
            uint32_t const tl0 = gt[0].tableLog, tl1 = gt[1].tableLog, tl2 = gt[2].tableLog;
            // ---- 5. bitstream scatter ----
            // order inside sequence n's packet (low -> high): [OF state bits][ML state bits][LL state bits] (none for n = nbSeq-1),
            // then LL extra, ML extra, OF extra.  Packets are laid out n = nbSeq-1 first.
            uint32_t totalBits = 0;
            {
                // first pass: total length, so that the area can be zeroed before scattering
                uint32_t mine = 0;
                for (uint32_t n = tid; n < nbSeq; n += kEntThreads) {
                    uint32_t const llc = ll_code(aLL[n]), mlc = ml_code(aML[n]), ofc = highbit32(aOF[n]);
                    const uint16_t* s16 = (const uint16_t*)(sb + n);
                    mine += c_LL_bits[llc] + c_ML_bits[mlc] + ofc + (s16[0] >> 12) + (s16[1] >> 12) + (s16[2] >> 12);
                }
                uint32_t tot; ent_scan_excl(mine, S.scanA, &tot);
                totalBits = tot + tl0 + tl1 + tl2;
            }
            uint32_t const streamSize = (totalBits + 1 + 7) / 8;             // BIT_closeCStream: 1-bit end mark
            for (uint32_t k = tid; k < streamSize; k += kEntThreads) dst[op + k] = 0;
            __syncthreads();
            {
                uint64_t const bit0 = (uint64_t)op * 8;
                uint32_t carry = 0;
                for (uint32_t hi = nbSeq; hi > 0; ) {
                    uint32_t const tile = min(hi, (uint32_t)kEntThreads);
                    uint32_t const lo = hi - tile;
                    bool const active = tid < tile;
                    uint32_t const n = hi - 1 - tid;                           // descending
                    uint32_t llv = 0, mlv = 0, ofv = 0, llb = 0, mlb = 0, ofb = 0, sL = 0, sM = 0, sO = 0;
                    if (active) {
                        llv = aLL[n]; mlv = aML[n]; ofv = aOF[n];
                        llb = c_LL_bits[ll_code(llv)]; mlb = c_ML_bits[ml_code(mlv)]; ofb = highbit32(ofv);
                        const uint16_t* s16 = (const uint16_t*)(sb + n);
                        sL = s16[0]; sO = s16[1]; sM = s16[2];
                    }
                    uint32_t const len = active ? (llb + mlb + ofb + (sL >> 12) + (sM >> 12) + (sO >> 12)) : 0;
                    uint32_t tot;
                    uint32_t const off = ent_scan_excl(len, S.scanA, &tot);
                    if (active) {
                        uint64_t pos = bit0 + carry + off;
                        put_bits(dstW, pos, sO & 0xFFF, sO >> 12); pos += sO >> 12;
                        put_bits(dstW, pos, sM & 0xFFF, sM >> 12); pos += sM >> 12;
                        put_bits(dstW, pos, sL & 0xFFF, sL >> 12); pos += sL >> 12;
                        put_bits(dstW, pos, llv, llb); pos += llb;
                        put_bits(dstW, pos, mlv, mlb); pos += mlb;
                        put_bits(dstW, pos, ofv, ofb);
                    }
                    carry += tot; hi = lo;
                }
                if (tid == 0) {     // FSE_flushCState x3 in the order ML, OF, LL, then the end mark (:694-700)
                    uint64_t pos = bit0 + carry;
                    put_bits(dstW, pos, cy.finalState[2], tl2); pos += tl2;
                    put_bits(dstW, pos, cy.finalState[1], tl1); pos += tl1;
                    put_bits(dstW, pos, cy.finalState[0], tl0); pos += tl0;
                    put_bits(dstW, pos, 1, 1);
                }
            }
            op += streamSize;
            if (lastCountSize && (lastCountSize + streamSize) < 4) raw = true;          // ZstdCompress.cs:3346-3350
        }
        __syncthreads();
        cSize = op - payload;
        // ZSTD_entropyCompressSeqStore gate (:3381-3389) and capacity of the slot
        if (cSize >= srcSize - ((srcSize >> 6) + 2)) raw = true;
    }
    __syncthreads();
    // ZSTD_compressBlock_internal :4563-4568: after the first block, a block of one repeated byte whose compressed form is tiny becomes an RLE block
    if (MB && !firstBlock && srcSize >= 7 && (raw ? 0u : cSize) < 25) {
        int diff = 0; uint8_t const b0 = src[0];
        for (uint32_t k = tid; k < srcSize; k += kEntThreads) diff |= src[k] != b0;
        if (!__syncthreads_or(diff)) { rle = true; raw = false; }
    }
    uint32_t const lastBit = lastBlock ? 1u : 0u;
    uint32_t total;
    if (raw) {          // ZSTD_noCompressBlock (ZstdCompressInternal.cs:102)
        for (uint32_t k = tid; k < srcSize; k += kEntThreads) dst[payload + k] = src[k];
        if (tid == 0) {
            uint32_t const h = lastBit + (0u << 1) + (srcSize << 3);
            blk[0] = (uint8_t)h; blk[1] = (uint8_t)(h >> 8); blk[2] = (uint8_t)(h >> 16);
        }
        total = payload + srcSize;
    } else if (rle) {   // bt_rle: one byte of content, the header carries the regenerated size (:4750)
        if (tid == 0) {
            uint32_t const h = lastBit + (1u << 1) + (srcSize << 3);
            blk[0] = (uint8_t)h; blk[1] = (uint8_t)(h >> 8); blk[2] = (uint8_t)(h >> 16);
            dst[payload] = src[0];
        }
        total = payload + 1;
    } else {
        if (tid == 0) {
            uint32_t const h = lastBit + (2u << 1) + (cSize << 3);
            blk[0] = (uint8_t)h; blk[1] = (uint8_t)(h >> 8); blk[2] = (uint8_t)(h >> 16);
        }
        total = payload + cSize;
    }
    if (MB && tid == 0) {
        // ZSTD_blockState_confirmRepcodesAndEntropyTables (:4575): only a compressed block (cSize > 1) moves the repcodes and the Huffman table on
        if (!raw && !rle) {
            it.rep[0] = it.repNext[0]; it.rep[1] = it.repNext[1];
            if (newHuf) { it.hufCur ^= 1u; it.hufRepeat = 1u; }
            it.fseValid = it.fseValidNext;
        }
        it.fseValid &= ~2u;             // the dictionary's offset table is only trusted for the first block (:4582-4589)
        it.outPos = total;
    }
    if (!lastBlock) return;
    // ZSTD_writeEpilogue (:5641-5652): optional content checksum over the whole frame, hashed by warp 0
    if (p.checksumFlag && tid < 32) {
        uint32_t const c = (uint32_t)xxh64_warp(p.src + it.srcOff, frameSize, tid);
        if (tid == 0) { dst[total] = (uint8_t)c; dst[total + 1] = (uint8_t)(c >> 8); dst[total + 2] = (uint8_t)(c >> 16); dst[total + 3] = (uint8_t)(c >> 24); }
    }
    if (tid == 0) p.results[item] = total + (p.checksumFlag ? 4u : 0u);
}

// ------------------------------------------------------------------------------------------------------------
//  FSE state chains (ZSTD_encodeSequences_body, ZstdCompressSequences.cs:585; FSE_encodeSymbol, Fse.cs:60): one LANE per chain.
//  A block has three chains (LL, OF, ML), each a walk over its sequences from the last to the first in which the state after
//  sequence n feeds sequence n-1: state -> nbBits -> next state.  8192 blocks give 24576 independent chains; 32 of them share
//  a warp (same instruction stream, every lane busy), their state tables (<= 1 KB each) sit in shared memory, the symbol
//  transforms (independent of the state) are fetched eight steps ahead through L1/L2.  Per sequence the chain leaves
//  value | nbBits << 12 in the slot that held the symbol code; the final state goes to EntCarry.
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t kChainTabStride = 1028;     // bytes per chain in shared memory: 512 x u16 + 4 (bank skew)
template <bool MB>
__global__ void __launch_bounds__(32) enc_fse_chain_kernel(EncPass p, const uint32_t* __restrict__ workList, uint32_t nItems)
{
    extern __shared__ __align__(16) uint8_t s_chain[];
    uint32_t const lane = threadIdx.x;
    uint32_t const base = blockIdx.x * 32;
    // cooperative, coalesced staging of the 32 state tables of this warp
    for (uint32_t t = 0; t < 32; t++) {
        uint32_t const ch = base + t, idx = ch / 3, k = ch % 3;
        if (idx >= nItems) break;
        uint32_t const item = MB ? workList[idx] : idx;
        if (!(p.carry[item].flags & 1u)) continue;
        const FseGTable& g = p.fseTabs[(size_t)item * 3 + k];
        uint32_t const words = max(1u, (1u << g.tableLog) >> 1);
        const uint32_t* const src = (const uint32_t*)g.stateTable; uint32_t* const dst = (uint32_t*)(s_chain + t * kChainTabStride);
        for (uint32_t w = lane; w < words; w += 32) dst[w] = src[w];
    }
    __syncwarp();
    uint32_t const ch = base + lane, idx = ch / 3, k = ch % 3;      // k: 0 LL, 1 OF, 2 ML (the u16 slot of the chain inside a sequence's 8 bytes)
    if (idx >= nItems) return;
    uint32_t const item = MB ? workList[idx] : idx;
    EntCarry& cy = p.carry[item];
    if (!(cy.flags & 1u)) return;
    const FseGTable& g = p.fseTabs[(size_t)item * 3 + k];
    const uint2* const gtt = (const uint2*)g.tt;                    // {deltaFindState, deltaNbBits}
    const uint16_t* const st = (const uint16_t*)(s_chain + lane * kChainTabStride);
    uint32_t const nbSeq = p.items[item].nbSeq;
    uint16_t* const out16 = (uint16_t*)(p.stateBits + (size_t)item * p.seqStride) + k;     // 4 x u16 per sequence
    uint32_t state;
    {   // FSE_initCState2 (Fse.cs:38) on the last sequence's symbol
        uint2 const tt = __ldg(gtt + out16[(size_t)(nbSeq - 1) * 4]);
        uint32_t const nbBitsOut = (tt.y + (1u << 15)) >> 16;
        uint32_t const value = (nbBitsOut << 16) - tt.y;
        state = st[(int32_t)(value >> nbBitsOut) + (int32_t)tt.x];
        out16[(size_t)(nbSeq - 1) * 4] = 0;
    }
    uint32_t n = nbSeq - 1;                                         // sequences n-1 .. 0 are still to do
    while (n >= 8) {
        uint32_t c[8]; uint2 t[8];
#pragma unroll
        for (int j = 0; j < 8; j++) c[j] = out16[(size_t)(n - 1 - j) * 4];
#pragma unroll
        for (int j = 0; j < 8; j++) t[j] = __ldg(gtt + c[j]);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t const nb = (state + t[j].y) >> 16;
            out16[(size_t)(n - 1 - j) * 4] = (uint16_t)((state & ((1u << nb) - 1)) | (nb << 12));
            state = st[(int32_t)(state >> nb) + (int32_t)t[j].x];
        }
        n -= 8;
    }
    while (n-- > 0) {
        uint2 const tt = __ldg(gtt + out16[(size_t)n * 4]);
        uint32_t const nb = (state + tt.y) >> 16;
        out16[(size_t)n * 4] = (uint16_t)((state & ((1u << nb) - 1)) | (nb << 12));
        state = st[(int32_t)(state >> nb) + (int32_t)tt.x];
    }
    cy.finalState[k] = state;       // flushed after the last packet
}

__global__ void enc_compact_kernel(const uint8_t* src, const uint64_t* srcOff, const uint64_t* sizes, const uint64_t* dstOff, uint8_t* dst)
{
    uint32_t const i = blockIdx.x;
    uint64_t const n = sizes[i];
    const uint8_t* s = src + srcOff[i]; uint8_t* d = dst + dstOff[i];
    for (uint64_t k = threadIdx.x; k < n; k += blockDim.x) d[k] = s[k];
}

// ------------------------------------------------------------------------------------------------------------
//  Host driver
// ------------------------------------------------------------------------------------------------------------
struct DBuf { void* p = nullptr; size_t cap = 0;
    bool ensure(size_t n) { if (n <= cap) return true; if (p) cudaFree(p); p = nullptr; cap = 0; size_t w = n + n / 8 + 256; if (cudaMalloc(&p, w) != cudaSuccess) { w = n + 256; if (cudaMalloc(&p, w) != cudaSuccess) { p = nullptr; return false; } } cap = w; return true; }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; } };
struct HBuf { void* p = nullptr; size_t cap = 0;
    bool ensure(size_t n) { if (n <= cap) return true; if (p) cudaFreeHost(p); p = nullptr; cap = 0; size_t w = n + n / 8 + 256; if (cudaHostAlloc(&p, w, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; return false; } cap = w; return true; }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; } };

struct EncArenaImpl {
    DBuf items, tables, seqLL, seqML, seqOF, lit, stateBits, results, compact, cSrcOff, cSizes, cDstOff, workLists, hufState, fseTabs, carry;
    HBuf hItems, hResults, hC, hWork;
    cudaEvent_t copied = nullptr;
    cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};       // enc_compress_device: before the match finder / after it / after the entropy stage of this arena's pass
    EncArenaImpl() { cudaEventCreateWithFlags(&copied, cudaEventDisableTiming); for (auto& e : tev) cudaEventCreate(&e); }
    ~EncArenaImpl() { if (copied) cudaEventDestroy(copied); for (auto& e : tev) if (e) cudaEventDestroy(e); }
};
static thread_local std::string t_encErr;
const char* enc_last_error() { return t_encErr.c_str(); }

// ---- digested dictionary (ZSTD_CDict_s): built on the device once per (dictionary, level) ----
struct EncDictImpl {
    DBuf raw, decHuf, decFse, info, stats, tables, huf, fse, out;
    uint32_t dictSize = 0, contentOff = 0, contentLen = 0, dictID = 0, hufRepeat = 0, fseValid = 0, rep[2] = {1, 4};
    CParams cp{};               // the CDict's parameters (ZSTD_cpm_createCDict: unknown source size)
    size_t tableEntries = 0;
};
void EncDict::release()
{
    if (!impl) return;
    DBuf* d[] = {&impl->raw, &impl->decHuf, &impl->decFse, &impl->info, &impl->stats, &impl->tables, &impl->huf, &impl->fse, &impl->out};
    for (auto* b : d) b->release();
    delete impl; impl = nullptr; level = 0; ready = false;
}
// ZSTD_initLocalDict (:1581) -> ZSTD_createCDict_advanced2 (:5933) -> ZSTD_initCDict_internal (:5826).  Returns 0 or a zstd error code.
size_t enc_dict_digest(EncDict& D, cudaStream_t stream, const void* dict, size_t dictSize, int level)
{
    D.release();
    if (dictSize > (1u << 30)) return (size_t)make_error(kMemoryAllocation);     // indices are position + 2 + content length in 32 bits; the reference trims such dictionaries (:5133-5147), this library refuses them
    CParams const cp = get_cparams_dict(level, kSrcSizeUnknown, dictSize, 2);
    if (cp.strategy == 0) return (size_t)make_error(kParameterUnsupported);       // a level outside ZSTD_fast / ZSTD_dfast for this dictionary size
    D.impl = new EncDictImpl();
    EncDictImpl& I = *D.impl;
    I.cp = cp; I.dictSize = (uint32_t)dictSize;
    I.tableEntries = ((size_t)1 << cp.hashLog) + (cp.strategy == 2 ? ((size_t)1 << cp.chainLog) : 0);
    auto fail = [&](ErrorCode e) { D.release(); return (size_t)make_error(e); };
    if (!I.raw.ensure(dictSize + 64) || !I.decHuf.ensure(kHufTableEntries * 2) || !I.decFse.ensure(kFseTableEntries * 4) || !I.info.ensure(kDictInfoWords * 4) ||
        !I.stats.ensure(kEncDictStatsWords * 4) || !I.tables.ensure(I.tableEntries * 4 + 16) || !I.huf.ensure(kHufStateSlot) || !I.fse.ensure(3 * sizeof(FseGTable)) || !I.out.ensure(16))
        return fail(kMemoryAllocation);
    uint32_t info[kDictInfoWords] = {}, out[4] = {};
    if (cudaMemsetAsync(I.raw.p, 0, dictSize + 64, stream) != cudaSuccess || cudaMemcpyAsync(I.raw.p, dict, dictSize, cudaMemcpyHostToDevice, stream) != cudaSuccess) return fail(kGeneric);
    cudaMemsetAsync(I.tables.p, 0, I.tableEntries * 4 + 16, stream); cudaMemsetAsync(I.huf.p, 0, kHufStateSlot, stream); cudaMemsetAsync(I.fse.p, 0, 3 * sizeof(FseGTable), stream);
    cudaMemsetAsync(I.stats.p, 0, kEncDictStatsWords * 4, stream);
    dec_launch_dict_setup((const uint8_t*)I.raw.p, (uint32_t)dictSize, (uint16_t*)I.decHuf.p, (uint32_t*)I.decFse.p, (uint32_t*)I.info.p, stream, (uint32_t*)I.stats.p);
    if (cudaMemcpyAsync(info, I.info.p, sizeof(info), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return fail(kGeneric);
    // a dictionary the reference cannot digest: ZSTD_createCDict_advanced2 returns NULL and ZSTD_initLocalDict reports memory_allocation (:1604-1607)
    if (info[0] != 1) return fail(kMemoryAllocation);
    I.contentOff = dictSize < 8 ? 0 : info[10];
    I.contentLen = dictSize < 8 ? 0 : info[11];                  // ZSTD_compress_insertDictionary (:5467): below 8 bytes nothing is loaded
    I.dictID = info[2] ? info[1] : 0;
    if (info[2]) { I.rep[0] = info[7]; I.rep[1] = info[8]; }
    enc_dict_build_kernel<<<1, 32, 0, stream>>>((const uint32_t*)I.stats.p, (const uint32_t*)I.info.p, (const uint8_t*)I.raw.p + I.contentOff, I.contentLen,
                                                 cp.strategy, cp.hashLog, cp.chainLog, cp.minMatch, (uint32_t*)I.tables.p, (uint8_t*)I.huf.p, (FseGTable*)I.fse.p, (uint32_t*)I.out.p);
    if (cudaMemcpyAsync(out, I.out.p, sizeof(out), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail(kGeneric);
    if (out[0] != 1) return fail(kMemoryAllocation);
    I.hufRepeat = out[1]; I.fseValid = out[2];
    D.level = level; D.ready = true;
    return 0;
}
size_t enc_max_frame_bytes() { return kEncMaxFrameBytes; }
void EncArena::release()
{
    if (!impl) return;
    DBuf* d[] = {&impl->items, &impl->tables, &impl->seqLL, &impl->seqML, &impl->seqOF, &impl->lit, &impl->stateBits, &impl->results, &impl->compact, &impl->cSrcOff, &impl->cSizes, &impl->cDstOff, &impl->workLists, &impl->hufState, &impl->fseTabs, &impl->carry};
    for (auto* b : d) b->release();
    impl->hItems.release(); impl->hResults.release(); impl->hC.release(); impl->hWork.release();
    delete impl; impl = nullptr;
}
const uint8_t* EncArena::compactBuf() const { return impl ? (const uint8_t*)impl->compact.p : nullptr; }

#define ENC_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { t_encErr = std::string(#call) + ": " + cudaGetErrorString(e_); fprintf(stderr, "[zstdb200] CUDA failure: %s\n", t_encErr.c_str()); return false; } } while (0)

constexpr size_t kEncMaxItemsPerPass = 8192;
// Frames one pass can hold: 8192 blocks of 128 KiB fill the arenas (0.9 MB of sequence / literal / state scratch per block); passes of smaller
// records hold proportionally more, up to 65536 (one launch overhead and one host round trip per pass is what small records are bound by)
size_t enc_pass_capacity(size_t maxBlockBytes)
{
    size_t const blk = std::max<size_t>(std::min<size_t>(maxBlockBytes, kBlockSizeMax), 16384);
    return std::min<size_t>(65536, kEncMaxItemsPerPass * (kBlockSizeMax / blk));
}

// Queues one pass (m <= kEncMaxItemsPerPass frames) without waiting: the frame descriptors and work lists are copied on
// `copyStream` (pass the stream that carries the bulk H2D of the same frames, so that the small copies do not queue
// behind later bulk transfers; `stream` must then wait for that stream's event before this call), all kernels
// run on `stream`, and the per-frame results are written straight into pinned host memory (enc_results).
// Frames of at most 128 KiB (the batch shape of BASELINE.json) take one match launch and one entropy launch.  If the pass
// holds larger frames, it runs in waves: wave b = block b of every frame that has one, match finder then entropy stage,
// because block b+1 of a frame needs the hash table, the confirmed repcodes and the Huffman table that block b leaves behind
// (ZSTD_compress_frameChunk, ZstdCompress.cs:4690).  A frame is a serial chain of ~25 ms per block; the batch is what is parallel.
// ev3 (optional): [0] before the match finder, [1] after it (first wave), [2] after the last entropy stage.
bool enc_enqueue(EncArena& A, cudaStream_t stream, cudaStream_t copyStream, size_t m, int level, int checksumFlag,
                 const uint8_t* d_src, const uint64_t* srcOff, const size_t* srcSize,
                 uint8_t* d_dst, const uint64_t* dstOff, const size_t* dstCap, cudaEvent_t* ev3, unsigned* launches, const EncDict* dict)
{
    const EncDictImpl* const DI = (dict && dict->ready) ? dict->impl : nullptr;
    static bool const trace = getenv("ZSTDB200_TRACE_ENQ") != nullptr;       // developer aid: host time of the stages of this function
    auto const tq0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (trace) fprintf(stderr, "[zstdb200] enqueue %-10s %.3f ms since entry\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tq0).count()); };
    if (!A.impl) A.impl = new EncArenaImpl();
    EncArenaImpl& I = *A.impl;
    size_t maxBlk = 0;
    for (size_t i = 0; i < m; i++) maxBlk = std::max(maxBlk, std::min<size_t>(srcSize[i], kBlockSizeMax));
    if (m > enc_pass_capacity(maxBlk)) { t_encErr = "pass too large"; return false; }
    // per-frame strides of the sequence / literal arenas follow the largest block of the pass: a pass of small records holds more of them
    uint32_t const seqStride = (uint32_t)((maxBlk / 4 + 1 + 3) & ~(size_t)3), litStride = (uint32_t)((maxBlk + 64 + 15) & ~(size_t)15);
    if (!I.hItems.ensure(m * sizeof(EncItem)) || !I.items.ensure(m * sizeof(EncItem)) || !I.hResults.ensure(m * 8)) { t_encErr = "out of memory (items)"; return false; }
    EncItem* hi = (EncItem*)I.hItems.p;
    uint64_t* const hr = (uint64_t*)I.hResults.p;
    size_t tableEntries = 0, maxWaves = 1, totalBlocks = 0;
    bool mb = false;
    std::vector<uint32_t> nBlk(m);
    // Everything but the offsets depends on the size alone: records of equal size (the usual batch) share one prepared descriptor
    EncItem proto; size_t protoSS = ~(size_t)0, protoTable = 0; uint32_t protoBlk = 0; uint64_t protoErr = 0;
    for (size_t i = 0; i < m; i++) {
        size_t const ss = srcSize[i];
        if (ss != protoSS) {
            EncItem& e = proto;
            memset(&e, 0, sizeof(e));
            protoSS = ss; protoTable = 0; protoBlk = 0; protoErr = 0;
            e.srcSize = (uint32_t)std::min<size_t>(ss, 0xFFFFFFF0u);
            CParams c = get_cparams(level, ss);
            if (DI) {
                // ZSTD_CCtx_init_compressStream2 (:6949) + ZSTD_resetCCtx_usingCDict (:2881): the frame's window comes from the level, the
                // source and the dictionary size; everything else from the CDict -- as it is when the CDict is copied, re-adjusted for
                // the source size when it is attached (small inputs: ZSTD_shouldAttachDict :2738, 8 KiB for ZSTD_fast, 16 KiB for ZSTD_dfast)
                bool const attach = ss <= (DI->cp.strategy == 1 ? 8u * 1024 : 16u * 1024);
                CParams const req = get_cparams_dict(level, ss, DI->dictSize, attach ? 1 : 0);
                c = attach ? adjust_cparams_dict(DI->cp, ss, DI->dictSize, 1) : DI->cp;
                if (req.strategy == 0) c.strategy = 0; else c.windowLog = req.windowLog;
                uint32_t const L = DI->contentLen, cdictEnd = 2 + L;
                e.dMode = attach ? 1u : 2u; e.dStep = c.targetLength + !c.targetLength;
                if (attach) {
                    if (L == 0) { e.dPrefix = 2; e.wLow = e.wDictLimit = 2; }                        // cdictLen == 0: nothing to attach (:2779)
                    else { e.dPrefix = cdictEnd; e.wLow = e.wDictLimit = cdictEnd; e.loadedDictEnd = cdictEnd; e.dms = 1; }
                } else {
                    e.dPrefix = cdictEnd; e.wLow = 2; e.wDictLimit = cdictEnd; e.loadedDictEnd = L ? cdictEnd : 0;
                    if (e.wDictLimit - e.wLow < 8) e.wLow = e.wDictLimit;                             // ZSTD_window_update (ZstdCompressInternal.cs:742)
                }
                e.hufRepeat = DI->hufRepeat; e.fseValid = DI->fseValid;
            }
            e.windowLog = c.windowLog; e.hashLog = c.hashLog; e.chainLog = c.chainLog; e.minMatch = c.minMatch; e.strategy = c.strategy;
            // ZSTD_fast with an acceleration factor: probe pairs `stepSize` apart (hasStep = targetLength > 1, ZstdFast.cs:101, :334) and
            // leave the literals uncompressed (ZSTD_literalsCompressionIsDisabled, ZstdCompressInternal.cs:483-498)
            e.stepSize = (c.strategy == 1 && c.targetLength > 1) ? c.targetLength + 1 : 2;
            e.rawLits = (c.strategy == 1 && c.targetLength > 0) ? 1u : 0u;
            e.nbSeq = 0; e.lastLL = (uint32_t)std::min<size_t>(ss, kBlockSizeMax);
            e.rep[0] = 1; e.rep[1] = 4;                               // repStartValue (ZstdInternal.cs:13)
            if (DI) { e.rep[0] = DI->rep[0]; e.rep[1] = DI->rep[1]; }   // the dictionary's repcodes
            if (c.strategy == 0) protoErr = make_error(kParameterUnsupported);                   // level 4 outside its dfast sizes
            else if (ss > kEncMaxFrameBytes) protoErr = make_error(kSrcSizeWrong);                // positions are 31-bit
            else {
                protoBlk = ss <= kBlockSizeMax ? 1u : (uint32_t)((ss + kBlockSizeMax - 1) / kBlockSizeMax);
                if (ss >= 7) protoTable = ((size_t)1 << c.hashLog) + (c.strategy == 2 ? ((size_t)1 << c.chainLog) : 0);   // below: raw block, no match finding, no table
            }
        }
        EncItem& e = hi[i];
        e = proto;
        e.srcOff = srcOff[i]; e.dstOff = dstOff[i]; e.dstCap = (uint32_t)std::min<size_t>(dstCap[i], 0xFFFFFFF0u);
        if (DI) mb = true;                                            // dictionary frames always run as block waves (entropy state)
        nBlk[i] = protoBlk;
        if (protoErr) { hr[i] = protoErr; mb = true; continue; }
        if (protoBlk > 1) mb = true;
        maxWaves = std::max<size_t>(maxWaves, protoBlk); totalBlocks += protoBlk;
        e.tableOff = (uint32_t)tableEntries;
        tableEntries += protoTable;
    }
    lap("items");
    if (tableEntries >= 0xFFFFFFFFull) { t_encErr = "hash-table arena exceeds 32-bit indexing"; return false; }
    // work lists: per wave [ZSTD_fast groups | frames with a dictionary (serial kernel) | ZSTD_dfast groups | entropy]
    struct WaveLists { size_t off[4]; uint32_t n[4]; };
    std::vector<WaveLists> waves(maxWaves);
    size_t const listCap = 2 * totalBlocks + 4;
    if (!I.hWork.ensure(listCap * 4) || !I.workLists.ensure(listCap * 4)) { t_encErr = "out of memory (work lists)"; return false; }
    uint32_t* const hw = (uint32_t*)I.hWork.p;
    {
        std::vector<uint32_t> alive; alive.reserve(m);
        for (size_t i = 0; i < m; i++) if (nBlk[i]) alive.push_back((uint32_t)i);
        size_t pos = 0;
        std::vector<uint32_t> tmp[4];
        for (size_t b = 0; b < maxWaves; b++) {
            for (auto& t : tmp) t.clear();
            size_t keep = 0;
            for (size_t q = 0; q < alive.size(); q++) {
                uint32_t const i = alive[q];
                if (nBlk[i] <= b) continue;
                alive[keep++] = i;
                size_t const ss = srcSize[i];
                size_t const bs = std::min<size_t>(kBlockSizeMax, ss - b * (size_t)kBlockSizeMax);
                tmp[3].push_back(i);
                if (ss < 7 || bs < 7) continue;                   // ZSTD_buildSeqStore: noCompress (:3438)
                tmp[DI ? 1 : (hi[i].strategy == 1 ? 0 : 2)].push_back(i);
            }
            alive.resize(keep);
            for (int k = 0; k < 4; k++) {
                waves[b].off[k] = pos; waves[b].n[k] = (uint32_t)tmp[k].size();
                if (!tmp[k].empty()) memcpy(hw + pos, tmp[k].data(), tmp[k].size() * 4);
                pos += tmp[k].size();
            }
        }
    }
    if (!I.tables.ensure(tableEntries * 4 + 16) || !I.seqLL.ensure(m * (size_t)seqStride * 4) || !I.seqML.ensure(m * (size_t)seqStride * 4) ||
        !I.seqOF.ensure(m * (size_t)seqStride * 4) || !I.lit.ensure(m * (size_t)litStride) || !I.stateBits.ensure(m * (size_t)seqStride * 8) ||
        (mb && !I.hufState.ensure(m * 2 * (size_t)kHufStateSlot)) || !I.fseTabs.ensure(m * 3 * sizeof(FseGTable)) || !I.carry.ensure(m * sizeof(EntCarry))) { t_encErr = "out of memory (arena)"; return false; }
    lap("lists+mem");
    ENC_CUDA(cudaMemcpyAsync(I.items.p, hi, m * sizeof(EncItem), cudaMemcpyHostToDevice, copyStream));
    ENC_CUDA(cudaMemcpyAsync(I.workLists.p, I.hWork.p, listCap * 4, cudaMemcpyHostToDevice, copyStream));
    if (copyStream != stream) {                                   // order the kernels after the two small copies
        ENC_CUDA(cudaEventRecord(I.copied, copyStream));
        ENC_CUDA(cudaStreamWaitEvent(stream, I.copied, 0));
    }
    if (ev3) ENC_CUDA(cudaEventRecord(ev3[0], stream));
    if (tableEntries) ENC_CUDA(cudaMemsetAsync(I.tables.p, 0, tableEntries * 4, stream));      // tables are zeroed per frame (ZstdCompress.cs:2472,2481)
    EncPass p;
    p.items = (EncItem*)I.items.p; p.nItems = (uint32_t)m; p.src = d_src; p.dst = d_dst; p.tables = (uint32_t*)I.tables.p;
    p.seqLL = (uint32_t*)I.seqLL.p; p.seqML = (uint32_t*)I.seqML.p; p.seqOF = (uint32_t*)I.seqOF.p; p.litBuf = (uint8_t*)I.lit.p;
    p.stateBits = (uint64_t*)I.stateBits.p; p.results = hr; p.hufState = (uint8_t*)I.hufState.p; p.fseTabs = (FseGTable*)I.fseTabs.p; p.carry = (EntCarry*)I.carry.p; p.checksumFlag = checksumFlag ? 1u : 0u;
    p.seqStride = seqStride; p.litStride = litStride;
    memset(&p.dict, 0, sizeof(p.dict));
    if (DI) {
        p.dict.content = (const uint8_t*)DI->raw.p + DI->contentOff; p.dict.contentLen = DI->contentLen; p.dict.dictID = DI->dictID;
        p.dict.tables = (const uint32_t*)DI->tables.p; p.dict.hashLog = DI->cp.hashLog; p.dict.chainLog = DI->cp.chainLog;
        p.dict.huf = (const uint8_t*)DI->huf.p; p.dict.fse = (const FseGTable*)DI->fse.p;
        uint32_t const entries = (uint32_t)DI->tableEntries;
        enc_dict_init_kernel<<<dim3((unsigned)m, std::max(1u, std::min(64u, entries / 1024))), 256, 0, stream>>>(p, (uint32_t)m, entries);
        *launches += 1;
    }
    const uint32_t* const dw = (const uint32_t*)I.workLists.p;
    *launches += tableEntries ? 1 : 0;
    for (size_t b = 0; b < maxWaves; b++) {
        WaveLists const& w = waves[b];
        uint32_t const wave = (uint32_t)b;
        // lanes per chunk: 16 measured best for ZSTD_fast (8: 49/52 ms, 16: 42/49 ms, 32: 63/72 ms per GiB Silesia-mix / text)
        if (mb) {
            if (w.n[0]) enc_match_group_kernel<16, true><<<(w.n[0] + 2 * kMatchWarps - 1) / (2 * kMatchWarps), 32 * kMatchWarps, 0, stream>>>(p, dw + w.off[0], w.n[0], wave);
            if (w.n[1]) {
                static int const forceSerial = []() { const char* e = getenv("ZSTDB200_DICT_SERIAL"); return e ? atoi(e) : 0; }();   // developer knob: everything through the serial kernel
                enc_match_dict_fast_group_kernel<16><<<(w.n[1] + 2 * kMatchWarps - 1) / (2 * kMatchWarps), 32 * kMatchWarps, 0, stream>>>(p, dw + w.off[1], w.n[1], wave, (uint32_t)forceSerial);
                enc_match_dict_dfast_group_kernel<16><<<(w.n[1] + 2 * kMatchWarps - 1) / (2 * kMatchWarps), 32 * kMatchWarps, 0, stream>>>(p, dw + w.off[1], w.n[1], wave);
                static int const forced = []() { const char* e = getenv("ZSTDB200_DICT_SHIFT"); return e ? atoi(e) : -1; }();      // developer knob
                uint32_t const shift = forced >= 0 ? (uint32_t)std::min(forced, 5) : (w.n[1] >= 131072 ? 0u : (w.n[1] >= 32768 ? 2u : 5u));
                uint64_t const threads = (uint64_t)w.n[1] << shift;
                enc_match_dict_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(p, dw + w.off[1], w.n[1], wave, shift);
            }
            if (w.n[2]) enc_match_dfast_group_kernel<16, true><<<(w.n[2] + 2 * kMatchWarps - 1) / (2 * kMatchWarps), 32 * kMatchWarps, 0, stream>>>(p, dw + w.off[2], w.n[2], wave);
            if (ev3 && b == 0) ENC_CUDA(cudaEventRecord(ev3[1], stream));
            if (w.n[3]) {
                enc_entropy_kernel<true, 0><<<w.n[3], kEntThreads, 0, stream>>>(p, dw + w.off[3], wave);
                enc_fse_chain_kernel<true><<<(3 * w.n[3] + 31) / 32, 32, 32 * kChainTabStride, stream>>>(p, dw + w.off[3], w.n[3]);
                enc_entropy_kernel<true, 1><<<w.n[3], kEntThreads, 0, stream>>>(p, dw + w.off[3], wave);
            }
        } else {
            static int const gsFast = []() { const char* e = getenv("ZSTDB200_MATCH_GS"); return e ? atoi(e) : 16; }();      // developer knob
            if (w.n[0] && gsFast == 8) enc_match_group_kernel<8, false><<<(w.n[0] + 4 * kMatchWarps - 1) / (4 * kMatchWarps), 32 * kMatchWarps, 0, stream>>>(p, dw + w.off[0], w.n[0], 0u);
            else if (w.n[0]) enc_match_group_kernel<16, false><<<(w.n[0] + 2 * kMatchWarps - 1) / (2 * kMatchWarps), 32 * kMatchWarps, 0, stream>>>(p, dw + w.off[0], w.n[0], 0u);
            if (w.n[2]) enc_match_dfast_group_kernel<16, false><<<(w.n[2] + 2 * kMatchWarps - 1) / (2 * kMatchWarps), 32 * kMatchWarps, 0, stream>>>(p, dw + w.off[2], w.n[2], 0u);
            if (ev3) ENC_CUDA(cudaEventRecord(ev3[1], stream));
            enc_entropy_kernel<false, 0><<<(unsigned)m, kEntThreads, 0, stream>>>(p, dw, 0u);
            enc_fse_chain_kernel<false><<<(unsigned)(3 * m + 31) / 32, 32, 32 * kChainTabStride, stream>>>(p, dw, (uint32_t)m);
            enc_entropy_kernel<false, 1><<<(unsigned)m, kEntThreads, 0, stream>>>(p, dw, 0u);
        }
        *launches += (w.n[0] ? 1 : 0) + (w.n[1] ? 3 : 0) + (w.n[2] ? 1 : 0) + ((mb ? w.n[3] : 1) ? 3 : 0);
    }
    if (ev3) ENC_CUDA(cudaEventRecord(ev3[2], stream));
    ENC_CUDA(cudaGetLastError());
    lap("launched");
    return true;
}
// Shared-memory carve-out of the encoder's kernels.  An SM keeps ONE L1 / shared-memory split while it has resident CTAs, and a CTA of a
// kernel launched with a different split waits until the SM has drained.  With the defaults (match finder: no shared memory -> maximum
// L1; entropy kernels: 13.5 KB per CTA x 15 CTAs -> maximum shared memory) the entropy stage of a sub-batch did not start before the LAST
// match kernel of the pipelined host path had finished (measured: profiles/r01_notes.md, session 4).  overlap = true gives every encoder
// kernel the same split so that they can share SMs; overlap = false restores the defaults, which are faster when nothing runs beside them
// (match 42.5 vs 44.0 ms, entropy 11.0 vs 12.7 ms per GiB at a 50 % split).
void enc_set_overlap_mode(bool overlap)
{
    static std::atomic<int> applied[64];                        // per device: 0 unknown, 1 defaults, 2 common split
    int dev = 0; cudaGetDevice(&dev);
    int const want = overlap ? 2 : 1;
    if (applied[dev & 63].load(std::memory_order_acquire) == want) return;
    static int const pct = getenv("ZSTDB200_ENC_CARVEOUT") ? atoi(getenv("ZSTDB200_ENC_CARVEOUT")) : 40;   // 100 KB of shared memory per SM: 28 % 77.4 ms, 40 % 74.0 ms, 50 % 75.4 ms per GiB host to host
    int const x = overlap ? pct : (int)cudaSharedmemCarveoutDefault;
    auto const A = cudaFuncAttributePreferredSharedMemoryCarveout;
    cudaFuncSetAttribute(enc_match_group_kernel<16, false>, A, x);
    cudaFuncSetAttribute(enc_match_group_kernel<8, false>, A, x);
    cudaFuncSetAttribute(enc_match_group_kernel<16, true>, A, x);
    cudaFuncSetAttribute(enc_match_dfast_group_kernel<16, false>, A, x);
    cudaFuncSetAttribute(enc_match_dfast_group_kernel<16, true>, A, x);
    cudaFuncSetAttribute(enc_match_dict_fast_group_kernel<16>, A, x);
    cudaFuncSetAttribute(enc_match_dict_dfast_group_kernel<16>, A, x);
    cudaFuncSetAttribute(enc_match_dict_kernel, A, x);
    cudaFuncSetAttribute(enc_entropy_kernel<false, 0>, A, x);
    cudaFuncSetAttribute(enc_entropy_kernel<false, 1>, A, x);
    cudaFuncSetAttribute(enc_entropy_kernel<true, 0>, A, x);
    cudaFuncSetAttribute(enc_entropy_kernel<true, 1>, A, x);
    cudaFuncSetAttribute(enc_fse_chain_kernel<false>, A, x);
    cudaFuncSetAttribute(enc_fse_chain_kernel<true>, A, x);
    cudaFuncSetAttribute(enc_compact_kernel, A, x);
    (void)cudaGetLastError();
    applied[dev & 63].store(want, std::memory_order_release);
}
const uint64_t* enc_results(const EncArena& A) { return A.impl ? (const uint64_t*)A.impl->hResults.p : nullptr; }

// More than one pass (n > 8192): with a second arena B the passes alternate between the two, so that the host's preparation of a pass
// (descriptors, work lists) and the read-back of the one before overlap the kernels -- many small records are otherwise bound by that.
bool enc_compress_device(EncArena& A, cudaStream_t stream, cudaEvent_t* ev, size_t n, int level, int checksumFlag,
                         const uint8_t* d_src, const uint64_t* srcOff, const size_t* srcSize,
                         uint8_t* d_dst, const uint64_t* dstOff, const size_t* dstCap, size_t* result,
                         float* timings, unsigned* launches, const EncDict* dict, EncArena* B)
{
    (void)ev;
    float msAll = 0, msMatch = 0, msEnt = 0;
    enc_set_overlap_mode(false);                 // the passes run one after the other on one stream: every kernel keeps its own L1 / shared-memory split
    size_t maxBlkAll = 0;
    for (size_t i = 0; i < n; i++) maxBlkAll = std::max(maxBlkAll, std::min<size_t>(srcSize[i], kBlockSizeMax));
    size_t const passCap = enc_pass_capacity(maxBlkAll);
    EncArena* const ar[2] = {&A, (B && n > passCap) ? B : &A};
    int const nAr = ar[1] != ar[0] ? 2 : 1;
    struct Pending { bool on = false; size_t base = 0, m = 0; } pend[2];
    for (auto* a : ar) if (!a->impl) a->impl = new EncArenaImpl();
    auto finish = [&](int k) -> bool {
        Pending& pd = pend[k];
        if (!pd.on) return true;
        EncArenaImpl& I = *ar[k]->impl;
        ENC_CUDA(cudaEventSynchronize(I.tev[2]));
        float t;
        cudaEventElapsedTime(&t, I.tev[0], I.tev[2]); msAll += t;
        cudaEventElapsedTime(&t, I.tev[0], I.tev[1]); msMatch += t;
        cudaEventElapsedTime(&t, I.tev[1], I.tev[2]); msEnt += t;
        const uint64_t* hr = enc_results(*ar[k]);
        for (size_t i = 0; i < pd.m; i++) {
            size_t const ss = srcSize[pd.base + i];
            result[pd.base + i] = ss > kEncMaxFrameBytes ? (size_t)make_error(kSrcSizeWrong) : (size_t)hr[i];
        }
        pd.on = false;
        return true;
    };
    int k = 0;
    for (size_t base = 0; base < n; base += passCap, k = (k + 1) % nAr) {
        size_t const m = std::min(passCap, n - base);
        if (!finish(k)) return false;
        if (!enc_enqueue(*ar[k], stream, stream, m, level, checksumFlag, d_src, srcOff + base, srcSize + base, d_dst, dstOff + base, dstCap + base, ar[k]->impl->tev, launches, dict)) return false;
        pend[k].on = true; pend[k].base = base; pend[k].m = m;
    }
    for (int q = 0; q < nAr; q++, k = (k + 1) % nAr) if (!finish(k)) return false;
    ENC_CUDA(cudaStreamSynchronize(stream));
    ENC_CUDA(cudaGetLastError());
    timings[1] = msAll; timings[8] = msMatch; timings[9] = msEnt;
    return true;
}

// Gathers n variable-size frames into the arena's dense buffer; the offset / size arrays stay in pinned host memory and
// are read by the kernel directly.
bool enc_compact_device(EncArena& A, cudaStream_t stream, size_t n, const uint8_t* d_dst, const uint64_t* dstOff,
                        const size_t* sizes, const uint64_t* cOff, size_t total, unsigned* launches)
{
    if (!A.impl) A.impl = new EncArenaImpl();
    EncArenaImpl& I = *A.impl;
    if (!I.compact.ensure(total + 16) || !I.hC.ensure(n * 24)) { t_encErr = "out of memory (compact)"; return false; }
    uint64_t* h = (uint64_t*)I.hC.p;
    for (size_t i = 0; i < n; i++) { h[i] = dstOff[i]; h[n + i] = is_error(sizes[i]) ? 0 : sizes[i]; h[2 * n + i] = cOff[i]; }
    enc_compact_kernel<<<(unsigned)n, 256, 0, stream>>>(d_dst, h, h + n, h + 2 * n, (uint8_t*)I.compact.p);
    *launches += 1;
    return true;
}

}  // namespace zb
