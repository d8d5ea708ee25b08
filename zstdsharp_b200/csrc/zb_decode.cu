// zb_decode.cu -- batch Zstandard frame decoder for sm_100a (product code; no CPU fallback).
//
// Replaces, for batches of independent frames, the reference's
//   ZSTD_decompressDCtx -> ZSTD_decompressMultiFrame -> ZSTD_decompressFrame        (ZstdDecompress.cs:1365,1216,1062)
//     -> ZSTD_decompressBlock_internal                                             (ZstdDecompressBlock.cs:3090)
//        -> ZSTD_decodeLiteralsBlock / HUF_decompress{1,4}X1_usingDTable           (:88, HufDecompress.cs:312,342)
//        -> ZSTD_decodeSeqHeaders / ZSTD_buildFSETable                             (:1845, :1571)
//        -> ZSTD_decompressSequences_body = ZSTD_decodeSequence + ZSTD_execSequence (:2668, :2360, :2187)
// Kernels (one wave = the k-th block of every item):
//   dec_scan_kernel      thread / item   counts blocks (number of waves), initialises cursors
//   dec_setup_kernel     warp   / item   frame+block+literal+sequence headers, Huffman and FSE table construction
//   dec_huf_kernel       lane   / stream HUF 4-stream (or single) literal decoding, tables staged in shared memory
//   dec_seq_kernel       lane   / item   3-state FSE sequence decoding (serial chain), tables staged in shared memory, records staged and flushed
//   dec_exec_kernel      warp   / item   literal + match copies through a shared-memory output tile, raw/RLE blocks, XXH64 check
//   dec_dict_kernel      thread          parses a loaded dictionary once (entropy tables, repcodes); dec_dict_prefill_kernel copies its
//                                        content in front of every item's output slot
#include <algorithm>
#include "zb_decode.cuh"

namespace zb {

// =====================================================================================================
//  Exact little restatement of BIT_DStream_t for the tiny FSE-compressed Huffman weight stream
//  (Bitstream.cs:172-425).  Only used on <= 128-byte streams by one lane; speed is irrelevant, the
//  end-of-stream semantics (FSE_decompress_usingDTable_generic, FseDecompress.cs:230) are what matters.
// =====================================================================================================
struct SmallDStream {
    uint64_t container; uint32_t consumed; int32_t ptr; int32_t size; const uint8_t* base;
    __device__ uint64_t load64(int32_t at) const {   // MEM_readLEST at base+at; bytes outside the stream read as 0
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) { int32_t a = at + i; uint32_t b = (a >= 0 && a < size) ? base[a] : 0u; v |= (uint64_t)b << (8 * i); }
        return v;
    }
    __device__ bool init(const uint8_t* b, int32_t n) {
        base = b; size = n;
        if (n < 1) return false;
        uint32_t last = b[n - 1];
        if (last == 0) return false;
        if (n >= 8) { ptr = n - 8; container = load64(ptr); consumed = 8 - highbit32(last); }
        else { ptr = 0; container = load64(0); consumed = 8 - highbit32(last) + (uint32_t)(8 - n) * 8; }
        return true;
    }
    __device__ uint32_t read(uint32_t nb) {   // BIT_readBits (masked shift, as in BIT_getMiddleBits)
        uint32_t start = 64u - consumed - nb;
        uint64_t v = (container >> (start & 63)) & ((1ull << nb) - 1);
        consumed += nb;
        return (uint32_t)v;
    }
    // BIT_reloadDStream: 0 unfinished, 1 endOfBuffer, 2 completed, 3 overflow
    __device__ int reload() {
        if (consumed > 64) return 3;
        if (ptr >= 8) { ptr -= (int32_t)(consumed >> 3); consumed &= 7; container = load64(ptr); return 0; }
        if (ptr == 0) return consumed < 64 ? 1 : 2;
        uint32_t nbBytes = consumed >> 3; int res = 0;
        if ((int32_t)nbBytes > ptr) { nbBytes = (uint32_t)ptr; res = 1; }
        ptr -= (int32_t)nbBytes; consumed -= nbBytes * 8; container = load64(ptr);
        return res;
    }
};

// FSE_readNCount_body (EntropyCommon.cs:52).  Clean forward bit reader; bits past the end read as zero and
// make the header invalid (the reference reaches the same verdict through its bitCount > 32 test).
// Returns bytes consumed, or 0 on error.
__device__ uint32_t fse_read_ncount(int16_t* norm, uint32_t* maxSVPtr, uint32_t* tableLogPtr, const uint8_t* src, uint32_t hbSize)
{
    uint32_t const maxSV1 = *maxSVPtr + 1;
    uint64_t acc = 0; uint32_t accBits = 0; uint32_t bytePos = 0; uint32_t bitsUsed = 0;
    auto fill = [&]() { while (accBits <= 56) { uint64_t b = bytePos < hbSize ? src[bytePos] : 0u; acc |= b << accBits; accBits += 8; bytePos++; } };
    auto take = [&](uint32_t nb) { acc >>= nb; accBits -= nb; bitsUsed += nb; };
    for (uint32_t s = 0; s < maxSV1; s++) norm[s] = 0;
    fill();
    int nbBits = (int)(acc & 0xF) + 5;
    if (nbBits > 15) return 0;
    take(4);
    *tableLogPtr = (uint32_t)nbBits;
    int remaining = (1 << nbBits) + 1;
    int threshold = 1 << nbBits;
    nbBits++;
    uint32_t charnum = 0; bool previous0 = false;
    for (;;) {
        if (previous0) {
            fill();
            while ((acc & 3) == 3) { charnum += 3; take(2); fill(); if (charnum >= maxSV1 + 64) return 0; }
            charnum += (uint32_t)(acc & 3); take(2);
            if (charnum >= maxSV1) break;
        }
        fill();
        {   int const max = (2 * threshold - 1) - remaining;
            int count;
            if ((int)(acc & (uint32_t)(threshold - 1)) < max) { count = (int)(acc & (uint32_t)(threshold - 1)); take((uint32_t)nbBits - 1); }
            else { count = (int)(acc & (uint32_t)(2 * threshold - 1)); if (count >= threshold) count -= max; take((uint32_t)nbBits); }
            count--;
            if (count >= 0) remaining -= count; else remaining += count;
            norm[charnum++] = (int16_t)count;
            previous0 = (count == 0);
            if (remaining < threshold) {
                if (remaining <= 1) break;
                nbBits = (int)highbit32((uint32_t)remaining) + 1;
                threshold = 1 << (nbBits - 1);
            }
            if (charnum >= maxSV1) break;
        }
    }
    if (remaining != 1) return 0;
    if (charnum > maxSV1) return 0;
    if (bitsUsed > 8 * hbSize) return 0;
    *maxSVPtr = charnum - 1;
    return (bitsUsed + 7) >> 3;
}

// Symbol spread shared by FSE_buildDTable_internal (FseDecompress.cs:25) and ZSTD_buildFSETable_body
// (ZstdDecompressBlock.cs:1571).  cell[] receives the symbol of every state, symNext[] the first "next" value.
__device__ bool fse_spread(uint8_t* cell, uint16_t* symNext, const int16_t* norm, uint32_t maxSV, uint32_t tableLog)
{
    uint32_t const tableSize = 1u << tableLog, tableMask = tableSize - 1;
    uint32_t const step = (tableSize >> 1) + (tableSize >> 3) + 3;
    uint32_t high = tableSize - 1, position = 0;
    for (uint32_t s = 0; s <= maxSV; s++) {
        if (norm[s] == -1) { cell[high--] = (uint8_t)s; symNext[s] = 1; } else symNext[s] = (uint16_t)norm[s];
    }
    for (uint32_t s = 0; s <= maxSV; s++) {
        for (int i = 0; i < norm[s]; i++) {
            cell[position] = (uint8_t)s;
            position = (position + step) & tableMask;
            while (position > high) position = (position + step) & tableMask;
        }
    }
    return position == 0;
}

// Per-warp scratch of the setup kernel
struct SetupScratch {
    uint8_t weights[260];
    uint8_t cell[512];
    uint16_t symNext[64];
    uint16_t start[260];
    int16_t norm[64];
    uint32_t rankStart[16];
    // sequence tables to be built by the whole warp after lane 0 has parsed the headers (index: 0 LL, 1 ML, 2 OF)
    int16_t normT[3][64];
    uint16_t cum[65];
    uint32_t tabAct[3];      // 0 nothing (repeat), 1 rle, 2 predefined, 3 compressed
    uint32_t tabArg[3];      // rle: symbol; compressed: maxSV | tableLog << 8
};

// HUF_readStats_body (EntropyCommon.cs:292) executed by one lane. Returns header size (iSize+1) or 0 on error.
__device__ uint32_t huf_read_stats(SetupScratch& sc, uint32_t* nbSymbolsPtr, uint32_t* tableLogPtr, const uint8_t* src, uint32_t srcSize)
{
    if (srcSize == 0) return 0;
    uint32_t iSize = src[0], oSize;
    if (iSize >= 128) {
        oSize = iSize - 127; iSize = (oSize + 1) / 2;
        if (iSize + 1 > srcSize) return 0;
        if (oSize >= 256) return 0;
        for (uint32_t n = 0; n < oSize; n += 2) { sc.weights[n] = src[1 + n / 2] >> 4; sc.weights[n + 1] = src[1 + n / 2] & 15; }
    } else {
        if (iSize + 1 > srcSize) return 0;
        // FSE_decompress_wksp_body (FseDecompress.cs:334), maxLog 6, dst capacity 255
        uint32_t tableLog, maxSV = 255;
        // norm for up to 256 symbols would not fit sc.norm; weights are <= 12 so any symbol > 12 is invalid anyway,
        // but the header may legally *mention* zero-probability symbols up to 255: parse into a local array.
        int16_t normW[256];
        uint32_t const hs = fse_read_ncount(normW, &maxSV, &tableLog, src + 1, iSize);
        if (hs == 0) return 0;
        if (tableLog > 6) return 0;
        // FSE_buildDTable_internal: <= 64 cells
        uint8_t cellS[64]; uint16_t next[256]; uint8_t cellBits[64]; uint16_t cellNew[64];
        {
            uint32_t const tableSize = 1u << tableLog, tableMask = tableSize - 1;
            uint32_t const step = (tableSize >> 1) + (tableSize >> 3) + 3;
            uint32_t high = tableSize - 1, position = 0;
            for (uint32_t s = 0; s <= maxSV; s++) { if (normW[s] == -1) { cellS[high--] = (uint8_t)s; next[s] = 1; } else next[s] = (uint16_t)normW[s]; }
            for (uint32_t s = 0; s <= maxSV; s++)
                for (int i = 0; i < normW[s]; i++) {
                    cellS[position] = (uint8_t)s; position = (position + step) & tableMask;
                    while (position > high) position = (position + step) & tableMask;
                }
            if (position != 0) return 0;
            for (uint32_t u = 0; u < tableSize; u++) {
                uint32_t const sym = cellS[u]; uint32_t const ns = next[sym]++;
                cellBits[u] = (uint8_t)(tableLog - highbit32(ns));
                cellNew[u] = (uint16_t)((ns << cellBits[u]) - tableSize);
            }
        }
        // FSE_decompress_usingDTable_generic (FseDecompress.cs:230), 64-bit variant
        SmallDStream bd;
        if (!bd.init(src + 1 + hs, (int32_t)(iSize - hs))) return 0;
        uint32_t s1 = bd.read(tableLog); bd.reload();
        uint32_t s2 = bd.read(tableLog); bd.reload();
        uint32_t const omax = 255; uint32_t op = 0;
        auto dec = [&](uint32_t& st) -> uint8_t { uint8_t sym = cellS[st]; uint32_t nb = cellBits[st]; uint32_t low = bd.read(nb); st = cellNew[st] + low; return sym; };
        for (; (bd.reload() == 0) && (op + 3 < omax); op += 4) {
            sc.weights[op] = dec(s1); sc.weights[op + 1] = dec(s2); sc.weights[op + 2] = dec(s1); sc.weights[op + 3] = dec(s2);
        }
        for (;;) {
            if (op + 2 > omax) return 0;
            sc.weights[op++] = dec(s1);
            if (bd.reload() == 3) { sc.weights[op++] = dec(s2); break; }
            if (op + 2 > omax) return 0;
            sc.weights[op++] = dec(s2);
            if (bd.reload() == 3) { sc.weights[op++] = dec(s1); break; }
        }
        oSize = op;
    }
    uint32_t rankStats[13]; for (int i = 0; i < 13; i++) rankStats[i] = 0;
    uint32_t weightTotal = 0;
    for (uint32_t n = 0; n < oSize; n++) {
        if (sc.weights[n] > kHufTableLogMax) return 0;
        rankStats[sc.weights[n]]++;
        weightTotal += (1u << sc.weights[n]) >> 1;
    }
    if (weightTotal == 0) return 0;
    uint32_t const tableLog = highbit32(weightTotal) + 1;
    if (tableLog > kHufTableLogMax) return 0;
    {
        uint32_t const total = 1u << tableLog, rest = total - weightTotal;
        uint32_t const verif = 1u << highbit32(rest), lastWeight = highbit32(rest) + 1;
        if (verif != rest) return 0;
        sc.weights[oSize] = (uint8_t)lastWeight;
        rankStats[lastWeight]++;
    }
    if ((rankStats[1] < 2) || (rankStats[1] & 1)) return 0;
    *nbSymbolsPtr = oSize + 1; *tableLogPtr = tableLog;
    // start index of every symbol in the decode table: symbols sorted by (weight, symbol) (HufDecompress.cs:131-252)
    uint32_t acc = 0;
    for (uint32_t w = 1; w <= tableLog; w++) { sc.rankStart[w] = acc; acc += rankStats[w] * ((1u << w) >> 1); }
    for (uint32_t s = 0; s < oSize + 1; s++) {
        uint32_t const w = sc.weights[s];
        if (w) { sc.start[s] = (uint16_t)sc.rankStart[w]; sc.rankStart[w] += (1u << w) >> 1; }
    }
    return iSize + 1;
}

// ZSTD_buildFSETable_body (ZstdDecompressBlock.cs:1571) into the compact entry format, one lane.
__device__ void build_seq_table(uint32_t* out, SetupScratch& sc, uint32_t maxSV, uint32_t tableLog, int kind /*0 LL 1 ML 2 OF*/)
{
    uint32_t const tableSize = 1u << tableLog;
    fse_spread(sc.cell, sc.symNext, sc.norm, maxSV, tableLog);
    for (uint32_t u = 0; u < tableSize; u++) {
        uint32_t const sym = sc.cell[u];
        uint32_t const ns = sc.symNext[sym]++;
        uint32_t const nbBits = tableLog - highbit32(ns);
        uint32_t const next = (ns << nbBits) - tableSize;
        uint32_t const add = kind == 0 ? c_LL_bits[sym] : (kind == 1 ? c_ML_bits[sym] : sym);
        out[u] = fse_pack(nbBits, add, sym, next);
    }
}

// Warp-parallel ZSTD_buildFSETable_body (ZstdDecompressBlock.cs:1571).  The serial algorithm walks the table with a
// fixed odd step and skips the cells above highThreshold; here every lane evaluates walk index j directly
// (cell = j*step mod size), a ballot-prefix gives the rank of the cell among the cells that are kept, and the rank is
// mapped to its symbol through the cumulated counts.  The second pass needs, for each cell, how many earlier cells hold
// the same symbol: __match_any_sync per 32 cells plus a running counter per symbol.
__device__ void build_seq_table_warp(uint32_t* out, SetupScratch& sc, const int16_t* norm, uint32_t maxSV, uint32_t tableLog, int kind, uint32_t lane)
{
    uint32_t const FULL = 0xFFFFFFFFu, ltMask = (1u << lane) - 1u;
    uint32_t const tableSize = 1u << tableLog, tableMask = tableSize - 1;
    uint32_t const step = (tableSize >> 1) + (tableSize >> 3) + 3;
    // pass 0: low-probability symbols take the top cells in symbol order; cumulated counts of the others
    uint32_t nLow = 0, run = 0;
    for (uint32_t s0 = 0; s0 <= maxSV; s0 += 32) {
        uint32_t const s = s0 + lane;
        int const nv = s <= maxSV ? (int)norm[s] : 0;
        bool const low = nv == -1;
        uint32_t const lowBallot = __ballot_sync(FULL, low);
        if (low) { sc.cell[tableSize - 1 - (nLow + __popc(lowBallot & ltMask))] = (uint8_t)s; }
        if (s <= maxSV) sc.symNext[s] = low ? 1 : (uint16_t)nv;
        uint32_t const c = nv > 0 ? (uint32_t)nv : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t const o = __shfl_up_sync(FULL, incl, d); if (lane >= (uint32_t)d) incl += o; }
        if (s <= maxSV) sc.cum[s] = (uint16_t)(run + incl - c);
        run += __shfl_sync(FULL, incl, 31);
        nLow += __popc(lowBallot);
    }
    if (lane == 0) sc.cum[maxSV + 1] = (uint16_t)run;
    __syncwarp();
    uint32_t const high = tableSize - 1 - nLow;
    // pass 1: spread
    uint32_t placed = 0;
    for (uint32_t j0 = 0; j0 < tableSize; j0 += 32) {
        uint32_t const j = j0 + lane;
        uint32_t const pos = (j * step) & tableMask;
        bool const keep = j < tableSize && pos <= high;
        uint32_t const kb = __ballot_sync(FULL, keep);
        if (keep) {
            uint32_t const k = placed + __popc(kb & ltMask);
            uint32_t lo = 0, hi = maxSV;                   // largest s with cum[s] <= k (symbols with count 0 share a value: take the last)
            while (lo < hi) { uint32_t const mid = (lo + hi + 1) >> 1; if (sc.cum[mid] <= k) lo = mid; else hi = mid - 1; }
            sc.cell[pos] = (uint8_t)lo;
        }
        placed += __popc(kb);
    }
    __syncwarp();
    // pass 2: entries in cell order
    for (uint32_t u0 = 0; u0 < tableSize; u0 += 32) {
        uint32_t const u = u0 + lane;
        bool const v = u < tableSize;
        uint32_t const sym = v ? sc.cell[u] : 0xFFu;
        uint32_t const peers = __match_any_sync(FULL, sym);
        uint32_t const base = v ? sc.symNext[sym] : 0u;
        __syncwarp();
        uint32_t const ns = base + __popc(peers & ltMask);
        if (v && (peers >> lane) == 1u) sc.symNext[sym] = (uint16_t)(base + __popc(peers));     // highest lane of the group
        __syncwarp();
        if (v) {
            uint32_t const nbBits = tableLog - highbit32(ns);
            uint32_t const next = (ns << nbBits) - tableSize;
            uint32_t const add = kind == 0 ? c_LL_bits[sym] : (kind == 1 ? c_ML_bits[sym] : sym);
            out[u] = fse_pack(nbBits, add, sym, next);
        }
    }
    __syncwarp();
}

__global__ void dec_default_tables_kernel(uint32_t* out)
{
    __shared__ SetupScratch sc;
    if (threadIdx.x != 0) return;
    for (int i = 0; i <= kMaxLL; i++) sc.norm[i] = c_LL_defaultNorm[i];
    build_seq_table(out + kFseLLOff, sc, kMaxLL, kLLDefaultNormLog, 0);
    for (int i = 0; i <= kMaxML; i++) sc.norm[i] = c_ML_defaultNorm[i];
    build_seq_table(out + kFseMLOff, sc, kMaxML, kMLDefaultNormLog, 1);
    for (int i = 0; i <= kDefaultMaxOff; i++) sc.norm[i] = c_OF_defaultNorm[i];
    build_seq_table(out + kFseOFOff, sc, kDefaultMaxOff, kOFDefaultNormLog, 2);
}
void dec_build_default_tables(uint32_t* d, cudaStream_t s) { dec_default_tables_kernel<<<1, 32, 0, s>>>(d); }

// ZSTD_decompress_insertDictionary / ZSTD_loadDEntropy (ZstdDecompress.cs:1880, :1770), one thread, once per loaded dictionary.
// A dictionary without the magic number is pure content.  The Huffman table is stored in the single-symbol form the literal
// kernel reads (the reference builds the double-symbol form of the same code).
// encStats (optional, for the encode side's CDict: enc_dict_build_kernel): [0..255] Huffman weights | [256] Huffman tableLog | [257] nbSymbols |
// then for OF, ML, LL: 64 normalized counts (sign-extended), maxSymbolValue, tableLog (kEncDictStatsWords words).
__global__ void dec_dict_kernel(const uint8_t* dict, uint32_t dictSize, uint16_t* hufOut, uint32_t* fseOut, uint32_t* info, uint32_t* encStats)
{
    __shared__ SetupScratch sc;
    if (threadIdx.x != 0) return;
    for (uint32_t i = 0; i < kDictInfoWords; i++) info[i] = 0;
    uint32_t pos = 0, status = 1;
    if (dictSize >= 8 && ld_le32(dict) == 0xEC30A437u) {
        info[1] = ld_le32(dict + 4);
        pos = 8;
        do {
            uint32_t nbSym = 0, tlog = 0;
            uint32_t const hs = huf_read_stats(sc, &nbSym, &tlog, dict + pos, dictSize - pos);
            if (hs == 0) { status = 2; break; }
            for (uint32_t s = 0; s < nbSym; s++) {
                uint32_t const w = sc.weights[s];
                if (!w) continue;
                uint32_t const len = (1u << w) >> 1, st = sc.start[s];
                for (uint32_t u = 0; u < len; u++) hufOut[st + u] = (uint16_t)((s << 8) | (tlog + 1 - w));
            }
            info[3] = tlog; pos += hs;
            if (encStats) { for (uint32_t q = 0; q < 256; q++) encStats[q] = q < nbSym ? sc.weights[q] : 0u; encStats[256] = tlog; encStats[257] = nbSym; }
            int const kinds[3] = {2, 1, 0};                         // stored in the order OF, ML, LL
            for (int q = 0; q < 3 && status == 1; q++) {
                int const kind = kinds[q];
                uint32_t const maxSym = kind == 0 ? kMaxLL : (kind == 1 ? kMaxML : kMaxOff), maxLog = kind == 2 ? kOffFSELog : kLLFSELog;
                uint32_t maxSV = maxSym, tableLog = 0;
                uint32_t const h = fse_read_ncount(sc.norm, &maxSV, &tableLog, dict + pos, dictSize - pos);
                if (h == 0 || maxSV > maxSym || tableLog > maxLog) { status = 2; break; }
                if (encStats) { uint32_t* const e = encStats + 258 + q * 66; for (uint32_t u = 0; u < 64; u++) e[u] = (uint32_t)(int32_t)(u <= maxSym ? sc.norm[u] : (int16_t)0); e[64] = maxSV; e[65] = tableLog; }
                build_seq_table(fseOut + (kind == 0 ? kFseLLOff : (kind == 1 ? kFseMLOff : kFseOFOff)), sc, maxSV, tableLog, kind);
                info[kind == 0 ? 4 : (kind == 1 ? 6 : 5)] = tableLog;
                pos += h;
            }
            if (status != 1) break;
            if (pos + 12 > dictSize) { status = 2; break; }
            uint32_t const contentSize = dictSize - (pos + 12);
            for (int i = 0; i < 3; i++) {
                uint32_t const r = ld_le32(dict + pos); pos += 4;
                if (r == 0 || r > contentSize) status = 2;
                info[7 + i] = r;
            }
            info[2] = 1;
        } while (0);
    }
    info[10] = pos; info[11] = status == 1 ? dictSize - pos : 0;
    info[0] = status;
}
void dec_launch_dict_setup(const uint8_t* d_dict, uint32_t dictSize, uint16_t* d_huf, uint32_t* d_fse, uint32_t* d_info, cudaStream_t s, uint32_t* d_encStats)
{ dec_dict_kernel<<<1, 32, 0, s>>>(d_dict, dictSize, d_huf, d_fse, d_info, d_encStats); }

__global__ void dec_dict_prefill_kernel(DecPass p, uint32_t contentOff, uint32_t contentSize)
{
    uint32_t const i = blockIdx.x;
    uint8_t* const d = p.dst + p.items[i].dstOff - contentSize;
    const uint8_t* const s = p.dictBytes + contentOff;
    for (uint32_t k = threadIdx.x; k < contentSize; k += blockDim.x) d[k] = s[k];
}
void dec_launch_dict_prefill(const DecPass& p, uint32_t contentOff, uint32_t contentSize, cudaStream_t s)
{ if (contentSize) dec_dict_prefill_kernel<<<p.nItems, 256, 0, s>>>(p, contentOff, contentSize); }

// =====================================================================================================
//  Frame walking helpers
// =====================================================================================================
__device__ __forceinline__ uint32_t frame_header_size(uint32_t fhd)   // ZSTD_frameHeaderSize_internal, ZstdDecompress.cs:427
{
    uint32_t const dictID = fhd & 3, single = (fhd >> 5) & 1, fcsId = fhd >> 6;
    uint32_t const did[4] = {0, 1, 2, 4}; uint32_t const fcs[4] = {0, 2, 4, 8};
    return 5 + !single + did[dictID] + fcs[fcsId] + (single && !fcsId);
}

// Skips skippable frames; decides whether the item is finished (ZSTD_decompressMultiFrame loop, ZstdDecompress.cs:1216-1321).
// Returns true when a regular frame (or garbage to be diagnosed) starts at *pos.
__device__ bool advance_frames(DecItem& it, const uint8_t* src, uint32_t* posIo)
{
    uint32_t pos = *posIo;
    for (;;) {
        uint32_t const rem = it.srcSize - pos;
        if (rem < 5) {
            if (rem != 0) { it.status = kStError; it.errCode = kSrcSizeWrong; } else it.status = kStDone;
            *posIo = pos; return false;
        }
        uint32_t const magic = ld_le32(src + pos);
        if ((magic & kMagicSkippableMask) == kMagicSkippableStart) {   // readSkippableFrameSize, :674
            if (rem < 8) { it.status = kStError; it.errCode = kSrcSizeWrong; *posIo = pos; return false; }
            uint32_t const sz = ld_le32(src + pos + 4);
            if ((uint32_t)(sz + 8) < sz) { it.status = kStError; it.errCode = kFrameParameterUnsupported; *posIo = pos; return false; }
            if ((uint64_t)sz + 8 > rem) { it.status = kStError; it.errCode = kSrcSizeWrong; *posIo = pos; return false; }
            pos += sz + 8;
            continue;
        }
        *posIo = pos; return true;
    }
}

// Thread per item: initialise the cursor and count blocks = number of waves this item needs (walks like
// ZSTD_findFrameSizeInfo, ZstdDecompress.cs:877; errors are diagnosed later by the setup kernel).
__global__ void dec_scan_kernel(DecPass p, const DecItemInit* init)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nItems) return;
    DecItem& it = p.items[i];
    DecItemInit const in = init[i];
    it.srcOff = in.srcOff; it.dstOff = in.dstOff; it.srcSize = in.srcSize; it.dstCap = in.dstCap;
    it.srcPos = 0; it.outPos = 0; it.frameStart = 0; it.status = kStRunning; it.errCode = 0; it.inFrame = 0;
    it.moreThan1Frame = 0; it.checksumFlag = 0; it.hasFcs = 0; it.fcs = 0; it.litEntropy = 0; it.fseEntropy = 0;
    it.blkType = kBlkNone; it.prefix = 0; it.hufX2 = 0;
    uint32_t const blocks = count_item_blocks(p.src + in.srcOff, in.srcSize);
    atomicMax(&p.counters[3], blocks);
}

// HUF_selectDecoder (HufDecompress.cs:1688; algoTime :1471-1681 = {tableTime, decode256Time} of the single- and the double-symbol
// decoder per compression-ratio bucket).  Which of the two the reference runs does not change the bytes of a valid stream, only the
// verdict on some damaged ones (huf_x2_replay below).
__device__ __constant__ uint16_t c_hufAlgoTime[16][4] = {
    {0, 0, 1, 1}, {0, 0, 1, 1}, {150, 216, 381, 119}, {170, 205, 514, 112}, {177, 199, 539, 110}, {197, 194, 644, 107},
    {221, 192, 735, 107}, {256, 189, 881, 106}, {359, 188, 1167, 109}, {582, 187, 1570, 114}, {688, 187, 1712, 122},
    {825, 186, 1965, 136}, {976, 185, 2131, 150}, {1180, 186, 2070, 175}, {1377, 185, 1731, 202}, {1412, 185, 1695, 202}};
__device__ uint32_t huf_select_decoder(uint32_t dstSize, uint32_t cSrcSize)
{
    uint32_t const Q = cSrcSize >= dstSize ? 15u : cSrcSize * 16u / dstSize;
    uint32_t const D256 = dstSize >> 8;
    uint32_t const t0 = c_hufAlgoTime[Q][0] + c_hufAlgoTime[Q][1] * D256;
    uint32_t t1 = c_hufAlgoTime[Q][2] + c_hufAlgoTime[Q][3] * D256;
    t1 += t1 >> 5;
    return t1 < t0 ? 1u : 0u;
}

// =====================================================================================================
//  Setup kernel: one warp per item; lane 0 parses, the warp fills the Huffman table.
// =====================================================================================================
constexpr int kSetupWarps = 4;

__device__ void setup_fail(DecItem& it, uint32_t code) { it.status = kStError; it.errCode = code; it.blkType = kBlkNone; }

// Parses the sequences section header and builds/keeps the three tables. Returns header bytes or 0xFFFFFFFF on error
// (error code in *err).  ZSTD_decodeSeqHeaders, ZstdDecompressBlock.cs:1845; ZSTD_buildSeqTable :1746.
__device__ uint32_t setup_seq_headers(DecItem& it, const DecPass& p, uint32_t item, SetupScratch& sc, const uint8_t* src, uint32_t srcSize, uint32_t* err)
{
    uint32_t ip = 0;
    if (srcSize < 1) { *err = kSrcSizeWrong; return 0xFFFFFFFFu; }
    uint32_t nbSeq = src[ip++];
    if (nbSeq == 0) { it.nbSeq = 0; if (srcSize != 1) { *err = kSrcSizeWrong; return 0xFFFFFFFFu; } return 1; }
    if (nbSeq > 0x7F) {
        if (nbSeq == 0xFF) { if (ip + 2 > srcSize) { *err = kSrcSizeWrong; return 0xFFFFFFFFu; } nbSeq = ld_le16(src + ip) + kLongNbSeq; ip += 2; }
        else { if (ip >= srcSize) { *err = kSrcSizeWrong; return 0xFFFFFFFFu; } nbSeq = ((nbSeq - 0x80) << 8) + src[ip++]; }
    }
    it.nbSeq = nbSeq;
    if (ip + 1 > srcSize) { *err = kSrcSizeWrong; return 0xFFFFFFFFu; }
    uint32_t const modes = src[ip++];
    *err = kCorruptionDetected;
    for (int k = 0; k < 3; k++) {   // order: LL, OF, ML
        uint32_t const type = k == 0 ? (modes >> 6) : (k == 1 ? ((modes >> 4) & 3) : ((modes >> 2) & 3));
        int const kind = k == 0 ? 0 : (k == 1 ? 2 : 1);
        uint32_t const maxSym = kind == 0 ? kMaxLL : (kind == 1 ? kMaxML : kMaxOff);
        uint32_t const maxLog = kind == 2 ? kOffFSELog : kLLFSELog;
        uint32_t* const logPtr = kind == 0 ? &it.llLog : (kind == 1 ? &it.mlLog : &it.ofLog);
        switch (type) {
        case 1: {   // set_rle
            if (ip >= srcSize) return 0xFFFFFFFFu;
            uint32_t const sym = src[ip++];
            if (sym > maxSym) return 0xFFFFFFFFu;
            sc.tabAct[kind] = 1; sc.tabArg[kind] = sym;
            *logPtr = 0;
            break; }
        case 0: {   // set_basic: predefined table
            sc.tabAct[kind] = 2;
            *logPtr = kind == 2 ? kOFDefaultNormLog : kLLDefaultNormLog;
            break; }
        case 3:     // set_repeat
            if (!it.fseEntropy) return 0xFFFFFFFFu;
            break;
        default: {  // set_compressed
            uint32_t maxSV = maxSym, tableLog;
            uint32_t const hs = fse_read_ncount(sc.normT[kind], &maxSV, &tableLog, src + ip, srcSize - ip);
            if (hs == 0) return 0xFFFFFFFFu;
            if (tableLog > maxLog) return 0xFFFFFFFFu;
            sc.tabAct[kind] = 3; sc.tabArg[kind] = maxSV | (tableLog << 8);
            *logPtr = tableLog;
            ip += hs;
            break; }
        }
    }
    *err = 0;
    return ip;
}

__global__ void __launch_bounds__(kSetupWarps * 32) dec_setup_kernel(DecPass p)
{
    __shared__ SetupScratch scratch[kSetupWarps];
    __shared__ uint32_t s_huf[kSetupWarps][3];   // [0] fill table? [1] nbSymbols [2] tableLog
    __shared__ uint32_t s_dict[kSetupWarps];     // a frame starts under a dictionary with entropy tables: copy them into the item's slots
    uint32_t const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t const item = blockIdx.x * kSetupWarps + warp;
    if (item >= p.nItems) return;
    DecItem& it = p.items[item];
    SetupScratch& sc = scratch[warp];
    const uint8_t* const src = p.src + it.srcOff;
    if (lane == 0) {
        s_huf[warp][0] = 0; s_dict[warp] = 0;
        sc.tabAct[0] = sc.tabAct[1] = sc.tabAct[2] = 0;
        do {
            if (it.status != kStRunning) { it.blkType = kBlkNone; break; }
            uint32_t pos = it.srcPos; uint32_t const size = it.srcSize;
            if (!it.inFrame) {
                if (!advance_frames(it, src, &pos)) { it.srcPos = pos; it.blkType = kBlkNone; break; }
                // ---- ZSTD_decompressFrame entry, ZstdDecompress.cs:1062 ----
                uint32_t const rem = size - pos; uint32_t err = 0;
                if (rem < 9) err = kSrcSizeWrong;
                uint32_t fhsize = 0, fhd = 0;
                if (!err) { fhd = src[pos + 4]; fhsize = frame_header_size(fhd); if (rem < fhsize + 3) err = kSrcSizeWrong; }
                if (!err && ld_le32(src + pos) != kMagic) err = kPrefixUnknown;                 // ZSTD_getFrameHeader_advanced :462
                if (!err && (fhd & 0x08)) err = kFrameParameterUnsupported;
                if (!err) {
                    uint32_t q = pos + 5; uint32_t const single = (fhd >> 5) & 1, fcsId = fhd >> 6, didCode = fhd & 3;
                    if (!single) { uint32_t const wl = (src[q++] >> 3) + 10; if (wl > 31) err = kWindowTooLarge; }
                    uint32_t dictID = 0;
                    if (didCode == 1) { dictID = src[q]; q += 1; } else if (didCode == 2) { dictID = ld_le16(src + q); q += 2; } else if (didCode == 3) { dictID = ld_le32(src + q); q += 4; }
                    uint64_t fcs = 0; uint32_t has = 1;
                    if (fcsId == 0) { if (single) fcs = src[q]; else has = 0; }
                    else if (fcsId == 1) fcs = ld_le16(src + q) + 256; else if (fcsId == 2) fcs = ld_le32(src + q); else fcs = ld_le64(src + q);
                    uint32_t const haveID = p.dictInfo ? p.dictInfo[1] : 0u;
                    if (!err && dictID != 0 && dictID != haveID) err = kDictionaryWrong;         // ZSTD_decodeFrameHeader :849
                    // the dictionary content is materialised in front of the item's output, i.e. in front of its FIRST frame only
                    if (!err && p.dictInfo && it.outPos != 0) err = kFrameParameterUnsupported;  // DESIGN.md, deviations
                    it.fcs = fcs; it.hasFcs = has; it.checksumFlag = (fhd >> 2) & 1;
                }
                if (err) {
                    if (err == kPrefixUnknown && it.moreThan1Frame) err = kSrcSizeWrong;          // :1282
                    setup_fail(it, err); break;
                }
                pos += fhsize;
                it.rep[0] = 1; it.rep[1] = 4; it.rep[2] = 8; it.litEntropy = 0; it.fseEntropy = 0;   // ZSTD_decompressBegin :1933
                it.frameStart = it.outPos; it.inFrame = 1; it.prefix = 0;
                if (p.dictInfo) {                                                                // ZSTD_decompressBegin_usingDict :1954
                    const uint32_t* const di = p.dictInfo;
                    it.prefix = di[11];
                    if (di[2]) {
                        it.rep[0] = di[7]; it.rep[1] = di[8]; it.rep[2] = di[9]; it.litEntropy = 1; it.fseEntropy = 1;
                        it.hufLog = di[3]; it.llLog = di[4]; it.ofLog = di[5]; it.mlLog = di[6];
                        it.hufX2 = 1;                                                            // HUF_readDTableX2_wksp, ZstdDecompress.cs:1786
                        s_dict[warp] = 1;
                    }
                }
            }
            // ---- block header: ZSTD_getcBlockSize, ZstdDecompressBlock.cs:19 ----
            uint32_t rem = size - pos;
            if (rem < 3) { setup_fail(it, kSrcSizeWrong); break; }
            uint32_t const h = ld_le24(src + pos);
            uint32_t const type = (h >> 1) & 3, cs = h >> 3;
            it.lastBlock = h & 1;
            if (type == 3) { setup_fail(it, kCorruptionDetected); break; }
            uint32_t const cBlockSize = type == kBlkRle ? 1 : cs;
            pos += 3; rem -= 3;
            if (cBlockSize > rem) { setup_fail(it, kSrcSizeWrong); break; }
            uint32_t const capLeft = it.dstCap - it.outPos;
            it.blkType = type; it.blkSrcOff = pos; it.blkSize = cs; it.srcPos = pos + cBlockSize;
            it.nbSeq = 0; it.blockOut = 0; it.seqLitEnd = 0; it.deferErr = 0; it.blkLimit = it.dstCap;
            if (type != kBlkCompressed) {
                if (cs > capLeft) setup_fail(it, kDstSizeTooSmall);     // ZSTD_copyRawBlock :1004 / ZSTD_setRleBlock :1029
                else if (cs) { uint32_t const slot = atomicAdd(&p.counters[4], 1u); p.rawList[slot] = item; }
                break;
            }
            // ---- compressed block: ZSTD_decompressBlock_internal, ZstdDecompressBlock.cs:3090 ----
            if (cBlockSize >= kBlockSizeMax) { setup_fail(it, kSrcSizeWrong); break; }
            const uint8_t* const b = src + pos; uint32_t const bsize = cBlockSize;
            // literals header: ZSTD_decodeLiteralsBlock :88
            if (bsize < 3) { setup_fail(it, kCorruptionDetected); break; }
            uint32_t const litEncType = b[0] & 3, lhlCode = (b[0] >> 2) & 3;
            uint32_t const expectedWrite = capLeft < kBlockSizeMax ? capLeft : kBlockSizeMax;
            uint32_t litSectionSize = 0; uint32_t err = 0;
            if (litEncType == 2 || litEncType == 3) {
                if (litEncType == 3 && !it.litEntropy) { setup_fail(it, kDictionaryCorrupted); break; }
                if (bsize < 5) { setup_fail(it, kCorruptionDetected); break; }
                uint32_t lhSize, litSize, litCSize, single = 0;
                uint32_t const lhc = ld_le32(b);
                if (lhlCode <= 1) { single = !lhlCode; lhSize = 3; litSize = (lhc >> 4) & 0x3FF; litCSize = (lhc >> 14) & 0x3FF; }
                else if (lhlCode == 2) { lhSize = 4; litSize = (lhc >> 4) & 0x3FFF; litCSize = lhc >> 18; }
                else { lhSize = 5; litSize = (lhc >> 4) & 0x3FFFF; litCSize = (lhc >> 22) + ((uint32_t)b[4] << 10); }
                if (litSize > kBlockSizeMax) err = kCorruptionDetected;
                else if (litCSize + lhSize > bsize) err = kCorruptionDetected;
                else if (expectedWrite < litSize) err = kDstSizeTooSmall;
                if (err) { setup_fail(it, err); break; }
                uint32_t hSize = 0;
                const uint8_t* const hsrc = b + lhSize;
                if (litEncType == 2) {
                    // HUF_decompress4X_hufOnly_wksp (HufDecompress.cs:1793) / HUF_decompress1X1_DCtx_wksp (:1766)
                    if (!single && (litSize == 0 || litCSize == 0)) { setup_fail(it, kCorruptionDetected); break; }
                    uint32_t nbSym = 0, tlog = 0;
                    hSize = huf_read_stats(sc, &nbSym, &tlog, hsrc, litCSize);
                    if (hSize == 0 || hSize >= litCSize) { setup_fail(it, kCorruptionDetected); break; }
                    it.hufLog = tlog; s_huf[warp][0] = 1; s_huf[warp][1] = nbSym; s_huf[warp][2] = tlog;
                    it.hufX2 = single ? 0u : huf_select_decoder(litSize, litCSize);      // ZstdDecompressBlock.cs:212-216
                }
                uint32_t const cOff = lhSize + hSize, cLen = litCSize - hSize;
                it.litType = kLitHuf; it.litSize = litSize;
                if (single) {
                    it.nStreams = 1; it.streamOff[0] = pos + cOff; it.streamLen[0] = cLen;
                    if (cLen < 1) { setup_fail(it, kCorruptionDetected); break; }
                } else {
                    // HUF_decompress4X1_usingDTable_internal_body, HufDecompress.cs:342
                    if (cLen < 10) { setup_fail(it, kCorruptionDetected); break; }
                    uint32_t const l1 = ld_le16(b + cOff), l2 = ld_le16(b + cOff + 2), l3 = ld_le16(b + cOff + 4);
                    if (l1 + l2 + l3 + 6 > cLen) { setup_fail(it, kCorruptionDetected); break; }
                    uint32_t const l4 = cLen - (l1 + l2 + l3 + 6);
                    uint32_t const seg = (litSize + 3) / 4;
                    if (3 * seg > litSize) { setup_fail(it, kCorruptionDetected); break; }
                    if (l1 < 1 || l2 < 1 || l3 < 1 || l4 < 1) { setup_fail(it, kCorruptionDetected); break; }   // BIT_initDStream: srcSize < 1
                    it.nStreams = 4;
                    it.streamOff[0] = pos + cOff + 6; it.streamLen[0] = l1;
                    it.streamOff[1] = it.streamOff[0] + l1; it.streamLen[1] = l2;
                    it.streamOff[2] = it.streamOff[1] + l2; it.streamLen[2] = l3;
                    it.streamOff[3] = it.streamOff[2] + l3; it.streamLen[3] = l4;
                }
                it.litEntropy = 1;
                litSectionSize = litCSize + lhSize;
            } else {
                uint32_t lhSize, litSize;
                if (lhlCode == 1) { lhSize = 2; litSize = ld_le16(b) >> 4; }
                else if (lhlCode == 3) { lhSize = 3; litSize = ld_le24(b) >> 4; }
                else { lhSize = 1; litSize = b[0] >> 3; }
                if (litEncType == 0) {          // set_basic (raw literals)
                    if (expectedWrite < litSize) err = kDstSizeTooSmall;
                    else if (litSize + lhSize > bsize) err = kCorruptionDetected;
                    if (err) { setup_fail(it, err); break; }
                    it.litType = kLitRaw; it.litSize = litSize; it.litOff = pos + lhSize;
                    litSectionSize = lhSize + litSize;
                } else {                        // set_rle
                    if (lhlCode == 3 && bsize < 4) err = kCorruptionDetected;
                    else if (litSize > kBlockSizeMax) err = kCorruptionDetected;
                    else if (expectedWrite < litSize) err = kDstSizeTooSmall;
                    if (err) { setup_fail(it, err); break; }
                    it.litType = kLitRle; it.litSize = litSize; it.litOff = pos + lhSize;
                    litSectionSize = lhSize + 1;
                }
            }
            // literal buffer placement (ZSTD_allocateLiteralsBuffer :44-73): with room behind the block the reference stores the literals in
            // dst at +128 KiB + 32 and the sequences may only write up to there; raw literals with 32 readable bytes behind them stay in src
            {
                bool const direct = it.litType == kLitRaw && (litSectionSize + 32 <= bsize);
                if (!direct && capLeft > kBlockSizeMax + 32 + it.litSize + 32) it.blkLimit = it.outPos + kBlockSizeMax + 32;
            }
            // An error from here on sits BEHIND the literal section: the reference decodes Huffman literals first (corruption_detected if
            // they are damaged) and only then parses the sequence section, so with Huffman literals the verdict is left to dec_huf
            bool const hufFirst = it.litType == kLitHuf;
            uint32_t late = 0;
            // sequences header
            uint32_t serr = 0;
            uint32_t const shs = setup_seq_headers(it, p, item, sc, b + litSectionSize, bsize - litSectionSize, &serr);
            if (shs == 0xFFFFFFFFu) late = serr;
            else {
                it.seqOff = pos + litSectionSize + shs; it.seqLen = bsize - litSectionSize - shs;
                if (it.nbSeq) {
                    if (it.nbSeq > kSeqCap) late = kCorruptionDetected;
                    else if (it.seqLen < 1) late = kCorruptionDetected;                          // BIT_initDStream error -> corruption (:2697)
                    else it.fseEntropy = 1;
                } else if (it.litSize > it.blkLimit - it.outPos) late = kDstSizeTooSmall;         // last literals copy :2748
            }
            if (late && !hufFirst) { setup_fail(it, late); break; }
            if (late) { it.deferErr = late; it.nbSeq = 0; sc.tabAct[0] = sc.tabAct[1] = sc.tabAct[2] = 0; }
            else if (it.nbSeq) { uint32_t const slot = atomicAdd(&p.counters[1], 1u); p.seqList[slot] = item; }
            if (hufFirst) { uint32_t const slot = atomicAdd(&p.counters[0], 1u); p.hufList[slot] = item; }
        } while (0);
    }
    __syncwarp();
    if (s_dict[warp]) {     // the dictionary's tables become the frame's "previous block" tables (repeat modes refer to them)
        const uint4* const sh = (const uint4*)p.dictHuf; uint4* const dh = (uint4*)(p.hufTable + (size_t)item * kHufTableEntries);
        for (uint32_t u = lane; u < kHufTableEntries * 2 / 16; u += 32) dh[u] = sh[u];
        const uint4* const sf = (const uint4*)p.dictFse; uint4* const df = (uint4*)(p.fseTable + (size_t)item * kFseTableEntries);
        for (uint32_t u = lane; u < kFseTableEntries * 4 / 16; u += 32) df[u] = sf[u];
    }
    __syncwarp();
    // sequence tables: built by the whole warp (ZSTD_buildSeqTable :1746)
    if (it.status == kStRunning && it.blkType == kBlkCompressed) {
        uint32_t* const tbl = p.fseTable + (size_t)item * kFseTableEntries;
        for (int kind = 0; kind < 3; kind++) {
            uint32_t const act = sc.tabAct[kind], arg = sc.tabArg[kind];
            uint32_t const off = kind == 0 ? kFseLLOff : (kind == 1 ? kFseMLOff : kFseOFOff);
            if (act == 1) {
                if (lane == 0) { uint32_t const add = kind == 0 ? c_LL_bits[arg] : (kind == 1 ? c_ML_bits[arg] : arg); tbl[off] = fse_pack(0, add, arg, 0); }
            } else if (act == 2) {
                uint32_t const n = kind == 2 ? (1u << kOFDefaultNormLog) : (1u << kLLDefaultNormLog);
                for (uint32_t u = lane; u < n; u += 32) tbl[off + u] = p.defaultFse[off + u];
            } else if (act == 3) {
                build_seq_table_warp(tbl + off, sc, sc.normT[kind], arg & 0xFF, arg >> 8, kind, lane);
            }
        }
    }
    __syncwarp();
    // warp-cooperative fill of the single-symbol Huffman table at its native tableLog: every symbol of weight w
    // owns (1<<w)>>1 consecutive cells, nbBits = tableLog+1-w (HUF_readDTableX1_wksp_bmi2, HufDecompress.cs:131-252)
    if (s_huf[warp][0] && it.status == kStRunning) {
        uint32_t const nbSym = s_huf[warp][1], tlog = s_huf[warp][2];
        uint16_t* const tab = p.hufTable + (size_t)item * kHufTableEntries;
        for (uint32_t s = 0; s < nbSym; s++) {
            uint32_t const w = sc.weights[s];
            if (!w) continue;
            uint32_t const len = (1u << w) >> 1, st = sc.start[s];
            uint16_t const e = (uint16_t)((s << 8) | (tlog + 1 - w));
            for (uint32_t u = lane; u < len; u += 32) tab[st + u] = e;
        }
    }
}

// =====================================================================================================
//  Backward bit reader (semantics of BIT_DStream_t, Bitstream.cs:172-425, for the hot loops).
//  A stream is consumed from its last byte downwards.  It is staged through a private shared-memory ring of 128 bytes
//  that mirrors the low bits of the global address (ring byte = global byte mod 128) and is filled with cp.async
//  several chunks ahead, so the decode loops never wait on HBM/L2.  The reader state is one integer: G, the
//  bit address (relative to the ring-aligned base) just above the next unread bit.  A read is two shared loads and
//  one funnel shift, with no branch: that keeps every lane of a warp on the same instruction stream, which is
//  what bounds these latency-limited kernels (profiles/r01_notes.md).
// =====================================================================================================
// Shared-memory loads by 32-bit shared address.  `volatile` + "memory": the compiler must see them as memory reads, ordered after the
// stores / cp.async waits that produce the data.  (Round 1 had plain `asm`: formally free to move across `cp.async.wait_group` and
// `__syncwarp()`; it happened to work until the assertion build of round 2 perturbed the schedule and the last sequences of some frames
// came out wrong, profiles/r02_notes.md.)
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t saddr) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr) : "memory"); return v; }

// top `nb` bits (0..31) of a left-justified word
__device__ __forceinline__ uint32_t top_bits(uint32_t x, uint32_t nb) { return __funnelshift_l(x, 0u, nb); }   // one SHF: high word of (0:x) << nb

// UNIFORM refill: ring of 8 chunks of 16 bytes, topped up by at most NF chunks per step, by every lane at the same
// instruction (predicated cp.async, no branch), D chunks ahead of the read position; the wait only covers groups
// committed more than W steps ago, so it never stalls in practice.
// Why: the first reader let every lane switch (32-byte) chunks on its own schedule.  A warp of independent streams then took
// the switch branch in most steps, and its `wait_group 0` waited -- through the warp-wide scoreboard -- for the cp.async that
// ANOTHER lane had issued one step earlier: 22 % of all stall samples of the sequence decoder sat on that wait
// (profiles/r01_notes.md; Silesia-mix decode 8.6 -> 7.0 ms with this reader).
// Contract: call step() once per step, all lanes together; a step consumes less than 16 bytes (one chunk per step keeps up)
// and reads only chunks cur .. cur-R of its starting position.  Chunk cur-R was requested when the reader was in chunk
// cur-R+D, i.e. more than (D-R-1)*16 bytes ago: W must be below that many bytes / the bytes a step can consume.
// The slot of chunk c-D is the slot of chunk c-D+8, which must lie above cur: D <= 6.
template <int D, int W>
struct BitRingU {
    static constexpr uint32_t CH = 16, NCH = 8, RB = CH * NCH, MASK = RB - 1;
    uint32_t sbase; const uint8_t* gbase; int32_t fetched, cLow; uint32_t gZero;
    __device__ __forceinline__ void fetch(int32_t c, bool on) const {
        uint32_t const sa = sbase + (((uint32_t)c * CH) & MASK);
        const uint8_t* const g = gbase + (ptrdiff_t)c * (ptrdiff_t)CH;
        uint32_t const pr = (on && c >= cLow) ? 1u : 0u;
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p cp.async.cg.shared.global [%0], [%1], 16; }" ::"r"(sa), "l"(g), "r"(pr) : "memory");
    }
    // Returns G of the stream start (top), or 0 when the last byte is zero (no end mark: corruption).  The caller follows up
    // with settle() once all lanes are back together.
    __device__ __forceinline__ uint32_t init(uint32_t ringShared, const uint8_t* first, uint32_t len) {
        sbase = ringShared;
        uintptr_t const a0 = (uintptr_t)first;
        // The base lies one ring length BELOW the aligned address, so that G >= 1024 for every stream: a stream without data bits (one
        // byte 0x01: all three sequence tables RLE) that starts on a 128-byte boundary would otherwise have G == 0, the failure value
        // (found by the soak, seed 777002: 1 valid frame of 40000 rejected).  Ring slots and peeks only use G modulo 1024 bits.
        gbase = (const uint8_t*)(a0 & ~(uintptr_t)MASK) - RB;
        uint32_t const rel = (uint32_t)(a0 & MASK) + RB;
        gZero = rel * 8;
        cLow = (int32_t)(rel / CH);
        fetched = cLow;
        uint32_t const lastByte = first[len - 1];
        if (lastByte == 0) return 0;
        uint32_t const G = gZero + (len - 1) * 8 + highbit32(lastByte);
        int32_t const cur = (int32_t)((G - 1) >> 3) / (int32_t)CH;
#pragma unroll
        for (int j = 0; j <= D; j++) fetch(cur - j, true);
        fetched = cur - D;
        return G;
    }
    __device__ __forceinline__ void idle() { sbase = 0; gbase = nullptr; fetched = 0; cLow = 0x7FFFFFFF; gZero = 0; }   // a lane without a stream
    static __device__ __forceinline__ void settle() { asm volatile("cp.async.commit_group;" ::: "memory"); asm volatile("cp.async.wait_group 0;" ::: "memory"); }
    template <int NF = 1>      // chunks a step may consume (NF * 16 bytes at most)
    __device__ __forceinline__ void step(uint32_t G, bool on) {
        int32_t const cur = (int32_t)(G - 1) >> 7;
#pragma unroll
        for (int f = 0; f < NF; f++) {
            bool const need = on && fetched > cur - D;
            fetched -= need ? 1 : 0;
            fetch(fetched, need);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(W) : "memory");
    }
    __device__ __forceinline__ uint32_t peek32(uint32_t G) const {
        uint32_t const o = ((G - 1) >> 3) & (MASK & ~3u);
        uint32_t const hi = lds32(sbase + o), lo = lds32(sbase + ((o - 4) & MASK));
        return __funnelshift_l(lo, hi, 0u - G);
    }
    __device__ __forceinline__ void peek64(uint32_t G, uint32_t& x0, uint32_t& x1) const {
        uint32_t const o = ((G - 1) >> 3) & (MASK & ~3u);
        uint32_t const w0 = lds32(sbase + o), w1 = lds32(sbase + ((o - 4) & MASK)), w2 = lds32(sbase + ((o - 8) & MASK));
        x0 = __funnelshift_l(w1, w0, 0u - G); x1 = __funnelshift_l(w2, w1, 0u - G);
    }
};

// =====================================================================================================
//  Huffman literal decoding: one lane per stream, 14 items (56 streams) per CTA.
//  HUF_decompress4X1_usingDTable_internal_body / HUF_decodeStreamX1 (HufDecompress.cs:342, :264)
//  The single-symbol table is split for shared memory: symbol bytes u8[2^log] and code lengths u8[2^(log-1)] -- cells
//  2j and 2j+1 always hold the same length, because every rank starts at an even cell (rankStats[1] is even,
//  EntropyCommon.cs:398).  3 KB per item for tableLog 11 instead of 4 KB, so that 56 items fit one SM and a batch of
//  8192 items is resident in ONE wave (any item needs a full stream time: 1.15 waves would cost 2x).
//  tableLog 12 (legal, never produced by zstd's encoder) is decoded from the table in HBM.
// =====================================================================================================
constexpr int kHufItemsPerCta = 14;
constexpr int kHufThreads = kHufItemsPerCta * 4;
constexpr uint32_t kHufSmemLog = 11;
constexpr uint32_t kHufSymBytes = 1u << kHufSmemLog, kHufLenBytes = 1u << (kHufSmemLog - 1), kHufItemBytes = kHufSymBytes + kHufLenBytes;
constexpr uint32_t kHufChunk = 32;              // 4 x 32 B of ring per stream (8 symbols consume <= 12 bytes, a refill reads 12 more)
constexpr uint32_t kHufSmemBytes = kHufThreads * 4 * kHufChunk + kHufItemsPerCta * kHufItemBytes;
__device__ __forceinline__ uint32_t lds8u(uint32_t saddr) { uint16_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=h"(v) : "r"(saddr) : "memory"); return v; }

__device__ __forceinline__ uint32_t lit_segment_stride(uint32_t litSize) { uint32_t const seg = (litSize + 3) / 4; return (seg + 15) & ~15u; }

// ---- Cold path: the reference's verdict on a DAMAGED literal section when it runs the double-symbol decoder ----
// On valid streams HUF_decompress4X2 yields the bytes of the single-symbol decoder, so the hot loop below never builds the
// double-symbol table.  On damaged streams the verdict can differ: HUF_decodeLastSymbolX2 (HufDecompress.cs:1022) clamps
// bitsConsumed when a two-symbol cell is used for the last symbol (a stream whose last code runs past its start is ACCEPTED), and
// the interleaved main loop of HUF_decompress4X2_usingDTable_internal_body (:1148) stops on stream 4 only, so streams 1-3 can
// overshoot their segment (:1322-1335).  When the hot loop finds a stream that is not consumed exactly and the table is of the
// double-symbol type, one thread replays the reference's control flow here (BIT_DStream_t of Bitstream.cs:172-425, cells derived
// from the single-symbol table: a cell holds two symbols when the second one fits the remaining dtLog - l1 bits, :789-886).
struct X2Bits {
    const uint8_t* s; uint32_t len; int32_t ptr; uint64_t c; uint32_t consumed;
    __device__ uint64_t rd(int32_t at) const { uint64_t v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | ((uint32_t)(at + i) < len ? s[at + i] : 0u); return v; }
    __device__ void init(const uint8_t* b, uint32_t n) {          // n >= 1 and b[n-1] != 0 were checked by the hot loop's reader
        s = b; len = n;
        uint32_t const hb = highbit32(b[n - 1]);
        if (n >= 8) { ptr = (int32_t)n - 8; c = rd(ptr); consumed = 8 - hb; }
        else { ptr = 0; c = rd(0); consumed = 8 - hb + (8 - n) * 8; }
    }
    __device__ uint32_t look(uint32_t nb) const { return (uint32_t)((c << (consumed & 63)) >> (64 - nb)); }
    __device__ int reloadFast() { if (ptr < 8) return 3; ptr -= (int32_t)(consumed >> 3); consumed &= 7; c = rd(ptr); return 0; }
    __device__ int reload() {                                     // 0 unfinished, 1 endOfBuffer, 2 completed, 3 overflow
        if (consumed > 64) return 3;
        if (ptr >= 8) return reloadFast();
        if (ptr == 0) return consumed < 64 ? 1 : 2;
        uint32_t nb = consumed >> 3; int res = 0;
        if ((int32_t)nb > ptr) { nb = (uint32_t)ptr; res = 1; }
        ptr -= (int32_t)nb; consumed -= nb * 8; c = rd(ptr);
        return res;
    }
    __device__ bool finished() const { return ptr == 0 && consumed == 64; }
};
struct X2Table {
    const uint16_t* tab; uint32_t nativeLog, dtLog, minBits;
    // cell of index val: returns length (1 or 2); sym = s1 | s2 << 8, nb = bits consumed
    __device__ uint32_t cell(uint32_t val, uint32_t& sym, uint32_t& nb) const {
        uint32_t const e1 = tab[val >> (dtLog - nativeLog)], l1 = e1 & 0xFF;
        sym = e1 >> 8; nb = l1;
        if (dtLog - l1 >= minBits) {
            uint32_t const e2 = tab[((val << l1) & ((1u << dtLog) - 1u)) >> (dtLog - nativeLog)], l2 = e2 & 0xFF;
            if (l2 <= dtLog - l1) { sym |= (e2 >> 8) << 8; nb = l1 + l2; return 2; }
        }
        return 1;
    }
};
struct X2Out { uint8_t* p; uint32_t pos, end; };                   // pos may overshoot end inside the main loop; bytes beyond end are dropped
__device__ void x2_symbol(X2Out& o, X2Bits& b, const X2Table& t)   // HUF_decodeSymbolX2 (:1012)
{
    uint32_t sym, nb; uint32_t const n = t.cell(b.look(t.dtLog), sym, nb);
    if (o.pos < o.end) o.p[o.pos] = (uint8_t)sym;
    if (n == 2 && o.pos + 1 < o.end) o.p[o.pos + 1] = (uint8_t)(sym >> 8);
    b.consumed += nb; o.pos += n;
}
__device__ void x2_stream(X2Out& o, X2Bits& b, const X2Table& t)   // HUF_decodeStreamX2 (:1047)
{
    if (o.end - o.pos >= 8) {
        if (t.dtLog <= 11) { while (b.reload() == 0 && o.pos + 9 < o.end) for (int k = 0; k < 5; k++) x2_symbol(o, b, t); }
        else { while (b.reload() == 0 && o.pos + 7 < o.end) for (int k = 0; k < 4; k++) x2_symbol(o, b, t); }
    } else b.reload();
    if (o.end - o.pos >= 2) {
        while (b.reload() == 0 && o.pos + 2 <= o.end) x2_symbol(o, b, t);
        while (o.pos + 2 <= o.end) x2_symbol(o, b, t);
    }
    if (o.pos < o.end) {                                           // HUF_decodeLastSymbolX2 (:1022)
        uint32_t sym, nb; uint32_t const n = t.cell(b.look(t.dtLog), sym, nb);
        o.p[o.pos++] = (uint8_t)sym;
        if (n == 1) b.consumed += nb;
        else if (b.consumed < 64) { b.consumed += nb; if (b.consumed > 64) b.consumed = 64; }
    }
}
// true = the reference accepts the literal section (and the literal buffer now holds what it decodes)
__device__ __noinline__ bool huf_x2_replay(const DecPass& p, uint32_t item, const DecItem& it)
{
    X2Table t; t.tab = p.hufTable + (size_t)item * kHufTableEntries; t.nativeLog = it.hufLog; t.dtLog = it.hufLog <= 11 ? 11u : 12u;
    t.minBits = t.tab[(1u << t.nativeLog) - 1u] & 0xFF;           // the last cells belong to the heaviest symbol = the shortest code
    const uint8_t* const src = p.src + it.srcOff;
    uint8_t* const lit = p.litBuf + (size_t)item * kLitStride;
    uint32_t const litSize = it.litSize;
    X2Bits b[4]; X2Out o[4];
    if (it.nStreams == 1) {
        b[0].init(src + it.streamOff[0], it.streamLen[0]);
        o[0].p = lit; o[0].pos = 0; o[0].end = litSize;
        x2_stream(o[0], b[0], t);
        return b[0].finished();
    }
    uint32_t const seg = (litSize + 3) / 4, stride = lit_segment_stride(litSize);
    for (int j = 0; j < 4; j++) {
        b[j].init(src + it.streamOff[j], it.streamLen[j]);
        o[j].p = lit + j * stride; o[j].pos = 0; o[j].end = j < 3 ? seg : litSize - 3 * seg;
    }
    if (o[3].end >= 8) {
        uint32_t endSignal = 1;
        while (endSignal && o[3].pos + 7 < o[3].end) {             // op4 < olimit
            for (int k = 0; k < 4; k++) x2_symbol(o[0], b[0], t);
            for (int k = 0; k < 4; k++) x2_symbol(o[1], b[1], t);
            endSignal &= b[0].reloadFast() == 0; endSignal &= b[1].reloadFast() == 0;
            for (int k = 0; k < 4; k++) x2_symbol(o[2], b[2], t);
            for (int k = 0; k < 4; k++) x2_symbol(o[3], b[3], t);
            endSignal &= b[2].reloadFast() == 0; endSignal &= b[3].reloadFast() == 0;
        }
    }
    if (o[0].pos > o[0].end || o[1].pos > o[1].end || o[2].pos > o[2].end) return false;      // op1 > opStart2 ... (:1322-1335)
    for (int j = 0; j < 4; j++) x2_stream(o[j], b[j], t);
    return b[0].finished() && b[1].finished() && b[2].finished() && b[3].finished();
}

__global__ void __launch_bounds__(kHufThreads) dec_huf_kernel(DecPass p)
{
    extern __shared__ __align__(256) uint8_t s_huf_raw[];    // [kHufThreads] rings of 128 B, then per item: u8 sym[2048] | u8 len[1024]
    uint8_t* const s_tab = s_huf_raw + kHufThreads * 4 * kHufChunk;
    uint32_t const nWork = p.counters[0];
    uint32_t const first = blockIdx.x * kHufItemsPerCta;
    if (first >= nWork) return;
    uint32_t const nHere = min((uint32_t)kHufItemsPerCta, nWork - first);
    for (uint32_t k = 0; k < nHere; k++) {
        uint32_t const item = p.hufList[first + k];
        uint32_t const log = p.items[item].hufLog;
        if (log > kHufSmemLog) continue;
        const uint32_t* const g = (const uint32_t*)(p.hufTable + (size_t)item * kHufTableEntries);      // two u16 cells {byte<<8 | nbBits} per word
        uint16_t* const sym = (uint16_t*)(s_tab + k * kHufItemBytes);
        uint8_t* const len = s_tab + k * kHufItemBytes + kHufSymBytes;
        uint32_t const pairs = max(1u, (1u << log) >> 1);
        for (uint32_t u = threadIdx.x; u < pairs; u += kHufThreads) {
            uint32_t const w = g[u];
            sym[u] = (uint16_t)(((w >> 8) & 0xFF) | ((w >> 24) << 8));
            len[u] = (uint8_t)(w & 0xFF);
        }
    }
    __shared__ uint32_t s_inexact[kHufItemsPerCta];
    if (threadIdx.x < kHufItemsPerCta) s_inexact[threadIdx.x] = 0;
    __syncthreads();
    uint32_t const slot = threadIdx.x >> 2, stream = threadIdx.x & 3;
    uint32_t const item = p.hufList[first + min(slot, nHere - 1)];
    DecItem& it = p.items[item];
    uint32_t const nStreams = it.nStreams;
    bool const active = slot < nHere && it.status == kStRunning && stream < nStreams;
    if (active) {
    uint32_t const litSize = it.litSize, log = it.hufLog;
    uint32_t const seg = (litSize + 3) / 4;
    uint32_t count, outOff;
    if (nStreams == 1) { count = litSize; outOff = 0; }
    else { count = stream < 3 ? seg : litSize - 3 * seg; outOff = stream * lit_segment_stride(litSize); }
    uint8_t* const out = p.litBuf + (size_t)item * kLitStride + outOff;
    const uint8_t* const src = p.src + it.srcOff;
    // one refill (up to two chunks) per 16 symbols (<= 22 bytes) in the main loop, per symbol in the tails; the end of a main
    // step reads down to chunk cur-2, which was requested more than 48 bytes = more than 2 steps ago
    BitRingU<6, 2> br;
    uint32_t G = br.init((uint32_t)__cvta_generic_to_shared(s_huf_raw) + threadIdx.x * (4 * kHufChunk), src + it.streamOff[stream], it.streamLen[stream]);
    br.settle();
    bool ok = G != 0;
    if (ok) {
        uint32_t const sh = 32 - log;
        uint32_t const symS = (uint32_t)__cvta_generic_to_shared(s_tab + slot * kHufItemBytes), lenS = symS + kHufSymBytes;
        int32_t const gz = (int32_t)br.gZero;
        uint32_t i = 0;
        if (log <= kHufSmemLog) {
            // 16 symbols -> one 16-byte store (out is 16-byte aligned by construction); 4 symbols (<= 44 bits) per window refill
            for (; i + 16 <= count && (int32_t)G >= gz; i += 16) {
                uint32_t v[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (q == 0) br.step<2>(G, true);
                    uint32_t x0, x1; br.peek64(G, x0, x1);
                    uint32_t acc = 0, used = 0;
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        uint32_t const idx = x0 >> sh;
                        uint32_t const nb = lds8u(lenS + (idx >> 1));
                        uint32_t const sy = lds8u(symS + idx);
                        x0 = __funnelshift_l(x1, x0, nb); x1 <<= nb; used += nb;
                        acc |= sy << (8 * r);
                    }
                    G -= used;
                    v[q] = acc;
                }
                ZB_ASSERTK(0, outOff + i + 16 <= kLitStride);
                *(uint4*)(out + i) = make_uint4(v[0], v[1], v[2], v[3]);
            }
            for (; i < count && (int32_t)G >= gz; i++) {
                br.step(G, true);
                uint32_t const idx = br.peek32(G) >> sh;
                G -= lds8u(lenS + (idx >> 1));
                out[i] = (uint8_t)lds8u(symS + idx);
            }
        } else {
            const uint16_t* const gtab = p.hufTable + (size_t)item * kHufTableEntries;
            for (; i < count && (int32_t)G >= gz; i++) {
                br.step(G, true);
                uint32_t const e = __ldg(gtab + (br.peek32(G) >> sh));
                G -= e & 0xFF;
                out[i] = (uint8_t)(e >> 8);
            }
        }
        ok = ((int32_t)G == gz) && (i == count);      // BIT_endOfDStream: the stream must be consumed exactly (:526-533)
    }
    if (!ok) s_inexact[slot] = G != 0 ? 1u : 2u;      // 2: BIT_initDStream itself fails (no end mark), whatever the decoder
    }
    __syncthreads();
    if (stream == 0 && slot < nHere && it.status == kStRunning && s_inexact[slot]) {
        // single-symbol decoder (HufDecompress.cs:526-533): corruption.  Double-symbol decoder: replay its control flow.
        bool const accept = s_inexact[slot] == 1 && it.hufX2 && huf_x2_replay(p, item, it);
        if (!accept) { it.errCode = kCorruptionDetected; it.status = kStError; }
    }
    if (stream == 0 && slot < nHere && it.status == kStRunning && it.deferErr) { it.errCode = it.deferErr; it.status = kStError; }    // the literals were sound
}

// =====================================================================================================
//  Sequence decoding: one lane per item; LL/ML/OF tables staged in shared memory (5 KB per item).
//  ZSTD_decompressSequences_body / ZSTD_decodeSequence (ZstdDecompressBlock.cs:2668, :2360) incl. the checks of
//  ZSTD_execSequenceEnd (:2083-2103), so that the exec kernel can copy without re-validating.
//  One step = three table loads, three bit-group reads (offset extra | matchLength+litLength extra | the three
//  state updates: each group is <= 32 bits), all branch-free.
// =====================================================================================================
constexpr int kSeqItemsPerCta = 14;            // 14 x (3.75 KB tables + 128 B ring) = 54.25 KB -> 4 CTAs = 56 items per SM: 8192 items are resident in ONE wave
constexpr uint32_t kSeqChunk = 32;             // 4 x 32 B of ring per item (a step consumes <= 12 bytes and reads <= 8 below)
// shared-memory table entry = 24 bits: u16 {nbBits 4 | addBits 5 | symbol 6 | next-state base bit 8} + u8 {next-state base bits 0..7}
constexpr uint32_t kSeqTab16Bytes = kFseTableEntries * 2, kSeqTab8Bytes = kFseTableEntries;
constexpr uint32_t kSeqItemBytes = kSeqTab16Bytes + kSeqTab8Bytes;
constexpr uint32_t kSeqFlush = 8;              // sequences per record flush
constexpr uint32_t kSeqSmemBytes = kSeqItemsPerCta * (kSeqItemBytes + 4 * kSeqChunk) + kSeqFlush * kSeqItemsPerCta * 8;
__device__ __forceinline__ uint32_t lds8(uint32_t saddr) { uint16_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=h"(v) : "r"(saddr) : "memory"); return v; }

// BIT_readBits / BIT_readBitsFast exactly as the reference behaves near, at and beyond the START of a backward bit stream
// (Bitstream.cs:189-264, 293-340, 381-424).  R = unread bits before the read.  While R >= 0 a read that runs past the start is filled
// with zero bits (the container is shifted left).  Once bitsConsumed has passed the 64-bit container (R < 0), BIT_reloadDStream
// reports `overflow` and changes nothing, and the shift count wraps (bitsConsumed & 63): the read returns bits of the stream's first
// eight bytes again.  The sequence decoder only checks the stream AFTER the last sequence (status >= completed,
// ZstdDecompressBlock.cs:2730), so a corrupted stream that runs dry early is still decoded and executed if its lengths and offsets
// happen to be valid; a valid stream gets here with its last few sequences (dec_seq_kernel's tail loop).
// The first 16 bytes of the stream held in registers (bytes past the stream's end read as zero): every read of the tail loop
// (R < kSeqTailBits = 96 unread bits) and of a stream shorter than its three initial states comes out of these two words.
struct NearStart {
    uint64_t lo, hi;                                                                               // stream bytes 0..7, 8..15 (little endian)
    __device__ __forceinline__ void load(const uint8_t* s, uint32_t len)
    {
        lo = hi = 0;
#pragma unroll
        for (int j = 7; j >= 0; j--) { lo = (lo << 8) | ((uint32_t)j < len ? s[j] : 0u); hi = (hi << 8) | ((uint32_t)(j + 8) < len ? s[j + 8] : 0u); }
    }
    // BIT_readBitsFast (also what BIT_readBits returns inside the stream): R <= 127
    __device__ __forceinline__ uint32_t bits(int32_t R, uint32_t nb) const
    {
        if (nb == 0) return 0;
        uint32_t const mask = (1u << nb) - 1u;                                                     // nb <= 31
        if (R <= 0) return (uint32_t)((lo << ((uint32_t)(-R) & 63u)) >> (64 - nb));                // R == 0: bitsConsumed == 64, shift count 0
        if (R < (int32_t)nb) return ((uint32_t)lo << (nb - (uint32_t)R)) & mask;                   // runs past the start: zero filled
        uint32_t const sft = (uint32_t)R - nb;                                                     // 0..127
        uint64_t const v = sft >= 64 ? hi >> (sft - 64) : (sft ? (lo >> sft) | (hi << (64 - sft)) : lo);
        return (uint32_t)v & mask;
    }
    // BIT_readBits (BIT_getMiddleBits): a read past the start takes bits from the TOP of the first eight bytes
    __device__ __forceinline__ uint32_t state_bits(int32_t R, uint32_t nb) const
    {
        if (nb == 0) return 0;
        if (R >= (int32_t)nb) return bits(R, nb);
        return (uint32_t)(lo >> ((uint32_t)(R - (int32_t)nb) & 63u)) & ((1u << nb) - 1u);
    }
};
constexpr int32_t kSeqTailBits = 96;      // a sequence reads at most 31 + 16 + 16 extra bits and 9 + 9 + 8 state bits

__global__ void __launch_bounds__(32) dec_seq_kernel(DecPass p)
{
    extern __shared__ __align__(256) uint8_t s_seq_raw[];   // [14] rings of 128 B | per item: u16[1280] | u8[1280] tables | record staging [8][14] x 8 B
    uint8_t* const s_tabs = s_seq_raw + kSeqItemsPerCta * 4 * kSeqChunk;
    uint2* const s_stage = (uint2*)(s_tabs + kSeqItemsPerCta * kSeqItemBytes);
    __shared__ uint32_t s_llBase[64], s_mlBase[64];
    uint32_t const nWork = p.counters[1];
    uint32_t const first = blockIdx.x * kSeqItemsPerCta;
    if (first >= nWork) return;
    uint32_t const nHere = min((uint32_t)kSeqItemsPerCta, nWork - first);
    uint32_t const lane = threadIdx.x;
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    for (uint32_t k = 0; k < nHere; k++) {
        uint32_t const item = p.seqList[first + k];
        const uint4* g = (const uint4*)(p.fseTable + (size_t)item * kFseTableEntries);
        uint2* const s16 = (uint2*)(s_tabs + k * kSeqItemBytes);
        uint32_t* const s8 = (uint32_t*)(s_tabs + k * kSeqItemBytes + kSeqTab16Bytes);
        for (uint32_t u = lane; u < kFseTableEntries / 4; u += 32) {       // 4 compact u32 entries -> 4 x u16 + 4 x u8
            uint4 const e = g[u];
            auto h = [](uint32_t x) { return (x & 0x7FFFu) | (((x >> 24) & 1u) << 15); };   // nb | add | sym | next bit 8
            s16[u] = make_uint2(h(e.x) | (h(e.y) << 16), h(e.z) | (h(e.w) << 16));
            s8[u] = ((e.x >> 16) & 0xFF) | (((e.y >> 16) & 0xFF) << 8) | (((e.z >> 16) & 0xFF) << 16) | (((e.w >> 16) & 0xFF) << 24);
        }
    }
    for (uint32_t u = lane; u < 64; u += 32) { s_llBase[u] = u < 36 ? c_LL_base[u] : 0u; s_mlBase[u] = u < 53 ? c_ML_base[u] : 0u; }
    __syncwarp();
    // Lanes 0..13 own one item each; all 32 lanes stay for the record flush (below).
    bool const mine = lane < nHere;
    uint32_t const item = p.seqList[first + (mine ? lane : 0)];
    DecItem& it = p.items[item];
    bool const live = mine && it.status == kStRunning;
    // state registers hold the entry INDEX (table offset included); the u16 part sits at t16 + 2*idx, the u8 part at t8 + idx
    uint32_t const slotL = mine ? lane : 0u;
    uint32_t const t16 = (uint32_t)__cvta_generic_to_shared(s_tabs + slotL * kSeqItemBytes), t8 = t16 + kSeqTab16Bytes;
    uint32_t const llBaseS = (uint32_t)__cvta_generic_to_shared(s_llBase), mlBaseS = (uint32_t)__cvta_generic_to_shared(s_mlBase);
    uint32_t const nbSeq = live ? it.nbSeq : 0u;
    // one refill (up to two chunks) per two sequences: they consume <= 178 bits and read down to chunk cur-2, which was requested
    // more than 48 bytes = more than 2 refills ago
    BitRingU<6, 2> br;
    br.idle();
    uint32_t err = 0, G = 1;
    if (live) {
        G = br.init((uint32_t)__cvta_generic_to_shared(s_seq_raw) + lane * (4 * kSeqChunk), p.src + it.srcOff + it.seqOff, it.seqLen);
        if (G == 0) { err = kCorruptionDetected; G = 1; }
    }
    br.settle();
    int32_t const gz = (int32_t)br.gZero;
    uint32_t rep0 = it.rep[0], rep1 = it.rep[1], rep2 = it.rep[2];
    uint32_t outPos = it.outPos; uint32_t const outPos0 = outPos, dstCap = it.blkLimit;      // the block's write limit (<= the item's capacity)
    uint32_t const frameStart = it.frameStart - it.prefix;      // virtualStart: the dictionary content sits right in front of the frame (wraps below 0: unsigned differences stay right)
    uint32_t litPos = 0; uint32_t const litSize = it.litSize;
    uint32_t aL = kFseLLOff, aO = kFseOFOff, aM = kFseMLOff;
    if (live && !err) {
        // ZSTD_initFseState x3 in the order LL, OF, ML (:2702-2704)
        uint32_t const llLog = it.llLog, ofLog = it.ofLog, mlLog = it.mlLog;
        uint32_t x = br.peek32(G);
        aL = kFseLLOff + top_bits(x, llLog); x <<= llLog;
        aO = kFseOFOff + top_bits(x, ofLog); x <<= ofLog;
        aM = kFseMLOff + top_bits(x, mlLog);
        if ((int32_t)(G - (llLog + ofLog + mlLog)) < gz) {           // fewer bits than the three initial states need: the reference reads on
            int32_t const R = (int32_t)G - gz;
            NearStart ns; ns.load(p.src + it.srcOff + it.seqOff, it.seqLen);
            aL = kFseLLOff + ns.state_bits(R, llLog);
            aO = kFseOFOff + ns.state_bits(R - (int32_t)llLog, ofLog);
            aM = kFseMLOff + ns.state_bits(R - (int32_t)(llLog + ofLog), mlLog);
        }
        G -= llLog + ofLog + mlLog;
    }
    uint32_t nDone = 0;                                 // sequences this lane has decoded
    uint32_t maxSeq = (live && !err) ? nbSeq : 0u;
#pragma unroll
    for (int d = 16; d; d >>= 1) maxSeq = max(maxSeq, __shfl_xor_sync(FULL, maxSeq, d));
    // Records go to HBM through a shared-memory staging block of kSeqFlush steps: one 8-byte store per lane and step to 14
    // different cache lines cost a quarter of this kernel (the chain's shared loads queue behind the scattered store in the
    // LSU); the flush writes each item's records of a block as contiguous bytes, 4 items per instruction.
    uint64_t const seqBase = (uint64_t)item * kSeqCap;
    for (uint32_t n0 = 0; n0 < maxSeq; n0 += kSeqFlush) {
        uint32_t cnt = 0;                                  // records this lane staged in the block
#pragma unroll 2
        for (uint32_t k = 0; k < kSeqFlush; k++) {
            uint32_t const n = n0 + k;
            bool const act = live && !err && n < nbSeq && (int32_t)G - gz >= kSeqTailBits;      // the last bits of a stream are left to the tail loop
            if ((k & 1) == 0) br.step<2>(G, act);
            if (act) {
                uint32_t const eL = lds16(t16 + 2 * aL), eO = lds16(t16 + 2 * aO), eM = lds16(t16 + 2 * aM);
                uint32_t const nL = lds8(t8 + aL), nO = lds8(t8 + aO), nM = lds8(t8 + aM);
                uint32_t const llBits = (eL >> 4) & 31, mlBits = (eM >> 4) & 31, ofBits = (eO >> 4) & 31;
                uint32_t const nbL = eL & 15, nbM = eM & 15, nbO = eO & 15;
                // stream order: offset extra, matchLength extra, litLength extra, then LL / ML / OF state bits (:2397-2480)
                uint32_t const G2 = G - (ofBits + mlBits + llBits), G3 = G2 - (nbL + nbM + nbO);
                uint32_t xA, xA1; br.peek64(G, xA, xA1);                      // offset extra, then matchLength + litLength extra: <= 63 bits
                uint32_t const xB = __funnelshift_l(xA1, xA, ofBits), xC = br.peek32(G2);
                uint32_t const llSym = (eL >> 9) & 63, mlSym = (eM >> 9) & 63;
                uint32_t const ofExtra = top_bits(xA, ofBits);
                uint32_t const ml = lds32(mlBaseS + mlSym * 4) + top_bits(xB, mlBits);
                uint32_t const ll = lds32(llBaseS + llSym * 4) + top_bits(xB << mlBits, llBits);     // mlBits <= 16
                aL = kFseLLOff + (nL | ((eL >> 15) << 8)) + top_bits(xC, nbL);
                uint32_t const xC2 = xC << nbL;
                aM = kFseMLOff + (nM | ((eM >> 15) << 8)) + top_bits(xC2, nbM);
                aO = kFseOFOff + (nO | ((eO >> 15) << 8)) + top_bits(xC2 << nbM, nbO);
                G = G3;
                uint32_t offset;
                {   // ZSTD_decodeSequence offset rules (:2397-2445), as selects
                    uint32_t const ll0 = (llSym == 0);                        // baseValue == 0 <=> code 0
                    uint32_t const ofv = 1u + ll0 + ofExtra;                  // only meaningful for ofBits == 1 (OF_base[1] = 1)
                    uint32_t t1 = (ofv == 3) ? rep0 - 1 : (ofv == 1 ? rep1 : rep2);
                    t1 += !t1;
                    uint32_t const big = ((1u << ofBits) - 3u) + ofExtra;     // OF_base[n] = (1<<n)-3 for n >= 2
                    uint32_t const t0 = ll0 ? rep1 : rep0;
                    offset = ofBits > 1 ? big : (ofBits == 0 ? t0 : t1);
                    // history update
                    bool const shift2 = (ofBits > 1) | ((ofBits == 1) & (ofv != 1));     // rep2 <- rep1
                    bool const shift1 = (ofBits > 0) | (ll0 != 0);                       // rep1 <- rep0
                    uint32_t const n2 = shift2 ? rep1 : rep2;
                    uint32_t const n1 = shift1 ? rep0 : rep1;
                    rep2 = n2; rep1 = n1; rep0 = offset;
                }
                // validity (ZSTD_execSequenceEnd order): output overflow, literal overrun, offset beyond frame start
                uint32_t const seqLen = ll + ml;
                bool const e1 = seqLen > dstCap - outPos, e2 = ll > litSize - litPos, e3 = offset > (outPos + ll) - frameStart;
                bool const e4 = (offset >> 30) != 0;                                             // offsets >= 1 GiB do not fit seq_pack
                if (e1 | e2 | e3 | e4) err = e1 ? kDstSizeTooSmall : kCorruptionDetected;
                ZB_ASSERTK(1, k * kSeqItemsPerCta + lane < kSeqFlush * kSeqItemsPerCta && n < kSeqCap);
                s_stage[k * kSeqItemsPerCta + lane] = seq_pack(ll, ml, offset);
                cnt = k + 1; nDone = n + 1;
                outPos += seqLen; litPos += ll;
            }
        }
        __syncwarp();
#pragma unroll
        for (uint32_t q = 0; q < (kSeqFlush * kSeqItemsPerCta + 31) / 32; q++) {
            uint32_t const idx = q * 32 + lane, slot = idx / kSeqFlush, k = idx % kSeqFlush;
            uint32_t const srcLane = slot < (uint32_t)kSeqItemsPerCta ? slot : 0u;
            uint32_t const c = __shfl_sync(FULL, cnt, srcLane);
            uint64_t const base = __shfl_sync(FULL, seqBase, srcLane);
            if (slot < (uint32_t)kSeqItemsPerCta && k < c) p.seq[base + n0 + k] = s_stage[k * kSeqItemsPerCta + slot];
        }
        __syncwarp();
    }
    // ---- tail: the sequences within kSeqTailBits of the stream start, bit reads as the reference does them there ----
    if (live && !err && nDone < nbSeq) {
        NearStart ns; ns.load(p.src + it.srcOff + it.seqOff, it.seqLen);
        for (uint32_t n = nDone; n < nbSeq && !err; n++) {
            uint32_t const eL = lds16(t16 + 2 * aL), eO = lds16(t16 + 2 * aO), eM = lds16(t16 + 2 * aM);
            uint32_t const nL = lds8(t8 + aL), nO = lds8(t8 + aO), nM = lds8(t8 + aM);
            uint32_t const llBits = (eL >> 4) & 31, mlBits = (eM >> 4) & 31, ofBits = (eO >> 4) & 31;
            uint32_t const nbL = eL & 15, nbM = eM & 15, nbO = eO & 15;
            uint32_t const llSym = (eL >> 9) & 63, mlSym = (eM >> 9) & 63;
            int32_t R = (int32_t)G - gz;
            uint32_t const ofExtra = ns.bits(R, ofBits); R -= (int32_t)ofBits;
            uint32_t const ml = lds32(mlBaseS + mlSym * 4) + ns.bits(R, mlBits); R -= (int32_t)mlBits;
            uint32_t const ll = lds32(llBaseS + llSym * 4) + ns.bits(R, llBits); R -= (int32_t)llBits;
            aL = kFseLLOff + (nL | ((eL >> 15) << 8)) + ns.state_bits(R, nbL); R -= (int32_t)nbL;
            aM = kFseMLOff + (nM | ((eM >> 15) << 8)) + ns.state_bits(R, nbM); R -= (int32_t)nbM;
            aO = kFseOFOff + (nO | ((eO >> 15) << 8)) + ns.state_bits(R, nbO); R -= (int32_t)nbO;
            G = (uint32_t)(R + gz);
            uint32_t offset;
            {   // ZSTD_decodeSequence offset rules (:2397-2445)
                uint32_t const ll0 = (llSym == 0);
                if (ofBits > 1) { offset = ((1u << ofBits) - 3u) + ofExtra; rep2 = rep1; rep1 = rep0; rep0 = offset; }
                else if (ofBits == 0) {
                    if (ll0) { offset = rep1; rep1 = rep0; rep0 = offset; } else offset = rep0;
                } else {
                    uint32_t const ofv = 1u + ll0 + ofExtra;
                    uint32_t t = (ofv == 3) ? rep0 - 1 : (ofv == 1 ? rep1 : rep2);
                    t += !t;
                    if (ofv != 1) rep2 = rep1;
                    rep1 = rep0; rep0 = offset = t;
                }
            }
            uint32_t const seqLen = ll + ml;
            bool const e1 = seqLen > dstCap - outPos, e2 = ll > litSize - litPos, e3 = offset > (outPos + ll) - frameStart;
            bool const e4 = (offset >> 30) != 0;
            if (e1 | e2 | e3 | e4) err = e1 ? kDstSizeTooSmall : kCorruptionDetected;
            ZB_ASSERTK(2, n < kSeqCap);
            p.seq[seqBase + n] = seq_pack(ll, ml, offset);
            outPos += seqLen; litPos += ll;
        }
    }
    if (!live) return;
    // the stream must not have unread bits left (BIT_reloadDStream >= completed, :2730)
    if (!err && (int32_t)G > gz) err = kCorruptionDetected;
    if (!err && (litSize - litPos) > dstCap - outPos) err = kDstSizeTooSmall;   // last literals, :2748
    if (err) { it.status = kStError; it.errCode = err; return; }
    it.rep[0] = rep0; it.rep[1] = rep1; it.rep[2] = rep2;
    it.seqLitEnd = litPos;
    it.blockOut = (outPos - outPos0) + (litSize - litPos);
}

// =====================================================================================================
//  Sequence execution + raw/RLE blocks: one warp per item (ZSTD_execSequence, ZstdDecompressBlock.cs:2187).
//  All items of a batch are resident at once (up to 64 warps per SM), so there is no block-level barrier anywhere:
//  the warp takes 32 sequences at a time (lane = sequence), a warp scan places them in a private shared-memory
//  tile, literals and matches are copied lane-parallel, matches whose source lies inside the same batch wait for
//  the lanes they depend on (rounds of ballots), long copies are taken over by the whole warp, and the tile is
//  flushed to HBM with 16-byte stores whenever the next batch does not fit.  Match sources behind the tile are read
//  back from HBM/L2 (they were flushed by this same warp earlier).
// =====================================================================================================
constexpr int kExecWarps = 2;                       // items per CTA
constexpr uint32_t kExecTileMem = 6656;             // shared memory per item; 32 warps x 6.5 KB (+1 KB reserved per CTA) = 224 KB per SM (48 warps x 4 KB: register spills, 6.2 ms)
constexpr uint32_t kExecTile = kExecTileMem - 16;   // usable bytes (the tile starts at the 16-byte phase of its HBM address)
constexpr uint32_t kExecLong = 40;                  // copies longer than this are done by the whole warp

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t const o = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= (uint32_t)d) v += o; }
    return v;
}

// forward copy of n bytes between two global buffers that do not overlap, any alignment, whole warp
__device__ void warp_copy_g2g(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane)
{
    uint32_t const head = min(n, (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15));
    if (lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    uint32_t const m = (uint32_t)((uintptr_t)src & 3), sh = m * 8;
    const uint32_t* const sw = (const uint32_t*)(src - m);
    uint32_t const nChunks = n >> 4;
    for (uint32_t c = lane; c < nChunks; c += 32) {
        const uint32_t* const w = sw + 4 * c;
        uint32_t const w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = m ? w[4] : 0u;
        uint4 v;
        v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh); v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
        *(uint4*)(dst + 16 * (size_t)c) = v;
    }
    uint32_t const done = nChunks << 4, rem = n - done;
    if (lane < rem) dst[done + lane] = src[done + lane];
}

__device__ void warp_fill_g(uint8_t* dst, uint32_t byte, uint32_t n, uint32_t lane)
{
    uint32_t const head = min(n, (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15));
    if (lane < head) dst[lane] = (uint8_t)byte;
    dst += head; n -= head;
    uint32_t const w = byte * 0x01010101u; uint4 const v = make_uint4(w, w, w, w);
    uint32_t const nChunks = n >> 4;
    for (uint32_t c = lane; c < nChunks; c += 32) *(uint4*)(dst + 16 * (size_t)c) = v;
    uint32_t const done = nChunks << 4, rem = n - done;
    if (lane < rem) dst[done + lane] = (uint8_t)byte;
}

// tile (shared) -> HBM; tile and g have the same address modulo 16
__device__ __forceinline__ void warp_flush_tile(uint8_t* g, const uint8_t* tile, uint32_t n, uint32_t lane)
{
    uint32_t const head = min(n, (uint32_t)((16 - ((uintptr_t)g & 15)) & 15));
    if (lane < head) g[lane] = tile[lane];
    g += head; tile += head; n -= head;
    uint32_t const nChunks = n >> 4;
    for (uint32_t c = lane; c < nChunks; c += 32) *(uint4*)(g + 16 * c) = *(const uint4*)(tile + 16 * c);
    uint32_t const done = nChunks << 4, rem = n - done;
    if (lane < rem) g[done + lane] = tile[done + lane];
}

// ---- per-lane copies into the tile: 4 bytes per step (unaligned source = two aligned words + funnel shift) ----
__device__ __forceinline__ uint32_t g_rd32(const uint8_t* p)        // global, any alignment; reads the two aligned words that hold p[0..3]
{
    uintptr_t const a = (uintptr_t)p; const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3); uint32_t const sh = (uint32_t)(a & 3) * 8;
    uint32_t const lo = w[0], hi = sh ? w[1] : 0u;
    return __funnelshift_r(lo, hi, sh);
}
// m (1..16) bytes out of four left-aligned words into the tile: 16 predicated byte stores, no loop
__device__ __forceinline__ void lane_put16(uint8_t* t, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t m)
{
    // two levels: most copies are at most 8 bytes long, and 16 flat predicated stores would all issue every time
    {   uint32_t const v[2] = {v0, v1};
#pragma unroll
        for (uint32_t i = 0; i < 8; i++) if (i < m) t[i] = (uint8_t)(v[i >> 2] >> ((i & 3) * 8)); }
    if (m > 8) {
        uint32_t const v[2] = {v2, v3};
#pragma unroll
        for (uint32_t i = 0; i < 8; i++) if (i + 8 < m) t[i + 8] = (uint8_t)(v[i >> 2] >> ((i & 3) * 8));
    }
}
// up to 16 bytes at g (any alignment): raw aligned words now, left-aligned words later (so that several loads overlap)
struct Raw16 { uint32_t w0, w1, w2, w3, w4, sh; };
__device__ __forceinline__ Raw16 lane_ld16(const uint8_t* g, uint32_t m)      // m = 0: nothing is read
{
    uintptr_t const a = (uintptr_t)g; const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    uint32_t const o = (uint32_t)(a & 3), need = m ? o + m : 0u;
    Raw16 r; r.sh = o * 8;
    r.w0 = need > 0 ? w[0] : 0u; r.w1 = need > 4 ? w[1] : 0u; r.w2 = need > 8 ? w[2] : 0u; r.w3 = need > 12 ? w[3] : 0u; r.w4 = need > 16 ? w[4] : 0u;
    return r;
}
__device__ __forceinline__ void lane_st16(uint8_t* t, const Raw16& r, uint32_t m)
{ lane_put16(t, __funnelshift_r(r.w0, r.w1, r.sh), __funnelshift_r(r.w1, r.w2, r.sh), __funnelshift_r(r.w2, r.w3, r.sh), __funnelshift_r(r.w3, r.w4, r.sh), m); }
// n bytes from global memory (any alignment) into the tile: the five aligned words of a 16-byte step are loaded
// together, so a copy costs one memory round trip per 16 bytes instead of one per byte
__device__ __forceinline__ void lane_copy_g2t(uint8_t* tile, uint32_t d, const uint8_t* g, uint32_t n)
{
    for (uint32_t k = 0; k < n; k += 16) {
        uint32_t const m = min(16u, n - k);
        uintptr_t const a = (uintptr_t)(g + k); const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
        uint32_t const o = (uint32_t)(a & 3), sh = o * 8, need = o + m;          // bytes [0, need) of the aligned words are wanted
        uint32_t const w0 = w[0], w1 = need > 4 ? w[1] : 0u, w2 = need > 8 ? w[2] : 0u, w3 = need > 12 ? w[3] : 0u, w4 = need > 16 ? w[4] : 0u;
        lane_put16(tile + d + k, __funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh), m);
    }
}
// forward copy inside the tile, distance d - s >= 4: every 4-byte read only touches bytes that are already final;
// with a distance >= 16 a whole 16-byte step is read before it is written
__device__ __forceinline__ void lane_copy_t2t(uint8_t* tile, uint32_t d, uint32_t s, uint32_t n)
{
    uint32_t k = 0;
    if (d - s >= 16) {
        for (; k < n; k += 16) {
            uint32_t const m = min(16u, n - k);
            uintptr_t const a = (uintptr_t)(tile + s + k); const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
            uint32_t const o = (uint32_t)(a & 3), sh = o * 8, need = o + m;
            uint32_t const w0 = w[0], w1 = need > 4 ? w[1] : 0u, w2 = need > 8 ? w[2] : 0u, w3 = need > 12 ? w[3] : 0u, w4 = need > 16 ? w[4] : 0u;
            lane_put16(tile + d + k, __funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh), m);
        }
        return;
    }
    for (; k + 4 <= n; k += 4) {
        uintptr_t const a = (uintptr_t)(tile + s + k); const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3); uint32_t const sh = (uint32_t)(a & 3) * 8;
        uint32_t const lo = w[0], hi = sh ? w[1] : 0u;
        uint32_t const v = __funnelshift_r(lo, hi, sh);
        tile[d + k] = (uint8_t)v; tile[d + k + 1] = (uint8_t)(v >> 8); tile[d + k + 2] = (uint8_t)(v >> 16); tile[d + k + 3] = (uint8_t)(v >> 24);
    }
    for (; k < n; k++) tile[d + k] = tile[s + k];
}

// Literal source of a block: raw bytes inside the frame, one repeated byte, or the Huffman output (4 padded segments)
struct LitSrc {
    const uint8_t* base; uint32_t type, rle, seg, pad;
    __device__ __forceinline__ uint32_t addr(uint32_t idx) const { uint32_t const s = (idx >= seg) + (idx >= 2 * seg) + (idx >= 3 * seg); return idx + s * pad; }
    // literals [idx, idx+n) -> global memory, whole warp
    __device__ void to_global(uint8_t* dst, uint32_t idx, uint32_t n, uint32_t lane) const {
        if (type == kLitRle) { warp_fill_g(dst, rle, n, lane); return; }
        while (n) {
            uint32_t const s = (idx >= seg) + (idx >= 2 * seg) + (idx >= 3 * seg);
            uint32_t const segEnd = s >= 3 ? 0xFFFFFFFFu : (s + 1) * seg;
            uint32_t const m = min(n, segEnd - idx);
            warp_copy_g2g(dst, base + idx + s * pad, m, lane);
            dst += m; idx += m; n -= m;
        }
    }
};

__global__ void __launch_bounds__(kExecWarps * 32) dec_exec_kernel(DecPass p)
{
    extern __shared__ __align__(16) uint8_t s_exec[];
    uint32_t const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t const item = blockIdx.x * kExecWarps + warp;
    if (item >= p.nItems) return;
    DecItem& it = p.items[item];
    if (it.status != kStRunning || it.blkType == kBlkNone) return;
    const uint8_t* const src = p.src + it.srcOff;
    uint8_t* const dst = p.dst + it.dstOff;
    uint32_t const outBase = it.outPos;
    uint32_t blockOut = 0;
    uint32_t const FULL = 0xFFFFFFFFu;

    if (it.blkType == kBlkRaw || it.blkType == kBlkRle) {
        blockOut = it.blkSize;                       // the bytes were moved by dec_rawcopy_kernel, which runs before this kernel
    } else {
        uint8_t* const tileMem = s_exec + warp * kExecTileMem;
        uint32_t const litSize = it.litSize;
        LitSrc L;
        L.type = it.litType;
        bool const fourSeg = L.type == kLitHuf && it.nStreams == 4;      // only the Huffman output buffer is segmented
        L.seg = fourSeg ? (litSize + 3) / 4 : 0x40000000u;
        L.pad = fourSeg ? lit_segment_stride(litSize) - (litSize + 3) / 4 : 0;
        L.base = L.type == kLitHuf ? p.litBuf + (size_t)item * kLitStride : src + it.litOff;
        L.rle = L.type == kLitRle ? src[it.litOff] : 0;
        uint32_t const nbSeq = it.nbSeq;
        const uint2* const aSeq = p.seq + (size_t)item * kSeqCap;
        uint32_t seqBase = 0, litPos = 0;
        uint32_t tileBase = outBase, fill = 0;                 // tile = output bytes [tileBase, tileBase + fill)
        uint8_t* tile = tileMem + ((uintptr_t)(dst + tileBase) & 15);
        uint2 rec = make_uint2(0u, 0u);
        if (lane < nbSeq) rec = aSeq[lane];
        while (seqBase < nbSeq) {
            uint32_t const n = seqBase + lane;
            uint32_t ll = 0, ml = 0, of = 0;
            if (n < nbSeq) seq_unpack(rec, ll, ml, of);
            uint32_t const len = ll + ml;
            uint32_t const oEnd = warp_incl_scan(len, lane), lEnd = warp_incl_scan(ll, lane);
            uint32_t const oStart = oEnd - len, lStart = lEnd - ll;
            uint32_t const fitMask = __ballot_sync(FULL, (n < nbSeq) && (fill + oEnd <= kExecTile));
            uint32_t const take = fitMask == FULL ? 32u : (uint32_t)__ffs((int)~fitMask) - 1u;   // leading sequences that fit (oEnd is monotone)
            if (take == 0) {
                {
                    // nothing fits: flush what we have, or (tile already empty) execute this one sequence straight in HBM
                    if (fill) {
                        ZB_ASSERTK(3, (uint64_t)tileBase + fill <= it.dstCap);
                        warp_flush_tile(dst + tileBase, tile, fill, lane);
                        tileBase += fill; fill = 0; tile = tileMem + ((uintptr_t)(dst + tileBase) & 15);
                        __syncwarp();
                        continue;
                    }
                    uint32_t const bl = __shfl_sync(FULL, ll, 0), bm = __shfl_sync(FULL, ml, 0), bo = __shfl_sync(FULL, of, 0);
                    __threadfence_block(); __syncwarp();
                    L.to_global(dst + tileBase, litPos, bl, lane);
                    __threadfence_block(); __syncwarp();
                    uint8_t* const md = dst + tileBase + bl;
                    if (bo >= 32 || bo >= bm) {
                        for (uint32_t c = 0; c < bm; c += 32) {          // every 32-byte step only reads bytes finished before it
                            uint32_t const k = c + lane;
                            if (k < bm) md[k] = __ldcg(md + ((int64_t)k - (int64_t)bo));
                            __threadfence_block(); __syncwarp();
                        }
                    } else {
                        for (uint32_t k = lane; k < bm; k += 32) md[k] = __ldcg(md - bo + (k % bo));   // periodic pattern
                    }
                    __threadfence_block(); __syncwarp();
                    tileBase += bl + bm; litPos += bl; seqBase += 1;
                    tile = tileMem + ((uintptr_t)(dst + tileBase) & 15);
                    rec = make_uint2(0u, 0u);
                    if (seqBase + lane < nbSeq) rec = aSeq[seqBase + lane];
                    continue;
                }
            }
            bool const mine = lane < take;
            ZB_ASSERTK(4, fill + __shfl_sync(FULL, oEnd, take - 1) <= kExecTile);                     // the batch fits the tile
            ZB_ASSERTK(5, !mine || (litPos + lEnd <= litSize && (uint64_t)tileBase + fill + oEnd <= it.dstCap));   // dec_seq validated every sequence
            // prefetch the next batch of records while this one is executed
            uint2 recNext = make_uint2(0u, 0u);
            if (seqBase + take + lane < nbSeq) recNext = aSeq[seqBase + take + lane];
            uint32_t const span = __shfl_sync(FULL, oEnd, take - 1), litSpan = __shfl_sync(FULL, lEnd, take - 1);
            // ---- literals -> tile; the part of every match that lies behind the tile (final bytes in HBM) is fetched in the
            // same round trip: it depends on nothing in this batch ----
            uint32_t const mDst = fill + oStart + ll;                    // tile-relative destination of my match
            int32_t const srcRel = (int32_t)mDst - (int32_t)of;          // tile-relative source (negative: behind the tile, in HBM)
            uint32_t const farN = (mine && ml && ml <= kExecLong && srcRel < 0) ? min(ml, (uint32_t)(-srcRel)) : 0u;
            uint32_t const farFast = farN <= 16 ? farN : 0u;             // longer far parts are copied in the rounds below
            ZB_ASSERTK(6, !(mine && ml) || (int64_t)tileBase + (int64_t)srcRel >= -(int64_t)it.prefix);   // a match never reaches in front of the frame (or its dictionary content)
            {
                uint32_t const d0 = fill + oStart;
                // the per-lane path below handles a run that crosses at most ONE segment boundary: with four streams over fewer than
                // 4 * kExecLong literals (legal, though the reference's encoder never emits it) a run can cross two, so those go the
                // whole-warp way, which maps every byte through L.addr()
                bool const llLong = ll > kExecLong || (ll && L.seg <= kExecLong);
                uint32_t longMask = __ballot_sync(FULL, mine && llLong);
                uint32_t const myLL = (mine && !llLong) ? ll : 0u;
                if (L.type == kLitRle) {
                    Raw16 const fm = lane_ld16(dst + tileBase + srcRel, farFast);
                    for (uint32_t k = 0; k < myLL; k++) tile[d0 + k] = (uint8_t)L.rle;
                    if (farFast) lane_st16(tile + mDst, fm, farFast);
                } else {
                    uint32_t const idx = litPos + lStart;
                    uint32_t const s = (idx >= L.seg) + (idx >= 2 * L.seg) + (idx >= 3 * L.seg);
                    uint32_t const a = idx + s * L.pad;
                    uint32_t const n1 = s >= 3 ? myLL : min(myLL, (s + 1) * L.seg - idx);      // bytes before the next segment boundary
                    uint32_t const l16 = min(n1, 16u);
                    Raw16 const lt = lane_ld16(L.base + a, l16);
                    Raw16 const fm = lane_ld16(dst + tileBase + srcRel, farFast);
                    if (l16) lane_st16(tile + d0, lt, l16);
                    if (farFast) lane_st16(tile + mDst, fm, farFast);
                    if (n1 > 16) lane_copy_g2t(tile, d0 + 16, L.base + a + 16, n1 - 16);
                    if (n1 < myLL) lane_copy_g2t(tile, d0 + n1, L.base + a + n1 + L.pad, myLL - n1);   // at most one boundary: llLong above
                }
                while (longMask) {
                    uint32_t const j = (uint32_t)__ffs((int)longMask) - 1u; longMask &= longMask - 1;
                    uint32_t const jl = __shfl_sync(FULL, ll, j), jd = __shfl_sync(FULL, d0, j), ji = __shfl_sync(FULL, litPos + lStart, j);
                    if (L.type == kLitRle) { for (uint32_t k = lane; k < jl; k += 32) tile[jd + k] = (uint8_t)L.rle; }
                    else { for (uint32_t k = lane; k < jl; k += 32) tile[jd + k] = __ldg(L.base + L.addr(ji + k)); }
                }
            }
            __syncwarp();
            // ---- matches ----
            uint32_t const nonSelf = of < ml ? of : ml;
            // lanes j < lane whose match output [mDst_j, oEnd_j) intersects my source [srcRel, srcRel + nonSelf)
            uint32_t depMask = 0;
            {
                int32_t const a = srcRel, b = srcRel + (int32_t)nonSelf;
                bool const inBatch = mine && ml && (b > (int32_t)fill);
                if (__any_sync(FULL, inBatch)) {
                    int32_t const myEnd = (int32_t)(fill + oEnd), myMD = (int32_t)mDst;
                    // jLo = first j with end_j > a ; jHi = last j with mDst_j < b   (both sorted by lane)
                    uint32_t lo = 0, hi = lane;
#pragma unroll
                    for (int stp = 0; stp < 5; stp++) {
                        uint32_t const mid = (lo + hi) >> 1;
                        int32_t const e = __shfl_sync(FULL, myEnd, mid & 31);
                        if (lo < hi) { if (e > a) hi = mid; else lo = mid + 1; }
                    }
                    uint32_t const jLo = lo;
                    lo = 0; hi = lane;
#pragma unroll
                    for (int stp = 0; stp < 5; stp++) {
                        uint32_t const mid = (lo + hi) >> 1;
                        int32_t const sft = __shfl_sync(FULL, myMD, mid & 31);
                        if (lo < hi) { if (sft < b) lo = mid + 1; else hi = mid; }
                    }
                    // lo = number of lanes j < lane with mDst_j < b
                    if (inBatch && lo > jLo) depMask = ((lo >= 32 ? 0u : (1u << lo)) - 1u) & ~((1u << jLo) - 1u);
                }
            }
            uint32_t done = ~(take >= 32 ? FULL : ((1u << take) - 1u));   // lanes outside the batch count as finished
            bool pending = mine;
            if (mine && ml == 0) pending = false;
            done |= __ballot_sync(FULL, mine && !pending);
            while (__any_sync(FULL, pending)) {
                bool const ready = pending && ((depMask & ~done) == 0);
                bool const isLong = ready && ml > kExecLong;
                if (ready && !isLong) {
                    const uint8_t* const gsrc = dst + tileBase;            // tile origin in HBM
                    uint32_t k0 = 0;
                    if (srcRel < 0) { k0 = min(ml, (uint32_t)(-srcRel)); if (!farFast) lane_copy_g2t(tile, mDst, gsrc + srcRel, k0); }
                    if (k0 < ml) {
                        if (of >= 4) lane_copy_t2t(tile, mDst + k0, (uint32_t)(srcRel + (int32_t)k0), ml - k0);
                        else { volatile uint8_t* const vt = tile; for (uint32_t k = k0; k < ml; k++) vt[mDst + k] = vt[(uint32_t)(srcRel + (int32_t)k)]; }   // offsets 1..3: byte by byte
                    }
                }
                uint32_t longMask = __ballot_sync(FULL, isLong);
                while (longMask) {
                    uint32_t const j = (uint32_t)__ffs((int)longMask) - 1u; longMask &= longMask - 1;
                    uint32_t const jm = __shfl_sync(FULL, ml, j), jo = __shfl_sync(FULL, of, j), jd = __shfl_sync(FULL, mDst, j);
                    int32_t const js = __shfl_sync(FULL, srcRel, j);
                    const uint8_t* const gsrc = dst + tileBase;
                    if (jo >= 32 || jo >= jm) {
                        for (uint32_t c = 0; c < jm; c += 32) {
                            uint32_t const k = c + lane;
                            if (k < jm) { int32_t const sp = js + (int32_t)k; tile[jd + k] = sp >= 0 ? tile[sp] : gsrc[sp]; }
                            __syncwarp();
                        }
                    } else {
                        for (uint32_t k = lane; k < jm; k += 32) { int32_t const sp = js + (int32_t)(k % jo); tile[jd + k] = sp >= 0 ? tile[sp] : gsrc[sp]; }
                    }
                }
                __syncwarp();
                done |= __ballot_sync(FULL, ready);
                if (ready) pending = false;
            }
            fill += span; litPos += litSpan; seqBase += take;
            rec = recNext;
        }
        // flush, then the last literals go straight to HBM (ZstdDecompressBlock.cs:2748-2760)
        __syncwarp();
        ZB_ASSERTK(7, (uint64_t)tileBase + fill + (litSize - litPos) <= it.dstCap && litPos <= litSize);
        if (fill) { warp_flush_tile(dst + tileBase, tile, fill, lane); tileBase += fill; }
        uint32_t const lastLL = litSize - litPos;
        L.to_global(dst + tileBase, litPos, lastLL, lane);
        blockOut = tileBase + lastLL - outBase;
    }
    __threadfence_block(); __syncwarp();
    // ---- end of block bookkeeping (ZSTD_decompressFrame loop tail, ZstdDecompress.cs:1156-1212) ----
    // frame checksum: low 32 bits of XXH64 over the regenerated frame (:1186-1207), hashed by the whole warp
    uint32_t checkCalc = 0;
    bool const wantCheck = it.lastBlock && it.checksumFlag && (it.srcSize - it.srcPos >= 4) &&
                           !(it.hasFcs && (uint64_t)(outBase + blockOut - it.frameStart) != it.fcs);
    if (wantCheck) checkCalc = (uint32_t)xxh64_warp(dst + it.frameStart, outBase + blockOut - it.frameStart, lane);
    if (lane == 0) {
        it.outPos = outBase + blockOut;
        if (it.lastBlock) {
            uint32_t err = 0;
            if (it.hasFcs && (uint64_t)(it.outPos - it.frameStart) != it.fcs) err = kCorruptionDetected;
            uint32_t pos = it.srcPos;
            if (!err && it.checksumFlag) {
                if (it.srcSize - pos < 4) err = kChecksumWrong;
                else { if (ld_le32(src + pos) != checkCalc) err = kChecksumWrong; pos += 4; }
            }
            if (err) { it.status = kStError; it.errCode = err; }
            else {
                it.inFrame = 0; it.moreThan1Frame = 1;
                advance_frames(it, src, &pos);
                it.srcPos = pos;
            }
        }
    }
}

// =====================================================================================================
//  Raw and RLE blocks (ZSTD_copyRawBlock / ZSTD_setRleBlock, ZstdDecompress.cs:1004, :1029): a streaming copy, the one part of the
//  decoder that is bound by HBM bandwidth.  Persistent CTAs walk (block, 64 KiB half) units (smaller units spent more time on the dependent descriptor loads than on the copy).  Destination chunks are 16-byte
//  aligned; the source is read as ALIGNED 16-byte words (the chunk and its successor, which the neighbouring thread also reads: an
//  L1 hit) and shifted into place, four chunks per thread in flight.  cp.async.bulk (TMA) cannot do this copy: it needs both
//  addresses 16-byte aligned, and a raw block sits 12 bytes into its frame while its destination is frame-aligned.
// =====================================================================================================
constexpr uint32_t kRawSeg = 65536, kRawSegsPerBlock = kBlockSizeMax / kRawSeg, kRawThreads = 256;
__global__ void __launch_bounds__(kRawThreads) dec_rawcopy_kernel(DecPass p)
{
    uint32_t const nUnits = p.counters[4] * kRawSegsPerBlock;
    for (uint32_t u = blockIdx.x; u < nUnits; u += gridDim.x) {
        DecItem const& it = p.items[p.rawList[u / kRawSegsPerBlock]];
        if (it.status != kStRunning) continue;
        uint32_t const lo = (u % kRawSegsPerBlock) * kRawSeg, size = it.blkSize;
        if (lo >= size) continue;
        uint32_t n = min(kRawSeg, size - lo);
        ZB_ASSERTK(8, (uint64_t)it.outPos + size <= it.dstCap && (it.blkType == kBlkRle || (uint64_t)it.blkSrcOff + size <= it.srcSize));
        uint8_t* dst = p.dst + it.dstOff + it.outPos + lo;
        uint32_t const t = threadIdx.x;
        uint32_t const head = min(n, (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15));
        if (it.blkType == kBlkRle) {
            uint32_t const byte = p.src[it.srcOff + it.blkSrcOff], w = byte * 0x01010101u;
            if (t < head) dst[t] = (uint8_t)byte;
            dst += head; n -= head;
            uint32_t const nChunks = n >> 4;
            for (uint32_t c = t; c < nChunks; c += kRawThreads) *(uint4*)(dst + 16 * (size_t)c) = make_uint4(w, w, w, w);
            uint32_t const done = nChunks << 4;
            if (t < n - done) dst[done + t] = (uint8_t)byte;
            continue;
        }
        const uint8_t* src = p.src + it.srcOff + it.blkSrcOff + lo;
        if (t < head) dst[t] = src[t];
        dst += head; src += head; n -= head;
        uint32_t const sh = (uint32_t)((uintptr_t)src & 15), w0 = sh >> 2, bsh = (sh & 3) * 8;
        const uint4* const s4 = (const uint4*)(src - sh);
        uint32_t const nChunks = n >> 4;
        for (uint32_t base = 0; base < nChunks; base += 4 * kRawThreads) {     // uniform trip count: the shuffles below need the whole warp
            uint32_t const c0 = base + t;
            uint4 a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t const c = c0 + kRawThreads * k;
                a[k] = (c < nChunks || (c == nChunks && sh)) ? __ldg(s4 + c) : make_uint4(0, 0, 0, 0);      // one chunk more: the last full chunk's successor
                b[k] = (c < nChunks && sh && (t & 31) == 31) ? __ldg(s4 + c + 1) : make_uint4(0, 0, 0, 0);    // the successor chunk is the next lane's chunk
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint4 const nx = make_uint4(__shfl_down_sync(0xFFFFFFFFu, a[k].x, 1), __shfl_down_sync(0xFFFFFFFFu, a[k].y, 1),
                                            __shfl_down_sync(0xFFFFFFFFu, a[k].z, 1), __shfl_down_sync(0xFFFFFFFFu, a[k].w, 1));
                if ((t & 31) != 31) b[k] = nx;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t const c = c0 + kRawThreads * k;
                uint32_t v0, v1, v2, v3, v4;
                switch (w0) {               // uniform for the unit
                case 0: v0 = a[k].x; v1 = a[k].y; v2 = a[k].z; v3 = a[k].w; v4 = b[k].x; break;
                case 1: v0 = a[k].y; v1 = a[k].z; v2 = a[k].w; v3 = b[k].x; v4 = b[k].y; break;
                case 2: v0 = a[k].z; v1 = a[k].w; v2 = b[k].x; v3 = b[k].y; v4 = b[k].z; break;
                default: v0 = a[k].w; v1 = b[k].x; v2 = b[k].y; v3 = b[k].z; v4 = b[k].w; break;
                }
                uint4 v;
                v.x = __funnelshift_r(v0, v1, bsh); v.y = __funnelshift_r(v1, v2, bsh); v.z = __funnelshift_r(v2, v3, bsh); v.w = __funnelshift_r(v3, v4, bsh);
                if (c < nChunks) __stcs((uint4*)(dst + 16 * (size_t)c), v);
            }
        }
        uint32_t const done = nChunks << 4;
        if (t < n - done) dst[done + t] = src[done + t];
    }
}

__global__ void dec_finish_kernel(DecPass p)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nItems) return;
    DecItem const& it = p.items[i];
    uint64_t r;
    if (it.status == kStDone) r = it.outPos;
    else if (it.status == kStError) r = make_error(it.errCode);
    else r = make_error(kGeneric);
    p.results[i] = r;
}

__global__ void dec_reset_counters_kernel(uint32_t* counters) { if (threadIdx.x < 3) counters[threadIdx.x] = 0; if (threadIdx.x == 4) counters[4] = 0; }

// opt-in shared memory sizes are a per-device function attribute: set them once per device
static void dec_set_attrs()
{
    static bool done[64] = {};
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || done[dev]) return;
    cudaFuncSetAttribute(dec_huf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHufSmemBytes);
    cudaFuncSetAttribute(dec_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqSmemBytes);
    cudaFuncSetAttribute(dec_exec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kExecWarps * kExecTileMem);
    // One shared-memory carve-out for every kernel of the pipeline: kernels of different sub-batches (streams) can only
    // share an SM when they agree on the L1/shared split, otherwise each switch waits for the SM to drain.
    cudaFuncSetAttribute(dec_scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(dec_reset_counters_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(dec_setup_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(dec_huf_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(dec_seq_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(dec_exec_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(dec_finish_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(dec_rawcopy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    done[dev] = true;
}

// persistent grid: 8 CTAs of 256 threads per SM at most, never more CTAs than there can be units
static unsigned raw_grid(const DecPass& p)
{
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    return (unsigned)std::min<uint64_t>((uint64_t)sms * 8, (uint64_t)p.nItems * kRawSegsPerBlock);
}

#ifdef ZB_DEBUG_ASSERTS
// developer aid of the assertion build (ZSTDB200_DUMP_ITEM=<item>): the last literals and the last output bytes of one item between kernels
__global__ void dec_debug_dump_kernel(DecPass p, uint32_t item, int stage)
{
    DecItem const& it = p.items[item];
    uint32_t const litSize = it.litSize, seg = (litSize + 3) / 4, stride = lit_segment_stride(litSize);
    const uint8_t* lb = p.litBuf + (size_t)item * kLitStride;
    printf("dump stage %d item %u: status %u litType %u litSize %u nbSeq %u outPos %u | last literals:", stage, item, it.status, it.litType, litSize, it.nbSeq, it.outPos);
    for (uint32_t i = litSize >= 12 ? litSize - 12 : 0; i < litSize; i++) { uint32_t const sgi = it.nStreams == 4 ? min(3u, i / seg) : 0u; printf(" %02x", lb[i + sgi * (stride - seg)]); }
    if (stage >= 2) {
        printf(" | last records:");
        for (uint32_t n = it.nbSeq >= 4 ? it.nbSeq - 4 : 0; n < it.nbSeq; n++) { uint32_t ll, ml, of; seq_unpack(p.seq[(size_t)item * kSeqCap + n], ll, ml, of); printf(" (%u,%u,%u)", ll, ml, of); }
    }
    printf(" | dst tail:");
    for (uint32_t i = it.dstCap >= 12 ? it.dstCap - 12 : 0; i < it.dstCap; i++) printf(" %02x", p.dst[it.dstOff + i]);
    printf("\n");
}
static void debug_dump(const DecPass& p, cudaStream_t s, int stage)
{
    static int const item = []() { const char* e = getenv("ZSTDB200_DUMP_ITEM"); return e ? atoi(e) : -1; }();
    if (item >= 0 && (uint32_t)item < p.nItems) dec_debug_dump_kernel<<<1, 1, 0, s>>>(p, (uint32_t)item, stage);
}
#else
static inline void debug_dump(const DecPass&, cudaStream_t, int) {}
#endif

void dec_launch_wave(const DecPass& p, cudaStream_t s)
{
    dec_reset_counters_kernel<<<1, 32, 0, s>>>(p.counters);
    dec_setup_kernel<<<(p.nItems + kSetupWarps - 1) / kSetupWarps, kSetupWarps * 32, 0, s>>>(p);
    dec_set_attrs();
    dec_huf_kernel<<<(p.nItems + kHufItemsPerCta - 1) / kHufItemsPerCta, kHufThreads, kHufSmemBytes, s>>>(p);
    debug_dump(p, s, 1);
    dec_seq_kernel<<<(p.nItems + kSeqItemsPerCta - 1) / kSeqItemsPerCta, 32, kSeqSmemBytes, s>>>(p);
    debug_dump(p, s, 2);
    dec_rawcopy_kernel<<<raw_grid(p), kRawThreads, 0, s>>>(p);
    dec_exec_kernel<<<(p.nItems + kExecWarps - 1) / kExecWarps, kExecWarps * 32, kExecWarps * kExecTileMem, s>>>(p);
    debug_dump(p, s, 3);
}

void dec_launch_wave_timed(const DecPass& p, cudaStream_t s, cudaEvent_t* ev)
{
    dec_reset_counters_kernel<<<1, 32, 0, s>>>(p.counters);
    cudaEventRecord(ev[0], s);
    dec_setup_kernel<<<(p.nItems + kSetupWarps - 1) / kSetupWarps, kSetupWarps * 32, 0, s>>>(p);
    cudaEventRecord(ev[1], s);
    dec_set_attrs();
    dec_huf_kernel<<<(p.nItems + kHufItemsPerCta - 1) / kHufItemsPerCta, kHufThreads, kHufSmemBytes, s>>>(p);
    cudaEventRecord(ev[2], s);
    dec_seq_kernel<<<(p.nItems + kSeqItemsPerCta - 1) / kSeqItemsPerCta, 32, kSeqSmemBytes, s>>>(p);
    cudaEventRecord(ev[3], s);
    dec_rawcopy_kernel<<<raw_grid(p), kRawThreads, 0, s>>>(p);
    dec_exec_kernel<<<(p.nItems + kExecWarps - 1) / kExecWarps, kExecWarps * 32, kExecWarps * kExecTileMem, s>>>(p);
    cudaEventRecord(ev[4], s);
}

void dec_launch_finish(const DecPass& p, cudaStream_t s)
{
    dec_finish_kernel<<<(p.nItems + 255) / 256, 256, 0, s>>>(p);
}

void dec_launch_scan_init(const DecPass& p, const void* init, cudaStream_t s)
{
    dec_scan_kernel<<<(p.nItems + 127) / 128, 128, 0, s>>>(p, (const DecItemInit*)init);
}

}  // namespace zb
