/*
 * zo_decode.c -- CPU ORACLE, decode side (test infrastructure only; see zo_common.h).
 *
 * Restates, in plain C, the reference's frame decoder for the no-dictionary one-shot path:
 *   ZSTD_decompressDCtx -> ZSTD_decompressMultiFrame -> ZSTD_decompressFrame
 *     -> ZSTD_decompressBlock_internal -> ZSTD_decodeLiteralsBlock / HUF_decompress{1,4}X1
 *        -> ZSTD_decodeSeqHeaders / ZSTD_buildFSETable -> ZSTD_decompressSequences_body
 *           (ZSTD_decodeSequence + ZSTD_execSequence)
 * Citations are file:line under /root/reference/src/ZstdSharp/Unsafe/.
 * The result-neutral CPU variants (split literal buffer, prefetching "Long" decoder: ZstdDecompressBlock.cs:2487, :2796)
 * are deliberately not restated.  The double-symbol Huffman decoder (HufDecompress.cs:652-1455) IS restated: on valid streams
 * it yields the bytes of the single-symbol decoder, but on damaged streams its verdict differs (HUF_decodeLastSymbolX2's
 * bit clamp, the op1 > opStart2 checks), and the reference picks it through HUF_selectDecoder (:1688).
 */
#include "zo_common.h"
#include "zo.h"

/* =====================================================================================
 *  Backward bit reader -- Bitstream.cs:172-425 (BIT_DStream_t), restated operation by operation
 * ===================================================================================== */
typedef struct { U64 bitContainer; unsigned bitsConsumed; const BYTE* ptr; const BYTE* start; const BYTE* limitPtr; } BIT_DStream_t;
typedef enum { BIT_DStream_unfinished = 0, BIT_DStream_endOfBuffer = 1, BIT_DStream_completed = 2, BIT_DStream_overflow = 3 } BIT_DStream_status;

/* Bitstream.cs:172 */
static size_t BIT_initDStream(BIT_DStream_t* bitD, const void* srcBuffer, size_t srcSize)
{
    const BYTE* src = (const BYTE*)srcBuffer;
    if (srcSize < 1) { memset(bitD, 0, sizeof(*bitD)); return ERROR(srcSize_wrong); }
    bitD->start = src;
    bitD->limitPtr = bitD->start + sizeof(U64);
    if (srcSize >= sizeof(U64)) {
        bitD->ptr = src + srcSize - sizeof(U64);
        bitD->bitContainer = MEM_read64(bitD->ptr);
        { BYTE const lastByte = src[srcSize - 1];
          bitD->bitsConsumed = lastByte ? 8 - BIT_highbit32(lastByte) : 0;
          if (lastByte == 0) return ERROR(GENERIC); }
    } else {
        bitD->ptr = bitD->start;
        bitD->bitContainer = src[0];
        switch (srcSize) {
        case 7: bitD->bitContainer += (U64)src[6] << 48; /* fall-through */
        case 6: bitD->bitContainer += (U64)src[5] << 40; /* fall-through */
        case 5: bitD->bitContainer += (U64)src[4] << 32; /* fall-through */
        case 4: bitD->bitContainer += (U64)src[3] << 24; /* fall-through */
        case 3: bitD->bitContainer += (U64)src[2] << 16; /* fall-through */
        case 2: bitD->bitContainer += (U64)src[1] << 8;  /* fall-through */
        default: break;
        }
        { BYTE const lastByte = src[srcSize - 1];
          bitD->bitsConsumed = lastByte ? 8 - BIT_highbit32(lastByte) : 0;
          if (lastByte == 0) return ERROR(corruption_detected); }
        bitD->bitsConsumed += (U32)(sizeof(U64) - srcSize) * 8;
    }
    return srcSize;
}
/* Bitstream.cs BIT_lookBits / BIT_getMiddleBits (shift count masked to the register width, as C# does) */
static inline size_t BIT_lookBits(const BIT_DStream_t* bitD, U32 nbBits)
{
    U32 const start = (U32)(64 - bitD->bitsConsumed - nbBits);
    return (size_t)((bitD->bitContainer >> (start & 63)) & ((((U64)1) << (nbBits & 63)) - 1));
}
static inline size_t BIT_lookBitsFast(const BIT_DStream_t* bitD, U32 nbBits)
{
    return (size_t)((bitD->bitContainer << (bitD->bitsConsumed & 63)) >> ((64 - nbBits) & 63));
}
static inline void BIT_skipBits(BIT_DStream_t* bitD, U32 nbBits) { bitD->bitsConsumed += nbBits; }
static inline size_t BIT_readBits(BIT_DStream_t* bitD, U32 nbBits) { size_t const v = BIT_lookBits(bitD, nbBits); BIT_skipBits(bitD, nbBits); return v; }
static inline size_t BIT_readBitsFast(BIT_DStream_t* bitD, U32 nbBits) { size_t const v = BIT_lookBitsFast(bitD, nbBits); BIT_skipBits(bitD, nbBits); return v; }
static inline BIT_DStream_status BIT_reloadDStreamFast(BIT_DStream_t* bitD)
{
    if (bitD->ptr < bitD->limitPtr) return BIT_DStream_overflow;
    bitD->ptr -= bitD->bitsConsumed >> 3;
    bitD->bitsConsumed &= 7;
    bitD->bitContainer = MEM_read64(bitD->ptr);
    return BIT_DStream_unfinished;
}
static inline BIT_DStream_status BIT_reloadDStream(BIT_DStream_t* bitD)
{
    if (bitD->bitsConsumed > 64) return BIT_DStream_overflow;
    if (bitD->ptr >= bitD->limitPtr) return BIT_reloadDStreamFast(bitD);
    if (bitD->ptr == bitD->start) {
        if (bitD->bitsConsumed < 64) return BIT_DStream_endOfBuffer;
        return BIT_DStream_completed;
    }
    {   U32 nbBytes = bitD->bitsConsumed >> 3;
        BIT_DStream_status result = BIT_DStream_unfinished;
        if (bitD->ptr - nbBytes < bitD->start) { nbBytes = (U32)(bitD->ptr - bitD->start); result = BIT_DStream_endOfBuffer; }
        bitD->ptr -= nbBytes;
        bitD->bitsConsumed -= nbBytes * 8;
        bitD->bitContainer = MEM_read64(bitD->ptr);
        return result;
    }
}
static inline unsigned BIT_endOfDStream(const BIT_DStream_t* d) { return (d->ptr == d->start) && (d->bitsConsumed == 64); }

/* =====================================================================================
 *  FSE normalized-count header -- EntropyCommon.cs:52-242 (FSE_readNCount_body)
 * ===================================================================================== */
static inline U32 cs_shr32(U32 v, int n) { return v >> (n & 31); }   /* C# uint >> int masks the count */

static size_t FSE_readNCount(S16* normalizedCounter, unsigned* maxSVPtr, unsigned* tableLogPtr, const void* headerBuffer, size_t hbSize)
{
    const BYTE* const istart = (const BYTE*)headerBuffer;
    const BYTE* const iend = istart + hbSize;
    const BYTE* ip = istart;
    int nbBits, remaining, threshold, bitCount;
    U32 bitStream;
    unsigned charnum = 0;
    unsigned const maxSV1 = *maxSVPtr + 1;
    int previous0 = 0;

    if (hbSize < 8) {   /* EntropyCommon.cs:66-84 : pad to 8 bytes and retry */
        BYTE buffer[8] = { 0 };
        memcpy(buffer, headerBuffer, hbSize);
        {   size_t const countSize = FSE_readNCount(normalizedCounter, maxSVPtr, tableLogPtr, buffer, sizeof(buffer));
            if (ERR_isError(countSize)) return countSize;
            if (countSize > hbSize) return ERROR(corruption_detected);
            return countSize;
        }
    }
    memset(normalizedCounter, 0, (*maxSVPtr + 1) * sizeof(normalizedCounter[0]));
    bitStream = MEM_read32(ip);
    nbBits = (int)((bitStream & 0xF) + 5);
    if (nbBits > 15) return ERROR(tableLog_tooLarge);
    bitStream >>= 4;
    bitCount = 4;
    *tableLogPtr = (unsigned)nbBits;
    remaining = (1 << nbBits) + 1;
    threshold = 1 << nbBits;
    nbBits++;

    for (;;) {
        if (previous0) {
            int repeats = (int)((U32)__builtin_ctz(~bitStream | 0x80000000) >> 1);
            while (repeats >= 12) {
                charnum += 3 * 12;
                if (ip <= iend - 7) { ip += 3; }
                else { bitCount -= (int)(8 * (iend - 7 - ip)); bitCount &= 31; ip = iend - 4; }
                bitStream = cs_shr32(MEM_read32(ip), bitCount);
                repeats = (int)((U32)__builtin_ctz(~bitStream | 0x80000000) >> 1);
            }
            charnum += 3 * (unsigned)repeats;
            bitStream = cs_shr32(bitStream, 2 * repeats);
            bitCount += 2 * repeats;
            charnum += bitStream & 3;
            bitCount += 2;
            if (charnum >= maxSV1) break;
            if ((ip <= iend - 7) || (ip + (bitCount >> 3) <= iend - 4)) { ip += bitCount >> 3; bitCount &= 7; }
            else { bitCount -= (int)(8 * (iend - 4 - ip)); bitCount &= 31; ip = iend - 4; }
            bitStream = cs_shr32(MEM_read32(ip), bitCount);
        }
        {   int const max = (2 * threshold - 1) - remaining;
            int count;
            if ((bitStream & (U32)(threshold - 1)) < (U32)max) {
                count = (int)(bitStream & (U32)(threshold - 1));
                bitCount += nbBits - 1;
            } else {
                count = (int)(bitStream & (U32)(2 * threshold - 1));
                if (count >= threshold) count -= max;
                bitCount += nbBits;
            }
            count--;
            if (count >= 0) remaining -= count; else remaining += count;
            normalizedCounter[charnum++] = (S16)count;
            previous0 = !count;
            if (remaining < threshold) {
                if (remaining <= 1) break;
                nbBits = (int)BIT_highbit32((U32)remaining) + 1;
                threshold = 1 << (nbBits - 1);
            }
            if (charnum >= maxSV1) break;
            if ((ip <= iend - 7) || (ip + (bitCount >> 3) <= iend - 4)) { ip += bitCount >> 3; bitCount &= 7; }
            else { bitCount -= (int)(8 * (iend - 4 - ip)); bitCount &= 31; ip = iend - 4; }
            bitStream = cs_shr32(MEM_read32(ip), bitCount);
        }
    }
    if (remaining != 1) return ERROR(corruption_detected);
    if (charnum > maxSV1) return ERROR(maxSymbolValue_tooSmall);
    if (bitCount > 32) return ERROR(corruption_detected);
    *maxSVPtr = charnum - 1;
    ip += (bitCount + 7) >> 3;
    return (size_t)(ip - istart);
}

/* =====================================================================================
 *  FSE decoding of Huffman weights -- FseDecompress.cs:25-176 (table), :230-312 (2-state decode),
 *  :334-397 (FSE_decompress_wksp_body)
 * ===================================================================================== */
typedef struct { U16 newState; BYTE symbol; BYTE nbBits; } FSE_decode_t;
typedef struct { unsigned tableLog; unsigned fastMode; FSE_decode_t cell[1 << 12]; } FSE_DTable;

/* symbol spread shared by FSE_buildDTable_internal (FseDecompress.cs:25) and ZSTD_buildFSETable_body
 * (ZstdDecompressBlock.cs:1571): low-prob symbols occupy the top cells, the others are stepped
 * through with step (size>>1)+(size>>3)+3.  The reference's "no low-prob" fast path (FseDecompress.cs:85-118)
 * produces the same cells as this generic walk. */
static size_t zo_fse_spread(BYTE* cellSymbol, U16* symbolNext, unsigned* fastMode, const S16* norm, unsigned maxSymbolValue, unsigned tableLog)
{
    U32 const tableSize = 1U << tableLog;
    U32 const tableMask = tableSize - 1;
    U32 const step = (tableSize >> 1) + (tableSize >> 3) + 3;
    U32 highThreshold = tableSize - 1;
    S16 const largeLimit = (S16)(1 << (tableLog - 1));
    U32 s, position = 0;
    *fastMode = 1;
    for (s = 0; s <= maxSymbolValue; s++) {
        if (norm[s] == -1) { cellSymbol[highThreshold--] = (BYTE)s; symbolNext[s] = 1; }
        else { if (norm[s] >= largeLimit) *fastMode = 0; symbolNext[s] = (U16)norm[s]; }
    }
    for (s = 0; s <= maxSymbolValue; s++) {
        int i;
        for (i = 0; i < norm[s]; i++) {
            cellSymbol[position] = (BYTE)s;
            position = (position + step) & tableMask;
            while (position > highThreshold) position = (position + step) & tableMask;
        }
    }
    if (position != 0) return ERROR(GENERIC);
    return 0;
}

static size_t FSE_buildDTable(FSE_DTable* dt, const S16* norm, unsigned maxSymbolValue, unsigned tableLog)
{
    BYTE cellSymbol[1 << 12]; U16 symbolNext[256]; U32 u;
    U32 const tableSize = 1U << tableLog;
    if (maxSymbolValue > 255) return ERROR(maxSymbolValue_tooLarge);
    if (tableLog > 12) return ERROR(tableLog_tooLarge);
    dt->tableLog = tableLog;
    CHECK_F(zo_fse_spread(cellSymbol, symbolNext, &dt->fastMode, norm, maxSymbolValue, tableLog));
    for (u = 0; u < tableSize; u++) {
        BYTE const symbol = cellSymbol[u];
        U32 const nextState = symbolNext[symbol]++;
        dt->cell[u].symbol = symbol;
        dt->cell[u].nbBits = (BYTE)(tableLog - BIT_highbit32(nextState));
        dt->cell[u].newState = (U16)((nextState << dt->cell[u].nbBits) - tableSize);
    }
    return 0;
}

typedef struct { size_t state; const FSE_decode_t* table; } FSE_DState_t;
static inline void FSE_initDState(FSE_DState_t* s, BIT_DStream_t* bitD, const FSE_DTable* dt)
{ s->state = BIT_readBits(bitD, dt->tableLog); BIT_reloadDStream(bitD); s->table = dt->cell; }          /* Fse.cs:98-107 */
static inline BYTE FSE_decodeSymbol(FSE_DState_t* s, BIT_DStream_t* bitD, int fast)                       /* Fse.cs:124-153 */
{
    FSE_decode_t const d = s->table[s->state];
    size_t const lowBits = fast ? BIT_readBitsFast(bitD, d.nbBits) : BIT_readBits(bitD, d.nbBits);
    s->state = d.newState + lowBits;
    return d.symbol;
}

/* FseDecompress.cs:230 FSE_decompress_usingDTable_generic (64-bit build: no intermediate reloads) */
static size_t FSE_decompress_usingDTable(BYTE* dst, size_t maxDstSize, const void* cSrc, size_t cSrcSize, const FSE_DTable* dt)
{
    BYTE* const ostart = dst; BYTE* op = ostart; BYTE* const omax = op + maxDstSize; BYTE* const olimit = omax - 3;
    int const fast = (int)dt->fastMode;
    BIT_DStream_t bitD; FSE_DState_t state1, state2;
    CHECK_F(BIT_initDStream(&bitD, cSrc, cSrcSize));
    FSE_initDState(&state1, &bitD, dt);
    FSE_initDState(&state2, &bitD, dt);
    for (; (BIT_reloadDStream(&bitD) == BIT_DStream_unfinished) && (op < olimit); op += 4) {
        op[0] = FSE_decodeSymbol(&state1, &bitD, fast);
        op[1] = FSE_decodeSymbol(&state2, &bitD, fast);
        op[2] = FSE_decodeSymbol(&state1, &bitD, fast);
        op[3] = FSE_decodeSymbol(&state2, &bitD, fast);
    }
    for (;;) {
        if (op > (omax - 2)) return ERROR(dstSize_tooSmall);
        *op++ = FSE_decodeSymbol(&state1, &bitD, fast);
        if (BIT_reloadDStream(&bitD) == BIT_DStream_overflow) { *op++ = FSE_decodeSymbol(&state2, &bitD, fast); break; }
        if (op > (omax - 2)) return ERROR(dstSize_tooSmall);
        *op++ = FSE_decodeSymbol(&state2, &bitD, fast);
        if (BIT_reloadDStream(&bitD) == BIT_DStream_overflow) { *op++ = FSE_decodeSymbol(&state1, &bitD, fast); break; }
    }
    return (size_t)(op - ostart);
}

/* FseDecompress.cs:334 FSE_decompress_wksp_body */
static size_t FSE_decompress_wksp(BYTE* dst, size_t dstCapacity, const void* cSrc, size_t cSrcSize, unsigned maxLog)
{
    const BYTE* ip = (const BYTE*)cSrc;
    S16 ncount[256]; unsigned tableLog; unsigned maxSymbolValue = 255;
    static __thread FSE_DTable dtable;
    {   size_t const NCountLength = FSE_readNCount(ncount, &maxSymbolValue, &tableLog, ip, cSrcSize);
        if (ERR_isError(NCountLength)) return NCountLength;
        if (tableLog > maxLog) return ERROR(tableLog_tooLarge);
        ip += NCountLength; cSrcSize -= NCountLength;
    }
    CHECK_F(FSE_buildDTable(&dtable, ncount, maxSymbolValue, tableLog));
    return FSE_decompress_usingDTable(dst, dstCapacity, ip, cSrcSize, &dtable);
}

/* =====================================================================================
 *  Huffman: weights (EntropyCommon.cs:292 HUF_readStats_body), X1 table (HufDecompress.cs:80),
 *  single-symbol stream decoding (HufDecompress.cs:254-537)
 * ===================================================================================== */
typedef struct { BYTE nbBits; BYTE byte; } HUF_DEltX1;
typedef struct { U16 sequence; BYTE nbBits; BYTE length; } HUF_DEltX2;   /* HUF_DEltX2.cs */
typedef struct { unsigned tableLog; unsigned tableType; HUF_DEltX1 dt[1 << 12]; HUF_DEltX2 dt2[1 << 12]; } HUF_DTable;

static size_t HUF_readStats(BYTE* huffWeight, size_t hwSize, U32* rankStats, U32* nbSymbolsPtr, U32* tableLogPtr, const void* src, size_t srcSize)
{
    U32 weightTotal; const BYTE* ip = (const BYTE*)src; size_t iSize, oSize;
    if (!srcSize) return ERROR(srcSize_wrong);
    iSize = ip[0];
    if (iSize >= 128) {     /* raw 4-bit weights */
        oSize = iSize - 127;
        iSize = (oSize + 1) / 2;
        if (iSize + 1 > srcSize) return ERROR(srcSize_wrong);
        if (oSize >= hwSize) return ERROR(corruption_detected);
        ip += 1;
        {   U32 n; for (n = 0; n < oSize; n += 2) { huffWeight[n] = ip[n / 2] >> 4; huffWeight[n + 1] = ip[n / 2] & 15; } }
    } else {                /* FSE-compressed weights */
        if (iSize + 1 > srcSize) return ERROR(srcSize_wrong);
        oSize = FSE_decompress_wksp(huffWeight, hwSize - 1, ip + 1, iSize, 6);
        if (ERR_isError(oSize)) return oSize;
    }
    memset(rankStats, 0, (HUF_TABLELOG_MAX + 1) * sizeof(U32));
    weightTotal = 0;
    {   U32 n; for (n = 0; n < oSize; n++) {
            if (huffWeight[n] > HUF_TABLELOG_MAX) return ERROR(corruption_detected);
            rankStats[huffWeight[n]]++;
            weightTotal += (1U << huffWeight[n]) >> 1;
    }   }
    if (weightTotal == 0) return ERROR(corruption_detected);
    {   U32 const tableLog = BIT_highbit32(weightTotal) + 1;
        if (tableLog > HUF_TABLELOG_MAX) return ERROR(corruption_detected);
        *tableLogPtr = tableLog;
        {   U32 const total = 1U << tableLog;
            U32 const rest = total - weightTotal;
            U32 const verif = 1U << BIT_highbit32(rest);
            U32 const lastWeight = BIT_highbit32(rest) + 1;
            if (verif != rest) return ERROR(corruption_detected);
            huffWeight[oSize] = (BYTE)lastWeight;
            rankStats[lastWeight]++;
    }   }
    if ((rankStats[1] < 2) || (rankStats[1] & 1)) return ERROR(corruption_detected);
    *nbSymbolsPtr = (U32)(oSize + 1);
    return iSize + 1;
}

/* HufDecompress.cs:80 HUF_readDTableX1_wksp_bmi2 : the table is always rescaled to log 11 (:102-105) */
static size_t HUF_readDTableX1(HUF_DTable* DTable, const void* src, size_t srcSize)
{
    BYTE huffWeight[HUF_SYMBOLVALUE_MAX + 2]; U32 rankVal[HUF_TABLELOG_MAX + 2]; U32 rankStart[HUF_TABLELOG_MAX + 2];
    U32 tableLog = 0, nbSymbols = 0;
    size_t const iSize = HUF_readStats(huffWeight, HUF_SYMBOLVALUE_MAX + 1, rankVal, &nbSymbols, &tableLog, src, srcSize);
    if (ERR_isError(iSize)) return iSize;
    {   U32 const maxTableLog = HUF_TABLELOG_MAX + 1;     /* dctx table is created with maxTableLog 12 */
        U32 const targetTableLog = maxTableLog < 11 ? maxTableLog : 11;
        if (tableLog < targetTableLog) {                   /* HUF_rescaleStats, HufDecompress.cs:44 */
            U32 const scale = targetTableLog - tableLog; U32 s;
            for (s = 0; s < nbSymbols; ++s) huffWeight[s] += (BYTE)((huffWeight[s] == 0) ? 0 : scale);
            for (s = targetTableLog; s > scale; --s) rankVal[s] = rankVal[s - scale];
            for (s = scale; s > 0; --s) rankVal[s] = 0;
            tableLog = targetTableLog;
        }
        if (tableLog > maxTableLog) return ERROR(tableLog_tooLarge);
        DTable->tableLog = tableLog; DTable->tableType = 0;
    }
    {   U32 n, nextRankStart = 0;
        for (n = 0; n < tableLog + 1; n++) { U32 const curr = nextRankStart; nextRankStart += rankVal[n]; rankStart[n] = curr; }
    }
    {   /* symbols sorted by (weight, symbol); each fills (1<<w)>>1 consecutive cells (HufDecompress.cs:131-252) */
        U32 w, s, uStart = 0;
        (void)rankStart;
        for (w = 1; w < tableLog + 1; ++w) {
            U32 const length = (1U << w) >> 1;
            BYTE const nbBits = (BYTE)(tableLog + 1 - w);
            for (s = 0; s < nbSymbols; ++s) {
                if (huffWeight[s] != w) continue;
                {   U32 u; for (u = 0; u < length; ++u) { DTable->dt[uStart + u].byte = (BYTE)s; DTable->dt[uStart + u].nbBits = nbBits; } }
                uStart += length;
            }
        }
    }
    return iSize;
}

static inline BYTE HUF_decodeSymbolX1(BIT_DStream_t* D, const HUF_DEltX1* dt, U32 dtLog)   /* HufDecompress.cs:254 */
{
    size_t const val = BIT_lookBitsFast(D, dtLog);
    BYTE const c = dt[val].byte;
    BIT_skipBits(D, dt[val].nbBits);
    return c;
}
static void HUF_decodeStreamX1(BYTE* p, BIT_DStream_t* bitD, BYTE* const pEnd, const HUF_DEltX1* dt, U32 dtLog)  /* :264 */
{
    if ((pEnd - p) > 3) {
        while ((BIT_reloadDStream(bitD) == BIT_DStream_unfinished) && (p < pEnd - 3)) {
            *p++ = HUF_decodeSymbolX1(bitD, dt, dtLog); *p++ = HUF_decodeSymbolX1(bitD, dt, dtLog);
            *p++ = HUF_decodeSymbolX1(bitD, dt, dtLog); *p++ = HUF_decodeSymbolX1(bitD, dt, dtLog);
        }
    } else {
        BIT_reloadDStream(bitD);
    }
    while (p < pEnd) *p++ = HUF_decodeSymbolX1(bitD, dt, dtLog);
}
static size_t HUF_decompress1X1_usingDTable(BYTE* dst, size_t dstSize, const void* cSrc, size_t cSrcSize, const HUF_DTable* DTable)  /* :312 */
{
    BIT_DStream_t bitD;
    CHECK_F(BIT_initDStream(&bitD, cSrc, cSrcSize));
    HUF_decodeStreamX1(dst, &bitD, dst + dstSize, DTable->dt, DTable->tableLog);
    if (!BIT_endOfDStream(&bitD)) return ERROR(corruption_detected);
    return dstSize;
}
/* HufDecompress.cs:342 HUF_decompress4X1_usingDTable_internal_body.  The reference interleaves the four
 * streams (:431-514); the symbols each stream yields, and the exact-consumption check (:526-533), are
 * the same as decoding the streams one after the other. */
static size_t HUF_decompress4X1_usingDTable(BYTE* dst, size_t dstSize, const void* cSrc, size_t cSrcSize, const HUF_DTable* DTable)
{
    if (cSrcSize < 10) return ERROR(corruption_detected);
    {   const BYTE* const istart = (const BYTE*)cSrc;
        BYTE* const ostart = dst; BYTE* const oend = ostart + dstSize;
        size_t const length1 = MEM_read16(istart), length2 = MEM_read16(istart + 2), length3 = MEM_read16(istart + 4);
        size_t const length4 = cSrcSize - (length1 + length2 + length3 + 6);
        const BYTE* const istart1 = istart + 6; const BYTE* const istart2 = istart1 + length1;
        const BYTE* const istart3 = istart2 + length2; const BYTE* const istart4 = istart3 + length3;
        size_t const segmentSize = (dstSize + 3) / 4;
        BYTE* const opStart2 = ostart + segmentSize; BYTE* const opStart3 = opStart2 + segmentSize; BYTE* const opStart4 = opStart3 + segmentSize;
        BIT_DStream_t b1, b2, b3, b4;
        if (length4 > cSrcSize) return ERROR(corruption_detected);
        if (opStart4 > oend) return ERROR(corruption_detected);
        CHECK_F(BIT_initDStream(&b1, istart1, length1));
        CHECK_F(BIT_initDStream(&b2, istart2, length2));
        CHECK_F(BIT_initDStream(&b3, istart3, length3));
        CHECK_F(BIT_initDStream(&b4, istart4, length4));
        HUF_decodeStreamX1(ostart, &b1, opStart2, DTable->dt, DTable->tableLog);
        HUF_decodeStreamX1(opStart2, &b2, opStart3, DTable->dt, DTable->tableLog);
        HUF_decodeStreamX1(opStart3, &b3, opStart4, DTable->dt, DTable->tableLog);
        HUF_decodeStreamX1(opStart4, &b4, oend, DTable->dt, DTable->tableLog);
        if (!(BIT_endOfDStream(&b1) & BIT_endOfDStream(&b2) & BIT_endOfDStream(&b3) & BIT_endOfDStream(&b4))) return ERROR(corruption_detected);
        return dstSize;
    }
}

/* ---- double-symbol decoder : HufDecompress.cs:652-1455 -------------------------------------------------------------
 * HUF_readDTableX2_wksp_bmi2 (:892-1010) builds, at log maxTableLog (11 when tableLog <= 11, else 12 : :934-937), a table
 * whose cell for an index holds the first symbol s1 (l1 bits) and, when the remaining maxTableLog - l1 bits decide a second
 * symbol s2 completely (l2 <= maxTableLog - l1, HUF_fillDTableX2 :848 / Level2 :789: the cells below rankVal[minWeight] are
 * the "skipped" single-symbol ones), both: {sequence = s1 | s2 << 8, nbBits = l1 + l2, length = 2}.  The canonical code
 * (cells ordered by weight, then symbol) is the one HUF_readDTableX1 lays out, so the cells are derived from that table. */
static size_t HUF_readDTableX2(HUF_DTable* DTable, const void* src, size_t srcSize)
{
    size_t const iSize = HUF_readDTableX1(DTable, src, srcSize);   /* same HUF_readStats verdicts; log 11 unless tableLog == 12 */
    if (ERR_isError(iSize)) return iSize;
    {   U32 const dtLog = DTable->tableLog; U32 const size = 1U << dtLog; U32 idx; U32 minBits = 32;
        for (idx = 0; idx < size; idx++) if (DTable->dt[idx].nbBits < minBits) minBits = DTable->dt[idx].nbBits;
        for (idx = 0; idx < size; idx++) {
            HUF_DEltX1 const e1 = DTable->dt[idx];
            HUF_DEltX2 e; e.sequence = e1.byte; e.nbBits = e1.nbBits; e.length = 1;
            if (dtLog - e1.nbBits >= minBits) {                       /* :870 enough room for a second symbol */
                HUF_DEltX1 const e2 = DTable->dt[(idx << e1.nbBits) & (size - 1)];
                if (e2.nbBits <= dtLog - e1.nbBits) { e.sequence = (U16)(e1.byte | (e2.byte << 8)); e.nbBits = (BYTE)(e1.nbBits + e2.nbBits); e.length = 2; }
            }
            DTable->dt2[idx] = e;
        }
        DTable->tableType = 1;
    }
    return iSize;
}
static inline U32 HUF_decodeSymbolX2(BYTE* op, BIT_DStream_t* D, const HUF_DEltX2* dt, U32 dtLog)   /* :1012 */
{
    size_t const val = BIT_lookBitsFast(D, dtLog);
    op[0] = (BYTE)dt[val].sequence; op[1] = (BYTE)(dt[val].sequence >> 8);
    BIT_skipBits(D, dt[val].nbBits);
    return dt[val].length;
}
static inline U32 HUF_decodeLastSymbolX2(BYTE* op, BIT_DStream_t* D, const HUF_DEltX2* dt, U32 dtLog)   /* :1022 */
{
    size_t const val = BIT_lookBitsFast(D, dtLog);
    op[0] = (BYTE)dt[val].sequence;
    if (dt[val].length == 1) BIT_skipBits(D, dt[val].nbBits);
    else if (D->bitsConsumed < 64) {
        BIT_skipBits(D, dt[val].nbBits);
        if (D->bitsConsumed > 64) D->bitsConsumed = 64;      /* the reference's "ugly hack": accepted although the pair ran past the start */
    }
    return 1;
}
/* :1047.  Writes two bytes per step: the reference's literal buffer has room past dstSize (WILDCOPY_OVERLENGTH); here the
 * caller's buffer is litBuffer[ZSTD_BLOCKSIZE_MAX + 32]. */
static size_t HUF_decodeStreamX2(BYTE* p, BIT_DStream_t* bitD, BYTE* const pEnd, const HUF_DEltX2* dt, U32 dtLog)
{
    BYTE* const pStart = p;
    if ((size_t)(pEnd - p) >= sizeof(size_t)) {
        if (dtLog <= 11) {
            while ((BIT_reloadDStream(bitD) == BIT_DStream_unfinished) && (p < pEnd - 9)) {
                p += HUF_decodeSymbolX2(p, bitD, dt, dtLog); p += HUF_decodeSymbolX2(p, bitD, dt, dtLog); p += HUF_decodeSymbolX2(p, bitD, dt, dtLog);
                p += HUF_decodeSymbolX2(p, bitD, dt, dtLog); p += HUF_decodeSymbolX2(p, bitD, dt, dtLog);
            }
        } else {
            while ((BIT_reloadDStream(bitD) == BIT_DStream_unfinished) && (p < pEnd - (sizeof(size_t) - 1))) {
                p += HUF_decodeSymbolX2(p, bitD, dt, dtLog); p += HUF_decodeSymbolX2(p, bitD, dt, dtLog);
                p += HUF_decodeSymbolX2(p, bitD, dt, dtLog); p += HUF_decodeSymbolX2(p, bitD, dt, dtLog);
            }
        }
    } else {
        BIT_reloadDStream(bitD);
    }
    if ((size_t)(pEnd - p) >= 2) {
        while ((BIT_reloadDStream(bitD) == BIT_DStream_unfinished) && (p <= pEnd - 2)) p += HUF_decodeSymbolX2(p, bitD, dt, dtLog);
        while (p <= pEnd - 2) p += HUF_decodeSymbolX2(p, bitD, dt, dtLog);
    }
    if (p < pEnd) p += HUF_decodeLastSymbolX2(p, bitD, dt, dtLog);
    return (size_t)(p - pStart);
}
static size_t HUF_decompress1X2_usingDTable(BYTE* dst, size_t dstSize, const void* cSrc, size_t cSrcSize, const HUF_DTable* DTable)  /* :1114 */
{
    BIT_DStream_t bitD;
    CHECK_F(BIT_initDStream(&bitD, cSrc, cSrcSize));
    HUF_decodeStreamX2(dst, &bitD, dst + dstSize, DTable->dt2, DTable->tableLog);
    if (!BIT_endOfDStream(&bitD)) return ERROR(corruption_detected);
    return dstSize;
}
/* :1148 HUF_decompress4X2_usingDTable_internal_body : the interleaved main loop matters here (its stop condition looks at
 * stream 4 only and the other three may overshoot their segment on damaged input, :1322-1335), so it is restated as is. */
static size_t HUF_decompress4X2_usingDTable(BYTE* dst, size_t dstSize, const void* cSrc, size_t cSrcSize, const HUF_DTable* DTable)
{
    if (cSrcSize < 10) return ERROR(corruption_detected);
    {   const BYTE* const istart = (const BYTE*)cSrc;
        BYTE* const ostart = dst; BYTE* const oend = ostart + dstSize; BYTE* const olimit = oend - (sizeof(size_t) - 1);
        const HUF_DEltX2* const dt = DTable->dt2; U32 const dtLog = DTable->tableLog;
        size_t const length1 = MEM_read16(istart), length2 = MEM_read16(istart + 2), length3 = MEM_read16(istart + 4);
        size_t const length4 = cSrcSize - (length1 + length2 + length3 + 6);
        const BYTE* const istart1 = istart + 6; const BYTE* const istart2 = istart1 + length1;
        const BYTE* const istart3 = istart2 + length2; const BYTE* const istart4 = istart3 + length3;
        size_t const segmentSize = (dstSize + 3) / 4;
        BYTE* const opStart2 = ostart + segmentSize; BYTE* const opStart3 = opStart2 + segmentSize; BYTE* const opStart4 = opStart3 + segmentSize;
        BYTE* op1 = ostart; BYTE* op2 = opStart2; BYTE* op3 = opStart3; BYTE* op4 = opStart4;
        U32 endSignal = 1;
        BIT_DStream_t b1, b2, b3, b4;
        if (length4 > cSrcSize) return ERROR(corruption_detected);
        if (opStart4 > oend) return ERROR(corruption_detected);
        CHECK_F(BIT_initDStream(&b1, istart1, length1));
        CHECK_F(BIT_initDStream(&b2, istart2, length2));
        CHECK_F(BIT_initDStream(&b3, istart3, length3));
        CHECK_F(BIT_initDStream(&b4, istart4, length4));
        if ((size_t)(oend - op4) >= sizeof(size_t)) {
            for ( ; endSignal & (U32)(op4 < olimit); ) {
                int k;
                for (k = 0; k < 4; k++) op1 += HUF_decodeSymbolX2(op1, &b1, dt, dtLog);
                for (k = 0; k < 4; k++) op2 += HUF_decodeSymbolX2(op2, &b2, dt, dtLog);
                endSignal &= (BIT_reloadDStreamFast(&b1) == BIT_DStream_unfinished);
                endSignal &= (BIT_reloadDStreamFast(&b2) == BIT_DStream_unfinished);
                for (k = 0; k < 4; k++) op3 += HUF_decodeSymbolX2(op3, &b3, dt, dtLog);
                for (k = 0; k < 4; k++) op4 += HUF_decodeSymbolX2(op4, &b4, dt, dtLog);
                endSignal &= (BIT_reloadDStreamFast(&b3) == BIT_DStream_unfinished);
                endSignal &= (BIT_reloadDStreamFast(&b4) == BIT_DStream_unfinished);
            }
        }
        if (op1 > opStart2) return ERROR(corruption_detected);
        if (op2 > opStart3) return ERROR(corruption_detected);
        if (op3 > opStart4) return ERROR(corruption_detected);
        HUF_decodeStreamX2(op1, &b1, opStart2, dt, dtLog);
        HUF_decodeStreamX2(op2, &b2, opStart3, dt, dtLog);
        HUF_decodeStreamX2(op3, &b3, opStart4, dt, dtLog);
        HUF_decodeStreamX2(op4, &b4, oend, dt, dtLog);
        if (!(BIT_endOfDStream(&b1) & BIT_endOfDStream(&b2) & BIT_endOfDStream(&b3) & BIT_endOfDStream(&b4))) return ERROR(corruption_detected);
        return dstSize;
    }
}
/* :1688 HUF_selectDecoder, algoTime :1471-1681 ({tableTime, decode256Time} of the single- and the double-symbol decoder) */
static const U16 zo_algoTime[16][4] = {
    {0, 0, 1, 1}, {0, 0, 1, 1}, {150, 216, 381, 119}, {170, 205, 514, 112}, {177, 199, 539, 110}, {197, 194, 644, 107},
    {221, 192, 735, 107}, {256, 189, 881, 106}, {359, 188, 1167, 109}, {582, 187, 1570, 114}, {688, 187, 1712, 122},
    {825, 186, 1965, 136}, {976, 185, 2131, 150}, {1180, 186, 2070, 175}, {1377, 185, 1731, 202}, {1412, 185, 1695, 202} };
static U32 HUF_selectDecoder(size_t dstSize, size_t cSrcSize)
{
    U32 const Q = (cSrcSize >= dstSize) ? 15 : (U32)(cSrcSize * 16 / dstSize);
    U32 const D256 = (U32)(dstSize >> 8);
    U32 const DTime0 = zo_algoTime[Q][0] + zo_algoTime[Q][1] * D256;
    U32 DTime1 = zo_algoTime[Q][2] + zo_algoTime[Q][3] * D256;
    DTime1 += DTime1 >> 5;
    return DTime1 < DTime0;
}

/* =====================================================================================
 *  Sequence symbol tables -- ZstdDecompressBlock.cs:1551 (rle), :1571 ZSTD_buildFSETable_body,
 *  default tables :398/:857/:1092 (identical to building from the *_defaultNorm arrays)
 * ===================================================================================== */
typedef struct { U16 nextState; BYTE nbAdditionalBits; BYTE nbBits; U32 baseValue; } ZSTD_seqSymbol;   /* ZSTD_seqSymbol.cs */
typedef struct { unsigned tableLog; ZSTD_seqSymbol t[1 << 9]; } seqTable_t;

static void ZSTD_buildFSETable(seqTable_t* dt, const S16* norm, unsigned maxSymbolValue, const U32* baseValue, const BYTE* nbAdditionalBits, unsigned tableLog)
{
    BYTE cellSymbol[1 << 9]; U16 symbolNext[MaxSeq + 1]; unsigned fastMode; U32 u;
    U32 const tableSize = 1U << tableLog;
    dt->tableLog = tableLog;
    (void)zo_fse_spread(cellSymbol, symbolNext, &fastMode, norm, maxSymbolValue, tableLog);
    for (u = 0; u < tableSize; u++) {
        U32 const symbol = cellSymbol[u];
        U32 const nextState = symbolNext[symbol]++;
        dt->t[u].nbBits = (BYTE)(tableLog - BIT_highbit32(nextState));
        dt->t[u].nextState = (U16)((nextState << dt->t[u].nbBits) - tableSize);
        dt->t[u].nbAdditionalBits = nbAdditionalBits[symbol];
        dt->t[u].baseValue = baseValue[symbol];
    }
}
static void ZSTD_buildSeqTable_rle(seqTable_t* dt, U32 baseValue, BYTE nbAddBits)
{
    dt->tableLog = 0; dt->t[0].nbBits = 0; dt->t[0].nextState = 0; dt->t[0].nbAdditionalBits = nbAddBits; dt->t[0].baseValue = baseValue;
}

/* =====================================================================================
 *  Decoder context (the fields of ZSTD_DCtx_s the one-shot no-dict path uses)
 * ===================================================================================== */
typedef struct {
    seqTable_t LLTable, OFTable, MLTable;         /* entropy.LLTable/OFTable/MLTable */
    seqTable_t LLdef, OFdef, MLdef;               /* LL/OF/ML_defaultDTable */
    const seqTable_t *LLTptr, *OFTptr, *MLTptr;
    HUF_DTable hufTable;
    U32 rep[3];
    int litEntropy, fseEntropy;
    size_t seqOutLimit;      /* how far the sequences of the current block may write (zo_litBufferPlacement) */
    const BYTE* litPtr; size_t litSize;
    BYTE litBuffer[ZSTD_BLOCKSIZE_MAX + 32];
    const BYTE* prefixStart;
    /* dictionary (ZstdDecompress.cs:1752-1931): content = external segment that ends where the frame's output begins */
    const BYTE* dictEnd; size_t dictContentSize; U32 loadedDictID;
    /* frame params */
    U64 frameContentSize; U64 windowSize; U32 blockSizeMax; U32 dictID; U32 checksumFlag; U32 headerSize;
} zo_DCtx;

static void zo_decompressBegin(zo_DCtx* d)    /* ZstdDecompress.cs:1933 ZSTD_decompressBegin */
{
    d->litEntropy = d->fseEntropy = 0;
    memcpy(d->rep, repStartValue, sizeof(repStartValue));
    d->LLTptr = &d->LLTable; d->MLTptr = &d->MLTable; d->OFTptr = &d->OFTable;
    d->prefixStart = NULL;
    d->dictEnd = NULL; d->dictContentSize = 0; d->loadedDictID = 0;
}

/* ---- literals : ZstdDecompressBlock.cs:88 ZSTD_decodeLiteralsBlock ---- */
/* ZSTD_allocateLiteralsBuffer (ZstdDecompressBlock.cs:44-73) + `oend` of ZSTD_decompressSequences_body (:2668).  The reference keeps the
 * literals of a block inside dst when there is room (ZSTD_in_dst: at dst + ZSTD_BLOCKSIZE_MAX + WILDCOPY_OVERLENGTH); the sequences of
 * the block may then only write up to that address, so a damaged block that regenerates more than 128 KiB + 32 bytes fails with
 * dstSize_tooSmall before anything else is looked at.  On valid frames the placement is invisible.  This oracle keeps its literals in a
 * buffer of its own and restates the limit.  NOT restated: the ZSTD_split placement (litSize > 64 KiB with a tight dst), whose
 * "output caught up with the literal buffer" checks (:2139-2153) and literal clobbering only show on damaged frames. */
static void zo_litBufferPlacement(zo_DCtx* dctx, size_t dstCapacity, size_t litSize, int directReference)
{
    if (!directReference && dstCapacity > ZSTD_BLOCKSIZE_MAX + 32 + litSize + 32) dctx->seqOutLimit = ZSTD_BLOCKSIZE_MAX + 32;   /* ZSTD_in_dst */
    else dctx->seqOutLimit = dstCapacity;                                                                                          /* ZSTD_not_in_dst (and ZSTD_split) */
}

static size_t zo_decodeLiteralsBlock(zo_DCtx* dctx, const void* src, size_t srcSize, void* dst, size_t dstCapacity)
{
    if (srcSize < MIN_CBLOCK_SIZE) return ERROR(corruption_detected);
    {   const BYTE* const istart = (const BYTE*)src;
        symbolEncodingType_e const litEncType = (symbolEncodingType_e)(istart[0] & 3);
        size_t const expectedWriteSize = ZSTD_BLOCKSIZE_MAX < dstCapacity ? ZSTD_BLOCKSIZE_MAX : dstCapacity;
        switch (litEncType) {
        case set_repeat:
            if (dctx->litEntropy == 0) return ERROR(dictionary_corrupted);
            /* fall-through */
        case set_compressed:
            if (srcSize < 5) return ERROR(corruption_detected);
            {   size_t lhSize, litSize, litCSize; U32 singleStream = 0;
                U32 const lhlCode = (istart[0] >> 2) & 3;
                U32 const lhc = MEM_read32(istart);
                size_t hufSuccess;
                switch (lhlCode) {
                case 0: case 1: default:
                    singleStream = !lhlCode; lhSize = 3; litSize = (lhc >> 4) & 0x3FF; litCSize = (lhc >> 14) & 0x3FF; break;
                case 2: lhSize = 4; litSize = (lhc >> 4) & 0x3FFF; litCSize = lhc >> 18; break;
                case 3: lhSize = 5; litSize = (lhc >> 4) & 0x3FFFF; litCSize = (lhc >> 22) + ((size_t)istart[4] << 10); break;
                }
                if (litSize > 0 && dst == NULL) return ERROR(dstSize_tooSmall);
                if (litSize > ZSTD_BLOCKSIZE_MAX) return ERROR(corruption_detected);
                if (litCSize + lhSize > srcSize) return ERROR(corruption_detected);
                if (expectedWriteSize < litSize) return ERROR(dstSize_tooSmall);
                zo_litBufferPlacement(dctx, dstCapacity, litSize, 0);
                if (litEncType == set_repeat) {
                    /* HUF_decompress{1,4}X_usingDTable_bmi2 (HufDecompress.cs:1759, :1786): by the type the table was built with */
                    int const x2 = dctx->hufTable.tableType != 0;
                    hufSuccess = singleStream ? (x2 ? HUF_decompress1X2_usingDTable : HUF_decompress1X1_usingDTable)(dctx->litBuffer, litSize, istart + lhSize, litCSize, &dctx->hufTable)
                                              : (x2 ? HUF_decompress4X2_usingDTable : HUF_decompress4X1_usingDTable)(dctx->litBuffer, litSize, istart + lhSize, litCSize, &dctx->hufTable);
                } else {
                    /* HUF_decompress{1X1_DCtx,4X_hufOnly}_wksp_bmi2 : HufDecompress.cs:1793 / :1774.
                     * hufOnly rejects dstSize==0 and cSrcSize==0 up front. */
                    const BYTE* ip = istart + lhSize; size_t cSrcSize = litCSize;
                    if (!singleStream && litSize == 0) hufSuccess = ERROR(dstSize_tooSmall);
                    else if (!singleStream && cSrcSize == 0) hufSuccess = ERROR(corruption_detected);
                    else {
                        /* single stream: always the single-symbol decoder (ZstdDecompressBlock.cs:212); four streams: HUF_selectDecoder */
                        int const x2 = !singleStream && HUF_selectDecoder(litSize, cSrcSize);
                        size_t const hSize = x2 ? HUF_readDTableX2(&dctx->hufTable, ip, cSrcSize) : HUF_readDTableX1(&dctx->hufTable, ip, cSrcSize);
                        if (ERR_isError(hSize)) hufSuccess = hSize;
                        else if (hSize >= cSrcSize) hufSuccess = ERROR(srcSize_wrong);
                        else {
                            ip += hSize; cSrcSize -= hSize;
                            hufSuccess = singleStream ? HUF_decompress1X1_usingDTable(dctx->litBuffer, litSize, ip, cSrcSize, &dctx->hufTable)
                                       : x2 ? HUF_decompress4X2_usingDTable(dctx->litBuffer, litSize, ip, cSrcSize, &dctx->hufTable)
                                            : HUF_decompress4X1_usingDTable(dctx->litBuffer, litSize, ip, cSrcSize, &dctx->hufTable);
                        }
                    }
                }
                if (ERR_isError(hufSuccess)) return ERROR(corruption_detected);
                dctx->litPtr = dctx->litBuffer; dctx->litSize = litSize; dctx->litEntropy = 1;
                return litCSize + lhSize;
            }
        case set_basic:
            {   size_t litSize, lhSize; U32 const lhlCode = (istart[0] >> 2) & 3;
                switch (lhlCode) {
                case 0: case 2: default: lhSize = 1; litSize = istart[0] >> 3; break;
                case 1: lhSize = 2; litSize = MEM_read16(istart) >> 4; break;
                case 3: lhSize = 3; litSize = MEM_readLE24(istart) >> 4; break;
                }
                if (litSize > 0 && dst == NULL) return ERROR(dstSize_tooSmall);
                if (expectedWriteSize < litSize) return ERROR(dstSize_tooSmall);
                zo_litBufferPlacement(dctx, dstCapacity, litSize, lhSize + litSize + 32 <= srcSize);     /* enough src behind them: referenced in place (:268-299) */
                if (litSize + lhSize > srcSize) return ERROR(corruption_detected);
                dctx->litPtr = istart + lhSize; dctx->litSize = litSize;
                return lhSize + litSize;
            }
        case set_rle:
            {   U32 const lhlCode = (istart[0] >> 2) & 3; size_t litSize, lhSize;
                switch (lhlCode) {
                case 0: case 2: default: lhSize = 1; litSize = istart[0] >> 3; break;
                case 1: lhSize = 2; litSize = MEM_read16(istart) >> 4; break;
                case 3: lhSize = 3; litSize = MEM_readLE24(istart) >> 4;
                        if (srcSize < 4) return ERROR(corruption_detected);
                        break;
                }
                if (litSize > 0 && dst == NULL) return ERROR(dstSize_tooSmall);
                if (litSize > ZSTD_BLOCKSIZE_MAX) return ERROR(corruption_detected);
                if (expectedWriteSize < litSize) return ERROR(dstSize_tooSmall);
                zo_litBufferPlacement(dctx, dstCapacity, litSize, 0);
                memset(dctx->litBuffer, istart[lhSize], litSize);
                dctx->litPtr = dctx->litBuffer; dctx->litSize = litSize;
                return lhSize + 1;
            }
        default:
            return ERROR(corruption_detected);
        }
    }
}

/* ---- sequence headers : ZstdDecompressBlock.cs:1746 ZSTD_buildSeqTable, :1845 ZSTD_decodeSeqHeaders ---- */
static size_t zo_buildSeqTable(seqTable_t* DTableSpace, const seqTable_t** DTablePtr, symbolEncodingType_e type, unsigned max, U32 maxLog,
                               const void* src, size_t srcSize, const U32* baseValue, const BYTE* nbAdditionalBits,
                               const seqTable_t* defaultTable, U32 flagRepeatTable)
{
    switch (type) {
    case set_rle:
        if (!srcSize) return ERROR(srcSize_wrong);
        if ((*(const BYTE*)src) > max) return ERROR(corruption_detected);
        {   U32 const symbol = *(const BYTE*)src;
            ZSTD_buildSeqTable_rle(DTableSpace, baseValue[symbol], nbAdditionalBits[symbol]); }
        *DTablePtr = DTableSpace;
        return 1;
    case set_basic:
        *DTablePtr = defaultTable;
        return 0;
    case set_repeat:
        if (!flagRepeatTable) return ERROR(corruption_detected);
        return 0;
    case set_compressed:
        {   unsigned tableLog; S16 norm[MaxSeq + 1];
            size_t const headerSize = FSE_readNCount(norm, &max, &tableLog, src, srcSize);
            if (ERR_isError(headerSize)) return ERROR(corruption_detected);
            if (tableLog > maxLog) return ERROR(corruption_detected);
            ZSTD_buildFSETable(DTableSpace, norm, max, baseValue, nbAdditionalBits, tableLog);
            *DTablePtr = DTableSpace;
            return headerSize;
        }
    default:
        return ERROR(GENERIC);
    }
}

static size_t zo_decodeSeqHeaders(zo_DCtx* dctx, int* nbSeqPtr, const void* src, size_t srcSize)
{
    const BYTE* const istart = (const BYTE*)src; const BYTE* const iend = istart + srcSize; const BYTE* ip = istart;
    int nbSeq;
    if (srcSize < 1) return ERROR(srcSize_wrong);
    nbSeq = *ip++;
    if (!nbSeq) { *nbSeqPtr = 0; if (srcSize != 1) return ERROR(srcSize_wrong); return 1; }
    if (nbSeq > 0x7F) {
        if (nbSeq == 0xFF) { if (ip + 2 > iend) return ERROR(srcSize_wrong); nbSeq = MEM_read16(ip) + LONGNBSEQ; ip += 2; }
        else { if (ip >= iend) return ERROR(srcSize_wrong); nbSeq = ((nbSeq - 0x80) << 8) + *ip++; }
    }
    *nbSeqPtr = nbSeq;
    if (ip + 1 > iend) return ERROR(srcSize_wrong);
    {   symbolEncodingType_e const LLtype = (symbolEncodingType_e)(*ip >> 6);
        symbolEncodingType_e const OFtype = (symbolEncodingType_e)((*ip >> 4) & 3);
        symbolEncodingType_e const MLtype = (symbolEncodingType_e)((*ip >> 2) & 3);
        ip++;
        {   size_t const llhSize = zo_buildSeqTable(&dctx->LLTable, &dctx->LLTptr, LLtype, MaxLL, LLFSELog, ip, (size_t)(iend - ip), LL_base, LL_bits, &dctx->LLdef, (U32)dctx->fseEntropy);
            if (ERR_isError(llhSize)) return ERROR(corruption_detected);
            ip += llhSize; }
        {   size_t const ofhSize = zo_buildSeqTable(&dctx->OFTable, &dctx->OFTptr, OFtype, MaxOff, OffFSELog, ip, (size_t)(iend - ip), OF_base, OF_bits, &dctx->OFdef, (U32)dctx->fseEntropy);
            if (ERR_isError(ofhSize)) return ERROR(corruption_detected);
            ip += ofhSize; }
        {   size_t const mlhSize = zo_buildSeqTable(&dctx->MLTable, &dctx->MLTptr, MLtype, MaxML, MLFSELog, ip, (size_t)(iend - ip), ML_base, ML_bits, &dctx->MLdef, (U32)dctx->fseEntropy);
            if (ERR_isError(mlhSize)) return ERROR(corruption_detected);
            ip += mlhSize; }
    }
    return (size_t)(ip - istart);
}

/* ---- sequences : ZstdDecompressBlock.cs:2341-2485 (decode), :2075-2262 (exec), :2668 (loop) ---- */
typedef struct { size_t litLength, matchLength, offset; } seq_t;
typedef struct { size_t state; const ZSTD_seqSymbol* table; } ZSTD_fseState;
typedef struct { BIT_DStream_t DStream; ZSTD_fseState stateLL, stateOffb, stateML; size_t prevOffset[3]; } seqState_t;

static void ZSTD_initFseState(ZSTD_fseState* s, BIT_DStream_t* bitD, const seqTable_t* dt)
{ s->state = BIT_readBits(bitD, dt->tableLog); BIT_reloadDStream(bitD); s->table = dt->t; }
static inline void ZSTD_updateFseStateWithDInfo(ZSTD_fseState* s, BIT_DStream_t* bitD, U16 nextState, U32 nbBits)
{ size_t const lowBits = BIT_readBits(bitD, nbBits); s->state = nextState + lowBits; }

static seq_t ZSTD_decodeSequence(seqState_t* seqState)   /* 64-bit build, longOffsets == 0 */
{
    seq_t seq;
    const ZSTD_seqSymbol* const llDInfo = seqState->stateLL.table + seqState->stateLL.state;
    const ZSTD_seqSymbol* const mlDInfo = seqState->stateML.table + seqState->stateML.state;
    const ZSTD_seqSymbol* const ofDInfo = seqState->stateOffb.table + seqState->stateOffb.state;
    seq.matchLength = mlDInfo->baseValue;
    seq.litLength = llDInfo->baseValue;
    {   U32 const ofBase = ofDInfo->baseValue;
        BYTE const llBits = llDInfo->nbAdditionalBits, mlBits = mlDInfo->nbAdditionalBits, ofBits = ofDInfo->nbAdditionalBits;
        BYTE const totalBits = (BYTE)(llBits + mlBits + ofBits);
        U16 const llNext = llDInfo->nextState, mlNext = mlDInfo->nextState, ofNext = ofDInfo->nextState;
        U32 const llnbBits = llDInfo->nbBits, mlnbBits = mlDInfo->nbBits, ofnbBits = ofDInfo->nbBits;
        {   size_t offset;
            if (ofBits > 1) {
                offset = ofBase + BIT_readBitsFast(&seqState->DStream, ofBits);
                seqState->prevOffset[2] = seqState->prevOffset[1];
                seqState->prevOffset[1] = seqState->prevOffset[0];
                seqState->prevOffset[0] = offset;
            } else {
                U32 const ll0 = (llDInfo->baseValue == 0);
                if (ofBits == 0) {
                    offset = seqState->prevOffset[ll0];
                    seqState->prevOffset[1] = seqState->prevOffset[!ll0];
                    seqState->prevOffset[0] = offset;
                } else {
                    offset = ofBase + ll0 + BIT_readBitsFast(&seqState->DStream, 1);
                    {   size_t temp = (offset == 3) ? seqState->prevOffset[0] - 1 : seqState->prevOffset[offset];
                        temp += !temp;
                        if (offset != 1) seqState->prevOffset[2] = seqState->prevOffset[1];
                        seqState->prevOffset[1] = seqState->prevOffset[0];
                        seqState->prevOffset[0] = offset = temp;
            }   }   }
            seq.offset = offset;
        }
        if (mlBits > 0) seq.matchLength += BIT_readBitsFast(&seqState->DStream, mlBits);
        if (totalBits >= 57 - (LLFSELog + MLFSELog + OffFSELog)) BIT_reloadDStream(&seqState->DStream);
        if (llBits > 0) seq.litLength += BIT_readBitsFast(&seqState->DStream, llBits);
        ZSTD_updateFseStateWithDInfo(&seqState->stateLL, &seqState->DStream, llNext, llnbBits);
        ZSTD_updateFseStateWithDInfo(&seqState->stateML, &seqState->DStream, mlNext, mlnbBits);
        ZSTD_updateFseStateWithDInfo(&seqState->stateOffb, &seqState->DStream, ofNext, ofnbBits);
    }
    return seq;
}

/* ZSTD_execSequence / ZSTD_execSequenceEnd semantics (checks in the order of :2083-2103); the copy itself is a
 * byte-forward copy, which is what wildcopy/overlapCopy8 compute for src-before-dst overlaps. */
static size_t zo_execSequence(BYTE* op, BYTE* const oend, seq_t sequence, const BYTE** litPtr, const BYTE* const litLimit, const BYTE* const prefixStart,
                              const BYTE* const dictEnd, size_t const dictContentSize)
{
    BYTE* const oLitEnd = op + sequence.litLength;
    size_t const sequenceLength = sequence.litLength + sequence.matchLength;
    const BYTE* match = oLitEnd - sequence.offset;
    if (sequenceLength > (size_t)(oend - op)) return ERROR(dstSize_tooSmall);
    if (sequence.litLength > (size_t)(litLimit - *litPtr)) return ERROR(corruption_detected);
    memmove(op, *litPtr, sequence.litLength);
    *litPtr += sequence.litLength;
    if (sequence.offset > (size_t)(oLitEnd - prefixStart)) {
        /* offset beyond prefix -> go into the dictionary segment (:2232-2252); virtualStart = prefixStart - dictContentSize */
        size_t const back = sequence.offset - (size_t)(oLitEnd - prefixStart);      /* bytes before prefixStart */
        if (back > dictContentSize) return ERROR(corruption_detected);
        match = dictEnd - back;
        if (back >= sequence.matchLength) { memmove(oLitEnd, match, sequence.matchLength); return sequenceLength; }
        memmove(oLitEnd, match, back);                                              /* span dictionary end and prefix start */
        {   BYTE* o = oLitEnd + back; size_t const ml = sequence.matchLength - back; size_t i;
            match = prefixStart;
            for (i = 0; i < ml; i++) o[i] = match[i]; }
        return sequenceLength;
    }
    {   BYTE* o = oLitEnd; size_t const ml = sequence.matchLength;
        if (sequence.offset >= ml) memcpy(o, match, ml);               /* no overlap */
        else if (sequence.offset >= 8) { size_t i = 0; for (; i + 8 <= ml; i += 8) memcpy(o + i, match + i, 8); for (; i < ml; i++) o[i] = match[i]; }
        else { size_t i; for (i = 0; i < ml; i++) o[i] = match[i]; }    /* byte-forward copy = ZSTD_overlapCopy8 + wildcopy result */
    }
    return sequenceLength;
}

typedef struct { uint32_t* triples; size_t cap; size_t n; } zo_seqTap;

static size_t zo_decompressSequences(zo_DCtx* dctx, void* dst, size_t maxDstSize, const void* seqStart, size_t seqSize, int nbSeq, zo_seqTap* tap)
{
    const BYTE* ip = (const BYTE*)seqStart; const BYTE* const iend = ip + seqSize;
    BYTE* const ostart = (BYTE*)dst; BYTE* const oend = ostart + (maxDstSize < dctx->seqOutLimit ? maxDstSize : dctx->seqOutLimit); BYTE* op = ostart;
    const BYTE* litPtr = dctx->litPtr; const BYTE* const litEnd = litPtr + dctx->litSize;
    const BYTE* const prefixStart = dctx->prefixStart;
    if (nbSeq) {
        seqState_t seqState;
        dctx->fseEntropy = 1;
        { U32 i; for (i = 0; i < 3; i++) seqState.prevOffset[i] = dctx->rep[i]; }
        if (ERR_isError(BIT_initDStream(&seqState.DStream, ip, (size_t)(iend - ip)))) return ERROR(corruption_detected);
        ZSTD_initFseState(&seqState.stateLL, &seqState.DStream, dctx->LLTptr);
        ZSTD_initFseState(&seqState.stateOffb, &seqState.DStream, dctx->OFTptr);
        ZSTD_initFseState(&seqState.stateML, &seqState.DStream, dctx->MLTptr);
        for (;;) {
            seq_t const sequence = ZSTD_decodeSequence(&seqState);
            size_t const oneSeqSize = zo_execSequence(op, oend, sequence, &litPtr, litEnd, prefixStart, dctx->dictEnd, dctx->dictContentSize);
            if (tap && tap->n < tap->cap) { tap->triples[3 * tap->n] = (U32)sequence.litLength; tap->triples[3 * tap->n + 1] = (U32)sequence.matchLength; tap->triples[3 * tap->n + 2] = (U32)sequence.offset; }
            if (tap) tap->n++;
            if (ERR_isError(oneSeqSize)) return oneSeqSize;
            op += oneSeqSize;
            if (--nbSeq == 0) break;
            BIT_reloadDStream(&seqState.DStream);
        }
        if (BIT_reloadDStream(&seqState.DStream) < BIT_DStream_completed) return ERROR(corruption_detected);
        { U32 i; for (i = 0; i < 3; i++) dctx->rep[i] = (U32)seqState.prevOffset[i]; }
    }
    {   size_t const lastLLSize = (size_t)(litEnd - litPtr);
        if (lastLLSize > (size_t)(oend - op)) return ERROR(dstSize_tooSmall);
        if (op != NULL) { memmove(op, litPtr, lastLLSize); op += lastLLSize; }
    }
    return (size_t)(op - ostart);
}

/* ZstdDecompressBlock.cs:3090 ZSTD_decompressBlock_internal */
static size_t zo_decompressBlock_internal(zo_DCtx* dctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize, zo_seqTap* tap)
{
    const BYTE* ip = (const BYTE*)src;
    if (srcSize >= ZSTD_BLOCKSIZE_MAX) return ERROR(srcSize_wrong);
    {   size_t const litCSize = zo_decodeLiteralsBlock(dctx, src, srcSize, dst, dstCapacity);
        if (ERR_isError(litCSize)) return litCSize;
        ip += litCSize; srcSize -= litCSize; }
    {   int nbSeq;
        size_t const seqHSize = zo_decodeSeqHeaders(dctx, &nbSeq, ip, srcSize);
        if (ERR_isError(seqHSize)) return seqHSize;
        ip += seqHSize; srcSize -= seqHSize;
        if (dst == NULL && nbSeq > 0) return ERROR(dstSize_tooSmall);
        return zo_decompressSequences(dctx, dst, dstCapacity, ip, srcSize, nbSeq, tap);
    }
}

/* ---- frame layer : ZstdDecompress.cs:427-650 (header), :19-47 of ZstdDecompressBlock.cs (block header) ---- */
static const size_t ZSTD_did_fieldSize[4] = { 0, 1, 2, 4 };
static const size_t ZSTD_fcs_fieldSize[4] = { 0, 2, 4, 8 };

static size_t zo_frameHeaderSize(const void* src, size_t srcSize)   /* :427 */
{
    size_t const minInputSize = 5;
    if (srcSize < minInputSize) return ERROR(srcSize_wrong);
    {   BYTE const fhd = ((const BYTE*)src)[minInputSize - 1];
        U32 const dictID = fhd & 3; U32 const singleSegment = (fhd >> 5) & 1; U32 const fcsId = fhd >> 6;
        return minInputSize + !singleSegment + ZSTD_did_fieldSize[dictID] + ZSTD_fcs_fieldSize[fcsId] + (singleSegment && !fcsId);
    }
}

typedef struct { U64 frameContentSize; U64 windowSize; U32 blockSizeMax; int isSkippable; U32 headerSize; U32 dictID; U32 checksumFlag; } zo_frameHeader;

/* :462 ZSTD_getFrameHeader_advanced : 0 ok, >0 wanted size, or error */
static size_t zo_getFrameHeader(zo_frameHeader* zfh, const void* src, size_t srcSize)
{
    const BYTE* ip = (const BYTE*)src; size_t const minInputSize = 5;
    memset(zfh, 0, sizeof(*zfh));
    if (srcSize < minInputSize) return minInputSize;
    if (src == NULL) return ERROR(GENERIC);
    if (MEM_read32(src) != ZSTD_MAGICNUMBER) {
        if ((MEM_read32(src) & ZSTD_MAGIC_SKIPPABLE_MASK) == ZSTD_MAGIC_SKIPPABLE_START) {
            if (srcSize < 8) return 8;
            memset(zfh, 0, sizeof(*zfh));
            zfh->frameContentSize = MEM_read32((const char*)src + 4);
            zfh->isSkippable = 1;
            return 0;
        }
        return ERROR(prefix_unknown);
    }
    {   size_t const fhsize = zo_frameHeaderSize(src, srcSize);
        if (srcSize < fhsize) return fhsize;
        zfh->headerSize = (U32)fhsize; }
    {   BYTE const fhdByte = ip[minInputSize - 1]; size_t pos = minInputSize;
        U32 const dictIDSizeCode = fhdByte & 3; U32 const checksumFlag = (fhdByte >> 2) & 1;
        U32 const singleSegment = (fhdByte >> 5) & 1; U32 const fcsID = fhdByte >> 6;
        U64 windowSize = 0; U32 dictID = 0; U64 frameContentSize = ZSTD_CONTENTSIZE_UNKNOWN;
        if ((fhdByte & 0x08) != 0) return ERROR(frameParameter_unsupported);
        if (!singleSegment) {
            BYTE const wlByte = ip[pos++];
            U32 const windowLog = (wlByte >> 3) + 10;
            if (windowLog > 31) return ERROR(frameParameter_windowTooLarge);
            windowSize = (1ULL << windowLog);
            windowSize += (windowSize >> 3) * (wlByte & 7);
        }
        switch (dictIDSizeCode) {
        default: case 0: break;
        case 1: dictID = ip[pos]; pos++; break;
        case 2: dictID = MEM_read16(ip + pos); pos += 2; break;
        case 3: dictID = MEM_read32(ip + pos); pos += 4; break;
        }
        switch (fcsID) {
        default: case 0: if (singleSegment) frameContentSize = ip[pos]; break;
        case 1: frameContentSize = MEM_read16(ip + pos) + 256; break;
        case 2: frameContentSize = MEM_read32(ip + pos); break;
        case 3: frameContentSize = MEM_read64(ip + pos); break;
        }
        if (singleSegment) windowSize = frameContentSize;
        zfh->frameContentSize = frameContentSize; zfh->windowSize = windowSize;
        zfh->blockSizeMax = (U32)(windowSize < ZSTD_BLOCKSIZE_MAX ? windowSize : ZSTD_BLOCKSIZE_MAX);
        zfh->dictID = dictID; zfh->checksumFlag = checksumFlag;
    }
    return 0;
}

typedef struct { blockType_e blockType; U32 lastBlock; U32 origSize; } blockProperties_t;
static size_t zo_getcBlockSize(const void* src, size_t srcSize, blockProperties_t* bp)   /* ZstdDecompressBlock.cs:19 */
{
    if (srcSize < ZSTD_blockHeaderSize) return ERROR(srcSize_wrong);
    {   U32 const cBlockHeader = MEM_readLE24(src); U32 const cSize = cBlockHeader >> 3;
        bp->lastBlock = cBlockHeader & 1; bp->blockType = (blockType_e)((cBlockHeader >> 1) & 3); bp->origSize = cSize;
        if (bp->blockType == bt_rle) return 1;
        if (bp->blockType == bt_reserved) return ERROR(corruption_detected);
        return cSize;
    }
}

static size_t zo_readSkippableFrameSize(const void* src, size_t srcSize)   /* ZstdDecompress.cs:652 */
{
    size_t const skippableHeaderSize = 8; U32 sizeU32;
    if (srcSize < skippableHeaderSize) return ERROR(srcSize_wrong);
    sizeU32 = MEM_read32((const BYTE*)src + 4);
    if ((U32)(sizeU32 + skippableHeaderSize) < sizeU32) return ERROR(frameParameter_unsupported);
    {   size_t const skippableSize = skippableHeaderSize + sizeU32;
        if (skippableSize > srcSize) return ERROR(srcSize_wrong);
        return skippableSize; }
}

/* ZstdDecompress.cs:877 ZSTD_findFrameSizeInfo */
static size_t zo_findFrameSizeInfo(const void* src, size_t srcSize, unsigned long long* decompressedBound)
{
    *decompressedBound = 0;
    if ((srcSize >= 8) && (MEM_read32(src) & ZSTD_MAGIC_SKIPPABLE_MASK) == ZSTD_MAGIC_SKIPPABLE_START)
        return zo_readSkippableFrameSize(src, srcSize);
    {   const BYTE* ip = (const BYTE*)src; const BYTE* const ipstart = ip; size_t remainingSize = srcSize; size_t nbBlocks = 0;
        zo_frameHeader zfh;
        {   size_t const ret = zo_getFrameHeader(&zfh, src, srcSize);
            if (ERR_isError(ret)) { *decompressedBound = ZSTD_CONTENTSIZE_ERROR; return ret; }
            if (ret > 0) { *decompressedBound = ZSTD_CONTENTSIZE_ERROR; return ERROR(srcSize_wrong); } }
        ip += zfh.headerSize; remainingSize -= zfh.headerSize;
        for (;;) {
            blockProperties_t bp;
            size_t const cBlockSize = zo_getcBlockSize(ip, remainingSize, &bp);
            if (ERR_isError(cBlockSize)) { *decompressedBound = ZSTD_CONTENTSIZE_ERROR; return cBlockSize; }
            if (ZSTD_blockHeaderSize + cBlockSize > remainingSize) { *decompressedBound = ZSTD_CONTENTSIZE_ERROR; return ERROR(srcSize_wrong); }
            ip += ZSTD_blockHeaderSize + cBlockSize; remainingSize -= ZSTD_blockHeaderSize + cBlockSize; nbBlocks++;
            if (bp.lastBlock) break;
        }
        if (zfh.checksumFlag) {
            if (remainingSize < 4) { *decompressedBound = ZSTD_CONTENTSIZE_ERROR; return ERROR(srcSize_wrong); }
            ip += 4;
        }
        *decompressedBound = (zfh.frameContentSize != ZSTD_CONTENTSIZE_UNKNOWN) ? zfh.frameContentSize : (unsigned long long)nbBlocks * zfh.blockSizeMax;
        return (size_t)(ip - ipstart);
    }
}

size_t zo_findFrameCompressedSize(const void* src, size_t srcSize) { unsigned long long b; return zo_findFrameSizeInfo(src, srcSize, &b); }

unsigned long long zo_decompressBound(const void* src, size_t srcSize)   /* ZstdDecompress.cs:971 */
{
    unsigned long long bound = 0;
    while (srcSize > 0) {
        unsigned long long decompressedBound;
        size_t const compressedSize = zo_findFrameSizeInfo(src, srcSize, &decompressedBound);
        if (ERR_isError(compressedSize) || decompressedBound == ZSTD_CONTENTSIZE_ERROR) return ZSTD_CONTENTSIZE_ERROR;
        src = (const BYTE*)src + compressedSize; srcSize -= compressedSize; bound += decompressedBound;
    }
    return bound;
}

/* ZstdDecompress.cs:1062 ZSTD_decompressFrame */
static size_t zo_decompressFrame(zo_DCtx* dctx, void* dst, size_t dstCapacity, const void** srcPtr, size_t* srcSizePtr, zo_seqTap* tap, int stopAfterFirstBlock)
{
    const BYTE* const istart = (const BYTE*)(*srcPtr); const BYTE* ip = istart;
    BYTE* const ostart = (BYTE*)dst; BYTE* const oend = dstCapacity != 0 ? ostart + dstCapacity : ostart; BYTE* op = ostart;
    size_t remainingSrcSize = *srcSizePtr;
    zo_frameHeader zfh;
    if (remainingSrcSize < 6 + ZSTD_blockHeaderSize) return ERROR(srcSize_wrong);
    {   size_t const frameHeaderSize = zo_frameHeaderSize(ip, 5);
        if (ERR_isError(frameHeaderSize)) return frameHeaderSize;
        if (remainingSrcSize < frameHeaderSize + ZSTD_blockHeaderSize) return ERROR(srcSize_wrong);
        {   size_t const result = zo_getFrameHeader(&zfh, ip, frameHeaderSize);      /* :834 ZSTD_decodeFrameHeader */
            if (ERR_isError(result)) return result;
            if (result > 0) return ERROR(srcSize_wrong);
            if (zfh.dictID != 0 && dctx->loadedDictID != zfh.dictID) return ERROR(dictionary_wrong); }   /* :849 */
        ip += frameHeaderSize; remainingSrcSize -= frameHeaderSize;
    }
    for (;;) {
        size_t decodedSize; blockProperties_t bp;
        size_t const cBlockSize = zo_getcBlockSize(ip, remainingSrcSize, &bp);
        if (ERR_isError(cBlockSize)) return cBlockSize;
        ip += ZSTD_blockHeaderSize; remainingSrcSize -= ZSTD_blockHeaderSize;
        if (cBlockSize > remainingSrcSize) return ERROR(srcSize_wrong);
        switch (bp.blockType) {
        case bt_compressed: decodedSize = zo_decompressBlock_internal(dctx, op, (size_t)(oend - op), ip, cBlockSize, tap); break;
        case bt_raw:                                                         /* :1004 ZSTD_copyRawBlock */
            if (cBlockSize > (size_t)(oend - op)) decodedSize = ERROR(dstSize_tooSmall);
            else if (op == NULL) decodedSize = cBlockSize == 0 ? 0 : ERROR(dstBuffer_null);
            else { memcpy(op, ip, cBlockSize); decodedSize = cBlockSize; }
            break;
        case bt_rle:                                                         /* :1029 ZSTD_setRleBlock */
            if (bp.origSize > (size_t)(oend - op)) decodedSize = ERROR(dstSize_tooSmall);
            else if (op == NULL) decodedSize = bp.origSize == 0 ? 0 : ERROR(dstBuffer_null);
            else { memset(op, *ip, bp.origSize); decodedSize = bp.origSize; }
            break;
        case bt_reserved: default: return ERROR(corruption_detected);
        }
        if (ERR_isError(decodedSize)) return decodedSize;
        if (decodedSize != 0) op += decodedSize;
        ip += cBlockSize; remainingSrcSize -= cBlockSize;
        if (stopAfterFirstBlock) return (size_t)(op - ostart);
        if (bp.lastBlock) break;
    }
    if (zfh.frameContentSize != ZSTD_CONTENTSIZE_UNKNOWN) {
        if ((U64)(op - ostart) != zfh.frameContentSize) return ERROR(corruption_detected);
    }
    if (zfh.checksumFlag) {
        if (remainingSrcSize < 4) return ERROR(checksum_wrong);
        {   U32 const checkCalc = (U32)zo_xxh64(ostart, (size_t)(op - ostart), 0);
            U32 const checkRead = MEM_read32(ip);
            if (checkRead != checkCalc) return ERROR(checksum_wrong); }
        ip += 4; remainingSrcSize -= 4;
    }
    *srcPtr = ip; *srcSizePtr = remainingSrcSize;
    return (size_t)(op - ostart);
}

static void zo_initDCtx(zo_DCtx* d)
{
    ZSTD_buildFSETable(&d->LLdef, LL_defaultNorm, MaxLL, LL_base, LL_bits, LL_DEFAULTNORMLOG);
    ZSTD_buildFSETable(&d->OFdef, OF_defaultNorm, DefaultMaxOff, OF_base, OF_bits, OF_DEFAULTNORMLOG);
    ZSTD_buildFSETable(&d->MLdef, ML_defaultNorm, MaxML, ML_base, ML_bits, ML_DEFAULTNORMLOG);
}

#include <stdlib.h>

void* zo_createDCtx(void) { zo_DCtx* const d = (zo_DCtx*)malloc(sizeof(zo_DCtx)); if (d) zo_initDCtx(d); return d; }
void zo_freeDCtx(void* ctx) { free(ctx); }

/* ZstdDecompress.cs:1216 ZSTD_decompressMultiFrame (no dictionary) */
size_t zo_decompressDCtx(void* ctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize)
{
    void* const dststart = dst; int moreThan1Frame = 0;
    zo_DCtx* const dctx = (zo_DCtx*)ctx;
    if (!dctx) return ERROR(memory_allocation);
    while (srcSize >= 5) {
        {   U32 const magicNumber = MEM_read32(src);
            if ((magicNumber & ZSTD_MAGIC_SKIPPABLE_MASK) == ZSTD_MAGIC_SKIPPABLE_START) {
                size_t const skippableSize = zo_readSkippableFrameSize(src, srcSize);
                if (ERR_isError(skippableSize)) return skippableSize;
                src = (const BYTE*)src + skippableSize; srcSize -= skippableSize;
                continue;
        }   }
        zo_decompressBegin(dctx);
        dctx->prefixStart = (const BYTE*)dst;      /* ZSTD_checkContinuity, ZstdDecompressBlock.cs:3166 */
        {   size_t const res = zo_decompressFrame(dctx, dst, dstCapacity, &src, &srcSize, NULL, 0);
            if ((zo_getErrorCode(res) == ZO_error_prefix_unknown) && (moreThan1Frame == 1)) return ERROR(srcSize_wrong);
            if (ERR_isError(res)) return res;
            if (res != 0) dst = (BYTE*)dst + res;
            dstCapacity -= res;
        }
        moreThan1Frame = 1;
    }
    if (srcSize) return ERROR(srcSize_wrong);
    return (size_t)((BYTE*)dst - (BYTE*)dststart);
}

/* ZstdDecompress.cs:1770 ZSTD_loadDEntropy + :1880 ZSTD_decompress_insertDictionary + :1752 ZSTD_refDictContent, applied at every
 * frame start as ZSTD_decompressBegin_usingDict does (:1954). */
static size_t zo_insertDictionary(zo_DCtx* dctx, const void* dict, size_t dictSize)
{
    const BYTE* dictPtr = (const BYTE*)dict; const BYTE* const dEnd = dictPtr + dictSize;
    if (dictSize >= 8 && MEM_read32(dict) == 0xEC30A437U) {
        dctx->loadedDictID = MEM_read32(dictPtr + 4);
        dictPtr += 8;
        {   size_t const hSize = HUF_readDTableX2(&dctx->hufTable, dictPtr, (size_t)(dEnd - dictPtr));     /* ZstdDecompress.cs:1786 */
            if (ERR_isError(hSize)) return ERROR(dictionary_corrupted);
            dictPtr += hSize; }
        {   S16 norm[MaxOff + 1]; unsigned maxV = MaxOff, log;
            size_t const h = FSE_readNCount(norm, &maxV, &log, dictPtr, (size_t)(dEnd - dictPtr));
            if (ERR_isError(h) || maxV > MaxOff || log > OffFSELog) return ERROR(dictionary_corrupted);
            ZSTD_buildFSETable(&dctx->OFTable, norm, maxV, OF_base, OF_bits, log);
            dictPtr += h; }
        {   S16 norm[MaxML + 1]; unsigned maxV = MaxML, log;
            size_t const h = FSE_readNCount(norm, &maxV, &log, dictPtr, (size_t)(dEnd - dictPtr));
            if (ERR_isError(h) || maxV > MaxML || log > MLFSELog) return ERROR(dictionary_corrupted);
            ZSTD_buildFSETable(&dctx->MLTable, norm, maxV, ML_base, ML_bits, log);
            dictPtr += h; }
        {   S16 norm[MaxLL + 1]; unsigned maxV = MaxLL, log;
            size_t const h = FSE_readNCount(norm, &maxV, &log, dictPtr, (size_t)(dEnd - dictPtr));
            if (ERR_isError(h) || maxV > MaxLL || log > LLFSELog) return ERROR(dictionary_corrupted);
            ZSTD_buildFSETable(&dctx->LLTable, norm, maxV, LL_base, LL_bits, log);
            dictPtr += h; }
        if (dictPtr + 12 > dEnd) return ERROR(dictionary_corrupted);
        {   size_t const contentSize = (size_t)(dEnd - (dictPtr + 12)); int i;
            for (i = 0; i < 3; i++) {
                U32 const r = MEM_read32(dictPtr); dictPtr += 4;
                if (r == 0 || r > contentSize) return ERROR(dictionary_corrupted);
                dctx->rep[i] = r;
        }   }
        dctx->litEntropy = dctx->fseEntropy = 1;
    }
    dctx->dictEnd = dEnd; dctx->dictContentSize = (size_t)(dEnd - dictPtr);
    return 0;
}

/* Decompressor.Unwrap after LoadDictionary: ZSTD_decompress_usingDict -> ZSTD_decompressMultiFrame (:1216) with a dictionary */
size_t zo_decompress_usingDict(void* dst, size_t dstCapacity, const void* src, size_t srcSize, const void* dict, size_t dictSize)
{
    void* const dststart = dst; int moreThan1Frame = 0; size_t result = 0;
    zo_DCtx* const dctx = (zo_DCtx*)zo_createDCtx();
    if (!dctx) return ERROR(memory_allocation);
    while (srcSize >= 5) {
        {   U32 const magicNumber = MEM_read32(src);
            if ((magicNumber & ZSTD_MAGIC_SKIPPABLE_MASK) == ZSTD_MAGIC_SKIPPABLE_START) {
                size_t const skippableSize = zo_readSkippableFrameSize(src, srcSize);
                if (ERR_isError(skippableSize)) { result = skippableSize; goto done; }
                src = (const BYTE*)src + skippableSize; srcSize -= skippableSize;
                continue;
        }   }
        zo_decompressBegin(dctx);
        if (dict && dictSize) { size_t const r = zo_insertDictionary(dctx, dict, dictSize); if (ERR_isError(r)) { result = r; goto done; } }
        dctx->prefixStart = (const BYTE*)dst;
        {   size_t const res = zo_decompressFrame(dctx, dst, dstCapacity, &src, &srcSize, NULL, 0);
            if ((zo_getErrorCode(res) == ZO_error_prefix_unknown) && (moreThan1Frame == 1)) { result = ERROR(srcSize_wrong); goto done; }
            if (ERR_isError(res)) { result = res; goto done; }
            if (res != 0) dst = (BYTE*)dst + res;
            dstCapacity -= res;
        }
        moreThan1Frame = 1;
    }
    result = srcSize ? ERROR(srcSize_wrong) : (size_t)((BYTE*)dst - (BYTE*)dststart);
done:
    zo_freeDCtx(dctx);
    return result;
}

size_t zo_decompress(void* dst, size_t dstCapacity, const void* src, size_t srcSize)
{
    void* const d = zo_createDCtx(); size_t r;
    if (!d) return ERROR(memory_allocation);
    r = zo_decompressDCtx(d, dst, dstCapacity, src, srcSize);
    zo_freeDCtx(d);
    return r;
}

size_t zo_decode_first_block_stages(const void* src, size_t srcSize, uint8_t* lits, size_t litCapacity, size_t* litSize,
                                    uint32_t* triples, size_t seqCapacity)
{
    zo_DCtx* const dctx = (zo_DCtx*)malloc(sizeof(zo_DCtx));
    BYTE* const out = (BYTE*)malloc(ZSTD_BLOCKSIZE_MAX + 64);
    zo_seqTap tap; size_t res;
    if (!dctx || !out) { free(dctx); free(out); return ERROR(memory_allocation); }
    tap.triples = triples; tap.cap = seqCapacity; tap.n = 0;
    zo_initDCtx(dctx); zo_decompressBegin(dctx); dctx->prefixStart = out;
    dctx->litSize = 0; dctx->litPtr = dctx->litBuffer;
    res = zo_decompressFrame(dctx, out, ZSTD_BLOCKSIZE_MAX, &src, &srcSize, &tap, 1);
    if (!ERR_isError(res)) {
        *litSize = dctx->litSize;
        if (dctx->litSize <= litCapacity && dctx->litSize) memcpy(lits, dctx->litPtr, dctx->litSize);
        res = tap.n;
    }
    free(dctx); free(out);
    return res;
}

/* ---- error helpers : ErrorPrivate.cs:10-184, ZstdCommon.cs:26-40 ---- */
unsigned zo_isError(size_t code) { return ERR_isError(code); }
int zo_getErrorCode(size_t code) { if (!ERR_isError(code)) return 0; return (int)(0 - code); }
const char* zo_getErrorName(size_t code)
{
    switch (zo_getErrorCode(code)) {
    case ZO_error_no_error: return "No error detected";
    case ZO_error_GENERIC: return "Error (generic)";
    case ZO_error_prefix_unknown: return "Unknown frame descriptor";
    case ZO_error_version_unsupported: return "Version not supported";
    case ZO_error_frameParameter_unsupported: return "Unsupported frame parameter";
    case ZO_error_frameParameter_windowTooLarge: return "Frame requires too much memory for decoding";
    case ZO_error_corruption_detected: return "Corrupted block detected";
    case ZO_error_checksum_wrong: return "Restored data doesn't match checksum";
    case ZO_error_parameter_unsupported: return "Unsupported parameter";
    case ZO_error_parameter_outOfBound: return "Parameter is out of bound";
    case ZO_error_init_missing: return "Context should be init first";
    case ZO_error_memory_allocation: return "Allocation error : not enough memory";
    case ZO_error_workSpace_tooSmall: return "workSpace buffer is not large enough";
    case ZO_error_stage_wrong: return "Operation not authorized at current processing stage";
    case ZO_error_tableLog_tooLarge: return "tableLog requires too much memory : unsupported";
    case ZO_error_maxSymbolValue_tooLarge: return "Unsupported max Symbol Value : too large";
    case ZO_error_maxSymbolValue_tooSmall: return "Specified maxSymbolValue is too small";
    case ZO_error_dictionary_corrupted: return "Dictionary is corrupted";
    case ZO_error_dictionary_wrong: return "Dictionary mismatch";
    case ZO_error_dictionaryCreation_failed: return "Cannot create Dictionary from provided samples";
    case ZO_error_dstSize_tooSmall: return "Destination buffer is too small";
    case ZO_error_srcSize_wrong: return "Src size is incorrect";
    case ZO_error_dstBuffer_null: return "Operation on NULL destination buffer";
    case ZO_error_frameIndex_tooLarge: return "Frame index is too large";
    case ZO_error_seekableIO: return "An I/O error occurred when reading/seeking";
    case ZO_error_dstBuffer_wrong: return "Destination buffer is wrong";
    case ZO_error_srcBuffer_wrong: return "Source buffer is wrong";
    default: return "Unspecified error code";
    }
}

/* entropy-header readers shared with the encode side's dictionary loader (zo_encode.c, ZSTD_loadCEntropy) */
size_t zo_FSE_readNCount(S16* normalizedCounter, unsigned* maxSVPtr, unsigned* tableLogPtr, const void* headerBuffer, size_t hbSize)
{ return FSE_readNCount(normalizedCounter, maxSVPtr, tableLogPtr, headerBuffer, hbSize); }
size_t zo_HUF_readStats(BYTE* huffWeight, size_t hwSize, U32* rankStats, U32* nbSymbolsPtr, U32* tableLogPtr, const void* src, size_t srcSize)
{ return HUF_readStats(huffWeight, hwSize, rankStats, nbSymbolsPtr, tableLogPtr, src, srcSize); }
