/*
 * zo_encode.c -- CPU ORACLE, encode side (test infrastructure only; see zo_common.h).
 *
 * Restates, in plain C, the reference's one-shot compressor for strategies ZSTD_fast and ZSTD_dfast (levels 1..3, the
 * negative levels, level 4 where it is still ZSTD_dfast), without and with a loaded dictionary:
 *   Compressor.Wrap -> ZSTD_compress2 -> ZSTD_compressEnd -> ZSTD_compress_frameChunk
 *     -> ZSTD_compressBlock_internal -> ZSTD_buildSeqStore (ZSTD_compressBlock_fast / _doubleFast)
 *        -> ZSTD_entropyCompressSeqStore (ZSTD_compressLiterals / HUF_*, ZSTD_buildSequencesStatistics / FSE_*,
 *           ZSTD_encodeSequences)
 * Citations are file:line under /root/reference/src/ZstdSharp/Unsafe/.
 */
#include <stdlib.h>
#include "zo_common.h"
#include "zo.h"

/* =====================================================================================
 *  Parameters -- Clevels.cs:8 (ZSTD_defaultCParameters rows 0..3 of the 4 size tables),
 *  ZstdCompress.cs:7891 ZSTD_getCParams_internal, :2023 ZSTD_adjustCParams_internal
 * ===================================================================================== */
typedef struct { U32 windowLog, chainLog, hashLog, searchLog, minMatch, targetLength; int strategy; } cParams_t;

static const cParams_t zo_defaultCParameters[4][5] = {   /* rows 0..4; row 4 only where it is still ZSTD_dfast (else strategy 0 = not restated) */
    {   /* "default" - for any srcSize > 256 KB            Clevels.cs:13-43  */
        { 19, 12, 13, 1, 6, 1, ZSTD_fast }, { 19, 13, 14, 1, 7, 0, ZSTD_fast },
        { 20, 15, 16, 1, 6, 0, ZSTD_fast }, { 21, 16, 17, 1, 5, 0, ZSTD_dfast }, { 21, 18, 18, 1, 5, 0, ZSTD_dfast } },
    {   /* for srcSize <= 256 KB                           Clevels.cs:246-276 */
        { 18, 12, 13, 1, 5, 1, ZSTD_fast }, { 18, 13, 14, 1, 6, 0, ZSTD_fast },
        { 18, 14, 14, 1, 5, 0, ZSTD_dfast }, { 18, 16, 16, 1, 4, 0, ZSTD_dfast }, { 18, 16, 17, 3, 5, 2, 0 /* ZSTD_greedy */ } },
    {   /* for srcSize <= 128 KB                           Clevels.cs:480-510 */
        { 17, 12, 12, 1, 5, 1, ZSTD_fast }, { 17, 12, 13, 1, 6, 0, ZSTD_fast },
        { 17, 13, 15, 1, 5, 0, ZSTD_fast }, { 17, 15, 16, 2, 5, 0, ZSTD_dfast }, { 17, 17, 17, 2, 4, 0, ZSTD_dfast } },
    {   /* for srcSize <= 16 KB                            Clevels.cs:713-743 */
        { 14, 12, 13, 1, 5, 1, ZSTD_fast }, { 14, 14, 15, 1, 5, 0, ZSTD_fast },
        { 14, 14, 15, 1, 4, 0, ZSTD_fast }, { 14, 14, 15, 2, 4, 0, ZSTD_dfast }, { 14, 14, 14, 4, 4, 2, 0 /* ZSTD_greedy */ } },
};

static cParams_t zo_adjustCParams(cParams_t cPar, U64 srcSize)   /* ZstdCompress.cs:2023, dictSize == 0, mode noAttachDict */
{
    U64 const maxWindowResize = 1ULL << 30;
    if (srcSize < maxWindowResize) {
        U32 const tSize = (U32)srcSize;
        U32 const hashSizeMin = 1 << 6;
        U32 const srcLog = (tSize < hashSizeMin) ? 6 : BIT_highbit32(tSize - 1) + 1;
        if (cPar.windowLog > srcLog) cPar.windowLog = srcLog;
    }
    {   U32 const dictAndWindowLog = cPar.windowLog;            /* ZSTD_dictAndWindowLog with dictSize 0 (:1985) */
        U32 const cycleLog = cPar.chainLog;                    /* ZSTD_cycleLog: strategy < btlazy2 */
        if (cPar.hashLog > dictAndWindowLog + 1) cPar.hashLog = dictAndWindowLog + 1;
        if (cycleLog > dictAndWindowLog) cPar.chainLog -= (cycleLog - dictAndWindowLog);
    }
    if (cPar.windowLog < 10) cPar.windowLog = 10;
    return cPar;
}

static int zo_getCParams_internal(cParams_t* out, int level, U64 srcSize)
{
    U64 const rSize = srcSize;
    U32 const tableID = (rSize <= 256 * 1024) + (rSize <= 128 * 1024) + (rSize <= 16 * 1024);
    int row;
    if (level == 0) row = 3;                 /* ZSTD_CLEVEL_DEFAULT */
    else if (level < 0) row = 0;             /* "entry 0 is baseline for fast mode" (ZstdCompress.cs:7901) */
    else if (level > 4) return -1;           /* outside the restated scope (ZSTD_fast / ZSTD_dfast levels) */
    else row = level;
    {   cParams_t cp = zo_defaultCParameters[tableID][row];
        if (cp.strategy == 0) return -1;     /* level 4 is ZSTD_greedy for this size */
        if (level < 0) {                     /* acceleration factor (:7918-7923); ZSTD_minCLevel() = -(1 << 17) */
            int const clamped = level < -(1 << 17) ? -(1 << 17) : level;
            cp.targetLength = (U32)(-clamped);
        }
        *out = zo_adjustCParams(cp, srcSize);
    }
    /* ZSTD_getCParamsFromCCtxParams (:2156) re-applies ZSTD_adjustCParams_internal: idempotent */
    *out = zo_adjustCParams(*out, srcSize);
    return 0;
}

void zo_getCParams(int level, size_t srcSize, unsigned out[7])
{
    cParams_t c; memset(&c, 0, sizeof(c));
    if (zo_getCParams_internal(&c, level, srcSize)) { memset(out, 0, 7 * sizeof(unsigned)); return; }
    out[0] = c.windowLog; out[1] = c.chainLog; out[2] = c.hashLog; out[3] = c.searchLog; out[4] = c.minMatch; out[5] = c.targetLength; out[6] = (unsigned)c.strategy;
}

size_t zo_compressBound(size_t srcSize)   /* ZstdCompress.cs:19 */
{
    return srcSize + (srcSize >> 8) + ((srcSize < (128 << 10)) ? (((128 << 10) - srcSize) >> 11) : 0);
}

/* =====================================================================================
 *  Forward bit writer -- Bitstream.cs:68-170 (BIT_CStream_t) and HufCompress.cs:862-981 (HUF_CStream_t).
 *  Both append fields LSB-first; they differ only in flush timing, which is not output-visible.  The
 *  overflow rule is shared: flushes clamp ptr to endPtr = start+capacity-8, and close returns 0 when
 *  ptr >= endPtr (Bitstream.cs:160-170, HufCompress.cs:964-981).
 * ===================================================================================== */
typedef struct { U64 acc; unsigned n; BYTE* ptr; BYTE* start; BYTE* endPtr; int overflow; } bw_t;

static size_t bw_init(bw_t* b, void* start, size_t capacity)
{
    b->acc = 0; b->n = 0; b->start = (BYTE*)start; b->ptr = b->start; b->overflow = 0;
    if (capacity <= sizeof(U64)) return ERROR(dstSize_tooSmall);
    b->endPtr = b->start + capacity - sizeof(U64);
    return 0;
}
static inline void bw_flush(bw_t* b)
{
    while (b->n >= 8) {
        if (b->ptr >= b->endPtr) b->overflow = 1; else *b->ptr++ = (BYTE)b->acc;
        b->acc >>= 8; b->n -= 8;
    }
}
static inline void bw_add(bw_t* b, U64 value, unsigned nbBits)
{
    if (nbBits == 0) return;
    b->acc |= (value & ((1ULL << nbBits) - 1)) << b->n;
    b->n += nbBits;
    bw_flush(b);
}
static size_t bw_close(bw_t* b)
{
    bw_add(b, 1, 1);
    if (b->overflow || b->ptr >= b->endPtr) return 0;
    if (b->n > 0) { *b->ptr = (BYTE)b->acc; return (size_t)(b->ptr - b->start) + 1; }
    return (size_t)(b->ptr - b->start);
}

/* =====================================================================================
 *  Histograms -- Hist.cs:17 HIST_count_simple (all variants give the same counts)
 * ===================================================================================== */
static unsigned HIST_count(unsigned* count, unsigned* maxSymbolValuePtr, const void* src, size_t srcSize)
{
    const BYTE* ip = (const BYTE*)src; const BYTE* const end = ip + srcSize;
    unsigned maxSymbolValue = *maxSymbolValuePtr; unsigned largestCount = 0;
    memset(count, 0, (maxSymbolValue + 1) * sizeof(*count));
    if (srcSize == 0) { *maxSymbolValuePtr = 0; return 0; }
    while (ip < end) count[*ip++]++;
    while (!count[maxSymbolValue]) maxSymbolValue--;
    *maxSymbolValuePtr = maxSymbolValue;
    { U32 s; for (s = 0; s <= maxSymbolValue; s++) if (count[s] > largestCount) largestCount = count[s]; }
    return largestCount;
}

/* =====================================================================================
 *  FSE compression tables -- FseCompress.cs
 * ===================================================================================== */
typedef struct { int deltaFindState; U32 deltaNbBits; } FSE_symbolCompressionTransform;   /* FSE_symbolCompressionTransform.cs */
typedef struct { U32 tableLog; U32 maxSymbolValue; U16 stateTable[1 << 9]; FSE_symbolCompressionTransform symbolTT[256]; } FSE_CTable;

static unsigned FSE_minTableLog(size_t srcSize, unsigned maxSymbolValue)   /* FseCompress.cs:384 */
{
    U32 const minBitsSrc = BIT_highbit32((U32)srcSize) + 1;
    U32 const minBitsSymbols = BIT_highbit32(maxSymbolValue) + 2;
    return minBitsSrc < minBitsSymbols ? minBitsSrc : minBitsSymbols;
}
static unsigned FSE_optimalTableLog_internal(unsigned maxTableLog, size_t srcSize, unsigned maxSymbolValue, unsigned minus)   /* :397 */
{
    U32 const maxBitsSrc = BIT_highbit32((U32)(srcSize - 1)) - minus;
    U32 tableLog = maxTableLog;
    U32 const minBits = FSE_minTableLog(srcSize, maxSymbolValue);
    if (tableLog == 0) tableLog = 11;
    if (maxBitsSrc < tableLog) tableLog = maxBitsSrc;
    if (minBits > tableLog) tableLog = minBits;
    if (tableLog < 5) tableLog = 5;
    if (tableLog > 12) tableLog = 12;
    return tableLog;
}
static unsigned FSE_optimalTableLog(unsigned maxTableLog, size_t srcSize, unsigned maxSymbolValue)
{ return FSE_optimalTableLog_internal(maxTableLog, srcSize, maxSymbolValue, 2); }

static size_t FSE_normalizeM2(S16* norm, U32 tableLog, const unsigned* count, size_t total, U32 maxSymbolValue, S16 lowProbCount)   /* :443 */
{
    S16 const NOT_YET_ASSIGNED = -2;
    U32 s, distributed = 0, ToDistribute;
    U32 const lowThreshold = (U32)(total >> tableLog);
    U32 lowOne = (U32)((total * 3) >> (tableLog + 1));
    for (s = 0; s <= maxSymbolValue; s++) {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = lowProbCount; distributed++; total -= count[s]; continue; }
        if (count[s] <= lowOne) { norm[s] = 1; distributed++; total -= count[s]; continue; }
        norm[s] = NOT_YET_ASSIGNED;
    }
    ToDistribute = (1U << tableLog) - distributed;
    if (ToDistribute == 0) return 0;
    if ((total / ToDistribute) > lowOne) {
        lowOne = (U32)((total * 3) / (ToDistribute * 2));
        for (s = 0; s <= maxSymbolValue; s++) {
            if ((norm[s] == NOT_YET_ASSIGNED) && (count[s] <= lowOne)) { norm[s] = 1; distributed++; total -= count[s]; continue; }
        }
        ToDistribute = (1U << tableLog) - distributed;
    }
    if (distributed == maxSymbolValue + 1) {
        U32 maxV = 0, maxC = 0;
        for (s = 0; s <= maxSymbolValue; s++) if (count[s] > maxC) { maxV = s; maxC = count[s]; }
        norm[maxV] += (S16)ToDistribute;
        return 0;
    }
    if (total == 0) {
        for (s = 0; ToDistribute > 0; s = (s + 1) % (maxSymbolValue + 1)) if (norm[s] > 0) { ToDistribute--; norm[s]++; }
        return 0;
    }
    {   U64 const vStepLog = 62 - tableLog;
        U64 const mid = (1ULL << (vStepLog - 1)) - 1;
        U64 const rStep = ((((U64)1 << vStepLog) * ToDistribute) + mid) / (U32)total;
        U64 tmpTotal = mid;
        for (s = 0; s <= maxSymbolValue; s++) {
            if (norm[s] == NOT_YET_ASSIGNED) {
                U64 const end = tmpTotal + (count[s] * rStep);
                U32 const sStart = (U32)(tmpTotal >> vStepLog);
                U32 const sEnd = (U32)(end >> vStepLog);
                U32 const weight = sEnd - sStart;
                if (weight < 1) return ERROR(GENERIC);
                norm[s] = (S16)weight;
                tmpTotal = end;
    }   }   }
    return 0;
}

static size_t FSE_normalizeCount(S16* normalizedCounter, unsigned tableLog, const unsigned* count, size_t total, unsigned maxSymbolValue, unsigned useLowProbCount)   /* :574 */
{
    if (tableLog == 0) tableLog = 11;
    if (tableLog < 5) return ERROR(GENERIC);
    if (tableLog > 12) return ERROR(tableLog_tooLarge);
    if (tableLog < FSE_minTableLog(total, maxSymbolValue)) return ERROR(GENERIC);
    {   S16 const lowProbCount = useLowProbCount ? -1 : 1;
        U64 const scale = 62 - tableLog;
        U64 const step = ((U64)1 << 62) / (U32)total;
        U64 const vStep = 1ULL << (scale - 20);
        int stillToDistribute = 1 << tableLog;
        unsigned s, largest = 0; S16 largestP = 0;
        U32 const lowThreshold = (U32)(total >> tableLog);
        for (s = 0; s <= maxSymbolValue; s++) {
            if (count[s] == total) return 0;   /* rle special case */
            if (count[s] == 0) { normalizedCounter[s] = 0; continue; }
            if (count[s] <= lowThreshold) { normalizedCounter[s] = lowProbCount; stillToDistribute--; }
            else {
                S16 proba = (S16)((count[s] * step) >> scale);
                if (proba < 8) {
                    U64 const restToBeat = vStep * rtbTable[proba];
                    proba += (count[s] * step) - ((U64)proba << scale) > restToBeat;
                }
                if (proba > largestP) { largestP = proba; largest = s; }
                normalizedCounter[s] = proba;
                stillToDistribute -= proba;
        }   }
        if (-stillToDistribute >= (normalizedCounter[largest] >> 1)) {
            size_t const errorCode = FSE_normalizeM2(normalizedCounter, tableLog, count, total, maxSymbolValue, lowProbCount);
            if (ERR_isError(errorCode)) return errorCode;
        } else normalizedCounter[largest] += (S16)stillToDistribute;
    }
    return tableLog;
}

static size_t FSE_NCountWriteBound(unsigned maxSymbolValue, unsigned tableLog)
{ size_t const maxHeaderSize = (((maxSymbolValue + 1) * tableLog + 4 + 2) / 8) + 1 + 2; return maxSymbolValue ? maxHeaderSize : 512; }

static size_t FSE_writeNCount_generic(void* header, size_t headerBufferSize, const S16* normalizedCounter, unsigned maxSymbolValue, unsigned tableLog, unsigned writeIsSafe)   /* :203 */
{
    BYTE* const ostart = (BYTE*)header; BYTE* out = ostart; BYTE* const oend = ostart + headerBufferSize;
    int nbBits; const int tableSize = 1 << tableLog; int remaining, threshold;
    U32 bitStream = 0; int bitCount = 0; unsigned symbol = 0; unsigned const alphabetSize = maxSymbolValue + 1; int previousIs0 = 0;
    bitStream += (tableLog - 5) << bitCount; bitCount += 4;
    remaining = tableSize + 1; threshold = tableSize; nbBits = (int)tableLog + 1;
    while ((symbol < alphabetSize) && (remaining > 1)) {
        if (previousIs0) {
            unsigned start = symbol;
            while ((symbol < alphabetSize) && !normalizedCounter[symbol]) symbol++;
            if (symbol == alphabetSize) break;
            while (symbol >= start + 24) {
                start += 24;
                bitStream += 0xFFFFU << bitCount;
                if ((!writeIsSafe) && (out > oend - 2)) return ERROR(dstSize_tooSmall);
                out[0] = (BYTE)bitStream; out[1] = (BYTE)(bitStream >> 8); out += 2; bitStream >>= 16;
            }
            while (symbol >= start + 3) { start += 3; bitStream += 3U << bitCount; bitCount += 2; }
            bitStream += (symbol - start) << bitCount; bitCount += 2;
            if (bitCount > 16) {
                if ((!writeIsSafe) && (out > oend - 2)) return ERROR(dstSize_tooSmall);
                out[0] = (BYTE)bitStream; out[1] = (BYTE)(bitStream >> 8); out += 2; bitStream >>= 16; bitCount -= 16;
        }   }
        {   int count = normalizedCounter[symbol++];
            int const max = (2 * threshold - 1) - remaining;
            remaining -= count < 0 ? -count : count;
            count++;
            if (count >= threshold) count += max;
            bitStream += (U32)count << bitCount;
            bitCount += nbBits;
            bitCount -= (count < max);
            previousIs0 = (count == 1);
            if (remaining < 1) return ERROR(GENERIC);
            while (remaining < threshold) { nbBits--; threshold >>= 1; }
        }
        if (bitCount > 16) {
            if ((!writeIsSafe) && (out > oend - 2)) return ERROR(dstSize_tooSmall);
            out[0] = (BYTE)bitStream; out[1] = (BYTE)(bitStream >> 8); out += 2; bitStream >>= 16; bitCount -= 16;
    }   }
    if (remaining != 1) return ERROR(GENERIC);
    if ((!writeIsSafe) && (out > oend - 2)) return ERROR(dstSize_tooSmall);
    out[0] = (BYTE)bitStream; out[1] = (BYTE)(bitStream >> 8);
    out += (bitCount + 7) / 8;
    return (size_t)(out - ostart);
}
static size_t FSE_writeNCount(void* buffer, size_t bufferSize, const S16* normalizedCounter, unsigned maxSymbolValue, unsigned tableLog)
{
    if (tableLog > 12) return ERROR(tableLog_tooLarge);
    if (tableLog < 5) return ERROR(GENERIC);
    if (bufferSize < FSE_NCountWriteBound(maxSymbolValue, tableLog)) {
        /* unsafe variant writes 2 bytes at a time with bound checks; give it a shadow buffer so that the
         * 2-byte store at the tail cannot touch memory past bufferSize, then copy back what it produced */
        BYTE tmp[600]; size_t r;
        if (bufferSize > sizeof(tmp) - 8) bufferSize = sizeof(tmp) - 8;
        r = FSE_writeNCount_generic(tmp, bufferSize, normalizedCounter, maxSymbolValue, tableLog, 0);
        if (!ERR_isError(r)) memcpy(buffer, tmp, r);
        return r;
    }
    return FSE_writeNCount_generic(buffer, bufferSize, normalizedCounter, maxSymbolValue, tableLog, 1);
}

/* FseCompress.cs:13 FSE_buildCTable_wksp */
static size_t FSE_buildCTable(FSE_CTable* ct, const S16* normalizedCounter, unsigned maxSymbolValue, unsigned tableLog)
{
    U32 const tableSize = 1U << tableLog; U32 const tableMask = tableSize - 1;
    U32 const step = (tableSize >> 1) + (tableSize >> 3) + 3;
    U32 const maxSV1 = maxSymbolValue + 1;
    U16 cumul[258]; BYTE tableSymbol[1 << 9];
    U32 highThreshold = tableSize - 1;
    ct->tableLog = tableLog; ct->maxSymbolValue = maxSymbolValue;
    {   U32 u; cumul[0] = 0;
        for (u = 1; u <= maxSV1; u++) {
            if (normalizedCounter[u - 1] == -1) { cumul[u] = (U16)(cumul[u - 1] + 1); tableSymbol[highThreshold--] = (BYTE)(u - 1); }
            else cumul[u] = (U16)(cumul[u - 1] + (U16)normalizedCounter[u - 1]);
        }
        cumul[maxSV1] = (U16)(tableSize + 1);
    }
    {   /* spread: the reference's fast path (:74-118) and slow path (:119-141) produce the same table */
        U32 position = 0, symbol;
        for (symbol = 0; symbol < maxSV1; symbol++) {
            int nbOccurrences; int const freq = normalizedCounter[symbol];
            for (nbOccurrences = 0; nbOccurrences < freq; nbOccurrences++) {
                tableSymbol[position] = (BYTE)symbol;
                position = (position + step) & tableMask;
                while (position > highThreshold) position = (position + step) & tableMask;
        }   }
    }
    {   U32 u; for (u = 0; u < tableSize; u++) { BYTE const s = tableSymbol[u]; ct->stateTable[cumul[s]++] = (U16)(tableSize + u); } }
    {   unsigned total = 0, s;
        for (s = 0; s <= maxSymbolValue; s++) {
            switch (normalizedCounter[s]) {
            case 0: ct->symbolTT[s].deltaNbBits = ((tableLog + 1) << 16) - (1 << tableLog); ct->symbolTT[s].deltaFindState = 0; break;
            case -1: case 1:
                ct->symbolTT[s].deltaNbBits = (tableLog << 16) - (1 << tableLog);
                ct->symbolTT[s].deltaFindState = (int)(total - 1); total++; break;
            default:
                {   U32 const maxBitsOut = tableLog - BIT_highbit32((U32)normalizedCounter[s] - 1);
                    U32 const minStatePlus = (U32)normalizedCounter[s] << maxBitsOut;
                    ct->symbolTT[s].deltaNbBits = (maxBitsOut << 16) - minStatePlus;
                    ct->symbolTT[s].deltaFindState = (int)(total - (unsigned)normalizedCounter[s]);
                    total += (unsigned)normalizedCounter[s];
    }   }   }   }
    return 0;
}
static void FSE_buildCTable_rle(FSE_CTable* ct, BYTE symbolValue)   /* :706 */
{
    ct->tableLog = 0; ct->maxSymbolValue = symbolValue; ct->stateTable[0] = 0; ct->stateTable[1] = 0;
    ct->symbolTT[symbolValue].deltaNbBits = 0; ct->symbolTT[symbolValue].deltaFindState = 0;
}

/* Fse.cs:10-96 */
typedef struct { ptrdiff_t value; const FSE_CTable* ct; unsigned stateLog; } FSE_CState_t;
static void FSE_initCState2(FSE_CState_t* st, const FSE_CTable* ct, U32 symbol)
{
    st->ct = ct; st->stateLog = ct->tableLog; st->value = (ptrdiff_t)1 << ct->tableLog;
    {   FSE_symbolCompressionTransform const tt = ct->symbolTT[symbol];
        U32 const nbBitsOut = (U32)((tt.deltaNbBits + (1 << 15)) >> 16);
        st->value = (ptrdiff_t)((nbBitsOut << 16) - tt.deltaNbBits);
        st->value = ct->stateTable[(st->value >> nbBitsOut) + tt.deltaFindState];
    }
}
static inline void FSE_encodeSymbol(bw_t* bitC, FSE_CState_t* st, unsigned symbol)
{
    FSE_symbolCompressionTransform const tt = st->ct->symbolTT[symbol];
    U32 const nbBitsOut = (U32)((st->value + tt.deltaNbBits) >> 16);
    bw_add(bitC, (U64)st->value, nbBitsOut);
    st->value = st->ct->stateTable[(st->value >> nbBitsOut) + tt.deltaFindState];
}
static inline void FSE_flushCState(bw_t* bitC, const FSE_CState_t* st) { bw_add(bitC, (U64)st->value, st->stateLog); }

/* FseCompress.cs:722 FSE_compress_usingCTable_generic (64-bit build) */
static size_t FSE_compress_usingCTable(void* dst, size_t dstSize, const void* src, size_t srcSize, const FSE_CTable* ct)
{
    const BYTE* const istart = (const BYTE*)src; const BYTE* const iend = istart + srcSize; const BYTE* ip = iend;
    bw_t bitC; FSE_CState_t CState1, CState2;
    if (srcSize <= 2) return 0;
    if (ERR_isError(bw_init(&bitC, dst, dstSize))) return 0;
    if (srcSize & 1) {
        FSE_initCState2(&CState1, ct, *--ip);
        FSE_initCState2(&CState2, ct, *--ip);
        FSE_encodeSymbol(&bitC, &CState1, *--ip);
    } else {
        FSE_initCState2(&CState2, ct, *--ip);
        FSE_initCState2(&CState1, ct, *--ip);
    }
    srcSize -= 2;
    if (srcSize & 2) { FSE_encodeSymbol(&bitC, &CState2, *--ip); FSE_encodeSymbol(&bitC, &CState1, *--ip); }
    while (ip > istart) {
        FSE_encodeSymbol(&bitC, &CState2, *--ip); FSE_encodeSymbol(&bitC, &CState1, *--ip);
        FSE_encodeSymbol(&bitC, &CState2, *--ip); FSE_encodeSymbol(&bitC, &CState1, *--ip);
    }
    FSE_flushCState(&bitC, &CState2);
    FSE_flushCState(&bitC, &CState1);
    return bw_close(&bitC);
}

/* =====================================================================================
 *  Huffman compression -- HufCompress.cs
 * ===================================================================================== */
typedef struct { BYTE nbBits[256]; U16 value[256]; U32 tableLog; } HUF_CTable;     /* HUF_CElt = nbBits | value<<(64-nbBits), :151-165 */
typedef enum { HUF_repeat_none = 0, HUF_repeat_check = 1, HUF_repeat_valid = 2 } HUF_repeat;
typedef struct { U32 count; U16 parent; BYTE byte; BYTE nbBits; } nodeElt;

/* HufCompress.cs:40 HUF_compressWeights */
static size_t HUF_compressWeights(void* dst, size_t dstSize, const BYTE* weightTable, size_t wtSize)
{
    BYTE* const ostart = (BYTE*)dst; BYTE* op = ostart; BYTE* const oend = ostart + dstSize;
    unsigned maxSymbolValue = 12; unsigned tableLog = 6;
    unsigned count[13]; S16 norm[13]; static __thread FSE_CTable CTable;
    if (wtSize <= 1) return 0;
    {   unsigned const maxCount = HIST_count(count, &maxSymbolValue, weightTable, wtSize);
        if (maxCount == wtSize) return 1;
        if (maxCount == 1) return 0; }
    tableLog = FSE_optimalTableLog(tableLog, wtSize, maxSymbolValue);
    CHECK_F(FSE_normalizeCount(norm, tableLog, count, wtSize, maxSymbolValue, 0));
    {   size_t const hSize = FSE_writeNCount(op, (size_t)(oend - op), norm, maxSymbolValue, tableLog);
        if (ERR_isError(hSize)) return hSize;
        op += hSize; }
    CHECK_F(FSE_buildCTable(&CTable, norm, maxSymbolValue, tableLog));
    {   size_t const cSize = FSE_compress_usingCTable(op, (size_t)(oend - op), weightTable, wtSize, &CTable);
        if (ERR_isError(cSize)) return cSize;
        if (cSize == 0) return 0;
        op += cSize; }
    return (size_t)(op - ostart);
}

/* HufCompress.cs:168 HUF_writeCTable_wksp */
static size_t HUF_writeCTable(void* dst, size_t maxDstSize, const HUF_CTable* CTable, unsigned maxSymbolValue, unsigned huffLog)
{
    BYTE bitsToWeight[HUF_TABLELOG_MAX + 1]; BYTE huffWeight[HUF_SYMBOLVALUE_MAX + 1];
    BYTE* op = (BYTE*)dst; U32 n;
    if (maxSymbolValue > HUF_SYMBOLVALUE_MAX) return ERROR(maxSymbolValue_tooLarge);
    bitsToWeight[0] = 0;
    for (n = 1; n < huffLog + 1; n++) bitsToWeight[n] = (BYTE)(huffLog + 1 - n);
    for (n = 0; n < maxSymbolValue; n++) huffWeight[n] = bitsToWeight[CTable->nbBits[n]];
    if (maxDstSize < 1) return ERROR(dstSize_tooSmall);
    {   size_t const hSize = HUF_compressWeights(op + 1, maxDstSize - 1, huffWeight, maxSymbolValue);
        if (ERR_isError(hSize)) return hSize;
        if ((hSize > 1) & (hSize < maxSymbolValue / 2)) { op[0] = (BYTE)hSize; return hSize + 1; } }
    if (maxSymbolValue > (256 - 128)) return ERROR(GENERIC);
    if (((maxSymbolValue + 1) / 2) + 1 > maxDstSize) return ERROR(dstSize_tooSmall);
    op[0] = (BYTE)(128 + (maxSymbolValue - 1));
    huffWeight[maxSymbolValue] = 0;
    for (n = 0; n < maxSymbolValue; n += 2) op[(n / 2) + 1] = (BYTE)((huffWeight[n] << 4) + huffWeight[n + 1]);
    return ((maxSymbolValue + 1) / 2) + 1;
}

/* HufCompress.cs:377 HUF_setMaxHeight */
static U32 HUF_setMaxHeight(nodeElt* huffNode, U32 lastNonNull, U32 maxNbBits)
{
    U32 const largestBits = huffNode[lastNonNull].nbBits;
    if (largestBits <= maxNbBits) return largestBits;
    {   int totalCost = 0;
        U32 const baseCost = 1U << (largestBits - maxNbBits);
        int n = (int)lastNonNull;
        while (huffNode[n].nbBits > maxNbBits) {
            totalCost += (int)(baseCost - (1U << (largestBits - huffNode[n].nbBits)));
            huffNode[n].nbBits = (BYTE)maxNbBits;
            n--;
        }
        while (huffNode[n].nbBits == maxNbBits) --n;
        totalCost >>= (largestBits - maxNbBits);
        {   U32 const noSymbol = 0xF0F0F0F0; U32 rankLast[HUF_TABLELOG_MAX + 2];
            memset(rankLast, 0xF0, sizeof(rankLast));
            {   U32 currentNbBits = maxNbBits; int pos;
                for (pos = n; pos >= 0; pos--) {
                    if (huffNode[pos].nbBits >= currentNbBits) continue;
                    currentNbBits = huffNode[pos].nbBits;
                    rankLast[maxNbBits - currentNbBits] = (U32)pos;
            }   }
            while (totalCost > 0) {
                U32 nBitsToDecrease = BIT_highbit32((U32)totalCost) + 1;
                for (; nBitsToDecrease > 1; nBitsToDecrease--) {
                    U32 const highPos = rankLast[nBitsToDecrease];
                    U32 const lowPos = rankLast[nBitsToDecrease - 1];
                    if (highPos == noSymbol) continue;
                    if (lowPos == noSymbol) break;
                    {   U32 const highTotal = huffNode[highPos].count;
                        U32 const lowTotal = 2 * huffNode[lowPos].count;
                        if (highTotal <= lowTotal) break;
                }   }
                while ((nBitsToDecrease <= HUF_TABLELOG_MAX) && (rankLast[nBitsToDecrease] == noSymbol)) nBitsToDecrease++;
                totalCost -= 1 << (nBitsToDecrease - 1);
                huffNode[rankLast[nBitsToDecrease]].nbBits++;
                if (rankLast[nBitsToDecrease - 1] == noSymbol) rankLast[nBitsToDecrease - 1] = rankLast[nBitsToDecrease];
                if (rankLast[nBitsToDecrease] == 0) rankLast[nBitsToDecrease] = noSymbol;
                else {
                    rankLast[nBitsToDecrease]--;
                    if (huffNode[rankLast[nBitsToDecrease]].nbBits != maxNbBits - nBitsToDecrease) rankLast[nBitsToDecrease] = noSymbol;
            }   }
            while (totalCost < 0) {
                if (rankLast[1] == noSymbol) {
                    while (huffNode[n].nbBits == maxNbBits) n--;
                    huffNode[n + 1].nbBits--;
                    rankLast[1] = (U32)(n + 1);
                    totalCost++;
                    continue;
                }
                huffNode[rankLast[1] + 1].nbBits--;
                rankLast[1]++;
                totalCost++;
    }   }   }
    return maxNbBits;
}

/* HufCompress.cs:520-687 HUF_getIndex / quicksort / HUF_sort */
#define RANK_POSITION_TABLE_SIZE 192
#define RANK_POSITION_LOG_BUCKETS_BEGIN ((RANK_POSITION_TABLE_SIZE - 1) - 32 - 1)   /* 158 */
#define RANK_POSITION_DISTINCT_COUNT_CUTOFF (RANK_POSITION_LOG_BUCKETS_BEGIN + 7)   /* + highbit32(158) = 165 */
static U32 HUF_getIndex(U32 const count)
{ return (count < RANK_POSITION_DISTINCT_COUNT_CUTOFF) ? count : BIT_highbit32(count) + RANK_POSITION_LOG_BUCKETS_BEGIN; }
static void HUF_swapNodes(nodeElt* a, nodeElt* b) { nodeElt tmp = *a; *a = *b; *b = tmp; }
static void HUF_insertionSort(nodeElt* huffNode, int const low, int const high)
{
    int i; int const size = high - low + 1;
    huffNode += low;
    for (i = 1; i < size; ++i) {
        nodeElt const key = huffNode[i]; int j = i - 1;
        while (j >= 0 && huffNode[j].count < key.count) { huffNode[j + 1] = huffNode[j]; j--; }
        huffNode[j + 1] = key;
    }
}
static int HUF_quickSortPartition(nodeElt* arr, int const low, int const high)
{
    U32 const pivot = arr[high].count; int i = low - 1; int j = low;
    for (; j < high; j++) if (arr[j].count > pivot) { i++; HUF_swapNodes(&arr[i], &arr[j]); }
    HUF_swapNodes(&arr[i + 1], &arr[high]);
    return i + 1;
}
static void HUF_simpleQuickSort(nodeElt* arr, int low, int high)
{
    int const kInsertionSortThreshold = 8;
    if (high - low < kInsertionSortThreshold) { HUF_insertionSort(arr, low, high); return; }
    while (low < high) {
        int const idx = HUF_quickSortPartition(arr, low, high);
        if (idx - low < high - idx) { HUF_simpleQuickSort(arr, low, idx - 1); low = idx + 1; }
        else { HUF_simpleQuickSort(arr, idx + 1, high); high = idx - 1; }
    }
}
typedef struct { U16 base; U16 curr; } rankPos;
static void HUF_sort(nodeElt* huffNode, const unsigned* count, U32 const maxSymbolValue, rankPos* rankPosition)
{
    U32 n; U32 const maxSymbolValue1 = maxSymbolValue + 1;
    memset(rankPosition, 0, sizeof(*rankPosition) * RANK_POSITION_TABLE_SIZE);
    for (n = 0; n < maxSymbolValue1; ++n) { U32 lowerRank = HUF_getIndex(count[n]); rankPosition[lowerRank].base++; }
    for (n = RANK_POSITION_TABLE_SIZE - 1; n > 0; --n) { rankPosition[n - 1].base += rankPosition[n].base; rankPosition[n - 1].curr = rankPosition[n - 1].base; }
    for (n = 0; n < maxSymbolValue1; ++n) {
        U32 const c = count[n]; U32 const r = HUF_getIndex(c) + 1; U32 const pos = rankPosition[r].curr++;
        huffNode[pos].count = c; huffNode[pos].byte = (BYTE)n;
    }
    for (n = RANK_POSITION_DISTINCT_COUNT_CUTOFF; n < RANK_POSITION_TABLE_SIZE - 1; ++n) {
        U32 const bucketSize = rankPosition[n].curr - rankPosition[n].base;
        U32 const bucketStartIdx = rankPosition[n].base;
        if (bucketSize > 1) HUF_simpleQuickSort(huffNode + bucketStartIdx, 0, (int)bucketSize - 1);
    }
}

#define STARTNODE (HUF_SYMBOLVALUE_MAX + 1)
static int HUF_buildTree(nodeElt* huffNode, U32 maxSymbolValue)   /* :689 */
{
    nodeElt* const huffNode0 = huffNode - 1;
    int nonNullRank, lowS, lowN; int nodeNb = STARTNODE; int n, nodeRoot;
    nonNullRank = (int)maxSymbolValue;
    while (huffNode[nonNullRank].count == 0) nonNullRank--;
    lowS = nonNullRank; nodeRoot = nodeNb + lowS - 1; lowN = nodeNb;
    huffNode[nodeNb].count = huffNode[lowS].count + huffNode[lowS - 1].count;
    huffNode[lowS].parent = huffNode[lowS - 1].parent = (U16)nodeNb;
    nodeNb++; lowS -= 2;
    for (n = nodeNb; n <= nodeRoot; n++) huffNode[n].count = (U32)(1U << 30);
    huffNode0[0].count = (U32)(1U << 31);
    while (nodeNb <= nodeRoot) {
        int const n1 = (huffNode[lowS].count < huffNode[lowN].count) ? lowS-- : lowN++;
        int const n2 = (huffNode[lowS].count < huffNode[lowN].count) ? lowS-- : lowN++;
        huffNode[nodeNb].count = huffNode[n1].count + huffNode[n2].count;
        huffNode[n1].parent = huffNode[n2].parent = (U16)nodeNb;
        nodeNb++;
    }
    huffNode[nodeRoot].nbBits = 0;
    for (n = nodeRoot - 1; n >= STARTNODE; n--) huffNode[n].nbBits = (BYTE)(huffNode[huffNode[n].parent].nbBits + 1);
    for (n = 0; n <= nonNullRank; n++) huffNode[n].nbBits = (BYTE)(huffNode[huffNode[n].parent].nbBits + 1);
    return nonNullRank;
}

static void HUF_buildCTableFromTree(HUF_CTable* CTable, const nodeElt* huffNode, int nonNullRank, U32 maxSymbolValue, U32 maxNbBits)   /* :750 */
{
    int n; U16 nbPerRank[HUF_TABLELOG_MAX + 1] = { 0 }; U16 valPerRank[HUF_TABLELOG_MAX + 1] = { 0 };
    int const alphabetSize = (int)(maxSymbolValue + 1);
    for (n = 0; n <= nonNullRank; n++) nbPerRank[huffNode[n].nbBits]++;
    {   U16 min = 0;
        for (n = (int)maxNbBits; n > 0; n--) { valPerRank[n] = min; min += nbPerRank[n]; min >>= 1; } }
    for (n = 0; n < alphabetSize; n++) CTable->nbBits[huffNode[n].byte] = huffNode[n].nbBits;
    for (n = 0; n < alphabetSize; n++) CTable->value[n] = CTable->nbBits[n] ? valPerRank[CTable->nbBits[n]]++ : 0;   /* HUF_setValue ignores nbBits==0 but still post-increments rank 0 */
    CTable->tableLog = maxNbBits;
}

static size_t HUF_buildCTable(HUF_CTable* CTable, const unsigned* count, U32 maxSymbolValue, U32 maxNbBits)   /* :790 HUF_buildCTable_wksp */
{
    nodeElt huffNodeTbl[2 * (HUF_SYMBOLVALUE_MAX + 1)]; rankPos rankPosition[RANK_POSITION_TABLE_SIZE];
    nodeElt* const huffNode0 = huffNodeTbl; nodeElt* const huffNode = huffNode0 + 1; int nonNullRank;
    if (maxNbBits == 0) maxNbBits = HUF_TABLELOG_DEFAULT;
    if (maxSymbolValue > HUF_SYMBOLVALUE_MAX) return ERROR(maxSymbolValue_tooLarge);
    memset(huffNode0, 0, sizeof(huffNodeTbl));
    memset(CTable, 0, sizeof(*CTable));
    HUF_sort(huffNode, count, maxSymbolValue, rankPosition);
    nonNullRank = HUF_buildTree(huffNode, maxSymbolValue);
    maxNbBits = HUF_setMaxHeight(huffNode, (U32)nonNullRank, maxNbBits);
    if (maxNbBits > HUF_TABLELOG_MAX) return ERROR(GENERIC);
    HUF_buildCTableFromTree(CTable, huffNode, nonNullRank, maxSymbolValue, maxNbBits);
    return maxNbBits;
}

static size_t HUF_estimateCompressedSize(const HUF_CTable* CTable, const unsigned* count, unsigned maxSymbolValue)
{ size_t nbBits = 0; int s; for (s = 0; s <= (int)maxSymbolValue; ++s) nbBits += (size_t)CTable->nbBits[s] * count[s]; return nbBits >> 3; }
static int HUF_validateCTable(const HUF_CTable* CTable, const unsigned* count, unsigned maxSymbolValue)
{ int bad = 0; int s; for (s = 0; s <= (int)maxSymbolValue; ++s) bad |= (count[s] != 0) & (CTable->nbBits[s] == 0); return !bad; }

/* HufCompress.cs:1056 HUF_compress1X_usingCTable_internal_body: symbols are appended last-to-first (:989-1054) */
static size_t HUF_compress1X_usingCTable(void* dst, size_t dstSize, const void* src, size_t srcSize, const HUF_CTable* CTable)
{
    const BYTE* ip = (const BYTE*)src; bw_t bitC; size_t n;
    if (dstSize < 8) return 0;
    if (ERR_isError(bw_init(&bitC, dst, dstSize))) return 0;
    for (n = srcSize; n > 0; n--) bw_add(&bitC, CTable->value[ip[n - 1]], CTable->nbBits[ip[n - 1]]);
    return bw_close(&bitC);
}
/* HufCompress.cs:1221 HUF_compress4X_usingCTable_internal */
static size_t HUF_compress4X_usingCTable(void* dst, size_t dstSize, const void* src, size_t srcSize, const HUF_CTable* CTable)
{
    size_t const segmentSize = (srcSize + 3) / 4;
    const BYTE* ip = (const BYTE*)src; const BYTE* const iend = ip + srcSize;
    BYTE* const ostart = (BYTE*)dst; BYTE* const oend = ostart + dstSize; BYTE* op = ostart; int i;
    if (dstSize < 6 + 1 + 1 + 1 + 8) return 0;
    if (srcSize < 12) return 0;
    op += 6;
    for (i = 0; i < 4; i++) {
        size_t const segLen = (i < 3) ? segmentSize : (size_t)(iend - ip);
        size_t const cSize = HUF_compress1X_usingCTable(op, (size_t)(oend - op), ip, segLen, CTable);
        if (ERR_isError(cSize)) return cSize;
        if (cSize == 0 || cSize > 65535) return 0;
        if (i < 3) MEM_write16(ostart + 2 * i, (U16)cSize);
        op += cSize; ip += segLen;
    }
    return (size_t)(op - ostart);
}
static size_t HUF_compressCTable_internal(BYTE* const ostart, BYTE* op, BYTE* const oend, const void* src, size_t srcSize, int singleStream, const HUF_CTable* CTable)   /* :1333 */
{
    size_t const cSize = singleStream ? HUF_compress1X_usingCTable(op, (size_t)(oend - op), src, srcSize, CTable)
                                      : HUF_compress4X_usingCTable(op, (size_t)(oend - op), src, srcSize, CTable);
    if (ERR_isError(cSize)) return cSize;
    if (cSize == 0) return 0;
    op += cSize;
    if ((size_t)(op - ostart) >= srcSize - 1) return 0;
    return (size_t)(op - ostart);
}

/* HufCompress.cs:1360 HUF_compress_internal (maxSymbolValue 255, huffLog 11) */
static size_t HUF_compress_internal(void* dst, size_t dstSize, const void* src, size_t srcSize, int singleStream,
                                    HUF_CTable* oldHufTable, HUF_repeat* repeat, int preferRepeat, unsigned suspectUncompressible)
{
    BYTE* const ostart = (BYTE*)dst; BYTE* const oend = ostart + dstSize; BYTE* op = ostart;
    unsigned count[HUF_SYMBOLVALUE_MAX + 1]; HUF_CTable CTable; unsigned maxSymbolValue = HUF_SYMBOLVALUE_MAX; unsigned huffLog = HUF_TABLELOG_DEFAULT;
    if (!srcSize) return 0;
    if (!dstSize) return 0;
    if (srcSize > ZSTD_BLOCKSIZE_MAX) return ERROR(srcSize_wrong);
    if (preferRepeat && repeat && *repeat == HUF_repeat_valid)
        return HUF_compressCTable_internal(ostart, op, oend, src, srcSize, singleStream, oldHufTable);
    if (suspectUncompressible && srcSize >= (4096 * 10)) {     /* SUSPECT_INCOMPRESSIBLE_SAMPLE_SIZE/RATIO, :1412-1446 */
        size_t largestTotal = 0;
        {   unsigned maxSymbolValueBegin = maxSymbolValue; largestTotal += HIST_count(count, &maxSymbolValueBegin, (const BYTE*)src, 4096); }
        {   unsigned maxSymbolValueEnd = maxSymbolValue; largestTotal += HIST_count(count, &maxSymbolValueEnd, (const BYTE*)src + srcSize - 4096, 4096); }
        if (largestTotal <= ((2 * 4096) >> 7) + 4) return 0;
    }
    {   size_t const largest = HIST_count(count, &maxSymbolValue, (const BYTE*)src, srcSize);
        if (largest == srcSize) { *ostart = ((const BYTE*)src)[0]; return 1; }
        if (largest <= (srcSize >> 7) + 4) return 0; }
    if (repeat && *repeat == HUF_repeat_check && !HUF_validateCTable(oldHufTable, count, maxSymbolValue)) *repeat = HUF_repeat_none;
    if (preferRepeat && repeat && *repeat != HUF_repeat_none)
        return HUF_compressCTable_internal(ostart, op, oend, src, srcSize, singleStream, oldHufTable);
    huffLog = FSE_optimalTableLog_internal(huffLog, srcSize, maxSymbolValue, 1);   /* HUF_optimalTableLog :12 */
    {   size_t const maxBits = HUF_buildCTable(&CTable, count, maxSymbolValue, huffLog);
        CHECK_F(maxBits);
        huffLog = (U32)maxBits; }
    {   size_t const hSize = HUF_writeCTable(op, dstSize, &CTable, maxSymbolValue, huffLog);
        if (ERR_isError(hSize)) return hSize;
        if (repeat && *repeat != HUF_repeat_none) {
            size_t const oldSize = HUF_estimateCompressedSize(oldHufTable, count, maxSymbolValue);
            size_t const newSize = HUF_estimateCompressedSize(&CTable, count, maxSymbolValue);
            if (oldSize <= hSize + newSize || hSize + 12 >= srcSize)
                return HUF_compressCTable_internal(ostart, op, oend, src, srcSize, singleStream, oldHufTable);
        }
        if (hSize + 12ul >= srcSize) return 0;
        op += hSize;
        if (repeat) *repeat = HUF_repeat_none;
        if (oldHufTable) memcpy(oldHufTable, &CTable, sizeof(CTable));
    }
    return HUF_compressCTable_internal(ostart, op, oend, src, srcSize, singleStream, &CTable);
}

/* =====================================================================================
 *  Literals section -- ZstdCompressLiterals.cs
 * ===================================================================================== */
typedef struct { HUF_CTable CTable; HUF_repeat repeatMode; } ZSTD_hufCTables_t;

static size_t ZSTD_minGain(size_t srcSize, int strat) { U32 const minlog = 6; (void)strat; return (srcSize >> minlog) + 2; }   /* ZstdCompressInternal.cs:137 (strat < btultra) */

static size_t ZSTD_noCompressLiterals(void* dst, size_t dstCapacity, const void* src, size_t srcSize)   /* :8 */
{
    BYTE* const ostart = (BYTE*)dst; U32 const flSize = 1 + (srcSize > 31) + (srcSize > 4095);
    if (srcSize + flSize > dstCapacity) return ERROR(dstSize_tooSmall);
    switch (flSize) {
    case 1: ostart[0] = (BYTE)((U32)set_basic + (srcSize << 3)); break;
    case 2: MEM_write16(ostart, (U16)((U32)set_basic + (1 << 2) + (srcSize << 4))); break;
    case 3: MEM_write32(ostart, (U32)((U32)set_basic + (3 << 2) + (srcSize << 4))); break;
    }
    memcpy(ostart + flSize, src, srcSize);
    return srcSize + flSize;
}
static size_t ZSTD_compressRleLiteralsBlock(void* dst, size_t dstCapacity, const void* src, size_t srcSize)   /* :49 */
{
    BYTE* const ostart = (BYTE*)dst; U32 const flSize = 1 + (srcSize > 31) + (srcSize > 4095);
    (void)dstCapacity;
    switch (flSize) {
    case 1: ostart[0] = (BYTE)((U32)set_rle + (srcSize << 3)); break;
    case 2: MEM_write16(ostart, (U16)((U32)set_rle + (1 << 2) + (srcSize << 4))); break;
    case 3: MEM_write32(ostart, (U32)((U32)set_rle + (3 << 2) + (srcSize << 4))); break;
    }
    ostart[flSize] = *(const BYTE*)src;
    return flSize + 1;
}
static size_t ZSTD_compressLiterals(const ZSTD_hufCTables_t* prevHuf, ZSTD_hufCTables_t* nextHuf, int strategy, int disableLiteralCompression,
                                    void* dst, size_t dstCapacity, const void* src, size_t srcSize, unsigned suspectUncompressible)   /* :86 */
{
    size_t const minGain = ZSTD_minGain(srcSize, strategy);
    size_t const lhSize = 3 + (srcSize >= 1024) + (srcSize >= 16384);
    BYTE* const ostart = (BYTE*)dst; U32 singleStream = srcSize < 256;
    symbolEncodingType_e hType = set_compressed; size_t cLitSize;
    memcpy(nextHuf, prevHuf, sizeof(*prevHuf));
    if (disableLiteralCompression) return ZSTD_noCompressLiterals(dst, dstCapacity, src, srcSize);      /* ZstdCompressLiterals.cs:100-101 */
    {   size_t const minLitSize = (prevHuf->repeatMode == HUF_repeat_valid) ? 6 : 63;
        if (srcSize <= minLitSize) return ZSTD_noCompressLiterals(dst, dstCapacity, src, srcSize); }
    if (dstCapacity < lhSize + 1) return ERROR(dstSize_tooSmall);
    {   HUF_repeat repeat = prevHuf->repeatMode;
        int const preferRepeat = strategy < ZSTD_lazy ? srcSize <= 1024 : 0;
        if (repeat == HUF_repeat_valid && lhSize == 3) singleStream = 1;
        cLitSize = HUF_compress_internal(ostart + lhSize, dstCapacity - lhSize, src, srcSize, (int)singleStream,
                                         &nextHuf->CTable, &repeat, preferRepeat, suspectUncompressible);
        if (repeat != HUF_repeat_none) hType = set_repeat;
    }
    if ((cLitSize == 0) || (cLitSize >= srcSize - minGain) || ERR_isError(cLitSize)) {
        memcpy(nextHuf, prevHuf, sizeof(*prevHuf));
        return ZSTD_noCompressLiterals(dst, dstCapacity, src, srcSize);
    }
    if (cLitSize == 1) {
        memcpy(nextHuf, prevHuf, sizeof(*prevHuf));
        return ZSTD_compressRleLiteralsBlock(dst, dstCapacity, src, srcSize);
    }
    if (hType == set_compressed) nextHuf->repeatMode = HUF_repeat_check;
    switch (lhSize) {
    case 3: { U32 const lhc = hType + ((!singleStream) << 2) + ((U32)srcSize << 4) + ((U32)cLitSize << 14); MEM_writeLE24(ostart, lhc); break; }
    case 4: { U32 const lhc = hType + (2 << 2) + ((U32)srcSize << 4) + ((U32)cLitSize << 18); MEM_write32(ostart, lhc); break; }
    case 5: { U32 const lhc = hType + (3 << 2) + ((U32)srcSize << 4) + ((U32)cLitSize << 22); MEM_write32(ostart, lhc); ostart[4] = (BYTE)(cLitSize >> 10); break; }
    }
    return lhSize + cLitSize;
}

/* =====================================================================================
 *  Sequences section -- ZstdCompressSequences.cs, ZstdCompress.cs:3069-3395
 * ===================================================================================== */
typedef enum { FSE_repeat_none = 0, FSE_repeat_check = 1, FSE_repeat_valid = 2 } FSE_repeat;
typedef struct { FSE_CTable offcodeCTable, matchlengthCTable, litlengthCTable; FSE_repeat offcode_repeatMode, matchlength_repeatMode, litlength_repeatMode; } ZSTD_fseCTables_t;
typedef struct { ZSTD_hufCTables_t huf; ZSTD_fseCTables_t fse; } ZSTD_entropyCTables_t;
typedef struct { ZSTD_entropyCTables_t entropy; U32 rep[3]; } ZSTD_compressedBlockState_t;

typedef struct {
    zo_seqDef* sequencesStart; zo_seqDef* sequences; BYTE* litStart; BYTE* lit;
    BYTE* llCode; BYTE* mlCode; BYTE* ofCode; size_t maxNbSeq; size_t maxNbLit;
    int longLengthType; U32 longLengthPos;
} seqStore_t;

static U32 ZSTD_LLcode(U32 litLength) { U32 const LL_deltaCode = 19; return (litLength > 63) ? BIT_highbit32(litLength) + LL_deltaCode : LL_Code[litLength]; }   /* ZstdCompressInternal.cs:20 */
static U32 ZSTD_MLcode(U32 mlBase) { U32 const ML_deltaCode = 36; return (mlBase > 127) ? BIT_highbit32(mlBase) + ML_deltaCode : ML_Code[mlBase]; }          /* :32 */

static void ZSTD_seqToCodes(const seqStore_t* seqStorePtr)   /* ZstdCompress.cs:3069 */
{
    const zo_seqDef* const sequences = seqStorePtr->sequencesStart;
    U32 const nbSeq = (U32)(seqStorePtr->sequences - seqStorePtr->sequencesStart); U32 u;
    for (u = 0; u < nbSeq; u++) {
        seqStorePtr->llCode[u] = (BYTE)ZSTD_LLcode(sequences[u].litLength);
        seqStorePtr->ofCode[u] = (BYTE)BIT_highbit32(sequences[u].offset);
        seqStorePtr->mlCode[u] = (BYTE)ZSTD_MLcode(sequences[u].matchLength);
    }
    if (seqStorePtr->longLengthType == 1) seqStorePtr->llCode[seqStorePtr->longLengthPos] = MaxLL;
    if (seqStorePtr->longLengthType == 2) seqStorePtr->mlCode[seqStorePtr->longLengthPos] = MaxML;
}

/* ZstdCompressSequences.cs:400 ZSTD_selectEncodingType, strategy < ZSTD_lazy branch only (fast/dfast) */
static symbolEncodingType_e ZSTD_selectEncodingType(FSE_repeat* repeatMode, size_t mostFrequent, size_t nbSeq, U32 defaultNormLog, int isDefaultAllowed, int strategy)
{
    if (mostFrequent == nbSeq) {
        *repeatMode = FSE_repeat_none;
        if (isDefaultAllowed && nbSeq <= 2) return set_basic;
        return set_rle;
    }
    if (isDefaultAllowed) {
        size_t const staticFse_nbSeq_max = 1000;
        size_t const mult = 10 - (size_t)strategy;
        size_t const baseLog = 3;
        size_t const dynamicFse_nbSeq_min = (((size_t)1 << defaultNormLog) * mult) >> baseLog;
        if ((*repeatMode == FSE_repeat_valid) && (nbSeq < staticFse_nbSeq_max)) return set_repeat;
        if ((nbSeq < dynamicFse_nbSeq_min) || (mostFrequent < (nbSeq >> (defaultNormLog - 1)))) { *repeatMode = FSE_repeat_none; return set_basic; }
    }
    *repeatMode = FSE_repeat_check;
    return set_compressed;
}

/* ZstdCompressSequences.cs:471 ZSTD_buildCTable */
static size_t ZSTD_buildCTable(void* dst, size_t dstCapacity, FSE_CTable* nextCTable, U32 FSELog, symbolEncodingType_e type, unsigned* count, U32 max,
                               const BYTE* codeTable, size_t nbSeq, const S16* defaultNorm, U32 defaultNormLog, U32 defaultMax, const FSE_CTable* prevCTable)
{
    BYTE* op = (BYTE*)dst; const BYTE* const oend = op + dstCapacity;
    switch (type) {
    case set_rle:
        FSE_buildCTable_rle(nextCTable, (BYTE)max);
        if (dstCapacity == 0) return ERROR(dstSize_tooSmall);
        *op = codeTable[0];
        return 1;
    case set_repeat:
        memcpy(nextCTable, prevCTable, sizeof(*prevCTable));
        return 0;
    case set_basic:
        CHECK_F(FSE_buildCTable(nextCTable, defaultNorm, defaultMax, defaultNormLog));
        return 0;
    case set_compressed: {
        S16 norm[MaxSeq + 1]; size_t nbSeq_1 = nbSeq;
        U32 const tableLog = FSE_optimalTableLog(FSELog, nbSeq, max);
        if (count[codeTable[nbSeq - 1]] > 1) { count[codeTable[nbSeq - 1]]--; nbSeq_1--; }
        CHECK_F(FSE_normalizeCount(norm, tableLog, count, nbSeq_1, max, nbSeq_1 >= 2048));   /* ZSTD_useLowProbCount :282 */
        {   size_t const NCountSize = FSE_writeNCount(op, (size_t)(oend - op), norm, max, tableLog);
            if (ERR_isError(NCountSize)) return NCountSize;
            CHECK_F(FSE_buildCTable(nextCTable, norm, max, tableLog));
            return NCountSize;
        }
    }
    default: return ERROR(GENERIC);
    }
}

/* ZstdCompressSequences.cs:585 ZSTD_encodeSequences_body (64-bit, longOffsets == 0: windowLog <= 57) */
static size_t ZSTD_encodeSequences(void* dst, size_t dstCapacity, const FSE_CTable* CTable_MatchLength, const BYTE* mlCodeTable,
                                   const FSE_CTable* CTable_OffsetBits, const BYTE* ofCodeTable, const FSE_CTable* CTable_LitLength, const BYTE* llCodeTable,
                                   const zo_seqDef* sequences, size_t nbSeq)
{
    bw_t blockStream; FSE_CState_t stateMatchLength, stateOffsetBits, stateLitLength;
    if (ERR_isError(bw_init(&blockStream, dst, dstCapacity))) return ERROR(dstSize_tooSmall);
    FSE_initCState2(&stateMatchLength, CTable_MatchLength, mlCodeTable[nbSeq - 1]);
    FSE_initCState2(&stateOffsetBits, CTable_OffsetBits, ofCodeTable[nbSeq - 1]);
    FSE_initCState2(&stateLitLength, CTable_LitLength, llCodeTable[nbSeq - 1]);
    bw_add(&blockStream, sequences[nbSeq - 1].litLength, LL_bits[llCodeTable[nbSeq - 1]]);
    bw_add(&blockStream, sequences[nbSeq - 1].matchLength, ML_bits[mlCodeTable[nbSeq - 1]]);
    bw_add(&blockStream, sequences[nbSeq - 1].offset, ofCodeTable[nbSeq - 1]);
    {   size_t n;
        for (n = nbSeq - 2; n < nbSeq; n--) {   /* intentional underflow */
            BYTE const llCode = llCodeTable[n]; BYTE const ofCode = ofCodeTable[n]; BYTE const mlCode = mlCodeTable[n];
            U32 const llBits = LL_bits[llCode]; U32 const ofBits = ofCode; U32 const mlBits = ML_bits[mlCode];
            FSE_encodeSymbol(&blockStream, &stateOffsetBits, ofCode);
            FSE_encodeSymbol(&blockStream, &stateMatchLength, mlCode);
            FSE_encodeSymbol(&blockStream, &stateLitLength, llCode);
            bw_add(&blockStream, sequences[n].litLength, llBits);
            bw_add(&blockStream, sequences[n].matchLength, mlBits);
            bw_add(&blockStream, sequences[n].offset, ofBits);
    }   }
    FSE_flushCState(&blockStream, &stateMatchLength);
    FSE_flushCState(&blockStream, &stateOffsetBits);
    FSE_flushCState(&blockStream, &stateLitLength);
    {   size_t const streamSize = bw_close(&blockStream);
        if (streamSize == 0) return ERROR(dstSize_tooSmall);
        return streamSize; }
}

/* ZstdCompress.cs:3127 ZSTD_buildSequencesStatistics + :3236 ZSTD_entropyCompressSeqStore_internal */
static size_t ZSTD_entropyCompressSeqStore_internal(seqStore_t* seqStorePtr, const ZSTD_entropyCTables_t* prevEntropy, ZSTD_entropyCTables_t* nextEntropy,
                                                    int strategy, int disableLit, void* dst, size_t dstCapacity)
{
    unsigned count[MaxSeq + 1];
    const zo_seqDef* const sequences = seqStorePtr->sequencesStart;
    size_t const nbSeq = (size_t)(seqStorePtr->sequences - seqStorePtr->sequencesStart);
    BYTE* const ostart = (BYTE*)dst; BYTE* const oend = ostart + dstCapacity; BYTE* op = ostart;
    size_t lastCountSize = 0;
    {   const BYTE* const literals = seqStorePtr->litStart;
        size_t const numLiterals = (size_t)(seqStorePtr->lit - seqStorePtr->litStart);
        unsigned const suspectUncompressible = (nbSeq == 0) || (numLiterals / nbSeq >= 20);    /* SUSPECT_UNCOMPRESSIBLE_LITERAL_RATIO */
        size_t const cSize = ZSTD_compressLiterals(&prevEntropy->huf, &nextEntropy->huf, strategy, disableLit, op, dstCapacity, literals, numLiterals, suspectUncompressible);
        if (ERR_isError(cSize)) return cSize;
        op += cSize;
    }
    if ((oend - op) < 3 + 1) return ERROR(dstSize_tooSmall);
    if (nbSeq < 128) *op++ = (BYTE)nbSeq;
    else if (nbSeq < LONGNBSEQ) { op[0] = (BYTE)((nbSeq >> 8) + 0x80); op[1] = (BYTE)nbSeq; op += 2; }
    else { op[0] = 0xFF; MEM_write16(op + 1, (U16)(nbSeq - LONGNBSEQ)); op += 3; }
    if (nbSeq == 0) { memcpy(&nextEntropy->fse, &prevEntropy->fse, sizeof(prevEntropy->fse)); return (size_t)(op - ostart); }
    {   BYTE* const seqHead = op++;
        U32 LLtype, Offtype, MLtype;
        ZSTD_seqToCodes(seqStorePtr);
        {   unsigned max = MaxLL;
            size_t const mostFrequent = HIST_count(count, &max, seqStorePtr->llCode, nbSeq);
            nextEntropy->fse.litlength_repeatMode = prevEntropy->fse.litlength_repeatMode;
            LLtype = ZSTD_selectEncodingType(&nextEntropy->fse.litlength_repeatMode, mostFrequent, nbSeq, LL_DEFAULTNORMLOG, 1, strategy);
            {   size_t const countSize = ZSTD_buildCTable(op, (size_t)(oend - op), &nextEntropy->fse.litlengthCTable, LLFSELog, (symbolEncodingType_e)LLtype, count, max,
                                                          seqStorePtr->llCode, nbSeq, LL_defaultNorm, LL_DEFAULTNORMLOG, MaxLL, &prevEntropy->fse.litlengthCTable);
                if (ERR_isError(countSize)) return countSize;
                if (LLtype == set_compressed) lastCountSize = countSize;
                op += countSize; } }
        {   unsigned max = MaxOff;
            size_t const mostFrequent = HIST_count(count, &max, seqStorePtr->ofCode, nbSeq);
            int const defaultAllowed = (max <= DefaultMaxOff);
            nextEntropy->fse.offcode_repeatMode = prevEntropy->fse.offcode_repeatMode;
            Offtype = ZSTD_selectEncodingType(&nextEntropy->fse.offcode_repeatMode, mostFrequent, nbSeq, OF_DEFAULTNORMLOG, defaultAllowed, strategy);
            {   size_t const countSize = ZSTD_buildCTable(op, (size_t)(oend - op), &nextEntropy->fse.offcodeCTable, OffFSELog, (symbolEncodingType_e)Offtype, count, max,
                                                          seqStorePtr->ofCode, nbSeq, OF_defaultNorm, OF_DEFAULTNORMLOG, DefaultMaxOff, &prevEntropy->fse.offcodeCTable);
                if (ERR_isError(countSize)) return countSize;
                if (Offtype == set_compressed) lastCountSize = countSize;
                op += countSize; } }
        {   unsigned max = MaxML;
            size_t const mostFrequent = HIST_count(count, &max, seqStorePtr->mlCode, nbSeq);
            nextEntropy->fse.matchlength_repeatMode = prevEntropy->fse.matchlength_repeatMode;
            MLtype = ZSTD_selectEncodingType(&nextEntropy->fse.matchlength_repeatMode, mostFrequent, nbSeq, ML_DEFAULTNORMLOG, 1, strategy);
            {   size_t const countSize = ZSTD_buildCTable(op, (size_t)(oend - op), &nextEntropy->fse.matchlengthCTable, MLFSELog, (symbolEncodingType_e)MLtype, count, max,
                                                          seqStorePtr->mlCode, nbSeq, ML_defaultNorm, ML_DEFAULTNORMLOG, MaxML, &prevEntropy->fse.matchlengthCTable);
                if (ERR_isError(countSize)) return countSize;
                if (MLtype == set_compressed) lastCountSize = countSize;
                op += countSize; } }
        *seqHead = (BYTE)((LLtype << 6) + (Offtype << 4) + (MLtype << 2));
    }
    {   size_t const bitstreamSize = ZSTD_encodeSequences(op, (size_t)(oend - op), &nextEntropy->fse.matchlengthCTable, seqStorePtr->mlCode,
                                                          &nextEntropy->fse.offcodeCTable, seqStorePtr->ofCode, &nextEntropy->fse.litlengthCTable, seqStorePtr->llCode, sequences, nbSeq);
        if (ERR_isError(bitstreamSize)) return bitstreamSize;
        op += bitstreamSize;
        if (lastCountSize && (lastCountSize + bitstreamSize) < 4) return 0;   /* :3346-3350 */
    }
    return (size_t)(op - ostart);
}

static size_t ZSTD_entropyCompressSeqStore(seqStore_t* seqStorePtr, const ZSTD_entropyCTables_t* prevEntropy, ZSTD_entropyCTables_t* nextEntropy,
                                           int strategy, int disableLit, void* dst, size_t dstCapacity, size_t srcSize)   /* :3357 */
{
    size_t const cSize = ZSTD_entropyCompressSeqStore_internal(seqStorePtr, prevEntropy, nextEntropy, strategy, disableLit, dst, dstCapacity);
    if (cSize == 0) return 0;
    if ((cSize == ERROR(dstSize_tooSmall)) & (srcSize <= dstCapacity)) return 0;
    if (ERR_isError(cSize)) return cSize;
    {   size_t const maxCSize = srcSize - ZSTD_minGain(srcSize, strategy);
        if (cSize >= maxCSize) return 0; }
    return cSize;
}

/* =====================================================================================
 *  Match finders -- ZstdFast.cs:96, ZstdDoubleFast.cs:51, helpers ZstdCompressInternal.cs:204-437
 * ===================================================================================== */
typedef struct { const BYTE* base; U32 dictLimit; U32 lowLimit; } ZSTD_window_t;
typedef struct ZSTD_matchState_s { ZSTD_window_t window; U32* hashTable; U32* chainTable; cParams_t cParams;
                                   U32 loadedDictEnd; const struct ZSTD_matchState_s* dictMatchState; U32 dictEndIndex; } ZSTD_matchState_t;   /* dictEndIndex: dms->window.nextSrc - base */

static const U32 prime4bytes = 2654435761U;
static const U64 prime5bytes = 889523592379ULL;
static const U64 prime6bytes = 227718039650203ULL;
static const U64 prime7bytes = 58295818150454627ULL;
static const U64 prime8bytes = 0xCF1BBCDCB7A56463ULL;
static size_t ZSTD_hashPtr(const void* p, U32 hBits, U32 mls)   /* ZstdCompressInternal.cs:340-437 */
{
    switch (mls) {
    default:
    case 4: return (MEM_read32(p) * prime4bytes) >> (32 - hBits);
    case 5: return (size_t)(((MEM_read64(p) << (64 - 40)) * prime5bytes) >> (64 - hBits));
    case 6: return (size_t)(((MEM_read64(p) << (64 - 48)) * prime6bytes) >> (64 - hBits));
    case 7: return (size_t)(((MEM_read64(p) << (64 - 56)) * prime7bytes) >> (64 - hBits));
    case 8: return (size_t)(((MEM_read64(p)) * prime8bytes) >> (64 - hBits));
    }
}
static size_t ZSTD_count(const BYTE* pIn, const BYTE* pMatch, const BYTE* const pInLimit)   /* :264 : common-prefix length bounded by pInLimit */
{
    const BYTE* const pStart = pIn;
    while (pIn + 8 <= pInLimit) {
        U64 const diff = MEM_read64(pMatch) ^ MEM_read64(pIn);
        if (diff) return (size_t)(pIn - pStart) + ((size_t)__builtin_ctzll(diff) >> 3);
        pIn += 8; pMatch += 8;
    }
    while (pIn < pInLimit && *pMatch == *pIn) { pIn++; pMatch++; }
    return (size_t)(pIn - pStart);
}
static void ZSTD_storeSeq(seqStore_t* seqStorePtr, size_t litLength, const BYTE* literals, U32 offCode, size_t mlBase)   /* :204 */
{
    memcpy(seqStorePtr->lit, literals, litLength);
    seqStorePtr->lit += litLength;
    if (litLength > 0xFFFF) { seqStorePtr->longLengthType = 1; seqStorePtr->longLengthPos = (U32)(seqStorePtr->sequences - seqStorePtr->sequencesStart); }
    seqStorePtr->sequences[0].litLength = (U16)litLength;
    seqStorePtr->sequences[0].offset = offCode + 1;
    if (mlBase > 0xFFFF) { seqStorePtr->longLengthType = 2; seqStorePtr->longLengthPos = (U32)(seqStorePtr->sequences - seqStorePtr->sequencesStart); }
    seqStorePtr->sequences[0].matchLength = (U16)mlBase;
    seqStorePtr->sequences++;
}
static U32 ZSTD_getLowestPrefixIndex(const ZSTD_matchState_t* ms, U32 curr, unsigned windowLog)   /* :802 */
{
    U32 const maxDistance = 1U << windowLog;
    U32 const lowestValid = ms->window.dictLimit;
    U32 const withinWindow = (curr - lowestValid > maxDistance) ? curr - maxDistance : lowestValid;
    return ms->loadedDictEnd != 0 ? lowestValid : withinWindow;
}
static U32 ZSTD_getLowestMatchIndex(const ZSTD_matchState_t* ms, U32 curr, unsigned windowLog)   /* :787 */
{
    U32 const maxDistance = 1U << windowLog;
    U32 const lowestValid = ms->window.lowLimit;
    U32 const withinWindow = (curr - lowestValid > maxDistance) ? curr - maxDistance : lowestValid;
    return ms->loadedDictEnd != 0 ? lowestValid : withinWindow;
}
#define ZSTD_REP_MOVE 2

/* ZstdFast.cs:96 ZSTD_compressBlock_fast_noDict_generic */
static size_t ZSTD_compressBlock_fast(ZSTD_matchState_t* ms, seqStore_t* seqStore, U32 rep[3], const void* src, size_t srcSize)
{
    const cParams_t* const cParams = &ms->cParams;
    U32* const hashTable = ms->hashTable; U32 const hlog = cParams->hashLog; U32 const mls = cParams->minMatch;
    size_t const stepSize = (cParams->targetLength > 1) ? (cParams->targetLength + !(cParams->targetLength) + 1) : 2;   /* hasStep = targetLength > 1 (:334) */
    const BYTE* const base = ms->window.base; const BYTE* const istart = (const BYTE*)src;
    U32 const endIndex = (U32)((size_t)(istart - base) + srcSize);
    U32 const prefixStartIndex = ZSTD_getLowestPrefixIndex(ms, endIndex, cParams->windowLog);
    const BYTE* const prefixStart = base + prefixStartIndex;
    const BYTE* const iend = istart + srcSize; const BYTE* const ilimit = iend - 8;
    const BYTE* anchor = istart; const BYTE* ip0 = istart; const BYTE* ip1; const BYTE* ip2; const BYTE* ip3;
    U32 current0; U32 rep_offset1 = rep[0]; U32 rep_offset2 = rep[1]; U32 offsetSaved = 0;
    size_t hash0, hash1; U32 idx, mval, offcode; const BYTE* match0; size_t mLength; size_t step; const BYTE* nextStep;
    size_t const kStepIncr = 1 << (8 - 1);

    ip0 += (ip0 == prefixStart);
    {   U32 const curr = (U32)(ip0 - base);
        U32 const windowLow = ZSTD_getLowestPrefixIndex(ms, curr, cParams->windowLog);
        U32 const maxRep = curr - windowLow;
        if (rep_offset2 > maxRep) { offsetSaved = rep_offset2; rep_offset2 = 0; }
        if (rep_offset1 > maxRep) { offsetSaved = rep_offset1; rep_offset1 = 0; }
    }
_start:
    step = stepSize; nextStep = ip0 + kStepIncr;
    ip1 = ip0 + 1; ip2 = ip0 + step; ip3 = ip2 + 1;
    if (ip3 >= ilimit) goto _cleanup;
    hash0 = ZSTD_hashPtr(ip0, hlog, mls); hash1 = ZSTD_hashPtr(ip1, hlog, mls);
    idx = hashTable[hash0];
    do {
        U32 const rval = MEM_read32(ip2 - rep_offset1);
        current0 = (U32)(ip0 - base);
        hashTable[hash0] = current0;
        if ((MEM_read32(ip2) == rval) & (rep_offset1 > 0)) {
            ip0 = ip2; match0 = ip0 - rep_offset1;
            mLength = ip0[-1] == match0[-1];
            ip0 -= mLength; match0 -= mLength;
            offcode = 0; mLength += 4;
            goto _match;
        }
        if (idx >= prefixStartIndex) mval = MEM_read32(base + idx); else mval = MEM_read32(ip0) ^ 1;
        if (MEM_read32(ip0) == mval) goto _offset;
        idx = hashTable[hash1];
        hash0 = hash1; hash1 = ZSTD_hashPtr(ip2, hlog, mls);
        ip0 = ip1; ip1 = ip2; ip2 = ip3;
        current0 = (U32)(ip0 - base);
        hashTable[hash0] = current0;
        if (idx >= prefixStartIndex) mval = MEM_read32(base + idx); else mval = MEM_read32(ip0) ^ 1;
        if (MEM_read32(ip0) == mval) goto _offset;
        idx = hashTable[hash1];
        hash0 = hash1; hash1 = ZSTD_hashPtr(ip2, hlog, mls);
        ip0 = ip1; ip1 = ip2; ip2 = ip0 + step; ip3 = ip1 + step;
        if (ip2 >= nextStep) { step++; nextStep += kStepIncr; }
    } while (ip3 < ilimit);
_cleanup:
    rep[0] = rep_offset1 ? rep_offset1 : offsetSaved;
    rep[1] = rep_offset2 ? rep_offset2 : offsetSaved;
    return (size_t)(iend - anchor);
_offset:
    match0 = base + idx;
    rep_offset2 = rep_offset1; rep_offset1 = (U32)(ip0 - match0);
    offcode = rep_offset1 + ZSTD_REP_MOVE;
    mLength = 4;
    while (((ip0 > anchor) & (match0 > prefixStart)) && (ip0[-1] == match0[-1])) { ip0--; match0--; mLength++; }
_match:
    mLength += ZSTD_count(ip0 + mLength, match0 + mLength, iend);
    ZSTD_storeSeq(seqStore, (size_t)(ip0 - anchor), anchor, offcode, mLength - MINMATCH);
    ip0 += mLength; anchor = ip0;
    if (ip1 < ip0) hashTable[hash1] = (U32)(ip1 - base);
    if (ip0 <= ilimit) {
        hashTable[ZSTD_hashPtr(base + current0 + 2, hlog, mls)] = current0 + 2;
        hashTable[ZSTD_hashPtr(ip0 - 2, hlog, mls)] = (U32)(ip0 - 2 - base);
        if (rep_offset2 > 0) {
            while ((ip0 <= ilimit) && (MEM_read32(ip0) == MEM_read32(ip0 - rep_offset2))) {
                size_t const rLength = ZSTD_count(ip0 + 4, ip0 + 4 - rep_offset2, iend) + 4;
                { U32 const tmpOff = rep_offset2; rep_offset2 = rep_offset1; rep_offset1 = tmpOff; }
                hashTable[ZSTD_hashPtr(ip0, hlog, mls)] = (U32)(ip0 - base);
                ip0 += rLength;
                ZSTD_storeSeq(seqStore, 0, anchor, 0, rLength - MINMATCH);
                anchor = ip0;
                continue;
    }   }   }
    goto _start;
}

/* ZstdDoubleFast.cs:51 ZSTD_compressBlock_doubleFast_noDict_generic */
static size_t ZSTD_compressBlock_doubleFast(ZSTD_matchState_t* ms, seqStore_t* seqStore, U32 rep[3], const void* src, size_t srcSize)
{
    const cParams_t* const cParams = &ms->cParams; U32 const mls = cParams->minMatch;
    U32* const hashLong = ms->hashTable; U32 const hBitsL = cParams->hashLog;
    U32* const hashSmall = ms->chainTable; U32 const hBitsS = cParams->chainLog;
    const BYTE* const base = ms->window.base; const BYTE* const istart = (const BYTE*)src; const BYTE* anchor = istart;
    U32 const endIndex = (U32)((size_t)(istart - base) + srcSize);
    U32 const prefixLowestIndex = ZSTD_getLowestPrefixIndex(ms, endIndex, cParams->windowLog);
    const BYTE* const prefixLowest = base + prefixLowestIndex;
    const BYTE* const iend = istart + srcSize; const BYTE* const ilimit = iend - 8;
    U32 offset_1 = rep[0], offset_2 = rep[1]; U32 offsetSaved = 0;
    size_t mLength; U32 offset; U32 curr = 0;
    size_t const kStepIncr = 1 << 8;
    const BYTE* nextStep; size_t step; size_t hl0, hl1 = 0; U32 idxl0, idxl1 = 0;
    const BYTE* matchl0; const BYTE* matchs0; const BYTE* matchl1 = NULL;
    const BYTE* ip = istart; const BYTE* ip1;

    ip += ((ip - prefixLowest) == 0);
    {   U32 const current = (U32)(ip - base);
        U32 const windowLow = ZSTD_getLowestPrefixIndex(ms, current, cParams->windowLog);
        U32 const maxRep = current - windowLow;
        if (offset_2 > maxRep) { offsetSaved = offset_2; offset_2 = 0; }
        if (offset_1 > maxRep) { offsetSaved = offset_1; offset_1 = 0; }
    }
    while (1) {
        step = 1; nextStep = ip + kStepIncr; ip1 = ip + step;
        if (ip1 > ilimit) goto _cleanup;
        hl0 = ZSTD_hashPtr(ip, hBitsL, 8);
        idxl0 = hashLong[hl0]; matchl0 = base + idxl0;
        do {
            size_t const hs0 = ZSTD_hashPtr(ip, hBitsS, mls);
            U32 const idxs0 = hashSmall[hs0];
            curr = (U32)(ip - base);
            matchs0 = base + idxs0;
            hashLong[hl0] = hashSmall[hs0] = curr;
            if ((offset_1 > 0) & (MEM_read32(ip + 1 - offset_1) == MEM_read32(ip + 1))) {
                mLength = ZSTD_count(ip + 1 + 4, ip + 1 + 4 - offset_1, iend) + 4;
                ip++;
                ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, 0, mLength - MINMATCH);
                goto _match_stored;
            }
            hl1 = ZSTD_hashPtr(ip1, hBitsL, 8);
            if (idxl0 > prefixLowestIndex) {
                if (MEM_read64(matchl0) == MEM_read64(ip)) {
                    mLength = ZSTD_count(ip + 8, matchl0 + 8, iend) + 8;
                    offset = (U32)(ip - matchl0);
                    while (((ip > anchor) & (matchl0 > prefixLowest)) && (ip[-1] == matchl0[-1])) { ip--; matchl0--; mLength++; }
                    goto _match_found;
            }   }
            idxl1 = hashLong[hl1]; matchl1 = base + idxl1;
            if (idxs0 > prefixLowestIndex) {
                if (MEM_read32(matchs0) == MEM_read32(ip)) goto _search_next_long;
            }
            if (ip1 >= nextStep) { step++; nextStep += kStepIncr; }
            ip = ip1; ip1 += step;
            hl0 = hl1; idxl0 = idxl1; matchl0 = matchl1;
        } while (ip1 <= ilimit);
_cleanup:
        rep[0] = offset_1 ? offset_1 : offsetSaved;
        rep[1] = offset_2 ? offset_2 : offsetSaved;
        return (size_t)(iend - anchor);
_search_next_long:
        if (idxl1 > prefixLowestIndex) {
            if (MEM_read64(matchl1) == MEM_read64(ip1)) {
                ip = ip1;
                mLength = ZSTD_count(ip + 8, matchl1 + 8, iend) + 8;
                offset = (U32)(ip - matchl1);
                while (((ip > anchor) & (matchl1 > prefixLowest)) && (ip[-1] == matchl1[-1])) { ip--; matchl1--; mLength++; }
                goto _match_found;
        }   }
        mLength = ZSTD_count(ip + 4, matchs0 + 4, iend) + 4;
        offset = (U32)(ip - matchs0);
        while (((ip > anchor) & (matchs0 > prefixLowest)) && (ip[-1] == matchs0[-1])) { ip--; matchs0--; mLength++; }
_match_found:
        offset_2 = offset_1; offset_1 = offset;
        if (step < 4) hashLong[hl1] = (U32)(ip1 - base);
        ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, offset + ZSTD_REP_MOVE, mLength - MINMATCH);
_match_stored:
        ip += mLength; anchor = ip;
        if (ip <= ilimit) {
            {   U32 const indexToInsert = curr + 2;
                hashLong[ZSTD_hashPtr(base + indexToInsert, hBitsL, 8)] = indexToInsert;
                hashLong[ZSTD_hashPtr(ip - 2, hBitsL, 8)] = (U32)(ip - 2 - base);
                hashSmall[ZSTD_hashPtr(base + indexToInsert, hBitsS, mls)] = indexToInsert;
                hashSmall[ZSTD_hashPtr(ip - 1, hBitsS, mls)] = (U32)(ip - 1 - base);
            }
            while ((ip <= ilimit) && ((offset_2 > 0) & (MEM_read32(ip) == MEM_read32(ip - offset_2)))) {
                size_t const rLength = ZSTD_count(ip + 4, ip + 4 - offset_2, iend) + 4;
                U32 const tmpOff = offset_2; offset_2 = offset_1; offset_1 = tmpOff;
                hashSmall[ZSTD_hashPtr(ip, hBitsS, mls)] = (U32)(ip - base);
                hashLong[ZSTD_hashPtr(ip, hBitsL, 8)] = (U32)(ip - base);
                ZSTD_storeSeq(seqStore, 0, anchor, 0, rLength - MINMATCH);
                ip += rLength; anchor = ip;
                continue;
    }   }   }
}


/* =====================================================================================
 *  Match finders with a dictionary -- ZstdFast.cs:390 (dictMatchState), :583 (extDict), ZstdDoubleFast.cs:250, :590.
 *  The oracle keeps the dictionary content and the input in ONE buffer (content directly in front of the input), so
 *  `dictBase + index` and `base + index` are the same address and ZSTD_count_2segments (ZstdCompressInternal.cs:283) is a
 *  plain ZSTD_count up to iend; everything else (which indices are eligible, where a backward extension stops, the 3-byte
 *  guard in front of the prefix start) is restated as the reference has it.
 * ===================================================================================== */
static size_t ZSTD_compressBlock_fast_dictMatchState(ZSTD_matchState_t* ms, seqStore_t* seqStore, U32 rep[3], const void* src, size_t srcSize)   /* ZstdFast.cs:390 */
{
    const cParams_t* const cParams = &ms->cParams;
    U32* const hashTable = ms->hashTable; U32 const hlog = cParams->hashLog; U32 const mls = cParams->minMatch;
    U32 const stepSize = cParams->targetLength + !(cParams->targetLength);
    const BYTE* const base = ms->window.base; const BYTE* const istart = (const BYTE*)src; const BYTE* ip = istart; const BYTE* anchor = istart;
    U32 const prefixStartIndex = ms->window.dictLimit; const BYTE* const prefixStart = base + prefixStartIndex;
    const BYTE* const iend = istart + srcSize; const BYTE* const ilimit = iend - 8;
    U32 offset_1 = rep[0], offset_2 = rep[1]; U32 const offsetSaved = 0;
    const ZSTD_matchState_t* const dms = ms->dictMatchState;
    const U32* const dictHashTable = dms->hashTable; U32 const dictStartIndex = dms->window.dictLimit;
    U32 const dictIndexDelta = prefixStartIndex - ms->dictEndIndex;                     /* dms index d is byte base[d + delta] */
    const BYTE* const dictStart = base + dictStartIndex + dictIndexDelta;
    U32 const dictAndPrefixLength = (U32)(ip - prefixStart) + (ms->dictEndIndex - dictStartIndex);
    U32 const dictHLog = dms->cParams.hashLog;
    ip += (dictAndPrefixLength == 0);
    while (ip < ilimit) {
        size_t mLength; size_t const h = ZSTD_hashPtr(ip, hlog, mls);
        U32 const curr = (U32)(ip - base); U32 const matchIndex = hashTable[h]; const BYTE* match = base + matchIndex;
        U32 const repIndex = curr + 1 - offset_1; const BYTE* const repMatch = base + repIndex;
        hashTable[h] = curr;
        if (((U32)((prefixStartIndex - 1) - repIndex) >= 3) && (MEM_read32(repMatch) == MEM_read32(ip + 1))) {
            mLength = ZSTD_count(ip + 1 + 4, repMatch + 4, iend) + 4;
            ip++;
            ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, 0, mLength - MINMATCH);
        } else if (matchIndex <= prefixStartIndex) {
            size_t const dictHash = ZSTD_hashPtr(ip, dictHLog, mls);
            U32 const dictMatchIndex = dictHashTable[dictHash]; const BYTE* dictMatch = base + dictMatchIndex + dictIndexDelta;
            if (dictMatchIndex <= dictStartIndex || MEM_read32(dictMatch) != MEM_read32(ip)) { ip += ((ip - anchor) >> 8) + stepSize; continue; }
            {   U32 const offset = (U32)(curr - dictMatchIndex - dictIndexDelta);
                mLength = ZSTD_count(ip + 4, dictMatch + 4, iend) + 4;
                while (((ip > anchor) & (dictMatch > dictStart)) && (ip[-1] == dictMatch[-1])) { ip--; dictMatch--; mLength++; }
                offset_2 = offset_1; offset_1 = offset;
                ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, offset + ZSTD_REP_MOVE, mLength - MINMATCH);
            }
        } else if (MEM_read32(match) != MEM_read32(ip)) { ip += ((ip - anchor) >> 8) + stepSize; continue; }
        else {
            U32 const offset = (U32)(ip - match);
            mLength = ZSTD_count(ip + 4, match + 4, iend) + 4;
            while (((ip > anchor) & (match > prefixStart)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
            offset_2 = offset_1; offset_1 = offset;
            ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, offset + ZSTD_REP_MOVE, mLength - MINMATCH);
        }
        ip += mLength; anchor = ip;
        if (ip <= ilimit) {
            hashTable[ZSTD_hashPtr(base + curr + 2, hlog, mls)] = curr + 2;
            hashTable[ZSTD_hashPtr(ip - 2, hlog, mls)] = (U32)(ip - 2 - base);
            while (ip <= ilimit) {
                U32 const current2 = (U32)(ip - base); U32 const repIndex2 = current2 - offset_2; const BYTE* const repMatch2 = base + repIndex2;
                if (((U32)((prefixStartIndex - 1) - repIndex2) >= 3) && (MEM_read32(repMatch2) == MEM_read32(ip))) {
                    size_t const repLength2 = ZSTD_count(ip + 4, repMatch2 + 4, iend) + 4;
                    U32 const tmpOffset = offset_2; offset_2 = offset_1; offset_1 = tmpOffset;
                    ZSTD_storeSeq(seqStore, 0, anchor, 0, repLength2 - MINMATCH);
                    hashTable[ZSTD_hashPtr(ip, hlog, mls)] = current2;
                    ip += repLength2; anchor = ip;
                    continue;
                }
                break;
    }   }   }
    rep[0] = offset_1 ? offset_1 : offsetSaved; rep[1] = offset_2 ? offset_2 : offsetSaved;
    return (size_t)(iend - anchor);
}

static size_t ZSTD_compressBlock_fast_extDict(ZSTD_matchState_t* ms, seqStore_t* seqStore, U32 rep[3], const void* src, size_t srcSize)   /* ZstdFast.cs:583 */
{
    const cParams_t* const cParams = &ms->cParams;
    U32* const hashTable = ms->hashTable; U32 const hlog = cParams->hashLog; U32 const mls = cParams->minMatch;
    U32 const stepSize = cParams->targetLength + !(cParams->targetLength);
    const BYTE* const base = ms->window.base; const BYTE* const istart = (const BYTE*)src; const BYTE* ip = istart; const BYTE* anchor = istart;
    U32 const endIndex = (U32)((size_t)(istart - base) + srcSize);
    U32 const lowLimit = ZSTD_getLowestMatchIndex(ms, endIndex, cParams->windowLog);
    U32 const dictStartIndex = lowLimit; const BYTE* const dictStart = base + dictStartIndex;
    U32 const dictLimit = ms->window.dictLimit; U32 const prefixStartIndex = dictLimit < lowLimit ? lowLimit : dictLimit;
    const BYTE* const prefixStart = base + prefixStartIndex;
    const BYTE* const iend = istart + srcSize; const BYTE* const ilimit = iend - 8;
    U32 offset_1 = rep[0], offset_2 = rep[1];
    if (prefixStartIndex == dictStartIndex) return ZSTD_compressBlock_fast(ms, seqStore, rep, src, srcSize);
    while (ip < ilimit) {
        size_t const h = ZSTD_hashPtr(ip, hlog, mls);
        U32 const matchIndex = hashTable[h]; const BYTE* match = base + matchIndex;
        U32 const curr = (U32)(ip - base); U32 const repIndex = curr + 1 - offset_1; const BYTE* const repMatch = base + repIndex;
        hashTable[h] = curr;
        if ((((U32)((prefixStartIndex - 1) - repIndex) >= 3) & (offset_1 <= curr + 1 - dictStartIndex)) && (MEM_read32(repMatch) == MEM_read32(ip + 1))) {
            size_t const rLength = ZSTD_count(ip + 1 + 4, repMatch + 4, iend) + 4;
            ip++;
            ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, 0, rLength - MINMATCH);
            ip += rLength; anchor = ip;
        } else {
            if ((matchIndex < dictStartIndex) || (MEM_read32(match) != MEM_read32(ip))) { ip += ((ip - anchor) >> 8) + stepSize; continue; }
            {   const BYTE* const lowMatchPtr = matchIndex < prefixStartIndex ? dictStart : prefixStart;
                U32 const offset = curr - matchIndex;
                size_t mLength = ZSTD_count(ip + 4, match + 4, iend) + 4;
                while (((ip > anchor) & (match > lowMatchPtr)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
                offset_2 = offset_1; offset_1 = offset;
                ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, offset + ZSTD_REP_MOVE, mLength - MINMATCH);
                ip += mLength; anchor = ip;
        }   }
        if (ip <= ilimit) {
            hashTable[ZSTD_hashPtr(base + curr + 2, hlog, mls)] = curr + 2;
            hashTable[ZSTD_hashPtr(ip - 2, hlog, mls)] = (U32)(ip - 2 - base);
            while (ip <= ilimit) {
                U32 const current2 = (U32)(ip - base); U32 const repIndex2 = current2 - offset_2; const BYTE* const repMatch2 = base + repIndex2;
                if ((((U32)((prefixStartIndex - 1) - repIndex2) >= 3) & (offset_2 <= curr - dictStartIndex)) && (MEM_read32(repMatch2) == MEM_read32(ip))) {   /* `curr`, not current2: as the reference has it (:676) */
                    size_t const repLength2 = ZSTD_count(ip + 4, repMatch2 + 4, iend) + 4;
                    U32 const tmpOffset = offset_2; offset_2 = offset_1; offset_1 = tmpOffset;
                    ZSTD_storeSeq(seqStore, 0, anchor, 0, repLength2 - MINMATCH);
                    hashTable[ZSTD_hashPtr(ip, hlog, mls)] = current2;
                    ip += repLength2; anchor = ip;
                    continue;
                }
                break;
    }   }   }
    rep[0] = offset_1; rep[1] = offset_2;
    return (size_t)(iend - anchor);
}

static size_t ZSTD_compressBlock_doubleFast_dictMatchState(ZSTD_matchState_t* ms, seqStore_t* seqStore, U32 rep[3], const void* src, size_t srcSize)   /* ZstdDoubleFast.cs:250 */
{
    const cParams_t* const cParams = &ms->cParams; U32 const mls = cParams->minMatch;
    U32* const hashLong = ms->hashTable; U32 const hBitsL = cParams->hashLog;
    U32* const hashSmall = ms->chainTable; U32 const hBitsS = cParams->chainLog;
    const BYTE* const base = ms->window.base; const BYTE* const istart = (const BYTE*)src; const BYTE* ip = istart; const BYTE* anchor = istart;
    U32 const endIndex = (U32)((size_t)(istart - base) + srcSize);
    U32 const prefixLowestIndex = ZSTD_getLowestPrefixIndex(ms, endIndex, cParams->windowLog);
    const BYTE* const prefixLowest = base + prefixLowestIndex;
    const BYTE* const iend = istart + srcSize; const BYTE* const ilimit = iend - 8;
    U32 offset_1 = rep[0], offset_2 = rep[1]; U32 const offsetSaved = 0;
    const ZSTD_matchState_t* const dms = ms->dictMatchState;
    const U32* const dictHashLong = dms->hashTable; const U32* const dictHashSmall = dms->chainTable;
    U32 const dictStartIndex = dms->window.dictLimit;
    U32 const dictIndexDelta = prefixLowestIndex - ms->dictEndIndex;
    const BYTE* const dictStart = base + dictStartIndex + dictIndexDelta;
    U32 const dictHBitsL = dms->cParams.hashLog; U32 const dictHBitsS = dms->cParams.chainLog;
    U32 const dictAndPrefixLength = (U32)(ip - prefixLowest) + (ms->dictEndIndex - dictStartIndex);
    ip += (dictAndPrefixLength == 0);
    while (ip < ilimit) {
        size_t mLength; U32 offset;
        size_t const h2 = ZSTD_hashPtr(ip, hBitsL, 8); size_t const h = ZSTD_hashPtr(ip, hBitsS, mls);
        size_t const dictHL = ZSTD_hashPtr(ip, dictHBitsL, 8); size_t const dictHS = ZSTD_hashPtr(ip, dictHBitsS, mls);
        U32 const curr = (U32)(ip - base);
        U32 const matchIndexL = hashLong[h2]; U32 matchIndexS = hashSmall[h];
        const BYTE* matchLong = base + matchIndexL; const BYTE* match = base + matchIndexS;
        U32 const repIndex = curr + 1 - offset_1; const BYTE* const repMatch = base + repIndex;
        hashLong[h2] = hashSmall[h] = curr;
        if (((U32)((prefixLowestIndex - 1) - repIndex) >= 3) && (MEM_read32(repMatch) == MEM_read32(ip + 1))) {
            mLength = ZSTD_count(ip + 1 + 4, repMatch + 4, iend) + 4;
            ip++;
            ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, 0, mLength - MINMATCH);
            goto _match_stored;
        }
        if (matchIndexL > prefixLowestIndex) {
            if (MEM_read64(matchLong) == MEM_read64(ip)) {
                mLength = ZSTD_count(ip + 8, matchLong + 8, iend) + 8;
                offset = (U32)(ip - matchLong);
                while (((ip > anchor) & (matchLong > prefixLowest)) && (ip[-1] == matchLong[-1])) { ip--; matchLong--; mLength++; }
                goto _match_found;
            }
        } else {
            U32 const dictMatchIndexL = dictHashLong[dictHL]; const BYTE* dictMatchL = base + dictMatchIndexL + dictIndexDelta;
            if (dictMatchL > dictStart && MEM_read64(dictMatchL) == MEM_read64(ip)) {
                mLength = ZSTD_count(ip + 8, dictMatchL + 8, iend) + 8;
                offset = (U32)(curr - dictMatchIndexL - dictIndexDelta);
                while (((ip > anchor) & (dictMatchL > dictStart)) && (ip[-1] == dictMatchL[-1])) { ip--; dictMatchL--; mLength++; }
                goto _match_found;
        }   }
        if (matchIndexS > prefixLowestIndex) {
            if (MEM_read32(match) == MEM_read32(ip)) goto _search_next_long;
        } else {
            U32 const dictMatchIndexS = dictHashSmall[dictHS];
            match = base + dictMatchIndexS + dictIndexDelta;
            matchIndexS = dictMatchIndexS + dictIndexDelta;
            if (match > dictStart && MEM_read32(match) == MEM_read32(ip)) goto _search_next_long;
        }
        ip += ((ip - anchor) >> 8) + 1;
        continue;
_search_next_long:
        {   size_t const hl3 = ZSTD_hashPtr(ip + 1, hBitsL, 8); size_t const dictHLNext = ZSTD_hashPtr(ip + 1, dictHBitsL, 8);
            U32 const matchIndexL3 = hashLong[hl3]; const BYTE* matchL3 = base + matchIndexL3;
            hashLong[hl3] = curr + 1;
            if (matchIndexL3 > prefixLowestIndex) {
                if (MEM_read64(matchL3) == MEM_read64(ip + 1)) {
                    mLength = ZSTD_count(ip + 9, matchL3 + 8, iend) + 8;
                    ip++;
                    offset = (U32)(ip - matchL3);
                    while (((ip > anchor) & (matchL3 > prefixLowest)) && (ip[-1] == matchL3[-1])) { ip--; matchL3--; mLength++; }
                    goto _match_found;
                }
            } else {
                U32 const dictMatchIndexL3 = dictHashLong[dictHLNext]; const BYTE* dictMatchL3 = base + dictMatchIndexL3 + dictIndexDelta;
                if (dictMatchL3 > dictStart && MEM_read64(dictMatchL3) == MEM_read64(ip + 1)) {
                    mLength = ZSTD_count(ip + 1 + 8, dictMatchL3 + 8, iend) + 8;
                    ip++;
                    offset = (U32)(curr + 1 - dictMatchIndexL3 - dictIndexDelta);
                    while (((ip > anchor) & (dictMatchL3 > dictStart)) && (ip[-1] == dictMatchL3[-1])) { ip--; dictMatchL3--; mLength++; }
                    goto _match_found;
        }   }   }
        if (matchIndexS < prefixLowestIndex) {
            mLength = ZSTD_count(ip + 4, match + 4, iend) + 4;
            offset = (U32)(curr - matchIndexS);
            while (((ip > anchor) & (match > dictStart)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
        } else {
            mLength = ZSTD_count(ip + 4, match + 4, iend) + 4;
            offset = (U32)(ip - match);
            while (((ip > anchor) & (match > prefixLowest)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
        }
_match_found:
        offset_2 = offset_1; offset_1 = offset;
        ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, offset + ZSTD_REP_MOVE, mLength - MINMATCH);
_match_stored:
        ip += mLength; anchor = ip;
        if (ip <= ilimit) {
            {   U32 const indexToInsert = curr + 2;
                hashLong[ZSTD_hashPtr(base + indexToInsert, hBitsL, 8)] = indexToInsert;
                hashLong[ZSTD_hashPtr(ip - 2, hBitsL, 8)] = (U32)(ip - 2 - base);
                hashSmall[ZSTD_hashPtr(base + indexToInsert, hBitsS, mls)] = indexToInsert;
                hashSmall[ZSTD_hashPtr(ip - 1, hBitsS, mls)] = (U32)(ip - 1 - base);
            }
            while (ip <= ilimit) {
                U32 const current2 = (U32)(ip - base); U32 const repIndex2 = current2 - offset_2; const BYTE* const repMatch2 = base + repIndex2;
                if (((U32)((prefixLowestIndex - 1) - repIndex2) >= 3) && (MEM_read32(repMatch2) == MEM_read32(ip))) {
                    size_t const repLength2 = ZSTD_count(ip + 4, repMatch2 + 4, iend) + 4;
                    U32 const tmpOffset = offset_2; offset_2 = offset_1; offset_1 = tmpOffset;
                    ZSTD_storeSeq(seqStore, 0, anchor, 0, repLength2 - MINMATCH);
                    hashSmall[ZSTD_hashPtr(ip, hBitsS, mls)] = current2;
                    hashLong[ZSTD_hashPtr(ip, hBitsL, 8)] = current2;
                    ip += repLength2; anchor = ip;
                    continue;
                }
                break;
    }   }   }
    rep[0] = offset_1 ? offset_1 : offsetSaved; rep[1] = offset_2 ? offset_2 : offsetSaved;
    return (size_t)(iend - anchor);
}

static size_t ZSTD_compressBlock_doubleFast_extDict(ZSTD_matchState_t* ms, seqStore_t* seqStore, U32 rep[3], const void* src, size_t srcSize)   /* ZstdDoubleFast.cs:590 */
{
    const cParams_t* const cParams = &ms->cParams; U32 const mls = cParams->minMatch;
    U32* const hashLong = ms->hashTable; U32 const hBitsL = cParams->hashLog;
    U32* const hashSmall = ms->chainTable; U32 const hBitsS = cParams->chainLog;
    const BYTE* const istart = (const BYTE*)src; const BYTE* ip = istart; const BYTE* anchor = istart;
    const BYTE* const iend = istart + srcSize; const BYTE* const ilimit = iend - 8;
    const BYTE* const base = ms->window.base;
    U32 const endIndex = (U32)((size_t)(istart - base) + srcSize);
    U32 const lowLimit = ZSTD_getLowestMatchIndex(ms, endIndex, cParams->windowLog);
    U32 const dictStartIndex = lowLimit; U32 const dictLimit = ms->window.dictLimit;
    U32 const prefixStartIndex = (dictLimit > lowLimit) ? dictLimit : lowLimit;
    const BYTE* const prefixStart = base + prefixStartIndex; const BYTE* const dictStart = base + dictStartIndex;
    U32 offset_1 = rep[0], offset_2 = rep[1];
    if (prefixStartIndex == dictStartIndex) return ZSTD_compressBlock_doubleFast(ms, seqStore, rep, src, srcSize);
    while (ip < ilimit) {
        size_t const hSmall = ZSTD_hashPtr(ip, hBitsS, mls); U32 const matchIndex = hashSmall[hSmall]; const BYTE* match = base + matchIndex;
        size_t const hLong = ZSTD_hashPtr(ip, hBitsL, 8); U32 const matchLongIndex = hashLong[hLong]; const BYTE* matchLong = base + matchLongIndex;
        U32 const curr = (U32)(ip - base); U32 const repIndex = curr + 1 - offset_1; const BYTE* const repMatch = base + repIndex;
        size_t mLength;
        hashSmall[hSmall] = hashLong[hLong] = curr;
        if ((((U32)((prefixStartIndex - 1) - repIndex) >= 3) & (offset_1 <= curr + 1 - dictStartIndex)) && (MEM_read32(repMatch) == MEM_read32(ip + 1))) {
            mLength = ZSTD_count(ip + 1 + 4, repMatch + 4, iend) + 4;
            ip++;
            ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, 0, mLength - MINMATCH);
        } else {
            if ((matchLongIndex > dictStartIndex) && (MEM_read64(matchLong) == MEM_read64(ip))) {
                const BYTE* const lowMatchPtr = matchLongIndex < prefixStartIndex ? dictStart : prefixStart;
                U32 offset;
                mLength = ZSTD_count(ip + 8, matchLong + 8, iend) + 8;
                offset = curr - matchLongIndex;
                while (((ip > anchor) & (matchLong > lowMatchPtr)) && (ip[-1] == matchLong[-1])) { ip--; matchLong--; mLength++; }
                offset_2 = offset_1; offset_1 = offset;
                ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, offset + ZSTD_REP_MOVE, mLength - MINMATCH);
            } else if ((matchIndex > dictStartIndex) && (MEM_read32(match) == MEM_read32(ip))) {
                size_t const h3 = ZSTD_hashPtr(ip + 1, hBitsL, 8); U32 const matchIndex3 = hashLong[h3]; const BYTE* match3 = base + matchIndex3;
                U32 offset;
                hashLong[h3] = curr + 1;
                if ((matchIndex3 > dictStartIndex) && (MEM_read64(match3) == MEM_read64(ip + 1))) {
                    const BYTE* const lowMatchPtr = matchIndex3 < prefixStartIndex ? dictStart : prefixStart;
                    mLength = ZSTD_count(ip + 9, match3 + 8, iend) + 8;
                    ip++;
                    offset = curr + 1 - matchIndex3;
                    while (((ip > anchor) & (match3 > lowMatchPtr)) && (ip[-1] == match3[-1])) { ip--; match3--; mLength++; }
                } else {
                    const BYTE* const lowMatchPtr = matchIndex < prefixStartIndex ? dictStart : prefixStart;
                    mLength = ZSTD_count(ip + 4, match + 4, iend) + 4;
                    offset = curr - matchIndex;
                    while (((ip > anchor) & (match > lowMatchPtr)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
                }
                offset_2 = offset_1; offset_1 = offset;
                ZSTD_storeSeq(seqStore, (size_t)(ip - anchor), anchor, offset + ZSTD_REP_MOVE, mLength - MINMATCH);
            } else { ip += ((ip - anchor) >> 8) + 1; continue; }
        }
        ip += mLength; anchor = ip;
        if (ip <= ilimit) {
            {   U32 const indexToInsert = curr + 2;
                hashLong[ZSTD_hashPtr(base + indexToInsert, hBitsL, 8)] = indexToInsert;
                hashLong[ZSTD_hashPtr(ip - 2, hBitsL, 8)] = (U32)(ip - 2 - base);
                hashSmall[ZSTD_hashPtr(base + indexToInsert, hBitsS, mls)] = indexToInsert;
                hashSmall[ZSTD_hashPtr(ip - 1, hBitsS, mls)] = (U32)(ip - 1 - base);
            }
            while (ip <= ilimit) {
                U32 const current2 = (U32)(ip - base); U32 const repIndex2 = current2 - offset_2; const BYTE* const repMatch2 = base + repIndex2;
                if ((((U32)((prefixStartIndex - 1) - repIndex2) >= 3) & (offset_2 <= current2 - dictStartIndex)) && (MEM_read32(repMatch2) == MEM_read32(ip))) {
                    size_t const repLength2 = ZSTD_count(ip + 4, repMatch2 + 4, iend) + 4;
                    U32 const tmpOffset = offset_2; offset_2 = offset_1; offset_1 = tmpOffset;
                    ZSTD_storeSeq(seqStore, 0, anchor, 0, repLength2 - MINMATCH);
                    hashSmall[ZSTD_hashPtr(ip, hBitsS, mls)] = current2;
                    hashLong[ZSTD_hashPtr(ip, hBitsL, 8)] = current2;
                    ip += repLength2; anchor = ip;
                    continue;
                }
                break;
    }   }   }
    rep[0] = offset_1; rep[1] = offset_2;
    return (size_t)(iend - anchor);
}

/* =====================================================================================
 *  Block / frame layer -- ZstdCompress.cs:3432, :4528, :4690, :4817, :5013, :5598, :5665, :7138
 * ===================================================================================== */
typedef struct {
    cParams_t cParams; ZSTD_matchState_t ms; seqStore_t seqStore;
    ZSTD_compressedBlockState_t blockStateA, blockStateB; ZSTD_compressedBlockState_t* prevCBlock; ZSTD_compressedBlockState_t* nextCBlock;
    int isFirstBlock; size_t blockSize;
} zo_CCtx;

static void ZSTD_reset_compressedBlockState(ZSTD_compressedBlockState_t* bs)   /* :2427 */
{
    int i; for (i = 0; i < 3; ++i) bs->rep[i] = repStartValue[i];
    bs->entropy.huf.repeatMode = HUF_repeat_none;
    bs->entropy.fse.offcode_repeatMode = FSE_repeat_none;
    bs->entropy.fse.matchlength_repeatMode = FSE_repeat_none;
    bs->entropy.fse.litlength_repeatMode = FSE_repeat_none;
}

static int ZSTD_isRLE(const BYTE* src, size_t length) { size_t i; for (i = 1; i < length; i++) if (src[i] != src[0]) return 0; return 1; }   /* :3671 */

static size_t ZSTD_buildSeqStore(zo_CCtx* zc, const void* src, size_t srcSize)   /* :3432 ; returns 1 = noCompress, 0 = compress */
{
    if (srcSize < MIN_CBLOCK_SIZE + ZSTD_blockHeaderSize + 1) return 1;
    zc->seqStore.lit = zc->seqStore.litStart; zc->seqStore.sequences = zc->seqStore.sequencesStart; zc->seqStore.longLengthType = 0;
    {   int i; for (i = 0; i < 3; ++i) zc->nextCBlock->rep[i] = zc->prevCBlock->rep[i]; }
    {   /* ZSTD_matchState_dictMode (ZstdCompressInternal.cs:576) + ZSTD_selectBlockCompressor (ZstdCompress.cs:3398) */
        int const extDict = zc->ms.window.lowLimit < zc->ms.window.dictLimit;
        int const dms = !extDict && zc->ms.dictMatchState != NULL;
        int const fast = (zc->cParams.strategy == ZSTD_fast);
        size_t const lastLLSize =
            extDict ? (fast ? ZSTD_compressBlock_fast_extDict(&zc->ms, &zc->seqStore, zc->nextCBlock->rep, src, srcSize)
                            : ZSTD_compressBlock_doubleFast_extDict(&zc->ms, &zc->seqStore, zc->nextCBlock->rep, src, srcSize))
          : dms     ? (fast ? ZSTD_compressBlock_fast_dictMatchState(&zc->ms, &zc->seqStore, zc->nextCBlock->rep, src, srcSize)
                            : ZSTD_compressBlock_doubleFast_dictMatchState(&zc->ms, &zc->seqStore, zc->nextCBlock->rep, src, srcSize))
          :           (fast ? ZSTD_compressBlock_fast(&zc->ms, &zc->seqStore, zc->nextCBlock->rep, src, srcSize)
                            : ZSTD_compressBlock_doubleFast(&zc->ms, &zc->seqStore, zc->nextCBlock->rep, src, srcSize));
        const BYTE* const lastLiterals = (const BYTE*)src + srcSize - lastLLSize;
        memcpy(zc->seqStore.lit, lastLiterals, lastLLSize);   /* ZSTD_storeLastLiterals */
        zc->seqStore.lit += lastLLSize;
    }
    return 0;
}

static size_t ZSTD_compressBlock_internal(zo_CCtx* zc, void* dst, size_t dstCapacity, const void* src, size_t srcSize, U32 frame)   /* :4528 */
{
    U32 const rleMaxLength = 25; size_t cSize; const BYTE* ip = (const BYTE*)src; BYTE* op = (BYTE*)dst;
    if (ZSTD_buildSeqStore(zc, src, srcSize)) { cSize = 0; goto out; }
    /* ZSTD_literalsCompressionIsDisabled (ZstdCompressInternal.cs:483-498), ZSTD_ps_auto: fast strategy with an acceleration factor */
    cSize = ZSTD_entropyCompressSeqStore(&zc->seqStore, &zc->prevCBlock->entropy, &zc->nextCBlock->entropy, zc->cParams.strategy,
                                         zc->cParams.strategy == ZSTD_fast && zc->cParams.targetLength > 0, dst, dstCapacity, srcSize);
    if (frame && !zc->isFirstBlock && cSize < rleMaxLength && ZSTD_isRLE(ip, srcSize)) { cSize = 1; op[0] = ip[0]; }
out:
    if (!ERR_isError(cSize) && cSize > 1) {   /* ZSTD_blockState_confirmRepcodesAndEntropyTables */
        ZSTD_compressedBlockState_t* const tmp = zc->prevCBlock; zc->prevCBlock = zc->nextCBlock; zc->nextCBlock = tmp;
    }
    if (zc->prevCBlock->entropy.fse.offcode_repeatMode == FSE_repeat_valid) zc->prevCBlock->entropy.fse.offcode_repeatMode = FSE_repeat_check;
    return cSize;
}

static size_t ZSTD_noCompressBlock(void* dst, size_t dstCapacity, const void* src, size_t srcSize, U32 lastBlock)   /* ZstdCompressInternal.cs:102 */
{
    U32 const cBlockHeader24 = lastBlock + (((U32)bt_raw) << 1) + (U32)(srcSize << 3);
    if (srcSize + ZSTD_blockHeaderSize > dstCapacity) return ERROR(dstSize_tooSmall);
    MEM_writeLE24(dst, cBlockHeader24);
    memcpy((BYTE*)dst + ZSTD_blockHeaderSize, src, srcSize);
    return ZSTD_blockHeaderSize + srcSize;
}

static size_t ZSTD_compress_frameChunk(zo_CCtx* cctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize)   /* :4690, lastFrameChunk = 1 */
{
    size_t blockSize = cctx->blockSize; size_t remaining = srcSize;
    const BYTE* ip = (const BYTE*)src; BYTE* const ostart = (BYTE*)dst; BYTE* op = ostart;
    U32 const maxDist = 1U << cctx->cParams.windowLog;
    while (remaining) {
        ZSTD_matchState_t* const ms = &cctx->ms;
        U32 const lastBlock = (blockSize >= remaining);
        if (dstCapacity < ZSTD_blockHeaderSize + MIN_CBLOCK_SIZE) return ERROR(dstSize_tooSmall);
        if (remaining < blockSize) blockSize = remaining;
        {   /* ZSTD_checkDictValidity(&ms->window, ip + blockSize, maxDist, ...) : ZstdCompressInternal.cs:697 */
            U32 const blockEndIdx = (U32)((ip + blockSize) - ms->window.base);
            if (blockEndIdx > ms->loadedDictEnd + maxDist) { ms->loadedDictEnd = 0; ms->dictMatchState = NULL; }
        }
        {   /* ZSTD_window_enforceMaxDist(&ms->window, ip, maxDist, ...) : ZstdCompressInternal.cs:659 */
            U32 const blockEndIdx = (U32)(ip - ms->window.base);
            if (blockEndIdx > maxDist + ms->loadedDictEnd) {
                U32 const newLowLimit = blockEndIdx - maxDist;
                if (ms->window.lowLimit < newLowLimit) ms->window.lowLimit = newLowLimit;
                if (ms->window.dictLimit < ms->window.lowLimit) ms->window.dictLimit = ms->window.lowLimit;
                ms->loadedDictEnd = 0; ms->dictMatchState = NULL;
        }   }
        {   size_t cSize = ZSTD_compressBlock_internal(cctx, op + ZSTD_blockHeaderSize, dstCapacity - ZSTD_blockHeaderSize, ip, blockSize, 1);
            if (ERR_isError(cSize)) return cSize;
            if (cSize == 0) {
                cSize = ZSTD_noCompressBlock(op, dstCapacity, ip, blockSize, lastBlock);
                if (ERR_isError(cSize)) return cSize;
            } else {
                U32 const cBlockHeader = cSize == 1 ? lastBlock + (((U32)bt_rle) << 1) + (U32)(blockSize << 3)
                                                    : lastBlock + (((U32)bt_compressed) << 1) + (U32)(cSize << 3);
                MEM_writeLE24(op, cBlockHeader);
                cSize += ZSTD_blockHeaderSize;
            }
            ip += blockSize; remaining -= blockSize; op += cSize; dstCapacity -= cSize;
            cctx->isFirstBlock = 0;
    }   }
    return (size_t)(op - ostart);
}

static size_t ZSTD_writeFrameHeader(void* dst, size_t dstCapacity, U32 windowLog, int checksumFlag, U64 pledgedSrcSize, U32 dictID)   /* :4817, contentSizeFlag=1, noDictIDFlag=0 */
{
    BYTE* const op = (BYTE*)dst;
    U32 const dictIDSizeCode = (dictID > 0) + (dictID >= 256) + (dictID >= 65536);
    U32 const windowSize = 1U << windowLog;
    U32 const singleSegment = (windowSize >= pledgedSrcSize);
    BYTE const windowLogByte = (BYTE)((windowLog - 10) << 3);
    U32 const fcsCode = (pledgedSrcSize >= 256) + (pledgedSrcSize >= 65536 + 256) + (pledgedSrcSize >= 0xFFFFFFFFU);
    BYTE const frameHeaderDescriptionByte = (BYTE)(dictIDSizeCode + ((checksumFlag > 0) << 2) + (singleSegment << 5) + (fcsCode << 6));
    size_t pos = 0;
    if (dstCapacity < 18) return ERROR(dstSize_tooSmall);
    MEM_write32(dst, ZSTD_MAGICNUMBER); pos = 4;
    op[pos++] = frameHeaderDescriptionByte;
    if (!singleSegment) op[pos++] = windowLogByte;
    switch (dictIDSizeCode) {
    default: case 0: break;
    case 1: op[pos] = (BYTE)dictID; pos++; break;
    case 2: MEM_write16(op + pos, (U16)dictID); pos += 2; break;
    case 3: MEM_write32(op + pos, dictID); pos += 4; break;
    }
    switch (fcsCode) {
    default: case 0: if (singleSegment) op[pos++] = (BYTE)pledgedSrcSize; break;
    case 1: MEM_write16(op + pos, (U16)(pledgedSrcSize - 256)); pos += 2; break;
    case 2: MEM_write32(op + pos, (U32)pledgedSrcSize); pos += 4; break;
    case 3: MEM_write64(op + pos, (U64)pledgedSrcSize); pos += 8; break;
    }
    return pos;
}

/* Context buffers are sized once for the largest supported geometry (ZSTD_resetCCtx_internal :2548 reuses the
 * workspace the same way); the match-finder tables are re-zeroed for every frame (:2472, :2481-2484). */
#define ZO_MAX_HASHLOG 18
#define ZO_MAX_CHAINLOG 18
static size_t zo_ctx_buffers(zo_CCtx* c)
{
    size_t const maxNbSeq = ZSTD_BLOCKSIZE_MAX / 3;
    if (c->seqStore.sequencesStart) return 0;
    c->seqStore.sequencesStart = (zo_seqDef*)malloc((maxNbSeq + 1) * sizeof(zo_seqDef));
    c->seqStore.litStart = (BYTE*)malloc(ZSTD_BLOCKSIZE_MAX + 32);
    c->seqStore.llCode = (BYTE*)malloc(maxNbSeq + 1); c->seqStore.mlCode = (BYTE*)malloc(maxNbSeq + 1); c->seqStore.ofCode = (BYTE*)malloc(maxNbSeq + 1);
    c->ms.hashTable = (U32*)malloc(((size_t)1 << ZO_MAX_HASHLOG) * sizeof(U32));
    c->ms.chainTable = (U32*)malloc(((size_t)1 << ZO_MAX_CHAINLOG) * sizeof(U32));
    if (!c->seqStore.sequencesStart || !c->seqStore.litStart || !c->seqStore.llCode || !c->seqStore.mlCode || !c->seqStore.ofCode || !c->ms.hashTable || !c->ms.chainTable)
        return ERROR(memory_allocation);
    return 0;
}
static size_t zo_ctx_alloc(zo_CCtx* c, size_t srcSize, int level)
{
    if (zo_getCParams_internal(&c->cParams, level, srcSize)) return ERROR(parameter_unsupported);
    CHECK_F(zo_ctx_buffers(c));
    {   /* ZSTD_resetCCtx_internal :2548 : windowSize = max(1, min(1<<wlog, pledged)); blockSize = min(128K, windowSize) */
        size_t const windowSize = (size_t)1 << c->cParams.windowLog;
        size_t ws = windowSize < srcSize ? windowSize : srcSize; if (ws < 1) ws = 1;
        c->blockSize = ws < ZSTD_BLOCKSIZE_MAX ? ws : ZSTD_BLOCKSIZE_MAX;
    }
    {   size_t const divider = (c->cParams.minMatch == 3) ? 3 : 4;
        c->seqStore.maxNbSeq = c->blockSize / divider; c->seqStore.maxNbLit = c->blockSize; }
    c->ms.cParams = c->cParams;
    memset(c->ms.hashTable, 0, ((size_t)1 << c->cParams.hashLog) * sizeof(U32));                         /* tables zeroed per frame: :2472,:2481 */
    if (c->cParams.strategy != ZSTD_fast) memset(c->ms.chainTable, 0, ((size_t)1 << c->cParams.chainLog) * sizeof(U32));
    c->prevCBlock = &c->blockStateA; c->nextCBlock = &c->blockStateB;
    ZSTD_reset_compressedBlockState(c->prevCBlock);
    c->isFirstBlock = 1;
    return 0;
}
static void zo_ctx_free(zo_CCtx* c)
{
    free(c->seqStore.sequencesStart); free(c->seqStore.litStart); free(c->seqStore.llCode); free(c->seqStore.mlCode); free(c->seqStore.ofCode);
    free(c->ms.hashTable); free(c->ms.chainTable);
}

void* zo_createCCtx(void) { return calloc(1, sizeof(zo_CCtx)); }
void zo_freeCCtx(void* ctx) { if (ctx) { zo_ctx_free((zo_CCtx*)ctx); free(ctx); } }

size_t zo_compressCCtx(void* ctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level, int checksumFlag)
{
    zo_CCtx* const c = (zo_CCtx*)ctx;
    BYTE* const ostart = (BYTE*)dst; BYTE* op = ostart; size_t result;
    if (!c) return ERROR(memory_allocation);
    result = zo_ctx_alloc(c, srcSize, level);
    if (ERR_isError(result)) return result;
    /* window: byte 0 of src gets index 2 (ZSTD_window_init + first ZSTD_window_update, ZstdCompressInternal.cs:723-768) */
    c->ms.window.base = (const BYTE*)src - 2; c->ms.window.dictLimit = 2; c->ms.window.lowLimit = 2;
    /* ZSTD_compressContinue_internal :5013 */
    {   size_t const fhSize = ZSTD_writeFrameHeader(op, dstCapacity, c->cParams.windowLog, checksumFlag, srcSize, 0);
        if (ERR_isError(fhSize)) return fhSize;
        op += fhSize; dstCapacity -= fhSize; }
    if (srcSize) {
        size_t const cSize = ZSTD_compress_frameChunk(c, op, dstCapacity, src, srcSize);
        if (ERR_isError(cSize)) return cSize;
        op += cSize; dstCapacity -= cSize;
    }
    /* ZSTD_writeEpilogue :5598 */
    if (srcSize == 0) {   /* stage != ending : write one empty last raw block */
        U32 const cBlockHeader24 = 1 + (((U32)bt_raw) << 1);
        if (dstCapacity < 4) return ERROR(dstSize_tooSmall);
        MEM_writeLE24(op, cBlockHeader24); op += ZSTD_blockHeaderSize; dstCapacity -= ZSTD_blockHeaderSize;
    }
    if (checksumFlag) {
        U32 const checksum = (U32)zo_xxh64(src, srcSize, 0);
        if (dstCapacity < 4) return ERROR(dstSize_tooSmall);
        MEM_write32(op, checksum); op += 4;
    }
    return (size_t)(op - ostart);
}

size_t zo_compress_advanced(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level, int checksumFlag)
{
    void* const c = zo_createCCtx(); size_t r;
    if (!c) return ERROR(memory_allocation);
    r = zo_compressCCtx(c, dst, dstCapacity, src, srcSize, level, checksumFlag);
    zo_freeCCtx(c);
    return r;
}

size_t zo_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level)
{ return zo_compress_advanced(dst, dstCapacity, src, srcSize, level, 0); }

/* =====================================================================================
 *  Dictionary compression: Compressor.LoadDictionary (Compressor.cs:43-56) + Wrap.
 *  ZSTD_CCtx_loadDictionary (ZstdCompress.cs:1683) only stores the bytes; the first ZSTD_compress2 digests them into a CDict
 *  built with the context's level (ZSTD_initLocalDict :1581 -> ZSTD_createCDict_advanced2 :5933 -> ZSTD_initCDict_internal
 *  :5826 -> ZSTD_compress_insertDictionary :5467, dtlm_full), and every frame starts from that CDict, either ATTACHED (small
 *  inputs: the CDict's tables are searched in place, ZSTD_resetCCtx_byAttachingCDict :2746) or COPIED (its tables become the
 *  working tables and the content an external dictionary segment, ZSTD_resetCCtx_byCopyingCDict :2803).
 * ===================================================================================== */
typedef enum { zo_cpm_noAttachDict = 0, zo_cpm_attachDict = 1, zo_cpm_createCDict = 2 } zo_cpm_e;

static U32 zo_dictAndWindowLog(U32 windowLog, U64 srcSize, U64 dictSize)   /* :1985 */
{
    U64 const maxWindowSize = 1ULL << 31;
    if (dictSize == 0) return windowLog;
    {   U64 const windowSize = 1ULL << windowLog; U64 const dictAndWindowSize = dictSize + windowSize;
        if (windowSize >= dictSize + srcSize) return windowLog;
        if (dictAndWindowSize >= maxWindowSize) return 31;
        return BIT_highbit32((U32)dictAndWindowSize - 1) + 1;
    }
}
static cParams_t zo_adjustCParams_dict(cParams_t cPar, U64 srcSize, size_t dictSize, zo_cpm_e mode)   /* :2023 */
{
    U64 const minSrcSize = 513; U64 const maxWindowResize = 1ULL << 30;
    if (mode == zo_cpm_createCDict) { if (dictSize && srcSize == ZSTD_CONTENTSIZE_UNKNOWN) srcSize = minSrcSize; }
    else if (mode == zo_cpm_attachDict) dictSize = 0;
    if ((srcSize < maxWindowResize) && (dictSize < maxWindowResize)) {
        U32 const tSize = (U32)(srcSize + dictSize); U32 const hashSizeMin = 1 << 6;
        U32 const srcLog = (tSize < hashSizeMin) ? 6 : BIT_highbit32(tSize - 1) + 1;
        if (cPar.windowLog > srcLog) cPar.windowLog = srcLog;
    }
    if (srcSize != ZSTD_CONTENTSIZE_UNKNOWN) {
        U32 const dictAndWindowLog = zo_dictAndWindowLog(cPar.windowLog, srcSize, dictSize);
        U32 const cycleLog = cPar.chainLog;                    /* ZSTD_cycleLog: strategy < btlazy2 */
        if (cPar.hashLog > dictAndWindowLog + 1) cPar.hashLog = dictAndWindowLog + 1;
        if (cycleLog > dictAndWindowLog) cPar.chainLog -= (cycleLog - dictAndWindowLog);
    }
    if (cPar.windowLog < 10) cPar.windowLog = 10;
    return cPar;
}
/* ZSTD_getCParamsFromCCtxParams :2156 = ZSTD_getCParams_internal :7891 (row size :7852) + a second adjust */
static int zo_getCParams_dict(cParams_t* out, int level, U64 srcSizeHint, size_t dictSize, zo_cpm_e mode)
{
    size_t const rowDict = (mode == zo_cpm_attachDict) ? 0 : dictSize;
    int const unknown = (srcSizeHint == ZSTD_CONTENTSIZE_UNKNOWN);
    size_t const addedSize = unknown && rowDict > 0 ? 500 : 0;
    U64 const rSize = unknown && rowDict == 0 ? ZSTD_CONTENTSIZE_UNKNOWN : srcSizeHint + rowDict + addedSize;   /* wraps for `unknown` with a dictionary, as in the reference */
    U32 const tableID = (rSize <= 256 * 1024) + (rSize <= 128 * 1024) + (rSize <= 16 * 1024);
    int row;
    if (level == 0) row = 3; else if (level < 0) row = 0; else if (level > 4) return -1; else row = level;
    {   cParams_t cp = zo_defaultCParameters[tableID][row];
        if (cp.strategy == 0) return -1;
        if (level < 0) { int const clamped = level < -(1 << 17) ? -(1 << 17) : level; cp.targetLength = (U32)(-clamped); }
        *out = zo_adjustCParams_dict(cp, srcSizeHint, dictSize, mode);
    }
    *out = zo_adjustCParams_dict(*out, srcSizeHint, dictSize, mode);
    return 0;
}

/* HufCompress.cs:249 HUF_readCTable */
static size_t zo_HUF_readCTable(HUF_CTable* CTable, unsigned* maxSymbolValuePtr, const void* src, size_t srcSize, unsigned* hasZeroWeights)
{
    BYTE huffWeight[HUF_SYMBOLVALUE_MAX + 1]; U32 rankVal[HUF_TABLELOG_MAX + 2]; U32 tableLog = 0, nbSymbols = 0;
    size_t const readSize = zo_HUF_readStats(huffWeight, HUF_SYMBOLVALUE_MAX + 1, rankVal, &nbSymbols, &tableLog, src, srcSize);
    if (ERR_isError(readSize)) return readSize;
    *hasZeroWeights = (rankVal[0] > 0);
    if (tableLog > HUF_TABLELOG_MAX) return ERROR(tableLog_tooLarge);
    if (nbSymbols > *maxSymbolValuePtr + 1) return ERROR(maxSymbolValue_tooSmall);
    memset(CTable, 0, sizeof(*CTable));
    CTable->tableLog = tableLog;
    {   U32 n; for (n = 0; n < nbSymbols; n++) { U32 const w = huffWeight[n]; CTable->nbBits[n] = (BYTE)(w ? tableLog + 1 - w : 0); } }
    {   U16 nbPerRank[HUF_TABLELOG_MAX + 2] = { 0 }; U16 valPerRank[HUF_TABLELOG_MAX + 2] = { 0 };
        {   U32 n; for (n = 0; n < nbSymbols; n++) nbPerRank[CTable->nbBits[n]]++; }
        valPerRank[tableLog + 1] = 0;
        {   U16 min = 0; U32 n; for (n = tableLog; n > 0; n--) { valPerRank[n] = min; min = (U16)(min + nbPerRank[n]); min >>= 1; } }
        {   U32 n; for (n = 0; n < nbSymbols; n++) CTable->value[n] = valPerRank[CTable->nbBits[n]]++; }
    }
    *maxSymbolValuePtr = nbSymbols - 1;
    return readSize;
}

static FSE_repeat ZSTD_dictNCountRepeat(const S16* normalizedCounter, unsigned dictMaxSymbolValue, unsigned maxSymbolValue)   /* :5239 */
{
    U32 s;
    if (dictMaxSymbolValue < maxSymbolValue) return FSE_repeat_check;
    for (s = 0; s <= maxSymbolValue; ++s) if (normalizedCounter[s] == 0) return FSE_repeat_check;
    return FSE_repeat_valid;
}

/* ZstdCompress.cs:5264 ZSTD_loadCEntropy: header size, or an error */
static size_t ZSTD_loadCEntropy(ZSTD_compressedBlockState_t* bs, const void* const dict, size_t dictSize)
{
    S16 offcodeNCount[MaxOff + 1]; unsigned offcodeMaxValue = MaxOff;
    const BYTE* dictPtr = (const BYTE*)dict; const BYTE* const dictEnd = dictPtr + dictSize;
    dictPtr += 8;
    bs->entropy.huf.repeatMode = HUF_repeat_check;
    {   unsigned maxSymbolValue = 255; unsigned hasZeroWeights = 1;
        size_t const hufHeaderSize = zo_HUF_readCTable(&bs->entropy.huf.CTable, &maxSymbolValue, dictPtr, (size_t)(dictEnd - dictPtr), &hasZeroWeights);
        if (!hasZeroWeights) bs->entropy.huf.repeatMode = HUF_repeat_valid;
        if (ERR_isError(hufHeaderSize)) return ERROR(dictionary_corrupted);
        if (maxSymbolValue < 255) return ERROR(dictionary_corrupted);
        dictPtr += hufHeaderSize;
    }
    {   unsigned offcodeLog;
        size_t const offcodeHeaderSize = zo_FSE_readNCount(offcodeNCount, &offcodeMaxValue, &offcodeLog, dictPtr, (size_t)(dictEnd - dictPtr));
        if (ERR_isError(offcodeHeaderSize)) return ERROR(dictionary_corrupted);
        if (offcodeLog > OffFSELog) return ERROR(dictionary_corrupted);
        if (ERR_isError(FSE_buildCTable(&bs->entropy.fse.offcodeCTable, offcodeNCount, MaxOff, offcodeLog))) return ERROR(dictionary_corrupted);
        dictPtr += offcodeHeaderSize;
    }
    {   S16 matchlengthNCount[MaxML + 1]; unsigned matchlengthMaxValue = MaxML, matchlengthLog;
        size_t const h = zo_FSE_readNCount(matchlengthNCount, &matchlengthMaxValue, &matchlengthLog, dictPtr, (size_t)(dictEnd - dictPtr));
        if (ERR_isError(h)) return ERROR(dictionary_corrupted);
        if (matchlengthLog > MLFSELog) return ERROR(dictionary_corrupted);
        if (ERR_isError(FSE_buildCTable(&bs->entropy.fse.matchlengthCTable, matchlengthNCount, matchlengthMaxValue, matchlengthLog))) return ERROR(dictionary_corrupted);
        bs->entropy.fse.matchlength_repeatMode = ZSTD_dictNCountRepeat(matchlengthNCount, matchlengthMaxValue, MaxML);
        dictPtr += h;
    }
    {   S16 litlengthNCount[MaxLL + 1]; unsigned litlengthMaxValue = MaxLL, litlengthLog;
        size_t const h = zo_FSE_readNCount(litlengthNCount, &litlengthMaxValue, &litlengthLog, dictPtr, (size_t)(dictEnd - dictPtr));
        if (ERR_isError(h)) return ERROR(dictionary_corrupted);
        if (litlengthLog > LLFSELog) return ERROR(dictionary_corrupted);
        if (ERR_isError(FSE_buildCTable(&bs->entropy.fse.litlengthCTable, litlengthNCount, litlengthMaxValue, litlengthLog))) return ERROR(dictionary_corrupted);
        bs->entropy.fse.litlength_repeatMode = ZSTD_dictNCountRepeat(litlengthNCount, litlengthMaxValue, MaxLL);
        dictPtr += h;
    }
    if (dictPtr + 12 > dictEnd) return ERROR(dictionary_corrupted);
    bs->rep[0] = MEM_read32(dictPtr + 0); bs->rep[1] = MEM_read32(dictPtr + 4); bs->rep[2] = MEM_read32(dictPtr + 8);
    dictPtr += 12;
    {   size_t const dictContentSize = (size_t)(dictEnd - dictPtr); U32 offcodeMax = MaxOff;
        if (dictContentSize <= ((U32)-1) - 128 * 1024) { U32 const maxOffset = (U32)dictContentSize + 128 * 1024; offcodeMax = BIT_highbit32(maxOffset); }
        bs->entropy.fse.offcode_repeatMode = ZSTD_dictNCountRepeat(offcodeNCount, offcodeMaxValue, offcodeMax < MaxOff ? offcodeMax : MaxOff);
        {   U32 u; for (u = 0; u < 3; u++) { if (bs->rep[u] == 0) return ERROR(dictionary_corrupted); if (bs->rep[u] > dictContentSize) return ERROR(dictionary_corrupted); } }
    }
    return (size_t)(dictPtr - (const BYTE*)dict);
}

/* ZstdFast.cs:9 ZSTD_fillHashTable / ZstdDoubleFast.cs:9 ZSTD_fillDoubleHashTable, dtlm_full, from index `from` to `end` */
static void ZSTD_fillHashTable_full(ZSTD_matchState_t* ms, U32 from, const BYTE* end)
{
    U32* const hashTable = ms->hashTable; U32 const hBits = ms->cParams.hashLog; U32 const mls = ms->cParams.minMatch;
    const BYTE* const base = ms->window.base; const BYTE* ip = base + from; const BYTE* const iend = end - 8;
    U32 const fastHashFillStep = 3;
    for (; ip + fastHashFillStep < iend + 2; ip += fastHashFillStep) {
        U32 const curr = (U32)(ip - base); U32 p;
        hashTable[ZSTD_hashPtr(ip, hBits, mls)] = curr;
        for (p = 1; p < fastHashFillStep; ++p) { size_t const hash = ZSTD_hashPtr(ip + p, hBits, mls); if (hashTable[hash] == 0) hashTable[hash] = curr + p; }
    }
}
static void ZSTD_fillDoubleHashTable_full(ZSTD_matchState_t* ms, U32 from, const BYTE* end)
{
    U32* const hashLarge = ms->hashTable; U32 const hBitsL = ms->cParams.hashLog; U32 const mls = ms->cParams.minMatch;
    U32* const hashSmall = ms->chainTable; U32 const hBitsS = ms->cParams.chainLog;
    const BYTE* const base = ms->window.base; const BYTE* ip = base + from; const BYTE* const iend = end - 8;
    U32 const fastHashFillStep = 3;
    for (; ip + fastHashFillStep - 1 <= iend; ip += fastHashFillStep) {
        U32 const curr = (U32)(ip - base); U32 i;
        for (i = 0; i < fastHashFillStep; ++i) {
            size_t const smHash = ZSTD_hashPtr(ip + i, hBitsS, mls); size_t const lgHash = ZSTD_hashPtr(ip + i, hBitsL, 8);
            if (i == 0) hashSmall[smHash] = curr + i;
            if (i == 0 || hashLarge[lgHash] == 0) hashLarge[lgHash] = curr + i;
        }
    }
}

static const size_t zo_attachDictSizeCutoffs[3] = { 8 * 1024, 8 * 1024, 16 * 1024 };   /* :2725, index = strategy (fast 1, dfast 2) */

/* One frame with a loaded dictionary. */
static size_t zo_compress_dict_internal(void* dst, size_t dstCapacity, const void* src, size_t srcSize, const void* dict, size_t dictSize,
                                        int level, int checksumFlag)
{
    zo_CCtx* const c = (zo_CCtx*)calloc(1, sizeof(zo_CCtx));
    ZSTD_matchState_t cdictMs; cParams_t cdictCP, reqCP, workCP;
    ZSTD_compressedBlockState_t cdictBs; U32 dictID = 0;
    const BYTE* content = (const BYTE*)dict; size_t contentLen = dictSize;
    BYTE* V = NULL; U32* cdictHash = NULL; U32* cdictChain = NULL; size_t result;
    BYTE* const ostart = (BYTE*)dst; BYTE* op = ostart;
    if (!c) return ERROR(memory_allocation);
    memset(&cdictMs, 0, sizeof(cdictMs));
#define ZO_DICT_FAIL(e) do { result = (e); goto _cleanup; } while (0)
    /* ---- the CDict: parameters for an unknown source size (ZSTD_cpm_createCDict) ---- */
    if (zo_getCParams_dict(&cdictCP, level, ZSTD_CONTENTSIZE_UNKNOWN, dictSize, zo_cpm_createCDict)) ZO_DICT_FAIL(ERROR(parameter_unsupported));
    ZSTD_reset_compressedBlockState(&cdictBs);
    if (dictSize >= 8 && MEM_read32(dict) == 0xEC30A437U) {      /* ZSTD_loadZstdDictionary :5424 */
        size_t const eSize = ZSTD_loadCEntropy(&cdictBs, dict, dictSize);
        if (ERR_isError(eSize)) ZO_DICT_FAIL(ERROR(memory_allocation));   /* ZSTD_createCDict_advanced2 returns NULL and ZSTD_initLocalDict (:1604) reports memory_allocation */
        dictID = MEM_read32((const BYTE*)dict + 4);
        content = (const BYTE*)dict + eSize; contentLen = dictSize - eSize;
    } else if (dictSize < 8) contentLen = 0;                     /* ZSTD_compress_insertDictionary :5467: nothing is loaded */
    V = (BYTE*)malloc(contentLen + srcSize + 32);
    cdictHash = (U32*)calloc((size_t)1 << cdictCP.hashLog, sizeof(U32));
    cdictChain = (U32*)calloc((size_t)1 << cdictCP.chainLog, sizeof(U32));
    if (!V || !cdictHash || !cdictChain) ZO_DICT_FAIL(ERROR(memory_allocation));
    memcpy(V, content, contentLen); memcpy(V + contentLen, src, srcSize); memset(V + contentLen + srcSize, 0, 32);
    cdictMs.cParams = cdictCP; cdictMs.hashTable = cdictHash; cdictMs.chainTable = cdictChain;
    cdictMs.window.base = V - 2; cdictMs.window.dictLimit = 2; cdictMs.window.lowLimit = 2;      /* ZSTD_window_init + first update */
    cdictMs.loadedDictEnd = contentLen ? (U32)(2 + contentLen) : 0;                               /* ZSTD_loadDictionaryContent :5126 */
    if (contentLen > 8) {
        if (cdictCP.strategy == ZSTD_fast) ZSTD_fillHashTable_full(&cdictMs, 2, V + contentLen);
        else ZSTD_fillDoubleHashTable_full(&cdictMs, 2, V + contentLen);
    }
    /* ---- the frame: ZSTD_CCtx_init_compressStream2 :6949, ZSTD_compressBegin_internal :5505 (cdict->compressionLevel == 0: always by CDict) ---- */
    {   int const attach = srcSize <= zo_attachDictSizeCutoffs[cdictCP.strategy];                 /* ZSTD_shouldAttachDict :2738 */
        U32 const cdictEnd = (U32)(2 + contentLen);
        if (zo_getCParams_dict(&reqCP, level, srcSize, dictSize, attach ? zo_cpm_attachDict : zo_cpm_noAttachDict)) ZO_DICT_FAIL(ERROR(parameter_unsupported));
        if (attach) { workCP = zo_adjustCParams_dict(cdictCP, srcSize, dictSize, zo_cpm_attachDict); workCP.windowLog = reqCP.windowLog; }
        else { workCP = cdictCP; workCP.windowLog = reqCP.windowLog; }
        c->cParams = workCP;
        CHECK_F(zo_ctx_buffers(c));
        {   size_t const windowSize = (size_t)1 << workCP.windowLog;
            size_t ws = windowSize < srcSize ? windowSize : srcSize; if (ws < 1) ws = 1;
            c->blockSize = ws < ZSTD_BLOCKSIZE_MAX ? ws : ZSTD_BLOCKSIZE_MAX; }
        c->seqStore.maxNbSeq = c->blockSize / ((workCP.minMatch == 3) ? 3 : 4); c->seqStore.maxNbLit = c->blockSize;
        c->ms.cParams = workCP;
        c->prevCBlock = &c->blockStateA; c->nextCBlock = &c->blockStateB;
        memcpy(c->prevCBlock, &cdictBs, sizeof(cdictBs));
        c->isFirstBlock = 1;
        if (attach) {
            memset(c->ms.hashTable, 0, ((size_t)1 << workCP.hashLog) * sizeof(U32));
            if (workCP.strategy != ZSTD_fast) memset(c->ms.chainTable, 0, ((size_t)1 << workCP.chainLog) * sizeof(U32));
            if (contentLen == 0) {                                                                  /* cdictLen == 0: nothing to attach */
                c->ms.window.base = V - 2; c->ms.window.dictLimit = 2; c->ms.window.lowLimit = 2;
            } else {
                c->ms.dictMatchState = &cdictMs; c->ms.dictEndIndex = cdictEnd;
                c->ms.window.base = V - 2; c->ms.window.dictLimit = cdictEnd; c->ms.window.lowLimit = cdictEnd;   /* ZSTD_window_clear, then the update */
                c->ms.loadedDictEnd = cdictEnd;
            }
        } else {
            memcpy(c->ms.hashTable, cdictHash, ((size_t)1 << workCP.hashLog) * sizeof(U32));
            if (workCP.strategy != ZSTD_fast) memcpy(c->ms.chainTable, cdictChain, ((size_t)1 << workCP.chainLog) * sizeof(U32));
            c->ms.loadedDictEnd = cdictMs.loadedDictEnd;
            /* window copied from the CDict, then ZSTD_window_update with the (non contiguous) source: the content becomes the extDict */
            c->ms.window.base = V - 2; c->ms.window.lowLimit = 2; c->ms.window.dictLimit = cdictEnd;
            if (c->ms.window.dictLimit - c->ms.window.lowLimit < 8) c->ms.window.lowLimit = c->ms.window.dictLimit;
        }
    }
    {   size_t const fhSize = ZSTD_writeFrameHeader(op, dstCapacity, workCP.windowLog, checksumFlag, srcSize, dictID);
        if (ERR_isError(fhSize)) ZO_DICT_FAIL(fhSize);
        op += fhSize; dstCapacity -= fhSize; }
    if (srcSize) {
        size_t const cSize = ZSTD_compress_frameChunk(c, op, dstCapacity, V + contentLen, srcSize);
        if (ERR_isError(cSize)) ZO_DICT_FAIL(cSize);
        op += cSize; dstCapacity -= cSize;
    } else {
        if (dstCapacity < 4) ZO_DICT_FAIL(ERROR(dstSize_tooSmall));
        MEM_writeLE24(op, 1 + (((U32)bt_raw) << 1)); op += ZSTD_blockHeaderSize; dstCapacity -= ZSTD_blockHeaderSize;
    }
    if (checksumFlag) {
        if (dstCapacity < 4) ZO_DICT_FAIL(ERROR(dstSize_tooSmall));
        MEM_write32(op, (U32)zo_xxh64(src, srcSize, 0)); op += 4;
    }
    result = (size_t)(op - ostart);
_cleanup:
    free(V); free(cdictHash); free(cdictChain);
    zo_ctx_free(c); free(c);
    return result;
#undef ZO_DICT_FAIL
}

size_t zo_compress_usingLoadedDict(void* dst, size_t dstCapacity, const void* src, size_t srcSize, const void* dict, size_t dictSize, int level, int checksumFlag)
{
    if (dict == NULL || dictSize == 0) return zo_compress_advanced(dst, dstCapacity, src, srcSize, level, checksumFlag);   /* ZSTD_CCtx_loadDictionary(NULL, 0) clears the dictionary */
    return zo_compress_dict_internal(dst, dstCapacity, src, srcSize, dict, dictSize, level, checksumFlag);
}

size_t zo_matchfinder_block(int level, const void* src, size_t srcSize, zo_seqDef* seqs, uint8_t* lits, size_t* litSize,
                            uint32_t longLength[2], uint32_t repOut[3])
{
    zo_CCtx* const c = (zo_CCtx*)calloc(1, sizeof(zo_CCtx)); size_t nbSeq;
    if (!c) return ERROR(memory_allocation);
    if (srcSize > ZSTD_BLOCKSIZE_MAX) { free(c); return ERROR(srcSize_wrong); }
    {   size_t const r = zo_ctx_alloc(c, srcSize, level); if (ERR_isError(r)) { free(c); return r; } }
    c->ms.window.base = (const BYTE*)src - 2; c->ms.window.dictLimit = 2; c->ms.window.lowLimit = 2;
    if (ZSTD_buildSeqStore(c, src, srcSize)) { nbSeq = 0; *litSize = 0; longLength[0] = longLength[1] = 0; memcpy(repOut, c->prevCBlock->rep, 12); }
    else {
        nbSeq = (size_t)(c->seqStore.sequences - c->seqStore.sequencesStart);
        memcpy(seqs, c->seqStore.sequencesStart, nbSeq * sizeof(zo_seqDef));
        *litSize = (size_t)(c->seqStore.lit - c->seqStore.litStart);
        memcpy(lits, c->seqStore.litStart, *litSize);
        longLength[0] = (U32)c->seqStore.longLengthType; longLength[1] = c->seqStore.longLengthPos;
        memcpy(repOut, c->nextCBlock->rep, 12);
    }
    zo_ctx_free(c); free(c);
    return nbSeq;
}
