// zb_decode.cuh -- device-side data layout of the batch decoder (product code).
//
// HBM layout for one decode pass over `n` items (an item = one caller buffer holding >= 1 concatenated frames):
//   DecItem      items[n]                  persistent per-item cursor + current-block descriptor
//   uint16_t     hufTable[n][4096]         single-symbol Huffman table, entry = byte<<8 | nbBits (HufDecompress.cs:80)
//   uint32_t     fseTable[n][1280]         compact LL(512) | ML(512) | OF(256) sequence tables (ZstdDecompressBlock.cs:1571)
//   uint8_t      litBuf[n][kLitStride]     regenerated literals, 4 segments each padded to 16 B
//   uint2        seq[n][kSeqCap]           decoded (litLength, matchLength, offset) records packed into 8 bytes
// Blocks of a frame are processed in "waves": wave k handles the k-th block of every item that still has one, so
// repeat-mode tables, rep codes and the window carry over through the persistent per-item state.
#pragma once
#include "zb_common.cuh"

namespace zb {

constexpr uint32_t kHufTableEntries = 1u << kHufTableLogMax;       // 4096 x u16
constexpr uint32_t kFseLLOff = 0, kFseMLOff = 512, kFseOFOff = 1024, kFseTableEntries = 1280;
constexpr uint32_t kLitStride = kBlockSizeMax + 64;                // 4 x 15 B of segment padding fits

enum : uint32_t { kStRunning = 0, kStDone = 1, kStError = 2 };
enum : uint32_t { kBlkRaw = 0, kBlkRle = 1, kBlkCompressed = 2, kBlkNone = 3 };
enum : uint32_t { kLitRaw = 0, kLitRle = 1, kLitHuf = 2 };

// Compact FSE decode entry: [0,4) nbBits | [4,9) nbAdditionalBits | [9,15) symbol | [16,26) nextState base
__host__ __device__ inline uint32_t fse_pack(uint32_t nbBits, uint32_t addBits, uint32_t sym, uint32_t next)
{ return nbBits | (addBits << 4) | (sym << 9) | (next << 16); }

// One decoded sequence in 8 bytes: litLength (17 bits) | matchLength-3 (17 bits) | offset (30 bits).  The sequence
// decoder rejects values outside these ranges (they cannot occur in a block that regenerates <= 128 KiB of a frame < 1 GiB).
__host__ __device__ inline uint2 seq_pack(uint32_t ll, uint32_t ml, uint32_t of)
{ uint32_t const m = ml - 3; uint2 r; r.x = ll | (m << 17); r.y = (m >> 15) | (of << 2); return r; }
__host__ __device__ inline void seq_unpack(uint2 r, uint32_t& ll, uint32_t& ml, uint32_t& of)
{ ll = r.x & 0x1FFFFu; ml = ((r.x >> 17) | ((r.y & 3u) << 15)) + 3u; of = r.y >> 2; }

struct __align__(16) DecItem {
    // ---- static for the pass ----
    uint64_t srcOff;        // byte offset of the item inside the packed source buffer
    uint64_t dstOff;        // byte offset of the item's output inside the destination buffer
    uint32_t srcSize;
    uint32_t dstCap;
    // ---- cursor (persists across waves) ----
    uint32_t srcPos;        // next unread byte of the item
    uint32_t outPos;        // bytes regenerated so far (all frames of the item)
    uint32_t frameStart;    // outPos at which the current frame began (prefixStart)
    uint32_t status;        // kStRunning / kStDone / kStError
    uint32_t errCode;
    uint32_t inFrame;
    uint32_t moreThan1Frame;
    uint32_t checksumFlag;
    uint32_t hasFcs;
    uint32_t hufX2;         // the current Huffman table is the reference's double-symbol form (HUF_selectDecoder / dictionary): only the
                            // verdict on damaged streams depends on it (dec_huf_kernel, huf_x2_replay)
    uint64_t fcs;
    uint32_t rep[3];
    uint32_t litEntropy;    // a Huffman table from an earlier block of this frame is available
    uint32_t fseEntropy;    // sequence tables from an earlier block of this frame are available
    uint32_t hufLog;
    uint32_t llLog, ofLog, mlLog;
    // ---- descriptor of the block handled in the current wave ----
    uint32_t blkType;
    uint32_t blkSrcOff;     // offset (in item) of the block content
    uint32_t blkSize;       // raw: byte count; rle: regenerated size; compressed: compressed size
    uint32_t lastBlock;
    uint32_t litType;
    uint32_t litSize;
    uint32_t litOff;        // raw: offset of literal bytes; rle: offset of the repeated byte
    uint32_t nStreams;
    uint32_t streamOff[4];
    uint32_t streamLen[4];
    uint32_t nbSeq;
    uint32_t seqOff;
    uint32_t seqLen;
    uint32_t blockOut;      // regenerated size of this block (literals-only part added by exec)
    uint32_t seqLitEnd;     // literals consumed by the sequences (set by seq_decode)
    uint32_t prefix;        // bytes of dictionary content that sit in front of the current frame's output (0 without a dictionary)
    uint32_t blkLimit;      // outPos up to which the sequences of this block may write: the reference keeps the block's literals inside
                            // dst when there is room (ZSTD_in_dst, at +128 KiB + 32) and its sequences stop there (ZstdDecompressBlock.cs:44-73, :2668)
    uint32_t deferErr;      // an error found behind a Huffman literal section: the reference decodes the literals first, so it only counts
                            // if they turn out to be sound (dec_huf_kernel reports it)
};

// Dictionary of a decode context (ZSTD_decompress_insertDictionary, ZstdDecompress.cs:1880): parsed once on the device.
// info[]: [0] 1 = usable, 2 = corrupted | [1] dictID | [2] entropy tables present | [3] hufLog | [4] llLog | [5] ofLog | [6] mlLog |
//         [7..9] repcodes | [10] offset of the content inside the dictionary | [11] content size
constexpr uint32_t kDictInfoWords = 16;

struct DecPass {
    DecItem* items;
    uint32_t nItems;
    const uint8_t* src;     // packed source bytes
    uint8_t* dst;           // destination bytes
    uint16_t* hufTable;
    uint32_t* fseTable;
    uint8_t* litBuf;
    uint2* seq;             // decoded sequences, packed by seq_pack()
    const uint32_t* defaultFse;  // predefined LL|ML|OF tables (ZstdDecompressBlock.cs:398/:857/:1092)
    uint32_t* hufList;      // item indices that need literal decoding this wave
    uint32_t* seqList;      // item indices that need sequence decoding this wave
    uint32_t* rawList;      // item indices whose block of this wave is raw or RLE (copied / filled by dec_rawcopy_kernel)
    uint32_t* counters;     // [0] hufCount [1] seqCount [2] running items after this wave [3] max blocks (scan) [4] rawCount
    uint64_t* results;      // per item: regenerated size or error code
    // dictionary (all null / 0 without one): tables in the layout of hufTable / fseTable, the raw dictionary bytes, info words
    const uint16_t* dictHuf; const uint32_t* dictFse; const uint8_t* dictBytes; const uint32_t* dictInfo;
};

// Number of blocks of the longest frame chain of an item = number of waves the item needs.  Walks headers like
// ZSTD_findFrameSizeInfo (ZstdDecompress.cs:877); malformed input simply stops the walk (errors are diagnosed by the
// setup kernel).  Shared by the device scan kernel and the host scheduler (which then needs no read-back).
__host__ __device__ inline uint32_t count_item_blocks(const uint8_t* src, uint32_t size)
{
    auto le32 = [](const uint8_t* q) { return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24); };
    uint32_t pos = 0, blocks = 0;
    for (;;) {
        uint32_t const rem = size - pos;
        if (rem < 5) break;
        uint32_t const magic = le32(src + pos);
        if ((magic & kMagicSkippableMask) == kMagicSkippableStart) {
            if (rem < 8) { blocks++; break; }
            uint32_t const sz = le32(src + pos + 4);
            if ((uint64_t)sz + 8 > rem) { blocks++; break; }
            pos += sz + 8; continue;
        }
        if (magic != kMagic) { blocks++; break; }                  // one more wave: dec_setup diagnoses what the walk stops at
        uint32_t const fhd = src[pos + 4];
        uint32_t const dictID = fhd & 3, single = (fhd >> 5) & 1, fcsId = fhd >> 6;
        uint32_t const hs = 5 + !single + (dictID == 3 ? 4 : dictID) + (fcsId == 0 ? 0 : (1u << fcsId)) + (single && !fcsId);
        if (rem < hs + 3) { blocks++; break; }
        pos += hs;
        bool bad = false;
        for (;;) {
            if (size - pos < 3) { bad = true; break; }
            uint32_t const h = (uint32_t)src[pos] | ((uint32_t)src[pos + 1] << 8) | ((uint32_t)src[pos + 2] << 16);
            uint32_t const type = (h >> 1) & 3, cs = h >> 3;
            uint32_t const csz = type == kBlkRle ? 1 : cs;
            blocks++;
            pos += 3;
            if (type == 3 || csz > size - pos) { bad = true; break; }
            pos += csz;
            if (h & 1) break;
        }
        if (bad) { blocks++; break; }
        if (fhd & 4) { if (size - pos < 4) break; pos += 4; }
    }
    return blocks ? blocks : 1;
}

// host-callable launchers (zb_decode.cu)
void dec_build_default_tables(uint32_t* d_defaultFse, cudaStream_t s);
struct DecItemInit { uint64_t srcOff; uint64_t dstOff; uint32_t srcSize; uint32_t dstCap; };
void dec_launch_scan_init(const DecPass& p, const void* d_init, cudaStream_t s);
void dec_launch_wave(const DecPass& p, cudaStream_t s);
void dec_launch_wave_timed(const DecPass& p, cudaStream_t s, cudaEvent_t* ev5);
void dec_launch_finish(const DecPass& p, cudaStream_t s);
// parses a dictionary that already sits in device memory (one thread; a one-off per ZSTD_DCtx_loadDictionary)
constexpr uint32_t kEncDictStatsWords = 258 + 3 * 66;     // what the encode side's CDict needs from the parse (dec_dict_kernel)
void dec_launch_dict_setup(const uint8_t* d_dict, uint32_t dictSize, uint16_t* d_huf, uint32_t* d_fse, uint32_t* d_info, cudaStream_t s, uint32_t* d_encStats = nullptr);
// copies the dictionary content in front of every item's output slot (the caller left p.dictInfo[11] bytes of headroom there)
void dec_launch_dict_prefill(const DecPass& p, uint32_t contentOff, uint32_t contentSize, cudaStream_t s);

}  // namespace zb
