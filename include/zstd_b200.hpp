/* zstd_b200.hpp — C++17 host-side mirror of ZstdSharp's managed wrapper classes over the C ABI of zstd_b200.h.
 *
 * The reference's host side is compiled C# (no .NET in this image), so the host layer above the C ABI is C++:
 * same class names, method names, argument meaning and error behaviour as
 *   src/ZstdSharp/Compressor.cs:6-150      (Compressor: Level, SetParameter, Wrap, TryWrap, GetCompressBound, Dispose)
 *   src/ZstdSharp/Decompressor.cs:6-140    (Decompressor: LoadDictionary, GetDecompressedSize, Unwrap, TryUnwrap, Dispose)
 *   src/ZstdSharp/ThrowHelper.cs:10-41     (EnsureZstdSuccess / EnsureContentSizeOk -> ZstdException(code, message))
 *   src/ZstdSharp/ZstdException.cs         (ZstdException.Code)
 * plus WrapBatch / UnwrapBatch over ZSTDB200_compressBatch / ZSTDB200_decompressBatch (many independent frames per call,
 * the shape the GPU is built for).  Header-only; link with -lzstdb200.  Nothing here computes: every byte of a frame is
 * produced or consumed by the CUDA kernels behind the C ABI, and a missing GPU surfaces as ZstdException(GENERIC).
 * One context per thread, as in the reference (ZstdNetTests.cs:498-522).
 */
#ifndef ZSTD_B200_HPP
#define ZSTD_B200_HPP

#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
#include <istream>
#include <ostream>
#include <cstring>
#include <algorithm>

#include "zstd_b200.h"

namespace ZstdSharp {

/* ZSTD_ErrorCode values used by the wrapper (ZstdSharp.Unsafe.ZSTD_ErrorCode; same numbers as zstd_errors.h 1.5.1). */
enum class ZSTD_ErrorCode : int {
    no_error = 0, GENERIC = 1, prefix_unknown = 10, version_unsupported = 12, frameParameter_unsupported = 14,
    frameParameter_windowTooLarge = 16, corruption_detected = 20, checksum_wrong = 22, dictionary_corrupted = 30,
    dictionary_wrong = 32, dictionaryCreation_failed = 34, parameter_unsupported = 40, parameter_outOfBound = 42,
    tableLog_tooLarge = 44, maxSymbolValue_tooLarge = 46, maxSymbolValue_tooSmall = 48, stage_wrong = 60,
    init_missing = 62, memory_allocation = 64, workSpace_tooSmall = 66, dstSize_tooSmall = 70, srcSize_wrong = 72,
    dstBuffer_null = 74
};

/* ZSTD_cParameter members this library accepts (Compressor.SetParameter); others answer parameter_unsupported. */
enum class ZSTD_cParameter : int {
    ZSTD_c_compressionLevel = 100, ZSTD_c_contentSizeFlag = 200, ZSTD_c_checksumFlag = 201,
    independentChunks = ZSTDB200_c_independentChunks   /* enum member cannot share the macro's name */
};

/* ZSTD_dParameter (Decompressor.SetParameter / GetParameter). */
enum class ZSTD_dParameter : int { ZSTD_d_windowLogMax = 100 };

class ZstdException : public std::runtime_error {
public:
    ZstdException(ZSTD_ErrorCode code, const std::string& message) : std::runtime_error(message), Code(code) {}
    const ZSTD_ErrorCode Code;
};

class ObjectDisposedException : public std::logic_error {
public:
    explicit ObjectDisposedException(const char* name) : std::logic_error(name) {}
};

namespace ThrowHelper {
/* ThrowHelper.cs:10-16, 36-40: an error return becomes ZstdException(code = 0 - returnValue, ZSTD_getErrorName). */
inline size_t EnsureZstdSuccess(size_t returnValue) {
    if (ZSTD_isError(returnValue))
        throw ZstdException(static_cast<ZSTD_ErrorCode>(static_cast<int>(size_t(0) - returnValue)),
                            ZSTD_getErrorName(returnValue));
    return returnValue;
}
/* ThrowHelper.cs:26-35. */
inline unsigned long long EnsureContentSizeOk(unsigned long long returnValue) {
    if (returnValue == 0ULL - 1)
        throw ZstdException(ZSTD_ErrorCode::GENERIC, "Decompressed content size is not specified");
    if (returnValue == 0ULL - 2)
        throw ZstdException(ZSTD_ErrorCode::GENERIC,
                            "Decompressed content size cannot be determined (e.g. invalid magic number, srcSize too small)");
    return returnValue;
}
inline bool IsDstSizeTooSmall(size_t returnValue) {
    return returnValue == size_t(0) - size_t(ZSTD_ErrorCode::dstSize_tooSmall);
}
}  // namespace ThrowHelper

/* Result of one item of a batch call: Size when Code == no_error. */
struct BatchResult {
    size_t Size;
    ZSTD_ErrorCode Code;
};

namespace detail {
inline std::vector<BatchResult> ToResults(const std::vector<size_t>& raw) {
    std::vector<BatchResult> out(raw.size());
    for (size_t i = 0; i < raw.size(); ++i) {
        if (ZSTD_isError(raw[i])) out[i] = {0, static_cast<ZSTD_ErrorCode>(static_cast<int>(size_t(0) - raw[i]))};
        else out[i] = {raw[i], ZSTD_ErrorCode::no_error};
    }
    return out;
}
}  // namespace detail

/* Compressor.cs:6-150. */
class Compressor {
public:
    static constexpr int DefaultCompressionLevel = 0;                                  /* Compressor.cs:10 */
    static constexpr int MinCompressionLevel = -(1 << 17), MaxCompressionLevel = 4;    /* ZSTD_fast / ZSTD_dfast levels (4: only where it is ZSTD_dfast) */

    explicit Compressor(int level = DefaultCompressionLevel) : cctx_(ZSTD_createCCtx()) {   /* Compressor.cs:58-65 */
        if (!cctx_) throw ZstdException(ZSTD_ErrorCode::GENERIC, "Failed to create cctx");
        try { Level(level); } catch (...) { ZSTD_freeCCtx(cctx_); cctx_ = nullptr; throw; }
    }
    ~Compressor() { Dispose(); }
    Compressor(const Compressor&) = delete;
    Compressor& operator=(const Compressor&) = delete;

    int Level() const { return level_; }
    void Level(int value) {                                                            /* Compressor.cs:16-27 */
        if (level_ != value) {
            SetParameter(ZSTD_cParameter::ZSTD_c_compressionLevel, value);
            level_ = value;
        }
    }
    void SetParameter(ZSTD_cParameter parameter, int value) {                          /* Compressor.cs:29-33 */
        EnsureNotDisposed();
        ThrowHelper::EnsureZstdSuccess(ZSTD_CCtx_setParameter(cctx_, static_cast<int>(parameter), value));
    }

    /* LoadDictionary(byte[] dict): null / empty drops the dictionary  (Compressor.cs:43-56). */
    void LoadDictionary(const void* dict, size_t dictLength) {
        EnsureNotDisposed();
        ThrowHelper::EnsureZstdSuccess(ZSTD_CCtx_loadDictionary(cctx_, dict, dict ? dictLength : 0));
    }
    int GetParameter(ZSTD_cParameter parameter) const {                                /* Compressor.cs:35-41 */
        EnsureNotDisposed();
        int value = 0;
        ThrowHelper::EnsureZstdSuccess(ZSTD_CCtx_getParameter(cctx_, static_cast<int>(parameter), &value));
        return value;
    }

    static int GetCompressBound(int length) { return static_cast<int>(ZSTD_compressBound(static_cast<size_t>(length))); }
    static uint64_t GetCompressBoundLong(uint64_t length) { return ZSTD_compressBound(static_cast<size_t>(length)); }

    /* Span<byte> Wrap(ReadOnlySpan<byte> src)  (Compressor.cs:78-83) */
    std::vector<uint8_t> Wrap(const void* src, size_t srcLength) {
        std::vector<uint8_t> dest(ZSTD_compressBound(srcLength));
        dest.resize(Wrap(src, srcLength, dest.data(), dest.size()));
        return dest;
    }
    std::vector<uint8_t> Wrap(const std::vector<uint8_t>& src) { return Wrap(src.data(), src.size()); }
    /* int Wrap(ReadOnlySpan<byte> src, Span<byte> dest)  (Compressor.cs:88-96) */
    size_t Wrap(const void* src, size_t srcLength, void* dest, size_t destLength) {
        EnsureNotDisposed();
        return ThrowHelper::EnsureZstdSuccess(ZSTD_compress2(cctx_, dest, destLength, src, srcLength));
    }
    /* bool TryWrap(src, dest, out written)  (Compressor.cs:107-126): false only for dstSize_tooSmall. */
    bool TryWrap(const void* src, size_t srcLength, void* dest, size_t destLength, size_t& written) {
        EnsureNotDisposed();
        size_t returnValue = ZSTD_compress2(cctx_, dest, destLength, src, srcLength);
        if (ThrowHelper::IsDstSizeTooSmall(returnValue)) { written = 0; return false; }
        written = ThrowHelper::EnsureZstdSuccess(returnValue);
        return true;
    }

    /* Many independent inputs in one pass over the GPU (ZSTDB200_compressBatch): item i becomes exactly the frame
     * Wrap(src[i]) returns.  A failing item carries its own code and does not fail the others; only a library-level
     * failure (no device, out of memory) throws. */
    std::vector<BatchResult> WrapBatch(const std::vector<const void*>& src, const std::vector<size_t>& srcLength,
                                       const std::vector<void*>& dest, const std::vector<size_t>& destLength) {
        EnsureNotDisposed();
        size_t n = src.size();
        if (srcLength.size() != n || dest.size() != n || destLength.size() != n)
            throw std::invalid_argument("WrapBatch: array lengths differ");
        std::vector<size_t> raw(n);
        ThrowHelper::EnsureZstdSuccess(ZSTDB200_compressBatch(cctx_, n, level_ == 0 ? 3 : level_, src.data(), srcLength.data(),
                                                              dest.data(), destLength.data(), raw.data()));
        return detail::ToResults(raw);
    }

    void Dispose() {                                                                   /* Compressor.cs:141-150 */
        if (cctx_) { ZSTD_freeCCtx(cctx_); cctx_ = nullptr; }
    }
    ZSTD_CCtx* Handle() const { return cctx_; }

private:
    void EnsureNotDisposed() const { if (!cctx_) throw ObjectDisposedException("Compressor"); }
    ZSTD_CCtx* cctx_;
    int level_ = DefaultCompressionLevel;
};

/* Decompressor.cs:6-140. */
class Decompressor {
public:
    Decompressor() : dctx_(ZSTD_createDCtx()) {                                        /* Decompressor.cs:10-15 */
        if (!dctx_) throw ZstdException(ZSTD_ErrorCode::GENERIC, "Failed to create dctx");
    }
    ~Decompressor() { Dispose(); }
    Decompressor(const Decompressor&) = delete;
    Decompressor& operator=(const Decompressor&) = delete;

    void SetParameter(ZSTD_dParameter parameter, int value) {                          /* Decompressor.cs:22-26 */
        EnsureNotDisposed();
        ThrowHelper::EnsureZstdSuccess(ZSTD_DCtx_setParameter(dctx_, static_cast<int>(parameter), value));
    }
    int GetParameter(ZSTD_dParameter parameter) const {                                /* Decompressor.cs:28-34 */
        EnsureNotDisposed();
        int value = 0;
        ThrowHelper::EnsureZstdSuccess(ZSTD_DCtx_getParameter(dctx_, static_cast<int>(parameter), &value));
        return value;
    }

    /* LoadDictionary(byte[] dict): null / empty drops the dictionary  (Decompressor.cs:36-48). */
    void LoadDictionary(const void* dict, size_t dictLength) {
        EnsureNotDisposed();
        ThrowHelper::EnsureZstdSuccess(ZSTD_DCtx_loadDictionary(dctx_, dict, dict ? dictLength : 0));
    }

    /* Decompressor.cs:50-54: ZSTD_decompressBound over all frames of src. */
    static uint64_t GetDecompressedSize(const void* src, size_t srcLength) {
        return ThrowHelper::EnsureContentSizeOk(ZSTD_decompressBound(src, srcLength));
    }

    /* Span<byte> Unwrap(ReadOnlySpan<byte> src, int maxDecompressedSize = int.MaxValue)  (Decompressor.cs:62-75) */
    std::vector<uint8_t> Unwrap(const void* src, size_t srcLength,
                                int maxDecompressedSize = std::numeric_limits<int>::max()) {
        uint64_t expectedDstSize = GetDecompressedSize(src, srcLength);
        if (expectedDstSize > static_cast<uint64_t>(maxDecompressedSize))
            throw ZstdException(ZSTD_ErrorCode::dstSize_tooSmall,
                                "Decompressed content size " + std::to_string(expectedDstSize) +
                                    " is greater than maxDecompressedSize " + std::to_string(maxDecompressedSize));
        std::vector<uint8_t> dest(expectedDstSize);
        dest.resize(Unwrap(src, srcLength, dest.data(), dest.size()));
        return dest;
    }
    std::vector<uint8_t> Unwrap(const std::vector<uint8_t>& src, int maxDecompressedSize = std::numeric_limits<int>::max()) {
        return Unwrap(src.data(), src.size(), maxDecompressedSize);
    }
    /* int Unwrap(ReadOnlySpan<byte> src, Span<byte> dest)  (Decompressor.cs:80-88) */
    size_t Unwrap(const void* src, size_t srcLength, void* dest, size_t destLength) {
        EnsureNotDisposed();
        return ThrowHelper::EnsureZstdSuccess(ZSTD_decompressDCtx(dctx_, dest, destLength, src, srcLength));
    }
    /* bool TryUnwrap(src, dest, out written)  (Decompressor.cs:96-115): false only for dstSize_tooSmall. */
    bool TryUnwrap(const void* src, size_t srcLength, void* dest, size_t destLength, size_t& written) {
        EnsureNotDisposed();
        size_t returnValue = ZSTD_decompressDCtx(dctx_, dest, destLength, src, srcLength);
        if (ThrowHelper::IsDstSizeTooSmall(returnValue)) { written = 0; return false; }
        written = ThrowHelper::EnsureZstdSuccess(returnValue);
        return true;
    }

    /* Many independent compressed buffers in one pass (ZSTDB200_decompressBatch); per-item sizes / error codes. */
    std::vector<BatchResult> UnwrapBatch(const std::vector<const void*>& src, const std::vector<size_t>& srcLength,
                                         const std::vector<void*>& dest, const std::vector<size_t>& destLength) {
        EnsureNotDisposed();
        size_t n = src.size();
        if (srcLength.size() != n || dest.size() != n || destLength.size() != n)
            throw std::invalid_argument("UnwrapBatch: array lengths differ");
        std::vector<size_t> raw(n);
        ThrowHelper::EnsureZstdSuccess(ZSTDB200_decompressBatch(dctx_, n, src.data(), srcLength.data(),
                                                                dest.data(), destLength.data(), raw.data()));
        return detail::ToResults(raw);
    }

    void Dispose() {                                                                   /* Decompressor.cs:117-130 */
        if (dctx_) { ZSTD_freeDCtx(dctx_); dctx_ = nullptr; }
    }
    ZSTD_DCtx* Handle() const { return dctx_; }

private:
    void EnsureNotDisposed() const { if (!dctx_) throw ObjectDisposedException("Decompressor"); }
    ZSTD_DCtx* dctx_;
};

/* The batch scheduler across the GPUs of one box (ZSTDB200_*BatchMulti, include/zstd_b200.h): WrapBatch / UnwrapBatch with the
 * meaning they have on Compressor / Decompressor; every item is handled by exactly one device, results in the caller's order.
 * Closest notion in the reference: many contexts used at once (ZstdNetTests.cs:498-522). */
class MultiCodec {
public:
    explicit MultiCodec(int nDevices = 0, int level = 1) : m_(ZSTDB200_createMulti(nDevices)), level_(level) {
        if (!m_) throw ZstdException(ZSTD_ErrorCode::GENERIC, std::string("Failed to create the multi-device scheduler: ") + ZSTDB200_lastErrorString());
    }
    ~MultiCodec() { Dispose(); }
    MultiCodec(const MultiCodec&) = delete;
    MultiCodec& operator=(const MultiCodec&) = delete;
    int DeviceCount() const { return ZSTDB200_multiDeviceCount(m_); }
    int Level() const { return level_; }
    void Level(int value) { level_ = value; }
    void SetParameter(ZSTD_cParameter parameter, int value) { ThrowHelper::EnsureZstdSuccess(ZSTDB200_multiSetParameter(m_, static_cast<int>(parameter), value)); }
    void LoadDictionary(const void* dict, size_t dictLength) { ThrowHelper::EnsureZstdSuccess(ZSTDB200_multiLoadDictionary(m_, dict, dict ? dictLength : 0)); }
    std::vector<BatchResult> WrapBatch(const std::vector<const void*>& src, const std::vector<size_t>& srcLength,
                                       const std::vector<void*>& dest, const std::vector<size_t>& destLength) {
        size_t n = src.size();
        if (srcLength.size() != n || dest.size() != n || destLength.size() != n) throw std::invalid_argument("WrapBatch: array lengths differ");
        std::vector<size_t> raw(n);
        ThrowHelper::EnsureZstdSuccess(ZSTDB200_compressBatchMulti(m_, n, level_ == 0 ? 3 : level_, src.data(), srcLength.data(), dest.data(), destLength.data(), raw.data()));
        return detail::ToResults(raw);
    }
    std::vector<BatchResult> UnwrapBatch(const std::vector<const void*>& src, const std::vector<size_t>& srcLength,
                                         const std::vector<void*>& dest, const std::vector<size_t>& destLength) {
        size_t n = src.size();
        if (srcLength.size() != n || dest.size() != n || destLength.size() != n) throw std::invalid_argument("UnwrapBatch: array lengths differ");
        std::vector<size_t> raw(n);
        ThrowHelper::EnsureZstdSuccess(ZSTDB200_decompressBatchMulti(m_, n, src.data(), srcLength.data(), dest.data(), destLength.data(), raw.data()));
        return detail::ToResults(raw);
    }
    void Dispose() { if (m_) { ZSTDB200_freeMulti(m_); m_ = nullptr; } }

private:
    ZSTDB200_Multi* m_;
    int level_;
};

/* Stream adapters over the batch API.  The reference's CompressionStream / DecompressionStream
 * (src/ZstdSharp/CompressionStream.cs:8-190, DecompressionStream.cs) drive zstd's serial streaming state machine; these keep the
 * classes' shape (constructor over an inner stream, Write / Flush / Read / Dispose, SetParameter) but cut the data into independent
 * frames of frameSize bytes and hand whole batches to WrapBatch / UnwrapBatch.  The output is format compatible: concatenated
 * frames are one valid zstd stream (ZSTD_decompressMultiFrame, Unsafe/ZstdDecompress.cs:1216), and DecompressionStream accepts any
 * concatenation of frames (from this class, the reference, the zstd CLI), cutting it at ZSTD_findFrameCompressedSize (:958). */
class CompressionStream {
public:
    CompressionStream(std::ostream& stream, int level = Compressor::DefaultCompressionLevel, size_t frameSize = 128 * 1024, size_t batchFrames = 1024)
        : inner_(stream), comp_(level), frame_(frameSize), batch_(batchFrames) {
        if (frameSize == 0 || frameSize > 128 * 1024 || batchFrames == 0) throw std::invalid_argument("frameSize / batchFrames");
    }
    ~CompressionStream() { try { Dispose(); } catch (...) {} }
    void SetParameter(ZSTD_cParameter parameter, int value) { EnsureNotDisposed(); comp_.SetParameter(parameter, value); }   /* CompressionStream.cs:46-50 */
    void LoadDictionary(const void* dict, size_t dictLength) { EnsureNotDisposed(); comp_.LoadDictionary(dict, dictLength); }  /* CompressionStream.cs:58-62 */
    void Write(const void* buffer, size_t count) {                                     /* CompressionStream.cs:130-150 */
        EnsureNotDisposed();
        const uint8_t* p = static_cast<const uint8_t*>(buffer);
        buf_.insert(buf_.end(), p, p + count);
        while (buf_.size() >= frame_ * batch_) Emit(frame_ * batch_);
    }
    void Flush() { EnsureNotDisposed(); Emit(buf_.size()); inner_.flush(); }           /* CompressionStream.cs:88-98: everything written so far becomes decodable */
    void Dispose() { if (!done_) { Emit(buf_.size()); inner_.flush(); done_ = true; } }

private:
    void EnsureNotDisposed() const { if (done_) throw ObjectDisposedException("CompressionStream"); }
    void Emit(size_t bytes) {
        if (bytes == 0) return;
        size_t const n = (bytes + frame_ - 1) / frame_;
        std::vector<const void*> src(n); std::vector<size_t> len(n), cap(n); std::vector<void*> dst(n);
        size_t total = 0;
        for (size_t i = 0; i < n; i++) { len[i] = std::min(frame_, bytes - i * frame_); cap[i] = ZSTD_compressBound(len[i]); total += cap[i]; }
        std::vector<uint8_t> out(total);
        size_t o = 0;
        for (size_t i = 0; i < n; i++) { src[i] = buf_.data() + i * frame_; dst[i] = out.data() + o; o += cap[i]; }
        std::vector<BatchResult> r = comp_.WrapBatch(src, len, dst, cap);
        for (size_t i = 0; i < n; i++) {
            if (r[i].Code != ZSTD_ErrorCode::no_error) throw ZstdException(r[i].Code, "compression of a stream frame failed");
            inner_.write(static_cast<const char*>(dst[i]), static_cast<std::streamsize>(r[i].Size));
        }
        buf_.erase(buf_.begin(), buf_.begin() + static_cast<std::ptrdiff_t>(bytes));
    }
    std::ostream& inner_;
    Compressor comp_;
    size_t frame_, batch_;
    std::vector<uint8_t> buf_;
    bool done_ = false;
};

class DecompressionStream {
public:
    explicit DecompressionStream(std::istream& stream, size_t batchBytes = size_t(64) << 20) : inner_(stream), batchBytes_(batchBytes) {}
    void SetParameter(ZSTD_dParameter parameter, int value) { dec_.SetParameter(parameter, value); }            /* DecompressionStream.cs:47-51 */
    void LoadDictionary(const void* dict, size_t dictLength) { dec_.LoadDictionary(dict, dictLength); }
    /* int Read(byte[] buffer, int offset, int count)  (DecompressionStream.cs:74-120): 0 at the end of the stream; a stream that
     * ends inside a frame throws (EndOfStreamException there, ZstdException(srcSize_wrong) here). */
    size_t Read(void* buffer, size_t count) {
        uint8_t* out = static_cast<uint8_t*>(buffer);
        size_t got = 0;
        while (got < count) {
            if (pos_ == ready_.size()) { if (!Refill()) break; }
            size_t const m = std::min(count - got, ready_.size() - pos_);
            std::memcpy(out + got, ready_.data() + pos_, m);
            pos_ += m; got += m;
        }
        return got;
    }

private:
    bool Refill() {
        ready_.clear(); pos_ = 0;
        // pull compressed bytes until at least one whole frame is buffered (or the inner stream ends)
        std::vector<size_t> cut;          // end offsets of whole frames inside pending_
        size_t at = 0;
        for (;;) {
            while (at < pending_.size()) {
                size_t const fs = ZSTD_findFrameCompressedSize(pending_.data() + at, pending_.size() - at);
                if (ZSTD_isError(fs)) break;
                at += fs; cut.push_back(at);
            }
            if (!cut.empty() && (at >= batchBytes_ || eof_)) break;
            if (eof_) break;
            size_t const old = pending_.size();
            pending_.resize(old + (size_t(1) << 20));
            inner_.read(reinterpret_cast<char*>(pending_.data() + old), static_cast<std::streamsize>(size_t(1) << 20));
            size_t const n = static_cast<size_t>(inner_.gcount());
            pending_.resize(old + n);
            if (n == 0) eof_ = true;
        }
        if (cut.empty()) {
            if (!pending_.empty()) {      // leftover bytes that are not a whole frame: report what ZSTD_findFrameCompressedSize says
                size_t const fs = ZSTD_findFrameCompressedSize(pending_.data(), pending_.size());
                pending_.clear();
                ThrowHelper::EnsureZstdSuccess(fs);
            }
            return false;
        }
        size_t const n = cut.size();
        std::vector<const void*> src(n); std::vector<size_t> len(n), cap(n); std::vector<void*> dst(n);
        size_t total = 0, prev = 0;
        for (size_t i = 0; i < n; i++) {
            src[i] = pending_.data() + prev; len[i] = cut[i] - prev; prev = cut[i];
            cap[i] = static_cast<size_t>(ThrowHelper::EnsureContentSizeOk(ZSTD_decompressBound(src[i], len[i])));
            total += cap[i];
        }
        std::vector<uint8_t> out(total ? total : 1);
        size_t o = 0;
        for (size_t i = 0; i < n; i++) { dst[i] = out.data() + o; o += cap[i]; }
        std::vector<BatchResult> r = dec_.UnwrapBatch(src, len, dst, cap);
        for (size_t i = 0; i < n; i++) {
            if (r[i].Code != ZSTD_ErrorCode::no_error) throw ZstdException(r[i].Code, "a frame of the stream failed to decode");
            ready_.insert(ready_.end(), static_cast<uint8_t*>(dst[i]), static_cast<uint8_t*>(dst[i]) + r[i].Size);
        }
        pending_.erase(pending_.begin(), pending_.begin() + static_cast<std::ptrdiff_t>(cut.back()));
        return true;
    }
    std::istream& inner_;
    Decompressor dec_;
    size_t batchBytes_;
    std::vector<uint8_t> pending_, ready_;
    size_t pos_ = 0;
    bool eof_ = false;
};

}  // namespace ZstdSharp

#endif /* ZSTD_B200_HPP */
