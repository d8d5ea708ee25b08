// C++ host-mirror tests (include/zstd_b200.hpp): the behavioural set of the reference's ZstdNetTests.cs, written against the
// same class / method names, with the CPU oracle (oracle/zo.h, test infrastructure) as the byte-exact checker.
//
//   host_mirror_test host   host-only checks (no GPU needed): bounds, content-size errors, disposed objects, and that compute
//                           calls FAIL LOUDLY without a device (no CPU fallback)
//   host_mirror_test gpu    the full set on cuda:0
//
// Built and run by tests/test_cpp_host.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <sstream>

#include "zstd_b200.hpp"
#include "zo.h"

using namespace ZstdSharp;
typedef std::vector<uint8_t> Bytes;

static int g_checks = 0;
#define CHECK(cond)                                                                    \
    do {                                                                               \
        ++g_checks;                                                                    \
        if (!(cond)) {                                                                 \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
            std::exit(1);                                                              \
        }                                                                              \
    } while (0)

template <class F>
static ZSTD_ErrorCode thrown_code(F f) {
    try { f(); } catch (const ZstdException& e) { return e.Code; }
    return ZSTD_ErrorCode::no_error;
}

// Deterministic inputs: Zipf-ish word text (compressible), byte ramp, noise.
static uint32_t rng_state = 0xD1C3u;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5; return rng_state; }
static Bytes text_like(size_t n) {
    static const char* words[] = {"the", "of", "and", "frame", "block", "sequence", "literal", "match", "offset", "huffman",
                                  "entropy", "table", "window", "stream", "decode", "a", "in", "to", "B200", "zstd"};
    Bytes out; out.reserve(n + 16);
    while (out.size() < n) {
        uint32_t r = rnd();
        const char* w = words[(r % 20) * ((r >> 8) % 20) / 20];
        out.insert(out.end(), w, w + std::strlen(w));
        out.push_back((r >> 20) % 11 == 0 ? '\n' : ' ');
    }
    out.resize(n);
    return out;
}
static Bytes ramp(size_t n) { Bytes b(n); for (size_t i = 0; i < n; ++i) b[i] = uint8_t(i); return b; }
static Bytes noise(size_t n) { Bytes b(n); for (size_t i = 0; i < n; ++i) b[i] = uint8_t(rnd() >> 11); return b; }

static Bytes oracle_compress(const Bytes& src, int level, int checksum = 0) {
    Bytes dst(zo_compressBound(src.size()));
    void* c = zo_createCCtx();
    size_t r = zo_compressCCtx(c, dst.data(), dst.size(), src.data(), src.size(), level, checksum);
    zo_freeCCtx(c);
    CHECK(!zo_isError(r));
    dst.resize(r);
    return dst;
}

static void host_only() {
    // Compressor.GetCompressBound (Compressor.cs:72-76) against the oracle's restatement
    for (int n : {0, 1, 100, 4096, 131072, 131073, 1 << 20})
        CHECK(size_t(Compressor::GetCompressBound(n)) == zo_compressBound(size_t(n)));
    // GetDecompressedSize: a good frame, garbage (ZstdNetTests.cs:169-179), truncated header
    Bytes data = text_like(70000);
    Bytes frame = oracle_compress(data, 1);
    CHECK(Decompressor::GetDecompressedSize(frame.data(), frame.size()) == data.size());
    Bytes junk = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
    CHECK(thrown_code([&] { Decompressor::GetDecompressedSize(junk.data(), junk.size()); }) == ZSTD_ErrorCode::GENERIC);
    CHECK(thrown_code([&] { Decompressor::GetDecompressedSize(frame.data(), 3); }) == ZSTD_ErrorCode::GENERIC);
    // Unwrap with maxDecompressedSize below the content size: dstSize_tooSmall before any work (ZstdNetTests.cs:217-237)
    Decompressor d;
    CHECK(thrown_code([&] { d.Unwrap(frame.data(), frame.size(), 1000); }) == ZSTD_ErrorCode::dstSize_tooSmall);
    // parameters (Compressor.cs:16-33)
    Compressor c(1);
    CHECK(c.Level() == 1);
    c.Level(3);
    CHECK(c.Level() == 3);
    c.SetParameter(ZSTD_cParameter::ZSTD_c_checksumFlag, 1);
    CHECK(thrown_code([&] { c.SetParameter(ZSTD_cParameter::ZSTD_c_checksumFlag, 2); }) == ZSTD_ErrorCode::parameter_outOfBound);
    CHECK(thrown_code([&] { c.Level(7); }) == ZSTD_ErrorCode::parameter_unsupported);
    CHECK(c.Level() == 3);
    // Compressor.GetParameter (Compressor.cs:35-41), Decompressor.SetParameter / GetParameter (Decompressor.cs:22-34)
    CHECK(c.GetParameter(ZSTD_cParameter::ZSTD_c_compressionLevel) == 3 && c.GetParameter(ZSTD_cParameter::ZSTD_c_checksumFlag) == 1);
    c.Level(0);                                                       // 0 is stored as ZSTD_CLEVEL_DEFAULT
    CHECK(c.GetParameter(ZSTD_cParameter::ZSTD_c_compressionLevel) == 3);
    c.Level(3);
    CHECK(d.GetParameter(ZSTD_dParameter::ZSTD_d_windowLogMax) == 27);
    d.SetParameter(ZSTD_dParameter::ZSTD_d_windowLogMax, 31);
    CHECK(d.GetParameter(ZSTD_dParameter::ZSTD_d_windowLogMax) == 31);
    d.SetParameter(ZSTD_dParameter::ZSTD_d_windowLogMax, 0);
    CHECK(d.GetParameter(ZSTD_dParameter::ZSTD_d_windowLogMax) == 27);
    CHECK(thrown_code([&] { d.SetParameter(ZSTD_dParameter::ZSTD_d_windowLogMax, 9); }) == ZSTD_ErrorCode::parameter_outOfBound);
    CHECK(thrown_code([&] { d.SetParameter(ZSTD_dParameter::ZSTD_d_windowLogMax, 32); }) == ZSTD_ErrorCode::parameter_outOfBound);
    CHECK(thrown_code([&] { d.SetParameter(static_cast<ZSTD_dParameter>(1002), 1); }) == ZSTD_ErrorCode::parameter_unsupported);
    // disposed objects (Compressor.cs / Decompressor.cs EnsureNotDisposed)
    c.Dispose();
    bool disposed = false;
    try { c.Wrap(data); } catch (const ObjectDisposedException&) { disposed = true; }
    CHECK(disposed);
    // without a device every compute call answers GENERIC: there is no CPU path behind this API
    if (ZSTDB200_deviceCount() == 0) {
        Compressor c2(1);
        CHECK(thrown_code([&] { c2.Wrap(data); }) == ZSTD_ErrorCode::GENERIC);
        CHECK(thrown_code([&] { d.Unwrap(frame); }) == ZSTD_ErrorCode::GENERIC);
        CHECK(std::strstr(ZSTDB200_lastErrorString(), "CUDA") != nullptr);
    }
}

static void round_trip(Compressor& c, Decompressor& d, const Bytes& data, int level, int checksum = 0) {
    Bytes frame = c.Wrap(data);
    CHECK(frame == oracle_compress(data, level, checksum));       // byte-identical to the reference algorithm
    Bytes back = d.Unwrap(frame);
    CHECK(back == data);
}

static void gpu_all() {
    CHECK(ZSTDB200_deviceCount() > 0);
    Decompressor d;
    // CompressAndDecompress_workCorrectly (ZstdNetTests.cs:26-41), levels 1..3 and the default level
    for (int level : {0, 1, 2, 3}) {
        Compressor c(level);
        round_trip(c, d, text_like(131072), level == 0 ? 3 : level);
    }
    Compressor c(1);
    // ..._onEmptyBuffer / _onOneByteBuffer (:459-478)
    round_trip(c, d, Bytes(), 1);
    round_trip(c, d, Bytes(1, 42), 1);
    // ..._onArraysOfDifferentSizes (:481-495): byte ramp of every size class, plus multi-block inputs
    for (size_t n : {2u, 3u, 9u, 10u, 11u, 255u, 256u, 257u, 1000u, 4096u, 65535u, 65536u, 65537u, 131071u, 131072u, 131073u, 300000u})
        round_trip(c, d, ramp(n), 1);
    round_trip(c, d, noise(131072), 1);                            // raw block
    round_trip(c, d, Bytes(131072, 0), 1);                         // RLE-like input (22-byte frame)
    round_trip(c, d, text_like(1 << 20), 1);                       // 8 blocks in one frame
    // CompressAndDecompress_worksCorrectly_advanced (:44-73): checksum adds 4 bytes
    {
        Bytes data = text_like(50000);
        Bytes plain = c.Wrap(data);
        c.SetParameter(ZSTD_cParameter::ZSTD_c_checksumFlag, 1);
        Bytes summed = c.Wrap(data);
        CHECK(summed.size() == plain.size() + 4);
        CHECK(summed == oracle_compress(data, 1, 1));
        CHECK(d.Unwrap(summed) == data);
        summed[summed.size() - 1] ^= 1;
        CHECK(thrown_code([&] { d.Unwrap(summed); }) == ZSTD_ErrorCode::checksum_wrong);
        c.SetParameter(ZSTD_cParameter::ZSTD_c_checksumFlag, 0);
    }
    // Compress_throwsDstSizeTooSmall / Compress_tryWrap (:402-430), Decompress_throwsDstSizeTooSmall (:433-456)
    {
        Bytes data = text_like(20000), frame = c.Wrap(data), small(20), out(data.size() - 1);
        size_t written = 123;
        CHECK(thrown_code([&] { c.Wrap(data.data(), data.size(), small.data(), small.size()); }) == ZSTD_ErrorCode::dstSize_tooSmall);
        CHECK(!c.TryWrap(data.data(), data.size(), small.data(), small.size(), written) && written == 0);
        CHECK(thrown_code([&] { d.Unwrap(frame.data(), frame.size(), out.data(), out.size()); }) == ZSTD_ErrorCode::dstSize_tooSmall);
        written = 123;
        CHECK(!d.TryUnwrap(frame.data(), frame.size(), out.data(), out.size(), written) && written == 0);
        // Compress_canWrite_toGivenBuffer / Decompress_canWrite_toGivenBuffer (:347-399): offsets inside larger buffers
        Bytes big(1000 + ZSTD_compressBound(data.size()), 0xEE);
        CHECK(c.TryWrap(data.data(), data.size(), big.data() + 1000, big.size() - 1000, written) && written == frame.size());
        CHECK(std::memcmp(big.data() + 1000, frame.data(), frame.size()) == 0 && big[999] == 0xEE && big[1000 + written] == 0xEE);
        Bytes bigOut(100 + data.size() + 100, 0xEE);
        CHECK(d.TryUnwrap(big.data() + 1000, written, bigOut.data() + 100, data.size(), written) && written == data.size());
        CHECK(std::memcmp(bigOut.data() + 100, data.data(), data.size()) == 0 && bigOut[99] == 0xEE && bigOut[100 + data.size()] == 0xEE);
    }
    // Decompress_throwsZstdException_onInvalidData (:169-179), _onMalformedDecompressedSize (:182-214)
    {
        Bytes junk = noise(100);
        junk[0] = 0;
        CHECK(thrown_code([&] { d.Unwrap(junk); }) == ZSTD_ErrorCode::GENERIC);
        Bytes data = text_like(4000), frame = c.Wrap(data);
        // frame header: magic(4) descriptor(1) [no window byte: single segment] FCS.  Claim one byte less / more.
        CHECK((frame[4] >> 6) == 1);                                  // 2-byte FCS field (value - 256)
        Bytes lessClaim = frame, moreClaim = frame;
        uint16_t fcs = uint16_t(frame[5] | (frame[6] << 8));
        lessClaim[5] = uint8_t((fcs - 1) & 0xFF); lessClaim[6] = uint8_t((fcs - 1) >> 8);
        moreClaim[5] = uint8_t((fcs + 1) & 0xFF); moreClaim[6] = uint8_t((fcs + 1) >> 8);
        Bytes dst(data.size() + 16);
        size_t r1 = zo_decompress(dst.data(), data.size() - 1, lessClaim.data(), lessClaim.size());
        size_t r2 = zo_decompress(dst.data(), data.size() + 1, moreClaim.data(), moreClaim.size());
        CHECK(zo_isError(r1) && zo_isError(r2));
        CHECK(int(thrown_code([&] { d.Unwrap(lessClaim); })) == zo_getErrorCode(r1));
        CHECK(int(thrown_code([&] { d.Unwrap(moreClaim); })) == zo_getErrorCode(r2));
    }
    // batch entry points: 300 items of mixed kinds and sizes, one bad item in the middle that must not poison the rest
    {
        const size_t n = 300;
        std::vector<Bytes> in(n), comp(n), out(n);
        std::vector<const void*> sp(n); std::vector<void*> dp(n); std::vector<size_t> ss(n), ds(n);
        for (size_t i = 0; i < n; ++i) {
            size_t len = (i % 7 == 0) ? 131072 : (rnd() % 140000);
            in[i] = (i % 5 == 0) ? noise(len) : (i % 5 == 1 ? ramp(len) : text_like(len));
            comp[i].resize(i == 150 ? 10 : ZSTD_compressBound(len));
            sp[i] = in[i].data(); ss[i] = len; dp[i] = comp[i].data(); ds[i] = comp[i].size();
        }
        std::vector<BatchResult> r = c.WrapBatch(sp, ss, dp, ds);
        for (size_t i = 0; i < n; ++i) {
            if (i == 150) { CHECK(r[i].Code == ZSTD_ErrorCode::dstSize_tooSmall); continue; }
            CHECK(r[i].Code == ZSTD_ErrorCode::no_error);
            comp[i].resize(r[i].Size);
            CHECK(comp[i] == oracle_compress(in[i], 1));
        }
        comp[150] = c.Wrap(in[150]);
        comp[151][comp[151].size() / 2] ^= 0x40;                       // corrupt one payload
        Bytes expect151(in[151].size());
        size_t o151 = zo_decompress(expect151.data(), expect151.size(), comp[151].data(), comp[151].size());
        std::vector<const void*> csp(n); std::vector<void*> odp(n); std::vector<size_t> css(n), ods(n);
        for (size_t i = 0; i < n; ++i) {
            out[i].assign(in[i].size() + 8, 0xEE);
            csp[i] = comp[i].data(); css[i] = comp[i].size(); odp[i] = out[i].data(); ods[i] = in[i].size();
        }
        std::vector<BatchResult> u = d.UnwrapBatch(csp, css, odp, ods);
        for (size_t i = 0; i < n; ++i) {
            if (i == 151) {
                if (zo_isError(o151)) CHECK(int(u[i].Code) == zo_getErrorCode(o151));
                else CHECK(u[i].Code == ZSTD_ErrorCode::no_error && u[i].Size == o151 &&
                           std::memcmp(out[i].data(), expect151.data(), o151) == 0);
                continue;
            }
            CHECK(u[i].Code == ZSTD_ErrorCode::no_error && u[i].Size == in[i].size());
            CHECK(std::memcmp(out[i].data(), in[i].data(), in[i].size()) == 0 && out[i][in[i].size()] == 0xEE);
        }
    }
    // ..._ifDifferentInstancesRunInDifferentThreads (:498-522): one context pair per thread
    {
        std::vector<Bytes> inputs;
        for (int t = 0; t < 4; ++t) inputs.push_back(text_like(100000 + 777 * t));
        std::vector<int> ok(4, 0);
        std::vector<std::thread> th;
        for (int t = 0; t < 4; ++t)
            th.emplace_back([&, t] {
                Compressor tc(1 + t % 3);
                Decompressor td;
                int good = 1;
                for (int k = 0; k < 6; ++k) {
                    Bytes f = tc.Wrap(inputs[t]);
                    good &= (f == oracle_compress(inputs[t], 1 + t % 3)) && (td.Unwrap(f) == inputs[t]);
                }
                ok[t] = good;
            });
        for (auto& x : th) x.join();
        for (int t = 0; t < 4; ++t) CHECK(ok[t]);
    }
    // Stream adapters (CompressionStream.cs / DecompressionStream.cs shape; ZstdNetSteamingTests.cs round trips): every frame the
    // adapter writes is the oracle's frame of that 128 KiB piece, the stream decodes back, foreign concatenations are accepted
    {
        Bytes data = text_like(5 * 131072 + 4321);
        std::ostringstream os;
        {
            CompressionStream cs(os, 1);
            cs.Write(data.data(), 100000);
            cs.Flush();                                                  // a flush closes the frame(s) written so far
            cs.Write(data.data() + 100000, data.size() - 100000);
            cs.Dispose();
            bool disposedThrows = false;
            try { cs.Write(data.data(), 1); } catch (const ObjectDisposedException&) { disposedThrows = true; }
            CHECK(disposedThrows);
        }
        std::string const z = os.str();
        Bytes expect;
        {
            Bytes a(data.begin(), data.begin() + 100000); Bytes f = oracle_compress(a, 1); expect.insert(expect.end(), f.begin(), f.end());
            for (size_t o = 100000; o < data.size(); o += 131072) {
                Bytes b(data.begin() + o, data.begin() + std::min(data.size(), o + 131072)); Bytes g = oracle_compress(b, 1);
                expect.insert(expect.end(), g.begin(), g.end());
            }
        }
        CHECK(Bytes(z.begin(), z.end()) == expect);
        std::istringstream is(z);
        DecompressionStream ds(is);
        Bytes back(data.size() + 100);
        size_t got = 0, r;
        while ((r = ds.Read(back.data() + got, std::min<size_t>(70001, back.size() - got))) != 0) got += r;
        CHECK(got == data.size() && std::memcmp(back.data(), data.data(), got) == 0);
        // a stream that ends inside a frame is an error (DecompressionStream.cs:108-113)
        std::istringstream cutIs(z.substr(0, z.size() - 7));
        DecompressionStream cutDs(cutIs);
        CHECK(thrown_code([&] { Bytes tmp(data.size() + 100); size_t g = 0, q; while ((q = cutDs.Read(tmp.data() + g, tmp.size() - g)) != 0) g += q; }) == ZSTD_ErrorCode::srcSize_wrong);
    }
    // Compressor.LoadDictionary / Decompressor.LoadDictionary with a raw-content dictionary (ZstdNetTests.cs:19-39, 95-134): the oracle's
    // bytes (Compressor.LoadDictionary + Wrap), the round trip, and the two failures the reference tests
    {
        Bytes const dict = text_like(20000), data = text_like(90000);
        Bytes expect(zo_compressBound(data.size()));
        size_t const er = zo_compress_usingLoadedDict(expect.data(), expect.size(), data.data(), data.size(), dict.data(), dict.size(), 3, 0);
        CHECK(!zo_isError(er)); expect.resize(er);
        Compressor c(3);
        c.LoadDictionary(dict.data(), dict.size());
        Bytes const f = c.Wrap(data);
        CHECK(f == expect);
        c.LoadDictionary(nullptr, 0);
        CHECK(c.Wrap(data) == oracle_compress(data, 3));
        Decompressor d;
        d.LoadDictionary(dict.data(), dict.size());
        CHECK(d.Unwrap(f) == data);
        d.LoadDictionary(nullptr, 0);
        CHECK(thrown_code([&] { d.Unwrap(f); }) != ZSTD_ErrorCode::no_error);             // DecompressWithoutDictionary_throwsZstdException_onDataCompressedWithIt
        std::string const other = "zstd supports raw-content dictionaries";
        d.LoadDictionary(other.data(), other.size());
        CHECK(thrown_code([&] { d.Unwrap(f); }) != ZSTD_ErrorCode::no_error);             // DecompressWithAnotherDictionary_throwsZstdException
    }
    // MultiCodec (ZSTDB200_*BatchMulti): however many devices are visible, the oracle's frames in the caller's order
    {
        MultiCodec m(0, 1);
        CHECK(m.DeviceCount() >= 1);
        size_t const n = 40;
        std::vector<Bytes> in(n), comp(n), out(n);
        std::vector<const void*> sp(n); std::vector<void*> dp(n); std::vector<size_t> ss(n), dsz(n);
        for (size_t i = 0; i < n; ++i) {
            in[i] = text_like(1000 + 3271 * i); comp[i].resize(ZSTD_compressBound(in[i].size()));
            sp[i] = in[i].data(); ss[i] = in[i].size(); dp[i] = comp[i].data(); dsz[i] = comp[i].size();
        }
        std::vector<BatchResult> r = m.WrapBatch(sp, ss, dp, dsz);
        for (size_t i = 0; i < n; ++i) { CHECK(r[i].Code == ZSTD_ErrorCode::no_error); comp[i].resize(r[i].Size); CHECK(comp[i] == oracle_compress(in[i], 1)); }
        std::vector<const void*> csp(n); std::vector<void*> odp(n); std::vector<size_t> css(n), ods(n);
        for (size_t i = 0; i < n; ++i) { out[i].resize(in[i].size()); csp[i] = comp[i].data(); css[i] = comp[i].size(); odp[i] = out[i].data(); ods[i] = out[i].size(); }
        std::vector<BatchResult> u = m.UnwrapBatch(csp, css, odp, ods);
        for (size_t i = 0; i < n; ++i) CHECK(u[i].Code == ZSTD_ErrorCode::no_error && u[i].Size == in[i].size() && out[i] == in[i]);
    }
}

int main(int argc, char** argv) {
    std::string mode = argc > 1 ? argv[1] : "host";
    host_only();
    if (mode == "gpu") gpu_all();
    std::printf("host_mirror_test %s: %d checks passed\n", mode.c_str(), g_checks);
    return 0;
}
