// zb_hostcopy.cpp -- host-side staging copy of the batch scheduler (product code, plain C++: no CUDA in this file).
//
// The host-pointer API moves pageable caller buffers (what ZstdSharp's `fixed (byte* ...)` over managed arrays hands over,
// Compressor.cs:91-95 / Decompressor.cs:62-88) through a pinned staging ring with host threads.  Those copies are bound by the
// host's memory bandwidth: a plain store first reads the destination line (write allocate), a non-temporal store does not, which
// removes a third of the traffic of a large copy whose destination is not read again soon.
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <immintrin.h>

namespace zb {

__attribute__((target("avx2"))) static void copy_stream_avx2(uint8_t* d, const uint8_t* s, size_t n)
{
    size_t const head = (32 - ((uintptr_t)d & 31)) & 31;
    if (head) { size_t const h = head < n ? head : n; memcpy(d, s, h); d += h; s += h; n -= h; }
    size_t const blocks = n / 128;
    for (size_t i = 0; i < blocks; i++) {
        __m256i const a = _mm256_loadu_si256((const __m256i*)(s)), b = _mm256_loadu_si256((const __m256i*)(s + 32));
        __m256i const c = _mm256_loadu_si256((const __m256i*)(s + 64)), e = _mm256_loadu_si256((const __m256i*)(s + 96));
        _mm256_stream_si256((__m256i*)(d), a); _mm256_stream_si256((__m256i*)(d + 32), b);
        _mm256_stream_si256((__m256i*)(d + 64), c); _mm256_stream_si256((__m256i*)(d + 96), e);
        s += 128; d += 128;
    }
    _mm_sfence();
    size_t const tail = n - blocks * 128;
    if (tail) memcpy(d, s, tail);
}

// Copies n bytes; large copies go around the cache on the store side when the CPU has AVX2.
void host_copy(void* dst, const void* src, size_t n)
{
    static int const useStream = []() { __builtin_cpu_init(); return __builtin_cpu_supports("avx2") ? 1 : 0; }();
    if (useStream && n >= 4096) copy_stream_avx2((uint8_t*)dst, (const uint8_t*)src, n);
    else memcpy(dst, src, n);
}

}  // namespace zb
