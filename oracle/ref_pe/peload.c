/* oracle/ref_pe/peload.c — TEST INFRASTRUCTURE, never linked into or loaded by the product library.
 *
 * Runs the reference's own native oracle on Linux: /root/reference/src/Zstd.Extern/libzstd.dll is the zstd 1.5.1
 * build that the reference's differential test asserts ZstdSharp equals byte for byte at every level
 * (src/ZstdSharp.Test/ZstdTest.cs:18-90, bound through src/Zstd.Extern/ExternMethods.cs:8-37).  It is a PE32+
 * x86-64 image with a static CRT whose only import library is KERNEL32, so it can be mapped by hand:
 *
 *   - the image bytes are embedded at build time (.incbin of the file where it lies; the output goes to oracle/_ref/,
 *     which is git-ignored: no reference bytes enter the history);
 *   - sections are copied to an anonymous RWX mapping, IMAGE_REL_BASED_DIR64 relocations applied;
 *   - DllMain / the CRT start-up is NOT run.  The CRT's malloc family reaches HeapAlloc(__acrt_heap, ..) with a null
 *     heap handle, which the stubs below ignore; memcpy/memset fall back to their SSE2 paths (__isa_available == 0);
 *   - every other import gets a trap thunk that names the import and aborts, so an unexpected dependency is loud;
 *   - Windows x64 code reads the TEB through gs: (__chkstk reads gs:[0x10], the stack limit): every calling thread
 *     gets a zeroed fake TEB via arch_prctl(ARCH_SET_GS) (user-space Linux does not use gs on x86-64);
 *   - exports are called through __attribute__((ms_abi)) function pointers.
 *
 * The ZREF_* wrappers are System V functions for ctypes / C callers; names follow the exports they forward to.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <asm/prctl.h>

#ifndef ZREF_DLL_PATH
#error "ZREF_DLL_PATH must name the reference's libzstd.dll"
#endif

__asm__(".section .rodata\n"
        ".balign 16\n"
        ".global zref_dll_image\n"
        "zref_dll_image:\n"
        ".incbin \"" ZREF_DLL_PATH "\"\n"
        ".global zref_dll_image_end\n"
        "zref_dll_image_end:\n"
        ".previous\n");
extern const unsigned char zref_dll_image[], zref_dll_image_end[];

#define MS __attribute__((ms_abi))

static uint8_t* g_base;          /* mapped image */
static uint32_t g_sizeOfImage;
static const char** g_importNames;
static int g_nImports;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static int g_loadError;
static __thread int t_gsReady;
static __thread uint64_t t_teb[512] __attribute__((aligned(64)));
static __thread uint32_t t_lastError;

static uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* ---- KERNEL32 stand-ins (ms_abi) -------------------------------------------------------------------------------- */
#define HEAP_ZERO_MEMORY 0x8u
static MS void* k32_HeapAlloc(void* heap, uint32_t flags, size_t n) {
    (void)heap;
    size_t* p = (size_t*)((flags & HEAP_ZERO_MEMORY) ? calloc(1, n + 16) : malloc(n + 16));
    if (!p) return NULL;
    p[0] = n;
    return p + 2;
}
static MS int k32_HeapFree(void* heap, uint32_t flags, void* q) {
    (void)heap; (void)flags;
    if (q) free((size_t*)q - 2);
    return 1;
}
static MS void* k32_HeapReAlloc(void* heap, uint32_t flags, void* q, size_t n) {
    if (!q) return k32_HeapAlloc(heap, flags, n);
    size_t* old = (size_t*)q - 2;
    size_t oldN = old[0];
    size_t* p = (size_t*)realloc(old, n + 16);
    if (!p) return NULL;
    p[0] = n;
    if ((flags & HEAP_ZERO_MEMORY) && n > oldN) memset((uint8_t*)(p + 2) + oldN, 0, n - oldN);
    return p + 2;
}
static MS size_t k32_HeapSize(void* heap, uint32_t flags, const void* q) {
    (void)heap; (void)flags;
    return ((const size_t*)q - 2)[0];
}
static MS void* k32_GetProcessHeap(void) { return (void*)0x1000; }
static MS uint32_t k32_GetLastError(void) { return t_lastError; }
static MS void k32_SetLastError(uint32_t e) { t_lastError = e; }
static MS int k32_QueryPerformanceCounter(int64_t* v) { *v = 0; return 1; }
static MS int k32_QueryPerformanceFrequency(int64_t* v) { *v = 10000000; return 1; }
static MS void* k32_EncodePointer(void* p) { return p; }
static MS int k32_IsProcessorFeaturePresent(uint32_t f) { (void)f; return 0; }
static MS int k32_IsDebuggerPresent(void) { return 0; }

static MS void k32_trap(int idx) {
    fprintf(stderr, "oracle/_ref: libzstd.dll called unsupported KERNEL32 import #%d (%s)\n", idx,
            (idx >= 0 && idx < g_nImports) ? g_importNames[idx] : "?");
    abort();
}

static const struct { const char* name; void* fn; } kStubs[] = {
    {"HeapAlloc", (void*)k32_HeapAlloc},
    {"HeapFree", (void*)k32_HeapFree},
    {"HeapReAlloc", (void*)k32_HeapReAlloc},
    {"HeapSize", (void*)k32_HeapSize},
    {"GetProcessHeap", (void*)k32_GetProcessHeap},
    {"GetLastError", (void*)k32_GetLastError},
    {"SetLastError", (void*)k32_SetLastError},
    {"QueryPerformanceCounter", (void*)k32_QueryPerformanceCounter},
    {"QueryPerformanceFrequency", (void*)k32_QueryPerformanceFrequency},
    {"EncodePointer", (void*)k32_EncodePointer},
    {"IsProcessorFeaturePresent", (void*)k32_IsProcessorFeaturePresent},
    {"IsDebuggerPresent", (void*)k32_IsDebuggerPresent},
};

/* ---- mapper ---------------------------------------------------------------------------------------------------- */
static void ensure_gs(void) {
    if (t_gsReady) return;
    t_teb[0x30 / 8] = (uint64_t)(uintptr_t)t_teb;   /* NT_TIB.Self */
    t_teb[0x10 / 8] = 0;                            /* StackLimit = 0: __chkstk never probes */
    if (syscall(SYS_arch_prctl, ARCH_SET_GS, (unsigned long)(uintptr_t)t_teb) != 0) {
        perror("oracle/_ref: arch_prctl(ARCH_SET_GS)");
        abort();
    }
    t_gsReady = 1;
}

static void load_image(void) {
    const uint8_t* img = zref_dll_image;
    size_t imgSize = (size_t)(zref_dll_image_end - zref_dll_image);
    g_loadError = 1;
    if (imgSize < 0x200 || img[0] != 'M' || img[1] != 'Z') return;
    uint32_t pe = rd32(img + 0x3c);
    if (pe + 24 + 240 > imgSize || rd32(img + pe) != 0x00004550u) return;
    if (rd16(img + pe + 4) != 0x8664) return;                      /* IMAGE_FILE_MACHINE_AMD64 */
    uint16_t nSec = rd16(img + pe + 6);
    uint16_t optSize = rd16(img + pe + 20);
    const uint8_t* opt = img + pe + 24;
    if (rd16(opt) != 0x20b) return;                                /* PE32+ */
    uint64_t imageBase = rd64(opt + 24);
    g_sizeOfImage = rd32(opt + 56);
    uint32_t sizeOfHeaders = rd32(opt + 60);
    uint32_t nDir = rd32(opt + 108);
    if (nDir < 6) return;
    uint32_t impRva = rd32(opt + 112 + 8 * 1);
    uint32_t relRva = rd32(opt + 112 + 8 * 5), relSize = rd32(opt + 112 + 8 * 5 + 4);

    uint8_t* base = (uint8_t*)mmap(NULL, g_sizeOfImage, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) { perror("oracle/_ref: mmap"); return; }
    memcpy(base, img, sizeOfHeaders);
    const uint8_t* sec = opt + optSize;
    for (int i = 0; i < nSec; i++, sec += 40) {
        uint32_t vsize = rd32(sec + 8), va = rd32(sec + 12), rawSize = rd32(sec + 16), rawPtr = rd32(sec + 20);
        uint32_t n = rawSize < vsize ? rawSize : vsize;
        if ((uint64_t)va + vsize > g_sizeOfImage || (uint64_t)rawPtr + n > imgSize) return;
        memcpy(base + va, img + rawPtr, n);
    }
    /* base relocations */
    int64_t delta = (int64_t)((uint64_t)(uintptr_t)base - imageBase);
    for (uint32_t off = 0; off + 8 <= relSize;) {
        uint32_t page = rd32(base + relRva + off), blk = rd32(base + relRva + off + 4);
        if (blk < 8) break;
        for (uint32_t k = 8; k + 2 <= blk; k += 2) {
            uint16_t e = rd16(base + relRva + off + k);
            int type = e >> 12;
            if (type == 10) {                                      /* IMAGE_REL_BASED_DIR64 */
                uint8_t* at = base + page + (e & 0xfff);
                uint64_t v = rd64(at) + (uint64_t)delta;
                memcpy(at, &v, 8);
            } else if (type != 0) {
                fprintf(stderr, "oracle/_ref: unsupported relocation type %d\n", type);
                return;
            }
        }
        off += blk;
    }
    /* imports: count, then one trap thunk per slot unless a stand-in exists */
    int total = 0;
    for (const uint8_t* d = base + impRva; rd32(d + 12); d += 20) {
        const uint8_t* ilt = base + (rd32(d) ? rd32(d) : rd32(d + 16));
        while (rd64(ilt)) { total++; ilt += 8; }
    }
    g_importNames = (const char**)calloc((size_t)total + 1, sizeof(char*));
    uint8_t* thunks = (uint8_t*)mmap(NULL, (size_t)total * 32 + 4096, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (thunks == MAP_FAILED || !g_importNames) return;
    int idx = 0;
    for (const uint8_t* d = base + impRva; rd32(d + 12); d += 20) {
        const uint8_t* ilt = base + (rd32(d) ? rd32(d) : rd32(d + 16));
        uint8_t* iat = base + rd32(d + 16);
        for (; rd64(ilt); ilt += 8, iat += 8, idx++) {
            uint64_t v = rd64(ilt);
            const char* name = (v >> 63) ? "(ordinal)" : (const char*)(base + (uint32_t)v + 2);
            g_importNames[idx] = name;
            void* target = NULL;
            for (size_t s = 0; s < sizeof(kStubs) / sizeof(kStubs[0]); s++)
                if (!strcmp(kStubs[s].name, name)) target = kStubs[s].fn;
            if (!target) {
                uint8_t* t = thunks + (size_t)idx * 32;
                uint64_t trap = (uint64_t)(uintptr_t)k32_trap;
                t[0] = 0xb9; memcpy(t + 1, &idx, 4);               /* mov ecx, idx     */
                t[5] = 0x48; t[6] = 0xb8; memcpy(t + 7, &trap, 8); /* mov rax, k32_trap */
                t[15] = 0xff; t[16] = 0xe0;                        /* jmp rax          */
                target = t;
            }
            uint64_t tv = (uint64_t)(uintptr_t)target;
            memcpy(iat, &tv, 8);
        }
    }
    g_nImports = idx;
    g_base = base;
    g_loadError = 0;
}

/* Looks an export up by name; NULL when absent. */
void* ZREF_sym(const char* want) {
    pthread_once(&g_once, load_image);
    if (g_loadError) return NULL;
    const uint8_t* base = g_base;
    uint32_t pe = rd32(base + 0x3c);
    const uint8_t* opt = base + pe + 24;
    uint32_t expRva = rd32(opt + 112);
    const uint8_t* e = base + expRva;
    uint32_t nNames = rd32(e + 24), funcs = rd32(e + 28), names = rd32(e + 32), ords = rd32(e + 36);
    for (uint32_t i = 0; i < nNames; i++) {
        const char* nm = (const char*)(base + rd32(base + names + 4 * i));
        if (!strcmp(nm, want)) {
            uint16_t o = rd16(base + ords + 2 * i);
            return (void*)(base + rd32(base + funcs + 4 * o));
        }
    }
    return NULL;
}

int ZREF_available(void) {
    pthread_once(&g_once, load_image);
    return !g_loadError;
}

static void* need(const char* name) {
    void* p = ZREF_sym(name);
    if (!p) { fprintf(stderr, "oracle/_ref: export %s not found (image load %s)\n", name, g_loadError ? "failed" : "ok"); abort(); }
    ensure_gs();
    return p;
}

/* ---- System V wrappers over the exports ---------------------------------------------------------------------------- */
#define FWD0(ret, name) \
    ret ZREF_##name(void) { static ret (MS *f)(void); if (!f) f = need("ZSTD_" #name); ensure_gs(); return f(); }
#define FWD1(ret, name, A) \
    ret ZREF_##name(A a) { static ret (MS *f)(A); if (!f) f = need("ZSTD_" #name); ensure_gs(); return f(a); }
#define FWD2(ret, name, A, B) \
    ret ZREF_##name(A a, B b) { static ret (MS *f)(A, B); if (!f) f = need("ZSTD_" #name); ensure_gs(); return f(a, b); }
#define FWD3(ret, name, A, B, C) \
    ret ZREF_##name(A a, B b, C c) { static ret (MS *f)(A, B, C); if (!f) f = need("ZSTD_" #name); ensure_gs(); return f(a, b, c); }
#define FWD4(ret, name, A, B, C, D) \
    ret ZREF_##name(A a, B b, C c, D d) { static ret (MS *f)(A, B, C, D); if (!f) f = need("ZSTD_" #name); ensure_gs(); return f(a, b, c, d); }
#define FWD5(ret, name, A, B, C, D, E) \
    ret ZREF_##name(A a, B b, C c, D d, E e) { static ret (MS *f)(A, B, C, D, E); if (!f) f = need("ZSTD_" #name); ensure_gs(); return f(a, b, c, d, e); }
#define FWD6(ret, name, A, B, C, D, E, F) \
    ret ZREF_##name(A a, B b, C c, D d, E e, F g) { static ret (MS *f)(A, B, C, D, E, F); if (!f) f = need("ZSTD_" #name); ensure_gs(); return f(a, b, c, d, e, g); }
#define FWD7(ret, name, A, B, C, D, E, F, G) \
    ret ZREF_##name(A a, B b, C c, D d, E e, F g, G h) { static ret (MS *f)(A, B, C, D, E, F, G); if (!f) f = need("ZSTD_" #name); ensure_gs(); return f(a, b, c, d, e, g, h); }

typedef void* P;
typedef const void* CP;

FWD0(unsigned, versionNumber)
FWD0(int, maxCLevel)
FWD0(int, minCLevel)
FWD1(size_t, compressBound, size_t)
FWD1(unsigned, isError, size_t)
FWD1(const char*, getErrorName, size_t)
FWD1(int, getErrorCode, size_t)
FWD0(P, createCCtx)
FWD1(size_t, freeCCtx, P)
FWD0(P, createDCtx)
FWD1(size_t, freeDCtx, P)
FWD6(size_t, compressCCtx, P, P, size_t, CP, size_t, int)
FWD5(size_t, compress2, P, P, size_t, CP, size_t)
FWD3(size_t, CCtx_setParameter, P, int, int)
FWD2(size_t, CCtx_reset, P, int)
FWD3(size_t, CCtx_loadDictionary, P, CP, size_t)
FWD2(size_t, CCtx_setPledgedSrcSize, P, unsigned long long)
FWD5(size_t, decompressDCtx, P, P, size_t, CP, size_t)
FWD7(size_t, decompress_usingDict, P, P, size_t, CP, size_t, CP, size_t)
FWD3(size_t, DCtx_loadDictionary, P, CP, size_t)
FWD2(size_t, DCtx_reset, P, int)
FWD3(size_t, DCtx_setParameter, P, int, int)
FWD2(unsigned long long, getFrameContentSize, CP, size_t)
FWD2(size_t, findFrameCompressedSize, CP, size_t)
FWD2(unsigned long long, decompressBound, CP, size_t)
FWD2(unsigned, getDictID_fromDict, CP, size_t)
FWD2(unsigned, getDictID_fromFrame, CP, size_t)

/* ZSTD_compress_usingDict has 8 arguments (cctx, dst, dstCap, src, srcSize, dict, dictSize, level). */
size_t ZREF_compress_usingDict8(P cctx, P dst, size_t cap, CP src, size_t n, CP dict, size_t dictSize, int level) {
    static size_t (MS *f)(P, P, size_t, CP, size_t, CP, size_t, int);
    if (!f) f = need("ZSTD_compress_usingDict");
    ensure_gs();
    return f(cctx, dst, cap, src, n, dict, dictSize, level);
}

/* ZSTD_getCParams returns a 28-byte struct {windowLog, chainLog, hashLog, searchLog, minMatch, targetLength, strategy}. */
typedef struct { unsigned v[7]; } zref_cparams;
void ZREF_getCParams(int level, unsigned long long srcSizeHint, size_t dictSize, unsigned out[7]) {
    static zref_cparams (MS *f)(int, unsigned long long, size_t);
    if (!f) f = need("ZSTD_getCParams");
    ensure_gs();
    zref_cparams c = f(level, srcSizeHint, dictSize);
    memcpy(out, c.v, sizeof c.v);
}

/* ZSTD_generateSequences(cctx, ZSTD_Sequence* out, outCap, src, srcSize): {offset, litLength, matchLength, rep} u32 x 4. */
size_t ZREF_generateSequences(P cctx, unsigned* outSeqs, size_t outCap, CP src, size_t n) {
    static size_t (MS *f)(P, unsigned*, size_t, CP, size_t);
    if (!f) f = need("ZSTD_generateSequences");
    ensure_gs();
    return f(cctx, outSeqs, outCap, src, n);
}

size_t ZREF_trainFromBuffer(P dict, size_t dictCap, CP samples, const size_t* sizes, unsigned nb) {
    static size_t (MS *f)(P, size_t, CP, const size_t*, unsigned);
    if (!f) f = need("ZDICT_trainFromBuffer");
    ensure_gs();
    return f(dict, dictCap, samples, sizes, nb);
}

/* One-shot helpers: level + optional checksum flag through the advanced API, as Compressor.Wrap drives it
 * (Compressor.cs:86-97: ZSTD_CCtx_setParameter(compressionLevel) once, then ZSTD_compress2). */
size_t ZREF_compress_level(P dst, size_t cap, CP src, size_t n, int level, int checksum) {
    P c = ZREF_createCCtx();
    if (!c) return (size_t)-64;
    ZREF_CCtx_setParameter(c, 100 /* ZSTD_c_compressionLevel */, level);
    if (checksum) ZREF_CCtx_setParameter(c, 201 /* ZSTD_c_checksumFlag */, 1);
    size_t r = ZREF_compress2(c, dst, cap, src, n);
    ZREF_freeCCtx(c);
    return r;
}
size_t ZREF_decompress(P dst, size_t cap, CP src, size_t n) {
    P d = ZREF_createDCtx();
    if (!d) return (size_t)-64;
    size_t r = ZREF_decompressDCtx(d, dst, cap, src, n);
    ZREF_freeDCtx(d);
    return r;
}

/* ---- multithreaded batch drivers (bench.py --impl reference / cpu_baseline with kind "reference") ----------------- */
typedef struct {
    int mode;                /* 0 decompress, 1 compress */
    int level, tid, nthreads;
    size_t n;
    const void* const* src; const size_t* srcSize;
    void* const* dst; const size_t* dstCap;
    size_t* result;
} zref_job;

static void* zref_worker(void* arg) {
    zref_job* j = (zref_job*)arg;
    ensure_gs();
    P ctx = j->mode ? ZREF_createCCtx() : ZREF_createDCtx();
    for (size_t i = (size_t)j->tid; i < j->n; i += (size_t)j->nthreads) {
        if (j->mode) j->result[i] = ZREF_compressCCtx(ctx, j->dst[i], j->dstCap[i], j->src[i], j->srcSize[i], j->level);
        else j->result[i] = ZREF_decompressDCtx(ctx, j->dst[i], j->dstCap[i], j->src[i], j->srcSize[i]);
    }
    if (j->mode) ZREF_freeCCtx(ctx); else ZREF_freeDCtx(ctx);
    return NULL;
}

static void zref_run(int mode, int level, size_t n, const void* const* src, const size_t* srcSize, void* const* dst,
                     const size_t* dstCap, size_t* result, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    zref_job jobs[256];
    pthread_once(&g_once, load_image);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (zref_job){mode, level, t, nthreads, n, src, srcSize, dst, dstCap, result};
        pthread_create(&th[t], NULL, zref_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

void ZREF_decompressBatchMT(size_t n, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap,
                            size_t* result, int nthreads) {
    zref_run(0, 0, n, src, srcSize, dst, dstCap, result, nthreads);
}
void ZREF_compressBatchMT(size_t n, const void* const* src, const size_t* srcSize, void* const* dst, const size_t* dstCap,
                          size_t* result, int level, int nthreads) {
    zref_run(1, level, n, src, srcSize, dst, dstCap, result, nthreads);
}
