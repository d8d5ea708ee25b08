"""Test-side access to the CPU oracle (oracle/) and to system libzstd 1.5.5 (second, independent checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use these.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "libzo.so")


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("zo_decode.c", "zo_encode.c", "zo_common.h", "zo.h")]
    if force or not os.path.exists(ORACLE_LIB) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_LIB) for s in srcs):
        os.makedirs(os.path.dirname(ORACLE_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-g", "-fPIC", "-std=gnu11", "-shared", "-o", ORACLE_LIB,
                               os.path.join(ORACLE_DIR, "zo_decode.c"), os.path.join(ORACLE_DIR, "zo_encode.c")])
    return ORACLE_LIB


def _sig(fn, res, args):
    fn.restype = res
    fn.argtypes = args


class Oracle:
    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        vp, st, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        _sig(L.zo_compress, st, [vp, st, vp, st, ci])
        _sig(L.zo_compress_advanced, st, [vp, st, vp, st, ci, ci])
        _sig(L.zo_decompress, st, [vp, st, vp, st])
        _sig(L.zo_decompress_usingDict, st, [vp, st, vp, st, vp, st])
        _sig(L.zo_createCCtx, vp, [])
        _sig(L.zo_freeCCtx, None, [vp])
        _sig(L.zo_compressCCtx, st, [vp, vp, st, vp, st, ci, ci])
        _sig(L.zo_createDCtx, vp, [])
        _sig(L.zo_freeDCtx, None, [vp])
        _sig(L.zo_decompressDCtx, st, [vp, vp, st, vp, st])
        _sig(L.zo_compressBound, st, [st])
        _sig(L.zo_decompressBound, ctypes.c_ulonglong, [vp, st])
        _sig(L.zo_isError, ctypes.c_uint, [st])
        _sig(L.zo_getErrorCode, ci, [st])
        _sig(L.zo_getErrorName, ctypes.c_char_p, [st])
        _sig(L.zo_getCParams, None, [ci, st, ctypes.POINTER(ctypes.c_uint)])
        _sig(L.zo_matchfinder_block, st, [ci, vp, st, vp, vp, ctypes.POINTER(st), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)])
        _sig(L.zo_decode_first_block_stages, st, [vp, st, vp, st, ctypes.POINTER(st), vp, st])
        _sig(L.zo_compress_usingLoadedDict, st, [vp, st, vp, st, vp, st, ci, ci])

    @staticmethod
    def _buf(data):
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        return a, (a.ctypes.data if a.size else 0)

    def compress_raw(self, data, level: int, cap: int | None = None, checksum: int = 0):
        a, p = self._buf(data)
        cap = self.lib.zo_compressBound(a.size) if cap is None else cap
        out = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.lib.zo_compress_advanced(out.ctypes.data, cap, p, a.size, level, checksum)
        return r, out

    def compress(self, data, level: int, checksum: int = 0) -> bytes:
        r, out = self.compress_raw(data, level, checksum=checksum)
        assert not self.lib.zo_isError(r), self.lib.zo_getErrorName(r)
        return out[:r].tobytes()

    def compress_loaded_dict_raw(self, data, level: int, dictionary, checksum: int = 0):
        """Compressor.LoadDictionary(dictionary) + Wrap(data) (Compressor.cs:43-56, 86-97)."""
        a, p = self._buf(data)
        d, dp = self._buf(dictionary)
        cap = self.lib.zo_compressBound(a.size)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.lib.zo_compress_usingLoadedDict(out.ctypes.data, cap, p, a.size, dp, d.size, level, checksum)
        return r, out

    def compress_loaded_dict(self, data, level: int, dictionary, checksum: int = 0) -> bytes:
        r, out = self.compress_loaded_dict_raw(data, level, dictionary, checksum)
        assert not self.lib.zo_isError(r), self.lib.zo_getErrorName(r)
        return out[:r].tobytes()

    def decompress_raw(self, frame, cap: int):
        a, p = self._buf(frame)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.lib.zo_decompress(out.ctypes.data, cap, p, a.size)
        return r, out

    def decompress(self, frame, cap: int) -> bytes:
        r, out = self.decompress_raw(frame, cap)
        assert not self.lib.zo_isError(r), self.lib.zo_getErrorName(r)
        return out[:r].tobytes()

    def decompress_using_dict_raw(self, frame, cap: int, dictionary):
        a, p = self._buf(frame)
        d, dp = self._buf(dictionary)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.lib.zo_decompress_usingDict(out.ctypes.data, cap, p, a.size, dp, d.size)
        return r, out

    def decompress_using_dict(self, frame, cap: int, dictionary) -> bytes:
        r, out = self.decompress_using_dict_raw(frame, cap, dictionary)
        assert not self.lib.zo_isError(r), self.lib.zo_getErrorName(r)
        return out[:r].tobytes()

    def error_code(self, rv: int) -> int:
        return self.lib.zo_getErrorCode(rv)

    def decompress_bound(self, frame) -> int:
        a, p = self._buf(frame)
        return int(self.lib.zo_decompressBound(p, a.size))

    def cparams(self, level: int, n: int):
        out = (ctypes.c_uint * 7)()
        self.lib.zo_getCParams(level, n, out)
        return tuple(out)

    def matchfinder(self, data, level: int):
        a, p = self._buf(data)
        seqs = np.zeros((a.size // 3 + 2, 2), dtype=np.uint32)   # zo_seqDef = {u32 offset; u16 ll; u16 ml}
        lits = np.zeros(a.size + 32, dtype=np.uint8)
        litSize = ctypes.c_size_t(0)
        ll = (ctypes.c_uint32 * 2)()
        rep = (ctypes.c_uint32 * 3)()
        n = self.lib.zo_matchfinder_block(level, p, a.size, seqs.ctypes.data, lits.ctypes.data, ctypes.byref(litSize), ll, rep)
        assert not self.lib.zo_isError(n)
        s = seqs[:n]
        off = s[:, 0].copy()
        llv = (s[:, 1] & 0xFFFF).astype(np.uint32)
        mlv = (s[:, 1] >> 16).astype(np.uint32)
        return off, llv, mlv, lits[:litSize.value].copy(), (ll[0], ll[1]), tuple(rep)

    def decode_stages(self, frame):
        a, p = self._buf(frame)
        lits = np.zeros(131072 + 64, dtype=np.uint8)
        litSize = ctypes.c_size_t(0)
        triples = np.zeros((65536, 3), dtype=np.uint32)
        n = self.lib.zo_decode_first_block_stages(p, a.size, lits.ctypes.data, lits.size, ctypes.byref(litSize), triples.ctypes.data, 65536)
        assert not self.lib.zo_isError(n), self.lib.zo_getErrorName(n)
        return lits[:litSize.value].copy(), triples[:n].copy()


class LibZstd:
    """System libzstd 1.5.5: the upstream C the reference is a translation of (at v1.5.1)."""

    def __init__(self):
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        self.lib = ctypes.CDLL(name)
        L = self.lib
        vp, st, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        _sig(L.ZSTD_compress, st, [vp, st, vp, st, ci])
        _sig(L.ZSTD_decompress, st, [vp, st, vp, st])
        _sig(L.ZSTD_compressBound, st, [st])
        _sig(L.ZSTD_isError, ctypes.c_uint, [st])
        _sig(L.ZSTD_getErrorCode, ci, [st])
        _sig(L.ZSTD_versionNumber, ctypes.c_uint, [])
        _sig(L.ZSTD_decompressBound, ctypes.c_ulonglong, [vp, st])
        _sig(L.ZSTD_createCCtx, vp, [])
        _sig(L.ZSTD_freeCCtx, st, [vp])
        _sig(L.ZSTD_CCtx_setParameter, st, [vp, ci, ci])
        _sig(L.ZSTD_compress2, st, [vp, vp, st, vp, st])
        _sig(L.ZSTD_compress_usingDict, st, [vp, vp, st, vp, st, vp, st, ci])
        _sig(L.ZSTD_createDCtx, vp, [])
        _sig(L.ZSTD_freeDCtx, st, [vp])
        _sig(L.ZSTD_decompress_usingDict, st, [vp, vp, st, vp, st, vp, st])
        _sig(L.ZDICT_trainFromBuffer, st, [vp, st, vp, ctypes.POINTER(st), ctypes.c_uint])
        _sig(L.ZDICT_isError, ctypes.c_uint, [st])

    def compress(self, data, level: int, checksum: int = 0) -> bytes:
        a, p = Oracle._buf(data)
        cap = self.lib.ZSTD_compressBound(a.size)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        if checksum:
            c = self.lib.ZSTD_createCCtx()
            self.lib.ZSTD_CCtx_setParameter(c, 100, level)
            self.lib.ZSTD_CCtx_setParameter(c, 201, 1)
            r = self.lib.ZSTD_compress2(c, out.ctypes.data, cap, p, a.size)
            self.lib.ZSTD_freeCCtx(c)
        else:
            r = self.lib.ZSTD_compress(out.ctypes.data, cap, p, a.size, level)
        assert not self.lib.ZSTD_isError(r)
        return out[:r].tobytes()

    def decompress_raw(self, frame, cap: int):
        a, p = Oracle._buf(frame)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.lib.ZSTD_decompress(out.ctypes.data, cap, p, a.size)
        return r, out

    def decompress(self, frame, cap: int) -> bytes:
        r, out = self.decompress_raw(frame, cap)
        assert not self.lib.ZSTD_isError(r)
        return out[:r].tobytes()

    def error_code(self, rv: int) -> int:
        return self.lib.ZSTD_getErrorCode(rv)

    # ---- dictionaries (test inputs only: ZDICT training and dictionary compression are not on the GPU path) ----
    def train_dictionary(self, samples, capacity: int) -> bytes:
        """ZDICT_trainFromBuffer: a zstd-format dictionary (magic 0xEC30A437, entropy tables, repcodes, content)."""
        blob = b"".join(bytes(x) for x in samples)
        sizes = (ctypes.c_size_t * len(samples))(*[len(bytes(x)) for x in samples])
        a = np.frombuffer(blob, dtype=np.uint8)
        out = np.empty(capacity, dtype=np.uint8)
        r = self.lib.ZDICT_trainFromBuffer(out.ctypes.data, capacity, a.ctypes.data, sizes, len(samples))
        assert not self.lib.ZDICT_isError(r), "ZDICT_trainFromBuffer failed"
        return out[:r].tobytes()

    def compress_using_dict(self, data, level: int, dictionary) -> bytes:
        a, p = Oracle._buf(data)
        d, dp = Oracle._buf(dictionary)
        cap = self.lib.ZSTD_compressBound(a.size)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        c = self.lib.ZSTD_createCCtx()
        r = self.lib.ZSTD_compress_usingDict(c, out.ctypes.data, cap, p, a.size, dp, d.size, level)
        self.lib.ZSTD_freeCCtx(c)
        assert not self.lib.ZSTD_isError(r)
        return out[:r].tobytes()

    def decompress_using_dict_raw(self, frame, cap: int, dictionary):
        a, p = Oracle._buf(frame)
        d, dp = Oracle._buf(dictionary)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        c = self.lib.ZSTD_createDCtx()
        r = self.lib.ZSTD_decompress_usingDict(c, out.ctypes.data, cap, p, a.size, dp, d.size)
        self.lib.ZSTD_freeDCtx(c)
        return r, out


REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libzstdref.so")
REF_DLL = "/root/reference/src/Zstd.Extern/libzstd.dll"


def build_ref() -> str | None:
    """oracle/_ref/libzstdref.so = PE mapper (oracle/ref_pe/peload.c) + the reference's libzstd.dll (zstd 1.5.1) embedded
    from where it lies.  Built here when /root/reference is present; on the GPU box only the prebuilt file exists."""
    src = os.path.join(ORACLE_DIR, "ref_pe", "peload.c")
    if os.path.exists(REF_DLL) and (not os.path.exists(REF_LIB) or os.path.getmtime(src) > os.path.getmtime(REF_LIB)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "_ref/libzstdref.so"])
    return REF_LIB if os.path.exists(REF_LIB) else None


class RefDll:
    """The reference's own native oracle: src/Zstd.Extern/libzstd.dll (zstd 1.5.1), the binary the reference's differential
    test compares ZstdSharp with byte for byte (ZstdTest.cs:18-90), run through oracle/ref_pe."""

    def __init__(self):
        path = build_ref()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libzstdref.so is not built and /root/reference is absent")
        self.lib = ctypes.CDLL(path)
        L = self.lib
        vp, st, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        assert L.ZREF_available() == 1, "PE image failed to load"
        _sig(L.ZREF_versionNumber, ctypes.c_uint, [])
        _sig(L.ZREF_compressBound, st, [st])
        _sig(L.ZREF_isError, ctypes.c_uint, [st])
        _sig(L.ZREF_getErrorCode, ci, [st])
        _sig(L.ZREF_getErrorName, ctypes.c_char_p, [st])
        _sig(L.ZREF_compress_level, st, [vp, st, vp, st, ci, ci])
        _sig(L.ZREF_decompress, st, [vp, st, vp, st])
        _sig(L.ZREF_createCCtx, vp, [])
        _sig(L.ZREF_freeCCtx, st, [vp])
        _sig(L.ZREF_createDCtx, vp, [])
        _sig(L.ZREF_freeDCtx, st, [vp])
        _sig(L.ZREF_compressCCtx, st, [vp, vp, st, vp, st, ci])
        _sig(L.ZREF_compress2, st, [vp, vp, st, vp, st])
        _sig(L.ZREF_CCtx_setParameter, st, [vp, ci, ci])
        _sig(L.ZREF_CCtx_loadDictionary, st, [vp, vp, st])
        _sig(L.ZREF_compress_usingDict8, st, [vp, vp, st, vp, st, vp, st, ci])
        _sig(L.ZREF_decompressDCtx, st, [vp, vp, st, vp, st])
        _sig(L.ZREF_decompress_usingDict, st, [vp, vp, st, vp, st, vp, st])
        _sig(L.ZREF_decompressBound, ctypes.c_ulonglong, [vp, st])
        _sig(L.ZREF_getFrameContentSize, ctypes.c_ulonglong, [vp, st])
        _sig(L.ZREF_findFrameCompressedSize, st, [vp, st])
        _sig(L.ZREF_getCParams, None, [ci, ctypes.c_ulonglong, st, ctypes.POINTER(ctypes.c_uint)])
        _sig(L.ZREF_generateSequences, st, [vp, vp, st, vp, st])
        _sig(L.ZREF_trainFromBuffer, st, [vp, st, vp, ctypes.POINTER(st), ctypes.c_uint])
        _sig(L.ZREF_decompressBatchMT, None, [st, vp, vp, vp, vp, vp, ci])
        _sig(L.ZREF_compressBatchMT, None, [st, vp, vp, vp, vp, vp, ci, ci])

    def version(self) -> int:
        return int(self.lib.ZREF_versionNumber())

    def compress_raw(self, data, level: int, cap: int | None = None, checksum: int = 0):
        a, p = Oracle._buf(data)
        cap = self.lib.ZREF_compressBound(a.size) if cap is None else cap
        out = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.lib.ZREF_compress_level(out.ctypes.data, cap, p, a.size, level, checksum)
        return r, out

    def compress(self, data, level: int, checksum: int = 0) -> bytes:
        r, out = self.compress_raw(data, level, checksum=checksum)
        assert not self.lib.ZREF_isError(r), self.lib.ZREF_getErrorName(r)
        return out[:r].tobytes()

    def decompress_raw(self, frame, cap: int):
        a, p = Oracle._buf(frame)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.lib.ZREF_decompress(out.ctypes.data, cap, p, a.size)
        return r, out

    def decompress(self, frame, cap: int) -> bytes:
        r, out = self.decompress_raw(frame, cap)
        assert not self.lib.ZREF_isError(r), self.lib.ZREF_getErrorName(r)
        return out[:r].tobytes()

    def error_code(self, rv: int) -> int:
        return self.lib.ZREF_getErrorCode(rv)

    def decompress_bound(self, frame) -> int:
        a, p = Oracle._buf(frame)
        return int(self.lib.ZREF_decompressBound(p, a.size))

    def cparams(self, level: int, n: int, dict_size: int = 0):
        out = (ctypes.c_uint * 7)()
        self.lib.ZREF_getCParams(level, n, dict_size, out)
        return tuple(out)

    def compress_using_dict(self, data, level: int, dictionary) -> bytes:
        a, p = Oracle._buf(data)
        d, dp = Oracle._buf(dictionary)
        cap = self.lib.ZREF_compressBound(a.size)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        c = self.lib.ZREF_createCCtx()
        r = self.lib.ZREF_compress_usingDict8(c, out.ctypes.data, cap, p, a.size, dp, d.size, level)
        self.lib.ZREF_freeCCtx(c)
        assert not self.lib.ZREF_isError(r), self.lib.ZREF_getErrorName(r)
        return out[:r].tobytes()

    def compress_loaded_dict(self, data, level: int, dictionary, checksum: int = 0) -> bytes:
        """Compressor.LoadDictionary + Wrap: ZSTD_CCtx_loadDictionary then ZSTD_compress2 (Compressor.cs:43-56, 86-97)."""
        a, p = Oracle._buf(data)
        d, dp = Oracle._buf(dictionary)
        cap = self.lib.ZREF_compressBound(a.size)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        c = self.lib.ZREF_createCCtx()
        self.lib.ZREF_CCtx_setParameter(c, 100, level)
        if checksum:
            self.lib.ZREF_CCtx_setParameter(c, 201, 1)
        r = self.lib.ZREF_CCtx_loadDictionary(c, dp, d.size)
        assert not self.lib.ZREF_isError(r), self.lib.ZREF_getErrorName(r)
        r = self.lib.ZREF_compress2(c, out.ctypes.data, cap, p, a.size)
        self.lib.ZREF_freeCCtx(c)
        assert not self.lib.ZREF_isError(r), self.lib.ZREF_getErrorName(r)
        return out[:r].tobytes()

    def decompress_using_dict_raw(self, frame, cap: int, dictionary):
        a, p = Oracle._buf(frame)
        d, dp = Oracle._buf(dictionary)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        c = self.lib.ZREF_createDCtx()
        r = self.lib.ZREF_decompress_usingDict(c, out.ctypes.data, cap, p, a.size, dp, d.size)
        self.lib.ZREF_freeDCtx(c)
        return r, out

    def generate_sequences(self, data, level: int):
        """ZSTD_generateSequences: (offset, litLength, matchLength, rep) per sequence, block delimiters included."""
        a, p = Oracle._buf(data)
        cap = a.size // 3 + 64
        seqs = np.zeros((cap, 4), dtype=np.uint32)
        c = self.lib.ZREF_createCCtx()
        self.lib.ZREF_CCtx_setParameter(c, 100, level)
        n = self.lib.ZREF_generateSequences(c, seqs.ctypes.data, cap, p, a.size)
        self.lib.ZREF_freeCCtx(c)
        assert not self.lib.ZREF_isError(n), self.lib.ZREF_getErrorName(n)
        return seqs[:n].copy()

    def train_dictionary(self, samples, capacity: int) -> bytes:
        blob = b"".join(bytes(x) for x in samples)
        sizes = (ctypes.c_size_t * len(samples))(*[len(bytes(x)) for x in samples])
        a = np.frombuffer(blob, dtype=np.uint8)
        out = np.empty(capacity, dtype=np.uint8)
        r = self.lib.ZREF_trainFromBuffer(out.ctypes.data, capacity, a.ctypes.data, sizes, len(samples))
        assert not self.lib.ZREF_isError(r), "ZDICT_trainFromBuffer failed"
        return out[:r].tobytes()


_ORACLE = None
_LIBZSTD = None
_REFDLL = None


def refdll_available() -> bool:
    return build_ref() is not None


def refdll() -> RefDll:
    global _REFDLL
    if _REFDLL is None:
        _REFDLL = RefDll()
    return _REFDLL


def oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def libzstd() -> LibZstd:
    global _LIBZSTD
    if _LIBZSTD is None:
        _LIBZSTD = LibZstd()
    return _LIBZSTD
