"""Host-side mirror of the reference's safe wrappers, on top of the C ABI.

Mirrors (same names, argument meaning and error behaviour):
  * ``Compressor``   -- /root/reference/src/ZstdSharp/Compressor.cs:6-164   (Level, SetParameter, Wrap, TryWrap, GetCompressBound)
  * ``Decompressor`` -- /root/reference/src/ZstdSharp/Decompressor.cs:6-149 (GetDecompressedSize, Unwrap, TryUnwrap)
  * ``ZstdException``-- /root/reference/src/ZstdSharp/ZstdException.cs, ThrowHelper.cs:10-41
plus the batch calls the reference does not have (``WrapBatch`` / ``UnwrapBatch``), which is where the GPU pays off.
The reference toolchain (.NET) is absent from this image, so the wrappers are Python; the C# binding a maintainer
would add is shown in INTEGRATION.md.
"""
from __future__ import annotations

import ctypes
import enum
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native

_lib = _native.lib

MAX_BYTE_ARRAY_LENGTH = 0x7FFFFFC7          # Constants.cs:7
CONTENTSIZE_UNKNOWN = (1 << 64) - 1
CONTENTSIZE_ERROR = (1 << 64) - 2


class ZSTD_ErrorCode(enum.IntEnum):         # Unsafe/ZSTD_ErrorCode.cs:5-35
    no_error = 0
    GENERIC = 1
    prefix_unknown = 10
    version_unsupported = 12
    frameParameter_unsupported = 14
    frameParameter_windowTooLarge = 16
    corruption_detected = 20
    checksum_wrong = 22
    dictionary_corrupted = 30
    dictionary_wrong = 32
    dictionaryCreation_failed = 34
    parameter_unsupported = 40
    parameter_outOfBound = 42
    tableLog_tooLarge = 44
    maxSymbolValue_tooLarge = 46
    maxSymbolValue_tooSmall = 48
    stage_wrong = 60
    init_missing = 62
    memory_allocation = 64
    workSpace_tooSmall = 66
    dstSize_tooSmall = 70
    srcSize_wrong = 72
    dstBuffer_null = 74
    frameIndex_tooLarge = 100
    seekableIO = 102
    dstBuffer_wrong = 104
    srcBuffer_wrong = 105
    maxCode = 120


class ZSTD_cParameter(enum.IntEnum):        # Unsafe/ZSTD_cParameter.cs (the ones the GPU path understands)
    ZSTD_c_compressionLevel = 100
    ZSTD_c_contentSizeFlag = 200
    ZSTD_c_checksumFlag = 201
    ZSTDB200_c_independentChunks = 10001     # new (include/zstd_b200.h): 1 = cut items above 128 KiB into independent frames


class ZSTD_dParameter(enum.IntEnum):        # Unsafe/ZSTD_dParameter.cs
    ZSTD_d_windowLogMax = 100


class ZstdException(Exception):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        try:
            self.Code = ZSTD_ErrorCode(code)
        except ValueError:
            self.Code = code


class ObjectDisposedException(Exception):
    pass


def is_error(rv: int) -> bool:
    return bool(_lib.ZSTD_isError(rv))


def error_code(rv: int) -> int:
    return ((1 << 64) - rv) if is_error(rv) else 0


def EnsureZstdSuccess(rv: int) -> int:      # ThrowHelper.cs:10-16
    if is_error(rv):
        raise ZstdException(error_code(rv), _lib.ZSTD_getErrorName(rv).decode())
    return rv


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8 or not buf.flags["C_CONTIGUOUS"]:
            buf = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
        return buf
    return np.frombuffer(buf, dtype=np.uint8)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data if a.size else 0


class Compressor:
    MinCompressionLevel = -(1 << 17)         # Compressor.cs:12 ZSTD_minCLevel(): the negative levels are ZSTD_fast with an acceleration factor
    MaxCompressionLevel = 4                  # GPU path: ZSTD_fast / ZSTD_dfast levels only (4 only for the input sizes where it is still ZSTD_dfast)
    DefaultCompressionLevel = 0              # Compressor.cs:10 (0 means level 3)

    def __init__(self, level: int = DefaultCompressionLevel):
        self._level = self.DefaultCompressionLevel
        self._cctx = _lib.ZSTD_createCCtx()
        if not self._cctx:
            raise ZstdException(ZSTD_ErrorCode.GENERIC, "Failed to create cctx")
        self.Level = level

    # -- parameters (Compressor.cs:16-41)
    @property
    def Level(self) -> int:
        return self._level

    @Level.setter
    def Level(self, value: int) -> None:
        if self._level != value:
            self._level = value
            self.SetParameter(ZSTD_cParameter.ZSTD_c_compressionLevel, value)

    def SetParameter(self, parameter: int, value: int) -> None:
        self._ensure()
        EnsureZstdSuccess(_lib.ZSTD_CCtx_setParameter(self._cctx, int(parameter), int(value)))

    def LoadDictionary(self, dictionary) -> None:    # Compressor.cs:43-56: null / empty removes the dictionary
        self._ensure()
        if dictionary is None or len(dictionary) == 0:
            EnsureZstdSuccess(_lib.ZSTD_CCtx_loadDictionary(self._cctx, 0, 0))
            return
        d = _as_u8(dictionary)
        EnsureZstdSuccess(_lib.ZSTD_CCtx_loadDictionary(self._cctx, _ptr(d), d.size))

    def GetParameter(self, parameter: int) -> int:   # Compressor.cs:35-41
        self._ensure()
        value = ctypes.c_int(0)
        EnsureZstdSuccess(_lib.ZSTD_CCtx_getParameter(self._cctx, int(parameter), ctypes.byref(value)))
        return value.value

    @staticmethod
    def GetCompressBound(length: int) -> int:
        return int(_lib.ZSTD_compressBound(length))

    # -- one-shot (Compressor.cs:78-126)
    def Wrap(self, src, dest: Optional[np.ndarray] = None):
        self._ensure()
        s = _as_u8(src)
        if dest is None:
            d = np.empty(self.GetCompressBound(s.size), dtype=np.uint8)
            n = EnsureZstdSuccess(_lib.ZSTD_compress2(self._cctx, _ptr(d), d.size, _ptr(s), s.size))
            return d[:n].tobytes()
        d = dest
        return EnsureZstdSuccess(_lib.ZSTD_compress2(self._cctx, _ptr(d), d.size, _ptr(s), s.size))

    def TryWrap(self, src, dest: np.ndarray) -> Tuple[bool, int]:
        self._ensure()
        s = _as_u8(src)
        rv = _lib.ZSTD_compress2(self._cctx, _ptr(dest), dest.size, _ptr(s), s.size)
        if is_error(rv) and error_code(rv) == ZSTD_ErrorCode.dstSize_tooSmall:
            return False, 0
        return True, EnsureZstdSuccess(rv)

    # -- batch (new): every chunk becomes its own frame, exactly as Wrap(chunk) would produce it
    def WrapBatch(self, chunks: Sequence) -> List[bytes]:
        self._ensure()
        srcs = [_as_u8(c) for c in chunks]
        n = len(srcs)
        if n == 0:
            return []
        caps = [self.GetCompressBound(s.size) for s in srcs]
        offs = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
        out = np.empty(int(offs[-1]), dtype=np.uint8)
        sp = (ctypes.c_void_p * n)(*[_ptr(s) for s in srcs])
        ss = (ctypes.c_size_t * n)(*[s.size for s in srcs])
        dp = (ctypes.c_void_p * n)(*[out.ctypes.data + int(o) for o in offs[:-1]])
        dc = (ctypes.c_size_t * n)(*caps)
        res = (ctypes.c_size_t * n)()
        level = 3 if self._level == 0 else self._level
        EnsureZstdSuccess(_lib.ZSTDB200_compressBatch(self._cctx, n, level, sp, ss, dp, dc, res))
        frames = []
        for i in range(n):
            EnsureZstdSuccess(res[i])
            frames.append(out[int(offs[i]):int(offs[i]) + res[i]].tobytes())
        return frames

    def timings(self) -> List[float]:
        t = (ctypes.c_float * _native.TIMING_SLOTS)()
        _lib.ZSTDB200_getLastTimings(self._cctx, t)
        return list(t)

    def launch_count(self) -> int:
        return int(_lib.ZSTDB200_getLastLaunchCount(self._cctx))

    @property
    def handle(self) -> int:
        return self._cctx

    def Dispose(self) -> None:
        if self._cctx:
            _lib.ZSTD_freeCCtx(self._cctx)
            self._cctx = None

    def __del__(self):
        try:
            self.Dispose()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.Dispose()

    def _ensure(self) -> None:
        if not self._cctx:
            raise ObjectDisposedException("Compressor")


class Decompressor:
    def __init__(self):
        self._dctx = _lib.ZSTD_createDCtx()
        if not self._dctx:
            raise ZstdException(ZSTD_ErrorCode.GENERIC, "Failed to create dctx")

    def SetParameter(self, parameter: int, value: int) -> None:   # Decompressor.cs:22-26
        self._ensure()
        EnsureZstdSuccess(_lib.ZSTD_DCtx_setParameter(self._dctx, int(parameter), int(value)))

    def GetParameter(self, parameter: int) -> int:   # Decompressor.cs:28-34
        self._ensure()
        value = ctypes.c_int(0)
        EnsureZstdSuccess(_lib.ZSTD_DCtx_getParameter(self._dctx, int(parameter), ctypes.byref(value)))
        return value.value

    def LoadDictionary(self, dictionary) -> None:   # Decompressor.cs:43-56: null / empty removes the dictionary
        self._ensure()
        if dictionary is None or len(dictionary) == 0:
            EnsureZstdSuccess(_lib.ZSTD_DCtx_loadDictionary(self._dctx, 0, 0))
            return
        d = _as_u8(dictionary)
        EnsureZstdSuccess(_lib.ZSTD_DCtx_loadDictionary(self._dctx, _ptr(d), d.size))

    @staticmethod
    def GetDecompressedSize(src) -> int:    # Decompressor.cs:50-54 + ThrowHelper.EnsureContentSizeOk :26-35
        s = _as_u8(src)
        rv = int(_lib.ZSTD_decompressBound(_ptr(s), s.size))
        if rv == CONTENTSIZE_UNKNOWN:
            raise ZstdException(ZSTD_ErrorCode.GENERIC, "Decompressed content size is not specified")
        if rv == CONTENTSIZE_ERROR:
            raise ZstdException(ZSTD_ErrorCode.GENERIC,
                                "Decompressed content size cannot be determined (e.g. invalid magic number, srcSize too small)")
        return rv

    def Unwrap(self, src, dest: Optional[np.ndarray] = None, maxDecompressedSize: int = 2 ** 31 - 1):
        self._ensure()
        s = _as_u8(src)
        if dest is None:                     # Decompressor.cs:62-78
            expected = self.GetDecompressedSize(s)
            if expected > maxDecompressedSize:
                raise ZstdException(ZSTD_ErrorCode.dstSize_tooSmall,
                                    f"Decompressed content size {expected} is greater than maxDecompressedSize {maxDecompressedSize}")
            if expected > MAX_BYTE_ARRAY_LENGTH:
                raise ZstdException(ZSTD_ErrorCode.dstSize_tooSmall,
                                    f"Decompressed content size {expected} is greater than max possible byte array size {MAX_BYTE_ARRAY_LENGTH}")
            d = np.empty(expected, dtype=np.uint8)
            n = EnsureZstdSuccess(_lib.ZSTD_decompressDCtx(self._dctx, _ptr(d), d.size, _ptr(s), s.size))
            return d[:n].tobytes()
        return EnsureZstdSuccess(_lib.ZSTD_decompressDCtx(self._dctx, _ptr(dest), dest.size, _ptr(s), s.size))

    def TryUnwrap(self, src, dest: np.ndarray) -> Tuple[bool, int]:
        self._ensure()
        s = _as_u8(src)
        rv = _lib.ZSTD_decompressDCtx(self._dctx, _ptr(dest), dest.size, _ptr(s), s.size)
        if is_error(rv) and error_code(rv) == ZSTD_ErrorCode.dstSize_tooSmall:
            return False, 0
        return True, EnsureZstdSuccess(rv)

    def UnwrapBatch(self, frames: Sequence, raise_on_error: bool = True, max_decompressed_size: int = MAX_BYTE_ARRAY_LENGTH,
                    capacity: Optional[int] = None):
        """Decompresses every element of ``frames`` as ``Unwrap`` would. Returns a list of bytes (or, with
        ``raise_on_error=False``, of bytes / ZstdException per item).  ``max_decompressed_size`` is Unwrap's guard
        (Decompressor.cs:47-59) applied per item: an item whose declared size exceeds it fails with dstSize_tooSmall and is
        not staged, so one frame declaring a huge content size never poisons the batch.  ``capacity`` instead gives every item
        exactly that many output bytes (what ``Unwrap(src, dest)`` with a caller buffer does), whatever its header declares."""
        self._ensure()
        srcs = [_as_u8(f) for f in frames]
        n = len(srcs)
        if n == 0:
            return []
        caps, oversize = [], []
        for s in srcs:
            if capacity is not None:
                caps.append(int(capacity)); oversize.append(False)
                continue
            b = int(_lib.ZSTD_decompressBound(_ptr(s), s.size))
            b = 0 if b >= CONTENTSIZE_ERROR else b
            oversize.append(b > max_decompressed_size)
            caps.append(0 if oversize[-1] else b)
        offs = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
        out = np.empty(max(int(offs[-1]), 1), dtype=np.uint8)
        sp = (ctypes.c_void_p * n)(*[_ptr(s) for s in srcs])
        ss = (ctypes.c_size_t * n)(*[0 if o else s.size for s, o in zip(srcs, oversize)])
        dp = (ctypes.c_void_p * n)(*[out.ctypes.data + int(o) for o in offs[:-1]])
        dc = (ctypes.c_size_t * n)(*caps)
        res = (ctypes.c_size_t * n)()
        EnsureZstdSuccess(_lib.ZSTDB200_decompressBatch(self._dctx, n, sp, ss, dp, dc, res))
        outs = []
        for i in range(n):
            code = int(ZSTD_ErrorCode.dstSize_tooSmall) if oversize[i] else (error_code(res[i]) if is_error(res[i]) else 0)
            if code:
                exc = ZstdException(code, _lib.ZSTD_getErrorName((1 << 64) - code).decode())
                if raise_on_error:
                    raise exc
                outs.append(exc)
            else:
                outs.append(out[int(offs[i]):int(offs[i]) + res[i]].tobytes())
        return outs

    def timings(self) -> List[float]:
        t = (ctypes.c_float * _native.TIMING_SLOTS)()
        _lib.ZSTDB200_getLastTimings(self._dctx, t)
        return list(t)

    def launch_count(self) -> int:
        return int(_lib.ZSTDB200_getLastLaunchCount(self._dctx))

    @property
    def handle(self) -> int:
        return self._dctx

    def Dispose(self) -> None:
        if self._dctx:
            _lib.ZSTD_freeDCtx(self._dctx)
            self._dctx = None

    def __del__(self):
        try:
            self.Dispose()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.Dispose()

    def _ensure(self) -> None:
        if not self._dctx:
            raise ObjectDisposedException("Decompressor")


class MultiCodec:
    """The batch scheduler across the GPUs of one box (BASELINE.json north_star: "partitioned across the 8 GPUs of one box by a
    host scatter", no collective): ``WrapBatch`` / ``UnwrapBatch`` with the meaning they have on ``Compressor`` /
    ``Decompressor``, every item handled by exactly one device, results in the caller's order (ZSTDB200_*BatchMulti)."""

    def __init__(self, n_devices: int = 0, level: int = 1):
        self._m = _lib.ZSTDB200_createMulti(int(n_devices))
        if not self._m:
            raise ZstdException(ZSTD_ErrorCode.GENERIC, "Failed to create the multi-device scheduler: " + _lib.ZSTDB200_lastErrorString().decode())
        self.Level = level

    @property
    def DeviceCount(self) -> int:
        return int(_lib.ZSTDB200_multiDeviceCount(self._m))

    def SetParameter(self, parameter: int, value: int) -> None:
        EnsureZstdSuccess(_lib.ZSTDB200_multiSetParameter(self._m, int(parameter), int(value)))

    def LoadDictionary(self, dictionary) -> None:
        if dictionary is None or len(dictionary) == 0:
            EnsureZstdSuccess(_lib.ZSTDB200_multiLoadDictionary(self._m, 0, 0))
            return
        d = _as_u8(dictionary)
        EnsureZstdSuccess(_lib.ZSTDB200_multiLoadDictionary(self._m, _ptr(d), d.size))

    def WrapBatch(self, chunks: Sequence) -> List[bytes]:
        srcs = [_as_u8(c) for c in chunks]
        n = len(srcs)
        if n == 0:
            return []
        caps = [Compressor.GetCompressBound(s.size) for s in srcs]
        offs = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
        out = np.empty(int(offs[-1]), dtype=np.uint8)
        sp = (ctypes.c_void_p * n)(*[_ptr(s) for s in srcs])
        ss = (ctypes.c_size_t * n)(*[s.size for s in srcs])
        dp = (ctypes.c_void_p * n)(*[out.ctypes.data + int(o) for o in offs[:-1]])
        dc = (ctypes.c_size_t * n)(*caps)
        res = (ctypes.c_size_t * n)()
        level = 3 if self.Level == 0 else self.Level
        EnsureZstdSuccess(_lib.ZSTDB200_compressBatchMulti(self._m, n, level, sp, ss, dp, dc, res))
        frames = []
        for i in range(n):
            EnsureZstdSuccess(res[i])
            frames.append(out[int(offs[i]):int(offs[i]) + res[i]].tobytes())
        return frames

    def UnwrapBatch(self, frames: Sequence, raise_on_error: bool = True):
        srcs = [_as_u8(f) for f in frames]
        n = len(srcs)
        if n == 0:
            return []
        caps = []
        for s in srcs:
            b = int(_lib.ZSTD_decompressBound(_ptr(s), s.size))
            caps.append(0 if b >= CONTENTSIZE_ERROR else b)
        offs = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
        out = np.empty(max(int(offs[-1]), 1), dtype=np.uint8)
        sp = (ctypes.c_void_p * n)(*[_ptr(s) for s in srcs])
        ss = (ctypes.c_size_t * n)(*[s.size for s in srcs])
        dp = (ctypes.c_void_p * n)(*[out.ctypes.data + int(o) for o in offs[:-1]])
        dc = (ctypes.c_size_t * n)(*caps)
        res = (ctypes.c_size_t * n)()
        EnsureZstdSuccess(_lib.ZSTDB200_decompressBatchMulti(self._m, n, sp, ss, dp, dc, res))
        outs = []
        for i in range(n):
            if is_error(res[i]):
                exc = ZstdException(error_code(res[i]), _lib.ZSTD_getErrorName(res[i]).decode())
                if raise_on_error:
                    raise exc
                outs.append(exc)
            else:
                outs.append(out[int(offs[i]):int(offs[i]) + res[i]].tobytes())
        return outs

    def Dispose(self) -> None:
        if self._m:
            _lib.ZSTDB200_freeMulti(self._m)
            self._m = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.Dispose()

    def __del__(self):
        try:
            self.Dispose()
        except Exception:
            pass
