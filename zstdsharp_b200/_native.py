"""ctypes binding of libzstdb200.so (the C ABI declared in include/zstd_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C zstdsharp_b200/csrc`` into
``zstdsharp_b200/_build/libzstdb200.so``.  There is deliberately no fallback: if the shared object is missing the
import fails loudly, and if no CUDA device is usable every compress/decompress call returns ZSTD_error_GENERIC.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZSTDB200_LIB selects another build of the same library (the assertion build: make -C zstdsharp_b200/csrc debug)
LIB_PATH = os.environ.get("ZSTDB200_LIB") or os.path.join(_HERE, "_build", "libzstdb200.so")

TIMING_SLOTS = 12

# every symbol include/zstd_b200.h declares
EXPORTED_SYMBOLS = [
    "ZSTD_createCCtx", "ZSTD_freeCCtx", "ZSTD_compressCCtx", "ZSTD_compress2", "ZSTD_createDCtx", "ZSTD_freeDCtx",
    "ZSTD_decompressDCtx", "ZSTD_compressBound", "ZSTD_CCtx_setParameter", "ZSTD_decompressBound", "ZSTD_isError",
    "ZSTD_getErrorName", "ZSTD_versionNumber", "ZSTD_versionString", "ZSTD_findFrameCompressedSize",
    "ZSTDB200_decompressBatch", "ZSTDB200_compressBatch", "ZSTDB200_decompressBatchDevice",
    "ZSTDB200_compressBatchDevice", "ZSTDB200_getLastTimings", "ZSTDB200_getLastLaunchCount",
    "ZSTDB200_lastErrorString", "ZSTDB200_deviceCount", "ZSTDB200_setStream", "ZSTD_DCtx_loadDictionary", "ZSTD_CCtx_loadDictionary",
    "ZSTD_CCtx_getParameter", "ZSTD_DCtx_setParameter", "ZSTD_DCtx_getParameter",
    "ZSTDB200_createMulti", "ZSTDB200_freeMulti", "ZSTDB200_multiDeviceCount", "ZSTDB200_multiSetParameter",
    "ZSTDB200_multiLoadDictionary", "ZSTDB200_decompressBatchMulti", "ZSTDB200_compressBatchMulti",
    "ZSTDB200_shardBounds", "ZSTDB200_bindThreadToDevice",
]


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C zstdsharp_b200/csrc`). zstdsharp_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)
    c_size_t, c_void_p, c_int = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int
    P = ctypes.POINTER
    lib.ZSTD_createCCtx.restype = c_void_p
    lib.ZSTD_createCCtx.argtypes = []
    lib.ZSTD_freeCCtx.restype = c_size_t
    lib.ZSTD_freeCCtx.argtypes = [c_void_p]
    lib.ZSTD_compressCCtx.restype = c_size_t
    lib.ZSTD_compressCCtx.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int]
    lib.ZSTD_compress2.restype = c_size_t
    lib.ZSTD_compress2.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]
    lib.ZSTD_createDCtx.restype = c_void_p
    lib.ZSTD_createDCtx.argtypes = []
    lib.ZSTD_freeDCtx.restype = c_size_t
    lib.ZSTD_freeDCtx.argtypes = [c_void_p]
    lib.ZSTD_decompressDCtx.restype = c_size_t
    lib.ZSTD_decompressDCtx.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]
    lib.ZSTD_DCtx_loadDictionary.restype = c_size_t
    lib.ZSTD_DCtx_loadDictionary.argtypes = [c_void_p, c_void_p, c_size_t]
    lib.ZSTD_CCtx_loadDictionary.restype = c_size_t
    lib.ZSTD_CCtx_loadDictionary.argtypes = [c_void_p, c_void_p, c_size_t]
    lib.ZSTD_findFrameCompressedSize.restype = c_size_t
    lib.ZSTD_findFrameCompressedSize.argtypes = [c_void_p, c_size_t]
    lib.ZSTD_compressBound.restype = c_size_t
    lib.ZSTD_compressBound.argtypes = [c_size_t]
    lib.ZSTD_CCtx_setParameter.restype = c_size_t
    lib.ZSTD_CCtx_setParameter.argtypes = [c_void_p, c_int, c_int]
    lib.ZSTD_CCtx_getParameter.restype = c_size_t
    lib.ZSTD_CCtx_getParameter.argtypes = [c_void_p, c_int, ctypes.POINTER(c_int)]
    lib.ZSTD_DCtx_setParameter.restype = c_size_t
    lib.ZSTD_DCtx_setParameter.argtypes = [c_void_p, c_int, c_int]
    lib.ZSTD_DCtx_getParameter.restype = c_size_t
    lib.ZSTD_DCtx_getParameter.argtypes = [c_void_p, c_int, ctypes.POINTER(c_int)]
    lib.ZSTD_decompressBound.restype = ctypes.c_ulonglong
    lib.ZSTD_decompressBound.argtypes = [c_void_p, c_size_t]
    lib.ZSTD_isError.restype = ctypes.c_uint
    lib.ZSTD_isError.argtypes = [c_size_t]
    lib.ZSTD_getErrorName.restype = ctypes.c_char_p
    lib.ZSTD_getErrorName.argtypes = [c_size_t]
    lib.ZSTD_versionNumber.restype = ctypes.c_uint
    lib.ZSTD_versionString.restype = ctypes.c_char_p
    batch_host = [c_void_p, c_size_t, P(c_void_p), P(c_size_t), P(c_void_p), P(c_size_t), P(c_size_t)]
    lib.ZSTDB200_decompressBatch.restype = c_size_t
    lib.ZSTDB200_decompressBatch.argtypes = batch_host
    lib.ZSTDB200_compressBatch.restype = c_size_t
    lib.ZSTDB200_compressBatch.argtypes = [c_void_p, c_size_t, c_int, P(c_void_p), P(c_size_t), P(c_void_p), P(c_size_t), P(c_size_t)]
    dev = [c_void_p, P(ctypes.c_uint64), P(c_size_t), c_void_p, P(ctypes.c_uint64), P(c_size_t), P(c_size_t)]
    lib.ZSTDB200_decompressBatchDevice.restype = c_size_t
    lib.ZSTDB200_decompressBatchDevice.argtypes = [c_void_p, c_size_t] + dev
    lib.ZSTDB200_compressBatchDevice.restype = c_size_t
    lib.ZSTDB200_compressBatchDevice.argtypes = [c_void_p, c_size_t, c_int] + dev
    lib.ZSTDB200_getLastTimings.restype = None
    lib.ZSTDB200_getLastTimings.argtypes = [c_void_p, P(ctypes.c_float)]
    lib.ZSTDB200_getLastLaunchCount.restype = ctypes.c_uint
    lib.ZSTDB200_getLastLaunchCount.argtypes = [c_void_p]
    lib.ZSTDB200_lastErrorString.restype = ctypes.c_char_p
    lib.ZSTDB200_deviceCount.restype = c_int
    lib.ZSTDB200_setStream.restype = c_size_t
    lib.ZSTDB200_setStream.argtypes = [c_void_p, c_void_p]
    lib.ZSTDB200_createMulti.restype = c_void_p
    lib.ZSTDB200_createMulti.argtypes = [c_int]
    lib.ZSTDB200_freeMulti.restype = c_size_t
    lib.ZSTDB200_freeMulti.argtypes = [c_void_p]
    lib.ZSTDB200_multiDeviceCount.restype = c_int
    lib.ZSTDB200_multiDeviceCount.argtypes = [c_void_p]
    lib.ZSTDB200_multiSetParameter.restype = c_size_t
    lib.ZSTDB200_multiSetParameter.argtypes = [c_void_p, c_int, c_int]
    lib.ZSTDB200_multiLoadDictionary.restype = c_size_t
    lib.ZSTDB200_multiLoadDictionary.argtypes = [c_void_p, c_void_p, c_size_t]
    lib.ZSTDB200_decompressBatchMulti.restype = c_size_t
    lib.ZSTDB200_decompressBatchMulti.argtypes = batch_host
    lib.ZSTDB200_compressBatchMulti.restype = c_size_t
    lib.ZSTDB200_compressBatchMulti.argtypes = [c_void_p, c_size_t, c_int, P(c_void_p), P(c_size_t), P(c_void_p), P(c_size_t), P(c_size_t)]
    lib.ZSTDB200_shardBounds.restype = None
    lib.ZSTDB200_shardBounds.argtypes = [c_size_t, P(c_size_t), c_int, P(c_size_t)]
    lib.ZSTDB200_bindThreadToDevice.restype = c_int
    lib.ZSTDB200_bindThreadToDevice.argtypes = [c_int]
    return lib


lib = _load()
