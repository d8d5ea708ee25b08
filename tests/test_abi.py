"""C-ABI surface: the shared library loads and exports every symbol include/zstd_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import numpy as np

from zstdsharp_b200 import _native, api
from zstdsharp_b200 import datagen as dg

from _oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "zstd_b200.h")).read()
    declared = set(re.findall(r"\b(ZSTD_\w+|ZSTDB200_\w+)\s*\(", hdr)) - {"ZSTDB200_API"}
    declared = {d for d in declared if not d.isupper()}
    assert len(declared) >= 22
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_native.EXPORTED_SYMBOLS)


def test_host_side_functions_without_gpu():
    lib = _native.lib
    o = oracle()
    for n in (0, 1, 100, 4096, 131072, 131073, 1 << 20):
        assert lib.ZSTD_compressBound(n) == o.lib.zo_compressBound(n)
    assert lib.ZSTD_versionNumber() == 10501 and lib.ZSTD_versionString() == b"1.5.1"
    assert lib.ZSTD_isError((1 << 64) - 70) == 1 and lib.ZSTD_isError(131072) == 0
    assert lib.ZSTD_getErrorName((1 << 64) - 70) == b"Destination buffer is too small" == o.lib.zo_getErrorName((1 << 64) - 70)
    for code in (1, 10, 12, 14, 16, 20, 22, 30, 32, 40, 42, 44, 46, 48, 60, 62, 64, 66, 70, 72, 74, 100, 102, 104, 105):
        assert lib.ZSTD_getErrorName((1 << 64) - code) == o.lib.zo_getErrorName((1 << 64) - code)
    # ZSTD_decompressBound is host logic (header walk): compare with the oracle on good, concatenated and broken input
    text = dg.text_like(3 * dg.FRAME)
    f1, f3 = o.compress(text[: dg.FRAME], 1), o.compress(text, 3)
    skippable = (0x184D2A50).to_bytes(4, "little") + (3).to_bytes(4, "little") + b"xyz"
    for blob in (f1, f3, f1 + skippable + f3, f1[:50], b"", b"\x01\x02\x03\x04\x05\x06", skippable):
        a = np.frombuffer(blob, dtype=np.uint8)
        got = lib.ZSTD_decompressBound(a.ctypes.data if a.size else 0, a.size)
        assert got == o.decompress_bound(blob), blob[:8]
    assert api.Decompressor.GetDecompressedSize(f3) == text.size


def test_parameter_surface():
    c = api.Compressor(1)
    assert c.Level == 1
    c.Level = 3
    c.SetParameter(api.ZSTD_cParameter.ZSTD_c_checksumFlag, 1)
    c.SetParameter(api.ZSTD_cParameter.ZSTD_c_checksumFlag, 0)
    try:
        c.SetParameter(api.ZSTD_cParameter.ZSTD_c_checksumFlag, 2)
        raise AssertionError("expected parameter_outOfBound")
    except api.ZstdException as e:
        assert e.Code == api.ZSTD_ErrorCode.parameter_outOfBound
    # new parameter of this library (include/zstd_b200.h): the value in the header, the Python mirror and the library agree
    hdr = open(os.path.join(ROOT, "include", "zstd_b200.h")).read()
    assert int(re.search(r"#define\s+ZSTDB200_c_independentChunks\s+(\d+)", hdr).group(1)) == api.ZSTD_cParameter.ZSTDB200_c_independentChunks
    c.SetParameter(api.ZSTD_cParameter.ZSTDB200_c_independentChunks, 1)
    c.SetParameter(api.ZSTD_cParameter.ZSTDB200_c_independentChunks, 0)
    try:
        c.SetParameter(api.ZSTD_cParameter.ZSTDB200_c_independentChunks, 2)
        raise AssertionError("expected parameter_outOfBound")
    except api.ZstdException as e:
        assert e.Code == api.ZSTD_ErrorCode.parameter_outOfBound
    for bad in ((api.ZSTD_cParameter.ZSTD_c_compressionLevel, 7), (160, 1)):
        try:
            c.SetParameter(*bad)
            raise AssertionError("expected parameter_unsupported")
        except api.ZstdException as e:
            assert e.Code == api.ZSTD_ErrorCode.parameter_unsupported
    # Compressor.GetParameter (Compressor.cs:35-41): level 0 reads back as ZSTD_CLEVEL_DEFAULT
    assert c.GetParameter(api.ZSTD_cParameter.ZSTD_c_compressionLevel) == 3
    c.Level = 0
    assert c.GetParameter(api.ZSTD_cParameter.ZSTD_c_compressionLevel) == 3
    c.Level = 2
    assert c.GetParameter(api.ZSTD_cParameter.ZSTD_c_compressionLevel) == 2
    assert c.GetParameter(api.ZSTD_cParameter.ZSTD_c_checksumFlag) == 0
    # Decompressor.SetParameter / GetParameter (Decompressor.cs:22-34; bounds U/ZstdDecompress.cs:2401-2407, default :2543-2546)
    d = api.Decompressor()
    W = api.ZSTD_dParameter.ZSTD_d_windowLogMax
    assert d.GetParameter(W) == 27
    d.SetParameter(W, 10)
    assert d.GetParameter(W) == 10
    d.SetParameter(W, 0)
    assert d.GetParameter(W) == 27
    for bad in (9, 32, -1):
        try:
            d.SetParameter(W, bad)
            raise AssertionError("expected parameter_outOfBound")
        except api.ZstdException as e:
            assert e.Code == api.ZSTD_ErrorCode.parameter_outOfBound
    c.Dispose()
    try:
        c.Wrap(b"abc")
        raise AssertionError("expected ObjectDisposedException")
    except api.ObjectDisposedException:
        pass


def test_no_silent_cpu_fallback():
    """Without a CUDA device the compute entry points must fail loudly (ZSTD_error_GENERIC), never emulate on the CPU."""
    if _native.lib.ZSTDB200_deviceCount() > 0:
        return
    d = api.Decompressor()
    try:
        d.Unwrap(bytes.fromhex("28b52ffd2000010000"))
        raise AssertionError("decompression must not succeed without a GPU")
    except api.ZstdException as e:
        assert e.Code == api.ZSTD_ErrorCode.GENERIC
    assert b"CUDA" in _native.lib.ZSTDB200_lastErrorString()
