"""GPU parity: every frame the GPU compressor emits must be byte-identical to the oracle's Compressor.Wrap output
(levels 1..3, one frame per chunk), must round-trip through libzstd 1.5.5 and through the GPU decoder.

Mirrors ZstdTest.CompressAndDecompressWithNative (src/ZstdSharp.Test/ZstdTest.cs:69-90: byte-identical output,
equal compressBound, equal decompressed bytes) with the oracle in the native library's role.
"""
import numpy as np
import pytest

from zstdsharp_b200 import datagen as dg

from _oracle import oracle, libzstd
from _cases import multiblock_inputs

pytestmark = pytest.mark.gpu
FRAME = dg.FRAME


@pytest.fixture(scope="module")
def comp():
    from zstdsharp_b200 import Compressor
    c = Compressor(1)
    yield c
    c.Dispose()


@pytest.fixture(scope="module")
def dec():
    from zstdsharp_b200 import Decompressor
    d = Decompressor()
    yield d
    d.Dispose()


def _chunks(data, size=FRAME):
    return [data[i:i + size] for i in range(0, data.size, size)]


def _first_diff(a: bytes, b: bytes) -> str:
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return f"first diff at byte {i} of {len(a)}/{len(b)}: {a[max(0,i-4):i+8].hex()} vs {b[max(0,i-4):i+8].hex()}"
    return f"length {len(a)} vs {len(b)}"


@pytest.mark.parametrize("workload", ["text", "silesia", "incompressible", "literal_heavy", "literal_mix"])
@pytest.mark.parametrize("level", [1, 2, 3])
def test_byte_identical_to_oracle(comp, dec, workload, level):
    o, z = oracle(), libzstd()
    data = dg.WORKLOADS[workload](12 * FRAME)
    chunks = _chunks(data)
    comp.Level = level
    frames = comp.WrapBatch(chunks)
    assert comp.launch_count() > 0
    for c, f in zip(chunks, frames):
        want = o.compress(c, level)
        assert f == want, _first_diff(f, want)
        assert z.decompress(f, FRAME) == c.tobytes()
    assert dec.UnwrapBatch(frames) == [c.tobytes() for c in chunks]


@pytest.mark.parametrize("level", [1, 3])
def test_sizes_sweep(comp, level):
    """ZstdNetTests.cs:456-496: empty, 1 byte, sizes 2..100000 step 3000 of (byte)i; plus parameter-bucket edges
    (16 KiB / 128 KiB tables, 1 KiB / 16 KiB literal headers, 256-literal single-stream limit, 40960 sampling gate)."""
    o = oracle()
    comp.Level = level
    sizes = [0, 1, 2, 3, 6, 7, 8, 63, 64, 65, 255, 256, 257, 1023, 1024, 1025, 4095, 4096, 16383, 16384, 16385,
             40959, 40960, 65535, 65536, 65791, 65792, 100001, 131071, 131072] + list(range(2, 100000, 3000))
    text = dg.text_like(2 * FRAME)
    srcs = [dg.byte_ramp(n) for n in sizes] + [text[:n] for n in sizes] + [dg.literal_heavy(FRAME)[:n] for n in sizes[:30]]
    frames = comp.WrapBatch(srcs)
    for s, f in zip(srcs, frames):
        want = o.compress(s, level)
        assert f == want, f"size {s.size}: " + _first_diff(f, want)


def test_special_inputs(comp):
    o = oracle()
    rng = np.random.default_rng(5)
    specials = [
        np.zeros(FRAME, dtype=np.uint8),                                   # one 131069-byte match (long-length escape)
        np.full(FRAME, 0x41, dtype=np.uint8),
        np.tile(np.frombuffer(b"abcdefgh", dtype=np.uint8), FRAME // 8),
        np.tile(rng.integers(0, 256, 1000, dtype=np.uint8), 132)[:FRAME],  # long repeats at distance 1000
        np.concatenate([rng.integers(0, 256, 70000, dtype=np.uint8), np.zeros(FRAME - 70000, dtype=np.uint8)]),   # litLength > 65535
        np.concatenate([np.zeros(60000, dtype=np.uint8), rng.integers(0, 256, FRAME - 60000, dtype=np.uint8)]),
        rng.integers(0, 4, FRAME, dtype=np.uint8),                         # ACGT-like, very short matches
        rng.integers(0, 2, FRAME, dtype=np.uint8),
        (np.arange(FRAME) // 7 % 251).astype(np.uint8),
    ]
    for level in (1, 2, 3):
        comp.Level = level
        frames = comp.WrapBatch(specials)
        for s, f in zip(specials, frames):
            want = o.compress(s, level)
            assert f == want, _first_diff(f, want)


def test_single_call_and_bounds(comp):
    from zstdsharp_b200 import ZstdException, ZSTD_ErrorCode, Compressor
    o = oracle()
    src = dg.text_like(FRAME)
    comp.Level = 1
    assert comp.Wrap(src) == o.compress(src, 1)
    assert Compressor.GetCompressBound(FRAME) == 131584 == o.lib.zo_compressBound(FRAME)
    for n in (0, 1, 100, 1000, 65536, 200000):
        assert Compressor.GetCompressBound(n) == o.lib.zo_compressBound(n)
    # destination too small -> code 70; TryWrap returns False (ZstdNetTests.cs:214-258)
    small = np.empty(20, dtype=np.uint8)
    with pytest.raises(ZstdException) as e:
        comp.Wrap(src, small)
    assert e.value.Code == ZSTD_ErrorCode.dstSize_tooSmall
    assert comp.TryWrap(src, small) == (False, 0)
    # level 0 means the default level 3 (ZstdCompress.cs:895-905)
    c0 = Compressor(0)
    assert c0.Wrap(src) == o.compress(src, 3)
    c0.Dispose()
    # levels the GPU path does not implement are refused, not silently downgraded
    with pytest.raises(ZstdException) as e:
        Compressor(5)
    assert e.value.Code == ZSTD_ErrorCode.parameter_unsupported


def test_frame_header_known_answers(comp):
    """Frame-header descriptor bytes (ZstdNetTests.cs:194-204 pins 0x60 for a small no-dict frame? -> single segment):
    128 KiB chunk: 28 B5 2F FD A0 00 00 02 00."""
    comp.Level = 1
    f = comp.Wrap(dg.text_like(FRAME))
    assert f[:9] == bytes.fromhex("28b52ffda000000200")
    g = comp.Wrap(bytes(range(100)))
    assert g[:6] == bytes.fromhex("28b52ffd2064")
    h = comp.Wrap(bytes(300))
    assert h[:7] == bytes.fromhex("28b52ffd602c00")


def test_checksum_flag(comp, dec):
    """ZSTD_c_checksumFlag = 1: 4 more bytes per frame (ZstdNetTests.cs:65), byte-identical to the oracle,
    accepted by libzstd (which verifies the XXH64 trailer) and by the GPU decoder."""
    from zstdsharp_b200 import ZSTD_cParameter
    o, z = oracle(), libzstd()
    data = dg.silesia_mix(6 * FRAME)
    chunks = _chunks(data) + [data[:n] for n in (0, 1, 5, 31, 32, 33, 100, 255, 256, 300, 4097, 70000)] + [dg.incompressible(FRAME)]
    comp.Level = 1
    plain = comp.WrapBatch(chunks)
    comp.SetParameter(ZSTD_cParameter.ZSTD_c_checksumFlag, 1)
    try:
        frames = comp.WrapBatch(chunks)
    finally:
        comp.SetParameter(ZSTD_cParameter.ZSTD_c_checksumFlag, 0)
    for c, f, q in zip(chunks, frames, plain):
        assert len(f) == len(q) + 4
        want = o.compress(c, 1, checksum=1)
        assert f == want, _first_diff(f, want)
        assert z.decompress(f, max(c.size, 1)) == c.tobytes()
    assert dec.UnwrapBatch(frames) == [c.tobytes() for c in chunks]


@pytest.mark.parametrize("level", [1, 3])
def test_multiblock_frames_byte_identical(comp, dec, level):
    """an input above 128 KiB becomes ONE multi-block frame with the reference's bytes
    (ZSTD_compress_frameChunk, ZstdCompress.cs:4690), checked against the oracle and decoded by libzstd and by the GPU."""
    o, z = oracle(), libzstd()
    cases = multiblock_inputs()
    names = list(cases)
    comp.Level = level
    frames = comp.WrapBatch([cases[k] for k in names])
    for k, f in zip(names, frames):
        want = o.compress(cases[k], level)
        assert f == want, f"{k} level {level}: " + _first_diff(f, want)
        assert z.decompress(f, cases[k].size) == cases[k].tobytes(), k
    assert dec.UnwrapBatch(frames) == [cases[k].tobytes() for k in names]
    # mixed pass: single-block and multi-block frames side by side, with the checksum trailer
    comp.SetParameter(201, 1)
    try:
        mix = [cases["text_200k"], cases["text_200k"][:FRAME], cases["zeros_300k"], cases["text_200k"][:100], cases["text_600k"]]
        frames = comp.WrapBatch(mix)
    finally:
        comp.SetParameter(201, 0)
    for c, f in zip(mix, frames):
        want = o.compress(c, level, checksum=1)
        assert f == want, _first_diff(f, want)
    assert dec.UnwrapBatch(frames) == [c.tobytes() for c in mix]


def test_wrap_larger_than_one_block(comp, dec):
    """BASELINE.json configs[0]: 10 MB of text-like data through Compressor.Wrap / Decompressor.Unwrap: one multi-block frame,
    byte-identical to the oracle's (the reference's) frame."""
    o, z = oracle(), libzstd()
    data = dg.text_like(80 * FRAME)[: 10 * 1000 * 1000]
    comp.Level = 1
    blob = comp.Wrap(data)
    want = o.compress(data, 1)
    assert blob == want, _first_diff(blob, want)
    assert dec.GetDecompressedSize(blob) == data.size
    assert z.decompress(blob, data.size) == data.tobytes()
    assert dec.Unwrap(blob) == data.tobytes()
    # capacity contract: compressBound(srcSize) is always enough, 1 byte less than needed is dstSize_tooSmall
    small = np.empty(len(blob) - 1, dtype=np.uint8)
    assert comp.TryWrap(data, small) == (False, 0)


def test_independent_chunks_parameter(comp, dec):
    """ZSTDB200_c_independentChunks = 1: inputs above 128 KiB are written as back-to-back independent 128 KiB frames
    (all pieces in parallel): any zstd decoder regenerates the input, ZSTD_decompressBound equals the input size, the first
    piece is byte-identical to Wrap(piece), the ratio stays within a few percent of the single frame."""
    o, z = oracle(), libzstd()
    data = dg.text_like(80 * FRAME)[: 10 * 1000 * 1000]
    comp.Level = 1
    comp.SetParameter(10001, 1)
    try:
        blob = comp.Wrap(data)
    finally:
        comp.SetParameter(10001, 0)
    assert dec.GetDecompressedSize(blob) == data.size
    assert z.decompress(blob, data.size) == data.tobytes()
    assert dec.Unwrap(blob) == data.tobytes()
    first = o.compress(data[:FRAME], 1)
    assert blob[:len(first)] == first
    single = o.compress(data, 1)                       # the reference's one multi-block frame (window spans the blocks)
    assert len(blob) < 1.06 * len(single)              # independent pieces cost a few percent of ratio, not more


def test_golden_vectors_on_gpu(comp):
    """The committed golden vectors (tests/golden/golden_vectors.json, written by the reference's own libzstd.dll 1.5.1 through oracle/ref_pe;
    inputs regenerated by make_golden.inputs()) against the GPU compressor directly, without the oracle in between."""
    import hashlib, importlib.util, json, os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    inputs = dict(mod.inputs())
    gold = json.load(open(os.path.join(here, "golden", "golden_vectors.json")))["vectors"]
    for level in (1, 2, 3):
        comp.Level = level
        vs = [v for v in gold if v["level"] == level]
        pieces, owner = [], []
        for v in vs:
            data = np.frombuffer(inputs[v["name"]], dtype=np.uint8)
            whole = data.size <= FRAME or "multiblock" in v["name"]
            ps = [data] if whole else _chunks(data)
            pieces += ps; owner += [v["name"]] * len(ps)
        frames = comp.WrapBatch(pieces)
        for v in vs:
            mine = [f for f, o in zip(frames, owner) if o == v["name"]]
            assert [len(f) for f in mine] == v["frame_sizes"], (v["name"], level)
            assert hashlib.sha256(b"".join(mine)).hexdigest() == v["frames_sha256"], (v["name"], level)
    comp.Level = 1


def test_pipelined_host_path_byte_identical(comp, dec):
    """More than 4096 host-adjacent chunks take the pipelined host path of ZSTDB200_compressBatch (sub-batches in flight on
    several streams, overlapping match / entropy kernels, per-sub-batch scatter): same bytes as one Wrap per chunk."""
    o = oracle()
    rng = np.random.default_rng(0xC0FFEE)
    data = dg.WORKLOADS["silesia"](20 * FRAME)
    cuts = np.sort(rng.choice(np.arange(1, data.size), size=4700, replace=False))
    bounds = np.concatenate([[0], cuts, [data.size]])
    chunks = [data[int(a):int(b)] for a, b in zip(bounds[:-1], bounds[1:])]
    assert len(chunks) > 4096
    comp.Level = 1
    frames = comp.WrapBatch(chunks)
    for c, f in zip(chunks, frames):
        want = o.compress(c, 1)
        assert f == want, _first_diff(f, want)
    assert dec.UnwrapBatch(frames) == [c.tobytes() for c in chunks]


def test_more_than_32768_contiguous_small_items():
    """one contiguous host run with more pieces than 4 pipelined sub-batches of 8192 can take (40000 x 4 KiB pages)
    used to fail the whole call with ZSTD_error_GENERIC.  Every page must come back as the oracle's frame."""
    import ctypes
    from zstdsharp_b200 import Compressor, _native
    lib = _native.lib
    o = oracle()
    n, page = 40000, 4096
    data = np.tile(dg.text_like(64 * FRAME), 20)[:n * page].copy()
    bound = Compressor.GetCompressBound(page)
    out = np.empty(n * bound, dtype=np.uint8)
    vp, st = ctypes.c_void_p, ctypes.c_size_t
    sp = (vp * n)(*[data.ctypes.data + i * page for i in range(n)]); ss = (st * n)(*([page] * n))
    dp = (vp * n)(*[out.ctypes.data + i * bound for i in range(n)]); dc = (st * n)(*([bound] * n)); res = (st * n)()
    with Compressor(1) as c:
        assert lib.ZSTDB200_compressBatch(c.handle, n, 1, sp, ss, dp, dc, res) == 0, lib.ZSTDB200_lastErrorString()
    for i in list(range(0, n, 997)) + [n - 1, 32767, 32768, 32769]:
        assert out[i * bound:i * bound + res[i]].tobytes() == o.compress(data[i * page:(i + 1) * page], 1), i


def test_compressCCtx_ignores_context_parameters():
    """ZSTD_compressCCtx derives everything from the level argument (ZstdCompress.cs:5772: contentSize 1, checksum 0), whatever
    ZSTD_CCtx_setParameter stored; ZSTD_compress2 honours the stored parameters."""
    from zstdsharp_b200 import Compressor, ZSTD_cParameter, _native
    lib = _native.lib
    o = oracle()
    src = dg.text_like(FRAME)
    cap = Compressor.GetCompressBound(FRAME)
    out = np.empty(cap, dtype=np.uint8)
    with Compressor(3) as c:
        c.SetParameter(ZSTD_cParameter.ZSTD_c_checksumFlag, 1)
        r = lib.ZSTD_compressCCtx(c.handle, out.ctypes.data, cap, src.ctypes.data, src.size, 1)
        assert out[:r].tobytes() == o.compress(src, 1, checksum=0)
        r = lib.ZSTD_compress2(c.handle, out.ctypes.data, cap, src.ctypes.data, src.size)
        assert out[:r].tobytes() == o.compress(src, 3, checksum=1)


@pytest.mark.parametrize("level", [-5, -4, -3, -2, -1, 4])
def test_negative_levels_and_level4_byte_identical(level):
    """Levels -5..-1 are ZSTD_fast with stepSize = targetLength + 1 and uncompressed literals (ZstdFast.cs:101, :334;
    ZstdCompress.cs:7918-7923; the reference tests them: ZstdTest.cs:64-67); level 4 is ZSTD_dfast {17,17,17,2,4,0} for
    16 KiB < n <= 128 KiB and {21,18,18,1,5,0} above 256 KiB (Clevels.cs row 4), ZSTD_greedy elsewhere -> parameter_unsupported."""
    from zstdsharp_b200 import Compressor, ZstdException
    o = oracle()
    inputs = []
    for wl in ("text", "silesia", "literal_mix", "incompressible"):
        data = dg.WORKLOADS[wl](3 * FRAME)
        inputs += [data[i * FRAME:(i + 1) * FRAME] for i in range(3)]
    text = dg.text_like(4 * FRAME)
    for n in (0, 1, 7, 8, 63, 64, 100, 1000, 4096, 16384, 16385, 20000, 65536, 100000, 131071, FRAME + 1, 200000, 300000, 3 * FRAME + 5):
        inputs.append(text[:n])
    ok = [a for a in inputs if not o.lib.zo_isError(o.compress_raw(a, level)[0])]
    assert len(ok) >= (12 if level == 4 else len(inputs))
    with Compressor(level) as c:
        frames = c.WrapBatch(ok)
        for a, f in zip(ok, frames):
            assert f == o.compress(a, level), (level, a.size)
        if level == 4:
            with pytest.raises(ZstdException) as e:
                c.Wrap(text[:1000])
            assert int(e.value.Code) == 40


@pytest.mark.parametrize("level", [1, 2, 3, -3])
def test_dictionary_compression_byte_identical(dec, level):
    """SURVEY 8f.4, encode side: Compressor.LoadDictionary + Wrap / WrapBatch (Compressor.cs:43-56 -> ZSTD_CCtx_loadDictionary, the CDict
    built by the first compression at the context's level, attached below 8 / 16 KiB and copied above, dictionary entropy tables
    and repcodes, dictionary id in the frame header).  Frames equal the oracle's (pinned to the reference DLL's
    ZSTD_CCtx_loadDictionary + ZSTD_compress2 bytes: tests/test_reference_pin.py) for zstd-format and raw dictionaries of 7 bytes
    .. 300 KB and inputs of 0 bytes .. 600 KB around every cut-off; they decode with the same dictionary on the GPU; a fresh
    context is used per (dictionary, level) like the reference's tests (ZstdNetTests.cs:19-39)."""
    from zstdsharp_b200 import Compressor
    from _dict_cases import compress_dictionaries, compress_payloads
    o = oracle()
    dicts = compress_dictionaries(libzstd())
    pays = compress_payloads()
    for name, d in dicts.items():
        c = Compressor(level)
        try:
            c.LoadDictionary(d)
            frames = c.WrapBatch(pays)
            assert c.launch_count() > 0
            for src, f in zip(pays, frames):
                want = o.compress_loaded_dict(src, level, d)
                assert f == want, (name, level, src.size, _first_diff(f, want))
            one = c.Wrap(pays[3])                                   # single-call path (ZSTD_compress2)
            assert bytes(one) == o.compress_loaded_dict(pays[3], level, d)
            c.LoadDictionary(None)                                  # back to no dictionary (ZSTD_CCtx_loadDictionary(NULL, 0))
            assert bytes(c.Wrap(pays[3])) == o.compress(pays[3], level)
        finally:
            c.Dispose()
        dec.LoadDictionary(d)
        try:
            assert dec.UnwrapBatch(frames) == [p.tobytes() for p in pays], name
        finally:
            dec.LoadDictionary(None)


def test_dictionary_tests_of_the_reference(dec):
    """ZstdNetTests.cs:19-39 (CompressAndDecompress_workCorrectly with a dictionary at the minimum / default / maximum level),
    :95-134 (decoding without the dictionary, or with another one, throws), plus the reference's verdict on a dictionary it
    cannot digest: LoadDictionary succeeds, every Wrap reports memory_allocation (ZstdCompress.cs:1604-1607; pinned on the DLL
    in tests/test_reference_pin.py)."""
    from zstdsharp_b200 import Compressor, Decompressor, ZstdException
    from _dict_cases import dictionaries
    dicts = dictionaries(libzstd())
    d = dicts["zdict_32k"]
    data = dg.text_like(3 * FRAME)[FRAME // 2:FRAME // 2 + 90_000]
    for level in (Compressor.MinCompressionLevel, Compressor.DefaultCompressionLevel, 3):
        c = Compressor(level); c.LoadDictionary(d)
        f = bytes(c.Wrap(data)); c.Dispose()
        dec.LoadDictionary(d)
        assert bytes(dec.Unwrap(f)) == data.tobytes()
        dec.LoadDictionary(None)
        with pytest.raises(ZstdException):                       # DecompressWithoutDictionary_throwsZstdException_onDataCompressedWithIt
            dec.Unwrap(f)
        dec.LoadDictionary(b"zstd supports raw-content dictionaries")
        with pytest.raises(ZstdException):                       # DecompressWithAnotherDictionary_throwsZstdException
            dec.Unwrap(f)
        dec.LoadDictionary(None)
    bad = bytearray(dicts["zdict_4k"]); bad[9] ^= 0xFF; bad[10] ^= 0x55; bad[12] ^= 0xFF
    c = Compressor(1); c.LoadDictionary(bytes(bad))
    with pytest.raises(ZstdException) as e:
        c.Wrap(data)
    assert int(e.value.Code) == 64
    c.LoadDictionary(dicts["zdict_4k"])                          # a good dictionary afterwards works
    assert bytes(c.Wrap(data)) == oracle().compress_loaded_dict(data, 1, dicts["zdict_4k"])
    c.Dispose()


def test_dictionary_digest_keeps_its_level_like_the_reference():
    """ZSTD_initLocalDict (ZstdCompress.cs:1581) builds the CDict once, with the level of the first compression after LoadDictionary; a later
    level change only moves the window (ZSTD_CCtx_init_compressStream2 :6949), the match finder keeps the CDict's parameters.  The same call
    sequence on one context of the reference DLL and on one GPU context gives the same bytes."""
    from _oracle import refdll, refdll_available
    if not refdll_available():
        pytest.skip("oracle/_ref not built")
    from zstdsharp_b200 import Compressor
    from _dict_cases import dictionaries
    r = refdll(); L = r.lib
    d = dictionaries(libzstd())["zdict_32k"]
    text = dg.text_like(4 * FRAME)
    inputs = [np.ascontiguousarray(text[1000:1000 + 6000]), np.ascontiguousarray(text[50_000:50_000 + 90_000]), np.ascontiguousarray(text[200_000:200_000 + 12_000])]
    for first, second in ((1, 3), (3, 1), (2, -3)):
        c = Compressor(first); c.LoadDictionary(d)
        rc = L.ZREF_createCCtx(); L.ZREF_CCtx_setParameter(rc, 100, first)
        db = np.frombuffer(d, dtype=np.uint8)
        assert not L.ZREF_isError(L.ZREF_CCtx_loadDictionary(rc, db.ctypes.data, db.size))
        out = np.empty(200_000, dtype=np.uint8)
        try:
            for k, level in enumerate((first, second, second, first)):
                src = inputs[k % len(inputs)]
                c.Level = level
                L.ZREF_CCtx_setParameter(rc, 100, level)
                n = L.ZREF_compress2(rc, out.ctypes.data, out.size, src.ctypes.data, src.size)
                assert not L.ZREF_isError(n)
                assert bytes(c.Wrap(src)) == out[:n].tobytes(), (first, second, k, level, src.size)
        finally:
            L.ZREF_freeCCtx(rc); c.Dispose()


def test_multi_device_codec_with_a_dictionary():
    """ZSTDB200_multiLoadDictionary loads the dictionary into every device's compression and decompression context: one WrapBatch /
    UnwrapBatch over all visible devices gives the oracle's frames in the caller's order and the inputs back."""
    from zstdsharp_b200 import MultiCodec
    from _dict_cases import dictionaries, compress_payloads
    o = oracle()
    d = dictionaries(libzstd())["zdict_4k"]
    pays = compress_payloads(n_random=8)
    m = MultiCodec(0, 3)
    try:
        m.LoadDictionary(d)
        frames = m.WrapBatch(pays)
        for src, f in zip(pays, frames):
            assert f == o.compress_loaded_dict(src, 3, d), src.size
        assert m.UnwrapBatch(frames) == [p.tobytes() for p in pays]
        m.LoadDictionary(None)
        assert m.WrapBatch(pays[:4]) == [o.compress(p, 3) for p in pays[:4]]
    finally:
        m.Dispose()


def test_dfast_write_of_the_last_visited_position(comp):
    """tests/_cases.py::dfast_last_window_case (found by the soak, seed 993001): the long-table write of the last visited position of a block
    must survive an unvisited position with the same 8 bytes in the same window."""
    from _cases import dfast_last_window_case
    o = oracle()
    a = dfast_last_window_case()
    for level in (1, 2, 3):
        comp.Level = level
        assert bytes(comp.Wrap(a)) == o.compress(a, level), level

