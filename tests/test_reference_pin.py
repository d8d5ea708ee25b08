"""Pins the CPU oracle to the REFERENCE'S OWN native oracle -- runs without a GPU.

src/Zstd.Extern/libzstd.dll (zstd 1.5.1, PE32+) is the binary that the reference's differential test requires ZstdSharp to
equal byte for byte at every level (src/ZstdSharp.Test/ZstdTest.cs:64-90: `compressedNative.SequenceEqual(compressedSharp)`,
`decompressedSharp.SequenceEqual(decompressedNative)`, levels -5..-1 and 1..22).  oracle/ref_pe maps it into this process
(oracle/_ref/libzstdref.so), so here the oracle -- the plain-C restatement of the C# -- is held to exactly that contract:
frames byte-identical at the levels the GPU path implements, decoded bytes identical for frames of every level, the same
verdict (error code included) on damaged frames, the same match-finder output (ZSTD_generateSequences as a stage oracle).

On the GPU box /root/reference does not exist; the prebuilt oracle/_ref/libzstdref.so travels with the snapshot.
"""
import numpy as np
import pytest

from zstdsharp_b200 import datagen as dg

from _oracle import oracle, refdll, refdll_available

FRAME = dg.FRAME
pytestmark = pytest.mark.skipif(not refdll_available(), reason="oracle/_ref/libzstdref.so not built (needs /root/reference)")

ALL_LEVELS = list(range(-5, 0)) + list(range(1, 23))          # ZstdTest.cs:64-67 LevelsData


def test_the_dll_is_zstd_1_5_1_and_loads_without_windows():
    r = refdll()
    assert r.version() == 10501                                # README.md: "Based on Zstandard v1.5.1"
    assert r.lib.ZREF_compressBound(FRAME) == 131584 == oracle().lib.zo_compressBound(FRAME)    # ZstdTest.cs:75-76
    for n in (0, 1, 255, 65536, 10_000_000):
        assert r.lib.ZREF_compressBound(n) == oracle().lib.zo_compressBound(n)


def test_cparams_match_the_dll():
    """ZSTD_getCParams of the DLL against the oracle's restatement of Clevels.cs + ZSTD_adjustCParams_internal."""
    o, r = oracle(), refdll()
    for level in (0, 1, 2, 3):
        for n in (1, 100, 4096, 16384, 16385, 65536, 131071, FRAME, FRAME + 1, 262144, 262145, 600_000, 10_000_000):
            assert o.cparams(level, n) == r.cparams(level, n), (level, n)


@pytest.mark.parametrize("workload", ["text", "silesia", "incompressible", "literal_heavy", "literal_mix"])
def test_encoder_byte_identical_to_the_dll(workload):
    """configs[1]/[2] shape: 128 KiB chunks, one frame each, levels 1..3."""
    o, r = oracle(), refdll()
    data = dg.WORKLOADS[workload](8 * FRAME)
    for i in range(0, data.size, FRAME):
        c = data[i:i + FRAME]
        for level in (1, 2, 3):
            assert o.compress(c, level) == r.compress(c, level), (workload, i, level)
        assert o.compress(c, 1, checksum=1) == r.compress(c, 1, checksum=1)


def test_encoder_size_sweep_matches_the_dll():
    """Every parameter-table bucket and the small-input corner cases (ZstdNetTests.cs:456-496 sizes, `(byte)i` data and text)."""
    o, r = oracle(), refdll()
    text = dg.text_like(2 * FRAME)
    sizes = [0, 1, 2, 6, 7, 8, 63, 64, 255, 256, 257, 1023, 1024, 4096, 16383, 16384, 16385, 40959, 40960, 65535, 65536,
             65791, 65792, 100001, 131071, 131072] + list(range(2, 100000, 9000))
    for n in sizes:
        for src in (dg.byte_ramp(n), text[:n]):
            for level in (1, 3):
                assert o.compress(src, level) == r.compress(src, level), (n, level)


def test_multiblock_frames_match_the_dll():
    """configs[0] (10 MB text, one frame) and the block-to-block shapes of tests/_cases.py."""
    from _cases import multiblock_inputs
    o, r = oracle(), refdll()
    data = dg.text_like(80 * FRAME)[: 10 * 1000 * 1000]
    for level in (1, 3):
        f = o.compress(data, level)
        assert f == r.compress(data, level)
        assert r.decompress(f, data.size) == data.tobytes() == o.decompress(f, data.size)
    for name, d in multiblock_inputs().items():
        for level in (1, 3):
            assert o.compress(d, level) == r.compress(d, level), (name, level)


def test_decoder_matches_the_dll_on_all_27_levels():
    """ZstdTest.cs:69-90 with the oracle in ZstdSharp's seat: frames written by the DLL at every level decode to the same bytes."""
    o, r = oracle(), refdll()
    for data in (dg.text_like(5 * FRAME)[:600_000], dg.silesia_mix(3 * FRAME)):
        for level in ALL_LEVELS:
            f = r.compress(data, level, checksum=level & 1)
            assert o.decompress(f, data.size) == data.tobytes() == r.decompress(f, data.size), level
            assert o.decompress_bound(f) == r.decompress_bound(f) == data.size


def _resolve(off_codes, lls, mls):
    """oracle seqStore entries (offset = offCode + 1: 1..3 repcodes, else offset + 3; ml = matchLength - 3) -> raw offsets,
    with the repcode rules of ZSTD_updateRep / ZSTD_copyBlockSequences (ZstdCompress.cs:3606-3660)."""
    rep = [1, 4, 8]
    out = []
    for oc, ll, ml in zip(off_codes.tolist(), lls.tolist(), mls.tolist()):
        if oc > 3:
            raw = oc - 3
            rep = [raw, rep[0], rep[1]]
        else:
            idx = oc - 1 + (1 if ll == 0 else 0)
            if idx == 0:
                raw = rep[0]
            else:
                raw = rep[0] - 1 if idx == 3 else rep[idx]
                rep = [raw, rep[0], rep[1]] if idx != 1 else [raw, rep[0], rep[2]]
        out.append((raw, ll, ml + 3))
    return out


@pytest.mark.parametrize("level", [1, 3])
def test_match_finder_stage_equals_generateSequences(level):
    """ZSTD_generateSequences of the DLL = the match finder's (offset, litLength, matchLength) list of a block: a stage oracle for
    ZSTD_compressBlock_fast / _doubleFast (ZstdFast.cs:96, ZstdDoubleFast.cs:51) independent of the entropy stage."""
    o, r = oracle(), refdll()
    for wl in ("text", "silesia", "literal_mix"):
        data = dg.WORKLOADS[wl](2 * FRAME)
        for i in range(0, data.size, FRAME):
            c = data[i:i + FRAME]
            off, ll, ml, lits, longLen, rep = o.matchfinder(c, level)
            ll = ll.copy(); ml = ml.copy()
            if longLen[0] == 1:
                ll[longLen[1]] += 0x10000
            elif longLen[0] == 2:
                ml[longLen[1]] += 0x10000
            mine = _resolve(off, ll, ml)
            seqs = r.generate_sequences(c, level)
            theirs = [(int(s[0]), int(s[1]), int(s[2])) for s in seqs if s[2] != 0]            # drop the block delimiter
            assert mine == theirs, (wl, i, level)
            delim = [s for s in seqs if s[2] == 0]
            assert len(delim) == 1 and int(delim[0][1]) == c.size - sum(a + b for _, a, b in mine)   # last literals


def test_damaged_frames_get_the_dlls_verdict():
    """Bit flips anywhere in a frame: same error code, or same size and bytes.  This is what pins the double-symbol Huffman
    decoder's acceptance rules (HufDecompress.cs:1022-1045, :1322-1335) and the bit reader's behaviour past a stream's start."""
    o, r = oracle(), refdll()
    n = 0
    for wl, seed in (("text", 1), ("silesia", 2), ("literal_heavy", 3), ("literal_mix", 4)):
        data = dg.WORKLOADS[wl](2 * FRAME)
        rng = np.random.default_rng(seed)
        for ci in range(2):
            src = data[ci * FRAME:(ci + 1) * FRAME]
            for level in (1, 3, 19):
                f = r.compress(src, level, checksum=ci)
                for _ in range(120):
                    g = bytearray(f)
                    for _k in range(int(rng.integers(1, 3))):
                        pos = int(rng.integers(0, len(g)))
                        g[pos] ^= 1 << int(rng.integers(0, 8))
                    ro, outo = o.decompress_raw(bytes(g), FRAME)
                    rr, outr = r.decompress_raw(bytes(g), FRAME)
                    assert o.error_code(ro) == r.error_code(rr), (wl, level, ci)
                    if not r.lib.ZREF_isError(rr):
                        assert ro == rr and outo[:ro].tobytes() == outr[:rr].tobytes()
                    n += 1
    assert n == 4 * 2 * 3 * 120


def test_known_answers_hold_for_the_dll():
    """The answers tests/test_oracle_pin.py takes from the reference's test suite, read off the DLL itself."""
    o, r = oracle(), refdll()
    src = dg.text_like(FRAME)
    f = r.compress(src, 1)
    assert f[:9] == bytes.fromhex("28b52ffda000000200")
    assert r.compress(bytes(300), 1)[4] == 0x60                                           # ZstdNetTests.cs:194-204
    assert r.compress(b"", 1) == bytes.fromhex("28b52ffd2000010000") == o.compress(b"", 1)
    assert r.error_code(r.compress_raw(src, 1, cap=20)[0]) == 70 == o.error_code(o.compress_raw(src, 1, cap=20)[0])
    assert r.error_code(r.decompress_raw(f, 20)[0]) == 70
    assert r.error_code(r.decompress_raw(bytes(range(1, 100)), 1000)[0]) == 10
    assert r.error_code(r.decompress_raw(f[:100], FRAME)[0]) == 72
    fc = r.compress(src, 1, checksum=1)
    assert len(fc) == len(f) + 4
    bad = bytearray(fc); bad[-1] ^= 0xFF
    assert r.error_code(r.decompress_raw(bytes(bad), FRAME)[0]) == 22 == o.error_code(o.decompress_raw(bytes(bad), FRAME)[0])


def test_dictionary_frames_of_the_dll_decode_in_the_oracle():
    """SURVEY 8f.4, decode side, pinned by the reference's binary: ZSTD_compress_usingDict frames (zstd-format and raw-content
    dictionaries) decoded by the oracle and by the DLL give the same bytes; wrong dictionary -> dictionary_wrong (32)."""
    from _dict_cases import dictionaries, payloads
    from _oracle import libzstd
    o, r = oracle(), refdll()
    dicts = dictionaries(libzstd())
    for name, d in dicts.items():
        for level in (1, 3, 9):
            for src in payloads():
                f = r.compress_using_dict(src, level, d)
                assert o.decompress_using_dict(f, src.size, d) == src.tobytes(), (name, level, src.size)
                rr, outr = r.decompress_using_dict_raw(f, src.size, d)
                assert rr == src.size and outr[:rr].tobytes() == src.tobytes()
    f = r.compress_using_dict(payloads()[3], 3, dicts["zdict_32k"])
    ro, _ = o.decompress_using_dict_raw(f, 70000, dicts["zdict_4k"])
    rr, _ = r.decompress_using_dict_raw(f, 70000, dicts["zdict_4k"])
    assert o.error_code(ro) == r.error_code(rr) == 32


def test_dictionary_compression_matches_the_dll():
    """SURVEY 8f.4, encode side: Compressor.LoadDictionary + Wrap (ZSTD_CCtx_loadDictionary, then ZSTD_compress2 digests the bytes
    into a CDict at the context's level: ZstdCompress.cs:1581, :5933, :5826; attach below 8 / 16 KiB, copy above: :2738-2881;
    match finders ZstdFast.cs:390 / :583, ZstdDoubleFast.cs:250 / :590; entropy tables and repcodes of the dictionary with
    their repeat modes: :5264).  The oracle writes the DLL's bytes for zstd-format and raw dictionaries of 7 bytes .. 300 KB,
    inputs of 0 bytes .. 600 KB around every cut-off, levels -3, 1, 2, 3, with and without checksum; the DLL decodes them."""
    from _dict_cases import compress_dictionaries, compress_payloads
    from _oracle import libzstd
    o, r = oracle(), refdll()
    dicts = compress_dictionaries(libzstd())
    pays = compress_payloads()
    n = 0
    for di, (name, d) in enumerate(dicts.items()):
        for pi, src in enumerate(pays):
            level = (1, 2, 3, -3)[(di + pi) % 4]
            want = r.compress_loaded_dict(src, level, d, checksum=(pi & 1))
            assert o.compress_loaded_dict(src, level, d, checksum=(pi & 1)) == want, (name, level, src.size)
            if pi % 5 == 0:
                rr, outr = r.decompress_using_dict_raw(want, src.size, d)
                assert rr == src.size and outr[:rr].tobytes() == src.tobytes()
            n += 1
    # every level on one dictionary / payload each side of the cut-offs
    for level in (-5, -1, 1, 2, 3):
        for src in (pays[3], pays[8], pays[10]):
            assert o.compress_loaded_dict(src, level, dicts["zdict_32k"]) == r.compress_loaded_dict(src, level, dicts["zdict_32k"]), (level, src.size)
    assert n >= 300
    # a dictionary whose entropy header is damaged: loading succeeds, every compression reports memory_allocation (64),
    # because ZSTD_createCDict_advanced2 returns NULL (ZstdCompress.cs:1604-1607)
    bad = bytearray(dicts["zdict_4k"]); bad[9] ^= 0xFF; bad[10] ^= 0x55; bad[12] ^= 0xFF
    L = r.lib
    c = L.ZREF_createCCtx(); L.ZREF_CCtx_setParameter(c, 100, 1)
    db = np.frombuffer(bytes(bad), dtype=np.uint8); out = np.empty(70000, dtype=np.uint8)
    assert not L.ZREF_isError(L.ZREF_CCtx_loadDictionary(c, db.ctypes.data, db.size))
    rv = L.ZREF_compress2(c, out.ctypes.data, out.size, pays[3].ctypes.data, pays[3].size)
    L.ZREF_freeCCtx(c)
    ro, _ = o.compress_loaded_dict_raw(pays[3], 1, bytes(bad))
    assert r.error_code(rv) == o.error_code(ro) == 64


def test_constructed_encoder_cases_match_the_dll():
    """tests/_cases.py::dfast_last_window_case (the input behind the round-2 fix of enc_match_dfast_group_kernel): the oracle writes the
    DLL's bytes for it at every restated level, so the GPU test that uses it is held to the reference."""
    from _cases import dfast_last_window_case
    o, r = oracle(), refdll()
    a = dfast_last_window_case()
    for level in (-5, 1, 2, 3):                                    # level 4 is ZSTD_greedy for this size: not restated
        assert o.compress(a, level) == r.compress(a, level), level


def test_handbuilt_tiny_four_stream_literals():
    """tests/_cases.py::handbuilt_small_4stream_frames: legal frames no zstd encoder writes (four Huffman streams over < 256
    literals), decoded by the oracle and by the DLL to the bytes the construction implies."""
    from _cases import handbuilt_small_4stream_frames
    o, r = oracle(), refdll()
    for f, expect in handbuilt_small_4stream_frames():
        assert o.decompress(f, len(expect)) == expect == r.decompress(f, len(expect))


@pytest.mark.parametrize("level", [-5, -4, -3, -2, -1, -17, 4])
def test_negative_levels_and_level4_match_the_dll(level):
    """ZstdTest.cs:64-67 runs levels -5..-1 too; level 4 is restated only for the sizes where it is still ZSTD_dfast."""
    from _cases import multiblock_inputs
    o, r = oracle(), refdll()
    n = 0
    for wl in ("text", "silesia", "literal_mix", "incompressible"):
        data = dg.WORKLOADS[wl](3 * FRAME)
        for i in range(0, data.size, FRAME):
            assert o.compress(data[i:i + FRAME], level) == r.compress(data[i:i + FRAME], level), (wl, i)
            n += 1
    text = dg.text_like(4 * FRAME)
    for size in (0, 1, 7, 8, 63, 64, 100, 1000, 4096, 16384, 16385, 20000, 65536, 100000, 131071, FRAME + 1, 200000, 300000, 3 * FRAME + 5):
        rv, _ = o.compress_raw(text[:size], level)
        if o.lib.zo_isError(rv):
            assert level == 4 and o.error_code(rv) == 40 and (size == 0 or r.cparams(4, size)[6] > 2)      # ZSTD_greedy: outside the restated scope
            continue
        if size:                                                   # ZSTD_getCParams reads a size hint of 0 as 'unknown'
            assert o.cparams(level, size) == r.cparams(level, size)
        assert o.compress(text[:size], level) == r.compress(text[:size], level), size
        n += 1
    for name, d in multiblock_inputs().items():
        if not o.lib.zo_isError(o.compress_raw(d, level)[0]):
            assert o.compress(d, level) == r.compress(d, level), name
            n += 1
    assert n >= 20


def test_error_codes_found_by_the_round2_soak():
    """tests/golden/soak_r02_frames.json against the DLL itself (the fixture's codes were read off it) and the oracle."""
    import json, os
    o, r = oracle(), refdll()
    for c in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "soak_r02_frames.json")))["cases"]:
        f = bytes.fromhex(c["frame_hex"])
        assert r.error_code(r.decompress_raw(f, c["capacity"])[0]) == c["error_code"] == o.error_code(o.decompress_raw(f, c["capacity"])[0]), c["name"]
