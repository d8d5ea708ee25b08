"""Stream adapters over the batch API (reference: CompressionStream.cs / DecompressionStream.cs and
ZstdNetSteamingTests.cs:32-45, 237-267 for the round-trip shape)."""
import io

import numpy as np
import pytest

from zstdsharp_b200 import datagen as dg

from _oracle import oracle, libzstd

FRAME = dg.FRAME


def test_split_frames_host_walk():
    """ZSTD_findFrameCompressedSize drives the cut of a concatenated stream (no GPU involved)."""
    from zstdsharp_b200.streams import split_frames
    from zstdsharp_b200 import ZstdException
    o = oracle()
    data = dg.text_like(3 * FRAME)
    frames = [o.compress(data[i * FRAME:(i + 1) * FRAME], 1, checksum=i & 1) for i in range(3)]
    frames.append(o.compress(data[:2 * FRAME + 5], 3))                     # multi-block frame
    skippable = (0x184D2A55).to_bytes(4, "little") + (3).to_bytes(4, "little") + b"abc"
    blob = frames[0] + skippable + b"".join(frames[1:])
    got, used = split_frames(blob)
    assert used == len(blob) and got == [frames[0], skippable] + frames[1:]
    got, used = split_frames(blob[:-7])                                     # truncated tail stays unconsumed
    assert got == [frames[0], skippable] + frames[1:3] and used == len(blob) - len(frames[3])
    with pytest.raises(ZstdException):
        split_frames(frames[0] + b"this is not a zstd frame at all")


@pytest.mark.gpu
def test_stream_round_trip_and_interop():
    from zstdsharp_b200.streams import CompressionStream, DecompressionStream
    z = libzstd()
    data = dg.silesia_mix(9 * FRAME)[: 8 * FRAME + 4321].tobytes()
    sink = io.BytesIO()
    with CompressionStream(sink, level=1, batch_frames=4) as cs:
        for i in range(0, len(data), 100000):                               # odd write sizes on purpose
            cs.Write(data[i:i + 100000])
    blob = sink.getvalue()
    assert len(blob) < len(data)
    assert z.decompress(blob, len(data)) == data                            # one valid zstd stream for any decoder
    with DecompressionStream(io.BytesIO(blob), batch_bytes=300000) as ds:
        out = bytearray()
        while True:
            piece = ds.Read(70001)
            if not piece:
                break
            out += piece
    assert bytes(out) == data
    # frames made elsewhere (libzstd, several levels, with and without checksum) through the same reader
    foreign = b"".join(z.compress(np.frombuffer(data[i:i + FRAME], dtype=np.uint8), lvl, checksum=lvl & 1)
                       for i, lvl in zip(range(0, 4 * FRAME, FRAME), (1, 3, 9, 19)))
    assert DecompressionStream(io.BytesIO(foreign)).readall() == data[:4 * FRAME]


@pytest.mark.gpu
def test_stream_round_trip_with_a_dictionary():
    """CompressionStream.LoadDictionary / DecompressionStream.LoadDictionary (CompressionStream.cs:58-62, DecompressionStream.cs:58-62;
    ZstdNetSteamingTests: the dictionary variants of the round trips): every frame of the stream is the oracle's dictionary frame of its
    piece, the stream reads back with the dictionary and fails without it."""
    from zstdsharp_b200.streams import CompressionStream, DecompressionStream
    from zstdsharp_b200 import ZstdException
    from _dict_cases import dictionaries
    o = oracle()
    d = dictionaries(libzstd())["zdict_32k"]
    data = dg.text_like(4 * FRAME)[FRAME // 3: FRAME // 3 + 2 * FRAME + 777].tobytes()
    sink = io.BytesIO()
    with CompressionStream(sink, level=3, batch_frames=2, leaveOpen=True) as cs:
        cs.LoadDictionary(d)
        cs.Write(data)
    z = sink.getvalue()
    expect = b"".join(o.compress_loaded_dict(np.frombuffer(data[i:i + FRAME], dtype=np.uint8), 3, d) for i in range(0, len(data), FRAME))
    assert z == expect
    ds = DecompressionStream(io.BytesIO(z))
    ds.LoadDictionary(d)
    assert ds.read() == data
    with pytest.raises(ZstdException):
        DecompressionStream(io.BytesIO(z)).read()
