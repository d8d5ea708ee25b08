"""Generates tests/golden/golden_vectors.json from the REFERENCE'S OWN native oracle.

src/Zstd.Extern/libzstd.dll (zstd 1.5.1) is the binary the reference's differential test asserts ZstdSharp equals byte for
byte at every level (src/ZstdSharp.Test/ZstdTest.cs:64-90).  It runs here through the PE mapper in oracle/ref_pe
(oracle/_ref/libzstdref.so), so the vectors are reference-held answers, not the output of some other zstd release.
The generator needs /root/reference (this container); the vectors travel.  For larger inputs only the SHA-256 of the
frame is stored.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from zstdsharp_b200 import datagen as dg  # noqa: E402

LEVELS = (1, 2, 3)


def inputs():
    rng = np.random.default_rng(20240229)
    yield "empty", b""
    yield "one_byte", b"\x2a"
    yield "abc", b"abc"
    yield "ramp_100", dg.byte_ramp(100).tobytes()
    yield "ramp_3002", dg.byte_ramp(3002).tobytes()
    yield "zeros_1000", bytes(1000)
    yield "text_700", dg.text_like(700).tobytes()
    yield "text_5000", dg.text_like(5000).tobytes()
    yield "skew_2000", dg.literal_heavy(2000).tobytes()
    yield "random_300", rng.integers(0, 256, 300, dtype=np.uint8).tobytes()
    yield "text_128k", dg.text_like(dg.FRAME).tobytes()
    yield "silesia_128k_x4", dg.silesia_mix(4 * dg.FRAME).tobytes()
    yield "zeros_128k", bytes(dg.FRAME)
    yield "literal_heavy_128k", dg.literal_heavy(dg.FRAME).tobytes()
    yield "text_1m_multiblock", dg.text_like(8 * dg.FRAME).tobytes()
    # multi-block frames that exercise the block-to-block state (same shapes as tests/_cases.py)
    text = dg.text_like(6 * dg.FRAME); rnd = dg.incompressible(dg.FRAME); zeros = np.zeros(3 * dg.FRAME, dtype=np.uint8)
    yield "multiblock_text_128k+6", text[:dg.FRAME + 6].tobytes()                                   # raw last block of 6 bytes
    yield "multiblock_text_600k", text[:600_000].tobytes()                                          # window smaller than the frame
    yield "multiblock_zeros_384k", zeros.tobytes()                                                  # RLE blocks after the first
    yield "multiblock_text_random_text", np.concatenate([text[:dg.FRAME + 5000], rnd, text[dg.FRAME:3 * dg.FRAME]]).tobytes()   # unconfirmed raw block
    yield "multiblock_literal_heavy_512k", dg.literal_heavy(4 * dg.FRAME).tobytes()                 # Huffman table reuse


def main():
    from _oracle import refdll
    z = refdll()
    assert z.version() == 10501
    out = {"generator": "reference src/Zstd.Extern/libzstd.dll through oracle/ref_pe", "libzstd_version": z.version(), "vectors": []}
    for name, data in inputs():
        whole = len(data) <= dg.FRAME or "multiblock" in name
        pieces = [data] if whole else [data[i:i + dg.FRAME] for i in range(0, len(data), dg.FRAME)]
        for level in LEVELS:
            frames = [z.compress(p, level) for p in pieces]
            blob = b"".join(frames)
            v = {"name": name, "level": level, "src_len": len(data), "src_sha256": hashlib.sha256(data).hexdigest(),
                 "frame_sizes": [len(f) for f in frames], "frames_sha256": hashlib.sha256(blob).hexdigest()}
            if len(blob) <= 1024:
                v["frames_hex"] = blob.hex()
            out["vectors"].append(v)
    with open(os.path.join(HERE, "golden_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["vectors"]), "vectors")


if __name__ == "__main__":
    main()
