"""Generates tests/golden/golden_vectors.json.

The reference (C#) cannot run in this image and its own fixtures are unusable here (`dickens` is missing, libzstd.dll
is a Windows binary: SURVEY.md section 8c), so the golden vectors are produced by the upstream C implementation the
reference is a mechanical translation of -- system libzstd (1.5.5) -- on small deterministic inputs, at levels 1..3
where 1.5.5 was verified byte-identical to the reference's 1.5.1 logic on this path.  For larger inputs only the
SHA-256 of the frame is stored.  Run:  python tests/golden/make_golden.py
"""
import ctypes
import ctypes.util
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from zstdsharp_b200 import datagen as dg  # noqa: E402


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
    z = ctypes.CDLL(ctypes.util.find_library("zstd") or "libzstd.so.1")
    z.ZSTD_compress.restype = ctypes.c_size_t
    z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    z.ZSTD_compressBound.restype = ctypes.c_size_t
    z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    z.ZSTD_versionNumber.restype = ctypes.c_uint
    out = {"generator": "system libzstd", "libzstd_version": int(z.ZSTD_versionNumber()), "vectors": []}
    for name, data in inputs():
        whole = len(data) <= dg.FRAME or "multiblock" in name
        pieces = [data] if whole else [data[i:i + dg.FRAME] for i in range(0, len(data), dg.FRAME)]
        for level in (1, 2, 3):
            frames = []
            for p in pieces:
                cap = z.ZSTD_compressBound(len(p))
                buf = ctypes.create_string_buffer(max(cap, 1))
                r = z.ZSTD_compress(buf, cap, p, len(p), level)
                frames.append(buf.raw[:r])
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
