"""Generates tests/golden/dict_golden.json: Compressor.LoadDictionary + Wrap frames written by the REFERENCE'S OWN libzstd.dll
(ZSTD_CCtx_loadDictionary + ZSTD_compress2 through oracle/ref_pe), as SHA-256 + length, for the dictionaries and inputs of
tests/_dict_cases.py.  The trained dictionaries come from system libzstd's ZDICT; their SHA-256 is stored so that a different
trainer cannot silently change the question.  Needs /root/reference (this container); the vectors travel.
Run:  python tests/golden/make_dict_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from _dict_cases import compress_dictionaries, compress_payloads  # noqa: E402
from _oracle import refdll, libzstd  # noqa: E402

LEVELS = (-3, 1, 2, 3)


def main():
    r = refdll()
    assert r.version() == 10501
    dicts = compress_dictionaries(libzstd())
    pays = compress_payloads()
    out = {"generator": "reference src/Zstd.Extern/libzstd.dll through oracle/ref_pe: ZSTD_CCtx_loadDictionary + ZSTD_compress2",
           "dictionaries": {k: hashlib.sha256(v).hexdigest() for k, v in dicts.items()}, "vectors": []}
    for name, d in dicts.items():
        for pi, src in enumerate(pays):
            for level in LEVELS:
                if (pi + LEVELS.index(level)) % 2:                    # half of the grid keeps the file small
                    continue
                f = r.compress_loaded_dict(src, level, d, checksum=pi & 1)
                out["vectors"].append({"dict": name, "payload": pi, "size": int(src.size), "level": level, "checksum": pi & 1,
                                       "frame_len": len(f), "frame_sha256": hashlib.sha256(f).hexdigest()})
    with open(os.path.join(HERE, "dict_golden.json"), "w") as fo:
        json.dump(out, fo, indent=0)
    print(len(out["vectors"]), "vectors")


if __name__ == "__main__":
    main()
