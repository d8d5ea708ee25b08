"""Developer probe: throughput of the dictionary compression path (serial match kernel, one thread per frame) next to the no-dictionary
path on the same inputs.  Usage: bench_dict.py   (needs a GPU; results go to stdout as JSON lines)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from zstdsharp_b200 import Compressor, Decompressor, datagen as dg
from _oracle import libzstd

text = dg.text_like(1 << 30)
samples = [text[i * 4000:(i + 1) * 4000].tobytes() for i in range(600)]
d32 = libzstd().train_dictionary(samples, 32768)
for rec, n in ((131072, 4096), (4096, 65536), (1024, 65536)):
    data = text[64 << 20:(64 << 20) + rec * n]
    chunks = [data[i * rec:(i + 1) * rec] for i in range(n)]
    for level in (1, 3):
        for use_dict in (False, True):
            c = Compressor(level)
            if use_dict: c.LoadDictionary(d32)
            frames = c.WrapBatch(chunks)            # warm-up (digest, arenas)
            t0 = time.perf_counter(); frames = c.WrapBatch(chunks); dt = time.perf_counter() - t0
            t = c.timings()
            out = sum(len(f) for f in frames)
            print(json.dumps({"record_bytes": rec, "records": n, "level": level, "dictionary": use_dict, "ratio": round(rec * n / out, 3),
                              "wall_gbs": round(rec * n / dt / 1e9, 3), "match_ms": round(t[8], 2), "entropy_ms": round(t[9], 2), "kernels_ms": round(t[1], 2)}), flush=True)
            if use_dict:
                dd = Decompressor(); dd.LoadDictionary(d32)
                back = dd.UnwrapBatch(frames)
                assert all(b == ch.tobytes() for b, ch in zip(back, chunks))
                dd.Dispose()
            c.Dispose()
