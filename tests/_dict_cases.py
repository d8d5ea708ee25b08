"""Dictionary inputs shared by the CPU (oracle pin) and GPU (parity) tests.  The dictionaries and the frames that use them
are produced by system libzstd (ZDICT_trainFromBuffer / ZSTD_compress_usingDict): dictionary training and dictionary
compression are not on the GPU path, only decompression is (Decompressor.LoadDictionary + Unwrap)."""
import numpy as np

from zstdsharp_b200 import datagen as dg


def dictionaries(z):
    text = dg.text_like(48 * dg.FRAME)
    samples = [text[i * 4000:(i + 1) * 4000].tobytes() for i in range(600)]
    return {
        "zdict_32k": z.train_dictionary(samples, 32768),                      # magic, entropy tables, repcodes, content
        "zdict_4k": z.train_dictionary(samples[:200], 4096),
        "raw_50k": text[2_000_000:2_050_000].tobytes(),                       # no magic: pure content
        "raw_7": b"abcdefg",                                                  # below 8 bytes: content by definition
    }


def payloads():
    text = dg.text_like(48 * dg.FRAME)
    sil = dg.silesia_mix(4 * dg.FRAME)
    return [text[3_000_000:3_000_000 + n] for n in (1, 100, 3000, 70000, dg.FRAME, 200_000, 400_000)] + \
           [text[2_010_000:2_010_000 + 30_000],                               # inside the raw dictionary: matches reach deep into it
            sil[:dg.FRAME], dg.literal_heavy(5000), np.zeros(0, dtype=np.uint8)]
