"""Dictionary inputs shared by the CPU (oracle pin) and GPU (parity) tests.  The dictionaries are produced by system libzstd
(ZDICT_trainFromBuffer: dictionary TRAINING is not on the GPU path); Decompressor.LoadDictionary + Unwrap and
Compressor.LoadDictionary + Wrap are."""
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


def compress_dictionaries(z):
    """dictionaries() plus the shapes that matter to the COMPRESSOR: content below / at the 8-byte floor, a dictionary larger than
    a block, one larger than the level-1 window of small inputs."""
    text = dg.text_like(48 * dg.FRAME)
    d = dict(dictionaries(z))
    d.update({"raw_8": b"abcdefgh", "raw_9": b"abcdefghi", "raw_110k": text[1_000_000:1_110_000].tobytes(), "raw_300k": text[:300_000].tobytes()})
    return d


def compress_payloads(seed=5, n_random=24):
    """Sizes around the attach / copy cut-offs of ZSTD_shouldAttachDict (8 KiB for ZSTD_fast, 16 KiB for ZSTD_dfast,
    ZstdCompress.cs:2725-2744), around one block, multi-block, plus random sizes; text and binary sources."""
    text = dg.text_like(48 * dg.FRAME)
    sil = dg.silesia_mix(8 * dg.FRAME)
    rng = np.random.default_rng(seed)
    sizes = [0, 1, 100, 8191, 8192, 8193, 16383, 16384, 16385, 131071, 131072, 131073, 262144, 600_000] + [int(rng.integers(0, 300_000)) for _ in range(n_random)]
    out = []
    for i, n in enumerate(sizes):
        srcsel = text if i % 3 else sil
        off = int(rng.integers(0, srcsel.size - n - 1))
        out.append(np.ascontiguousarray(srcsel[off:off + n]))
    out.append(np.ascontiguousarray(text[2_010_000:2_040_000]))              # inside raw_50k: matches reach deep into the dictionary
    # inputs that repeat their FIRST bytes: with an attached dictionary a candidate at the first position of the input counts as an empty
    # cell (`matchIndex <= prefixStartIndex`, ZstdFast.cs:441) -- the soak found the group kernels matching against it (seed 991177)
    out.append(np.resize(text[5000:5037], 4095).copy())
    out.append(np.resize(np.arange(7, dtype=np.uint8), 3000).copy())
    out.append(np.resize(sil[100:164], 12000).copy())
    out.append(np.resize(text[7000:7300], 40000).copy())
    return out
