"""Input shapes shared by the CPU (oracle pin) and GPU (parity) tests."""
import numpy as np

from zstdsharp_b200 import datagen as dg

FRAME = dg.FRAME


def multiblock_inputs():
    """Inputs above 128 KiB: what changes from block to block of a frame is the window (matches into earlier blocks, offsets up
    to windowSize), the repcodes (only a compressed block confirms them), the Huffman table (set_repeat / new table / raw),
    the RLE block type (allowed after the first block) and the last-block bit."""
    text = dg.text_like(24 * FRAME)
    sil = dg.silesia_mix(24 * FRAME)
    rnd = dg.incompressible(3 * FRAME)
    lit = dg.literal_heavy(4 * FRAME)
    zeros = np.zeros(3 * FRAME, dtype=np.uint8)
    cases = {
        "text_128k+1": text[:FRAME + 1],
        "text_128k+6": text[:FRAME + 6],                  # last block below MIN_CBLOCK_SIZE + 3 + 1: raw
        "text_128k+40": text[:FRAME + 40],                # last block takes the serial match kernel
        "text_200k": text[:200_000],
        "text_256k": text[:2 * FRAME],                    # last size of the 256 KiB parameter table
        "text_256k+1": text[:2 * FRAME + 1],              # windowLog 19, hashLog 14 / 17
        "text_600k": text[:600_000],                      # window (512 KiB at level 1) smaller than the frame: no singleSegment
        "text_3m": text[:3_000_000],
        "silesia_3m": sil[:3_000_000],
        "zeros_384k": zeros,                              # RLE blocks after the first one
        "zeros_300k": zeros[:300_000],
        "text|random|text": np.concatenate([text[:FRAME + 5000], rnd[:FRAME], text[FRAME:3 * FRAME]]),   # raw block in the middle: repcodes / table not confirmed
        "random_300k": rnd[:300_000],
        "literal_heavy_512k": lit,                        # Huffman table reuse (set_repeat) vs new table
        "zeros|text": np.concatenate([zeros[:FRAME], text[:FRAME + 100]]),
        "text|zeros|text": np.concatenate([text[:FRAME], zeros[:FRAME], text[:70000]]),
        "period_1m": np.tile(text[:50_000], 20),          # long matches at offsets beyond one block
        "ramp_1m": dg.byte_ramp(1_000_000),
        "tiny_tail_literals": np.concatenate([text[:FRAME], lit[:900]]),   # last block literals <= 1024: preferRepeat
    }
    return cases
