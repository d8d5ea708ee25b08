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


def handbuilt_small_4stream_frames():
    """Hand-built frames whose literal section has FOUR Huffman streams although it is tiny (segments of 20 / 7 literals; the
    reference's encoder never does this below 256 literals, other encoders may) and whose sequences take literal runs of
    <= 40 bytes across TWO segment boundaries.  Returns [(frame, expected_bytes)]."""
    code = {1: "000", 2: "001", 0: "01", 3: "1"}           # weights {0: 2, 1: 1, 2: 1, 3: 3 (implied)}, tableLog 3

    def stream(symbols):
        bits = "1" + "".join(code[s] for s in symbols)      # end mark, then the symbols in decoding order
        bits = "0" * (-len(bits) % 8) + bits
        return int(bits, 2).to_bytes(len(bits) // 8, "little")

    out = []
    rng = np.random.default_rng(5)
    for lit_size, lls in ((80, (32, 38)), (28, (13, 13)), (80, (39, 39))):
        lits = rng.choice([0, 1, 2, 3], size=lit_size, p=[0.3, 0.1, 0.1, 0.5]).astype(np.uint8)
        seg = (lit_size + 3) // 4
        streams = [stream(lits[i * seg:min(lit_size, (i + 1) * seg)].tolist()) for i in range(4)]
        tree = bytes([127 + 3, 0x21, 0x10])
        jump = b"".join(len(s).to_bytes(2, "little") for s in streams[:3])
        payload = tree + jump + b"".join(streams)
        lit_hdr = (2 | (1 << 2) | (lit_size << 4) | (len(payload) << 14)).to_bytes(3, "little")
        if lls[0] >= 32:        # LL code 22: 32 + 3 extra bits; RLE tables: no state bits
            ll_code, extra = 22, "".join(format(ll - 32, "03b") for ll in lls)
        else:                   # LL codes below 16 have no extra bits: both runs must then be equal ... use code = ll directly twice
            ll_code, extra, lls = lls[0], "", (lls[0], lls[0])
        bits = "1" + extra
        bits = "0" * (-len(bits) % 8) + bits
        seq = bytes([2, 0x54, ll_code, 0, 0]) + int(bits, 2).to_bytes(len(bits) // 8, "little")
        block = lit_hdr + payload + seq
        expect = bytearray()
        pos = 0
        for ll in lls:
            expect += lits[pos:pos + ll].tobytes(); pos += ll
            expect += bytes([expect[-1]]) * 3                # matchLength code 0 = 3 bytes at repeat offset 1
        expect += lits[pos:].tobytes()
        hdr = (1 | (2 << 1) | (len(block) << 3)).to_bytes(3, "little")
        frame = bytes.fromhex("28b52ffd") + bytes([0x20, len(expect)]) + hdr + block
        out.append((frame, bytes(expect)))
    return out


def dfast_last_window_case():
    """A two-block input for ZSTD_dfast (levels 2 / 3 at this size) found by the round-2 soak (seed 993001): the first block is incompressible, so
    the search step has grown to ~440 bytes at its end; the last position the reference's loop still visits (130562) and the first one that
    fails the loop condition but still has 8 readable bytes (131005) hold the same 8 bytes.  The reference writes 130562 into the long table;
    a window-parallel match finder must not let the unvisited position shadow that write.  The second block repeats those bytes behind a decoy
    that owns the short-table bucket, so only the long table finds the 24-byte match."""
    rng = np.random.default_rng(20261019)
    B = 131072
    a = rng.integers(0, 256, size=B + 400, dtype=np.uint8)
    ip, step, next_step, last = 1, 1, 1 + 256, None                  # ZstdDoubleFast.cs:100-165: positions of an all-literal block
    while ip + step <= B - 8:
        last = ip
        ip1 = ip + step
        if ip1 >= next_step:
            step += 1; next_step += 256
        ip = ip1
    assert last == 130562 and ip == 131005
    marker = a[last:last + 24].copy()
    a[ip:ip + 16] = marker[:16]
    a[B + 40:B + 46] = marker[:6]                                     # decoy: same first bytes, different continuation
    a[B + 86:B + 110] = marker
    return a
