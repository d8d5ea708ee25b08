"""Reproducer kept from the round-2 assertion-build hunt (profiles/r02_notes.md, "assertion build"): 16 text-like 128 KiB frames whose last
sequences are decoded by dec_seq_kernel's tail loop; prints the items whose decoded bytes differ and where.  Select the library under
test with ZSTDB200_LIB (e.g. zstdsharp_b200/_build/libzstdb200_dbg.so); ZSTDB200_DUMP_ITEM=<i> makes the assertion build print that
item's last literals / records / output bytes between kernels.  Needs a GPU; uses the oracle only as the checker."""
import sys, os
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from zstdsharp_b200 import datagen as dg, api
from _oracle import oracle
o = oracle(); F = dg.FRAME
data = dg.WORKLOADS['text'](16 * F)
chunks = [data[i * F:(i + 1) * F] for i in range(16)]
frames = [o.compress(c, 3) for c in chunks]
lits, tr = o.decode_stages(frames[0])
print('oracle last 6 sequences of item 0:', tr[-6:].tolist(), 'nbSeq', len(tr), 'litSize', lits.size, flush=True)
with api.Decompressor() as d:
    outs = d.UnwrapBatch(frames, raise_on_error=False)
bad = [i for i, (a, c) in enumerate(zip(outs, chunks)) if a != c.tobytes()]
print('bad items', bad)
for i in bad:
    a = np.frombuffer(outs[i], dtype=np.uint8); d = np.nonzero(a != chunks[i])[0]
    print('  item', i, 'wrong bytes at', d[0], '..', d[-1], 'got', a[d[0]-4:].tobytes().hex(), 'want', chunks[i][d[0]-4:].tobytes().hex(), 'next item head', chunks[(i+1) % 16][:8].tobytes().hex())
