"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line (developer tool)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict(); cur = None; hdr = None; fname = ''
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; iI = hdr.index('Instructions Executed'); iS = hdr.index('# Samples'); iT = hdr.index('Thread Instructions Executed'); continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[0] not in ('', '-') and r[2] == '-':
        cur = (fname, r[0], r[1].strip()[:110]); agg.setdefault(cur, [0, 0, 0]); continue
    if cur is None: continue
    try:
        agg[cur][0] += int(r[iI]); agg[cur][1] += int(r[iS]); agg[cur][2] += int(r[iT])
    except ValueError: pass
tot = [sum(v[k] for v in agg.values()) for k in range(3)]
print('total inst %d samples %d thread-inst %d' % tuple(tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%5.1f%% smp %5.1f%% inst act %4.1f | %s:%s %s' % (100 * v[1] / max(tot[1], 1), 100 * v[0] / max(tot[0], 1), v[2] / max(v[0], 1), k[0], k[1], k[2]))
