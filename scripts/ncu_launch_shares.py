"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list (developer tool)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    ns = float(r[-1].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.3f | %.1f%% |" % (k[:90], v[0], v[1] / 1e6, 100 * v[1] / tot))
print("\nTotal device time in the captured launches: %.1f ms." % (tot / 1e6))
