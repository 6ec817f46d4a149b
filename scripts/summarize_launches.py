"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last f/g evaluation, per kernel and in
launch order.  python scripts/summarize_launches.py launches.csv > profiles/xxx.md"""
import collections
import csv
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = []
for r in rows[1:]:
    try:
        data.append((r[ki], float(r[vi].replace(",", ""))))
    except ValueError:
        pass
starts = [i for i, (k, v) in enumerate(data) if "conv1_1_fwd" in k or "conv1_1_tc_kernel" in k]
last = data[starts[-1]:]
tot = sum(v for _, v in last)
print("# ncu launch list: %s" % path)
print()
print("Last f/g evaluation: %d launches, %.1f us summed (cold-cache, serialised: compare shares, not absolutes)" % (len(last), tot / 1000))
print()
print("| kernel | launches | us | share |")
print("|---|---|---|---|")
agg = collections.OrderedDict()
for k, v in last:
    kk = k.split("(")[0].replace("void ", "").replace("stx::<unnamed>::", "")
    agg.setdefault(kk, [0.0, 0])
    agg[kk][0] += v
    agg[kk][1] += 1
for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("| `%s` | %d | %.1f | %.1f%% |" % (k, c, v / 1000, 100 * v / tot))
print()
print("Launch order (us): " + " ".join("%s:%.0f" % (k.split("(")[0].split("::")[-1].replace("_kernel", ""), v / 1000) for k, v in last))
