"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/summarize_launches.py in.csv out_prefix
Writes <out_prefix>_lp_launches.csv (this library's kernels only, in launch order) and <out_prefix>_summary.txt."""
import collections
import csv
import sys

src, out = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, idc, gs, bs = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "ID", "Grid Size", "Block Size"))
ours = [r for r in rows[hi + 1:] if len(r) > mv and "lp::" in r[kn]]
if not ours:  # list taken with `--kernel-name-base demangled -k regex:lp::`: already filtered, names come without the namespace
    ours = [r for r in rows[hi + 1:] if len(r) > mv]
with open(out + "_lp_launches.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["ID", "Kernel", "Grid", "Block", "gpu__time_duration.sum [ns]"])
    for r in ours:
        w.writerow([r[idc], r[kn].split("(")[0].replace("void ", ""), r[gs], r[bs], r[mv]])
agg = collections.OrderedDict()
for r in ours:
    a = agg.setdefault(r[kn].split("(")[0].replace("void ", ""), [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(out + "_summary.txt", "w") as f:
    f.write(f"source: {src}\n{len(ours)} launches of liblitparrot_b200 kernels, {tot / 1e3:.1f} us device time (ncu: cold cache, serialised; compare SHARES)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:60s} n={v[0]:5d} share={100 * v[1] / tot:5.1f}%  avg={v[1] / v[0] / 1e3:8.2f} us\n")
print(open(out + "_summary.txt").read())
