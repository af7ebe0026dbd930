"""Print the most-sampled SASS lines of a `ncu --page source --csv` dump: python tools/ncu_top.py file.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[i_samp].isdigit()]
tot = sum(int(r[i_samp] or 0) for r in body)
print("total samples", tot)
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(body, key=lambda r: -int(r[i_samp] or 0))[:n]:
    st = {hdr[i][6:]: int(r[i]) for i in stall_cols if r[i] and int(r[i])}
    print(f"{int(r[i_samp]):6d} {100 * int(r[i_samp]) / max(tot, 1):5.1f}% exec={r[i_exec]:>8s} {r[i_src][:90]:90s} {st}")
