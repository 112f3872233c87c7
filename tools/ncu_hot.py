"""Top stall-sample instructions from `ncu -i rep --page source --csv [--kernel-id :::N]` output.
    ncu -i X.ncu-rep --page source --csv --kernel-id :::2 > src.csv; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for r in rows[hi + 1:]:
    if len(r) <= i_ex or not r[i_s].isdigit():
        continue
    data.append((int(r[i_s]), r[i_src].strip(), int(r[i_ex] or 0)))
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "instructions", len(data))
for s, src, ex in sorted(data, key=lambda d: -d[0])[:top]:
    print("%6d %5.1f%%  ex=%8d  %s" % (s, 100 * s / tot, ex, src[:120]))
