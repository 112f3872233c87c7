"""Stall samples per CUDA source line from `ncu -i rep --page source --csv --print-source cuda,sass`.
    python tools/ncu_lines.py file.csv [top] [file-substring]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = sys.argv[3] if len(sys.argv) > 3 else None
per = collections.Counter()
text = {}
cur_file = None
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        cur_file = r[1]
        continue
    if "Warp Stall Sampling (All Samples)" in r:
        hdr = r
        i_line, i_src, i_s = r.index("Line No"), r.index("Source"), r.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) <= i_s or not r[i_s].isdigit():
        continue
    if r[i_line].isdigit():           # a CUDA source row carrying the sum of its SASS rows
        key = (cur_file, int(r[i_line]))
        per[key] += int(r[i_s])
        text[key] = r[i_src].strip()
tot = sum(per.values()) or 1
print("total samples", tot)
for (f, ln), s in per.most_common(top):
    if want and want not in (f or ""):
        continue
    print("%6d %5.1f%%  %s:%d  %s" % (s, 100 * s / tot, (f or "?").split("/")[-1], ln, text[(f, ln)][:110]))
