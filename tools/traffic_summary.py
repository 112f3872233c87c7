"""Sum DRAM traffic per kernel class over the LAST step of an ncu launch list captured with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum  ->  profiles/r2_traffic.json
    python tools/traffic_summary.py gpurun_out/traffic.csv profiles/r2_traffic.json"""
import collections
import csv
import json
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
by_id = collections.OrderedDict()
for r in rows:
    d = by_id.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "nsecond": 1e-3}.get(unit, 1)
    d[r["Metric Name"]] = v * scale
ks = list(by_id.values())
starts = [i for i, k in enumerate(ks) if "logmel" in k["name"]]
step = ks[starts[-1]:]
agg = collections.OrderedDict()
for k in step:
    m = re.search(r"([A-Za-z_0-9]+_kernel)", k["name"])
    n = m.group(1) if m else k["name"][:40]
    a = agg.setdefault(n, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    a["launches"] += 1
    a["us"] += k.get("gpu__time_duration.sum", 0.0)
    a["dram_bytes"] += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
out = {"note": "ncu --clock-control none, tools/profile_step.py (B=256), --cache-control none (warm L2), last of 2 steps; launches serialised by ncu",
       "per_kernel": agg,
       "conv_gemm_bytes_per_step": sum(a["dram_bytes"] for n, a in agg.items() if n.startswith("conv_gemm")),
       "logmel_bytes_per_launch": sum(a["dram_bytes"] for n, a in agg.items() if n.startswith("logmel")),
       "eval_bytes_per_launch": agg.get("eval_l1_pck_kernel", {}).get("dram_bytes")}
json.dump(out, open(sys.argv[2], "w"), indent=1)
for n, a in agg.items():
    print("%-32s n=%3d %9.1f us %10.1f MB" % (n, a["launches"], a["us"], a["dram_bytes"] / 1e6))
