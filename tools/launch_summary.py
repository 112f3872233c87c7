"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the LAST
hot-path step (from the last logmel launch to the end).   python tools/launch_summary.py <csv> [--list]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def label(full):
    m = re.search(r"conv_gemm_kernel<(?:\(int\))?(\d+)", full)
    if m:
        return "conv_gemm_kernel<%s>" % m.group(1)
    m = re.search(r"logmel512_kernel<(\w+)", full)
    if m:
        return "logmel512_kernel<%s>" % m.group(1)
    m = re.search(r"([A-Za-z_0-9]+_kernel)", full)
    if m:
        return m.group(1)
    name = re.sub(r"\(.*", "", full)
    return name.split("::")[-1][:48]


def main():
    rows = [r for r in load(sys.argv[1]) if r.get("Metric Name", "gpu__time_duration.sum") == "gpu__time_duration.sum"]
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
    for r in rows:
        r["Metric Value"] = str(float(r["Metric Value"].replace(",", "")) * scale.get(r.get("Metric Unit", "ns"), 1e-3) * 1000.0)
    ks = [(label(r["Kernel Name"]), float(r["Metric Value"]) / 1000.0, r["Grid Size"], r["Kernel Name"]) for r in rows]
    starts = [i for i, k in enumerate(ks) if "logmel" in k[3]]
    step = ks[starts[-1]:]
    total = sum(t for _, t, _, _ in step)
    agg = collections.OrderedDict()
    for n, t, _, _ in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    print("last step: %d launches, %.1f us (cold-cache, serialised: compare SHARES)" % (len(step), total))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s n=%3d %9.1f us %5.1f%%" % (k, c, t, 100 * t / total))
    if "--list" in sys.argv:
        for i, (n, t, g, _) in enumerate(step):
            print("%3d %-40s %9.1f us grid=%s" % (i, n, t, g))


if __name__ == "__main__":
    main()
