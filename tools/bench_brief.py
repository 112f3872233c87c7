"""Run bench.py and print a one-line digest (dev aid)."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline"] + sys.argv[1:], capture_output=True, text=True)
for l in out.stdout.splitlines():
    if l.startswith("{"):
        d = json.loads(l)
        r = d["roofline"]
        print("value %.0f clips/s  %.3f ms/step | e2e %.0f | gemm %.3f ms (%.0f TF/s, frac %.3f) other %.3f ms | mel %.0f GB/s eval %.0f GB/s | launches %d | clocks %s" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], r["gemm_ms_per_step"], r["achieved"], r["frac"], r["other_ms_per_step"],
            d["roofline_mel"]["achieved"], d["roofline_eval"]["achieved"], d["gpu_launches"], d["clocks"]))
    else:
        print(l)
print(out.stderr[-2000:])
