"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel lines kept under profiles/.
    ncu -i rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem"), ("smsp__inst_executed.sum", "warp_inst")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    name = name.split("(")[0].split("::")[-1]
    parts = []
    for key, short in WANT:
        if key in idx and r[idx[key]] != "":
            v = r[idx[key]]
            try:
                f = float(v.replace(",", ""))
                v = ("%.4g" % f)
            except ValueError:
                pass
            u = units[idx[key]]
            parts.append("%s=%s%s" % (short, v, (" " + u) if u and u not in ("%",) and short in ("time", "dram_rd", "dram_wr", "dyn_smem") else ""))
    print("%-28s %s" % (name[:28], "  ".join(parts)))
