"""Summarise an .ncu-rep (run here, no GPU): python tools/ncu_summary.py report.ncu-rep [launch_index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2 + idx]
m = {h: (v, u) for h, u, v in zip(hdr, units, data)}
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for k in keys:
    if k in m: print(f"{k:75s} {m[k][0]} {m[k][1]}")
st = sorted(((float(v[0]), k) for k, v in m.items() if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("_not_issued") and v[0].replace('.','',1).isdigit()), reverse=True)
tot = sum(v for v, _ in st) or 1
print("stall samples:", ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_','')}={100*v/tot:.0f}%" for v, k in st[:9]))
