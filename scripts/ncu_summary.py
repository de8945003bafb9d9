"""Summarise an ncu report (.ncu-rep) into the handful of counters DESIGN.md / profiles/ quote.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [updates_per_launch] > profiles/xxx.txt"""
import csv, subprocess, sys, io
from collections import Counter

rep = sys.argv[1]
updates = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_global_red.sum"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name"))
    for k in KEYS:
        if k in d: print(f"  {k:75s} {d[k]:>18s} {u[k]}")
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            if float(d[k] or 0) >= 0.2: print(f"  {k:75s} {d[k]:>18s}")
    if updates:
        wi = float(d["smsp__inst_executed.sum"]); th = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"])
        rd = float(d["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_write.sum"]]
        print(f"  derived: thread-instructions per update = {wi*th/updates:.1f}; DRAM bytes per update = {(rd+wr)/updates:.1f} "
              f"(algorithmic 32); dram traffic per launch = {(rd+wr)/1e9:.3f} GB")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
try:
    h = rows[1]; ia, ii = h.index("Source"), h.index("Instructions Executed")
    c = Counter(); tot = 0
    for r in rows[2:]:
        if len(r) <= ii or not r[ii].isdigit(): continue
        t = r[ia].split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(r[ii]); tot += int(r[ii])
    print("  SASS mix (share of executed warp instructions):", ", ".join(f"{o} {n/tot*100:.1f}%" for o, n in c.most_common(16)))
except Exception as e:
    print("  (no SASS page:", e, ")")
