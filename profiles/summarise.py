"""Turn an .ncu-rep (ncu --set full) into the small text summary committed under profiles/.
Usage: python profiles/summarise.py gpurun_out/prof_X.ncu-rep profiles/r01_X.txt"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none : {rep.split('/')[-1]}\n")
    for r in rows[2:]:
        f.write(f"\n## {r[h.index('Kernel Name')]}\n")
        for k in keys:
            if k in h:
                f.write(f"{k:72s} {r[h.index(k)]} {units[h.index(k)]}\n")
        st = []
        for i, name in enumerate(h):
            if name.startswith("smsp__average_warps_issue_stalled") and name.endswith("per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), name.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        f.write("stall reasons (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]) + "\n")
print("wrote", out)
