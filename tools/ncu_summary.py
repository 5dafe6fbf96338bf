"""Text summary of an `ncu --set full` report: one block per kernel launch with the figures DESIGN.md
quotes (duration, registers, DRAM bytes, pipe use, occupancy, stall cycles per issued instruction).
usage: python tools/ncu_summary.py report.ncu-rep [nodes_per_launch]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nodes = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
WANT = [
    ("duration", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("regs/thread", "launch__registers_per_thread"), ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("dram % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("alu pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("lsu pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("active threads/warp instr", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("warps/scheduler", "smsp__warps_active.avg.per_cycle_active"),
    ("eligible warps/scheduler", "smsp__warps_eligible.avg.per_cycle_active"),
    ("local loads (spills)", "smsp__sass_inst_executed_op_local_ld.sum"),
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for r in data:
    print("## " + r[col["Kernel Name"]])
    for label, key in WANT:
        if key in col and r[col[key]] != "":
            print(f"  {label:30s} {r[col[key]]} {units[col[key]]}")
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            v = num(r[i])
            if v and v >= 0.02:
                stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    print("  stall cycles per issued instr  " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))
    if nodes:
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = num(r[col[key]])
            tot += (v or 0.0) * TO_BYTES.get(units[col[key]], 1.0)
        print(f"  dram bytes per node            {tot / nodes:.0f} B (read+write)")
        wi = num(r[col['smsp__inst_executed.sum']])
        if wi:
            print(f"  thread instructions per node   {32 * wi / nodes:.0f}")
    print()
