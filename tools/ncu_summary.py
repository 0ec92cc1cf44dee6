#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--csv out.csv]
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__shared_mem_per_block_static", "static smem"),
    ("launch__occupancy_limit_registers", "occ lim regs (CTAs)"), ("launch__occupancy_limit_shared_mem", "occ lim smem (CTAs)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instr"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor pipe instr"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
]
STALLS = "smsp__average_warps_issue_stalled_{}_per_issue_active.ratio"
STALL_NAMES = ["barrier", "short_scoreboard", "long_scoreboard", "wait", "not_selected", "math_pipe_throttle", "mio_throttle",
               "lg_throttle", "branch_resolving", "dispatch_stall", "no_instruction", "membar", "drain", "imc_miss", "tex_throttle",
               "sleeping", "selected"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0]
        lines.append(f"== {name}  (launch id {r[h.index('ID')]})")
        for k, label in KEYS:
            if k in h and r[h.index(k)] not in ("", "n/a"):
                lines.append(f"   {label:28s} {r[h.index(k)]} {units[h.index(k)]}")
        st = []
        for n in STALL_NAMES:
            k = STALLS.format(n)
            if k in h:
                try:
                    v = float(r[h.index(k)])
                except ValueError:
                    continue
                if v >= 0.05:
                    st.append((v, n))
        lines.append("   stalls (warps per issue):    " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)))
    print("\n".join(lines))
    if "--csv" in sys.argv:
        open(sys.argv[sys.argv.index("--csv") + 1], "w").write(out)


if __name__ == "__main__":
    main()
