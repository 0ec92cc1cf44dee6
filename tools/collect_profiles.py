#!/usr/bin/env python
"""Turn the scratch outputs of tools/make_profiles.sh (gpurun_out/) into the committed evidence under profiles/:
    <tag>_launches.csv          the ncu launch list (per-launch gpu__time_duration; cold-cache, serialised)
    <tag>_ncu_<name>.txt        tools/ncu_summary.py of each `ncu --set full` capture
    <tag>_traffic.json          per kernel: time, dram bytes read + written per launch (what bench.py's
                                roofline.traffic quotes), issue-active, tensor-pipe-active
    python tools/collect_profiles.py <tag>"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
traffic = {}
if os.path.exists(os.path.join(src, f"launches_{tag}.csv")):
    shutil.copy(os.path.join(src, f"launches_{tag}.csv"), os.path.join(dst, f"{tag}_launches.csv"))
for name in ("blend_bwd", "blend_fwd", "decode", "rest", "optim"):
    rep = os.path.join(src, f"prof_{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(dst, f"{tag}_ncu_{name}.txt"), "w").write(txt)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]

    def get(r, k):
        if k not in h:
            return None
        v = num(r[h.index(k)])
        return None if v is None else v * UNIT.get(units[h.index(k)], 1.0)

    for r in rows[2:]:
        k = r[h.index("Kernel Name")].split("(")[0]
        k = k.replace("void ", "").replace("splatco::", "")
        if k.startswith("blend_bwd2_kernel"):        # bench.py looks the blend kernels up by stage name
            k = "blend_bwd_kernel"
        elif k.startswith("blend_fwd2_kernel"):
            k = "blend_fwd_kernel"
        e = traffic.setdefault(k, {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0, "issue_active_pct": 0.0, "tensor_active_pct": 0.0,
                                   "warp_instructions": 0.0, "threads_per_instruction": 0.0})
        e["launches"] += 1
        e["time_us"] += get(r, "gpu__time_duration.sum") or 0.0
        e["dram_bytes"] += (get(r, "dram__bytes_read.sum") or 0.0) + (get(r, "dram__bytes_write.sum") or 0.0)
        e["issue_active_pct"] += get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0.0
        e["tensor_active_pct"] += get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0.0
        e["warp_instructions"] += get(r, "smsp__inst_executed.sum") or 0.0
        e["threads_per_instruction"] += get(r, "smsp__thread_inst_executed_per_inst_executed.ratio") or 0.0
for k, e in traffic.items():
    n = e.pop("launches")
    traffic[k] = {"captured_launches": n, **{kk: round(v / n, 3) for kk, v in e.items()}}
json.dump(traffic, open(os.path.join(dst, f"{tag}_traffic.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(traffic, indent=1, sort_keys=True))
