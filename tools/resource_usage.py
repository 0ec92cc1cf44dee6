#!/usr/bin/env python
"""Per-kernel registers / stack / static shared memory of the built library (cuobjdump --dump-resource-usage), as a table.
    python tools/resource_usage.py > profiles/<round>_resource_usage.txt
Runs without a GPU."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "splatco_b200", "libsplatco_b200.so")


def main():
    lines = subprocess.run(["cuobjdump", "--dump-resource-usage", SO], capture_output=True, text=True, check=True).stdout.splitlines()
    names, rows = [], []
    for i, ln in enumerate(lines[:-1]):
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            kv = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", lines[i + 1]))
            names.append(m.group(1))
            rows.append([int(kv.get(k, 0)) for k in ("REG", "STACK", "LOCAL", "SHARED")])
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    out = []
    for d, r in zip(dem, rows):
        d = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", d.strip()).replace("void ", "").replace("splatco::", "")
        out.append((d, *r))
    out.sort()
    print("# cuobjdump --dump-resource-usage splatco_b200/libsplatco_b200.so (sm_100a)")
    print("# smem = STATIC shared memory (the decode MLP and blend_bwd2 kernels add dynamic shared memory at launch);")
    print("# stack = per-thread stack bytes (spills or local arrays)")
    print(f"{'kernel':84s} {'regs':>5s} {'stack':>6s} {'local':>6s} {'smem':>7s}")
    for r in out:
        print(f"{r[0][:84]:84s} {r[1]:5d} {r[2]:6d} {r[3]:6d} {r[4]:7d}")


if __name__ == "__main__":
    main()
