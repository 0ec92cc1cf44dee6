"""Probe of the MN-major tcgen05 operand path (splatco_tc_wgrad_selftest): every variant runs in its own
process (a malformed descriptor can trap the context), errors vs an fp64 reference are printed as JSON lines."""
import ctypes as C
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(N, variant):
    import numpy as np
    import torch
    from splatco_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(3)
    At = torch.randn(128, 128, generator=g)
    Bt = torch.randn(128, N, generator=g)
    Ad, Bd = At.cuda(), Bt.cuda()
    Cd = torch.full((128, N), float("nan"), device="cuda")
    rc = L.splatco_tc_wgrad_selftest(N, C.c_void_p(Ad.data_ptr()), C.c_void_p(Bd.data_ptr()), C.c_void_p(Cd.data_ptr()),
                                     variant, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, L.splatco_last_error()
    torch.cuda.synchronize()
    ref = (At.double().t() @ Bt.double()).numpy()
    got = Cd.cpu().numpy().astype(np.float64)
    err = float(np.nanmax(np.abs(got - ref)) / np.sqrt(128)) if np.isfinite(got).any() else float("nan")
    print(json.dumps({"N": N, "variant": variant, "err": err, "nan": int(np.isnan(got).sum())}))


if __name__ == "__main__":
    if len(sys.argv) == 3:
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        for N in (96,):
            for variant in [0, 8, 1, 2, 3, 5, 6, 7, 1 | 128, 2 | 128, 3 | 128, 5 | 128, 6 | 128, 7 | 128, 3 | 16, 3 | 32, 3 | 64, 3 | 96, 7 | 16, 7 | 32]:
                try:
                    r = subprocess.run([sys.executable, __file__, str(N), str(variant)], capture_output=True, text=True, timeout=120)
                    out = r.stdout.strip().splitlines()
                    print(out[-1] if out else json.dumps({"N": N, "variant": variant, "rc": r.returncode, "stderr": r.stderr[-300:]}))
                except subprocess.TimeoutExpired:
                    print(json.dumps({"N": N, "variant": variant, "timeout": True}))
                sys.stdout.flush()
