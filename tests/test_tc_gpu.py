"""GPU self-test of the tcgen05 primitives (splatco_b200/csrc/tc.cuh): 3xTF32 tile GEMM vs an fp64
reference.  3xTF32 must reach ~fp32 accuracy (1e-5 relative to the row/column norms); the plain
single-pass TF32 variant is only checked to be TF32-accurate (sanity of the descriptor layout)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(M, N, K, variant, seed=0):
    from splatco_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    Ad, Bd = A.cuda(), B.cuda()
    Cd = torch.full((M, N), float("nan"), device="cuda")
    rc = L.splatco_tc_gemm_selftest(M, N, K, C.c_void_p(Ad.data_ptr()), C.c_void_p(Bd.data_ptr()),
                                    C.c_void_p(Cd.data_ptr()), variant, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, L.splatco_last_error()
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t()).numpy()
    scale = np.sqrt(K)
    return np.abs(Cd.cpu().numpy().astype(np.float64) - ref).max() / scale


@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (300, 32, 72), (256, 96, 100), (1000, 112, 96), (128, 96, 8), (130, 110, 99)])
def test_3xtf32_tile_gemm_fp32_accurate(M, N, K):
    err = run(M, N, K, variant=0)
    assert err < 2e-5, err      # ~4e-6 measured: 3xTF32 keeps ~21 mantissa bits per product (single pass: ~3.5e-3)


def test_single_pass_tf32_is_tf32_accurate():
    err = run(256, 96, 104, variant=2)
    assert 1e-6 < err < 5e-3, err
