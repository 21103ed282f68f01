"""Unit test of the tcgen05/TMEM helpers (csrc/umma.cuh) through the C ABI's bring-up entry point."""
import ctypes as C

import pytest
import torch

from b200taco import _cabi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [64, 256, 1792])
def test_umma_gemm_matches_fp32_matmul(K):
    lib = _cabi.lib()
    g = torch.Generator().manual_seed(K)
    A = (torch.randn(128, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    Bm = (torch.randn(32, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    D = torch.zeros(128, 32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check(lib.b200taco_debug_umma_gemm(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), K, 0, st), "debug_umma_gemm")
    torch.cuda.synchronize()
    ref = A.float() @ Bm.float().t()
    err = float((D - ref).abs().max())
    print("K", K, "max err", err, "ref max", float(ref.abs().max()))
    assert err <= 1e-4 * max(1.0, float(ref.abs().max()))
