"""Hardware self-test of the tcgen05 / TMEM / TMA SW128 plumbing (csrc/umma.cuh) that the
tensor-core correlation path is built on: a bf16 GEMM through fn2b200_debug_umma_gemm."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [64, 256])
def test_umma_gemm_matches_cpu(K):
    from flownet2_b200._lib import LIB, check
    g = torch.Generator().manual_seed(K)
    A = torch.randn(128, K, generator=g).bfloat16()
    B = torch.randn(144, K, generator=g).bfloat16()
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.full((128, 144), float("nan"), device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(LIB.fn2b200_debug_umma_gemm(ctypes.c_void_p(Ad.data_ptr()), ctypes.c_void_p(Bd.data_ptr()),
                                      ctypes.c_void_p(D.data_ptr()), K, st), "debug_umma_gemm")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (D.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err
