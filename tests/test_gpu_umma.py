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


@pytest.mark.parametrize("mode", [-144, -145])
def test_umma_gemm_noswizzle_a_mnmajor_b(mode):
    """Operand forms of the tensor-core backward: A built by threads (K-major, no swizzle),
    B^T = [K][N] row-major consumed as an MN-major SW128 operand."""
    from flownet2_b200._lib import LIB, check
    g = torch.Generator().manual_seed(7)
    A = torch.randn(128, 144, generator=g).bfloat16()
    Bt = torch.randn(144, 64, generator=g).bfloat16()
    Ad, Bd = A.cuda(), Bt.cuda()
    D = torch.full((128, 64), float("nan"), device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(LIB.fn2b200_debug_umma_gemm(ctypes.c_void_p(Ad.data_ptr()), ctypes.c_void_p(Bd.data_ptr()),
                                      ctypes.c_void_p(D.data_ptr()), mode, st), "debug_umma_gemm(2)")
    torch.cuda.synchronize()
    ref = A.double() @ Bt.double()
    err = (D.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err
