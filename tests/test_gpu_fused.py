"""GPU tests of the fused / tensor-core path: the tcgen05 plumbing in isolation, then the fused frame renderer against the
op-by-op path (which is itself parity-checked against the reference kernels in test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("K,N", [(16, 16), (32, 64), (64, 64), (64, 16), (64, 80), (80, 64), (96, 64), (128, 128)])
def test_tcgen05_gemm_tile_matches_torch_fp32(K, N):
    """out = A @ W^T with fp16 operands and fp32 accumulation in TMEM vs a plain PyTorch fp32 matmul of the same fp16 values."""
    from radnerf_b200 import abi as L
    g = torch.Generator(device="cpu").manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g).half().to(DEV)
    W = torch.randn(N, K, generator=g).half().to(DEV)
    out = torch.full((128, N), float("nan"), device=DEV)
    L.check(L.lib().rn_selftest_umma(L.ptr(A), L.ptr(W), L.ptr(out), K, N, L.cur_stream()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    assert torch.isfinite(out).all()
    # products of fp16 values are exact in fp32; only the summation order differs
    assert (out - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
