"""-m gpu: the raw tcgen05 GEMM core (asn_gemm_bf16_tn) against an fp64 matmul of the same
bf16-rounded operands.  fp32 accumulation of exact bf16 products: tolerance 2e-5 norm-wise."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from gpu_util import gpu

pytestmark = gpu


@pytest.mark.parametrize("M,N,K,split", [
    (128, 128, 64, 1), (128, 256, 128, 1), (256, 128, 512, 1), (300, 200, 264, 1),
    (1000, 688, 1024, 1), (2048, 176, 4096, 4), (64, 48, 72, 1), (129, 257, 2048, 3),
])
def test_gemm_bf16_tn(M, N, K, split):
    from adaptsegnet_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    c = ops.gemm_bf16_tn(a, b, split_k=split)
    torch.cuda.synchronize()
    got = c.sum(0).double().cpu().numpy()
    ref = (a.double() @ b.double().t()).cpu().numpy()
    assert rel_err(got, ref) < 2e-5


@pytest.mark.parametrize("M,N,K,split", [
    (128, 128, 64, 1), (64, 64, 200, 1), (256, 192, 512, 1), (264, 200, 300, 1),
    (2048, 688, 14651, 5), (1024, 688, 4225, 9), (72, 48, 1000, 2),
])
def test_gemm_bf16_nt_mn(M, N, K, split):
    """C = A[K,M]^T . B[K,N] with both operands MN-major (the ASPP / discriminator weight-gradient form)."""
    from adaptsegnet_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(K, M, device="cuda").to(torch.bfloat16)
    b = torch.randn(K, N, device="cuda").to(torch.bfloat16)
    c = ops.gemm_bf16_nt_mn(a, b, split_k=split)
    torch.cuda.synchronize()
    got = c.sum(0).double().cpu().numpy()
    ref = (a.double().t() @ b.double()).cpu().numpy()
    assert rel_err(got, ref) < 2e-5
