"""The tcgen05 / TMEM / TMA GEMM (csrc/gemm_tc.cu) through vitrs_gemm_bf16 against fp32 matmul.

bf16 x bf16 products are exact in fp32, so the only difference from the fp32 reference of the
same (bf16-rounded) operands is accumulation order: the fp32-output path must agree to ~1e-5
relative, the bf16-output path to one bf16 rounding (2^-8).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (128, 256, 128), (256, 512, 768), (200, 192, 192), (197 * 3, 576, 192), (1000, 768, 3072),
          (130, 64, 256), (64, 48, 48), (1576, 2304, 768), (3072, 768, 4000), (8, 8, 8), (129, 264, 72)]
LAYOUTS = [(0, 0), (0, 1), (1, 1), (1, 0)]


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_bf16_out(vitrs, M, N, K, a_mn, b_mn):
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operands need an extent that is a multiple of 8 (TMA 16-byte rule)")
    g = torch.Generator(device="cuda").manual_seed(M * 31 + N * 7 + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    want = A.float() @ B.float().t()
    Am = A.t().contiguous() if a_mn else A  # MN-major: stored [K, M]
    Bm = B.t().contiguous() if b_mn else B
    D = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    vitrs.gemm_bf16(D, Am, Bm, M, N, K, M if a_mn else K, N if b_mn else K, N, a_mn, b_mn, 0)
    torch.cuda.synchronize()
    err = (D.float() - want).abs().max().item() / want.abs().max().item()
    assert err <= 2.0 ** -7, err


@pytest.mark.parametrize("a_mn,b_mn", [(1, 1), (0, 0)])
@pytest.mark.parametrize("M,N,K", [(768, 192, 197 * 16), (2304, 768, 25216), (192, 768, 1000), (64, 48, 520), (3072, 768, 8192)])
def test_gemm_f32_accumulate(vitrs, M, N, K, a_mn, b_mn):
    """dweight-style: fp32 output, added into (split-K with vector reductions)."""
    g = torch.Generator(device="cuda").manual_seed(K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    D0 = torch.randn(M, N, device="cuda", generator=g)
    want = D0.double() + A.double() @ B.double().t()
    Am = A.t().contiguous() if a_mn else A
    Bm = B.t().contiguous() if b_mn else B
    D = D0.clone()
    vitrs.gemm_bf16(D, Am, Bm, M, N, K, M if a_mn else K, N if b_mn else K, N, a_mn, b_mn, 1)
    torch.cuda.synchronize()
    err = (D.double() - want).abs().max().item() / want.abs().max().item()
    assert err <= 2e-5, err
