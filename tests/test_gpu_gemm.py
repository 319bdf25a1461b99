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


def _gelu(x):
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))


def _gelu_grad(x):
    u = 0.7978845608028654 * (x + 0.044715 * x ** 3)
    th = torch.tanh(u)
    return 0.5 * (1.0 + th) + x * 0.5 * (1.0 - th * th) * 0.7978845608028654 * (1.0 + 3.0 * 0.044715 * x * x)


@pytest.mark.parametrize("epi", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("M,N,K,b_mn", [(256, 512, 768, 0), (200, 192, 192, 0), (129, 264, 72, 1), (1576, 3072, 768, 0), (1000, 768, 3072, 1),
                                        (64, 48, 48, 0), (130, 64, 256, 0)])
def test_gemm_fused_epilogues(vitrs, M, N, K, b_mn, epi):
    """Each fused epilogue of the training step (vitrs_gemm_bf16_fused) against the unfused fp32 op sequence
    (matmul_forward train_vit.rs:384, gelu_forward :482, residual_forward :376, gelu_backward :639)."""
    g = torch.Generator(device="cuda").manual_seed(M + N * 3 + K * 5 + epi)
    A = (torch.randn(M, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    aux = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    acc = A.float() @ B.float().t()
    Bm = B.t().contiguous() if b_mn else B
    D = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    D2 = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    vitrs.gemm_bf16_fused(D, D2 if epi == 2 else None, aux if epi in (3, 4) else None, bias if epi != 4 else None, None, A, Bm, M, N, K,
                          K, N if b_mn else K, N, 0, b_mn, epi)
    torch.cuda.synchronize()
    if epi == 1:
        want = acc + bias
    elif epi == 2:
        want = acc + bias
    elif epi == 3:
        want = acc + bias + aux.float()
    elif epi == 8:
        want = _gelu((acc + bias).to(torch.bfloat16).float())  # gelu_forward of the pre-activation as it would have been stored
    else:
        want = acc * _gelu_grad(aux.float())
    scale = want.abs().max().item()
    assert (D.float() - want).abs().max().item() <= 2.0 ** -7 * scale
    if epi == 2:
        # gelu_forward consumes the stored (bf16) pre-activation, as the unfused op would
        want2 = _gelu(D.float())
        assert (D2.float() - want2).abs().max().item() <= 2.0 ** -7 * max(want2.abs().max().item(), 1.0)
    else:
        assert (D2 == 7.0).all()


@pytest.mark.parametrize("M,N,K", [(768, 768, 1576), (2304, 768, 4000), (192, 256, 600)])
def test_gemm_fused_colsum(vitrs, M, N, K):
    """dweight-style GEMM (both operands MN-major) with the bias gradient riding along: a_colsum[m] += sum_k A(m,k) (tv:548-550)."""
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)   # dout^T
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    Am, Bm = A.t().contiguous(), B.t().contiguous()
    cs0 = torch.randn(M, device="cuda", generator=g)
    cs = cs0.clone()
    D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    vitrs.gemm_bf16_fused(D, None, None, None, cs, Am, Bm, M, N, K, M, N, N, 1, 1, 1)
    torch.cuda.synchronize()
    want = A.float() @ B.float().t()
    assert (D.float() - want).abs().max().item() <= 2.0 ** -7 * want.abs().max().item()
    want_cs = cs0.double() + A.double().sum(dim=1)
    assert (cs.double() - want_cs).abs().max().item() <= 1e-4 * max(want_cs.abs().max().item(), 1.0)
