"""GPU operators through the C ABI against the CPU oracle.

First the reference's own op tests (tests/vit_tests.rs) verbatim through libvitrs.so, then
random-input parity of every op in fp32 verify mode (<= 1e-4 relative, BASELINE.json) and in
bf16 production mode (<= 2e-2 relative).  "relative" = max|a-b| / max|b| over the tensor.
"""
import numpy as np
import pytest
import torch

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
f32 = np.float32
TOL_F32 = 1e-4   # north_star: fp32 verify mode within 1e-4 relative
TOL_BF16 = 2e-2  # north_star: bf16 production mode within 2e-2


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def host(t):
    return t.float().cpu().numpy()


def relerr(got, want):
    want = np.asarray(want, np.float64)
    got = np.asarray(got, np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


# ---- the reference's own cases (tests/vit_tests.rs) through the C ABI ---------------------------
def test_ref_residual_forward(vitrs):
    n = 10  # vit_tests.rs:92-101
    out = torch.zeros(n, device="cuda")
    vitrs.residual_forward(out, torch.full((n,), 1.0, device="cuda"), torch.full((n,), 2.0, device="cuda"), n)
    assert torch.equal(out.cpu(), torch.full((n,), 3.0))


def test_ref_matmul_forward(vitrs):
    b, t, c, oc = 2, 3, 4, 5  # vit_tests.rs:104-132 (true value 11.0)
    out = torch.zeros(b * t * oc, device="cuda")
    inp, w, bias = (torch.full((n,), v, device="cuda") for n, v in ((b * t * c, 1.0), (oc * c, 2.0), (oc, 3.0)))
    vitrs.matmul_forward(out, inp, w, bias, b, t, c, oc)
    assert torch.equal(out.cpu(), torch.full((b * t * oc,), 11.0))
    vitrs.matmul_forward(out, inp, w, None, b, t, c, oc)  # null bias (train_vit.rs:388)
    assert torch.equal(out.cpu(), torch.full((b * t * oc,), 8.0))


def test_ref_attention_forward(vitrs):
    b, t, c, nh = 2, 3, 4, 2  # vit_tests.rs:135-160
    inp = torch.ones(b * t * 3 * c, device="cuda")
    out = torch.zeros(b * t * c, device="cuda")
    preatt = torch.zeros(b * nh * t * t, device="cuda")
    att = torch.zeros(b * nh * t * t, device="cuda")
    vitrs.attention_forward(out, preatt, att, inp, b, t, c, nh, causal=1)
    assert np.allclose(host(out), 1.0, atol=1e-6)
    hs = c // nh
    pre, a = host(preatt).reshape(b, nh, t, t), host(att).reshape(b, nh, t, t)
    for tq in range(t):
        assert np.allclose(pre[:, :, tq, :tq + 1], hs / np.sqrt(hs))
        assert np.allclose(a[:, :, tq, :tq + 1], 1.0 / (tq + 1), atol=1e-6)
        assert np.all(a[:, :, tq, tq + 1:] == 0)


def test_ref_layernorm_forward(vitrs):
    b, t, c = 2, 3, 4  # vit_tests.rs:163-190
    out, mean, rstd = (torch.zeros(n, device="cuda") for n in (b * t * c, b * t, b * t))
    vitrs.layernorm_forward(out, mean, rstd, torch.ones(b * t * c, device="cuda"), torch.full((c,), 2.0, device="cuda"),
                            torch.full((c,), 3.0, device="cuda"), b, t, c)
    assert np.array_equal(host(out), np.full(b * t * c, 3.0, f32))
    assert np.array_equal(host(mean), np.ones(b * t, f32))
    assert np.allclose(host(rstd), 1.0 / np.sqrt(1e-5), rtol=1e-6)


def test_ref_gelu_forward(vitrs):
    out = torch.zeros(10, device="cuda")  # vit_tests.rs:193-201
    vitrs.gelu_forward(out, torch.ones(10, device="cuda"), 10)
    assert np.allclose(host(out), 0.841192, atol=1e-6)


def test_ref_softmax_forward(vitrs):
    b, t, v = 2, 3, 4  # vit_tests.rs:204-230
    probs = torch.zeros(b * t * v, device="cuda")
    vitrs.softmax_forward(probs, torch.ones(b * t * v, device="cuda"), b, t, v)
    rows = host(probs).reshape(b * t, v)
    assert np.all(np.abs(rows.sum(axis=1) - 1.0) < 1e-6) and np.allclose(rows, 0.25)


# ---- random-input parity, both modes -------------------------------------------------------------
MODES = [("f32", torch.float32, TOL_F32), ("bf16", torch.bfloat16, TOL_BF16)]


def prep(a, dtype):
    """Host array as the op will see it (rounded to bf16 in production mode) and its device copy."""
    h = bf16_round(a) if dtype == torch.bfloat16 else np.ascontiguousarray(a, f32)
    return h, dev(h, dtype)


@pytest.mark.parametrize("mode,dtype,tol", MODES)
@pytest.mark.parametrize("n", [10, 4096 + 3, 1 << 20])
def test_residual_and_gelu(vitrs, mode, dtype, tol, n):
    rng = np.random.default_rng(n)
    (ha, a), (hb, b_), (hg, g) = (prep(rng.standard_normal(n).astype(f32) * 2, dtype) for _ in range(3))
    out = torch.zeros(n, device="cuda", dtype=dtype)
    vitrs.residual_forward(out, a, b_, n)
    want = np.zeros(n, f32); po.residual_forward(want, ha, hb, n)
    assert relerr(host(out), want) <= tol
    vitrs.gelu_forward(out, a, n)
    po.gelu_forward(want, ha, n)
    assert relerr(host(out), want) <= tol
    hd, d = prep(rng.standard_normal(n).astype(f32), dtype)
    wantd = hd.copy(); po.gelu_backward(wantd, ha, hg, n)  # accumulates
    vitrs.gelu_backward(d, a, g, n)
    assert relerr(host(d), wantd) <= tol
    hd1, d1 = prep(rng.standard_normal(n).astype(f32), dtype)
    hd2, d2 = prep(rng.standard_normal(n).astype(f32), dtype)
    w1, w2 = hd1.copy(), hd2.copy(); po.residual_backward(w1, w2, hg, n)
    vitrs.residual_backward(d1, d2, g, n)
    assert relerr(host(d1), w1) <= tol and relerr(host(d2), w2) <= tol


@pytest.mark.parametrize("mode,dtype,tol", MODES)
@pytest.mark.parametrize("b,t,c", [(2, 3, 4), (2, 65, 64), (3, 197, 192), (1, 50, 768), (2, 7, 1000)])
def test_layernorm(vitrs, mode, dtype, tol, b, t, c):
    rng = np.random.default_rng(c)
    hx, x = prep(rng.standard_normal(b * t * c).astype(f32) * 1.5 + 0.3, dtype)
    w = (1 + 0.1 * rng.standard_normal(c)).astype(f32); bias = (0.1 * rng.standard_normal(c)).astype(f32)
    out = torch.zeros(b * t * c, device="cuda", dtype=dtype)
    mean, rstd = torch.zeros(b * t, device="cuda"), torch.zeros(b * t, device="cuda")
    vitrs.layernorm_forward(out, mean, rstd, x, dev(w), dev(bias), b, t, c)
    wo, wm, wr = np.zeros(b * t * c, f32), np.zeros(b * t, f32), np.zeros(b * t, f32)
    po.layernorm_forward(wo, wm, wr, hx, w, bias, b, t, c)
    assert relerr(host(out), wo) <= tol and relerr(host(mean), wm) <= 1e-5 and relerr(host(rstd), wr) <= 1e-5
    # backward accumulates into dinp / dweight / dbias (train_vit.rs:626-633)
    hg, g = prep(rng.standard_normal(b * t * c).astype(f32), dtype)
    hdx, dx = prep(rng.standard_normal(b * t * c).astype(f32) * 0.1, dtype)
    dw0, db0 = rng.standard_normal(c).astype(f32), rng.standard_normal(c).astype(f32)
    dw, db = dev(dw0), dev(db0)
    vitrs.layernorm_backward(dx, dw, db, g, x, dev(w), mean, rstd, b, t, c)
    wdx, wdw, wdb = hdx.copy(), dw0.copy(), db0.copy()
    po.layernorm_backward(wdx, wdw, wdb, hg, hx, w, wm, wr, b, t, c)
    assert relerr(host(dx), wdx) <= tol
    assert relerr(host(dw), wdw) <= 1e-4 * (50 if dtype == torch.bfloat16 else 1)
    assert relerr(host(db), wdb) <= 1e-4


@pytest.mark.parametrize("mode,dtype,tol", MODES)
@pytest.mark.parametrize("b,t,c,oc", [(2, 3, 4, 5), (2, 65, 64, 192), (1, 197, 192, 576), (2, 130, 256, 64), (1, 300, 768, 3072)])
def test_matmul(vitrs, mode, dtype, tol, b, t, c, oc):
    rng = np.random.default_rng(oc)
    hx, x = prep(rng.standard_normal(b * t * c).astype(f32), dtype)
    hw, w = prep(rng.standard_normal(oc * c).astype(f32) * 0.05, dtype)
    bias = rng.standard_normal(oc).astype(f32)
    out = torch.zeros(b * t * oc, device="cuda", dtype=dtype)
    vitrs.matmul_forward(out, x, w, dev(bias), b, t, c, oc)
    want = np.zeros(b * t * oc, f32); po.matmul_forward(want, hx, hw, bias, b, t, c, oc)
    assert relerr(host(out), want) <= tol
    vitrs.matmul_forward(out, x, w, None, b, t, c, oc)
    po.matmul_forward(want, hx, hw, None, b, t, c, oc)
    assert relerr(host(out), want) <= tol
    # backward: all three outputs accumulate; dbias may be NULL (train_vit.rs:548)
    hg, g = prep(rng.standard_normal(b * t * oc).astype(f32), dtype)
    hdx, dx = prep(rng.standard_normal(b * t * c).astype(f32), dtype)
    dw0, db0 = rng.standard_normal(oc * c).astype(f32), rng.standard_normal(oc).astype(f32)
    dw, db = dev(dw0), dev(db0)
    vitrs.matmul_backward(dx, dw, db, g, x, w, b, t, c, oc)
    wdx, wdw, wdb = hdx.copy(), dw0.copy(), db0.copy()
    po.matmul_backward(wdx, wdw, wdb, hg, hx, hw, b, t, c, oc)
    assert relerr(host(dx), wdx) <= tol
    assert relerr(host(dw), wdw) <= (1e-4 if dtype == torch.float32 else 2e-3)  # fp32 accumulation of bf16 products
    assert relerr(host(db), wdb) <= 1e-4
    dw2 = dev(dw0)
    vitrs.matmul_backward(None, dw2, None, g, x, w, b, t, c, oc)
    assert relerr(host(dw2), wdw) <= (1e-4 if dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("causal", [0, 1])
@pytest.mark.parametrize("b,t,c,nh", [(2, 3, 4, 2), (2, 65, 64, 4), (1, 197, 192, 3), (2, 40, 256, 2)])
def test_attention_f32(vitrs, causal, b, t, c, nh):
    rng = np.random.default_rng(t)
    qkv = rng.standard_normal(b * t * 3 * c).astype(f32)
    out = torch.zeros(b * t * c, device="cuda")
    preatt, att = torch.zeros(b * nh * t * t, device="cuda"), torch.zeros(b * nh * t * t, device="cuda")
    vitrs.attention_forward(out, preatt, att, dev(qkv), b, t, c, nh, causal=causal)
    wo, wp, wa = np.zeros(b * t * c, f32), np.zeros(b * nh * t * t, f32), np.zeros(b * nh * t * t, f32)
    po.attention_forward(wo, wp, wa, qkv, b, t, c, nh, causal=causal)
    assert relerr(host(out), wo) <= TOL_F32 and relerr(host(preatt), wp) <= TOL_F32 and relerr(host(att), wa) <= TOL_F32
    dout = rng.standard_normal(b * t * c).astype(f32)
    dqkv0 = rng.standard_normal(b * t * 3 * c).astype(f32)
    dqkv, dpre, datt = dev(dqkv0), torch.zeros_like(preatt), torch.zeros_like(att)
    vitrs.attention_backward(dqkv, dpre, datt, dev(dout), dev(qkv), att, b, t, c, nh, causal=causal)
    wd, wdp, wda = dqkv0.copy(), np.zeros_like(wp), np.zeros_like(wa)
    po.attention_backward(wd, wdp, wda, dout, qkv, wa, b, t, c, nh, causal=causal)
    assert relerr(host(dqkv), wd) <= TOL_F32
    assert relerr(host(datt), wda) <= TOL_F32 and relerr(host(dpre), wdp) <= TOL_F32
    # NULL preatt/att/dpreatt/datt are legal at this ABI (the fused path keeps none of them)
    out2 = torch.zeros_like(out)
    vitrs.attention_forward(out2, None, None, dev(qkv), b, t, c, nh, causal=causal)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("causal", [0, 1])
@pytest.mark.parametrize("b,t,c,nh", [(2, 65, 64, 4), (2, 197, 192, 3), (1, 50, 128, 2), (1, 300, 64, 1), (2, 128, 128, 2), (1, 785, 128, 2),
                                      (3, 17, 64, 1), (1, 256, 64, 1), (1, 257, 64, 1)])
def test_attention_bf16(vitrs, causal, b, t, c, nh):
    """Fused (lse-only) attention: hs = 64 runs the tensor-core kernel, others the SIMT one."""
    rng = np.random.default_rng(t + c)
    hq, qkv = prep(rng.standard_normal(b * t * 3 * c).astype(f32), torch.bfloat16)
    out = torch.zeros(b * t * c, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(b * nh * t, device="cuda")
    vitrs.attention_forward(out, lse, None, qkv, b, t, c, nh, causal=causal)
    wo, wp, wa = np.zeros(b * t * c, f32), np.zeros(b * nh * t * t, f32), np.zeros(b * nh * t * t, f32)
    po.attention_forward(wo, wp, wa, hq, b, t, c, nh, causal=causal)
    assert relerr(host(out), wo) <= TOL_BF16
    pre = wp.reshape(b * nh, t, t).astype(np.float64)
    if causal:
        pre = np.where(np.tril(np.ones((t, t), bool)), pre, -np.inf)
    want_lse = np.log(np.exp(pre - pre.max(-1, keepdims=True)).sum(-1)) + pre.max(-1)
    assert np.abs(host(lse).reshape(b * nh, t) - want_lse).max() <= 2e-2
    hdo, dout = prep(rng.standard_normal(b * t * c).astype(f32), torch.bfloat16)
    dqkv = torch.zeros(b * t * 3 * c, device="cuda", dtype=torch.bfloat16)
    vitrs.attention_backward_bf16(dqkv, dout, out, lse, qkv, b, t, c, nh, causal=causal)
    wd, wdp, wda = np.zeros(b * t * 3 * c, f32), np.zeros_like(wp), np.zeros_like(wa)
    po.attention_backward(wd, wdp, wda, hdo, hq, wa, b, t, c, nh, causal=causal)
    assert relerr(host(dqkv), wd) <= TOL_BF16


@pytest.mark.parametrize("t", [16, 64, 128, 160, 192, 197, 256])
def test_attention_bf16_many_heads_per_cta(vitrs, t):
    """768 (batch, head) pairs: the persistent backward kernel walks ~5 heads per CTA (operand tiles refilled for the next head
    while the current one is still computing).  Reference: fp32 autograd of the same attention on the bf16-rounded inputs —
    the oracle's O(T^3) backward would take minutes here; its formula is pinned by test_attention_bf16 on small shapes."""
    b, c, nh = 64, 768, 12
    g = torch.Generator(device="cuda").manual_seed(t * 7 + b)
    qkv = (torch.randn(b, t, 3 * c, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    dout = (torch.randn(b, t, c, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    out = torch.zeros(b, t, c, device="cuda", dtype=torch.bfloat16)
    dqkv = torch.zeros(b, t, 3 * c, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(b * nh * t, device="cuda")
    vitrs.attention_forward(out, lse, None, qkv, b, t, c, nh, causal=0)
    vitrs.attention_backward_bf16(dqkv, dout, out, lse, qkv, b, t, c, nh, causal=0)
    torch.cuda.synchronize()
    x = qkv.float().requires_grad_(True)
    q, k, v = x.split(c, dim=2)
    q, k, v = (z.view(b, t, nh, 64).transpose(1, 2) for z in (q, k, v))
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    y = (att @ v).transpose(1, 2).reshape(b, t, c)
    y.backward(dout.float())
    assert (out.float() - y.detach()).abs().max().item() <= TOL_BF16 * y.abs().max().item()
    assert (dqkv.float() - x.grad).abs().max().item() <= TOL_BF16 * x.grad.abs().max().item()
    # the += contract of the ABI: a second call doubles the result (same rounding: bf16 sums of equal values are exact)
    first = dqkv.clone()
    vitrs.attention_backward_bf16(dqkv, dout, out, lse, qkv, b, t, c, nh, causal=0)
    assert (dqkv.float() - 2 * first.float()).abs().max().item() <= 2.0 ** -7 * first.float().abs().max().item()


def test_softmax_crossentropy(vitrs):
    b, t, v = 8, 1, 1000
    rng = np.random.default_rng(5)
    logits = (rng.standard_normal(b * t * v) * 3).astype(f32)
    targets = rng.integers(0, v, b * t).astype(np.int32)
    probs = torch.zeros(b * t * v, device="cuda")
    vitrs.softmax_forward(probs, dev(logits), b, t, v)
    wp = np.zeros(b * t * v, f32); po.softmax_forward(wp, logits, b, t, v)
    assert relerr(host(probs), wp) <= TOL_F32
    losses = torch.zeros(b * t, device="cuda")
    vitrs.crossentropy_forward(losses, probs, dev(targets), b, t, v)
    wl = np.zeros(b * t, f32); po.crossentropy_forward(wl, wp, targets, b, t, v)
    assert relerr(host(losses), wl) <= TOL_F32
    dl0 = rng.standard_normal(b * t * v).astype(f32)
    dlosses = np.full(b * t, 1.0 / (b * t), f32)
    dlog = dev(dl0)
    vitrs.crossentropy_softmax_backward(dlog, dev(dlosses), probs, dev(targets), b, t, v)
    wd = dl0.copy(); po.crossentropy_softmax_backward(wd, dlosses, wp, targets, b, t, v)
    assert relerr(host(dlog), wd) <= TOL_F32


def test_encoder_and_patch_embed(vitrs):
    rng = np.random.default_rng(9)
    b, t, c, vocab = 2, 5, 8, 11
    inputs = rng.integers(0, vocab, b * t).astype(np.int32)
    wte, wpe = rng.standard_normal(vocab * c).astype(f32), rng.standard_normal(t * c).astype(f32)
    enc = torch.zeros(b * t * c, device="cuda")
    vitrs.encoder_forward(enc, dev(inputs), dev(wte), dev(wpe), b, t, c)
    want = np.zeros(b * t * c, f32); po.encoder_forward(want, inputs, wte, wpe, b, t, c)
    assert np.array_equal(host(enc), want)
    denc = rng.standard_normal(b * t * c).astype(f32)
    dwte, dwpe = torch.zeros(vocab * c, device="cuda"), torch.zeros(t * c, device="cuda")
    vitrs.encoder_backward(dwte, dwpe, dev(denc), dev(inputs), b, t, c)
    w1, w2 = np.zeros(vocab * c, f32), np.zeros(t * c, f32); po.encoder_backward(w1, w2, denc, inputs, b, t, c)
    assert relerr(host(dwte), w1) <= 1e-6 and relerr(host(dwpe), w2) <= 1e-6

    for b, img, patch, c in [(2, 32, 4, 64), (3, 64, 16, 192), (1, 32, 8, 40)]:
        g = img // patch; t = g * g + 1; kdim = 3 * patch * patch
        images = rng.uniform(-1, 1, b * 3 * img * img).astype(f32)
        pw, pb = rng.standard_normal(c * kdim).astype(f32) * 0.05, rng.standard_normal(c).astype(f32)
        cls, wpe = rng.standard_normal(c).astype(f32), rng.standard_normal(t * c).astype(f32)
        enc = torch.zeros(b * t * c, device="cuda")
        vitrs.patch_embed_forward(enc, dev(images), dev(pw), dev(pb), dev(cls), dev(wpe), b, img, patch, c)
        want = np.zeros(b * t * c, f32); po.patch_embed_forward(want, images, pw, pb, cls, wpe, b, img, patch, c)
        assert relerr(host(enc), want) <= TOL_F32
        denc = rng.standard_normal(b * t * c).astype(f32)
        outs = [torch.zeros(n, device="cuda") for n in (c * kdim, c, c, t * c)]
        vitrs.patch_embed_backward(*outs, dev(denc), dev(images), b, img, patch, c)
        wants = [np.zeros(n, f32) for n in (c * kdim, c, c, t * c)]
        po.patch_embed_backward(*wants, denc, images, b, img, patch, c)
        for o, w_ in zip(outs, wants):
            assert relerr(host(o), w_) <= TOL_F32


def test_adamw_sgd_and_init(vitrs):
    rng = np.random.default_rng(3)
    n = 100003
    p0, g = rng.standard_normal(n).astype(f32), rng.standard_normal(n).astype(f32) * 0.1
    p, m, v = dev(p0), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    shadow = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    wp, wm, wv = p0.copy(), np.zeros(n, f32), np.zeros(n, f32)
    for step in (1, 2, 3):
        vitrs.adamw_step(p, dev(g), m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, shadow=shadow)
        po.adamw_step(wp, g, wm, wv, 1e-3, 0.9, 0.999, 1e-8, 0.01, step)
    assert np.abs(host(p) - wp).max() <= 1e-6 and relerr(host(m), wm) <= 1e-5 and relerr(host(v), wv) <= 1e-5
    assert torch.equal(shadow, p.to(torch.bfloat16))
    vitrs.sgd_step(p, dev(g), 0.1)
    po.sgd_step(wp, g, 0.1)
    assert np.abs(host(p) - wp).max() <= 1e-6
    # the counter generator is bit-identical to the oracle's (DEVIATIONS D9)
    u = torch.zeros(4097, device="cuda")
    vitrs.fill_uniform(u, 1337, 6, 0.0, 0.02)
    assert np.array_equal(host(u), po.fill_uniform(4097, 1337, 6, 0.0, 0.02))
    vitrs.fill_uniform(u, 1337, 1000, -1.0, 1.0)
    assert np.array_equal(host(u), po.fill_uniform(4097, 1337, 1000, -1.0, 1.0))


def test_empty_inputs_are_noops(vitrs):
    z = torch.zeros(0, device="cuda")
    one = torch.ones(4, device="cuda")
    vitrs.residual_forward(z, z, z, 0)
    vitrs.gelu_forward(z, z, 0)
    vitrs.matmul_forward(z, z, one, None, 0, 3, 2, 2)
    vitrs.layernorm_forward(z, z, z, z, one, one, 0, 3, 4)
    torch.cuda.synchronize()


@pytest.mark.parametrize("qa,ka", [(10.0, 12.0), (20.0, 24.0), (40.0, 40.0)])
@pytest.mark.parametrize("t,peak_key", [(197, 150), (256, 255), (160, 40), (197, 100), (197, 196), (144, 143), (300, 290), (417, 200)])
def test_attention_bf16_late_peak_moves_the_exponent_reference(vitrs, t, peak_key, qa, ka):
    """The persistent forward takes its exponent reference from the first 32 keys of each half of a row and never looks at a
    maximum again: only a chunk whose probabilities sum past 2^64 moves the reference and rescales what was written
    (attention_tc.cu, single-pass softmax).  Here one late key scores qa * ka / 8 nats above every other for every query:
    15 nats stays inside the range (probabilities up to 2^22 against the reference), 60 nats (2^87) takes the rescale path,
    200 nats overflows the exponential itself.  The peak sits in a full chunk of the second half, in the last (masked) chunk,
    in the first half, and in the first half's 16-column remainder; T > 256 runs the streaming kernel, where a late peak also
    rescales the O accumulator of the earlier key tiles.  The result must be the exact softmax (oracle)."""
    b, c, nh = 2, 128, 2
    rng = np.random.default_rng(t)
    x = (rng.standard_normal((b, t, 3, nh, 64)) * 0.5).astype(f32)
    u = np.full(64, 1.0 / 8.0, f32)                     # unit vector
    x[:, :, 0] += qa * u                                 # every query has a component qa along u
    x[:, peak_key, 1] = ka * u                           # one key has ka along u: q.k / sqrt(64) = qa * ka / 8
    hq, qkv = prep(x.reshape(-1), torch.bfloat16)
    out = torch.zeros(b * t * c, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(b * nh * t, device="cuda")
    vitrs.attention_forward(out, lse, None, qkv, b, t, c, nh, causal=0)
    wo, wp, wa = np.zeros(b * t * c, f32), np.zeros(b * nh * t * t, f32), np.zeros(b * nh * t * t, f32)
    po.attention_forward(wo, wp, wa, hq, b, t, c, nh, causal=0)
    assert relerr(host(out), wo) <= TOL_BF16
    pre = wp.reshape(b * nh, t, t).astype(np.float64)
    want_lse = np.log(np.exp(pre - pre.max(-1, keepdims=True)).sum(-1)) + pre.max(-1)
    assert np.abs(host(lse).reshape(b * nh, t) - want_lse).max() <= 2e-2 * max(1.0, qa * ka / 120.0)  # (lse itself grows with the peak)
    assert (wa.reshape(b * nh, t, t)[:, :, peak_key] > 0.9).all()  # the case is what it claims to be


def test_two_contexts_on_one_device_do_not_lower_each_others_shared_memory_opt_in(vitrs):
    """The dynamic shared memory opt-in belongs to (kernel, device), not to a context: a second context that launches the
    persistent attention backward with ONE key tile (less shared memory) must not leave the first context's two-key-tile
    launch without its opt-in (regression: 'launch -> invalid argument' in the full suite)."""
    c, nh = 128, 2
    other = vitrs.Context(0)

    def bwd(t, ctx):
        b = 2
        qkv = (torch.randn(b, t, 3 * c, device="cuda") * 0.5).to(torch.bfloat16)
        dout = (torch.randn(b, t, c, device="cuda") * 0.1).to(torch.bfloat16)
        out = torch.zeros(b, t, c, device="cuda", dtype=torch.bfloat16)
        dqkv = torch.zeros(b, t, 3 * c, device="cuda", dtype=torch.bfloat16)
        lse = torch.zeros(b * nh * t, device="cuda")
        vitrs.attention_forward(out, lse, None, qkv, b, t, c, nh, causal=0, ctx=ctx)
        vitrs.attention_backward_bf16(dqkv, dout, out, lse, qkv, b, t, c, nh, causal=0, ctx=ctx)
        torch.cuda.synchronize()
        assert torch.isfinite(dqkv.float()).all()

    bwd(197, None)    # default context: two key tiles
    bwd(64, other)    # second context: one key tile
    bwd(197, None)    # default context again
    bwd(197, other)
    other.close()
