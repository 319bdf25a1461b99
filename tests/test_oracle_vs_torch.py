"""Pins the oracle with PyTorch CPU fp32 (independent implementation) and finite differences.

The reference ships no golden vectors, so the oracle is pinned here before any
GPU result is compared with it.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pyoracle as po
from tests import torch_ref

f32 = np.float32
RNG = np.random.default_rng(1234)


def rnd(*shape, scale=1.0):
    return (RNG.standard_normal(shape) * scale).astype(f32)


@pytest.mark.parametrize("b,t,c,oc,bias", [(2, 3, 4, 5, True), (3, 7, 32, 48, True), (2, 5, 16, 8, False)])
def test_matmul_fwd_bwd(b, t, c, oc, bias):
    inp, w, dout = rnd(b * t, c), rnd(oc, c), rnd(b * t, oc)
    bv = rnd(oc) if bias else None
    out = np.zeros((b * t, oc), f32)
    po.matmul_forward(out, inp, w, bv, b, t, c, oc)
    ti, tw = torch.tensor(inp, requires_grad=True), torch.tensor(w, requires_grad=True)
    tb = torch.tensor(bv, requires_grad=True) if bias else None
    tout = F.linear(ti, tw, tb)
    assert np.allclose(out, tout.detach().numpy(), rtol=1e-5, atol=1e-5)
    tout.backward(torch.tensor(dout))
    dinp, dw = np.zeros_like(inp), np.zeros_like(w)
    db = np.zeros(oc, f32) if bias else None
    po.matmul_backward(dinp, dw, db, dout, inp, w, b, t, c, oc)
    assert np.allclose(dinp, ti.grad.numpy(), rtol=1e-5, atol=1e-5)
    assert np.allclose(dw, tw.grad.numpy(), rtol=1e-5, atol=1e-5)
    if bias:
        assert np.allclose(db, tb.grad.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("b,t,c", [(2, 3, 4), (4, 9, 64), (1, 5, 192)])
def test_layernorm_fwd_bwd(b, t, c):
    x, w, bias, dout = rnd(b * t, c), rnd(c), rnd(c), rnd(b * t, c)
    out, mean, rstd = np.zeros_like(x), np.zeros(b * t, f32), np.zeros(b * t, f32)
    po.layernorm_forward(out, mean, rstd, x, w, bias, b, t, c)
    tx, tw, tb = (torch.tensor(a, requires_grad=True) for a in (x, w, bias))
    tout = F.layer_norm(tx, (c,), tw, tb, 1e-5)
    assert np.allclose(out, tout.detach().numpy(), rtol=1e-5, atol=1e-5)
    assert np.allclose(mean, x.mean(axis=1), atol=1e-6)
    assert np.allclose(rstd, 1.0 / np.sqrt(x.var(axis=1) + 1e-5), rtol=1e-5)
    tout.backward(torch.tensor(dout))
    dx, dw, db = np.zeros_like(x), np.zeros_like(w), np.zeros_like(bias)
    po.layernorm_backward(dx, dw, db, dout, x, w, mean, rstd, b, t, c)
    assert np.allclose(dx, tx.grad.numpy(), rtol=1e-4, atol=1e-5)
    assert np.allclose(dw, tw.grad.numpy(), rtol=1e-4, atol=1e-5)
    assert np.allclose(db, tb.grad.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("causal", [0, 1])
@pytest.mark.parametrize("b,t,c,nh", [(2, 3, 4, 2), (2, 17, 64, 4), (1, 65, 64, 1)])
def test_attention_fwd_bwd(b, t, c, nh, causal):
    hs = c // nh
    qkv, dout = rnd(b, t, 3 * c), rnd(b, t, c)
    out = np.zeros((b, t, c), f32)
    preatt, att = np.zeros((b, nh, t, t), f32), np.zeros((b, nh, t, t), f32)
    po.attention_forward(out, preatt, att, qkv, b, t, c, nh, causal)
    tq = torch.tensor(qkv, requires_grad=True)
    q, k, v = tq.split(c, dim=2)
    q, k, v = (z.view(b, t, nh, hs).transpose(1, 2) for z in (q, k, v))
    ty = F.scaled_dot_product_attention(q, k, v, is_causal=bool(causal)).transpose(1, 2).reshape(b, t, c)
    assert np.allclose(out, ty.detach().numpy(), rtol=1e-5, atol=1e-5)
    assert np.allclose(att.sum(axis=-1), 1.0, atol=1e-5)
    ty.backward(torch.tensor(dout))
    dqkv = np.zeros_like(qkv)
    dpre, datt = np.zeros_like(preatt), np.zeros_like(att)
    po.attention_backward(dqkv, dpre, datt, dout, qkv, att, b, t, c, nh, causal)
    assert np.allclose(dqkv, tq.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_gelu_fwd_bwd_and_finite_difference():
    x = np.concatenate([rnd(1000, scale=2.0).ravel(), np.array([-2, -1, -0.5, 0, 0.5, 1, 2], f32)])
    n = x.size
    out = np.zeros(n, f32)
    po.gelu_forward(out, x, n)
    tx = torch.tensor(x, requires_grad=True)
    ty = F.gelu(tx, approximate="tanh")
    assert np.allclose(out, ty.detach().numpy(), rtol=1e-5, atol=1e-6)
    dout = rnd(n)
    ty.backward(torch.tensor(dout))
    dx = np.zeros(n, f32)
    po.gelu_backward(dx, x, dout, n)
    assert np.allclose(dx, tx.grad.numpy(), rtol=1e-4, atol=1e-5)
    # central finite difference of the oracle's own forward in float64 (Q3 evidence)
    x64 = x.astype(np.float64)
    g = lambda z: 0.5 * z * (1 + np.tanh(np.sqrt(2 / np.pi) * (z + 0.044715 * z ** 3)))
    fd = (g(x64 + 1e-6) - g(x64 - 1e-6)) / 2e-6
    ones = np.ones(n, f32)
    dx1 = np.zeros(n, f32)
    po.gelu_backward(dx1, x, ones, n)
    assert np.allclose(dx1, fd, rtol=1e-4, atol=1e-5)


def test_softmax_crossentropy_fwd_bwd():
    b, v = 6, 10
    logits = rnd(b, v, scale=3.0)
    targets = RNG.integers(0, v, b).astype(np.int32)
    probs, losses = np.zeros((b, v), f32), np.zeros(b, f32)
    po.softmax_forward(probs, logits, b, 1, v)
    po.crossentropy_forward(losses, probs, targets, b, 1, v)
    tl = torch.tensor(logits, requires_grad=True)
    tlosses = F.cross_entropy(tl, torch.tensor(targets, dtype=torch.long), reduction="none")
    assert np.allclose(probs, F.softmax(tl, dim=1).detach().numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(losses, tlosses.detach().numpy(), rtol=1e-5, atol=1e-6)
    tlosses.mean().backward()
    dlogits = np.zeros((b, v), f32)
    po.crossentropy_softmax_backward(dlogits, np.full(b, 1.0 / b, f32), probs, targets, b, 1, v)
    assert np.allclose(dlogits, tl.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_patch_embed_fwd_bwd():
    b, img, patch, c = 2, 8, 4, 6
    g = img // patch
    t = g * g + 1
    images = rnd(b, 3, img, img)
    pw, pb, cls, wpe = rnd(c, 3 * patch * patch), rnd(c), rnd(c), rnd(t, c)
    enc = np.zeros((b, t, c), f32)
    po.patch_embed_forward(enc, images, pw, pb, cls, wpe, b, img, patch, c)
    tpw, tpb, tcls, twpe = (torch.tensor(a, requires_grad=True) for a in (pw, pb, cls, wpe))
    tok = F.conv2d(torch.tensor(images), tpw.view(c, 3, patch, patch), tpb, stride=patch).flatten(2).transpose(1, 2)
    tenc = torch.cat([tcls.view(1, 1, c).expand(b, 1, c), tok], dim=1) + twpe.view(1, t, c)
    assert np.allclose(enc, tenc.detach().numpy(), rtol=1e-5, atol=1e-5)
    denc = rnd(b, t, c)
    tenc.backward(torch.tensor(denc))
    dpw, dpb, dcls, dwpe = np.zeros_like(pw), np.zeros_like(pb), np.zeros_like(cls), np.zeros_like(wpe)
    po.patch_embed_backward(dpw, dpb, dcls, dwpe, denc, images, b, img, patch, c)
    for got, want in ((dpw, tpw), (dpb, tpb), (dcls, tcls), (dwpe, twpe)):
        assert np.allclose(got, want.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_encoder_fwd_bwd():
    b, t, c, v = 2, 5, 8, 11
    wte, wpe = rnd(v, c), rnd(t, c)
    inputs = RNG.integers(0, v, (b, t)).astype(np.int32)
    enc = np.zeros((b, t, c), f32)
    po.encoder_forward(enc, inputs, wte, wpe, b, t, c)
    assert np.allclose(enc, wte[inputs] + wpe[None])
    denc = rnd(b, t, c)
    dwte, dwpe = np.zeros_like(wte), np.zeros_like(wpe)
    po.encoder_backward(dwte, dwpe, denc, inputs, b, t, c)
    want = np.zeros_like(wte)
    np.add.at(want, inputs.ravel(), denc.reshape(-1, c))
    assert np.allclose(dwte, want, atol=1e-6) and np.allclose(dwpe, denc.sum(axis=0), atol=1e-6)


def test_adamw_matches_torch():
    n = 1000
    p0, g = rnd(n), rnd(n)
    p, m, v = p0.copy(), np.zeros(n, f32), np.zeros(n, f32)
    tp = torch.tensor(p0, requires_grad=True)
    opt = torch.optim.AdamW([tp], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    for step in range(1, 6):
        gs = (g * step).astype(f32)
        po.adamw_step(p, gs, m, v, 1e-2, 0.9, 0.999, 1e-8, 0.01, step)
        tp.grad = torch.tensor(gs)
        opt.step()
    assert np.allclose(p, tp.detach().numpy(), rtol=1e-5, atol=1e-6)
    # the reference's SGD (train_vit.rs:737-743)
    q = p0.copy()
    po.sgd_step(q, g, 0.1)
    assert np.allclose(q, p0 - np.float32(0.1) * g)


@pytest.mark.parametrize("cfg_name,b,causal", [("tiny", 4, 0), ("tiny", 2, 1)])
def test_model_matches_torch(cfg_name, b, causal):
    cfg = po.CONFIGS[cfg_name]
    m = po.ViT(cfg_name, seed=1337, causal=causal)
    images, labels = po.synthetic_batch(cfg, b)
    loss = m.forward(images, labels)
    m.zero_grad()
    m.backward()
    p = torch_ref.params_from_flat(m.params_flat(), cfg, m.param_sizes, po.PARAM_NAMES)
    logits, tloss, _ = torch_ref.forward(p, cfg, images, labels, causal=bool(causal))
    tloss.backward()
    assert abs(loss - tloss.item()) < 1e-5 * max(1.0, abs(tloss.item()))
    assert np.allclose(m.act("logits").reshape(b, -1), logits.detach().numpy(), rtol=1e-4, atol=1e-6)
    for name in po.PARAM_NAMES:
        got, want = m.grad(name), p[name].grad.numpy().ravel()
        denom = np.abs(want).max() + 1e-12
        assert np.abs(got - want).max() / denom < 1e-4, name


# The shapes the GPU parity tests hold the production kernels to (tests/test_gpu_parity_configs.py): 64-wide heads at T = 197
# and T = 785.  The oracle is the reference there, so it is pinned there too.
TI16_2L = dict(image_size=224, patch_size=16, channels=192, num_layers=2, num_heads=3, num_classes=16)
B8_1L = dict(image_size=224, patch_size=8, channels=128, num_layers=1, num_heads=2, num_classes=16)
MID_HS64 = dict(image_size=64, patch_size=16, channels=256, num_layers=2, num_heads=4, num_classes=16)


@pytest.mark.parametrize("cfg,b,init_mode,tol", [(TI16_2L, 2, 1, 1e-4), (B8_1L, 1, 1, 1e-4), (MID_HS64, 4, 0, 5e-4), (TI16_2L, 2, 0, 5e-4)],
                         ids=["t197-hs64", "t785-hs64", "mid-refinit", "t197-refinit"])
def test_model_matches_torch_on_parity_shapes(cfg, b, init_mode, tol):
    """Loss, logits and every parameter gradient against PyTorch CPU fp32.  init_mode 0 is the reference's all-positive init
    (rusty_vit.rs:864-903), where two correct fp32 evaluations differ by up to ~1e-4 (DEVIATIONS D15): held to 5e-4."""
    m = po.ViT(cfg, seed=1337, init_mode=init_mode)
    images, labels = po.synthetic_batch(cfg, b)
    loss = m.forward(images, labels)
    m.zero_grad()
    m.backward()
    p = torch_ref.params_from_flat(m.params_flat(), cfg, m.param_sizes, po.PARAM_NAMES)
    logits, tloss, _ = torch_ref.forward(p, cfg, images, labels)
    tloss.backward()
    assert abs(loss - tloss.item()) < 1e-5 * max(1.0, abs(tloss.item()))
    want = logits.detach().numpy()
    assert np.abs(m.act("logits").reshape(b, -1) - want).max() <= tol * np.abs(want).max()
    for name in po.PARAM_NAMES:
        got, want = m.grad(name), p[name].grad.numpy().ravel()
        assert np.abs(got - want).max() <= tol * (np.abs(want).max() + 1e-12), name


def test_model_training_curve_matches_torch():
    """5 AdamW steps: oracle and torch produce the same loss sequence."""
    cfg = po.CONFIGS["tiny"]
    m = po.ViT("tiny")
    p = torch_ref.params_from_flat(m.params_flat(), cfg, m.param_sizes, po.PARAM_NAMES)
    flat = torch.cat([p[n].detach().ravel() for n in po.PARAM_NAMES]).requires_grad_(True)
    opt = torch.optim.AdamW([flat], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    for step in range(5):
        images, labels = po.synthetic_batch(cfg, 8, step=step)
        lo = m.forward(images, labels)
        m.zero_grad(); m.backward(); m.update(1e-3)
        off, pp = 0, {}
        for n, s in zip(po.PARAM_NAMES, m.param_sizes):
            pp[n] = flat[off:off + s]; off += s
        _, tl, _ = torch_ref.forward(pp, cfg, images, labels)
        opt.zero_grad(); tl.backward(); opt.step()
        assert abs(lo - tl.item()) < 2e-5 * max(1.0, tl.item()), (step, lo, tl.item())


def test_model_gradient_finite_difference():
    """Whole-model directional derivative in float64-ish: loss(p + h d) - loss(p - h d)."""
    cfg = po.CONFIGS["tiny"]
    m = po.ViT("tiny")
    images, labels = po.synthetic_batch(cfg, 2)
    m.forward(images, labels); m.zero_grad(); m.backward()
    g = m.grads_flat().copy()
    d = RNG.standard_normal(g.size).astype(f32)
    d /= np.linalg.norm(d)
    p0 = m.params_flat().copy()
    h = 1e-2
    m.params_flat()[:] = p0 + h * d
    lp = m.forward(images, labels)
    m.params_flat()[:] = p0 - h * d
    lm = m.forward(images, labels)
    fd = (lp - lm) / (2 * h)
    an = float(np.dot(g.astype(np.float64), d.astype(np.float64)))
    assert abs(fd - an) < 2e-2 * max(abs(an), 1e-3), (fd, an)


def test_dloss_scale_data_parallel_identity():
    """Two half-batches with dloss = 1/B_global sum to the full-batch gradient."""
    cfg = po.CONFIGS["tiny"]
    images, labels = po.synthetic_batch(cfg, 4)
    full = po.ViT("tiny")
    full.forward(images, labels); full.zero_grad(); full.backward()
    acc = np.zeros(full.num_parameters, np.float64)
    for r in range(2):
        part = po.ViT("tiny")
        part.set_dloss_scale(1.0 / 4)
        part.forward(images[2 * r:2 * r + 2], labels[2 * r:2 * r + 2])
        part.zero_grad(); part.backward()
        acc += part.grads_flat()
    assert np.allclose(acc, full.grads_flat(), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("cfg", [MID_HS64, dict(image_size=32, patch_size=8, channels=128, num_layers=2, num_heads=2, num_classes=10)],
                         ids=["c256", "c128"])
def test_reference_init_is_ill_conditioned_in_fp32(cfg):
    """DEVIATIONS D15, measured: against a float64 evaluation of the same step, the oracle's own fp32 rounding error is ~1e-6
    under the symmetric init (init_mode 1) and 30-80x that under the reference's all-positive init U[0,1)*0.02
    (rusty_vit.rs:864-903, init_mode 0), where every matmul output carries a large component common to all channels that the
    next LayerNorm cancels.  Two correct fp32 implementations therefore cannot be held to 1e-4 of each other there: the fp32
    parity gates use 1e-4 under init_mode 1 and 5e-4 on the cancellation-prone tensors under init_mode 0."""
    def worst_error(init_mode):
        m = po.ViT(cfg, seed=1337, init_mode=init_mode)
        images, labels = po.synthetic_batch(cfg, 4)
        m.forward(images, labels); m.zero_grad(); m.backward()
        p32 = torch_ref.params_from_flat(m.params_flat(), cfg, m.param_sizes, po.PARAM_NAMES)
        p64 = {k: v.detach().double().requires_grad_(True) for k, v in p32.items()}
        logits, loss, _ = torch_ref.forward(p64, cfg, images.astype(np.float64), labels)
        loss.backward()
        rel = lambda a, b: np.abs(np.asarray(a, np.float64).ravel() - b.ravel()).max() / max(np.abs(b).max(), 1e-30)
        errs = [rel(m.act("logits"), logits.detach().numpy())] + [rel(m.grad(n), p64[n].grad.numpy()) for n in po.PARAM_NAMES]
        return max(errs)
    symmetric, reference = worst_error(1), worst_error(0)
    assert symmetric <= 5e-6, symmetric
    assert reference <= 5e-4, reference
    assert reference >= 10 * symmetric, (reference, symmetric)


@pytest.mark.parametrize("cfg,b,floor_logits,floor_worst", [(MID_HS64, 16, 2.1e-2, 7.9e-2), ("ti16", 4, 5.9e-2, 1.2e-1)], ids=["c256", "ti16"])
def test_bf16_storage_floor_under_the_reference_init(cfg, b, floor_logits, floor_worst):
    """DEVIATIONS D14, measured: an IDEAL bf16-storage evaluation of the step (exact fp32 arithmetic, values rounded only where
    the production mode stores bf16: torch_ref.forward_bf16_storage) is already 2e-2 .. 1.2e-1 away from the fp32 oracle under
    the reference's all-positive init, on logits and the LayerNorm gradients — no bf16 implementation can meet 2e-2 there — while
    under the symmetric init the same evaluation stays inside 2e-2 on every tensor.  The GPU parity tests
    (tests/test_gpu_parity_configs.py) therefore hold init_mode 1 to 2e-2 and init_mode 0 to max(2e-2, 3 x this floor)."""
    c = po.CONFIGS[cfg] if isinstance(cfg, str) else cfg

    def floors(init_mode):
        ref = po.ViT(cfg, seed=1337, init_mode=init_mode)
        images, labels = po.synthetic_batch(c, b)
        ref.forward(images, labels); ref.zero_grad(); ref.backward()
        p = torch_ref.params_from_flat(ref.params_flat(), c, ref.param_sizes, po.PARAM_NAMES)
        logits, loss, _ = torch_ref.forward_bf16_storage(p, c, images, labels)
        loss.backward()
        rel = lambda a, w: np.abs(np.asarray(a, np.float64).ravel() - np.asarray(w, np.float64).ravel()).max() / np.abs(w).max()
        out = {"logits": rel(logits.detach().numpy(), ref.act("logits"))}
        out.update({"d" + n: rel(p[n].grad.numpy(), ref.grad(n)) for n in po.PARAM_NAMES})
        return out
    sym, refinit = floors(1), floors(0)
    assert max(sym.values()) <= 2e-2, sym
    assert abs(refinit["logits"] - floor_logits) <= 0.1 * floor_logits, refinit["logits"]
    assert abs(max(refinit.values()) - floor_worst) <= 0.1 * floor_worst, max(refinit.values())
    assert max(refinit, key=refinit.get) in ("dln1w", "dln2w")
