"""The reference's own ten tests (tests/vit_tests.rs) carried over against the oracle.

Shapes and inputs are the reference's; where the reference only asserts "not all zero" the
true value is asserted instead, and the one wrong expectation (35.0, vit_tests.rs:126-131)
is replaced by the arithmetic answer 3 + 4*1*2 = 11.0.
"""
import numpy as np

from oracle import pyoracle as po

f32 = np.float32


def test_param_count_gpt2_shape():
    # vit_tests.rs:10-15 — the 16 reference tensor sizes (rusty_vit.rs:105-122) for the
    # GPT-2 124M shape sum to 124 439 808.  Recomputed from the reference's size rules.
    maxT, V, L, C = 1024, 50257, 12, 768
    sizes = [V * C, maxT * C, L * C, L * C, L * 3 * C * C, L * 3 * C, L * C * C, L * C, L * C, L * C,
             L * 4 * C * C, L * 4 * C, L * C * 4 * C, L * C, C, C]
    assert sum(sizes) == 124439808


def test_param_count_vit_b16():
    m = po.ViT("b16")
    # patchw 768*768 + patchb + cls + wpe 197*768 + 12 blocks + lnf + head
    C, L = 768, 12
    want = C * 768 + C + C + 197 * C + L * (2 * C + 3 * C * C + 3 * C + C * C + C + 2 * C + 4 * C * C + 4 * C + 4 * C * C + C) + 2 * C + 1000 * C + 1000
    assert m.num_parameters == want == 86567656


def test_residual_forward():
    # vit_tests.rs:92-101 — exact
    n = 10
    out = np.zeros(n, f32)
    po.residual_forward(out, np.full(n, 1.0, f32), np.full(n, 2.0, f32), n)
    assert np.array_equal(out, np.full(n, 3.0, f32))


def test_matmul_forward():
    # vit_tests.rs:104-132 — b2 t3 c4 oc5, inp 1, weight 2, bias 3 -> 11.0
    b, t, c, oc = 2, 3, 4, 5
    out = np.zeros(b * t * oc, f32)
    po.matmul_forward(out, np.ones(b * t * c, f32), np.full(oc * c, 2.0, f32), np.full(oc, 3.0, f32), b, t, c, oc)
    assert np.array_equal(out, np.full(b * t * oc, 11.0, f32))
    # null bias is legal (train_vit.rs:388)
    po.matmul_forward(out, np.ones(b * t * c, f32), np.full(oc * c, 2.0, f32), None, b, t, c, oc)
    assert np.array_equal(out, np.full(b * t * oc, 8.0, f32))


def test_attention_forward():
    # vit_tests.rs:135-160 — b2 t3 c4 nh2 all-ones qkv (causal, as the reference)
    b, t, c, nh = 2, 3, 4, 2
    inp = np.ones(b * t * 3 * c, f32)
    out = np.zeros(b * t * c, f32)
    preatt = np.zeros(b * nh * t * t, f32)
    att = np.zeros(b * nh * t * t, f32)
    po.attention_forward(out, preatt, att, inp, b, t, c, nh)
    # all-ones V: any convex combination is exactly 1 (this is what Q2 breaks in the reference)
    assert np.allclose(out, 1.0, atol=1e-6)
    hs = c // nh
    pre = preatt.reshape(b, nh, t, t)
    a = att.reshape(b, nh, t, t)
    for tq in range(t):
        assert np.allclose(pre[:, :, tq, :tq + 1], hs / np.sqrt(hs))
        assert np.allclose(a[:, :, tq, :tq + 1], 1.0 / (tq + 1), atol=1e-6)
        assert np.all(a[:, :, tq, tq + 1:] == 0)


def test_layernorm_forward():
    # vit_tests.rs:163-190 — all-ones input, w 2, bias 3: out 3, mean 1, rstd 1/sqrt(1e-5)
    b, t, c = 2, 3, 4
    out, mean, rstd = np.zeros(b * t * c, f32), np.zeros(b * t, f32), np.zeros(b * t, f32)
    po.layernorm_forward(out, mean, rstd, np.ones(b * t * c, f32), np.full(c, 2.0, f32), np.full(c, 3.0, f32), b, t, c)
    assert np.array_equal(out, np.full(b * t * c, 3.0, f32))
    assert np.array_equal(mean, np.ones(b * t, f32))
    assert np.allclose(rstd, 1.0 / np.sqrt(1e-5), rtol=1e-6)


def test_gelu_forward():
    # vit_tests.rs:193-201 — gelu(1.0) = 0.841192
    out = np.zeros(10, f32)
    po.gelu_forward(out, np.ones(10, f32), 10)
    assert np.allclose(out, 0.841192, atol=1e-6)


def test_softmax_forward():
    # vit_tests.rs:204-230 — uniform logits, each row sums to 1 within 1e-6
    b, t, v = 2, 3, 4
    probs = np.zeros(b * t * v, f32)
    po.softmax_forward(probs, np.ones(b * t * v, f32), b, t, v)
    rows = probs.reshape(b * t, v)
    assert np.all(np.abs(rows.sum(axis=1) - 1.0) < 1e-6)
    assert np.allclose(rows, 0.25)


def test_forward_pass_loss_positive():
    # vit_tests.rs:19-50 — mean_loss > 0; encoded/logits/probs/losses not all zero
    m = po.ViT("tiny")
    images, labels = po.synthetic_batch(po.CONFIGS["tiny"], 4)
    loss = m.forward(images, labels)
    assert loss > 0.0
    for name in ("encoded", "logits", "probs", "losses"):
        assert np.any(m.act(name) != 0.0)
    # no targets -> mean_loss = -1 (rusty_vit.rs:348-350)
    assert m.forward(images, None) == -1.0


def test_backward_pass_all_grads_touched():
    # vit_tests.rs:53-89 — every gradient view exists; here: every tensor receives gradient
    m = po.ViT("tiny")
    images, labels = po.synthetic_batch(po.CONFIGS["tiny"], 4)
    m.forward(images, labels)
    m.zero_grad()
    m.backward()
    for name in po.PARAM_NAMES:
        assert np.any(m.grad(name) != 0.0), name
    for name in ("encoded", "logits", "losses"):
        assert np.any(m.grad_act(name) != 0.0), name


def test_backward_accumulates():
    # train_vit.rs:538,549,552: backward ops add into their outputs
    b, t, c, oc = 1, 2, 3, 2
    rng = np.random.default_rng(0)
    inp, w, dout = (rng.standard_normal(n).astype(f32) for n in (b * t * c, oc * c, b * t * oc))
    d1 = [np.zeros(b * t * c, f32), np.zeros(oc * c, f32), np.zeros(oc, f32)]
    po.matmul_backward(*d1, dout, inp, w, b, t, c, oc)
    d2 = [x.copy() for x in d1]
    po.matmul_backward(*d2, dout, inp, w, b, t, c, oc)
    for a, bb in zip(d1, d2):
        assert np.allclose(bb, 2 * a, rtol=1e-6)


def test_rng_numpy_matches_c():
    got = po.rand_u01(1337, 6, 1000)
    want = np.array([po.lib().vit_rand_u01(1337, 6, i) for i in range(1000)], f32)
    assert np.array_equal(got, want)
    assert 0.0 <= got.min() and got.max() < 1.0
