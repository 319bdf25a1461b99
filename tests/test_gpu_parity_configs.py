"""bf16 production step against the oracle on the BASELINE.json model configs.

The toy shapes of test_gpu_model.py never reach the kernels the headline run uses.  Here one full
step (ViT::forward + backward, rusty_vit.rs:269-449) runs on ViT-Ti/16 and ViT-B/16 at T = 197 and
on a two-block ViT-B/8 slice at T = 785 — the CTA-pair tcgen05 GEMM with every fused epilogue,
attn_fwd_tc2 / attn_bwd_persist (T <= 256) and the streaming attention kernels (T = 785), EPI_ROWDOT —
and logits, loss and EVERY parameter gradient are held to the north-star's 2e-2 (tensor-level
relative error, max |a - b| / max |b|).  The 100-step AdamW loss curve is repeated on an hs = 64,
T = 197 model so those kernels see 100 consecutive optimiser steps, and one case runs under the
reference's own all-positive init (rusty_vit.rs:864-903, init_mode 0).
"""
import numpy as np
import pytest
import torch

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
TOL_BF16 = 2e-2
FLOOR_FACTOR = 3.0  # init_mode 0 only, see test_bf16_step_matches_oracle_on_model_configs

# ViT-B/8 (T = 785, C = 768, NH = 12) cut to two blocks so the oracle finishes in seconds
B8_2L = dict(image_size=224, patch_size=8, channels=768, num_layers=2, num_heads=12, num_classes=1000)
# T = 197 with three 64-wide heads (ViT-Ti/16's block shape), two blocks, 16 classes: the loss-curve model
TI16_2L = dict(image_size=224, patch_size=16, channels=192, num_layers=2, num_heads=3, num_classes=16)
MID_HS64 = dict(image_size=64, patch_size=16, channels=256, num_layers=2, num_heads=4, num_classes=16)


def headline_ctx(vitrs):
    """A context that keeps the GEMM tile choice of the headline run (CTA pairs on 256 x 256 tiles) whatever the problem size:
    at batch 2-4 the library would otherwise pick its small-problem tiles, and these tests exist to put the production kernels
    in front of the oracle.  (The default choice is what the inference-engine and toy-shape tests run.)"""
    import os
    os.environ["VITRS_GEMM_NO_SMALL"] = "1"
    try:
        return vitrs.Context(0)  # switches are read when the context is created
    finally:
        os.environ.pop("VITRS_GEMM_NO_SMALL", None)


def relerr(got, want):
    got, want = np.asarray(got, np.float64).ravel(), np.asarray(want, np.float64).ravel()
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


def to_dev(images, labels):
    return torch.from_numpy(images).cuda(), torch.from_numpy(labels).cuda()


def bf16_storage_floor(cfg, ref, images, labels):
    """Tensor-level relative error of an IDEAL bf16-storage evaluation (fp32 arithmetic, values rounded only where the
    production mode stores bf16; tests/torch_ref.py) against the fp32 oracle: what the precision itself costs."""
    from tests import torch_ref
    p = torch_ref.params_from_flat(ref.params_flat(), cfg, ref.param_sizes, po.PARAM_NAMES)
    logits, loss, _ = torch_ref.forward_bf16_storage(p, cfg, images, labels)
    loss.backward()
    floor = {"logits": relerr(logits.detach().numpy(), ref.act("logits"))}
    for name in po.PARAM_NAMES:
        floor["d" + name] = relerr(p[name].grad.numpy(), ref.grad(name))
    return floor


@pytest.mark.parametrize("cfg_name,b,init_mode", [("ti16", 4, 1), ("b16", 2, 1), (B8_2L, 2, 1),
                                                  ("ti16", 4, 0), (MID_HS64, 16, 0)],
                         ids=["ti16-b4", "b16-b2", "b8x2-b2", "ti16-b4-refinit", "mid-b16-refinit"])
def test_bf16_step_matches_oracle_on_model_configs(vitrs, cfg_name, b, init_mode):
    """init_mode 1 (symmetric weights): logits, loss and every parameter gradient within 2e-2.

    init_mode 0 is the reference's own init, U[0,1)*0.02 (rusty_vit.rs:864-903): every weight is positive, so each
    matmul output carries a large component common to all its channels that the next LayerNorm removes again; a
    bf16 tensor holds 8 significant bits of that common component, not of the signal.  The loss still agrees to
    2e-2, but an IDEAL bf16-storage evaluation (exact arithmetic, rounding only where bf16 is stored) is already
    2e-2 .. 1.2e-1 away from fp32 on logits and the LayerNorm gradients (DEVIATIONS D14).  There the bar is: no
    tensor worse than 2e-2 or FLOOR_FACTOR x that precision floor, whichever is larger.  (The floor is one realisation of
    the rounding noise and the fused kernels round at other, not more, places than the unfused emulation — measured
    GPU / floor ratios on the two cases: 0.8 .. 2.7, the largest on the LayerNorm bias gradients.)"""
    cfg = po.CONFIGS[cfg_name] if isinstance(cfg_name, str) else cfg_name
    ref = po.ViT(cfg_name, seed=1337, init_mode=init_mode)
    m = vitrs.ViT(cfg_name, max_batch=b, mode=vitrs.MODE_BF16, seed=1337, init_mode=init_mode, ctx=headline_ctx(vitrs))
    assert np.array_equal(m.params_flat().cpu().numpy(), ref.params_flat())
    images, labels = po.synthetic_batch(cfg, b)
    ref_loss = ref.forward(images, labels); ref.zero_grad(); ref.backward()
    m.zero_grad(); m.forward(*to_dev(images, labels)); m.backward()
    assert abs(m.mean_loss - ref_loss) <= TOL_BF16 * abs(ref_loss), (m.mean_loss, ref_loss)
    report = {"logits": relerr(m.act("logits").cpu().numpy(), ref.act("logits"))}
    for name in po.PARAM_NAMES:
        report["d" + name] = relerr(m.grad(name).cpu().numpy(), ref.grad(name))
    bound = {k: TOL_BF16 for k in report}
    if init_mode == 0:
        floor = bf16_storage_floor(cfg, ref, images, labels)
        bound = {k: max(TOL_BF16, FLOOR_FACTOR * floor[k]) for k in report}
        print("bf16 storage floor", {k: f"{v:.2e}" for k, v in floor.items()})
    else:
        # activations: the north-star gates logits, loss and gradients; intermediate tensors are gated where no rounding has
        # accumulated yet (the first block) and printed for the last block, where bf16 storage of the residual stream over up
        # to 12 blocks puts qkv / atty / the residual stream itself at 1e-2 .. 2.1e-2 of their per-block maximum
        L = cfg["num_layers"]
        for name in ("encoded", "qkv", "atty", "residual2", "fch", "residual3"):
            got = m.act(name).float().cpu().numpy()
            want = ref.act(name)[:got.size]
            per = got.size // (1 if name == "encoded" else L)
            report[name + "[0]"] = relerr(got[:per], want[:per])
            bound[name + "[0]"] = TOL_BF16
            if name != "encoded":
                print(f"{name}[{L - 1}] {relerr(got[-per:], want[-per:]):.2e}", end="  ")
        print()
    print("gpu vs oracle", {k: f"{v:.2e}" for k, v in report.items()})
    bad = {k: f"{v:.3e} > {bound[k]:.3e}" for k, v in report.items() if not v <= bound[k]}
    assert not bad, bad
    m.close()


@pytest.mark.parametrize("cfg,b", [(TI16_2L, 4), (MID_HS64, 16)], ids=["t197-hs64", "mid-hs64"])
def test_bf16_loss_curve_100_steps_hs64(vitrs, cfg, b):
    """100 AdamW steps through the tensor-core attention kernels and the CTA-pair GEMM vs the oracle's curve."""
    steps, lr = 100, 3e-4
    ref = po.ViT(cfg, seed=1337, init_mode=1)
    m = vitrs.ViT(cfg, max_batch=b, mode=vitrs.MODE_BF16, seed=1337, init_mode=1, ctx=headline_ctx(vitrs))
    ref_curve, got_curve = [], []
    for step in range(steps):
        images, labels = po.synthetic_batch(cfg, b, step=step % 8)  # 8 repeating batches: the loss must fall
        ref_curve.append(ref.forward(images, labels)); ref.zero_grad(); ref.backward(); ref.update(lr)
        m.train_step(*to_dev(images, labels), lr)
        got_curve.append(m.mean_loss)
    ref_curve, got_curve = np.array(ref_curve), np.array(got_curve)
    assert ref_curve[-4:].mean() < ref_curve[:4].mean() - 0.05, ref_curve
    dev_ = np.abs(got_curve - ref_curve)
    msg = f"max {dev_.max():.4f} mean {dev_.mean():.4f} last {got_curve[-4:]} vs {ref_curve[-4:]}"
    assert dev_.max() <= TOL_BF16 * ref_curve.max(), msg
    m.close()


@pytest.mark.parametrize("cfg_name,b", [("b16", 2), ("ti16", 4)], ids=["b16-b2", "ti16-b4"])
def test_inference_engine_logits_match_oracle(vitrs, cfg_name, b):
    """The graph-replayed, activation-free forward (vitrs_infer_*, rusty_vit.rs:339-350) against the oracle's logits at 2e-2."""
    ref = po.ViT(cfg_name, seed=1337, init_mode=1)
    m = vitrs.ViT(cfg_name, max_batch=1, mode=vitrs.MODE_BF16, seed=1337, init_mode=1)  # a serving model: no batch-sized arena
    eng = vitrs.InferenceEngine(m, max_batch=b)
    images, labels = po.synthetic_batch(po.CONFIGS[cfg_name], b)
    ref.forward(images, None)
    want = ref.act("logits").reshape(b, -1)
    x = torch.from_numpy(images).cuda()
    for call in range(3):
        logits, probs = eng.forward(x)
        assert relerr(logits.cpu().numpy(), want) <= TOL_BF16, call
    assert np.abs(probs.sum(dim=1).cpu().numpy() - 1.0).max() < 1e-5
    assert relerr(eng.forward_host(images), want) <= TOL_BF16
    eng.close(); m.close()
