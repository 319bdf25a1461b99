"""Generates tests/golden/*.npz from PyTorch CPU fp32 (the independent implementation).

The reference itself cannot be executed (no Rust toolchain, sources do not compile —
SURVEY §0.3), so these vectors come from torch library ops on the shared synthetic inputs.
Run:  python tests/golden/make_golden.py     (deterministic; commit the outputs)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402  (numpy generator + config table only)
from tests import torch_ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def init_params_numpy(cfg_name, seed=1337):
    """D9 init restated in numpy: weights U[0,1)*0.02, LN gains 1, biases 0."""
    cfgc = po.make_config(cfg_name)
    sizes = (po.C.c_size_t * 20)()
    po.lib().vit_param_sizes(po.C.byref(cfgc), sizes)
    sizes = list(sizes)
    flat = np.zeros(sum(sizes), np.float32)
    off = 0
    for i, (n, s) in enumerate(zip(po.PARAM_NAMES, sizes)):
        if n in ("patchw", "cls", "wpe", "qkvw", "attprojw", "fcw", "fcprojw", "headw"):
            flat[off:off + s] = po.fill_uniform(s, seed, i, 0.0, 0.02)
        elif n in ("ln1w", "ln2w", "lnfw"):
            flat[off:off + s] = 1.0
        off += s
    return flat, sizes


def model_case(cfg_name, b, causal, steps):
    cfg = po.CONFIGS[cfg_name]
    flat, sizes = init_params_numpy(cfg_name)
    tflat = torch.tensor(flat, requires_grad=True)
    opt = torch.optim.AdamW([tflat], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    out = {"losses": []}
    for step in range(steps):
        images, labels = po.synthetic_batch(cfg, b, step=step)
        p, off = {}, 0
        for n, s in zip(po.PARAM_NAMES, sizes):
            p[n] = tflat[off:off + s]; off += s
        logits, loss, acts = torch_ref.forward(p, cfg, images, labels, causal=causal)
        opt.zero_grad(); loss.backward()
        if step == 0:
            out["logits"] = logits.detach().numpy()
            out["grads"] = tflat.grad.numpy().copy()
            out["encoded"] = acts["encoded"].detach().numpy()
            out["qkv0"] = acts["qkv0"].detach().numpy()
            out["atty0"] = acts["atty0"].detach().numpy()
        out["losses"].append(loss.item())
        opt.step()
    out["losses"] = np.array(out["losses"], np.float64)
    out["params_after"] = tflat.detach().numpy().copy()
    return out


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(1)
    np.savez_compressed(os.path.join(HERE, "tiny_b4_noncausal.npz"), **model_case("tiny", 4, False, 10))
    np.savez_compressed(os.path.join(HERE, "tiny_b2_causal.npz"), **model_case("tiny", 2, True, 2))
    print("written")
