"""Independent PyTorch CPU fp32 implementation of the ViT step (second oracle).

Used to pin oracle/vit_oracle.c: it shares no code with it (torch library ops: conv2d,
layer_norm, gelu(tanh), scaled_dot_product_attention, cross_entropy, optim.AdamW).
"""
import numpy as np
import torch
import torch.nn.functional as F


def params_from_flat(flat, cfg, sizes, names):
    out, off = {}, 0
    for n, s in zip(names, sizes):
        out[n] = torch.tensor(np.array(flat[off:off + s]), dtype=torch.float32, requires_grad=True)
        off += s
    return out


def forward(p, cfg, images, labels, causal=False, dloss_scale=None):
    C, L, NH = cfg["channels"], cfg["num_layers"], cfg["num_heads"]
    ps, V = cfg["patch_size"], cfg["num_classes"]
    B = images.shape[0]
    x = torch.as_tensor(images, dtype=p["patchw"].dtype)  # (float64 parameters give a float64 evaluation: the conditioning tests)
    w = p["patchw"].view(C, 3, ps, ps)
    tok = F.conv2d(x, w, p["patchb"], stride=ps)                 # [B,C,g,g]
    tok = tok.flatten(2).transpose(1, 2)                          # [B,N,C]
    cls = p["cls"].view(1, 1, C).expand(B, 1, C)
    h = torch.cat([cls, tok], dim=1)
    T = h.shape[1]
    h = h + p["wpe"].view(1, T, C)
    acts = {"encoded": h}
    hs = C // NH
    for l in range(L):
        ln1 = F.layer_norm(h, (C,), p["ln1w"].view(L, C)[l], p["ln1b"].view(L, C)[l], 1e-5)
        qkv = F.linear(ln1, p["qkvw"].view(L, 3 * C, C)[l], p["qkvb"].view(L, 3 * C)[l])
        q, k, v = qkv.split(C, dim=2)
        q = q.view(B, T, NH, hs).transpose(1, 2)
        k = k.view(B, T, NH, hs).transpose(1, 2)
        v = v.view(B, T, NH, hs).transpose(1, 2)
        y = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
        y = y.transpose(1, 2).reshape(B, T, C)
        h = h + F.linear(y, p["attprojw"].view(L, C, C)[l], p["attprojb"].view(L, C)[l])
        ln2 = F.layer_norm(h, (C,), p["ln2w"].view(L, C)[l], p["ln2b"].view(L, C)[l], 1e-5)
        f = F.linear(ln2, p["fcw"].view(L, 4 * C, C)[l], p["fcb"].view(L, 4 * C)[l])
        f = F.gelu(f, approximate="tanh")
        h = h + F.linear(f, p["fcprojw"].view(L, C, 4 * C)[l], p["fcprojb"].view(L, C)[l])
        if l == 0:
            acts["qkv0"], acts["atty0"] = qkv, y
    acts["residual3_last"] = h
    lnf = F.layer_norm(h[:, 0, :], (C,), p["lnfw"], p["lnfb"], 1e-5)
    logits = F.linear(lnf, p["headw"].view(V, C), p["headb"])
    acts["logits"] = logits
    if labels is None:
        return logits, None, acts
    losses = F.cross_entropy(logits, torch.as_tensor(labels, dtype=torch.long), reduction="none")
    loss = losses.mean()
    return logits, loss, acts


# ---- idealised bf16 storage (error floor of the production mode) ---------------------------------
class _RoundBf16(torch.autograd.Function):
    """Value stored as bf16 and read back: rounds the activation forward and its gradient backward."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _ShadowBf16(torch.autograd.Function):
    """bf16 shadow of an fp32 master weight: rounded when read, gradient kept in fp32."""

    @staticmethod
    def forward(ctx, w):
        return w.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def forward_bf16_storage(p, cfg, images, labels):
    """The same step with every tensor the production mode keeps in bf16 rounded to bf16 where it is
    stored (activations, their gradients, the weight shadows) and exact fp32 arithmetic in between.

    No kernel can do better than this with bf16 storage, so its distance from the fp32 evaluation is
    the error floor that the bf16 GPU path is compared against when the weights are ill conditioned
    (the reference's all-positive init, DEVIATIONS D14)."""
    r, rw = _RoundBf16.apply, _ShadowBf16.apply
    C, L, NH = cfg["channels"], cfg["num_layers"], cfg["num_heads"]
    ps, V = cfg["patch_size"], cfg["num_classes"]
    B = images.shape[0]
    x = r(torch.as_tensor(images, dtype=torch.float32))
    tok = F.conv2d(x, rw(p["patchw"]).view(C, 3, ps, ps), p["patchb"], stride=ps).flatten(2).transpose(1, 2)
    h = torch.cat([p["cls"].view(1, 1, C).expand(B, 1, C), tok], dim=1)
    T = h.shape[1]
    h = r(h + p["wpe"].view(1, T, C))
    acts = {"encoded": h}
    hs = C // NH
    for l in range(L):
        ln1 = r(F.layer_norm(h, (C,), p["ln1w"].view(L, C)[l], p["ln1b"].view(L, C)[l], 1e-5))
        qkv = r(F.linear(ln1, rw(p["qkvw"]).view(L, 3 * C, C)[l], p["qkvb"].view(L, 3 * C)[l]))
        q, k, v = [t.view(B, T, NH, hs).transpose(1, 2) for t in qkv.split(C, dim=2)]
        y = r(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, C))
        h = r(h + F.linear(y, rw(p["attprojw"]).view(L, C, C)[l], p["attprojb"].view(L, C)[l]))
        ln2 = r(F.layer_norm(h, (C,), p["ln2w"].view(L, C)[l], p["ln2b"].view(L, C)[l], 1e-5))
        f = r(F.linear(ln2, rw(p["fcw"]).view(L, 4 * C, C)[l], p["fcb"].view(L, 4 * C)[l]))
        g = r(F.gelu(f, approximate="tanh"))
        h = r(h + F.linear(g, rw(p["fcprojw"]).view(L, C, 4 * C)[l], p["fcprojb"].view(L, C)[l]))
        if l == 0:
            acts["qkv0"], acts["atty0"] = qkv, y
    acts["residual3_last"] = h
    lnf = F.layer_norm(h[:, 0, :], (C,), p["lnfw"], p["lnfb"], 1e-5)
    logits = F.linear(lnf, p["headw"].view(V, C), p["headb"])
    acts["logits"] = logits
    loss = F.cross_entropy(logits, torch.as_tensor(labels, dtype=torch.long))
    return logits, loss, acts
