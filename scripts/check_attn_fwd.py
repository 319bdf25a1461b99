"""Diagnostic: attention forward vs an fp32 torch reference, overall and for the weights of individual late keys."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
torch.manual_seed(0)
for (b, t, nh) in ((4, 197, 3), (64, 197, 12), (2, 256, 2), (2, 208, 2), (2, 160, 2)):
    c = nh * 64
    qkv = (torch.randn(b, t, 3, nh, 64, device="cuda") * 0.7)
    # V = one-hot of (key - (t - 64)): out[q, d] is the attention weight of key t - 64 + d
    v = torch.zeros(b, t, nh, 64, device="cuda")
    for d in range(64):
        v[:, t - 64 + d, :, d] = 1.0
    qkv[:, :, 2] = v
    x = qkv.reshape(b, t, 3 * c).to(torch.bfloat16).contiguous()
    out = torch.zeros(b, t, c, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(b * nh * t, device="cuda")
    pkg.attention_forward(out, lse, None, x, b, t, c, nh, causal=0)
    torch.cuda.synchronize()
    xf = x.float().view(b, t, 3, nh, 64)
    q, k, vv = (xf[:, :, i].transpose(1, 2) for i in range(3))
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    want = (att @ vv).transpose(1, 2).reshape(b, t, c)
    err = (out.float() - want).abs().view(b, t, nh, 64)
    per_key = err.amax(dim=(0, 1, 2))
    print(f"b{b} t{t} nh{nh}: max err {err.max().item():.3e} (max weight {want.max().item():.3f}); per key (last 64), worst 8:",
          [(t - 64 + int(i), f"{per_key[i].item():.1e}") for i in per_key.argsort(descending=True)[:8]])
