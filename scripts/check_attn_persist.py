"""Attention backward at many heads per CTA (the persistent kernel's multi-head path): prints a digest of dqkv for seeded
inputs so that two runs (persistent / VITRS_ATTN_BWD_NOPERSIST=1) can be compared, and checks against a torch reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
b, t, c, nh = int(os.environ.get("B", 64)), int(os.environ.get("T", 197)), 768, 12
g = torch.Generator(device="cuda").manual_seed(t * 7 + b)
qkv = (torch.randn(b, t, 3 * c, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
dout = (torch.randn(b, t, c, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
out = torch.zeros(b, t, c, device="cuda", dtype=torch.bfloat16)
dqkv = torch.zeros(b, t, 3 * c, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(b * nh * t, device="cuda")
pkg.attention_forward(out, lse, None, qkv, b, t, c, nh, causal=0)
pkg.attention_backward_bf16(dqkv, dout, out, lse, qkv, b, t, c, nh, causal=0)
torch.cuda.synchronize()
# torch reference (fp32 autograd on the bf16-rounded inputs)
x = qkv.float().requires_grad_(True)
q, k, v = x.split(c, dim=2)
q = q.view(b, t, nh, 64).transpose(1, 2); k = k.view(b, t, nh, 64).transpose(1, 2); v = v.view(b, t, nh, 64).transpose(1, 2)
att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
y = (att @ v).transpose(1, 2).reshape(b, t, c)
y.backward(dout.float())
err = (dqkv.float() - x.grad).abs().max().item() / x.grad.abs().max().item()
print(f"T={t} B={b} digest {dqkv.float().sum().item():.6e} {dqkv.float().abs().sum().item():.6e} relerr_vs_torch {err:.3e}", flush=True)
