"""Microbenchmark of LayerNorm forward / backward on the ViT-B/16 activation shape [1024*197, 768] bf16 (HBM-bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
B, T, C = 1024, 197, int(os.environ.get("C", 768))
R = B * T
x = torch.randn(R, C, device="cuda").to(torch.bfloat16)
dy = torch.randn(R, C, device="cuda").to(torch.bfloat16)
dx = torch.randn(R, C, device="cuda").to(torch.bfloat16)
y = torch.empty_like(x)
w = torch.rand(C, device="cuda") + 0.5
b = torch.randn(C, device="cuda")
mean = torch.empty(R, device="cuda"); rstd = torch.empty(R, device="cuda")
dw = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tag = os.environ.get("TAG", "")
ms = timeit(lambda: pkg.layernorm_forward(y, mean, rstd, x, w, b, B, T, C))
print(f"{tag} ln_fwd: {ms*1e3:.1f} us  {2*R*C*2/ms/1e6:.0f} GB/s")
ms = timeit(lambda: pkg.layernorm_backward(dx, dw, db, dy, x, w, mean, rstd, B, T, C))
print(f"{tag} ln_bwd: {ms*1e3:.1f} us  {4*R*C*2/ms/1e6:.0f} GB/s")
