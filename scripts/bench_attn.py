"""Microbenchmark of the fused attention ops through the C ABI (CUDA events, warm L2 irrelevant: 1.2 GB of qkv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
b, t, c, nh = int(os.environ.get("B", 1024)), int(os.environ.get("T", 197)), 768, 12
qkv = (torch.randn(b * t * 3 * c, device="cuda") * 0.5).to(torch.bfloat16)
dout = (torch.randn(b * t * c, device="cuda") * 0.1).to(torch.bfloat16)
out = torch.zeros(b * t * c, device="cuda", dtype=torch.bfloat16)
dqkv = torch.zeros(b * t * 3 * c, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(b * nh * t, device="cuda")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fwd = timeit(lambda: pkg.attention_forward(out, lse, None, qkv, b, t, c, nh, causal=0))
bwd = timeit(lambda: pkg.attention_backward_bf16(dqkv, dout, out, lse, qkv, b, t, c, nh, causal=0))
flops_f = 4.0 * t * t * 64 * nh * b
print(f"{os.environ.get('TAG','')} B={b} T={t}: fwd {fwd*1e3:.0f} us ({flops_f/fwd/1e9:.0f} TF/s)  bwd {bwd*1e3:.0f} us ({2.5*flops_f/bwd/1e9:.0f} TF/s)")
