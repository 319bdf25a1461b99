"""Experiment: weight-gradient GEMM (both operands MN-major, K = B*T) time vs split count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
K = 201728
def run(M, N, label):
    A = torch.randn(K, M, device="cuda").to(torch.bfloat16)   # [K, M] = MN-major A
    B = torch.randn(K, N, device="cuda").to(torch.bfloat16)
    D = torch.zeros(M, N, device="cuda")
    for splits in (0, 1, 2, 3, 4, 8, 16):
        if splits: os.environ["VITRS_GEMM_SPLITS"] = str(splits)
        else: os.environ.pop("VITRS_GEMM_SPLITS", None)
        for _ in range(2): pkg.gemm_bf16(D, A, B, M, N, K, M, N, N, 1, 1, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): pkg.gemm_bf16(D, A, B, M, N, K, M, N, N, 1, 1, 1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{label} M={M} N={N} splits={splits or 'auto'}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.0f} TFLOP/s  (min DRAM {(M+N)*K*2/1e9:.2f} GB -> {(M+N)*K*2/ms/1e6:.0f} GB/s if read once)", flush=True)
run(768, 3072, "fcproj-dW")
run(3072, 768, "fc-dW")
run(768, 768, "proj-dW")
