"""Microbenchmark of the raw tcgen05 GEMM entry (bf16 out, no epilogue math) on the ViT-B/16 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
M = 201728
for (N, K) in ((2304, 768), (768, 768), (3072, 768), (768, 3072), (768, 2304)):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): pkg.gemm_bf16(D, A, B, M, N, K, K, K, N, 0, 0, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): pkg.gemm_bf16(D, A, B, M, N, K, K, K, N, 0, 0, 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{os.environ.get('TAG','')} N={N} K={K}: {ms*1e3:.0f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
    if os.environ.get('CUBLAS'):
        for _ in range(3): torch.matmul(A, B.t(), out=D)
        e0.record()
        for _ in range(5): torch.matmul(A, B.t(), out=D)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"cublas N={N} K={K}: {ms*1e3:.0f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
    del A, B, D

if os.environ.get('FUSED'):
    # the fused epilogues of the fc / fcproj-dX GEMMs (N=3072, K=768)
    N, K = 3072, 768
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16) * 0.03
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    aux = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    D2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for epi, name in ((1, "bias"), (2, "bias_gelu"), (2, "bias_gelu_same_dst"), (3, "bias_residual"), (4, "gelu_bwd")):
        args = (D, (D if "same" in name else D2) if epi == 2 else None, aux if epi in (3, 4) else None, bias if epi != 4 else None, None, A, B, M, N, K, K, K, N, 0, 0, epi)
        for _ in range(3): pkg.gemm_bf16_fused(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): pkg.gemm_bf16_fused(*args)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{os.environ.get('TAG','')} fused {name} N={N} K={K}: {ms*1e3:.0f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
