"""Timeline of CTA 0 of the persistent attention kernels from the diagnostic build (make -C vit.rs_b200/csrc trace):
SM-clock stamps at the pipeline's hand-over points, printed per head as cycle offsets.  KERNEL=bwd|fwd, B = batch."""
import ctypes as C
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["VITRS_LIB"] = os.path.join(ROOT, "vit.rs_b200", "libvitrs_trace.so")
os.environ.setdefault("VITRS_ATTN_BWD_OVERWRITE", "1")
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
lib = pkg.lib()
b, t, c, nh = int(os.environ.get("B", 128)), int(os.environ.get("T", 197)), 768, 12
which = os.environ.get("KERNEL", "bwd")
qkv = (torch.randn(b * t * 3 * c, device="cuda") * 0.5).to(torch.bfloat16)
dout = (torch.randn(b * t * c, device="cuda") * 0.1).to(torch.bfloat16)
out = torch.zeros(b * t * c, device="cuda", dtype=torch.bfloat16)
dqkv = torch.zeros(b * t * 3 * c, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(b * nh * t, device="cuda")
buf = (C.c_ulonglong * (1 << 16))()
n = C.c_uint()
def run():
    pkg.attention_forward(out, lse, None, qkv, b, t, c, nh, causal=0)
    if which == "bwd":
        lib.vitrs_debug_trace_read(buf, C.byref(n))  # drop the forward's stamps
        pkg.attention_backward_bf16(dqkv, dout, out, lse, qkv, b, t, c, nh, causal=0)
for _ in range(3):
    run()
    lib.vitrs_debug_trace_read(buf, C.byref(n))
run()
lib.vitrs_debug_trace_read(buf, C.byref(n))
ev = sorted(((buf[i] >> 24, (buf[i] >> 16) & 255, (buf[i] >> 8) & 255, buf[i] & 255) for i in range(n.value)))
t0 = ev[0][0]
names = {1: "head", 2: "wait s_full", 3: "got s_full+ds_free", 4: "P/dS done", 5: "store_acc begin", 6: "got acc_full", 7: "store_acc end",
         8: "wait dq_full", 9: "got dq_full", 10: "head end", 20: "I wait p_full", 21: "I got p_full", 22: "I wait load0", 23: "I got load0",
         24: "I issued", 30: "L wait kv0", 31: "L got kv0", 32: "L got q0", 33: "L wait kv1", 34: "L got kv1",
         41: "wait s_main", 42: "got s_main", 43: "P done", 44: "got O", 45: "O staged", 46: "O stored", 50: "I PV g0", 51: "I PV g1",
         52: "I main g0", 53: "I main g1", 54: "I tail g0", 55: "I tail g1",
         60: "first chunk loaded", 61: "reference agreed", 62: "chunk done", 63: "P stores done", 64: "halves met"}
print(f"{n.value} stamps, span {ev[-1][0] - t0} cycles, heads per CTA = {b * nh / 148:.1f}")
# print the full timeline of a steady-state window: from the 3rd 'head'(1)/(41) event of the first SIMT warp
lo = int(os.environ.get("FROM", 0)); hi = int(os.environ.get("TO", 60000))
warps = [int(w) for w in os.environ.get("WARPS", "0,4,8,12" if which == "bwd" else "1,2,6").split(",")]
last = {}
for clk, w, e, a in ev:
    rel = clk - t0
    if rel < lo or rel > hi or w not in warps: continue
    d = rel - last.get(w, rel)
    last[w] = rel
    print(f"{rel:8d} (+{d:5d}) w{w:<2d} {names.get(e, e)} {a}")
