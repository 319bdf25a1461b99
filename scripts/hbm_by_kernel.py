"""The bandwidth-bound kernels of one training step against the HBM roofline: algorithmic bytes per launch
over the time of the committed ncu launch list (isolated clocks), as a fraction of the measured copy bandwidth
(MEASURED_PEAKS.json hbm_gbs, else the 6 553.9 GB/s this pool's B200s measured).  No GPU needed.
Usage: python scripts/hbm_by_kernel.py b16 1024 profiles/r2_launches_vitb16_b1024.csv"""
import collections, csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
pkg = ge.load_package()
cfg_name, B, path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
cfg = pkg.CONFIGS[cfg_name]
peak = 6553.9
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
T = (cfg["image_size"] // cfg["patch_size"]) ** 2 + 1
C, NH, V, img, kdim = cfg["channels"], cfg["num_heads"], cfg["num_classes"], cfg["image_size"], 3 * cfg["patch_size"] ** 2
M = B * T
n_params = pkg.model_footprint(cfg_name, B)["num_parameters"]
# kernel-name prefix -> (algorithmic bytes per launch, what they are); bf16 activations (2 B), fp32 statistics / parameters (4 B)
BYTES = {
    "ln_fwd_kernel<__nv_bfloat16": (2 * M * C * 2 + 8 * M, "row in, row out, mean / rstd out"),
    "ln_bwd_kernel<__nv_bfloat16": (4 * M * C * 2, "dout, inp in; dinp in and out (residual accumulation fused)"),
    "adamw_kernel": (30 * n_params, "fp32 p, g, m, v in; p, m, v out; bf16 shadow out"),
    "im2col_kernel": (B * 3 * img * img * 4 + M * kdim * 2, "fp32 images in, bf16 patch rows out"),
    "patch_bwd_reduce_kernel": (M * C * 2, "bf16 dencoded in (column sums into dwpe / dcls / dpatchb)"),
    "attn_fwd_persist_kernel": (M * 3 * C * 2 + M * C * 2 + B * NH * T * 4, "qkv in, out and lse out"),
    "attn_bwd_persist_kernel": (M * 3 * C * 2 + M * C * 2 + M * 3 * C * 2 + 2 * B * NH * T * 4, "qkv, dout, lse, D in; dqkv out"),
    "attn_fwd_stream_kernel": (M * 3 * C * 2 + M * C * 2 + B * NH * T * 4, "qkv in, out and lse out (K / V re-read per query tile through L2)"),
    "head_loss_kernel": (3 * B * V * 4, "logits in, probs and dlogits out"),
}
rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
ki, vi = rows[0].index("Kernel Name"), rows[0].index("Metric Value")
times = collections.OrderedDict()
total = 0.0
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[ki]).replace("<unnamed>::", "")
    us = float(r[vi]) / 1e3
    total += us
    for prefix in BYTES:
        if name.startswith(prefix):
            times.setdefault(prefix, []).append(us)
print(f"# {os.path.relpath(path, ROOT)}: bandwidth-bound kernels of one vit-{cfg_name} step at batch {B} (isolated clocks); HBM peak {peak:.1f} GB/s (measured copy)")
print(f"{'kernel':34s}  n  {'mean us':>8s} {'alg. MB':>9s} {'GB/s':>8s} {'of peak':>8s} {'share of step':>14s}  bytes counted")
for prefix, v in times.items():
    nbytes, what = BYTES[prefix]
    mean = sum(v) / len(v)
    gbs = nbytes / (mean * 1e-6) / 1e9
    print(f"{prefix.replace('__nv_bfloat16', 'bf16'):34s} {len(v):2d}  {mean:8.1f} {nbytes / 1e6:9.1f} {gbs:8.0f} {gbs / peak:8.2f} {100 * sum(v) / total:13.1f}%  {what}")
