"""Per GEMM shape of one training step: launches, mean time and TFLOP/s, from a committed ncu launch list (profiles/r2_launches_*.csv:
`ncu --metrics gpu__time_duration.sum --clock-control none`, each kernel replayed alone at isolated clocks) joined with the host-side
plan of the step (vitrs.step_gemms / vitrs.gemm_plan — no GPU needed).  Usage: python scripts/gemm_by_shape.py b16 1024 profiles/r2_launches_vitb16_b1024.csv"""
import collections, csv, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
pkg = ge.load_package()
cfg, batch, path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
ki, vi = rows[0].index("Kernel Name"), rows[0].index("Metric Value")
ran = [(re.search(r"gemm_tc_kernel<([^>]*)>", r[ki]).group(1), float(r[vi]) / 1e3) for r in rows[1:] if "gemm_tc_kernel" in r[ki]]
first = next(i for i, (t, _) in enumerate(ran) if t.endswith("1"))  # the step starts at its patch-embedding GEMM
gemms = pkg.step_gemms(cfg, batch)
times = collections.OrderedDict()
for g, (_, us) in zip(gemms, ran[first:first + len(gemms)]):
    times.setdefault(g[0], []).append(us)
print(f"# {os.path.relpath(path, ROOT)}: the {len(gemms)} tcgen05 GEMM launches of one vit-{cfg} step at batch {batch}, by shape (isolated clocks)")
print(f"{'gemm':12s} {'M':>7s} {'N':>5s} {'K':>7s}  n  {'mean us':>8s} {'TFLOP/s':>8s}  tiles x splits -> CTAs")
total_us = total_fl = 0.0
for name, v in times.items():
    _, M, N, K, a_mn, b_mn, epi = next(g for g in gemms if g[0] == name)
    p = pkg.gemm_plan(M, N, K, a_mn, b_mn, epi)
    mean = sum(v) / len(v)
    fl = 2.0 * M * N * K
    total_us += sum(v); total_fl += fl * len(v)
    print(f"{name:12s} {M:7d} {N:5d} {K:7d} {len(v):2d}  {mean:8.1f} {fl / (mean * 1e-6) / 1e12:8.1f}  {p['tiles']} x {p['splits']} -> {p['grid']}")
print(f"{'all':12s} {'':21s} {sum(len(v) for v in times.values()):3d} {total_us:8.1f} {total_fl / (total_us * 1e-6) / 1e12:8.1f}")
