"""Kernel timeline of two training steps through CUPTI (torch.profiler): busy time, idle gaps between consecutive
kernels on the stream, and the per-kernel in-step durations (power-capped clocks, unlike ncu's isolated replays)."""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import __graft_entry__ as ge
pkg = ge.load_package()
cfgname = os.environ.get("CFG", "b16")
B = int(os.environ.get("BATCH", 1024))
ctx = pkg.Context(0)
model = pkg.ViT(cfgname, max_batch=B, mode=pkg.MODE_BF16, seed=1337, init_mode=1, ctx=ctx)
model.set_dloss_scale(1.0 / B)
cfg = pkg.CONFIGS[cfgname]
x = torch.empty(B, 3, cfg["image_size"], cfg["image_size"], device="cuda")
pkg.fill_uniform(x, 1337, 1000, -1.0, 1.0, ctx=ctx)
y = torch.randint(0, cfg["num_classes"], (B,), device="cuda", dtype=torch.int32)
for _ in range(4): model.train_step(x, y, 1e-4)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): model.train_step(x, y, 1e-4)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace.json")
ev = [e for e in json.load(open("/tmp/trace.json"))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# the middle step: between the 1st and 2nd adamw
adam = [i for i, e in enumerate(ev) if "adamw" in e["name"]]
step = ev[adam[0] + 1: adam[1] + 1]
t0, t1 = step[0]["ts"], step[-1]["ts"] + step[-1]["dur"]
busy = sum(e["dur"] for e in step)
gaps = [step[i + 1]["ts"] - (step[i]["ts"] + step[i]["dur"]) for i in range(len(step) - 1)]
print(f"step span {t1 - t0:.0f} us, kernels {len(step)}, busy {busy:.0f} us, gaps total {sum(g for g in gaps if g > 0):.0f} us, "
      f"overlap {-sum(g for g in gaps if g < 0):.0f} us, max gap {max(gaps):.1f} us, median gap {sorted(gaps)[len(gaps)//2]:.2f} us")
agg = collections.defaultdict(lambda: [0.0, 0])
for e in step:
    import re
    nm = e["name"].replace("(anonymous namespace)::", "").replace("void ", "")
    k = re.split(r"\(", nm)[0][:70]
    agg[k][0] += e["dur"]; agg[k][1] += 1
for k, (d, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"{d:10.0f} us {100*d/(t1-t0):5.1f}%  n={n:3d} avg={d/n:8.1f}  {k}")
big = sorted(((g, i) for i, g in enumerate(gaps)), reverse=True)[:8]
for g, i in big:
    print(f"gap {g:7.1f} us after {step[i]['name'][:50]} before {step[i+1]['name'][:50]}")
