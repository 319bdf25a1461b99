"""CUPTI kernel timeline of one data-parallel training step (rank 0 of N; launch under torch.distributed.run).

Names the overlap loss of the gradient exchange: for every NCCL kernel of the step its span, and for the GEMMs that run
while an NCCL kernel is resident their duration against the same GEMM shapes' duration in the forward-free part of the step
(no collective in flight).  Output: one text report (stdout of rank 0)."""
import os, sys, json, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import __graft_entry__ as ge
pkg = ge.load_package()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = pkg.Context(local)
if world > 1:
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    mc = int(os.environ.get("MAX_CTAS", -1))
    ctx.comm_init(uid[0], rank, world, max_ctas=None if mc < 0 else mc)
cfgname, B = os.environ.get("CFG", "b16"), int(os.environ.get("BATCH", 1024))
model = pkg.ViT(cfgname, max_batch=B, mode=pkg.MODE_BF16, seed=1337, init_mode=1, ctx=ctx)
model.set_dloss_scale(1.0 / (B * world))
if os.environ.get("ZERO1"):
    model.enable_zero1()
cfg = pkg.CONFIGS[cfgname]
x = torch.empty(B, 3, cfg["image_size"], cfg["image_size"], device="cuda")
pkg.fill_uniform(x, 1337 + rank, 1000, -1.0, 1.0, ctx=ctx)
y = torch.randint(0, cfg["num_classes"], (B,), device="cuda", dtype=torch.int32)
for _ in range(4): model.train_step(x, y, 1e-4)
torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize()
if rank == 0:
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): model.train_step(x, y, 1e-4)
        torch.cuda.synchronize()
    prof.export_chrome_trace("/tmp/trace_dp.json")
else:
    for _ in range(3): model.train_step(x, y, 1e-4)
    torch.cuda.synchronize()
if world > 1: dist.barrier()
if rank == 0:
    ev = [e for e in json.load(open("/tmp/trace_dp.json"))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    short = lambda n: re.split(r"\(", n.replace("(anonymous namespace)::", "").replace("void ", ""))[0][:60]
    main = [e for e in ev if "nccl" not in e["name"].lower()]
    adam = [i for i, e in enumerate(main) if "adamw" in e["name"]]
    t0 = main[adam[0]]["ts"] + main[adam[0]]["dur"]
    t1 = main[adam[1]]["ts"] + main[adam[1]]["dur"]
    step = [e for e in ev if t0 <= e["ts"] < t1]
    nccl = [e for e in step if "nccl" in e["name"].lower()]
    print(f"world {world}  step span {t1 - t0:.0f} us  kernels {len(step)}  nccl kernels {len(nccl)}  nccl busy {sum(e['dur'] for e in nccl):.0f} us")
    for e in nccl:
        print(f"  nccl at +{e['ts'] - t0:9.0f} us  dur {e['dur']:8.0f} us  grid {e['args'].get('grid')}  block {e['args'].get('block')}  {short(e['name'])}")
    def overlapped(e):
        return any(n["ts"] < e["ts"] + e["dur"] and e["ts"] < n["ts"] + n["dur"] for n in nccl)
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0])
    for e in step:
        if "nccl" in e["name"].lower(): continue
        k = short(e["name"]) + " grid" + str(e["args"].get("grid"))
        a = agg[k]
        if overlapped(e): a[2] += e["dur"]; a[3] += 1
        else: a[0] += e["dur"]; a[1] += 1
    print("kernel (by grid): alone n / avg us   |   while an NCCL kernel is resident n / avg us")
    for k, (d, n, do, no) in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][2]))[:14]:
        print(f"  {k:78s} {n:4d} {d / max(n, 1):9.1f}   | {no:4d} {do / max(no, 1):9.1f}")
    tot_alone = sum(v[0] for v in agg.values()); tot_ov = sum(v[2] for v in agg.values())
    print(f"compute busy {tot_alone + tot_ov:.0f} us of {t1 - t0:.0f} us span ({tot_ov:.0f} us of it under a resident NCCL kernel)")
model.close()
if world > 1: dist.destroy_process_group()
