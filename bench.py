#!/usr/bin/env python
"""bench.py — ViT training throughput (forward + backward + AdamW) through libvitrs.so on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config b16] [--batch 1024]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU oracle (the reference cannot be built) on the host cores

One JSON line on stdout (rank 0).  `value` is whole-job images/s with the batch resident in HBM,
timed with CUDA events (max over ranks) between barriers; `e2e` is the same metric through
ViT.train_step_host with pinned HOST buffers (H2D of the batch and D2H of the loss inside the
timed region, next batch prefetched on the copy stream).  `roofline` is the tcgen05 GEMM kernel
(the dominant kernel): algorithmic flops / CUDA-event time per launch, summed over one step.
`cpu_baseline` is the oracle port timed on the host cores on a bounded sample (rank 0, N = 1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ViT train images/sec (fwd+bwd+AdamW)"
UNIT = "images/s"

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line on stdout when the box sets
# NCCL_DEBUG), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved real stdout.
_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def isolate_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")  # B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_throughput(cfg_name, batch, steps, warmup, threads=None):
    """images/s of the CPU oracle (oracle/vit_oracle.c) doing full training steps on the host cores."""
    from oracle import pyoracle as po
    if threads:
        po.lib().vit_oracle_set_threads(threads)
    cores = po.lib().vit_oracle_num_threads()
    cfg = po.CONFIGS[cfg_name]
    m = po.ViT(cfg_name, seed=1337, init_mode=1)
    images, labels = po.synthetic_batch(cfg, batch)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        m.forward(images, labels); m.zero_grad(); m.backward(); m.update(1e-3)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, cores, total / len(times)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path.  ViT.rs itself cannot be compiled
    (no Rust toolchain, sources invalid), so this is the oracle port with all host threads."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    # size the per-step sample so that the run ends within a few minutes: one calibration image first
    t0 = time.perf_counter()
    # all host threads, stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ips1, cores, _ = cpu_oracle_throughput(args.config, 1, 1, 0, threads=threads)
    calib = time.perf_counter() - t0
    budget = 150.0
    per_step = max(1, min(8, int(budget / max(1e-3, (args.steps + args.warmup) / ips1))))
    ips, cores, sec_per_step = cpu_oracle_throughput(args.config, per_step, args.steps, args.warmup, threads=threads)
    line = {"impl": "reference", "metric": METRIC, "value": round(ips, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(sec_per_step * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"vit-{args.config} train step (fwd+bwd+AdamW), CPU oracle port of ViT.rs",
                       "sample": f"{per_step} image(s) per step", "calibration_s": round(calib, 2)},
            "cpu_baseline": {"value": round(ips, 4), "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x {per_step} image(s), vit-{args.config}, fp32 fwd+bwd+AdamW"},
            "e2e": {"value": round(ips, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vitrs", choices=["vitrs", "reference"])
    ap.add_argument("--config", default="b16", choices=["tiny", "ti16", "s16", "b16", "b8"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 1024 for b16/s16, 256 for ti16/b8)")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--comm", default="bf16", choices=["bf16", "f32"], help="gradient exchange payload (N > 1)")
    ap.add_argument("--zero1", action="store_true", help="ZeRO-1: reduce-scatter, AdamW on the shard, all-gather of the bf16 weights")
    ap.add_argument("--nccl-max-ctas", type=int, default=-1, help="cap on NCCL thread blocks per collective (-1: library default)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (global batch 1024 split over the ranks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "vitrs" else args.warmup
    isolate_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libvitrs.so has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(local_rank)
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world, max_ctas=None if args.nccl_max_ctas < 0 else args.nccl_max_ctas)

    cfg = pkg.CONFIGS[args.config]
    per_gpu = args.batch or {"b16": 1024, "s16": 1024, "ti16": 256, "b8": 256, "tiny": 8}[args.config]
    global_batch = per_gpu * world
    mode = pkg.MODE_BF16 if args.mode == "bf16" else pkg.MODE_F32
    model = pkg.ViT(args.config, max_batch=per_gpu, mode=mode, seed=1337, init_mode=1, ctx=ctx)
    model.set_dloss_scale(1.0 / global_batch)
    if mode == pkg.MODE_BF16:
        model.set_comm_dtype(args.comm)
    if args.zero1:
        model.enable_zero1()
    lr = 1e-4

    # synthetic data: images U[-1,1) from the shared counter generator, labels uniform; two batches
    img = cfg["image_size"]
    d_images, d_labels, h_images, h_labels = [], [], [], []
    for i in range(2):
        x = torch.empty(per_gpu, 3, img, img, device="cuda")
        pkg.fill_uniform(x, 1337 + rank, 1000 + 2 * i, -1.0, 1.0, ctx=ctx)
        y = torch.randint(0, cfg["num_classes"], (per_gpu,), device="cuda", dtype=torch.int32)
        d_images.append(x); d_labels.append(y)
        if not args.no_e2e:
            h_images.append(x.cpu().pin_memory()); h_labels.append(y.cpu().pin_memory())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident throughput ----------------------------------------------------------------
    # (the library records a CUDA graph of the step the second time a batch buffer comes by: two untimed passes over both
    # buffers first, so that no capture falls into a timed region; the W warm-up steps follow)
    for s in range(4):
        model.train_step(d_images[s % 2], d_labels[s % 2], lr)
    for s in range(args.warmup):
        model.train_step(d_images[s % 2], d_labels[s % 2], lr)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        model.train_step(d_images[s % 2], d_labels[s % 2], lr)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    loss = model.mean_loss
    ms_per_step = ms_total / args.steps
    value = global_batch * args.steps / (ms_total / 1e3)

    # ---- end to end through the public API with host buffers -------------------------------------------
    e2e = None
    if not args.no_e2e:
        for s in range(6):
            model.train_step_host(h_images[s % 2], h_labels[s % 2], lr)
        barrier()
        model.prefetch_host(h_images[0], h_labels[0])
        e0.record()
        for s in range(args.steps):
            if s + 1 < args.steps:
                model.prefetch_host(h_images[(s + 1) % 2], h_labels[(s + 1) % 2])
            last_loss = model.train_step_host(h_images[s % 2], h_labels[s % 2], lr)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": round(global_batch * args.steps / (ms_e2e / 1e3), 2), "unit": UNIT,
               "h2d_bytes_per_step": int(world * (h_images[0].numel() * 4 + h_labels[0].numel() * 4)),
               "d2h_bytes_per_step": 4 * world, "ms_per_step": round(ms_e2e / args.steps, 3), "loss": round(last_loss, 5)}

        # the same through the raw-image entry point (uint8 host batches, normalisation fused into im2col): extra key, the
        # headline `e2e` above keeps the fp32 synthetic images of SURVEY 8-d
        h_u8 = [torch.randint(0, 256, (per_gpu, 3, img, img), dtype=torch.uint8).pin_memory() for _ in range(2)]
        for s in range(6):
            model.train_step_host_u8(h_u8[s % 2], h_labels[s % 2], lr)
        barrier()
        model.prefetch_host_u8(h_u8[0], h_labels[0])
        e0.record()
        for s in range(args.steps):
            if s + 1 < args.steps:
                model.prefetch_host_u8(h_u8[(s + 1) % 2], h_labels[(s + 1) % 2])
            model.train_step_host_u8(h_u8[s % 2], h_labels[s % 2], lr)
        e1.record()
        barrier()
        ms_u8 = max_over_ranks(e0.elapsed_time(e1))
        e2e["uint8_images"] = {"value": round(global_batch * args.steps / (ms_u8 / 1e3), 2), "ms_per_step": round(ms_u8 / args.steps, 3),
                               "h2d_bytes_per_step": int(world * (h_u8[0].numel() + h_labels[0].numel() * 4))}
        del h_u8

    # ---- inference forward (no targets: logits only, rusty_vit.rs:339-350), extra key ------------------------------------
    for s in range(2):
        model.forward(d_images[s % 2], None)
    barrier()
    e0.record()
    for s in range(args.steps):
        model.forward(d_images[s % 2], None)
    e1.record()
    barrier()
    ms_inf = max_over_ranks(e0.elapsed_time(e1))
    inference = {"value": round(global_batch * args.steps / (ms_inf / 1e3), 2), "unit": "images/s", "ms_per_batch": round(ms_inf / args.steps, 3)}

    # the same through the inference engine (vitrs_infer_*: ping-pong workspace, CUDA-graph replay) incl. small-batch latency
    # from HOST buffers (H2D + forward + D2H of the logits, synchronised): the serving figures of SURVEY 8-f.3
    if world == 1 and mode == pkg.MODE_BF16:
        eng = pkg.InferenceEngine(model, max_batch=per_gpu)
        for s in range(3):
            eng.forward(d_images[s % 2])
        barrier()
        e0.record()
        for s in range(args.steps):
            eng.forward(d_images[s % 2])
        e1.record()
        barrier()
        ms_eng = e0.elapsed_time(e1)
        inference["engine"] = {"value": round(per_gpu * args.steps / (ms_eng / 1e3), 2), "ms_per_batch": round(ms_eng / args.steps, 3),
                               "workspace_bytes": eng.stats()["workspace_bytes"], "latency_ms_host_to_host": {}}
        import numpy as np
        for lb in (1, 8, 64):
            if lb > per_gpu:
                continue
            hx = np.ascontiguousarray(d_images[0][:lb].cpu().numpy())
            out = np.empty((lb, cfg["num_classes"]), dtype=np.float32)
            for _ in range(5):
                eng.forward_host(hx, out)
            ts = []
            for _ in range(30):
                t0 = time.perf_counter()
                eng.forward_host(hx, out)
                ts.append((time.perf_counter() - t0) * 1e3)
            ts.sort()
            inference["engine"]["latency_ms_host_to_host"][str(lb)] = {"median": round(ts[len(ts) // 2], 3), "p90": round(ts[int(len(ts) * 0.9)], 3)}
        eng.close()

    # ---- roofline of the dominant kernel: every tcgen05 GEMM launch of one step, CUDA events per launch ----
    peaks = measured_peaks()
    ctx.profile_begin()
    model.train_step(d_images[0], d_labels[0], lr)
    gemm_ms, gemm_flops, gemm_launches = ctx.profile_end()
    barrier()
    flops_per_image = pkg.train_flops_per_image(args.config)
    roofline = None
    if gemm_ms > 0:
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_tc_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.config)
        roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05/TMEM/TMA bf16 GEMM)", "achieved": round(achieved, 1),
                    "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": round(achieved / peaks["tf_sustained"], 4),
                    "peak_kind": f"{peaks['source']} cuBLAS bf16 sustained (kernel timed inside a long step)",
                    "launches_per_step": gemm_launches, "gemm_ms_per_step": round(gemm_ms, 3),
                    "gemm_share_of_step": round(gemm_ms / ms_per_step, 4), "traffic": traffic,
                    "traffic_source": "profiles/gemm_tc_traffic.json: dram bytes per launch from an ncu --set full capture of the same step "
                                      "(committed file, NOT measured in this run)" if traffic is not None else None,
                    "step_tflops": round(value / world * flops_per_image / 1e12, 1),
                    "step_frac_of_peak": round(value / world * flops_per_image / 1e12 / peaks["tf_sustained"], 4),
                    "step_frac_of_nominal_2250": round(value / world * flops_per_image / 1e12 / 2250.0, 4)}

    # ---- CPU baseline: the oracle port on the host cores, bounded sample (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            # one calibration image, then one step sized for about 12 s of CPU work (bounded at 16 images): a batch large
            # enough for the oracle's threads to spread over
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            ips1, cores, sec1 = cpu_oracle_throughput(args.config, 1, 1, 0, threads=threads)
            sample_batch = max(1, min(16, int(12.0 / max(sec1, 1e-3))))
            ips, cores, sec = cpu_oracle_throughput(args.config, sample_batch, 1, 0, threads=threads)
            cpu = {"value": round(ips, 4), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"1 step x {sample_batch} image(s), vit-{args.config}, fp32 fwd+bwd+AdamW, {sec:.1f} s "
                             f"(after a {sec1:.1f} s calibration image)"}
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    # ---- strong scaling (BASELINE's "batch 1024" read as the GLOBAL batch): same model, 1024 / N images per rank ----
    strong = None
    if world > 1 and not args.no_strong and args.config in ("b16", "s16") and per_gpu % world == 0:
        sb = per_gpu // world
        model.set_dloss_scale(1.0 / per_gpu)
        xs = [d_images[i][:sb].contiguous() for i in range(2)]
        ys = [d_labels[i][:sb].contiguous() for i in range(2)]
        for s_ in range(3):
            model.train_step(xs[s_ % 2], ys[s_ % 2], lr)
        barrier()
        e0.record()
        for s_ in range(args.steps):
            model.train_step(xs[s_ % 2], ys[s_ % 2], lr)
        e1.record()
        barrier()
        ms_strong = max_over_ranks(e0.elapsed_time(e1))
        strong = {"global_batch": per_gpu, "per_gpu_batch": sb, "value": round(per_gpu * args.steps / (ms_strong / 1e3), 2), "unit": UNIT,
                  "ms_per_step": round(ms_strong / args.steps, 3)}
        model.set_dloss_scale(1.0 / global_batch)

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.mode, "data": "synthetic",
                "config": {"workload": f"vit-{args.config} {img}x{img} train step (fwd+bwd+AdamW), per-GPU batch {per_gpu}",
                           "global_batch": global_batch, "per_gpu_batch": per_gpu, "tokens": (img // cfg['patch_size']) ** 2 + 1,
                           "parallelism": f"dp{world}" + ("+zero1" if args.zero1 else ""), "grad_exchange": args.comm if world > 1 else None,
                           "train_gflop_per_image": round(flops_per_image / 1e9, 3),
                           "l2_policy": "inputs larger than L2 (activations are GBs per step); no flush needed",
                           "final_loss": round(loss, 5)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "inference_forward": inference, "strong_scaling": strong,
                "optimizer_state_bytes_per_rank": model.optimizer_state_bytes,
                "step_graph_replays": model.step_graph_replays}
        emit(line)
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
