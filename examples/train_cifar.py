#!/usr/bin/env python
"""Train a ViT on CIFAR-layout record files through libvitrs.so — the loop the reference leaves to its caller
(`ViT::forward` / `backward` / `optimizer_step`, rusty_vit.rs:269-449, :949): loader -> step -> checkpoint -> serving check.

    python examples/train_cifar.py --data data_batch_1.bin ... --config ti16 --batch 256 --epochs 2
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_cifar.py --data ... --zero1

One process per GPU.  Every rank opens the same files with the same seed and takes its own batches of each round
(vitrs_loader_open_transform: random crop + flip, bilinear resize to the model's input on `--workers` host threads); gradients are
exchanged per bucket inside `backward` (bf16 NCCL all-reduce, or reduce-scatter with `--zero1`).  Needs a B200: there is no CPU path.
`--plan` prints the memory footprint and the GEMM routing of the chosen configuration and exits (no GPU needed).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", nargs="+", default=[], help="CIFAR binary record files (label byte(s) + 3 x S x S uint8)")
    ap.add_argument("--record-size", type=int, default=32)
    ap.add_argument("--label-bytes", type=int, default=1, help="1: CIFAR-10, 2: CIFAR-100 (the fine label is used)")
    ap.add_argument("--config", default="ti16", choices=["tiny", "ti16", "s16", "b16", "b8"])
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--batch", type=int, default=256, help="per GPU")
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--weight-decay", type=float, default=0.05)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--zero1", action="store_true")
    ap.add_argument("--checkpoint", default="vit.ckpt")
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--plan", action="store_true")
    args = ap.parse_args()

    pkg = ge.load_package()
    cfg = dict(pkg.CONFIGS[args.config], num_classes=args.classes)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))

    if args.plan:  # host arithmetic only
        f = pkg.model_footprint(cfg, args.batch, world=world, zero1=args.zero1)
        print(json.dumps({k: (round(v / 2 ** 30, 3) if k not in ("num_parameters", "train_flops_per_image") else v) for k, v in f.items()}))
        print("largest batch that fits 180 GB:", pkg.max_batch_for(cfg, world=world, zero1=args.zero1))
        seen = set()
        for name, *shape in pkg.step_gemms(cfg, args.batch):
            if name not in seen:
                seen.add(name)
                print(f"{name:12s} {shape[:3]} ->", pkg.gemm_plan(*shape))
        return

    if not args.data:
        ap.error("--data is required (CIFAR binary record files)")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    ctx = pkg.Context(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    if args.resume and os.path.exists(args.checkpoint):
        model = pkg.ViT.build_from_checkpoint(args.checkpoint, max_batch=args.batch, ctx=ctx)
    else:
        model = pkg.ViT(cfg, max_batch=args.batch, mode=pkg.MODE_BF16, seed=1337, init_mode=1, ctx=ctx)
    model.set_dloss_scale(1.0 / (args.batch * world))  # mean over the GLOBAL batch (rusty_vit.rs:366 generalised)
    if args.zero1:
        model.enable_zero1()
    model.set_input_norm((0.4914, 0.4822, 0.4465), (0.2470, 0.2435, 0.2616))  # CIFAR-10 channel statistics, applied on the device

    loader = pkg.RecordLoader(args.data, image_size=args.record_size, batch=args.batch, label_bytes=args.label_bytes, shuffle=True,
                              seed=1337, drop_last=True, ctx=ctx, rank=rank, world=world, out_size=cfg["image_size"],
                              workers=args.workers, random_flip=True, crop_pad=args.record_size // 8)
    steps = args.epochs * loader.batches_per_epoch
    if rank == 0:
        print(f"{loader.num_records} records, {loader.batches_per_epoch} steps per epoch on each of {world} rank(s), "
              f"{loader.num_classes_seen} classes seen; optimiser state {model.optimizer_state_bytes / 2 ** 20:.0f} MiB per rank")
    t0 = time.perf_counter()
    for step in range(steps):
        loss, b = loader.train_step(model, args.lr, weight_decay=args.weight_decay)  # H2D of batch n + 1 overlaps step n
        if rank == 0 and (step % 50 == 0 or step == steps - 1):
            dt = time.perf_counter() - t0
            print(f"step {step:6d}  loss {loss:.4f}  {(step + 1) * b * world / dt:9.0f} images/s")
        if world > 1 and step % 200 == 0:
            ctx.comm_async_error()  # raises when a peer or the fabric has failed
    # every rank calls it (under ZeRO-1 the shards are gathered collectively); only rank 0's bytes are kept
    model.save_checkpoint(args.checkpoint if rank == 0 else os.devnull)

    if rank == 0:  # serving check: the graph-replayed forward on a few records, host buffers in, logits out
        import numpy as np
        eng = pkg.InferenceEngine(model, max_batch=64)
        check = pkg.RecordLoader(args.data[:1], image_size=args.record_size, batch=64, label_bytes=args.label_bytes, shuffle=False,
                                 pinned=False, out_size=cfg["image_size"], workers=args.workers)
        img, lab, _ = check.next()
        logits = eng.forward_host_u8(np.ascontiguousarray(img))
        print(f"top-1 on the first {len(lab)} training records: {(logits.argmax(axis=1) == lab).mean():.3f}")
        check.close(); eng.close()
    loader.close(); model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
