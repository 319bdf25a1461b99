"""Data-parallel host logic on CPU, world_size 2 over gloo.

The GPU path shards the batch, uses dloss = 1/B_global on every rank and SUM-all-reduces the
gradient buffer in buckets.  Here the same arithmetic runs with the CPU oracle as the compute
and torch.distributed(gloo) as the exchange: 2-rank gradients, loss and post-AdamW weights must
equal the single-process result on the whole batch.  The bucket schedule the library uses
(vitrs_grad_bucket, host arithmetic only) is walked exactly as model.cu walks it.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as po

B_GLOBAL = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, buckets, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = po.CONFIGS["tiny"]
    images, labels = po.synthetic_batch(cfg, B_GLOBAL)
    per = B_GLOBAL // world
    sl = slice(rank * per, (rank + 1) * per)
    m = po.ViT("tiny", seed=1337, init_mode=1)
    m.set_dloss_scale(1.0 / B_GLOBAL)
    local_mean = m.forward(images[sl], labels[sl])
    m.zero_grad(); m.backward()
    grads = torch.from_numpy(m.grads_flat())  # aliases the oracle's buffer
    for bucket in buckets:                    # the library's exchange order, slice by slice
        for off, cnt in bucket:
            dist.all_reduce(grads[off:off + cnt], op=dist.ReduceOp.SUM)
    loss = torch.tensor([local_mean * per / B_GLOBAL], dtype=torch.float64)  # sum(losses)/B_global, as the GPU path reports it
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    m.update(1e-3)
    if rank == 0:
        q.put((m.grads_flat().copy(), loss.item(), m.params_flat().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_schedule_covers_every_gradient_once(vitrs):
    for name in ("tiny", "ti16", "b16"):
        buckets = vitrs.grad_buckets(name)
        cfg = vitrs.CONFIGS[name]
        assert len(buckets) == cfg["num_layers"] + 2 and all(len(b) == 12 for b in buckets[1:-1]) and len(buckets[-1]) == 2
        n = po.ViT(name).num_parameters if name == "tiny" else None
        cover = {}
        total = 0
        for b in buckets:
            for off, cnt in b:
                assert cnt > 0 and off not in cover
                cover[off] = cnt
                total += cnt
        end = 0
        for off in sorted(cover):  # contiguous, non-overlapping
            assert off == end
            end = off + cover[off]
        assert total == end and (n is None or total == n)


def test_two_rank_gradients_equal_single_process(vitrs):
    buckets = vitrs.grad_buckets("tiny")
    cfg = po.CONFIGS["tiny"]
    images, labels = po.synthetic_batch(cfg, B_GLOBAL)
    ref = po.ViT("tiny", seed=1337, init_mode=1)
    ref_loss = ref.forward(images, labels)
    ref.zero_grad(); ref.backward()
    ref_grads = ref.grads_flat().copy()
    ref.update(1e-3)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, buckets, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads, loss, params = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.abs(grads - ref_grads).max() / np.abs(ref_grads).max() < 1e-5  # summation order only
    assert abs(loss - ref_loss) < 1e-5
    d = np.abs(params - ref.params_flat())
    assert d.max() < 2e-4 and np.percentile(d, 99.9) < 1e-5


def test_zero1_partition_arithmetic(vitrs):
    """ZeRO-1 host logic (vitrs_zero_partition): the bucket regions tile the exchange buffer; a region is the bucket's big slices
    (GEMM weight matrices) padded to 8 * world and cut into world equal, 16-byte aligned shards, followed by its small slices
    (replicated) padded to 8; together the regions cover every gradient element exactly once."""
    for name in ("tiny", "ti16", "b16"):
        buckets = vitrs.grad_buckets(name, with_kind=True)
        for world in (1, 2, 4, 8):
            part = vitrs.zero_partition(name, world)
            assert len(part) == len(buckets)
            z = 0
            for (zo, zl, zb, sh), slices in zip(part, buckets):
                nb = sum(c for _, c, big in slices if big)
                ns = sum(c for _, c, big in slices if not big)
                assert zo == z and zb >= nb and zb - nb < 8 * world and zb % (8 * world) == 0 and sh * world == zb and sh % 8 == 0
                assert zl - zb >= ns and zl - zb - ns < 8
                z += zl
        big = sum(c for b in buckets for _, c, k in b if k)
        total = sum(c for b in buckets for _, c, _ in b)
        assert big / total > (0.9 if name != "tiny" else 0.5)  # what ZeRO-1 shards: the bulk of the model


def _zero_worker(rank, world, port, q):
    """Reduce-scatter + sharded AdamW + all-gather over gloo on the CPU oracle's buffers: the arithmetic of model.cu's ZeRO-1
    path (big slices sharded in Z order; small slices all-reduced and updated on every rank)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    vitrs = ge.load_package()
    cfg = po.CONFIGS["tiny"]
    images, labels = po.synthetic_batch(cfg, B_GLOBAL)
    per = B_GLOBAL // world
    sl = slice(rank * per, (rank + 1) * per)
    m = po.ViT("tiny", seed=1337, init_mode=1)
    m.set_dloss_scale(1.0 / B_GLOBAL)
    m.forward(images[sl], labels[sl]); m.zero_grad(); m.backward()
    grads, params = m.grads_flat(), m.params_flat()
    buckets, part = vitrs.grad_buckets("tiny", with_kind=True), vitrs.zero_partition("tiny", world)
    new_params = params.copy()
    hp = (1e-3, 0.9, 0.999, 1e-8, 0.01, 1)
    for slices, (zo, zl, zb, sh) in zip(buckets, part):
        big = [np.arange(o, o + c) for o, c, k in slices if k]
        small = [np.arange(o, o + c) for o, c, k in slices if not k]
        if big:
            idx = np.concatenate(big)                                          # Z order of the bucket's big part
            region = np.zeros(zb, np.float32); region[:idx.size] = grads[idx]
            # gloo has no reduce_scatter: all_reduce and keep this rank's shard (the same sums)
            dist.all_reduce(torch.from_numpy(region), op=dist.ReduceOp.SUM)
            g = region[rank * sh:(rank + 1) * sh].copy()
            pz = np.zeros(zb, np.float32); pz[:idx.size] = params[idx]
            p = pz[rank * sh:(rank + 1) * sh].copy()
            mom, var = np.zeros_like(p), np.zeros_like(p)
            po.adamw_step(p, g, mom, var, *hp)                                  # AdamW on the shard only
            gathered = [torch.zeros(sh) for _ in range(world)]
            dist.all_gather(gathered, torch.from_numpy(p))
            new_params[idx] = torch.cat(gathered).numpy()[:idx.size]
        if small:
            idx = np.concatenate(small)
            g = grads[idx].copy()
            dist.all_reduce(torch.from_numpy(g), op=dist.ReduceOp.SUM)
            p = params[idx].copy()
            mom, var = np.zeros_like(p), np.zeros_like(p)
            po.adamw_step(p, g, mom, var, *hp)                                  # replicated
            new_params[idx] = p
    if rank == 0:
        q.put(new_params)
    dist.barrier()
    dist.destroy_process_group()


def test_zero1_two_rank_step_equals_single_process(vitrs):
    cfg = po.CONFIGS["tiny"]
    images, labels = po.synthetic_batch(cfg, B_GLOBAL)
    ref = po.ViT("tiny", seed=1337, init_mode=1)
    ref.forward(images, labels); ref.zero_grad(); ref.backward(); ref.update(1e-3)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_zero_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    params = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = np.abs(params - ref.params_flat())
    assert d.max() < 2e-4 and np.percentile(d, 99.9) < 1e-5
