"""Host-side planning (include/vitrs.h, "planning on the host"): the sizing and routing arithmetic the library itself uses,
checked without a device.

vitrs_model_footprint / vitrs_infer_footprint are built from the functions vitrs_model_create and vitrs_infer_create allocate
with (param_sizes_of, arena_layout, zplan_sizes: csrc/model.cu; workspace_layout: csrc/infer.cu); vitrs_gemm_plan is the
function gemm_tc_bf16 (csrc/gemm_tc.cu) takes its tile / CTA-pair / split-K / SIMT-fallback decisions from.  The tests restate
the documented rules independently (DESIGN.md sections 3-4; the tensor lists of rusty_vit.rs:105-122 and :150-174) and pin what
every GEMM of the BASELINE.json configs is routed to.
"""
import random

import pytest

SM = 148
TRAIN = (("b16", 1024), ("s16", 1024), ("ti16", 256), ("b8", 256))


def dims(vitrs, cfg):
    d = vitrs.CONFIGS[cfg] if isinstance(cfg, str) else cfg
    t = (d["image_size"] // d["patch_size"]) ** 2 + 1
    return d, t, d["channels"], d["num_layers"], d["num_heads"], d["num_classes"], 3 * d["patch_size"] ** 2


def param_count(vitrs, cfg):
    """ParameterTensors (rusty_vit.rs:105-122) with the ViT tensors of DEVIATIONS D7."""
    _, t, c, l, _, v, k = dims(vitrs, cfg)
    return c * k + c + c + t * c + l * (2 * c + 3 * c * c + 3 * c + c * c + c + 2 * c + 4 * c * c + 4 * c + 4 * c * c + c) + 2 * c + v * c + v


def arena_bytes(vitrs, cfg, batch, mode, grad):
    """ActivationTensors (rusty_vit.rs:150-174, batch factor restored): fp32 statistics / score tensors / head, everything else
    in the mode's element type; production mode never materialises preatt / att / attproj / fcproj and keeps only the head
    tensors in its gradient arena; every view starts on a 256-byte boundary."""
    _, t, c, l, nh, v, _ = dims(vitrs, cfg)
    per = dict(encoded=t * c, ln1=l * t * c, ln1_mean=l * t, ln1_rstd=l * t, qkv=l * t * 3 * c, atty=l * t * c, preatt=l * nh * t * t,
               att=l * nh * t * t, attproj=l * t * c, residual2=l * t * c, ln2=l * t * c, ln2_mean=l * t, ln2_rstd=l * t,
               fch=l * t * 4 * c, fch_gelu=l * t * 4 * c, fcproj=l * t * c, residual3=l * t * c, lnf=c, lnf_mean=1, lnf_rstd=1,
               logits=v, probs=v, losses=1)
    head = ("lnf", "lnf_mean", "lnf_rstd", "logits", "probs", "losses")
    f32_always = ("ln1_mean", "ln1_rstd", "ln2_mean", "ln2_rstd", "preatt", "att") + head
    total = 0
    for name in vitrs.ACT_NAMES:
        elem = 4 if (mode == vitrs.MODE_F32 or name in f32_always) else 2
        if mode == vitrs.MODE_BF16 and (name in ("preatt", "att", "attproj", "fcproj") or (grad and name not in head)):
            elem = 0
        total += (per[name] * batch * elem + 255) // 256 * 256
    return total


@pytest.mark.parametrize("cfg,batch", TRAIN + (("tiny", 8),))
@pytest.mark.parametrize("mode", ["bf16", "f32"])
def test_model_footprint_matches_the_documented_layout(vitrs, cfg, batch, mode):
    mode = vitrs.MODE_BF16 if mode == "bf16" else vitrs.MODE_F32
    f = vitrs.model_footprint(cfg, batch, mode)
    d, t, c, l, nh, v, k = dims(vitrs, cfg)
    n = param_count(vitrs, cfg)
    assert f["num_parameters"] == n
    assert (f["weights_f32"], f["grads_f32"], f["adam_moments"]) == (4 * n, 4 * n, 8 * n)
    assert f["weights_bf16"] == (2 * n if mode == vitrs.MODE_BF16 else 0)
    assert f["zero1_master_shard"] == 0 and f["exchange_buffer"] == 0
    assert f["activations"] == arena_bytes(vitrs, cfg, batch, mode, False)
    extra = 2 * batch * t * c * 2 + batch * t * 4 * c * 2 + batch * v * 4 + batch * c * 4 if mode == vitrs.MODE_BF16 else 0
    assert f["activation_grads"] == arena_bytes(vitrs, cfg, batch, mode, True) + extra
    esz = 2 if mode == vitrs.MODE_BF16 else 4
    assert f["workspace"] == 4 * l * batch * nh * t + 8 * batch * c + esz * batch * t * k + 8
    assert f["staging"] == 2 * (batch * 3 * d["image_size"] ** 2 * 4 + batch * 4)
    assert f["total"] == sum(x for key, x in f.items() if key not in ("num_parameters", "total", "train_flops_per_image"))
    assert f["train_flops_per_image"] == vitrs.train_flops_per_image(cfg)


def test_headline_configs_fit_a_b200(vitrs):
    """180 GB of HBM3e per GPU: the bench configurations and what is left above them."""
    hbm = 180 * 10 ** 9
    for cfg, batch in TRAIN:
        f = vitrs.model_footprint(cfg, batch)
        assert f["total"] < 0.5 * hbm, (cfg, f["total"])
        assert vitrs.max_batch_for(cfg) >= 2 * batch
    # ViT-B/16 at batch 1024: ~60 GB of bf16 activations (preatt / att alone would be 46 GB in bf16)
    f = vitrs.model_footprint("b16", 1024)
    assert 59.5e9 < f["activations"] < 60.5e9
    fv = vitrs.model_footprint("b16", 64, vitrs.MODE_F32)  # verify mode keeps both arenas whole, scores included
    assert fv["activations"] == fv["activation_grads"] > 11e9
    # max_batch_for is the largest batch that fits, exactly
    b = vitrs.max_batch_for("b16", hbm_bytes=40 * 10 ** 9, reserve=0)
    assert vitrs.model_footprint("b16", b)["total"] <= 40e9 < vitrs.model_footprint("b16", b + 1)["total"]


@pytest.mark.parametrize("world", [2, 8])
def test_footprint_data_parallel_and_zero1(vitrs, world):
    n = param_count(vitrs, "b16")
    rep = vitrs.model_footprint("b16", 1024, world=world)
    z = vitrs.model_footprint("b16", 1024, world=world, zero1=True)
    part = vitrs.zero_partition("b16", world)
    z_size = sum(zl for _, zl, _, _ in part)
    shards = sum(sh for _, _, _, sh in part)
    assert rep["exchange_buffer"] == z["exchange_buffer"] == 2 * z_size >= 2 * n
    assert rep["adam_moments"] == 8 * n and rep["zero1_master_shard"] == 0
    assert z["zero1_master_shard"] == 4 * (shards + 4)
    small = n - sum(cnt for bucket in vitrs.grad_buckets("b16", with_kind=True) for _, cnt, big in bucket if big)
    assert 8 * (shards + 4) + 8 * small <= z["adam_moments"] <= 8 * (shards + 4) + 8 * (small + 16)
    # master weights + both moments of the GEMM weight matrices are cut into `world` shards; the small
    # tensors keep replicated moments (their master weights are the flat fp32 buffer itself)
    state = z["zero1_master_shard"] + z["adam_moments"]
    assert abs(state - (12 * (n - small) / world + 8 * small)) < 1e5
    with pytest.raises(vitrs.VitrsError):
        vitrs.model_footprint("b16", 8, vitrs.MODE_F32, world=world, zero1=True)  # ZeRO-1 is a production-mode feature


def test_footprint_rejects_what_model_create_rejects(vitrs):
    bad = dict(vitrs.CONFIGS["tiny"])
    for key, val in (("patch_size", 5), ("channels", 60), ("num_heads", 5), ("num_layers", 0), ("num_layers", 63), ("num_classes", 0)):
        with pytest.raises(vitrs.VitrsError):
            vitrs.model_footprint(dict(bad, **{key: val}), 4)
    with pytest.raises(vitrs.VitrsError):
        vitrs.model_footprint("tiny", 0)


def test_infer_footprint(vitrs):
    ws, stage = vitrs.infer_footprint("b16", 1024)
    assert ws == 3434184704  # 3.4 GB against the training arena's 59.9 GB
    assert stage == 1024 * 3 * 224 * 224 * 4
    assert ws < vitrs.model_footprint("b16", 1024)["activations"] / 15
    prev = 0
    for b in (1, 8, 64, 256, 1024):
        w, _ = vitrs.infer_footprint("b16", b)
        assert w > prev
        prev = w
    # ping-pong buffers: independent of the depth of the model
    deep = dict(vitrs.CONFIGS["b16"], num_layers=24)
    assert vitrs.infer_footprint(deep, 64) == vitrs.infer_footprint("b16", 64)


# ---- GEMM routing ---------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("cfg,batch", TRAIN)
def test_every_gemm_of_the_bench_configs_runs_on_cta_pairs(vitrs, cfg, batch):
    """At the bench batch sizes every bf16 GEMM of the step is the persistent tcgen05 kernel on [256 x 256] CTA-pair tiles with a
    six-stage ring; the one exception is recorded in DESIGN.md section 7 (ViT-B/8's patch embedding, K = 192)."""
    for name, M, N, K, a_mn, b_mn, epi in vitrs.step_gemms(cfg, batch):
        p = vitrs.gemm_plan(M, N, K, a_mn, b_mn, epi, SM)
        if (cfg, name) == ("b8", "patch"):
            assert p["kernel"] == "simt" and (p["tile_m"], p["tile_n"]) == (64, 64), p
            lifted = vitrs.gemm_plan(M, N, K, a_mn, b_mn, epi, SM, vitrs.PLAN_PATCH_TC)
            assert (lifted["kernel"], lifted["cta_group"], lifted["tile_n"], lifted["k_blocks_per_split"]) == ("tcgen05", 2, 256, 3)
            continue
        assert (p["kernel"], p["tile_m"], p["tile_n"], p["cta_group"], p["stages"]) == ("tcgen05", 256, 256, 2, 6), (name, p)
        assert p["tiles"] == -(-M // 256) * -(-N // 256)
        assert p["grid"] % 2 == 0 and 2 <= p["grid"] <= SM
        kb = -(-K // 64)
        if epi != vitrs.EPI_ACCUM_F32:
            assert p["splits"] == 1 and p["k_blocks_per_split"] == kb and p["grid"] == SM, (name, p)
        else:
            # split-K: every split is non-empty, together they cover K, and one wave covers >= 92 % of the SM pairs unless the
            # cap (32 splits, >= 16 K blocks each) is reached first
            assert (p["splits"] - 1) * p["k_blocks_per_split"] < kb <= p["splits"] * p["k_blocks_per_split"], (name, p)
            units = p["tiles"] * p["splits"]
            waves = -(-units // (SM // 2))
            assert units / (waves * (SM // 2)) >= 0.92 or p["splits"] == min(32, kb // 16), (name, p)
            assert p["grid"] == 2 * min(units, SM // 2)


def test_small_problems_take_small_tiles(vitrs):
    """Inference at batch 1-64 (M = 197 .. 12 608 rows): tiles shrink as far as that puts more CTAs on the chip; the
    parity-config tests pin the headline tiles with VITRS_GEMM_NO_SMALL."""
    c = 768
    p1 = vitrs.gemm_plan(197, 3 * c, c, epilogue=vitrs.EPI_BIAS)  # batch 1: 2 x 9 pair tiles would use 36 SMs
    assert (p1["tile_m"], p1["tile_n"], p1["cta_group"]) == (128, 128, 1) and p1["grid"] == 2 * 18
    p8 = vitrs.gemm_plan(8 * 197, c, c, epilogue=vitrs.EPI_BIAS_RESIDUAL)
    assert (p8["tile_m"], p8["tile_n"], p8["cta_group"], p8["stages"]) == (128, 128, 1, 6)
    p64 = vitrs.gemm_plan(64 * 197, 3 * c, c, epilogue=vitrs.EPI_BIAS)
    assert (p64["tile_m"], p64["tile_n"], p64["cta_group"]) == (256, 256, 2) and p64["grid"] == SM
    keep = vitrs.gemm_plan(197, 3 * c, c, epilogue=vitrs.EPI_BIAS, flags=vitrs.PLAN_NO_SMALL)
    assert (keep["tile_m"], keep["tile_n"], keep["cta_group"], keep["grid"]) == (256, 256, 2, 2 * 9)
    single = vitrs.gemm_plan(201728, 3 * c, c, epilogue=vitrs.EPI_BIAS, flags=vitrs.PLAN_SINGLE_CTA)
    assert (single["tile_m"], single["tile_n"], single["cta_group"], single["stages"], single["grid"]) == (128, 256, 1, 4, SM)
    # the weight gradients fill the chip by splitting K instead, whatever their extent
    dw = vitrs.gemm_plan(192, 192, 50432, 1, 1, vitrs.EPI_ACCUM_F32)
    assert (dw["tile_m"], dw["tile_n"], dw["cta_group"], dw["splits"]) == (256, 256, 2, 32)
    # the patch-embedding epilogue exists for K-major operands only
    assert vitrs.gemm_plan(1024, 768, 768, 0, 1, vitrs.EPI_PATCH)["kernel"] == "simt"
    assert vitrs.gemm_plan(201728, 768, 768, epilogue=vitrs.EPI_PATCH)["kernel"] == "tcgen05"
    assert vitrs.gemm_plan(1570, 768, 192, epilogue=vitrs.EPI_PATCH)["tile_n"] == 128  # patch 8 at test batch sizes: single-CTA tiles


def test_shapes_the_tensor_core_kernel_cannot_take_go_to_simt(vitrs):
    """TMA needs 16-byte inner extents (8 bf16): everything else is the SIMT kernel's, [32 x 32] tiles when [64 x 64] would
    leave SMs without a CTA."""
    for M, N, K, a_mn, b_mn in ((100, 10, 64, 0, 0), (64, 64, 4, 0, 0), (100, 64, 64, 1, 0), (64, 100, 64, 0, 1), (64, 64, 60, 0, 0)):
        p = vitrs.gemm_plan(M, N, K, a_mn, b_mn)
        assert p["kernel"] == "simt" and p["stages"] == 0 and p["splits"] == 1, (M, N, K, p)
    assert vitrs.gemm_plan(256, 1004, 768)["tile_m"] == 32  # 4 x 16 tiles of 64 < 148 SMs
    assert vitrs.gemm_plan(4096, 1004, 768)["tile_m"] == 64
    assert vitrs.gemm_plan(100, 96, 64, epilogue=vitrs.EPI_ROWDOT)["kernel"] == "simt"  # head slices are 64 columns wide
    for bad in ((0, 8, 8), (8, 0, 8), (8, 8, 0)):
        with pytest.raises(vitrs.VitrsError):
            vitrs.gemm_plan(*bad)


def test_plan_invariants_on_random_shapes(vitrs):
    rng = random.Random(1337)
    for _ in range(400):
        M, N, K = (8 * rng.randint(1, 4000) for _ in range(3))
        epi = rng.choice([vitrs.EPI_NONE, vitrs.EPI_BIAS, vitrs.EPI_BIAS_GELU, vitrs.EPI_ACCUM_F32])
        a_mn = b_mn = int(epi == vitrs.EPI_ACCUM_F32)
        sm = rng.choice([148, 132, 64])
        p = vitrs.gemm_plan(M, N, K, a_mn, b_mn, epi, sm, rng.choice([0, vitrs.PLAN_NO_SMALL, vitrs.PLAN_SINGLE_CTA]))
        assert p["kernel"] == "tcgen05"
        assert (p["tile_m"], p["tile_n"], p["cta_group"], p["stages"]) in ((256, 256, 2, 6), (128, 256, 1, 4), (128, 128, 1, 6))
        assert p["tiles"] == -(-M // p["tile_m"]) * -(-N // p["tile_n"])
        kb = -(-K // 64)
        assert (p["splits"] - 1) * p["k_blocks_per_split"] < kb <= p["splits"] * p["k_blocks_per_split"]
        assert p["splits"] == 1 or epi == vitrs.EPI_ACCUM_F32
        slots = sm // p["cta_group"]
        assert p["grid"] == p["cta_group"] * min(p["tiles"] * p["splits"], slots)
        # 2 accumulators of tile_n fp32 columns fit the 512 columns of tensor memory; the ring + staging fit 227 KB of shared memory
        assert 2 * p["tile_n"] <= 512
        stage_bytes = 128 * 64 * 2 + p["tile_n"] // p["cta_group"] * 64 * 2
        assert p["stages"] * stage_bytes + 8 * 32 * 128 + 512 + 1024 <= 227 * 1024


# ---- the plan against what a GPU ran ------------------------------------------------------------------------------------------


@pytest.mark.parametrize("cfg,batch,fname", [("b16", 1024, "r2_launches_vitb16_b1024.csv"), ("ti16", 256, "r2_launches_vitti16_b256.csv")])
def test_plan_reproduces_the_committed_ncu_launch_lists(vitrs, cfg, batch, fname):
    """profiles/r2_launches_*.csv are ncu launch lists of one training step on a B200: the kernel instantiation
    (gemm_tc_kernel<tile_n, stages, A MN-major, B MN-major, CTAs per tile, patch epilogue>) and the grid of every GEMM launch, in
    order, must be what the host-side plan says for the step's GEMM sequence."""
    import csv
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", fname)
    rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
    ki, gi = rows[0].index("Kernel Name"), rows[0].index("Grid Size")
    ran = []
    for r in rows[1:]:
        m = re.search(r"gemm_tc_kernel<([^>]*)>", r[ki])
        if m:
            ran.append((tuple(int(x) for x in m.group(1).split(",")), int(re.match(r"\((\d+)", r[gi]).group(1))))
    first = next(i for i, (targs, _) in enumerate(ran) if targs[5] == 1)  # the step starts at its patch-embedding GEMM
    want = []
    for name, M, N, K, a_mn, b_mn, epi in vitrs.step_gemms(cfg, batch):
        p = vitrs.gemm_plan(M, N, K, a_mn, b_mn, epi, SM)
        assert p["kernel"] == "tcgen05", name
        want.append(((p["tile_n"], p["stages"], a_mn, b_mn, p["cta_group"], int(epi == vitrs.EPI_PATCH)), p["grid"]))
    got = ran[first:first + len(want)]
    assert len(got) == len(want) == 2 + 12 * vitrs.CONFIGS[cfg]["num_layers"]
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, (i, vitrs.step_gemms(cfg, batch)[i][0], g, w)
