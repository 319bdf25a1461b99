"""The committed GPU evidence (profiles/) is internally consistent and agrees with what the library computes on the host.

Nothing here runs a kernel: each bench line of profiles/r2_configs.json (printed by bench.py on a B200) is re-derived from its own
fields — throughput from the step time, the roofline fraction from achieved / peak, the end-to-end byte counts from the batch
shape, the algorithmic flops from the model configuration — and compared with the host-side footprint / plan functions, so a
hand-edited or stale number does not survive."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = json.load(open(os.path.join(ROOT, "profiles", "r2_configs.json")))["lines"]


def cfg_of(line):
    return line["config"]["workload"].split()[0].replace("vit-", "")


@pytest.mark.parametrize("key", sorted(LINES))
def test_bench_line_is_self_consistent(vitrs, key):
    line = LINES[key]
    cfg_name = cfg_of(line)
    cfg = vitrs.CONFIGS[cfg_name]
    world, per_gpu = line["n_gpus"], line["config"]["per_gpu_batch"]
    assert line["metric"] == "ViT train images/sec (fwd+bwd+AdamW)" and line["unit"] == "images/s" and line["data"] == "synthetic"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["warmup"] >= 3 and line["steps"] >= 8
    assert line["config"]["global_batch"] == per_gpu * world
    assert line["config"]["tokens"] == (cfg["image_size"] // cfg["patch_size"]) ** 2 + 1
    # throughput = images per step / step time
    assert abs(line["value"] - per_gpu * world / (line["ms_per_step"] / 1e3)) <= 2e-3 * line["value"]
    flops = vitrs.train_flops_per_image(cfg_name)
    assert abs(line["config"]["train_gflop_per_image"] - flops / 1e9) < 1e-3
    # clocks were sampled during the timed region and show no thermal or hardware slowdown (a power cap is noted, not rejected)
    if line["steps"] * line["ms_per_step"] > 500:  # (the 200 ms sampler needs a timed region of some length: not the 24 ms tiny run)
        assert line["clocks"]["sm_mhz"] and line["clocks"]["sm_max_mhz"] >= line["clocks"]["sm_mhz"]
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    steps_run = line["steps"]
    if line["dtype"] == "bf16":
        r = line["roofline"]
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 2e-3
        assert abs(r["step_tflops"] - line["value"] / world * flops / 1e12) <= 2e-3 * r["step_tflops"] + 0.06
        assert abs(r["gemm_share_of_step"] - r["gemm_ms_per_step"] / line["ms_per_step"]) < 2e-3
        # every tcgen05 GEMM of the step was timed: the count the host-side plan gives for this configuration
        planned = sum(1 for g in vitrs.step_gemms(cfg_name, per_gpu) if vitrs.gemm_plan(*g[1:])["kernel"] == "tcgen05")
        assert r["launches_per_step"] in (planned, planned + 1), (r["launches_per_step"], planned)  # (+1: ViT-B/8 lines predate the SIMT routing of its patch GEMM)
        # the achieved rate is the algorithmic GEMM work of the step over the measured GEMM time
        gemm_flops = sum(2.0 * M * N * K for _, M, N, K, *_ in vitrs.step_gemms(cfg_name, per_gpu))
        assert abs(r["achieved"] - gemm_flops / (r["gemm_ms_per_step"] / 1e3) / 1e12) <= 0.02 * r["achieved"]
        assert line["optimizer_state_bytes_per_rank"] == 12 * vitrs.model_footprint(cfg_name, per_gpu)["num_parameters"]
        assert line["gpu_launches"] >= steps_run * 200
    if line.get("e2e"):
        e = line["e2e"]
        img = cfg["image_size"]
        assert e["h2d_bytes_per_step"] == world * (per_gpu * 3 * img * img * 4 + per_gpu * 4) and e["d2h_bytes_per_step"] == 4 * world
        assert e["value"] <= 1.02 * line["value"]  # host buffers cannot be faster than resident ones (2 % timing noise)
        assert abs(e["value"] - per_gpu * world / (e["ms_per_step"] / 1e3)) <= 2e-3 * e["value"]
        assert e["uint8_images"]["h2d_bytes_per_step"] == world * (per_gpu * 3 * img * img + per_gpu * 4)
    eng = (line.get("inference_forward") or {}).get("engine")
    if eng:
        assert eng["workspace_bytes"] == vitrs.infer_footprint(cfg_name, per_gpu)[0]
        assert eng["value"] > 2.5 * line["value"] / world  # a forward is a third of a training step's flops
    if line.get("cpu_baseline"):
        assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "image" in line["cpu_baseline"]["sample"]
    if line.get("strong_scaling"):
        s = line["strong_scaling"]
        assert s["per_gpu_batch"] * world == s["global_batch"] == per_gpu
        assert abs(s["value"] - s["global_batch"] / (s["ms_per_step"] / 1e3)) <= 2e-3 * s["value"]


def test_quoted_headline_figures_match_the_committed_lines():
    """DESIGN.md section 7 quotes these; a figure in the text that the evidence file does not hold is a defect of the text."""
    text = open(os.path.join(ROOT, "DESIGN.md")).read()
    for key, fmt in (("cfg_b16", "{:,.0f}"), ("cfg_ti16", "{:,.0f}"), ("scale_b16_n8", "{:,.0f}"), ("scale_s16_n8", "{:,.0f}"),
                     ("scale_b8_n8", "{:,.0f}"), ("scale_b16_n2", "{:,.0f}")):
        quoted = fmt.format(LINES[key]["value"]).replace(",", " ")
        assert quoted in text, (key, quoted)
