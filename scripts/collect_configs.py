"""Assemble profiles/r2_configs.json: the bench line of every BASELINE.json config measured this round (gpurun_out/cfg_*.json from
scripts/round2_evidence.sh, gpurun_out/scale_*.json from the multi-GPU visits), with the derived fraction of the measured peak."""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {"note": "one bench.py JSON line per config; single-GPU lines from scripts/round2_evidence.sh, multi-GPU lines from the scale visits "
               "(every line was printed by an unprofiled run)", "lines": {}}
def first_json(path):
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                pass
    return None
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "cfg_*.json")) + glob.glob(os.path.join(ROOT, "gpurun_out", "scale_*.json"))):
    d = first_json(path)
    if d:
        key = os.path.basename(path)[:-5]
        out["lines"][key] = d
        r = d.get("roofline") or {}
        print(f"{key:24s} {d['value']:10.1f} img/s  n={d['n_gpus']}  {d['ms_per_step']:8.2f} ms  step {r.get('step_frac_of_peak')} of sustained, gemm {r.get('frac')}")
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_configs.json"), "w"), indent=1)
