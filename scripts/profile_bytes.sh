#!/bin/bash
# DRAM bytes and time of every kernel of one training step (ncu, two metrics): the memory roofline of the small configs,
# whose GEMMs (K = 192 / 384) are HBM-bound.  CFG = config, output gpurun_out/bytes_<cfg>.csv
mkdir -p gpurun_out
CFG=${CFG:-ti16}
CMD="python bench.py --config $CFG --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
VITRS_NO_STEP_GRAPH=1 $CMD > gpurun_out/plain_bytes.log 2>&1 || { tail -5 gpurun_out/plain_bytes.log; exit 1; }
# 4 priming + 3 warm-up steps of 231 launches precede the timed step (+ ~20 set-up kernels)
VITRS_NO_STEP_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s ${SKIP:-1637} -c 232 --csv \
  --log-file gpurun_out/bytes_$CFG.csv $CMD > gpurun_out/ncu_bytes.log 2>&1
echo "bytes=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.DictReader(l for l in open("gpurun_out/bytes_$CFG.csv") if not l.startswith("=="))]
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
def val(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
for r in rows:
    k = r["Kernel Name"].split("(")[0][-60:]
    if r["Metric Name"] == "gpu__time_duration.sum": agg[k][0] += val(r); agg[k][2] += 1
    else: agg[k][1] += val(r)
T = sum(a[0] for a in agg.values()); B = sum(a[1] for a in agg.values())
print(f"$CFG: {sum(a[2] for a in agg.values())} launches, {T/1e3:.2f} ms (isolated clocks), {B/1e9:.2f} GB of DRAM traffic -> {B/T/1e6:.2f} TB/s average")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"  {a[0]:9.1f} us  {a[1]/1e9:7.3f} GB  {a[1]/max(a[0],1e-9)/1e6:5.2f} TB/s  n={a[2]:3d}  {k}")
PY
