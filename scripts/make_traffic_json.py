"""profiles/gemm_tc_traffic.json from the ncu --set full CSVs of one forward and one backward block.

traffic = dram__bytes_read.sum + dram__bytes_write.sum of the tcgen05 GEMM launches, averaged per launch over
the 15 block GEMMs captured (12 identical blocks make up 144 of the step's 146 GEMM launches)."""
import csv, json, sys
tot, n, rows_out = 0.0, 0, []
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    def gb(d, k):
        v = float(d[idx[k]]); u = units[idx[k]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
    for d in data:
        if "gemm_tc_kernel" not in d[idx["Kernel Name"]]:
            continue
        b = gb(d, "dram__bytes_read.sum") + gb(d, "dram__bytes_write.sum")
        tot += b; n += 1
        rows_out.append({"kernel": d[idx["Kernel Name"]].split("(")[0].strip(), "dram_bytes": b,
                         "tensor_pipe_pct": float(d[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
                         "time": d[idx["gpu__time_duration.sum"]] + units[idx["gpu__time_duration.sum"]]})
cfg = sys.argv[1]
out = {cfg: tot / n, "launches_captured": n, "unit": "bytes per launch (mean over one block's GEMM launches)", "launches": rows_out}
json.dump(out, open("profiles/gemm_tc_traffic.json", "w"), indent=1)
print(cfg, n, "launches, mean traffic", tot / n / 1e9, "GB")
