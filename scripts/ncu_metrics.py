"""Print the headline metrics of every kernel in an ncu report (ncu -i X --page raw --csv)."""
import csv, subprocess, sys
out = open(sys.argv[1]).read() if sys.argv[1].endswith(".csv") else subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("smsp__inst_executed.sum", "inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf")]
for d in data:
    name = d[idx["Kernel Name"]][:70]
    parts = []
    for k, short in want:
        if k in idx:
            parts.append(f"{short}={d[idx[k]]}{units[idx[k]] if units[idx[k]] not in ('%', 'inst', 'register/thread', '') else ''}")
    print(name, "grid", d[idx["Grid Size"]] if "Grid Size" in idx else "", "\n   ", "  ".join(parts))
