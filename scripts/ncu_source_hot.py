"""Top SASS instructions by warp-stall samples from `ncu -i X --page source --csv`, with stall-reason breakdown."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[idx["# Samples"]]) for r in data)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[idx[h]]) for r in data) for h in reasons}
print("total samples", tot)
print("by reason:", ", ".join(f"{k[6:]}={v} ({100*v/tot:.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    r = data[i]
    rs = sorted(((int(r[idx[h]]), h[6:]) for h in reasons), reverse=True)[:2]
    print(f"{i:5d} {int(r[idx['# Samples']]):6d} {100*int(r[idx['# Samples']])/tot:5.1f}%  {r[idx['Source']].strip()[:70]:70s} {rs}")
