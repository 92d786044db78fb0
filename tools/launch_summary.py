"""tools/launch_summary.py <ncu launch list .csv> -- per-kernel summary (launches, total, mean, share of
the GPU time) of an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list."""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
agg = OrderedDict()
for r in rows:
    name = r[4].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"\(.*", "", name).replace("void ", "").strip()
    name = re.sub(r"<.*", lambda m: m.group(0) if "gather_units" in name or "gather_lists" in name or "scatter" in name else "", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14])
tot = sum(a[1] for a in agg.values())
print("kernel,launches,total_ns,mean_ns,share_of_gpu_time")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{n},{t:.0f},{t / n:.0f},{t / tot:.4f}")
