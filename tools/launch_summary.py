"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the step."""
import collections, csv, sys
path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0      # steps 0: one step per letterbox launch
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg[row["Kernel Name"]][0] += 1; agg[row["Kernel Name"]][1] += v; tot += v
if steps <= 0:
    steps = max(1, sum(n for k, (n, _) in agg.items() if "letterbox_kernel" in k))
print(f"{path}: {tot/steps:.1f} us of kernel time per step ({steps} steps, cold-cache serialised ncu timings)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{t/tot*100:5.1f}%  {t/steps:9.1f} us/step  launches/step={n//steps:3d}  {k[:100]}")
