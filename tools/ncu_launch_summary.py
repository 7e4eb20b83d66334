"""Summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel: launches, ms per forward, share.
usage: python tools/ncu_launch_summary.py launches.csv [forwards]"""
import collections
import csv
import sys

fw = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
agg = collections.OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)
    a = agg.setdefault(r["Kernel Name"][:80], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches (%d fwd) | ms per forward | share |\n|---|---|---|---|" % fw)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1] / fw:.3f} | {100 * a[1] / tot:.1f}% |")
print(f"\nTotal {tot / fw:.2f} ms per forward.")
