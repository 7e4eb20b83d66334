"""Parse an ncu --csv metrics log (dram bytes + duration per launch) of tools/profile_step.py and write
profiles/ncu_traffic.json: per kernel class of ONE forward, DRAM bytes per launch (read + write) and launches.
usage: python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/ncu_traffic.json"""
import collections
import csv
import json
import re
import sys

CLASS = [("window_attn_", "window_attn"), ("stem_conv2", "stem_conv2"), ("mlp_fused", "mlp_fused"), ("linear_tc", "linear"), ("layernorm_nchw", "layernorm_nchw"),
         ("MergeRows", "patch_merge_ln"), ("layernorm_rows", "layernorm"), ("stem_conv1", "stem_conv1")]
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    d["unit_" + r["Metric Name"]] = r["Metric Unit"]
def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
per = collections.OrderedDict()
items = list(launch.values())
# keep the second forward only: the last occurrence block (kernels after the last stem / first-LN launch)
starts = [i for i, d in enumerate(items) if "stem_conv1" in d["name"]]
items = items[starts[-1]:] if starts else items
after_conv2 = False
for d in items:
    cls = next((c for k, c in CLASS if k in d["name"]), None)
    if cls is None:
        continue
    if cls == "linear" and after_conv2:          # the launch right after stem conv2 is the patch conv (same GEMM kernel)
        cls = "patch_conv"
    after_conv2 = cls == "stem_conv2"
    b = to_bytes(d.get("dram__bytes_read.sum", 0), d.get("unit_dram__bytes_read.sum", "byte")) + \
        to_bytes(d.get("dram__bytes_write.sum", 0), d.get("unit_dram__bytes_write.sum", "byte"))
    p = per.setdefault(cls, {"launches": 0, "dram_bytes": 0.0})
    p["launches"] += 1
    p["dram_bytes"] += b
out = {c: {"launches_per_forward": p["launches"], "dram_bytes_per_forward": p["dram_bytes"],
           "dram_bytes_per_launch": p["dram_bytes"] / p["launches"]} for c, p in per.items()}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
