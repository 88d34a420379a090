"""profiles/r2_launches_step.csv (ncu launch list of the bench step) -> profiles/r2_umma_traffic.json: DRAM bytes per launch of the dense
CTA-pair kernel and its share of the step's device time (bench.py reads the file for `roofline.traffic`)."""
import collections
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "profiles", "r2_launches_step.csv")
rows = list(csv.DictReader([l for l in open(src) if l.startswith('"')]))
per = collections.defaultdict(dict)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
for r in rows:
    per[r["ID"]]["name"] = r["Kernel Name"]
    per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
total_us = sum(p.get("gpu__time_duration.sum", 0.0) for p in per.values())
pair = [p for p in per.values() if "pair2" in p["name"]]
out = {"model": "1b",
       "dram_bytes_per_launch": sum(p["dram__bytes_read.sum"] + p["dram__bytes_write.sum"] for p in pair) / len(pair),
       "launches": len(pair),
       "avg_us_per_launch_under_ncu": sum(p["gpu__time_duration.sum"] for p in pair) / len(pair),
       "kernel_share_of_step_device_time": sum(p["gpu__time_duration.sum"] for p in pair) / total_us,
       "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over the launches of the bench steps "
                 "(profiles/r2_launches_step.csv; scripts/derive_traffic.py)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_umma_traffic.json"), "w"), indent=1)
print(out)
