"""Summarise an ncu launch list (the `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv` pass of scripts/gpu_round.sh) of scripts/profile_step.py:

    python scripts/summarize_launches.py gpurun_out/launches_v12.csv [--out profiles/r01_solve_traffic.json]

Takes the LAST step of the capture (everything after the last carrier RHS launch), prints per kernel class the number of
launches, the sum of the isolated (cold-cache, serialised) durations and their share of the step, the DRAM bytes, and
writes the DRAM bytes of the solve kernels -- the `traffic` figure of bench.py's roofline object."""
import argparse
import csv
import json
import re
from collections import OrderedDict

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--out", default=None)
a = ap.parse_args()

rows = []
with open(a.csv) as f:
    lines = [l for l in f if l.startswith('"')]
reader = csv.DictReader(lines)
launch = OrderedDict()
for r in reader:
    k = int(r["ID"])
    d = launch.setdefault(k, {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        d[r["Metric Name"]] = v * scale
    else:
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        d["us"] = v * scale
ids = list(launch)
rhs = [k for k in ids if "carrier_rhs" in launch[k]["name"]]
start = rhs[-1]
step = [launch[k] for k in ids if k >= start]


def cls(name):
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name[:40]


table = OrderedDict()
for d in step:
    t = table.setdefault(cls(d["name"]), {"launches": 0, "us": 0.0, "dram": 0.0})
    t["launches"] += 1
    t["us"] += d.get("us", 0.0)
    t["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
total_us = sum(t["us"] for t in table.values())
print(f"last step of the capture: {len(step)} launches, sum of isolated durations {total_us / 1e3:.3f} ms")
for k, t in table.items():
    print(f"  {k:34s} {t['launches']:4d} launches {t['us']:10.1f} us {100 * t['us'] / total_us:5.1f} %  "
          f"{t['dram'] / 1e6:10.1f} MB DRAM")
solve = {k: t for k, t in table.items() if re.search(r"level_kernel|ell_|gather", k)}
import hashlib  # noqa: E402
import os  # noqa: E402


def kernel_source_hash():
    """sha256 of the sources of the solve kernels: bench.py refuses a traffic file taken from other kernels"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    for rel in ("pecs_b200/csrc/cuda/solve_kernels.cu", "pecs_b200/csrc/cuda/solve_kernels.cuh", "pecs_b200/csrc/cuda/schur_kernels.cu"):
        h.update(open(os.path.join(root, rel), "rb").read())
    return h.hexdigest()[:16]


out = {"kernel_source_sha16": kernel_source_hash(),
       "dram_bytes_per_step": sum(t["dram"] for t in solve.values()),
       "launches": sum(t["launches"] for t in solve.values()),
       "sum_isolated_durations_ms": sum(t["us"] for t in solve.values()) / 1e3,
       "share_of_step_isolated": sum(t["us"] for t in solve.values()) / total_us,
       "source": f"{a.csv} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                 "--clock-control none), last step of the capture: forward/backward level kernels + ELL kernels of "
                 "the five solves", "per_kernel": table}
print(json.dumps({k: v for k, v in out.items() if k != "per_kernel"}, indent=1))
if a.out:
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
