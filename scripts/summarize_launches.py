"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST
bench step (the launches between the last two adam_kernel launches).
usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launch_list_step_summary.txt"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
  lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
  if r.get("Metric Name") == "gpu__time_duration.sum":
    rows.append((r["Kernel Name"], float(r["Metric Value"]) * (1e-6 if r["Metric Unit"] == "ns" else 1e-3)))
adam = [i for i, (k, _) in enumerate(rows) if "adam_kernel" in k or "adam_dev_kernel" in k or "adam_guarded_kernel" in k]
if len(adam) >= 2:
  rows = rows[adam[-2] + 1:adam[-1] + 1]
tot = sum(ms for _, ms in rows)
agg = collections.OrderedDict()
for k, ms in rows:
  k = re.sub(r"^void ", "", k)
  k = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", k)
  k = re.sub(r"\(.*$", "", k)
  a = agg.setdefault(k, [0.0, 0])
  a[0] += ms; a[1] += 1
print("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches): one bench.py step")
print(f"launches in step: {len(rows)}   sum of kernel times: {tot:.3f} ms")
for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
  print(f"  {ms:8.3f} ms  {100 * ms / tot:5.1f}%  x{n:4d}  {k[:150]}")
