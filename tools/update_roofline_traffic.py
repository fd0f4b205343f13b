#!/usr/bin/env python
"""profiles/roofline_traffic.json <- one `ncu --set full` capture: dram__bytes_read.sum + dram__bytes_write.sum of the
captured launch, stored with a hash of the kernel's source files so bench.py quotes it only while those sources are
unchanged.   usage: update_roofline_traffic.py <workload> <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <capture name> <src files...>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

workload, raw, capture, files = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
rows = list(csv.reader(open(raw)))
d = dict(zip(rows[0], zip(rows[-1], rows[1])))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = sum(float(d[k][0]) * scale[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
try:
  cur = json.load(open(path))
except Exception:
  cur = {}
cur = {k: v for k, v in cur.items() if isinstance(v, dict)}
cur[workload] = {"dram_bytes_per_launch": int(tot), "capture": capture, "kernel": d["Kernel Name"][0][:60],
                 "kernel_sources": files, "kernel_source_sha16": bench.kernel_source_hash(files),
                 "tensor_pipe_pct_of_elapsed": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]),
                 "gpu_time_ms_under_ncu": d["gpu__time_duration.sum"][0] + " " + d["gpu__time_duration.sum"][1]}
json.dump(cur, open(path, "w"), indent=1)
print(json.dumps(cur[workload]))
