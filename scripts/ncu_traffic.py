#!/usr/bin/env python
"""profiles/ncu_traffic.json: DRAM bytes per launch of the dominant kernels, read from the committed
`ncu --set full` raw pages (profiles/*_raw.csv), keyed by kernel name and tagged with the hash of the
kernel's source file at the time -- bench.py reports roofline.traffic from it and prints null when
the source changed after the capture.

    python scripts/ncu_traffic.py ell_tma_pipe_kernel<1>=profiles/r1_ell_tma_pipe_c2_raw.csv:gpu-spmv_b200/csrc/ell_kernels.cu ...
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(path):
    rows = list(csv.reader(open(os.path.join(ROOT, path))))
    hdr, units = rows[0], rows[1]
    out = []
    for data in rows[2:]:
        total = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            total += float(data[i].replace(",", "")) * UNIT[units[i]]
        out.append((data[hdr.index("Kernel Name")], total, float(data[hdr.index("gpu__time_duration.sum")].replace(",", ""))))
    return out


table_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
table = json.load(open(table_path)) if os.path.exists(table_path) else {}
for arg in sys.argv[1:]:
    kernel, rest = arg.split("=", 1)
    capture, source = rest.split(":", 1)
    launches = [l for l in dram_bytes(capture) if kernel.split("<")[0] in l[0]]
    assert launches, f"{kernel}: no launch in {capture}"
    digest = hashlib.sha256(open(os.path.join(ROOT, source), "rb").read()).hexdigest()[:16]
    table[kernel] = {"dram_bytes_per_launch": sum(l[1] for l in launches) / len(launches), "launches_in_capture": len(launches),
                     "gpu_time_us_under_ncu": sum(l[2] for l in launches) / len(launches),
                     "capture": capture, "source": source, "source_sha256_16": digest}
json.dump(table, open(table_path, "w"), indent=1, sort_keys=True)
print(json.dumps(table, indent=1, sort_keys=True))
