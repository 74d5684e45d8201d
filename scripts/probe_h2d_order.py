#!/usr/bin/env python
"""In which order does one cudaMemcpyAsync land in device memory, and does polling its destination disturb it?
(spmv_b200_probe_h2d_order)  -> profiles/r2_host_gated.txt"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from _load_pkg import load_pkg
sp = load_pkg()
torch.zeros(1, device="cuda")
n = 1 << 24
xh = torch.rand(n).pin_memory()
names = {-1: "no polling", 0: "ld.relaxed.sys", 1: "ld.relaxed.gpu", 2: "ld.volatile"}
for samples, mode, sleep in ((16, -1, 0), (16, -1, 0), (16, 0, 0), (16, 1, 0), (16, 2, 0), (16, 0, 1000), (16, 1, 1000), (1024, 1, 0), (1024, 1, 1000), (1024, 0, 1000), (16, -1, 0)):
    out = (C.c_longlong * (samples + 1))()
    rc = sp.lib.spmv_b200_probe_h2d_order(xh.data_ptr(), n, samples, out, mode, sleep)
    us = [v / 1e3 for v in out]
    step = max(1, samples // 16)
    print(f"{samples:5d} pollers, {names[mode]:15s} sleep {sleep:5d} ns: copy {us[samples]:7.0f} us by events; arrivals [us] " +
          " ".join(f"{v:.0f}" for v in us[:samples:step]), flush=True)
