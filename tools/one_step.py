#!/usr/bin/env python3
"""one warm-up + one witness step at 2^log_n (for ncu launch lists): python tools/one_step.py [log_n] [curve]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
eg = load_package()
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = eg.Context(sys.argv[2] if len(sys.argv) > 2 else "pallas", 0)
n = 1 << log_n
dev = torch.device("cuda", 0)
d_s = torch.empty(n * 32, dtype=torch.uint8, device=dev)
d_p = torch.empty(n * 96, dtype=torch.uint8, device=dev)
ctx.dev_synth_inputs(0xEA6E0002, n, d_s.data_ptr(), d_p.data_ptr())
for _ in range(2):
    r = ctx.compute_lhs_witness_ptr(d_s.data_ptr(), d_p.data_ptr(), n, 5, eg.CANONICAL, device=True)
    print("device ms", r.device_ms)
    r.free()
