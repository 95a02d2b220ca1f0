"""per-scope device ms of one witness step, median over a few steps (low-noise view of one kernel group): python tools/scope_ms.py [scope ...]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
eg = load_package()
ctx = eg.Context("pallas", 0)
n = 1 << 20
dev = torch.device("cuda", 0)
d_s = torch.empty(n * 32, dtype=torch.uint8, device=dev); d_p = torch.empty(n * 96, dtype=torch.uint8, device=dev)
ctx.dev_synth_inputs(0xEA6E0002, n, d_s.data_ptr(), d_p.data_ptr())
ctx.set_profiling(1)
rows = {}
tot = []
for it in range(6):
    ctx.profile_reset()
    r = ctx.compute_lhs_witness_ptr(d_s.data_ptr(), d_p.data_ptr(), n, 5, eg.CANONICAL, device=True)
    tot.append(r.device_ms); r.free()
    if it >= 2:
        for e in ctx.profile():
            rows.setdefault(e["kernel"], []).append(e["ms"])
want = sys.argv[1:]
print("step", "%.1f" % statistics.median(tot[2:]), " ".join("%s=%.2f" % (k, statistics.median(v)) for k, v in rows.items() if not want or k in want))
