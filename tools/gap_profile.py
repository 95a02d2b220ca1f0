"""where the main stream idles: per-scope gaps of one step (profiling mode 2), largest first: python tools/gap_profile.py"""
import os, sys
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
for _ in range(2):
    ctx.compute_lhs_witness_ptr(d_s.data_ptr(), d_p.data_ptr(), n, 5, eg.CANONICAL, device=True).free()
ctx.set_profiling(2)
ctx.profile_reset()
steps = 3
tot = 0
for _ in range(steps):
    r = ctx.compute_lhs_witness_ptr(d_s.data_ptr(), d_p.data_ptr(), n, 5, eg.CANONICAL, device=True); tot += r.device_ms; r.free()
prof = ctx.profile()
gaps = sorted(((e["ms"] / steps, e["kernel"]) for e in prof if e["kernel"].startswith("idle~")), reverse=True)
print("step %.1f ms; idle between main-stream scopes %.1f ms; largest:" % (tot / steps, sum(g[0] for g in gaps)), ", ".join("%s %.2f" % (k[12:], v) for v, k in gaps[:10]))
