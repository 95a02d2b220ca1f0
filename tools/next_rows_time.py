"""kernel-side timing of the section 8f rows (profiling scopes of the engine): prepare_scalar_witness, result_eval, naive witness, MSM"""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_package
eg = load_package()
ctx = eg.Context("pallas", 0)
n = 1 << 20
S, P = ctx.synth_inputs(7, n)
d = eg.num_digits(eg.PALLAS, 5)
ctx.set_profiling(True)
for _ in range(2):
    ctx.profile_reset()
    t0 = time.perf_counter()
    out = ctx.prepare_scalar_witness(S, 5, d, 8, eg.PSW_INTENDED)
    wall = time.perf_counter() - t0
    pr = {e["kernel"]: e for e in ctx.profile()}
print("prepare_scalar_witness 2^20 scalars, base 5, logtable 8: kernel %.3f ms (%.0f GB/s of %d MB written), negbase %.3f ms, whole call %.0f ms"
      % (pr["scalar_witness"]["ms"], out.nbytes / pr["scalar_witness"]["ms"] / 1e6, out.nbytes >> 20, pr["negbase"]["ms"], wall * 1e3))
res = ctx.compute_lhs_witness(S, P, 5, eg.CANONICAL)
for _ in range(2):
    ctx.profile_reset()
    t0 = time.perf_counter()
    v = res.ev(P[:4])
    wall = time.perf_counter() - t0
    pr = {e["kernel"]: e for e in ctx.profile()}
print("result_eval: 56 functions (%.2f GB of coefficients) at 4 points: kernels %.2f ms, whole call %.1f ms"
      % (res.total_bytes() / 1e9, pr["result_eval"]["ms"], wall * 1e3))
res.free()
t0 = time.perf_counter(); _, ms = ctx.best_multiexp(S, P, with_time=True); print("best_multiexp 2^20: device %.2f ms" % ms)
