import sys
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_package
eg = load_package()
ctx = eg.Context("pallas", 0)
for log_n in (16, 20, 22):
    n = 1 << log_n
    S, P = ctx.synth_inputs(5, n)
    ms = min(ctx.best_multiexp(S, P, with_time=True)[1] for _ in range(3))
    print("best_multiexp 2^%d Pallas (127-bit scalars, 32 byte windows): %.2f ms device = %.1f M points/s" % (log_n, ms, n / ms / 1e3))
