"""small end-to-end run for compute-sanitizer (one tool per gpurun call): all three curves, a few sizes incl. multi-pass NTT"""
import sys
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_package
eg = load_package()
for curve in ("pallas", "vesta", "grumpkin"):
    ctx = eg.Context(curve, 0)
    for n in (1, 7, 300, 2600):
        S, P = ctx.synth_inputs(11 + n, n)
        r = ctx.compute_lhs_witness(S, P, 5, eg.CANONICAL | eg.KEEP_DIGITS)
        assert r.num_functions == r.d
        r.free()
    S, P = ctx.synth_inputs(5, 40)
    ctx.compute_lhs_witness(S, P, 17, eg.RAW_TREE).free()
    f, out = ctx.compute_divisor_witness_partial(P[:33])
    ctx.eval_function(f, P[:33])
    ctx.close()
print("sanitize run finished")
