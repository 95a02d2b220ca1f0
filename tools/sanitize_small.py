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
    # rows next to the path: K1 at other bases, prepare_scalar_witness, naive arrangement, padded rows,
    # evaluation of a result at points, windowed MSM
    S, P = ctx.synth_inputs(9, 301)
    for base in (2, 3, 17, 255):
        ctx.negbase_decompose(S, base)
    ctx.negbase_decompose(S[:300], 5)
    d = eg.num_digits(eg.CURVE_IDS[curve], 5)
    ctx.prepare_scalar_witness(S, 5, d, 8, eg.PSW_INTENDED)
    ctx.prepare_scalar_witness(S[:7], 5, d, 1, eg.PSW_FAITHFUL)
    _, out = ctx.compute_divisor_witness_partial(P[:64])
    import numpy as np
    closing = np.zeros((1, 12), dtype=np.uint64)
    closing[0, :8] = out
    closing[0, 8:12] = eg.half_pow(eg.CURVE_IDS[curve], 0)   # z = Montgomery 1; partial output is -(sum), so this closes the list
    ctx.compute_divisor_witness_naive(np.concatenate([P[:64], closing]))
    r = ctx.compute_lhs_witness(S[:151], P[:151], 5, eg.CANONICAL)
    r.ev(P[:3])
    r.padded(151, 5)
    r.free()
    ctx.best_multiexp(S[:200], P[:200])
    ctx.close()
print("sanitize run finished")
