"""GPU tests at BASELINE.json sizes (pytest -m gpu): config 2 (2^16) with ALL functions against the oracle, config 3 (2^20)
against the committed SHA-256 of the oracle's full-size run (every digit, carry and coefficient) and through
size-independent properties (carry == independent MSM on the oracle side, functions vanish on their points, degrees,
the norm identity)."""
import json
import os

import numpy as np
import pytest

import pyref

pytestmark = pytest.mark.gpu


def trim(arr):
    n = len(arr)
    while n and not arr[n - 1].any():
        n -= 1
    return arr[:n]


def build_tmp(mult, digits, carries, i, base):
    """tmp_i of the reference (src/argument_witness_calc.rs:110-127) as (m,12) Jacobian rows with z = 1"""
    n = digits.shape[0]
    one = None
    sel = digits[:, i] != 0
    idx = np.nonzero(sel)[0]
    rows = mult[idx, digits[idx, i].astype(np.int64) - 1]  # (m, 8)
    return rows, idx


def affine_to_jac(aff, one):
    out = np.zeros((len(aff), 12), dtype=np.uint64)
    out[:, :8] = aff
    nz = aff.any(axis=1)
    out[nz, 8:12] = one
    return out


def neg_affine(row, p, oracle):
    pt = oracle.unpack_affine(row, p)[0]
    if pt is None:
        return np.zeros(8, dtype=np.uint64)
    return oracle.pack_points([(pt[0], (-pt[1]) % p)], p)[0][:8]


@pytest.mark.parametrize("cname,log_n", [("pallas", 16), ("vesta", 14)])
def test_config2_all_functions_vs_oracle(gpu_ctx, oracle, eagen, cname, log_n):
    """BASELINE config 2: digits, carries and ALL d canonical (a, b) bit-exact against the oracle's own
    full compute_lhs_witness on the same inputs (about 40 s of oracle at 2^16 on the box's host cores)."""
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    n, base = 1 << log_n, 5
    oracle.set_threads(os.cpu_count() or 1)
    S, P = ctx.synth_inputs(0xEA6E0001, n)
    res = ctx.compute_lhs_witness(S, P, base, eagen.CANONICAL | eagen.KEEP_DIGITS)
    ro = oracle.lhs_witness(cv.id, S, P, base)
    assert (res.digits == ro.digits).all()
    assert (res.carries == ro.carries).all() and (res.carry == ro.carry).all()
    assert res.num_functions == ro.d == len(ro.ca)
    for k in range(ro.d):
        f = res.function(k)
        assert f.a.shape == ro.ca[k].shape and (f.a == ro.ca[k]).all(), k
        assert f.b.shape == ro.cb[k].shape and (f.b == ro.cb[k]).all(), k
    res.free()


@pytest.mark.parametrize("cname", ["pallas", "vesta", "grumpkin"])
def test_synthetic_inputs_match_the_oracle_restatement(gpu_ctx, oracle, cname):
    """the golden hashes below were computed from oracle_synth_inputs: the device generator must describe the same scalars and the
    same points (it emits Jacobian triples with non-trivial z, the oracle z = 1; compare the affine points)"""
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    for seed, n in ((0xEA6E0002, 1031), (7, 64)):
        S, P = ctx.synth_inputs(seed, n)
        So, Po = oracle.synth_inputs(cv.id, seed, n)
        assert (S == So).all()
        aff = ctx.precompute_multiplicities(P, 2)[:, 0, :]          # 1 * P_j, affine
        assert (aff == Po[:, :8]).all()


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("fname", sorted(f for f in os.listdir(GOLDEN) if f.startswith("lhs_") and f.endswith("_hashes.json")))
def test_full_size_hashes(gpu_ctx, eagen, fname):
    """BASELINE config 3 (and any other committed full-size run): every digit, every carry and every coefficient of all d
    canonical functions against the SHA-256 record of the oracle's own full-size run (tools/golden_full_size.py; 2^20 Pallas
    points took the oracle 30 min on 8 cores).  This pins every kernel change at the metric's size."""
    from hashes import witness_hashes
    with open(os.path.join(GOLDEN, fname)) as f:
        gold = json.load(f)
    ctx = gpu_ctx(gold["curve"])
    n = 1 << gold["log_n"]
    S, P = ctx.synth_inputs(int(gold["seed"], 0), n)
    res = ctx.compute_lhs_witness(S, P, gold["base"], eagen.CANONICAL | eagen.KEEP_DIGITS)
    assert res.d == gold["d"] == res.num_functions
    fa, fb = [], []
    for k in range(res.d):
        f = res.function(k)
        fa.append(f.a)
        fb.append(f.b)
    rec = witness_hashes(res.digits, res.carries, fa, fb)
    res.free()
    assert rec["digits"] == gold["digits"]
    assert rec["carries"] == gold["carries"]
    for k, (g, h) in enumerate(zip(gold["functions"], rec["functions"])):
        assert (g["la"], g["lb"]) == (h["la"], h["lb"]), k
        assert g["sha256"] == h["sha256"], k


@pytest.mark.parametrize("cname,log_n", [("grumpkin", 13)])
def test_config2_sampled_trees_vs_oracle(gpu_ctx, oracle, eagen, cname, log_n):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    n, base = 1 << log_n, 5
    S, P = ctx.synth_inputs(0xEA6E0001, n)
    res = ctx.compute_lhs_witness(S, P, base, eagen.CANONICAL | eagen.KEEP_DIGITS)
    ro = oracle.lhs_witness(cv.id, S, P, base, with_functions=False)
    assert (res.digits == ro.digits).all()
    assert (res.carries == ro.carries).all() and (res.carry == ro.carry).all()
    mult = ctx.precompute_multiplicities(P, base)
    one = oracle.pack_felts([1], cv.p)[0]
    d = ro.d
    for i in (1, 17, d - 1):
        rows, _ = build_tmp(mult, ro.digits, ro.carries, i, base)
        parts = []
        if i and ro.carries[i - 1].any():
            parts.append(np.tile(neg_affine(ro.carries[i - 1], cv.p, oracle), (base, 1)))
        parts += [rows, neg_affine(ro.carries[i], cv.p, oracle).reshape(1, 8)]
        tmp = affine_to_jac(np.concatenate(parts), one)
        rt = oracle.divisor_witness(cv.id, tmp)
        f = res.function(d - 1 - i)
        assert f.a.shape == rt.ca[0].shape and (f.a == rt.ca[0]).all()
        assert f.b.shape == rt.cb[0].shape and (f.b == rt.cb[0]).all()


def test_config3_2pow20_properties(gpu_ctx, oracle, eagen):
    cv, ctx = pyref.Curve("pallas"), gpu_ctx("pallas")
    n, base = 1 << 20, 5
    S, P = ctx.synth_inputs(0xEA6E0002, n)
    res = ctx.compute_lhs_witness(S, P, base, eagen.CANONICAL | eagen.KEEP_DIGITS)
    d = res.d
    digits, carries = res.digits, res.carries
    # digits: exact against the oracle on a 4096-scalar sample, round trip on all of them via numpy in float-free form
    samp = np.random.default_rng(1).integers(0, n, size=4096)
    ro = oracle.lhs_witness(cv.id, S[samp], P[samp], base, with_functions=False)
    assert (digits[samp] == ro.digits).all()
    # final carry == independent MSM (oracle, double-and-add over the full 2^20 points, threaded)
    assert (res.carry == oracle.msm_naive(cv.id, S, P)).all()
    # per-position structure: degrees and vanishing on sampled points of tmp_i
    mult = ctx.precompute_multiplicities(P, base)
    one = oracle.pack_felts([1], cv.p)[0]
    for i in (2, 29, d - 1):  # position 0 is empty for scalars < 2^127 (5^55 > 2^127): f = 1 there
        rows, idx = build_tmp(mult, digits, carries, i, base)
        extra = (base if (i and carries[i - 1].any()) else 0) + (1 if carries[i].any() else 0)
        npts = len(rows) + extra
        f = res.function(d - 1 - i)
        assert len(f.a) == npts // 2 + 1 and len(f.b) == (npts - 3) // 2 + 1
        pick = np.random.default_rng(i).integers(0, len(rows), size=512)
        pts = [rows[pick], neg_affine(carries[i], cv.p, oracle).reshape(1, 8)]
        if i:
            pts.append(neg_affine(carries[i - 1], cv.p, oracle).reshape(1, 8))
        vals = ctx.eval_function(f, affine_to_jac(np.concatenate(pts), one))
        assert not vals.any()
        # the function must NOT vanish on an unrelated point
        other = ctx.eval_function(f, ctx.synth_inputs(0x0BADC0DE, 4)[1])
        assert other.any(axis=1).all()
    f0 = res.function(d - 1)  # iteration 0: tmp = [-carry] = [O]  ->  the constant 1
    assert len(f0.a) == 1 and len(f0.b) == 0 and (f0.a[0] == one).all()


@pytest.mark.parametrize("cname,log_n", [("pallas", 20), ("grumpkin", 18), ("vesta", 17)])
def test_config3_norm_identity(gpu_ctx, oracle, eagen, cname, log_n):
    """Full-size check of EVERY coefficient: for the monic divisor witness f = a + y b of the n points of tmp_i,
    f(Q) f(-Q) = a(x)^2 - (x^3 + b) b(x)^2 = (-1)^n prod_i (x - x(P_i)) at any curve point Q = (x, y).  Both sides are
    evaluated at random points: the left one on the device from the resident coefficients (eagen_result_eval), the right one with
    Python integers over all points of the list (~0.84 M at 2^20; Grumpkin exercises the generic, non-sparse modulus at scale) (Schwartz-Zippel: a wrong coefficient anywhere fails with probability
    ~ 1 - 2^-230)."""
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    p = cv.p
    n, base = 1 << log_n, 5
    S, P = ctx.synth_inputs(0xEA6E0002, n)
    res = ctx.compute_lhs_witness(S, P, base, eagen.CANONICAL | eagen.KEEP_DIGITS)
    d, digits, carries = res.d, res.digits, res.carries
    mult = ctx.precompute_multiplicities(P, base)
    rng = pyref.SplitMix64(2020 + log_n)
    Q = pyref.random_point(rng, cv)
    QJ = oracle.pack_points([Q, cv.neg(Q)], p)
    vals = res.ev(QJ)                                   # (d, 2, 4): f_k(Q), f_k(-Q) for every digit position
    rinv = pow(pyref.R, -1, p)
    xq_m = Q[0] * pyref.R % p                           # x(Q) in Montgomery form: (xq_m - xm_i) = R (x(Q) - x_i)
    for i in (2, 29, d - 1):
        rows, _ = build_tmp(mult, digits, carries, i, base)
        xs = rows[:, :4].astype(object)
        xm = xs[:, 0] + (xs[:, 1] << 64) + (xs[:, 2] << 128) + (xs[:, 3] << 192)
        extra = []
        if i and carries[i - 1].any():
            extra += [oracle.unpack_affine(carries[i - 1], p)[0][0]] * base      # b copies of -carry_{i-1}: same x as carry_{i-1}
        if carries[i].any():
            extra.append(oracle.unpack_affine(carries[i], p)[0][0])
        npts = len(xm) + len(extra)
        acc = 1
        for v in xm:
            acc = acc * (xq_m - v) % p
        acc = acc * pow(rinv, len(xm), p) % p
        for x in extra:
            acc = acc * (Q[0] - x) % p
        want = acc if npts % 2 == 0 else (-acc) % p
        k = d - 1 - i
        fq, fmq = oracle.unpack_felts(vals[k, 0], p)[0], oracle.unpack_felts(vals[k, 1], p)[0]
        assert fq * fmq % p == want, i
    res.free()
