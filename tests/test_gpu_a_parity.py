"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
inputs.  Bit-exact everywhere: digits as bytes, curve points as affine Montgomery bytes, polynomials as the
trimmed raw coefficients (EAGEN_RAW_TREE) and as the canonical monic form (EAGEN_CANONICAL)."""
import json
import os

import numpy as np
import pytest

import pyref

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CURVES = ["pallas", "vesta", "grumpkin"]


def trim(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    n = len(arr)
    while n and not arr[n - 1].any():
        n -= 1
    return arr[:n]


def same(a, b):
    a, b = np.asarray(a, dtype=np.uint64).reshape(-1, 4), np.asarray(b, dtype=np.uint64).reshape(-1, 4)
    return a.shape == b.shape and (a == b).all()


def gen(cv, n, seed):
    rng = pyref.SplitMix64(seed)
    p0, dl = pyref.random_point(rng, cv), pyref.random_point(rng, cv)
    pts = [p0]
    for _ in range(n - 1):
        pts.append(cv.add(pts[-1], dl))
    return pts, [pyref.random_scalar(rng, cv) for _ in range(n)]


# ---- K9 / helpers -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cname", CURVES)
def test_batch_invert(gpu_ctx, oracle, cname):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    rng = pyref.SplitMix64(1)
    for n in (1, 2, 17, 1024, 1025, 20000, 300000):
        vals = [rng.next_bits(4) % cv.p for _ in range(min(n, 3000))]
        vals[0] = 0 if n > 2 else vals[0]
        vals = (vals * (n // len(vals) + 1))[:n]
        if n > 10:
            vals[7] = 0
            vals[n - 1] = 1
        arr = oracle.pack_felts(vals[:3000], cv.p)
        arr = np.tile(arr, (n // len(arr) + 1, 1))[:n].copy()
        if n > 10:
            arr[7] = 0
            arr[n - 1] = oracle.pack_felts([1], cv.p)[0]
        got = ctx.batch_invert(arr)
        idx = sorted(set([0, 1, 7, n // 2, n - 2, n - 1]) & set(range(n)))
        for i in idx:
            want = oracle.field_op(cv.id if cname != "vesta" else 1, 3, arr[i]) if cname != "grumpkin" else oracle.field_op(2, 3, arr[i])
            assert (got[i] == want).all(), (n, i)
        # x * x^-1 == 1 everywhere (checked through the GPU product itself for speed): spot check 200 random entries
        one = oracle.pack_felts([1], cv.p)[0]
        fid = {"pallas": 0, "vesta": 1, "grumpkin": 2}[cname]
        for i in np.random.default_rng(n).integers(0, n, size=min(n, 200)):
            if arr[i].any():
                assert (oracle.field_op(fid, 2, arr[i], got[i]) == one).all()
            else:
                assert not got[i].any()


@pytest.mark.parametrize("cname", CURVES)
def test_ntt_matches_best_fft(gpu_ctx, oracle, cname):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    fid = {"pallas": 0, "vesta": 1, "grumpkin": 2}[cname]
    rng = pyref.SplitMix64(2)
    for log_n in (1, 2, 3, 5, 9, 10, 11, 12, 13, 16):
        n = 1 << log_n
        base = oracle.pack_felts([rng.next_bits(4) % cv.p for _ in range(min(n, 512))], cv.p)
        a = np.tile(base, (n // len(base) + 1, 1))[:n].copy()
        a[:, 0] ^= np.arange(n, dtype=np.uint64)  # distinct, still < p in the top limb
        f = ctx.ntt(a)
        assert same(f, oracle.fft(fid, a)), log_n
        assert same(ctx.ntt(f, inverse=True), oracle.fft(fid, f, inverse=True)), log_n


@pytest.mark.parametrize("cname", CURVES)
def test_poly_mul(gpu_ctx, oracle, cname):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    fid = {"pallas": 0, "vesta": 1, "grumpkin": 2}[cname]
    rng = pyref.SplitMix64(3)
    for la, lb in ((1, 1), (2, 1), (3, 5), (31, 32), (100, 423), (1500, 2500), (0, 4), (0, 0), (5, 0)):
        a = oracle.pack_felts([rng.next_bits(4) % cv.p for _ in range(la)], cv.p)
        b = oracle.pack_felts([rng.next_bits(4) % cv.p for _ in range(lb)], cv.p)
        got = ctx.poly_mul(a, b)
        want = oracle.poly_mul(fid, a, b, 0)
        assert same(got, want), (la, lb)


# ---- K1 ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("base", [2, 3, 5, 16, 17, 255])
def test_negbase_digits(gpu_ctx, oracle, eagen, cname, base):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    rng = pyref.SplitMix64(base)
    sq = pyref.isqrt(cv.q) + 2
    d = pyref.num_digits(cv, base)
    maxrep = sum((base - 1) * base ** i for i in range(0, d, 2))
    vals = [0, 1, 2, base - 1, base, base + 1, min(sq - 1, maxrep), min(sq - 2, maxrep), 2 ** 64, 2 ** 126 + 12345]
    vals += [min(rng.next_bits(2) % sq, maxrep) for _ in range(3000)]
    got = ctx.negbase_decompose(oracle.pack_felts(vals, cv.q), base)
    assert got.shape == (len(vals), d)
    for v, row in zip(vals, got):
        ref = pyref.negbase_decompose(v, base)
        assert row.tolist() == ([0] * (d - len(ref)) + ref[::-1])
    with pytest.raises(eagen.EagenError) as e:
        ctx.negbase_decompose(oracle.pack_felts([1, sq], cv.q), base)
    assert e.value.status == eagen.E_RANGE
    if maxrep < sq - 1:  # the reference would silently truncate here (src/argument_witness_calc.rs:99)
        with pytest.raises(eagen.EagenError) as e:
            ctx.negbase_decompose(oracle.pack_felts([sq - 1], cv.q), base)
        assert e.value.status == eagen.E_DIGITS


# ---- K2 ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cname", CURVES)
def test_multiples(gpu_ctx, oracle, cname):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    pts, _ = gen(cv, 40, 5)
    pts[3] = None
    rng = pyref.SplitMix64(8)
    zs = [rng.next_bits(4) % cv.p for _ in pts]
    for base in (2, 5, 17):
        got = ctx.precompute_multiplicities(oracle.pack_points(pts, cv.p, zs), base)
        for j, P in enumerate(pts):
            assert oracle.unpack_affine(got[j], cv.p) == [cv.mul(k, P) for k in range(1, base)]


# ---- divisor witnesses -------------------------------------------------------------------------------------------
def _pts(case):
    return [None if P is None else (int(P[0], 16), int(P[1], 16)) for P in case["points"]]


def _felts(lst):
    return [int(x, 16) for x in lst]


with open(os.path.join(G, "witness_small.json")) as _f:
    SMALL = json.load(_f)


@pytest.mark.parametrize("case", SMALL, ids=lambda c: c["curve"] + "-" + c["name"])
def test_golden_small_cases(gpu_ctx, oracle, eagen, case):
    cv, ctx = pyref.Curve(case["curve"]), gpu_ctx(case["curve"])
    pts = _pts(case)
    zs = [(11 * i + 5) % cv.p for i in range(len(pts))]
    P = oracle.pack_points(pts, cv.p, zs)
    if case["kind"] == "lhs":
        S = oracle.pack_felts(_felts(case["scalars"]), cv.q)
        raw = ctx.compute_lhs_witness(S, P, case["base"], eagen.RAW_TREE | eagen.KEEP_DIGITS)
        can = ctx.compute_lhs_witness(S, P, case["base"], eagen.CANONICAL)
        assert raw.digits.tolist() == case["digits"]
        want = [None if c is None else (int(c[0], 16), int(c[1], 16)) for c in case["carries"]]
        assert oracle.unpack_affine(raw.carries, cv.p) == want
        assert oracle.unpack_affine(raw.carry, cv.p)[0] == want[-1]
        assert raw.num_functions == len(case["raw"])
        for k in range(raw.num_functions):
            fr, fc = raw.function(k), can.function(k)
            assert oracle.unpack_felts(fr.a, cv.p) == oracle.unpack_felts(trim(oracle.pack_felts(_felts(case["raw"][k][0]), cv.p)), cv.p)
            assert oracle.unpack_felts(fr.b, cv.p) == oracle.unpack_felts(trim(oracle.pack_felts(_felts(case["raw"][k][1]), cv.p)), cv.p)
            assert oracle.unpack_felts(fc.a, cv.p) == _felts(case["canonical"][k][0])
            assert oracle.unpack_felts(fc.b, cv.p) == _felts(case["canonical"][k][1])
    else:
        fr, out = ctx.compute_divisor_witness_partial(P, eagen.RAW_TREE)
        fc, _ = ctx.compute_divisor_witness_partial(P, eagen.CANONICAL)
        want_out = None if case["output"] is None else (int(case["output"][0], 16), int(case["output"][1], 16))
        assert oracle.unpack_affine(out, cv.p)[0] == want_out
        assert same(fr.a, trim(oracle.pack_felts(_felts(case["raw"][0]), cv.p)))
        assert same(fr.b, trim(oracle.pack_felts(_felts(case["raw"][1]), cv.p)))
        assert oracle.unpack_felts(fc.a, cv.p) == _felts(case["canonical"][0])
        assert oracle.unpack_felts(fc.b, cv.p) == _felts(case["canonical"][1])


@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 33, 64, 65, 200, 1031, 2050, 5000])
def test_divisor_witness_vs_oracle(gpu_ctx, oracle, eagen, cname, n):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    pts, _ = gen(cv, n, 1000 + n)
    P = oracle.pack_points(pts, cv.p)
    ro = oracle.divisor_witness(cv.id, P, partial=True)
    fr, out = ctx.compute_divisor_witness_partial(P, eagen.RAW_TREE)
    assert (out == ro.output).all()
    assert same(fr.a, trim(ro.a[0])) and same(fr.b, trim(ro.b[0]))
    fc, _ = ctx.compute_divisor_witness_partial(P, eagen.CANONICAL)
    assert same(fc.a, ro.ca[0]) and same(fc.b, ro.cb[0])
    with pytest.raises(eagen.EagenError) as e:
        ctx.compute_divisor_witness(P)
    assert e.value.status == eagen.E_SUM_NONZERO


def test_divisor_witness_reference_test_shape_10000(gpu_ctx, oracle, eagen):
    """randpoints_witness_test at the reference's own size (src/regular_functions_utils.rs:653-659): 10 000 copies of one Grumpkin
    point plus minus their sum, and the same shape with 10 000 distinct points; raw and canonical form against the oracle"""
    cv, ctx = pyref.Curve("grumpkin"), gpu_ctx("grumpkin")
    pts, _ = gen(cv, 10000, 653)
    a = pts[0]
    acc = None
    for q in pts:
        acc = cv.add(acc, q)
    for v in ([a] * 10000 + [cv.neg(cv.mul(10000, a))], pts + [cv.neg(acc)]):
        P = oracle.pack_points(v, cv.p)
        ro = oracle.divisor_witness(cv.id, P)
        fr = ctx.compute_divisor_witness(P, eagen.RAW_TREE)
        assert same(fr.a, trim(ro.a[0])) and same(fr.b, trim(ro.b[0]))
        fc = ctx.compute_divisor_witness(P, eagen.CANONICAL)
        assert same(fc.a, ro.ca[0]) and same(fc.b, ro.cb[0])
        assert not ctx.eval_function(fc, P[:: 97]).any()


def test_divisor_witness_degenerate_geometry(gpu_ctx, oracle, eagen):
    """repeated points (the reference's tests use ONE point 10 000 times), P/-P pairs, identities"""
    cv, ctx = pyref.Curve("grumpkin"), gpu_ctx("grumpkin")
    pts, _ = gen(cv, 3, 77)
    a = pts[0]
    vectors = [
        [a] * 1000 + [cv.neg(cv.mul(1000, a))],                      # randpoints_witness_test shape
        [None, None, None, a, a, cv.neg(a), None, cv.neg(a), a, cv.neg(a)],  # witness_with_zeros_test
        [a, cv.neg(a)] * 37,
        [None] * 9,
        [None] * 5 + [a, cv.neg(a)] + [None] * 6,
        [a, a, a, a, cv.neg(cv.mul(4, a))],
        [pts[1], None, pts[2], None, None, cv.neg(cv.add(pts[1], pts[2]))],
    ]
    for v in vectors:
        P = oracle.pack_points(v, cv.p)
        ro = oracle.divisor_witness(cv.id, P)
        fr = ctx.compute_divisor_witness(P, eagen.RAW_TREE)
        assert same(fr.a, trim(ro.a[0])) and same(fr.b, trim(ro.b[0])), v[:4]
        fc = ctx.compute_divisor_witness(P, eagen.CANONICAL)
        assert same(fc.a, ro.ca[0]) and same(fc.b, ro.cb[0])
        vals = ctx.eval_function(fc, P)
        assert not vals.any()


# ---- the whole path ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("n,base", [(1, 5), (2, 5), (37, 5), (300, 5), (1024, 5), (130, 2), (130, 3), (130, 17), (40, 255)])
def test_lhs_witness_vs_oracle(gpu_ctx, oracle, eagen, cname, n, base):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    pts, sc = gen(cv, n, 31 * n + base)
    if base == 2:  # keep scalars inside what d digits can hold (reference truncates above it)
        d = pyref.num_digits(cv, 2)
        lim = sum(2 ** i for i in range(0, d, 2))
        sc = [min(s, lim) for s in sc]
    S, P = oracle.pack_felts(sc, cv.q), oracle.pack_points(pts, cv.p)
    ro = oracle.lhs_witness(cv.id, S, P, base)
    raw = ctx.compute_lhs_witness(S, P, base, eagen.RAW_TREE | eagen.KEEP_DIGITS)
    assert (raw.digits == ro.digits).all()
    assert (raw.carries == ro.carries).all() and (raw.carry == ro.carry).all()
    assert (raw.carry == oracle.msm_naive(cv.id, S, P)).all()
    can = ctx.compute_lhs_witness(S, P, base, eagen.CANONICAL)
    assert raw.num_functions == ro.d == len(ro.a)
    for k in range(ro.d):
        fr, fc = raw.function(k), can.function(k)
        assert same(fr.a, trim(ro.a[k])) and same(fr.b, trim(ro.b[k])), k
        assert same(fc.a, ro.ca[k]) and same(fc.b, ro.cb[k]), k


def test_lhs_repeated_point_reference_test_shape(gpu_ctx, oracle, eagen):
    """lhs_test: 10 000 copies of one Grumpkin point and one scalar (reference: src/argument_witness_calc.rs:138-148)"""
    cv, ctx = pyref.Curve("grumpkin"), gpu_ctx("grumpkin")
    pts, sc = gen(cv, 1, 4242)
    n = 10000
    S, P = oracle.pack_felts(sc * n, cv.q), oracle.pack_points(pts * n, cv.p)
    res = ctx.compute_lhs_witness(S, P, 5, eagen.CANONICAL | eagen.KEEP_DIGITS)
    assert oracle.unpack_affine(res.carry, cv.p)[0] == cv.mul(sc[0] * n % cv.q, pts[0])
    ro = oracle.lhs_witness(cv.id, S, P, 5, with_functions=False)
    assert (res.digits == ro.digits).all() and (res.carries == ro.carries).all()
    # compare ALL 56 functions with the oracle's tree on the same tmp list
    carries = oracle.unpack_affine(ro.carries, cv.p)
    mult = [cv.mul(k, pts[0]) for k in range(1, 5)]
    for i in range(ro.d):
        prev = carries[i - 1] if i else None
        dg = int(ro.digits[0][i])
        tmp = ([cv.neg(prev)] * 5 if prev is not None else []) + ([mult[dg - 1]] * n if dg else []) + [cv.neg(carries[i])]
        rt = oracle.divisor_witness(cv.id, oracle.pack_points(tmp, cv.p))
        f = res.function(ro.d - 1 - i)
        assert same(f.a, rt.ca[0]) and same(f.b, rt.cb[0])


def test_lhs_errors(gpu_ctx, oracle, eagen):
    cv, ctx = pyref.Curve("pallas"), gpu_ctx("pallas")
    pts, sc = gen(cv, 3, 9)
    P = oracle.pack_points(pts, cv.p)
    with pytest.raises(eagen.EagenError) as e:
        ctx.compute_lhs_witness(oracle.pack_felts(sc[:2], cv.q), P, 5)
    assert e.value.status == eagen.E_LEN
    with pytest.raises(eagen.EagenError) as e:
        ctx.compute_lhs_witness(oracle.pack_felts([1, 2 ** 127 + 2, 3], cv.q), P, 5)
    assert e.value.status == eagen.E_RANGE
    # the context stays usable after an error
    r = ctx.compute_lhs_witness(oracle.pack_felts(sc, cv.q), P, 5)
    assert r.num_functions == 56


def test_lhs_witness_streamed_output(gpu_ctx, oracle, eagen):
    """eagen_lhs_witness_stream: same functions as the handle-based call, laid out in fixed slots of the caller's buffer"""
    ctx = gpu_ctx("pallas")
    n = 3000
    S, P = ctx.synth_inputs(7, n)
    ref = ctx.compute_lhs_witness(S, P, 5)
    a_s, b_s, total = ctx.stream_layout(n, 5)
    assert total == ref.d * (a_s + b_s) * 32
    buf = np.zeros(total // 8, dtype=np.uint64)
    r = ctx.compute_lhs_witness_stream(S.ctypes.data, P.ctypes.data, n, 5, buf.ctypes.data, total)
    assert (r.carry == ref.carry).all() and r.num_functions == ref.num_functions
    slots = buf.reshape(ref.d, a_s + b_s, 4)
    for k in range(ref.d):
        fa, fb = ref.poly(k, 0), ref.poly(k, 1)
        assert len(r.poly(k, 0)) == len(fa) and len(r.poly(k, 1)) == len(fb)
        assert (slots[k, :len(fa)] == fa).all() and (slots[k, a_s:a_s + len(fb)] == fb).all()
    with pytest.raises(eagen.EagenError) as e:
        ctx.compute_lhs_witness_stream(S.ctypes.data, P.ctypes.data, n, 5, buf.ctypes.data, total - 32)
    assert e.value.status == eagen.E_LEN


@pytest.mark.parametrize("cname", CURVES)
def test_synthetic_inputs_are_valid_for_every_curve(gpu_ctx, oracle, eagen, cname):
    """the device-side generator stays below isqrt(order)+2 on every curve (126 bits on Grumpkin) and yields curve points"""
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    S, P = ctx.synth_inputs(0xEA6E0000, 400)
    sq = pyref.isqrt(cv.q) + 2
    assert all(s < sq for s in oracle.unpack_felts(S, cv.q))
    assert len({bytes(r) for r in P}) == 400
    res = ctx.compute_lhs_witness(S, P, 5, eagen.CANONICAL)
    assert (res.carry == oracle.msm_naive(cv.id, S, P)).all()


def test_empty_and_identity_inputs(gpu_ctx, oracle, eagen):
    """n = 0 (every tmp list is [O]: all functions are the constant 1), all-zero scalars, all-identity points"""
    cv, ctx = pyref.Curve("pallas"), gpu_ctx("pallas")
    one = oracle.pack_felts([1], cv.p)[0]
    r = ctx.compute_lhs_witness(np.zeros((0, 4), np.uint64), np.zeros((0, 12), np.uint64), 5, eagen.KEEP_DIGITS)
    assert r.num_functions == 56 and not r.carry.any()
    for k in range(56):
        f = r.function(k)
        assert len(f.a) == 1 and len(f.b) == 0 and (f.a[0] == one).all()
    pts, sc = gen(cv, 6, 3)
    P = oracle.pack_points(pts, cv.p)
    for S, Pp in ((oracle.pack_felts([0] * 6, cv.q), P), (oracle.pack_felts(sc, cv.q), np.zeros((6, 12), np.uint64))):
        ro = oracle.lhs_witness(cv.id, S, Pp, 5)
        rg = ctx.compute_lhs_witness(S, Pp, 5, eagen.CANONICAL)
        assert (rg.carries == ro.carries).all()
        for k in range(ro.d):
            f = rg.function(k)
            assert same(f.a, ro.ca[k]) and same(f.b, ro.cb[k])
    # compute_divisor_witness of the empty list is the constant 1 (reference: src/regular_functions_utils.rs:455)
    f = ctx.compute_divisor_witness(np.zeros((0, 12), np.uint64))
    assert len(f.a) == 1 and (f.a[0] == one).all() and len(f.b) == 0


@pytest.mark.parametrize("cname", CURVES)
def test_best_multiexp_full_width_scalars(gpu_ctx, oracle, eagen, cname):
    """the windowed bucket MSM (reference tests' cross-check `best_multiexp`) against double-and-add on the oracle side"""
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    rng = pyref.SplitMix64(77)
    for n in (1, 2, 100, 3000):
        pts, _ = gen(cv, n, 500 + n)
        sc = [rng.next_bits(4) % cv.q for _ in range(n)]
        if n >= 100:
            sc[3], sc[4], pts[5] = 0, cv.q - 1, None
        zs = [rng.next_bits(4) % cv.p or 1 for _ in range(n)]
        S, P = oracle.pack_felts(sc, cv.q), oracle.pack_points(pts, cv.p, zs)
        assert (ctx.best_multiexp(S, P) == oracle.msm_naive(cv.id, S, P)).all(), n
    assert not ctx.best_multiexp(np.zeros((0, 4), np.uint64), np.zeros((0, 12), np.uint64)).any()
    # and the identity the reference's lhs_test asserts: witness carry == best_multiexp (src/argument_witness_calc.rs:144-147)
    S, P = ctx.synth_inputs(31337, 2000)
    assert (ctx.compute_lhs_witness(S, P, 5, eagen.NO_FUNCTIONS).carry == ctx.best_multiexp(S, P)).all()
