"""GPU parity tests (pytest -m gpu) for the SURVEY.md section 8f rows next to the hot path, called through the C ABI and
compared bit for bit with the CPU oracle: prepare_scalar_witness (reference: src/negbase_utils.rs:79-124),
compute_divisor_witness_naive (src/regular_functions_utils.rs:483-551) and the division-free K1 at every base."""
import numpy as np
import pytest

import pyref

pytestmark = pytest.mark.gpu
CURVES = ["pallas", "vesta", "grumpkin"]


def gen_points(cv, n, seed):
    rng = pyref.SplitMix64(seed)
    p0, dl = pyref.random_point(rng, cv), pyref.random_point(rng, cv)
    pts = [p0]
    for _ in range(n - 1):
        pts.append(cv.add(pts[-1], dl))
    return pts


def close_sum(pts, cv):
    s = None
    for P in pts:
        s = cv.add(s, P)
    return pts + [cv.neg(s)]


def trim_rows(arr):
    """strip trailing zero coefficients (the product returns raw functions trimmed)"""
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    n = len(arr)
    while n and not arr[n - 1].any():
        n -= 1
    return arr[:n]


def psw_rows(arr):
    """structured (base, num_limbs+1) array of one scalar -> the tuple form oracle_lib.prepare_scalar_witness returns"""
    rows = []
    for i in range(arr.shape[0]):
        row = []
        for j in range(arr.shape[1]):
            e = arr[i, j]
            assert int(e["zero"]) == 0
            v = int(e["lo"]) | (int(e["hi"]) << 64)
            sv = v - (1 << 128) if v >> 127 else v
            kind = int(e["kind"])
            row.append(("scalar", v) if kind == 0 else (("bucket", sv) if kind == 1 else ("limb", sv, int(e["mask"]))))
        rows.append(row)
    return rows


# ---- K1 at every base -------------------------------------------------------------------------------------
@pytest.mark.parametrize("cname", CURVES)
def test_negbase_every_base(gpu_ctx, oracle, cname):
    """digits of edge and random scalars for bases 2..255, ragged n (byte-store path) and n % 4 == 0 (32-bit-store path)"""
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    sq = pyref.isqrt(cv.q) + 2
    rng = pyref.SplitMix64(11)
    for base in list(range(2, 34)) + [63, 64, 100, 127, 128, 200, 254, 255]:
        d = pyref.num_digits(cv, base)
        for n in (131, 260):
            xs = [0, 1, base - 1, base, sq - 1, sq - 2, (1 << 64) - 1, 1 << 64][: n]
            xs += [rng.next_bits(2) % sq for _ in range(n - len(xs))]
            xs = [x for x in xs if len(pyref.negbase_decompose(x, base)) <= d]
            got = ctx.negbase_decompose(oracle.pack_felts(xs, cv.q), base)
            assert got.shape == (len(xs), d)
            for i in list(range(10)) + [len(xs) // 2, len(xs) - 1]:
                ref = pyref.negbase_decompose(xs[i], base)
                assert list(got[i]) == [0] * (d - len(ref)) + ref[::-1], (base, n, i)


# ---- prepare_scalar_witness --------------------------------------------------------------------------------
@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("mode", [0, 1])
def test_prepare_scalar_witness(gpu_ctx, oracle, eagen, cname, mode):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    sq = pyref.isqrt(cv.q) + 2
    rng = pyref.SplitMix64(900 + mode)
    for base in (2, 3, 5, 16, 17, 255):
        d = pyref.num_digits(cv, base)
        for logtable in (1, 3, 8, 13, 24):
            num_limbs = (d + logtable - 1) // logtable
            xs = [0, 1, base, sq - 1] + [rng.next_bits(2) % sq for _ in range(5)]
            xs = [x for x in xs if len(pyref.negbase_decompose(x, base)) <= d]
            ok, refs = [], []
            for x in xs:
                try:
                    refs.append(oracle.prepare_scalar_witness(x, base, d, logtable, mode))
                    ok.append(x)
                except oracle.OracleError:
                    # faithful mode: the reference indexes out of bounds -> the product reports EAGEN_E_ARG for the batch
                    with pytest.raises(eagen.EagenError) as ei:
                        ctx.prepare_scalar_witness(oracle.pack_felts([x], cv.q), base, d, logtable, mode)
                    assert ei.value.status == eagen.E_ARG
            if not ok:
                continue
            got = ctx.prepare_scalar_witness(oracle.pack_felts(ok, cv.q), base, d, logtable, mode)
            assert got.shape == (len(ok), base, num_limbs + 1)
            for i, ref in enumerate(refs):
                assert psw_rows(got[i]) == ref, (base, logtable, ok[i])


def test_prepare_scalar_witness_batch_and_errors(gpu_ctx, oracle, eagen):
    """a batch that spans several blocks; the reference's assert on the digit count (:81); empty input"""
    cv, ctx = pyref.Curve("pallas"), gpu_ctx("pallas")
    sq = pyref.isqrt(cv.q) + 2
    rng = pyref.SplitMix64(4)
    xs = [rng.next_bits(2) % sq for _ in range(1000)]
    d = pyref.num_digits(cv, 5)
    got = ctx.prepare_scalar_witness(oracle.pack_felts(xs, cv.q), 5, d, 8, eagen.PSW_INTENDED)
    for i in (0, 1, 127, 128, 500, 999):
        assert psw_rows(got[i]) == oracle.prepare_scalar_witness(xs[i], 5, d, 8, 1)
    with pytest.raises(eagen.EagenError) as ei:   # sq - 1 needs more than 10 digits
        ctx.prepare_scalar_witness(oracle.pack_felts([sq - 1], cv.q), 5, 10, 8, eagen.PSW_INTENDED)
    assert ei.value.status == eagen.E_DIGITS
    with pytest.raises(eagen.EagenError) as ei:
        ctx.prepare_scalar_witness(oracle.pack_felts([sq], cv.q), 5, d, 8, eagen.PSW_INTENDED)
    assert ei.value.status == eagen.E_RANGE
    assert ctx.prepare_scalar_witness(np.zeros((0, 4), np.uint64), 5, d, 8).shape[0] == 0


# ---- compute_divisor_witness_naive ---------------------------------------------------------------------------
@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 8, 33, 100, 1000])
def test_divisor_witness_naive(gpu_ctx, oracle, cname, n):
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    rng = pyref.SplitMix64(n)
    pts = close_sum(gen_points(cv, n, 3 * n + 1), cv)
    zs = [rng.next_bits(4) % cv.p or 1 for _ in pts]
    P = oracle.pack_points(pts, cv.p, zs)
    pos, neg = ctx.compute_divisor_witness_naive(P)
    rpos, rneg = oracle.divisor_witness_naive(cv.id, P)
    assert pos.shape == rpos.shape and (pos == rpos).all()
    assert neg.shape == rneg.shape and (neg == rneg).all()


def test_divisor_witness_naive_degenerate(gpu_ctx, oracle, eagen):
    cv, ctx = pyref.Curve("pallas"), gpu_ctx("pallas")
    g = gen_points(cv, 6, 9)
    cases = [
        [g[0], None, cv.neg(g[0])],
        [None, g[0], g[1], None, None, cv.neg(cv.add(g[0], g[1]))],
        [g[0], g[0], g[0], cv.neg(cv.mul(3, g[0]))],
        [g[0], cv.neg(g[0]), g[1], cv.neg(g[1])],
        close_sum([g[2]] * 9, cv),
        close_sum([g[3]] * 2000, cv),
        [None, None],
        [],
    ]
    for pts in cases:
        P = oracle.pack_points(pts, cv.p) if pts else np.zeros((0, 12), np.uint64)
        pos, neg = ctx.compute_divisor_witness_naive(P)
        rpos, rneg = oracle.divisor_witness_naive(cv.id, P)
        assert pos.shape == rpos.shape and (pos == rpos).all(), len(pts)
        assert neg.shape == rneg.shape and (neg == rneg).all(), len(pts)
    with pytest.raises(eagen.EagenError) as ei:
        ctx.compute_divisor_witness_naive(oracle.pack_points([g[0], g[1]], cv.p))
    assert ei.value.status == eagen.E_SUM_NONZERO


# ---- circuit-facing layouts: padded rows and evaluation of every f_k at challenge points ----------------------
@pytest.mark.parametrize("cname", CURVES)
def test_result_eval_and_padded_rows(gpu_ctx, oracle, eagen, cname):
    """RegularFunction::ev of all d functions on the device (src/regular_functions_utils.rs:228-237) against the oracle's
    Horner evaluation of the copied coefficients, and the zero-padded a_size / b_size rows of src/config.rs:641-642"""
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    for n, flags in ((5001, eagen.CANONICAL), (300, eagen.RAW_TREE), (1, eagen.CANONICAL)):
        S, P = ctx.synth_inputs(4242 + n, n)
        res = ctx.compute_lhs_witness(S, P, 5, flags)
        rng = pyref.SplitMix64(n)
        q = gen_points(cv, 3, 77 + n) + [None]
        zs = [rng.next_bits(4) % cv.p or 1 for _ in q]
        Q = np.concatenate([oracle.pack_points(q, cv.p, zs), P[:2]])        # random points, the identity, two input points
        got = res.ev(Q)
        assert got.shape == (res.num_functions, len(Q), 4)
        fns = res.functions()
        for k in (0, 1, res.d // 2, res.d - 1):
            for j in range(len(Q)):
                if j == 3:
                    assert not got[k, j].any()       # identity -> 0 by convention
                    continue
                want = oracle.eval_function(cv.id, fns[k].a, fns[k].b, Q[j])
                assert (got[k, j] == np.asarray(want, dtype=np.uint64).reshape(4)).all(), (n, k, j)
        # padded rows: n odd -> every list has at most n + base + 1 points, so every function fits
        if n % 2 == 1 and flags == eagen.CANONICAL:
            a_size, b_size = eagen.circuit_sizes(n, 5)
            A, B = res.padded(n, 5)
            assert A.shape == (res.d, a_size, 4) and B.shape == (res.d, b_size, 4)
            for k in range(res.d):
                la, lb = len(fns[k].a), len(fns[k].b)
                assert (A[k, :la] == fns[k].a).all() and not A[k, la:].any()
                assert (B[k, :lb] == fns[k].b).all() and not B[k, lb:].any()
        if n > 1:   # rows too short for these functions
            with pytest.raises(eagen.EagenError) as ei:
                res.padded(0, 2)
            assert ei.value.status == eagen.E_LEN
        res.free()


# ---- domain collisions: x of an intermediate output point on the evaluation domain ------------------------------------
def generator(cv):
    """published generators: Pallas / Vesta (-1, 2), Grumpkin (1, sqrt(-16)); x = -1 and x = 1 lie on EVERY power-of-two domain"""
    if cv.name == "grumpkin":
        return (1, cv.sqrt((1 - 17) % cv.p))
    return (cv.p - 1, 2)


@pytest.mark.parametrize("cname", CURVES)
def test_domain_collision_falls_back_to_isomorphic_curve(gpu_ctx, oracle, eagen, cname):
    """Natural inputs built from the curve generator make an output point's x hit the evaluation domain (pointwise division by
    zero).  Both the canonical and the raw witness must still equal the oracle's: the tree is rebuilt on y^2 = x^3 + u^6 b and
    mapped back (the raw function through its homogeneity degree in u)."""
    cv, ctx = pyref.Curve(cname), gpu_ctx(cname)
    G = generator(cv)
    assert cv.on_curve(G)
    before = ctx.fallback_count()
    # (1) the smallest lhs witness that collides: one point, the generator, scalar 6 = (1, 4, 1) in base -5
    for sc, pts in (([6], [G]), ([6, 11, 3, 124, 0, 77], [G] * 6), (list(range(1, 41)), [G] * 40),
                    ([7 * j + 1 for j in range(60)], [cv.mul(j % 5 + 1, G) for j in range(60)])):
        S, P = oracle.pack_felts(sc, cv.q), oracle.pack_points(pts, cv.p)
        ro = oracle.lhs_witness(cv.id, S, P, 5)
        rg = ctx.compute_lhs_witness(S, P, 5, eagen.CANONICAL)
        rr = ctx.compute_lhs_witness(S, P, 5, eagen.RAW_TREE)
        assert (rg.carries == ro.carries).all() and (rg.carry == ro.carry).all() and (rr.carries == ro.carries).all()
        for k in range(ro.d):
            f, fr = rg.function(k), rr.function(k)
            assert f.a.shape == ro.ca[k].shape and (f.a == ro.ca[k]).all(), (len(sc), k)
            assert f.b.shape == ro.cb[k].shape and (f.b == ro.cb[k]).all(), (len(sc), k)
            ta, tb = trim_rows(ro.a[k]), trim_rows(ro.b[k])
            assert fr.a.shape == ta.shape and (fr.a == ta).all() and fr.b.shape == tb.shape and (fr.b == tb).all(), ("raw", len(sc), k)
        rg.free()
        rr.free()
    assert ctx.fallback_count() > before, "these inputs are expected to exercise the fallback"
    # (2) stand-alone divisor witness whose two child outputs are +G and -G: the denominator (x - x_G)^2 vanishes on the domain
    lst = [cv.mul(2, G), cv.neg(cv.mul(3, G)), G]
    P = oracle.pack_points(lst, cv.p)
    before = ctx.fallback_count()
    f = ctx.compute_divisor_witness(P)
    r = oracle.divisor_witness(cv.id, P)
    assert f.a.shape == r.ca[0].shape and (f.a == r.ca[0]).all() and f.b.shape == r.cb[0].shape and (f.b == r.cb[0]).all()
    assert ctx.fallback_count() == before + 1
    fr = ctx.compute_divisor_witness(P, eagen.RAW_TREE)
    ta, tb = trim_rows(r.a[0]), trim_rows(r.b[0])
    assert fr.a.shape == ta.shape and (fr.a == ta).all() and fr.b.shape == tb.shape and (fr.b == tb).all()
    assert ctx.fallback_count() == before + 2
    # (3) partial form: the output point comes back on the original curve
    lst = [cv.mul(2, G), cv.neg(cv.mul(3, G)), G, cv.mul(5, G), cv.mul(9, G)]
    P = oracle.pack_points(lst, cv.p)
    f, out = ctx.compute_divisor_witness_partial(P)
    r = oracle.divisor_witness(cv.id, P, partial=True)
    assert (f.a == r.ca[0]).all() and (f.b == r.cb[0]).all() and (out == r.output).all()
    fr, out = ctx.compute_divisor_witness_partial(P, eagen.RAW_TREE)
    assert (fr.a == trim_rows(r.a[0])).all() and (fr.b == trim_rows(r.b[0])).all() and (out == r.output).all()
