"""CPU tests (not gpu) for the SURVEY.md section 8f rows beyond the hot path: prepare_scalar_witness
(reference: src/negbase_utils.rs:79-124) and compute_divisor_witness_naive (src/regular_functions_utils.rs:483-551).
The C++ oracle is checked against the independent Python big-int restatement and against properties of the result."""
import pytest

import pyref

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


# ---- prepare_scalar_witness -----------------------------------------------------------------------------
def test_psw_known_small_case():
    # 7 in base -5: 7 = 2 + (-5)*(-1) -> digits [2, 4, 1] (2 - 20 + 25)
    assert pyref.negbase_decompose(7, 5) == [2, 4, 1]
    rows = pyref.prepare_scalar_witness(7, 5, 4, 2, intended=True)
    assert rows[0][0] == ("scalar", 7)
    assert rows[2][0] == ("bucket", 1) and rows[4][0] == ("bucket", -5) and rows[1][0] == ("bucket", 25) and rows[3][0] == ("bucket", 0)
    # limbs of two digits: limb 1 = digits 0,1 ; limb 2 = digits 2,3
    assert rows[0][1] == ("limb", 1 - 5, 0b11) and rows[0][2] == ("limb", 1, 0b01)
    assert rows[2][1] == ("limb", 1, 0b01) and rows[4][1] == ("limb", -5, 0b10) and rows[1][2] == ("limb", 1, 0b01)
    # the buckets recombine to the scalar: sum_k k * bucket_k
    assert sum(k * rows[k][0][1] for k in range(1, 5)) == 7


@pytest.mark.parametrize("mode", [0, 1])
def test_psw_oracle_matches_pyref(oracle, mode):
    rng = pyref.SplitMix64(77 + mode)
    cv = pyref.Curve("pallas")
    sq = pyref.isqrt(cv.q) + 2
    for base in (2, 3, 5, 16, 17, 255):
        d = pyref.num_digits(cv, base)
        for logtable in (1, 3, 8, 13, 24):
            num_limbs = (d + logtable - 1) // logtable
            for x in [0, 1, base, sq - 1] + [rng.next_bits(2) % sq for _ in range(4)]:
                try:
                    ref = pyref.prepare_scalar_witness(x, base, d, logtable, intended=bool(mode))
                except IndexError:
                    with pytest.raises(oracle.OracleError):
                        oracle.prepare_scalar_witness(x, base, d, logtable, mode)
                    continue
                got = oracle.prepare_scalar_witness(x, base, d, logtable, mode)
                assert got == ref, (base, logtable, x)
                assert len(got) == base and len(got[0]) == num_limbs + 1


def test_psw_intended_mode_recombines(oracle):
    """intended semantics: sum_k k * sum_limbs value * (-b)^(logtable*(limb-1)) == scalar, and row 0 masks mark the non-zero digits"""
    rng = pyref.SplitMix64(5)
    cv = pyref.Curve("vesta")
    sq = pyref.isqrt(cv.q) + 2
    base, logtable = 5, 8
    d = pyref.num_digits(cv, base)
    for _ in range(20):
        x = rng.next_bits(2) % sq
        rows = oracle.prepare_scalar_witness(x, base, d, logtable, 1)
        digits = pyref.negbase_decompose(x, base)
        total = 0
        for k in range(1, base):
            assert rows[k][0][1] == sum((-base) ** i for i, dg in enumerate(digits) if dg == k)
            for l in range(1, len(rows[k])):
                total += k * rows[k][l][1] * (-base) ** (logtable * (l - 1))
        assert total == x
        for l in range(1, len(rows[0])):
            want = sum(1 << (i % logtable) for i, dg in enumerate(digits) if dg and i // logtable == l - 1)
            assert rows[0][l][2] == want


# ---- compute_divisor_witness_naive ------------------------------------------------------------------------
@pytest.mark.parametrize("cname", CURVES)
@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 8, 33, 100])
def test_naive_oracle_matches_pyref(oracle, cname, n):
    cv = pyref.Curve(cname)
    pts = close_sum(gen_points(cv, n, 3 * n + 1), cv)
    rpos, rneg = pyref.divisor_witness_naive(pts, cv)
    pos, neg = oracle.divisor_witness_naive(cv.id, oracle.pack_points(pts, cv.p))
    assert [tuple(oracle.unpack_felts(l, cv.p)) for l in pos] == rpos
    assert [tuple(oracle.unpack_felts(l, cv.p)) for l in neg] == rneg


def test_naive_degenerate_inputs(oracle):
    """identity points in the list (skipped as `inc1`, accepted as partner), repeated points (tangent fallback), P / -P pairs"""
    cv = pyref.Curve("pallas")
    g = gen_points(cv, 6, 9)
    cases = [
        [g[0], None, cv.neg(g[0])],
        [None, g[0], g[1], None, None, cv.neg(cv.add(g[0], g[1]))],
        [g[0], g[0], g[0], cv.neg(cv.mul(3, g[0]))],
        [g[0], cv.neg(g[0]), g[1], cv.neg(g[1])],
        close_sum([g[2]] * 9, cv),
        [None, None],
        [],
    ]
    for pts in cases:
        rpos, rneg = pyref.divisor_witness_naive(pts, cv)
        pos, neg = oracle.divisor_witness_naive(cv.id, oracle.pack_points(pts, cv.p))
        assert [tuple(oracle.unpack_felts(l, cv.p)) for l in pos] == rpos
        assert [tuple(oracle.unpack_felts(l, cv.p)) for l in neg] == rneg
    with pytest.raises(oracle.OracleError):
        oracle.divisor_witness_naive(cv.id, oracle.pack_points([g[0], g[1]], cv.p))


def test_naive_arrangement_is_the_divisor_witness(oracle):
    """prod(pos lines) / prod(neg lines) has the same divisor as compute_divisor_witness: the quotient of the two is the same
    constant at every point of the curve"""
    cv = pyref.Curve("grumpkin")
    p = cv.p
    pts = close_sum(gen_points(cv, 21, 4), cv)
    P = oracle.pack_points(pts, cv.p)
    pos, neg = oracle.divisor_witness_naive(cv.id, P)
    r = oracle.divisor_witness(cv.id, P)
    fa, fb = oracle.unpack_felts(r.a[0], p), oracle.unpack_felts(r.b[0], p)

    def ev_lines(lines, X, Y):
        v = 1
        for l in lines:
            lx, ly, lz = oracle.unpack_felts(l, p)
            v = v * ((lz + lx * X + ly * Y) % p) % p
        return v

    ratios = set()
    for Q in gen_points(cv, 5, 1234):
        X, Y = Q
        f = (pyref.peval(fa, X, p) + Y * pyref.peval(fb, X, p)) % p
        arr = ev_lines(pos, X, Y) * pow(ev_lines(neg, X, Y), p - 2, p) % p
        ratios.add(arr * pow(f, p - 2, p) % p)
    assert len(ratios) == 1 and 0 not in ratios
