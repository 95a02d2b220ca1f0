"""CPU tests: the reference's own property tests re-expressed against the oracle (not gpu).

reference tests mirrored: lhs_test (src/argument_witness_calc.rs:138-148), negbase_test
(src/negbase_utils.rs:126-134), poly_test / linefunc_test / randpoints_witness_test /
witness_with_zeros_test (src/regular_functions_utils.rs:554-579,636-671).
"""
import numpy as np
import pytest

import pyref

CURVES = ["pallas", "vesta", "grumpkin"]


def gen(cv, n, seed):
    rng = pyref.SplitMix64(seed)
    p0, dl = pyref.random_point(rng, cv), pyref.random_point(rng, cv)
    pts = [p0]
    for _ in range(n - 1):
        pts.append(cv.add(pts[-1], dl))
    return pts, [pyref.random_scalar(rng, cv) for _ in range(n)]


@pytest.mark.parametrize("cname", CURVES)
def test_lhs_carry_equals_msm_and_functions_vanish(oracle, cname):
    cv = pyref.Curve(cname)
    n = 200
    pts, sc = gen(cv, n, 11)
    S, P = oracle.pack_felts(sc, cv.q), oracle.pack_points(pts, cv.p)
    r = oracle.lhs_witness(cv.id, S, P, 5)
    assert (oracle.msm_naive(cv.id, S, P) == r.carry).all()
    # rebuild tmp_k from digits / carries and check f_k vanishes on every point of it
    d = r.d
    carries = oracle.unpack_affine(r.carries, cv.p)
    mult = [[cv.mul(k, Pj) for k in range(1, 5)] for Pj in pts]
    for i in (0, 1, d // 2, d - 1):
        prev = carries[i - 1] if i else None
        tmp = ([cv.neg(prev)] * 5 if prev is not None else []) + \
              [mult[j][r.digits[j][i] - 1] for j in range(n) if r.digits[j][i]] + [cv.neg(carries[i])]
        k = d - 1 - i  # ret.reverse()
        for T in tmp:
            if T is None:
                continue
            v = oracle.eval_function(cv.id, r.a[k], r.b[k], oracle.pack_points([T], cv.p)[0])
            assert not v.any()
        nn = sum(1 for T in tmp if T is not None)
        ca, cb = r.ca[k], r.cb[k]
        assert len(ca) == nn // 2 + 1 and len(cb) == max((nn - 3) // 2 + 1, 0)


def test_lhs_repeated_point_like_reference_test(oracle):
    """lhs_test shape: ONE point and ONE scalar repeated (reference: src/argument_witness_calc.rs:141-145)"""
    cv = pyref.Curve("grumpkin")
    pts, sc = gen(cv, 1, 5)
    n = 300
    S, P = oracle.pack_felts(sc * n, cv.q), oracle.pack_points(pts * n, cv.p)
    r = oracle.lhs_witness(cv.id, S, P, 5)
    assert (oracle.msm_naive(cv.id, S, P) == r.carry).all()
    assert oracle.unpack_affine(r.carry, cv.p)[0] == cv.mul(sc[0] * n % cv.q, pts[0])


def test_negbase_roundtrip(oracle):
    rng = pyref.SplitMix64(3)
    for _ in range(200):
        x = rng.next_bits(4) >> (rng.next() % 250)
        for base in (2, 5, 17, 255):
            for v in (x, -x):
                dg = oracle.negbase_decompose(v, base)
                assert all(0 <= t < base for t in dg)
                assert sum(t * (-base) ** i for i, t in enumerate(dg)) == v


def test_poly_identities(oracle):
    """poly_test: sizes 100 x 423 take the FFT route (reference: src/regular_functions_utils.rs:554-579)"""
    for field, fid in (("bn256_fr", 2), ("pallas_fp", 0), ("pallas_fq", 1)):
        p = pyref.FIELDS[field]
        rng = pyref.SplitMix64(9)
        a = [rng.next_bits(4) % p for _ in range(100)]
        b = [rng.next_bits(4) % p for _ in range(423)]
        A, B = oracle.pack_felts(a, p), oracle.pack_felts(b, p)
        best = oracle.unpack_felts(oracle.poly_mul(fid, A, B, 0), p)
        naive = oracle.unpack_felts(oracle.poly_mul(fid, A, B, 1), p)
        fftp = oracle.unpack_felts(oracle.poly_mul(fid, A, B, 2), p)
        assert best == naive == fftp and len(best) == 522
        t = rng.next_bits(4) % p
        assert pyref.peval(best, t, p) == pyref.peval(a, t, p) * pyref.peval(b, t, p) % p
        # forward + inverse transform round trip, scaled by 2^-k (mul_fft's convention)
        v = oracle.pack_felts(a[:64], p)
        back = oracle.unpack_felts(oracle.fft(fid, oracle.fft(fid, v), inverse=True), p)
        assert [x * pow(64, -1, p) % p for x in back] == a[:64]


@pytest.mark.parametrize("cname", CURVES)
def test_linefunc_and_zero_vectors(oracle, cname):
    cv = pyref.Curve(cname)
    pts, _ = gen(cv, 3, 21)
    p1, p2 = pts[0], pts[1]
    p3 = cv.neg(cv.add(p1, p2))
    r = oracle.divisor_witness(cv.id, oracle.pack_points([p1, p2, p3], cv.p))
    for T in (p1, p2, p3):
        assert not oracle.eval_function(cv.id, r.a[0], r.b[0], oracle.pack_points([T], cv.p)[0]).any()
    a = pts[2]
    zeros = [None, None, None, a, a, cv.neg(a), None, cv.neg(a), a, cv.neg(a)]
    r = oracle.divisor_witness(cv.id, oracle.pack_points(zeros, cv.p))
    for T in zeros:
        if T is not None:
            assert not oracle.eval_function(cv.id, r.a[0], r.b[0], oracle.pack_points([T], cv.p)[0]).any()
    with pytest.raises(oracle.OracleError):
        oracle.divisor_witness(cv.id, oracle.pack_points([p1, p2], cv.p))


def test_randpoints_witness_repeated_point(oracle):
    """randpoints_witness_test shape: one point repeated + (-sum) (reference: src/regular_functions_utils.rs:650-662)"""
    cv = pyref.Curve("grumpkin")
    pts, _ = gen(cv, 1, 33)
    n = 1000
    last = cv.neg(cv.mul(n, pts[0]))
    r = oracle.divisor_witness(cv.id, oracle.pack_points(pts * n + [last], cv.p))
    for T in (pts[0], last):
        assert not oracle.eval_function(cv.id, r.a[0], r.b[0], oracle.pack_points([T], cv.p)[0]).any()


def test_error_paths(oracle):
    cv = pyref.Curve("pallas")
    pts, sc = gen(cv, 2, 1)
    P = oracle.pack_points(pts, cv.p)
    with pytest.raises(oracle.OracleError):  # scalar >= sqrt(p)+2 (reference: src/argument_witness_calc.rs:97)
        oracle.lhs_witness(cv.id, oracle.pack_felts([2 ** 127 + 2, 1], cv.q), P, 5)
    with pytest.raises(oracle.OracleError):  # base 2 needs more than d digits for the top of the range
        oracle.lhs_witness(2, oracle.pack_felts([pyref.isqrt(pyref.FIELDS["bn256_fq"]) + 1, 1], pyref.FIELDS["bn256_fq"]),
                           oracle.pack_points(gen(pyref.Curve("grumpkin"), 2, 1)[0], pyref.FIELDS["bn256_fr"]), 2)


def test_curve_parameters_pair_base_and_scalar_fields():
    """The published generators lie on y^2 = x^3 + b and have the scalar field's order: Pallas / Vesta (-1, 2) with b = 5
    (pasta_curves), Grumpkin (1, sqrt(-16)) with b = -17 (halo2curves) -- pins which field is C::Base and which is C::Scalar."""
    for cname in ("pallas", "vesta", "grumpkin"):
        cv = pyref.Curve(cname)
        if cname == "grumpkin":
            y = cv.sqrt((1 - 17) % cv.p)
            assert y is not None
            G = (1, y)
        else:
            G = (cv.p - 1, 2)
        assert cv.on_curve(G)
        assert cv.mul(cv.q, G) is None and cv.mul(cv.q - 1, G) == cv.neg(G)
        assert pyref.isqrt(cv.q) + 2 < 1 << 128   # the witness path's scalar range fits 128 bits (K1 relies on it)
