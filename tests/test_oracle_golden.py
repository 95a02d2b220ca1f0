"""CPU tests: the oracle against the committed golden fixtures (not gpu).

Pins (a) the oracle's field arithmetic / Montgomery form to the reference's own bn256::Fr FFT tables
(reference: src/precomputed_fft_data.rs:3-216), (b) negbase known answers, (c) digits, carries and
raw + canonical polynomials to the independent Python restatement's vectors.
"""
import json
import os

import numpy as np
import pytest

import pyref

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_bn256_fr_fft_tables_match_reference(oracle):
    t = load("bn256_fr_fft_tables.json")
    fid = pyref.FIELD_ID["bn256_fr"]
    for k in range(64):
        assert oracle.omega_pow(fid, k).tobytes().hex() == t["omega_pow"][k], k
        assert oracle.omega_pow_inv(fid, k).tobytes().hex() == t["omega_pow_inv"][k], k
        assert oracle.half_pow(fid, k).tobytes().hex() == t["half_pow"][k], k


def test_reference_tables_are_consistent():
    """independent numeric check of the fixture itself: omega = 7^((r-1)/2^28), squaring chain, 2^-k"""
    t = load("bn256_fr_fft_tables.json")
    p = pyref.FIELDS["bn256_fr"]
    rinv = pow(pyref.R, -1, p)
    om = [int.from_bytes(bytes.fromhex(h), "little") * rinv % p for h in t["omega_pow"]]
    omi = [int.from_bytes(bytes.fromhex(h), "little") * rinv % p for h in t["omega_pow_inv"]]
    hp = [int.from_bytes(bytes.fromhex(h), "little") * rinv % p for h in t["half_pow"]]
    assert om[0] == pow(7, (p - 1) >> 28, p)
    for k in range(63):
        assert om[k + 1] == om[k] * om[k] % p
        assert om[k] * omi[k] % p == 1
        assert hp[k] * pow(2, k, p) % p == 1
    assert om[27] == p - 1 and om[28] == 1


def test_negbase_known_answers(oracle):
    ka = load("negbase_known_answers.json")
    for c in ka["negbase"]:
        x = int(c["x"])
        got = oracle.negbase_decompose(x, c["base"])
        assert got == c["digits"], c
        assert sum(d * (-c["base"]) ** i for i, d in enumerate(got)) == x
    # SURVEY section 8c spot values
    assert oracle.negbase_decompose(5, 5) == [0, 4, 1]
    assert oracle.negbase_decompose(6, 5) == [1, 4, 1]
    assert oracle.negbase_decompose(24, 5) == [4, 1, 1]
    assert oracle.negbase_decompose(25, 5) == [0, 0, 1]
    assert oracle.negbase_decompose(123456789, 17) == [1, 15, 11, 15, 0, 15, 6]
    assert "".join(map(str, oracle.negbase_decompose(123456789, 5))) == "4321142012331"
    assert len(oracle.negbase_decompose(2 ** 127 + 1, 5)) == 55
    for c in ka["table_entry"]:
        p = pyref.FIELDS[c["field"]]
        got = oracle.unpack_felts(oracle.table_entry_by_id(pyref.FIELD_ID[c["field"]], c["base"], c["id"]), p)[0]
        assert "%064x" % got == c["value"]
    p = pyref.FIELDS["pallas_fp"]
    vals = {1: -5, 2: 25, 3: 20, 5: -130, 11: 645}
    for idx, v in vals.items():
        assert oracle.unpack_felts(oracle.table_entry_by_id(0, 5, idx), p)[0] == v % p


def test_num_digits(oracle):
    for name in ("pallas", "vesta", "grumpkin"):
        cv = pyref.Curve(name)
        for base in (2, 3, 4, 5, 16, 17, 255):
            assert oracle.num_digits(cv.id, base) == pyref.num_digits(cv, base)
    assert oracle.num_digits(0, 5) == 56 and oracle.num_digits(0, 2) == 129 and oracle.num_digits(0, 4) == 65


def _pts(case):
    return [None if P is None else (int(P[0], 16), int(P[1], 16)) for P in case["points"]]


def _felts(lst):
    return [int(x, 16) for x in lst]


@pytest.mark.parametrize("case", load("witness_small.json"), ids=lambda c: c["curve"] + "-" + c["name"])
def test_witness_small(oracle, case):
    cv = pyref.Curve(case["curve"])
    pts = _pts(case)
    # Jacobian inputs with non-trivial z: the result must not depend on the representation
    zs = [(7 * i + 3) % cv.p for i in range(len(pts))]
    P = oracle.pack_points(pts, cv.p, zs)
    if case["kind"] == "lhs":
        sc = _felts(case["scalars"])
        r = oracle.lhs_witness(cv.id, oracle.pack_felts(sc, cv.q), P, case["base"])
        assert r.digits.tolist() == case["digits"]
        want = [None if c is None else (int(c[0], 16), int(c[1], 16)) for c in case["carries"]]
        assert oracle.unpack_affine(r.carries, cv.p) == want
        assert oracle.unpack_affine(r.carry, cv.p)[0] == want[-1]
        for k in range(len(case["raw"])):
            assert oracle.unpack_felts(r.a[k], cv.p) == _felts(case["raw"][k][0])
            assert oracle.unpack_felts(r.b[k], cv.p) == _felts(case["raw"][k][1])
            assert oracle.unpack_felts(r.ca[k], cv.p) == _felts(case["canonical"][k][0])
            assert oracle.unpack_felts(r.cb[k], cv.p) == _felts(case["canonical"][k][1])
    else:
        r = oracle.divisor_witness(cv.id, P, partial=True)
        out = None if case["output"] is None else (int(case["output"][0], 16), int(case["output"][1], 16))
        assert oracle.unpack_affine(r.output, cv.p)[0] == out
        assert oracle.unpack_felts(r.a[0], cv.p) == _felts(case["raw"][0])
        assert oracle.unpack_felts(r.b[0], cv.p) == _felts(case["raw"][1])
        assert oracle.unpack_felts(r.ca[0], cv.p) == _felts(case["canonical"][0])
        assert oracle.unpack_felts(r.cb[0], cv.p) == _felts(case["canonical"][1])
