#!/usr/bin/env python3
"""Generate the committed golden fixtures (run in the build container, where /root/reference exists).

1. bn256_fr_fft_tables.json -- the ONLY constant data the reference holds for this path: the three
   64-entry Montgomery-form tables of `impl FftPrecomp for bn256::Fr`
   (reference: src/precomputed_fft_data.rs:3-216), parsed from the Rust source as data.
2. negbase_known_answers.json -- negbase_decompose / table_entry_by_id known answers
   (reference: src/negbase_utils.rs:20-36,58-77) computed by the independent Python restatement.
3. witness_small.json -- small compute_lhs_witness / compute_divisor_witness cases on Pallas, Vesta and
   Grumpkin computed by tests/pyref.py (independent big-int implementation), including the reference's
   own edge-case vector `witness_with_zeros_test` (reference: src/regular_functions_utils.rs:664-671)
   and an all-equal-points case mirroring `lhs_test` (reference: src/argument_witness_calc.rs:138-148).

Usage: python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pyref  # noqa: E402

REF = "/root/reference/src/precomputed_fft_data.rs"


def parse_tables():
    src = open(REF).read()
    out = {}
    for fn in ("omega_pow", "omega_pow_inv", "half_pow"):
        m = re.search(r"fn %s\(.*?\{(.*?)_\s*=>" % fn, src, re.S)
        body = m.group(1)
        rows = re.findall(r"(\d+)\s*=>\s*\[([^\]]*)\]", body)
        tab = {}
        for k, lst in rows:
            b = bytes(int(x) for x in lst.split(","))
            assert len(b) == 32
            tab[int(k)] = b.hex()
        out[fn] = [tab[k] for k in sorted(tab)]
        assert sorted(tab) == list(range(len(tab)))
    return out


def hx(v):
    return "%064x" % v


def pt(P):
    return None if P is None else [hx(P[0]), hx(P[1])]


def witness_case(name, cv, scalars, pts, base):
    digits, carries, fns = pyref.lhs_witness(scalars, pts, base, cv)
    canon = [pyref.canonicalize(f, cv.p) for f in fns]
    return dict(name=name, kind="lhs", curve=cv.name, base=base,
                scalars=[hx(s) for s in scalars], points=[pt(P) for P in pts],
                digits=digits, carries=[pt(c) for c in carries],
                raw=[[[hx(c) for c in f[0]], [hx(c) for c in f[1]]] for f in fns],
                canonical=[[[hx(c) for c in f[0]], [hx(c) for c in f[1]]] for f in canon])


def divisor_case(name, cv, pts):
    f, out = pyref.divisor_witness_partial(pts, cv)
    c = pyref.canonicalize(f, cv.p)
    return dict(name=name, kind="divisor", curve=cv.name, points=[pt(P) for P in pts], output=pt(out),
                raw=[[hx(x) for x in f[0]], [hx(x) for x in f[1]]],
                canonical=[[hx(x) for x in c[0]], [hx(x) for x in c[1]]])


def main():
    with open(os.path.join(HERE, "bn256_fr_fft_tables.json"), "w") as f:
        json.dump(dict(source="reference: src/precomputed_fft_data.rs:3-216 (raw Montgomery bytes, little endian)",
                       **parse_tables()), f, indent=0)

    ka = dict(negbase=[], table_entry=[])
    for x, b in [(5, 5), (6, 5), (24, 5), (25, 5), (123456789, 5), (123456789, 17), (0, 5), (2 ** 127 + 1, 5),
                 (2 ** 127 + 1, 2), (2 ** 127 + 1, 255), (-1, 5), (-123456789, 7), (2 ** 200 + 12345, 3), (1, 2)]:
        ka["negbase"].append(dict(x=str(x), base=b, digits=pyref.negbase_decompose(x, b)))
    for field in ("pallas_fp", "bn256_fr"):
        p = pyref.FIELDS[field]
        for base, idx in [(5, 1), (5, 2), (5, 3), (5, 5), (5, 11), (17, 1023)]:
            acc, bits = 0, bin(idx)[2:]
            for bit in bits:
                acc = (acc + int(bit)) * (-base) % p
            ka["table_entry"].append(dict(field=field, base=base, id=idx, value=hx(acc)))
    with open(os.path.join(HERE, "negbase_known_answers.json"), "w") as f:
        json.dump(ka, f, indent=0)

    cases = []
    for ci, cname in enumerate(("pallas", "vesta", "grumpkin")):
        cv = pyref.Curve(cname)
        rng = pyref.SplitMix64(0xEA6E0000 + ci)
        n = 9
        pts = [pyref.random_point(rng, cv) for _ in range(n)]
        sc = [pyref.random_scalar(rng, cv) for _ in range(n)]
        cases.append(witness_case("random9", cv, sc, pts, 5))
        if cname == "pallas":
            # mirrors lhs_test: one point, one scalar, repeated
            cases.append(witness_case("all_equal6", cv, [sc[0]] * 6, [pts[0]] * 6, 5))
            # edge scalars and an identity point, other bases
            edge_sc = [0, 1, 2 ** 127 + 1, 5, sc[1]]
            edge_pts = [pts[0], pts[1], pts[2], None, cv.neg(pts[0])]
            cases.append(witness_case("edge_base5", cv, edge_sc, edge_pts, 5))
            cases.append(witness_case("base3", cv, sc[:4], pts[:4], 3))
            cases.append(witness_case("base17", cv, sc[:5], pts[:5], 17))
        a = pts[0]
        zeros = [None, None, None, a, a, cv.neg(a), None, cv.neg(a), a, cv.neg(a)]
        cases.append(divisor_case("witness_with_zeros", cv, zeros))
        s = None
        for P in pts[:7]:
            s = cv.add(s, P)
        cases.append(divisor_case("seven_plus_negsum", cv, pts[:7] + [cv.neg(s)]))
        cases.append(divisor_case("partial_five", cv, pts[:5]))
    cases.append(divisor_case("all_identity4", pyref.Curve("pallas"), [None] * 4))
    cvp = pyref.Curve("pallas")
    q = pyref.random_point(pyref.SplitMix64(7), cvp)
    cases.append(divisor_case("alternating", cvp, [q, cvp.neg(q)] * 4))
    with open(os.path.join(HERE, "witness_small.json"), "w") as f:
        json.dump(cases, f)
    print("cases:", len(cases))


if __name__ == "__main__":
    main()
