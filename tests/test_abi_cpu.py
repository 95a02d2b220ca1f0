"""CPU tests of the C-ABI library (not gpu): it loads, exports every symbol include/*.h declares, refuses to run
without a device (no CPU fallback), and its __host__ __device__ arithmetic -- the very source the kernels are
compiled from -- matches the oracle when run on the host through the self-test hooks."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eagen_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(eagen):
    L = eagen.lib()
    decl = declared_functions("eagen_msm.h")
    assert len(decl) >= 30
    for name in decl + declared_functions("eagen_msm_selftest.h"):
        assert hasattr(L, name), name
    assert sorted(eagen.ABI_SYMBOLS) == decl


def test_no_device_means_error_not_fallback(eagen):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(eagen.EagenError) as e:
        eagen.Context("pallas", 0)
    assert e.value.status == eagen.E_NO_DEVICE


def test_num_digits_matches_reference_formula(eagen):
    for name in ("pallas", "vesta", "grumpkin"):
        cv = pyref.Curve(name)
        for base in (2, 3, 4, 5, 16, 17, 255):
            assert eagen.num_digits(cv.id, base) == pyref.num_digits(cv, base)


def test_fft_precomp_matches_reference_tables_and_oracle(eagen, oracle):
    import json
    t = json.load(open(os.path.join(ROOT, "tests", "golden", "bn256_fr_fft_tables.json")))
    for k in range(64):
        assert eagen.omega_pow(eagen.GRUMPKIN, k).tobytes().hex() == t["omega_pow"][k]
        assert eagen.omega_pow_inv(eagen.GRUMPKIN, k).tobytes().hex() == t["omega_pow_inv"][k]
        assert eagen.half_pow(eagen.GRUMPKIN, k).tobytes().hex() == t["half_pow"][k]
    for cname, fid in (("pallas", 0), ("vesta", 1)):
        for k in (0, 1, 5, 31, 32, 40):
            assert (eagen.omega_pow(eagen.CURVE_IDS[cname], k) == oracle.omega_pow(fid, k)).all()
            assert (eagen.omega_pow_inv(eagen.CURVE_IDS[cname], k) == oracle.omega_pow_inv(fid, k)).all()
            assert (eagen.half_pow(eagen.CURVE_IDS[cname], k) == oracle.half_pow(fid, k)).all()


@pytest.mark.parametrize("field", ["pallas_fp", "pallas_fq", "bn256_fr", "bn256_fq"])
def test_host_field_arithmetic_matches_oracle(eagen, oracle, field):
    p, fid = pyref.FIELDS[field], pyref.FIELD_ID[field]
    rng = pyref.SplitMix64(100 + fid)
    vals = [0, 1, p - 1, p - 2, 2, (1 << 255) % p, (1 << 128) - 1, (1 << 256) % p, p >> 1] + [rng.next_bits(4) % p for _ in range(400)]
    arr = oracle.pack_felts(vals, p)
    for i in range(len(vals)):
        a, b = arr[i], arr[(i * 7 + 3) % len(vals)]
        for op in (0, 1, 2):
            assert (eagen.selftest_field(fid, op, a, b) == oracle.field_op(fid, op, a, b)).all(), (field, op, i)
        # op 6: the device's carry-chain Montgomery product, carry flag emulated on the host
        assert (eagen.selftest_field(fid, 6, a, b) == oracle.field_op(fid, 2, a, b)).all(), (field, 'chain', i)
        assert (eagen.selftest_field(fid, 3, a) == oracle.field_op(fid, 3, a)).all()
        assert (eagen.selftest_field(fid, 5, a) == oracle.field_op(fid, 5, a)).all()
    # raw canonical limbs -> Montgomery
    raw = np.array([(vals[9] >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    assert (eagen.selftest_field(fid, 4, raw) == arr[9]).all()


@pytest.mark.parametrize("cname", ["pallas", "vesta", "grumpkin"])
def test_host_complete_curve_formulas_match_oracle(eagen, oracle, cname):
    cv = pyref.Curve(cname)
    rng = pyref.SplitMix64(55)
    P, Q = pyref.random_point(rng, cv), pyref.random_point(rng, cv)
    zs = [rng.next_bits(4) % cv.p for _ in range(2)]
    cases = [(P, Q), (P, P), (P, cv.neg(P)), (None, Q), (P, None), (None, None)]
    for A, B in cases:
        ja, jb = oracle.pack_points([A], cv.p, zs[:1])[0], oracle.pack_points([B], cv.p, zs[1:])[0]
        want = cv.add(A, B)
        got = oracle.unpack_affine(eagen.selftest_curve(cv.id, 0, ja, jb), cv.p)[0]
        assert got == want
        if B is not None:
            jb1 = oracle.pack_points([B], cv.p)[0]
            assert oracle.unpack_affine(eagen.selftest_curve(cv.id, 2, ja, jb1[:8]), cv.p)[0] == want
        assert oracle.unpack_affine(eagen.selftest_curve(cv.id, 1, ja), cv.p)[0] == cv.add(A, A)
    jp = oracle.pack_points([P], cv.p, zs[:1])[0]
    for k in (0, 1, 2, 3, 5, 17, 255):
        assert oracle.unpack_affine(eagen.selftest_curve(cv.id, 3, jp, None, k), cv.p)[0] == cv.mul(k, P)


def test_negbase_constants(eagen):
    """K1's offset trick: digits of x in base -b = base-b digits of x + K with odd positions complemented; and the
    fixed-point reciprocal that replaces every division"""
    for cname in ("pallas", "grumpkin"):
        cv = pyref.Curve(cname)
        for base in (2, 3, 5, 17, 255):
            prm = eagen.selftest_negbase_params(cv.id, base)
            d = pyref.num_digits(cv, base)
            assert prm["d"] == d and prm["sq"] == pyref.isqrt(cv.q) + 2
            assert prm["K"] == sum((base - 1) * base ** i for i in range(1, d, 2)) and prm["bd"] == base ** d
            assert prm["inv"] == -(-(1 << 288) // base ** d) and prm["bd"] < 1 << 143
            assert prm["words"] == (d + 3) // 4 and prm["group"] in (1, 2, 4) and base ** prm["group"] <= 1024
            rng = pyref.SplitMix64(base)
            for _ in range(50):
                x = rng.next_bits(2) % prm["sq"]
                y = x + prm["K"]
                if y >= prm["bd"]:
                    continue
                e = [(y // base ** i) % base for i in range(d)]
                dg = [(base - 1 - e[i]) if i & 1 else e[i] for i in range(d)]
                ref = pyref.negbase_decompose(x, base)
                assert dg == ref + [0] * (d - len(ref))


def test_negbase_kernel_arithmetic_on_host(eagen, oracle):
    """the kernel's own per-scalar source (Montgomery -> canonical, division-free digit extraction, complement table) run on
    the host against the independent big-int negbase_decompose: every base, edge scalars and random ones"""
    for cname in ("pallas", "vesta", "grumpkin"):
        cv = pyref.Curve(cname)
        sq = pyref.isqrt(cv.q) + 2
        rng = pyref.SplitMix64(hash(cname) & 0xffff)
        for base in list(range(2, 40)) + [63, 64, 100, 127, 128, 200, 254, 255]:
            d = pyref.num_digits(cv, base)
            xs = [0, 1, 2, base - 1, base, base + 1, base ** 2, sq - 1, sq - 2, (1 << 64) - 1, 1 << 64, base ** (d - 2), base ** (d - 2) - 1]
            xs += [rng.next_bits(2) % sq for _ in range(12)]
            for x in xs:
                if x >= sq:
                    continue
                ref = pyref.negbase_decompose(x, base)
                got, kerr = eagen.selftest_negbase_digits(cv.id, base, oracle.pack_felts([x], cv.q)[0])
                if len(ref) > d:
                    assert kerr == 2
                    continue
                assert kerr == 0 and list(got) == ([0] * (d - len(ref)) + ref[::-1]), (cname, base, x)
        # out of range: >= isqrt(order) + 2
        _, kerr = eagen.selftest_negbase_digits(cv.id, 5, oracle.pack_felts([sq], cv.q)[0])
        assert kerr == 1


def test_ntt_pass_plan(eagen):
    for t in range(1, 33):
        plan = eagen.selftest_ntt_plan(t)
        stages = []
        for hi, lo in plan:
            assert 1 <= hi - lo + 1 <= 10
            stages += list(range(hi, lo - 1, -1))
        assert stages == list(range(t - 1, -1, -1))


def test_cpp_host_mirror_compiles_links_and_runs(eagen, tmp_path):
    """include/eagen_msm.hpp (the C++ mirror of the reference API) builds against the library and its GPU-free calls work"""
    import subprocess
    exe = str(tmp_path / "host_mirror_check")
    libdir = os.path.dirname(eagen.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_check.cpp"),
                           "-L", libdir, "-leagen_msm", "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_table_entry_by_id_matches_oracle_and_known_answers(eagen, oracle):
    p = pyref.FIELDS["pallas_fp"]
    for idx, v in {0: 0, 1: -5, 2: 25, 3: 20, 5: -130, 11: 645}.items():
        assert oracle.unpack_felts(eagen.table_entry_by_id(eagen.PALLAS, 5, idx), p)[0] == v % p
    for cname, fid in (("pallas", 0), ("vesta", 1), ("grumpkin", 2)):
        for base, idx in ((5, 1023), (17, 77), (2, 32767), (255, 9)):
            assert (eagen.table_entry_by_id(eagen.CURVE_IDS[cname], base, idx) == oracle.table_entry_by_id(fid, base, idx)).all()


def test_challenge_point_helpers(eagen, oracle):
    """to_curve_x / y_from_x / slope (reference: src/config.rs:163-187): host-side helpers of the C ABI against Python ints"""
    for cname in ("pallas", "vesta", "grumpkin"):
        cv = pyref.Curve(cname)
        p = cv.p
        rng = pyref.SplitMix64(len(cname))
        seen = set()
        for _ in range(24):
            x = rng.next_bits(4) % p
            rhs = (x ** 3 + cv.b) % p
            want_sq, want_y = pyref.sqrt_ff(cv.base_field, rhs)
            X = oracle.pack_felts([x], p)[0]
            y, is_sq = eagen.y_from_x(cv.id, X)
            yv = oracle.unpack_felts(y, p)[0]
            assert bool(is_sq) == want_sq and yv == want_y
            seen.add(want_sq)
            if want_sq:
                assert yv * yv % p == rhs
                assert (eagen.to_curve_x(cv.id, X) == X).all()
                sl = oracle.unpack_felts(eagen.slope(cv.id, X, y), p)[0]
                assert sl == 3 * x * x * pow(2 * yv, -1, p) % p
            else:
                with pytest.raises(eagen.EagenError) as ei:
                    eagen.to_curve_x(cv.id, X)
                assert ei.value.status == eagen.E_DOMAIN
        assert seen == {True, False}
        with pytest.raises(eagen.EagenError) as ei:
            eagen.slope(cv.id, oracle.pack_felts([3], p)[0], oracle.pack_felts([0], p)[0])
        assert ei.value.status == eagen.E_DOMAIN
    assert eagen.circuit_sizes(1000, 5) == (503, 503) and eagen.circuit_sizes(2, 2) == (3, 2)


def test_pasta_fft_precomp_matches_published_root_of_unity(eagen, oracle):
    """The reference only carries FftPrecomp tables for bn256::Fr; the Pasta tables are new.  Their seed must be the
    PrimeField::ROOT_OF_UNITY pasta_curves publishes (= 5^((p-1)/2^32), the recipe of src/scripts.rs:44-70), and the table
    entries its squaring chain, its inverse and powers of 1/2."""
    published = {
        "pallas": 0x2bce74deac30ebda362120830561f81aea322bf2b7bb7584bdad6fabd87ea32f,   # pasta_curves Fp::ROOT_OF_UNITY
        "vesta": 0x2de6a9b8746d3f589e5c4dfd492ae26e9bb97ea3c106f049a70e2c1102b6d05f,    # pasta_curves Fq::ROOT_OF_UNITY
    }
    for cname, root in published.items():
        cv = pyref.Curve(cname)
        p = cv.p
        assert root == pow(5, (p - 1) >> 32, p) and pow(root, 1 << 31, p) == p - 1
        for k in (0, 1, 7, 31, 32, 63):
            w = oracle.unpack_felts(eagen.omega_pow(cv.id, k), p)[0]
            wi = oracle.unpack_felts(eagen.omega_pow_inv(cv.id, k), p)[0]
            h = oracle.unpack_felts(eagen.half_pow(cv.id, k), p)[0]
            assert w == pow(root, 1 << min(k, 40), p) and w * wi % p == 1 and h * pow(2, k, p) % p == 1


def test_fft_precomp_large_exponents_terminate(eagen):
    """half_pow(exp) must be square-and-multiply (a 2^60 exponent returns at once) and omega_pow(k) is the identity from
    k = S on instead of silently clamping"""
    cv = pyref.Curve("pallas")
    h = eagen.half_pow(eagen.PALLAS, (1 << 60) + 12345)
    want = pow(pow(2, -1, cv.p), (1 << 60) + 12345, cv.p)
    import oracle_lib
    assert oracle_lib.unpack_felts(h, cv.p)[0] == want
    one = oracle_lib.pack_felts([1], cv.p)[0]
    for k in (32, 33, 1000, (1 << 40)):
        assert (eagen.omega_pow(eagen.PALLAS, k) == one).all()
        assert (eagen.omega_pow_inv(eagen.PALLAS, k) == one).all()


def test_sharding_plan_entry_points_need_no_device(eagen):
    """eagen_position_range / eagen_lhs_witness_sharded_layout are pure host arithmetic: the d positions are partitioned into
    contiguous, balanced ranges, and a rank's streamed buffer holds exactly its positions' slots"""
    for d in (33, 56, 65, 129):
        for world in (1, 2, 3, 4, 8):
            rs = [eagen.position_range(r, world, d) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == d and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    with pytest.raises(eagen.EagenError):
        eagen.position_range(8, 8, 56)
    L = eagen.lib()
    d = eagen.num_digits(eagen.PALLAS, 5)
    a, b, tot = C.c_size_t(), C.c_size_t(), C.c_size_t()
    whole = 0
    for r in range(8):
        assert L.eagen_lhs_witness_sharded_layout(eagen.PALLAS, 1 << 20, C.c_uint8(5), r, 8, C.byref(a), C.byref(b), C.byref(tot)) == 0
        lo, hi = eagen.position_range(r, 8, d)
        assert tot.value == (hi - lo) * (a.value + b.value) * 32
        whole += tot.value
    a1, b1, t1 = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert L.eagen_lhs_witness_stream_layout(eagen.PALLAS, 1 << 20, C.c_uint8(5), C.byref(a1), C.byref(b1), C.byref(t1)) == 0
    assert (a1.value, b1.value) == (a.value, b.value) and t1.value == whole


def test_oracle_synthetic_inputs_are_valid(oracle):
    """the CPU restatement of the synthetic inputs (the golden hashes' inputs): scalars below isqrt(order)+2, points on the curve,
    pairwise distinct, a function of the seed only"""
    for name in ("pallas", "vesta", "grumpkin"):
        cv = pyref.Curve(name)
        S, P = oracle.synth_inputs(cv.id, 0xEA6E0002, 64)
        S2, P2 = oracle.synth_inputs(cv.id, 0xEA6E0002, 32)
        assert (S[:32] == S2).all() and (P[:32] == P2).all()
        import math
        lim = math.isqrt(cv.q) + 2
        assert all(s < lim for s in oracle.unpack_felts(S, cv.q))
        pts = oracle.unpack_affine(P[:, :8], cv.p)
        assert len(set(pts)) == 64
        for x, y in pts:
            assert (y * y - x * x * x - cv.b) % cv.p == 0


@pytest.mark.parametrize("field", ["pallas_fp", "pallas_fq", "bn256_fr"])
def test_lazy_reduction_arithmetic_on_the_host(eagen, field):
    """the transform's butterflies work on values in [0, 2p) (csrc/field.cuh: mul_lazy / add_lazy / sub_lazy): run the very source on
    the host (emulated carry flag) and check range and residue class against Python integers, including the corners where a + b
    carries out of 256 bits (4p > 2^256 for the Pasta moduli) and where a - b borrows"""
    fid = {"pallas_fp": 0, "pallas_fq": 1, "bn256_fr": 2}[field]
    p = {"pallas_fp": pyref.Curve("pallas").p, "pallas_fq": pyref.Curve("pallas").q, "bn256_fr": pyref.Curve("grumpkin").p}[field]
    R = 1 << 256
    rinv = pow(R, -1, p)

    def words(v):
        return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)

    def val(w):
        return sum(int(w[i]) << (64 * i) for i in range(4))
    rng = pyref.SplitMix64(2026)
    corner = [0, 1, p - 1, p, p + 1, 2 * p - 1, 2 * p - 2, (2 * p) >> 1, R - 2 * p if R - 2 * p < 2 * p else p]
    vals = corner + [rng.next_bits(4) % (2 * p) for _ in range(40)]
    for i, a in enumerate(vals):
        for b in (vals[(i * 7 + 3) % len(vals)], vals[(i * 5 + 1) % len(vals)], 2 * p - 1, 0):
            s = val(eagen.selftest_field(fid, 8, words(a), words(b)))
            assert s < 2 * p and (s - a - b) % p == 0, (field, "add", a, b)
            d = val(eagen.selftest_field(fid, 9, words(a), words(b)))
            assert d < 2 * p and (d - a + b) % p == 0, (field, "sub", a, b)
            w = b % p                                                   # twiddles are canonical
            m = val(eagen.selftest_field(fid, 7, words(a), words(w)))
            assert m < 2 * p and (m - a * w * rinv) % p == 0, (field, "mul", a, w)
            n = val(eagen.selftest_field(fid, 10, words(m)))
            assert n == a * w * rinv % p, (field, "normalise", a, w)
