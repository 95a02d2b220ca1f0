"""ctypes loader for the CPU oracle (oracle/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None
U64P = C.POINTER(C.c_uint64)
U8P = C.POINTER(C.c_uint8)


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_last_error.restype = C.c_char_p
        _LIB.oracle_result_seconds.restype = C.c_double
        _LIB.oracle_result_poly_len.restype = C.c_size_t
        _LIB.oracle_result_num_functions.restype = C.c_size_t
        _LIB.oracle_result_poly_len.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        _LIB.oracle_result_poly_copy.argtypes = [C.c_void_p, C.c_size_t, C.c_int, U64P]
        for f in ("oracle_result_free", "oracle_result_d", "oracle_result_seconds", "oracle_result_num_functions"):
            getattr(_LIB, f).argtypes = [C.c_void_p]
        _LIB.oracle_result_digits.argtypes = [C.c_void_p, U8P]
        _LIB.oracle_result_carry.argtypes = [C.c_void_p, U64P]
        _LIB.oracle_result_carries.argtypes = [C.c_void_p, U64P]
    return _LIB


class OracleError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise OracleError(lib().oracle_last_error().decode())


def _p64(a):
    return a.ctypes.data_as(U64P)


def pack_felts(vals, p):
    """list of canonical ints -> (n,4) uint64 Montgomery limbs"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = pyref.to_mont_words(v % p, p)
    return out


def unpack_felts(arr, p):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    rinv = pow(pyref.R, -1, p)
    return [sum(int(w[i]) << (64 * i) for i in range(4)) * rinv % p for w in arr]


def pack_points(pts, p, zs=None):
    """affine tuples / None -> (n,12) Jacobian Montgomery; zs optionally re-randomises z"""
    out = np.zeros((len(pts), 12), dtype=np.uint64)
    for i, P in enumerate(pts):
        if P is None:
            continue
        z = 1 if zs is None else zs[i]
        out[i, 0:4] = pyref.to_mont_words(P[0] * z * z % p, p)
        out[i, 4:8] = pyref.to_mont_words(P[1] * z * z * z % p, p)
        out[i, 8:12] = pyref.to_mont_words(z, p)
    return out


def unpack_affine(arr, p):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 8)
    out = []
    for row in arr:
        if not row.any():
            out.append(None)
        else:
            x, y = unpack_felts(row.reshape(2, 4), p)
            out.append((x, y))
    return out


def set_threads(n):
    _chk(lib().oracle_set_threads(int(n)))


def get_threads():
    return lib().oracle_get_threads()


def num_digits(curve_id, base):
    d = C.c_uint()
    _chk(lib().oracle_num_digits(curve_id, C.c_uint8(base), C.byref(d)))
    return d.value


def negbase_decompose(x, base):
    mag = abs(x)
    w = np.array([(mag >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    out = np.zeros(300, dtype=np.uint8)
    n = C.c_size_t()
    _chk(lib().oracle_negbase_decompose(_p64(w), int(x < 0), C.c_uint8(base), out.ctypes.data_as(U8P), C.byref(n)))
    return [int(v) for v in out[: n.value]]


def table_entry_by_id(field_id, base, idx):
    out = np.zeros(4, dtype=np.uint64)
    _chk(lib().oracle_table_entry_by_id(field_id, C.c_uint8(base), C.c_size_t(idx), _p64(out)))
    return out


def field_op(field_id, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    _chk(lib().oracle_field_op(field_id, op, _p64(a), None if bb is None else _p64(bb), _p64(out)))
    return out


def omega_pow(field_id, k):
    return field_op(field_id, 6, np.zeros(4, np.uint64), np.array([k, 0, 0, 0], np.uint64))


def omega_pow_inv(field_id, k):
    return field_op(field_id, 7, np.zeros(4, np.uint64), np.array([k, 0, 0, 0], np.uint64))


def half_pow(field_id, k):
    return field_op(field_id, 8, np.zeros(4, np.uint64), np.array([k, 0, 0, 0], np.uint64))


def poly_mul(field_id, a, b, mode=0):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    n = len(a) + len(b) - 1 if len(a) + len(b) else 0
    out = np.zeros((max(n, 1), 4), dtype=np.uint64)
    _chk(lib().oracle_poly_mul(field_id, _p64(a), C.c_size_t(len(a)), _p64(b), C.c_size_t(len(b)), _p64(out), mode))
    return out[:n]


def fft(field_id, a, inverse=False):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4).copy()
    log_n = (len(a) - 1).bit_length()
    assert 1 << log_n == len(a)
    _chk(lib().oracle_fft(field_id, _p64(a), C.c_uint(log_n), int(inverse)))
    return a


def msm_naive(curve_id, scalars, pts):
    out = np.zeros(8, dtype=np.uint64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    pts = np.ascontiguousarray(pts, dtype=np.uint64)
    _chk(lib().oracle_msm_naive(curve_id, _p64(scalars), _p64(pts), C.c_size_t(len(pts)), _p64(out)))
    return out


def eval_function(curve_id, a, b, pt):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    pt = np.ascontiguousarray(pt, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    rc = lib().oracle_eval_function(curve_id, _p64(a), C.c_size_t(len(a)), _p64(b), C.c_size_t(len(b)), _p64(pt), _p64(out))
    if rc < 0:
        _chk(rc)
    return None if rc == 1 else out


class Result:
    """digits (n,d) MSD first; carries (d,8); carry (8,); a/b raw and ca/cb canonical lists of (len,4)"""

    def __init__(self, h, n, with_digits=True):
        L = lib()
        self.d = L.oracle_result_d(h)
        self.seconds = L.oracle_result_seconds(h)
        nf = L.oracle_result_num_functions(h)
        if with_digits and self.d:
            self.digits = np.zeros((n, self.d), dtype=np.uint8)
            L.oracle_result_digits(h, self.digits.ctypes.data_as(U8P))
            self.carries = np.zeros((self.d, 8), dtype=np.uint64)
            L.oracle_result_carries(h, _p64(self.carries))
            self.carry = np.zeros(8, dtype=np.uint64)
            L.oracle_result_carry(h, _p64(self.carry))
        polys = []
        for which in range(4):
            lst = []
            for k in range(nf):
                ln = L.oracle_result_poly_len(h, k, which)
                arr = np.zeros((max(ln, 1), 4), dtype=np.uint64)
                L.oracle_result_poly_copy(h, k, which, _p64(arr))
                lst.append(arr[:ln])
            polys.append(lst)
        self.a, self.b, self.ca, self.cb = polys
        L.oracle_result_free(h)


def synth_inputs(curve_id, seed, n):
    """the product's synthetic inputs (k_synth_inputs) restated on the CPU: (n,4) scalars, (n,12) points with z = 1"""
    S = np.zeros((n, 4), dtype=np.uint64)
    P = np.zeros((n, 12), dtype=np.uint64)
    _chk(lib().oracle_synth_inputs(curve_id, C.c_uint64(seed), C.c_size_t(n), _p64(S), _p64(P)))
    return S, P


def lhs_witness(curve_id, scalars, pts, base, with_functions=True):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    pts = np.ascontiguousarray(pts, dtype=np.uint64)
    h = C.c_void_p()
    _chk(lib().oracle_lhs_witness(curve_id, _p64(scalars), _p64(pts), C.c_size_t(len(pts)), C.c_uint8(base),
                                  int(with_functions), C.byref(h)))
    return Result(h, len(pts))


def divisor_witness(curve_id, pts, partial=False):
    pts = np.ascontiguousarray(pts, dtype=np.uint64)
    h = C.c_void_p()
    out_pt = np.zeros(8, dtype=np.uint64)
    _chk(lib().oracle_divisor_witness(curve_id, _p64(pts), C.c_size_t(len(pts)), int(partial), _p64(out_pt), C.byref(h)))
    r = Result(h, len(pts), with_digits=False)
    r.output = out_pt
    return r


def prepare_scalar_witness(sc, base, num_digits, logtable, mode=0):
    """oracle restatement of src/negbase_utils.rs:79-124 -> rows[base][num_limbs+1] like pyref.prepare_scalar_witness"""
    w = np.array([(sc >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    num_limbs = (num_digits + logtable - 1) // logtable
    out = np.zeros((base, num_limbs + 1, 4), dtype=np.uint64)
    _chk(lib().oracle_prepare_scalar_witness(_p64(w), C.c_uint8(base), C.c_size_t(num_digits), C.c_size_t(logtable), int(mode), _p64(out)))
    rows = []
    for i in range(base):
        row = []
        for j in range(num_limbs + 1):
            v = int(out[i, j, 0]) | (int(out[i, j, 1]) << 64)
            mask, kind = int(out[i, j, 2]) & 0xFFFFFFFF, int(out[i, j, 2]) >> 32
            sv = v - (1 << 128) if v >> 127 else v
            row.append(("scalar", v) if kind == 0 else (("bucket", sv) if kind == 1 else ("limb", sv, mask)))
        rows.append(row)
    return rows


def divisor_witness_naive(curve_id, pts):
    """oracle restatement of src/regular_functions_utils.rs:483-551 -> (pos, neg) arrays (k, 3, 4): lx | ly | lz"""
    pts = np.ascontiguousarray(pts, dtype=np.uint64)
    n = len(pts)
    pos, neg = np.zeros((max(n, 1), 3, 4), dtype=np.uint64), np.zeros((max(n, 1), 3, 4), dtype=np.uint64)
    npos, nneg = C.c_size_t(len(pos)), C.c_size_t(len(neg))
    _chk(lib().oracle_divisor_witness_naive(curve_id, _p64(pts), C.c_size_t(n), _p64(pos), C.byref(npos), _p64(neg), C.byref(nneg)))
    return pos[: npos.value], neg[: nneg.value]
