"""B200-native Liam-Eagen MSM witness engine -- Python host mirror over the C ABI (ctypes).

The product is `libeagen_msm.so` (hand-written sm_100a kernels behind include/eagen_msm.h).  This module is
the Python-side mirror of the reference crate's public functions for the path, used by tests/ and bench.py:

    compute_lhs_witness            reference: src/argument_witness_calc.rs:87-136
    negbase_decompose (batched)    reference: src/negbase_utils.rs:20-36
    precompute_multiplicities      reference: src/argument_witness_calc.rs:43-51
    compute_divisor_witness(_partial)  reference: src/regular_functions_utils.rs:453-480
    poly_mul / ntt / omega_pow ... reference: src/regular_functions_utils.rs:17-24,102-129,209-216

There is NO CPU fallback: if the shared library is missing, or no CUDA device is visible, calls raise.
Arrays are numpy uint64 in the ABI layout (Montgomery limbs; see include/eagen_msm.h).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeagen_msm.so")

PALLAS, VESTA, GRUMPKIN = 0, 1, 2
CURVE_IDS = {"pallas": PALLAS, "vesta": VESTA, "grumpkin": GRUMPKIN}

CANONICAL, RAW_TREE, PARTIAL, NO_FUNCTIONS, KEEP_DIGITS = 0, 1, 2, 4, 8
POLY_A, POLY_B = 0, 1

OK = 0
E_ARG, E_LEN, E_RANGE, E_SUM_NONZERO, E_NTT_TOO_LARGE, E_CUDA, E_NCCL, E_DIGITS, E_DOMAIN, E_NO_DEVICE, E_EMPTY = range(-1, -12, -1)

U64P = C.POINTER(C.c_uint64)
U8P = C.POINTER(C.c_uint8)

PSW_FAITHFUL, PSW_INTENDED = 0, 1
# one Entry of prepare_scalar_witness (32 bytes)
PSW_DTYPE = np.dtype([("lo", "<u8"), ("hi", "<u8"), ("mask", "<u4"), ("kind", "<u4"), ("zero", "<u8")])

# every symbol include/eagen_msm.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "eagen_ctx_create", "eagen_ctx_destroy", "eagen_last_error", "eagen_status_string", "eagen_launch_count", "eagen_fallback_count",
    "eagen_num_digits", "eagen_negbase_decompose", "eagen_precompute_multiplicities", "eagen_lhs_witness",
    "eagen_divisor_witness", "eagen_result_num_digits", "eagen_result_num_functions", "eagen_result_poly_len",
    "eagen_result_poly_copy", "eagen_result_carry", "eagen_result_carries", "eagen_result_digits",
    "eagen_result_copy_all", "eagen_result_total_bytes", "eagen_result_device_ms", "eagen_result_free",
    "eagen_poly_mul", "eagen_ntt", "eagen_fft_precomp", "eagen_batch_invert", "eagen_eval_function",
    "eagen_dev_shard_sums", "eagen_dev_carry_chain", "eagen_dev_trees", "eagen_dev_lhs_witness",
    "eagen_result_device_view", "eagen_synth_inputs", "eagen_dev_synth_inputs",
    "eagen_set_profiling", "eagen_profile_reset", "eagen_profile_json", "eagen_microbench",
    "eagen_dev_negbase", "eagen_dev_ntt", "eagen_lhs_witness_stream", "eagen_lhs_witness_stream_layout",
    "eagen_table_entry_by_id", "eagen_msm", "eagen_prepare_scalar_witness", "eagen_divisor_witness_naive",
    "eagen_circuit_sizes", "eagen_result_copy_padded", "eagen_result_eval", "eagen_to_curve_x", "eagen_y_from_x", "eagen_slope",
    "eagen_ctx_set_stream_split", "eagen_comm_unique_id", "eagen_comm_init", "eagen_comm_init_all", "eagen_comm_destroy",
    "eagen_comm_size", "eagen_comm_rank", "eagen_position_range", "eagen_lhs_witness_sharded_layout", "eagen_lhs_witness_sharded",
    "eagen_dev_lhs_witness_sharded", "eagen_result_first_function",
]
COMM_ID_BYTES = 128
SELFTEST_SYMBOLS = ["eagen_selftest_field", "eagen_selftest_curve", "eagen_selftest_negbase_params", "eagen_selftest_negbase_digits",
                    "eagen_selftest_ntt_plan"]


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it and ships it to the other ranks through any side channel)"""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib().eagen_comm_unique_id(buf)
    if rc != 0:
        raise EagenError(rc, lib().eagen_last_error(None).decode())
    return buf.raw


def comm_init_all(ctxs):
    """single-process multi-GPU: one NCCL communicator over the contexts' devices (the sharded call then needs one host thread per context)"""
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    rc = lib().eagen_comm_init_all(arr, len(ctxs))
    if rc != 0:
        raise EagenError(rc, lib().eagen_last_error(ctxs[0]._h).decode())


def position_range(rank, nranks, d):
    b, e = C.c_uint32(), C.c_uint32()
    rc = lib().eagen_position_range(rank, nranks, d, C.byref(b), C.byref(e))
    if rc != 0:
        raise EagenError(rc, "eagen_position_range: bad arguments")
    return b.value, e.value


class EagenError(RuntimeError):
    """A non-zero status from the C ABI (the reference panics in the same situations)."""

    def __init__(self, status, message):
        super().__init__("eagen status %d: %s" % (status, message))
        self.status = status


_lib = None


def lib():
    """Load libeagen_msm.so; fail loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libeagen_msm.so is missing: run `python halo2-liam-eagen-msm_b200/build.py` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.eagen_last_error.restype = C.c_char_p
        L.eagen_last_error.argtypes = [C.c_void_p]
        L.eagen_status_string.restype = C.c_char_p
        L.eagen_launch_count.restype = C.c_uint64
        L.eagen_launch_count.argtypes = [C.c_void_p]
        L.eagen_fallback_count.restype = C.c_uint64
        L.eagen_fallback_count.argtypes = [C.c_void_p]
        L.eagen_ctx_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.eagen_ctx_destroy.argtypes = [C.c_void_p]
        L.eagen_num_digits.argtypes = [C.c_int, C.c_uint8, C.POINTER(C.c_uint32)]
        L.eagen_negbase_decompose.argtypes = [C.c_void_p, U64P, C.c_size_t, C.c_uint8, U8P]
        L.eagen_precompute_multiplicities.argtypes = [C.c_void_p, U64P, C.c_size_t, C.c_uint8, U64P]
        L.eagen_lhs_witness.argtypes = [C.c_void_p, U64P, U64P, C.c_size_t, C.c_uint8, C.c_uint32, C.POINTER(C.c_void_p)]
        L.eagen_lhs_witness_stream_layout.argtypes = [C.c_int, C.c_size_t, C.c_uint8, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.eagen_lhs_witness_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.c_void_p, C.c_size_t,
                                               C.POINTER(C.c_void_p)]
        L.eagen_msm.argtypes = [C.c_void_p, U64P, U64P, C.c_size_t, U64P, C.POINTER(C.c_double)]
        L.eagen_dev_lhs_witness.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.POINTER(C.c_void_p)]
        L.eagen_divisor_witness.argtypes = [C.c_void_p, U64P, C.c_size_t, C.c_uint32, U64P, C.POINTER(C.c_void_p)]
        L.eagen_result_num_digits.restype = C.c_uint32
        L.eagen_result_num_digits.argtypes = [C.c_void_p]
        L.eagen_result_num_functions.restype = C.c_size_t
        L.eagen_result_num_functions.argtypes = [C.c_void_p]
        L.eagen_result_poly_len.restype = C.c_size_t
        L.eagen_result_poly_len.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        L.eagen_result_poly_copy.argtypes = [C.c_void_p, C.c_size_t, C.c_int, U64P]
        L.eagen_result_carry.argtypes = [C.c_void_p, U64P]
        L.eagen_result_carries.argtypes = [C.c_void_p, U64P]
        L.eagen_result_digits.argtypes = [C.c_void_p, U8P]
        L.eagen_result_total_bytes.restype = C.c_size_t
        L.eagen_result_total_bytes.argtypes = [C.c_void_p]
        L.eagen_result_copy_all.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.eagen_result_device_ms.restype = C.c_double
        L.eagen_result_device_ms.argtypes = [C.c_void_p]
        L.eagen_result_free.argtypes = [C.c_void_p]
        L.eagen_poly_mul.argtypes = [C.c_void_p, U64P, C.c_size_t, U64P, C.c_size_t, U64P]
        L.eagen_ntt.argtypes = [C.c_void_p, U64P, C.c_uint32, C.c_int]
        L.eagen_fft_precomp.argtypes = [C.c_int, C.c_int, C.c_uint64, U64P]
        L.eagen_batch_invert.argtypes = [C.c_void_p, U64P, C.c_size_t]
        L.eagen_eval_function.argtypes = [C.c_void_p, U64P, C.c_size_t, U64P, C.c_size_t, U64P, C.c_size_t, U64P]
        L.eagen_dev_shard_sums.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint8, C.c_void_p, C.c_void_p, C.c_void_p]
        L.eagen_dev_negbase.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint8, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.eagen_dev_ntt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        L.eagen_dev_carry_chain.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint8, C.c_void_p]
        L.eagen_dev_trees.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.c_uint32,
                                      C.c_uint32, C.POINTER(C.c_void_p)]
        L.eagen_result_device_view.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.eagen_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.eagen_microbench.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.eagen_profile_reset.argtypes = [C.c_void_p]
        L.eagen_profile_json.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.eagen_synth_inputs.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t, U64P, U64P]
        L.eagen_dev_synth_inputs.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p]
        L.eagen_prepare_scalar_witness.argtypes = [C.c_void_p, U64P, C.c_size_t, C.c_uint8, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t]
        L.eagen_divisor_witness_naive.argtypes = [C.c_void_p, U64P, C.c_size_t, U64P, C.POINTER(C.c_size_t), U64P, C.POINTER(C.c_size_t)]
        L.eagen_circuit_sizes.argtypes = [C.c_size_t, C.c_uint8, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.eagen_result_copy_padded.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, U64P, U64P]
        L.eagen_result_eval.argtypes = [C.c_void_p, C.c_void_p, U64P, C.c_size_t, U64P]
        L.eagen_to_curve_x.argtypes = [C.c_int, U64P, U64P]
        L.eagen_y_from_x.argtypes = [C.c_int, U64P, U64P, C.POINTER(C.c_int)]
        L.eagen_slope.argtypes = [C.c_int, U64P, U64P]
        L.eagen_ctx_set_stream_split.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_int]
        L.eagen_comm_unique_id.argtypes = [C.c_void_p]
        L.eagen_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.eagen_comm_init_all.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.eagen_comm_destroy.argtypes = [C.c_void_p]
        L.eagen_comm_size.argtypes = [C.c_void_p]
        L.eagen_comm_rank.argtypes = [C.c_void_p]
        L.eagen_position_range.argtypes = [C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.eagen_lhs_witness_sharded_layout.argtypes = [C.c_int, C.c_size_t, C.c_uint8, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                                       C.POINTER(C.c_size_t)]
        L.eagen_lhs_witness_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.c_void_p, C.c_size_t,
                                                C.POINTER(C.c_void_p)]
        L.eagen_dev_lhs_witness_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.POINTER(C.c_void_p)]
        L.eagen_result_first_function.restype = C.c_size_t
        L.eagen_result_first_function.argtypes = [C.c_void_p]
        L.eagen_selftest_field.argtypes = [C.c_int, C.c_int, U64P, U64P, U64P]
        L.eagen_selftest_curve.argtypes = [C.c_int, C.c_int, U64P, U64P, C.c_uint32, U64P]
        L.eagen_selftest_negbase_params.argtypes = [C.c_int, C.c_uint8, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.eagen_selftest_negbase_digits.argtypes = [C.c_int, C.c_uint8, U64P, U8P, C.POINTER(C.c_int)]
        L.eagen_selftest_ntt_plan.argtypes = [C.c_int, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _p64(a):
    return a.ctypes.data_as(U64P)


def _arr(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a if cols is None else a.reshape(-1, cols)


def num_digits(curve, base):
    """d = logb_ceil(isqrt(order)+2, base) + 1   (reference: src/argument_witness_calc.rs:89-91)"""
    d = C.c_uint32()
    rc = lib().eagen_num_digits(curve, C.c_uint8(base), C.byref(d))
    if rc:
        raise EagenError(rc, lib().eagen_status_string(rc).decode())
    return d.value


def fft_precomp(curve, which, exp):
    out = np.zeros(4, dtype=np.uint64)
    rc = lib().eagen_fft_precomp(curve, which, C.c_uint64(exp), _p64(out))
    if rc:
        raise EagenError(rc, lib().eagen_status_string(rc).decode())
    return out


def table_entry_by_id(curve, base, idx):
    """reference: src/negbase_utils.rs:58-77 over the curve's base field"""
    out = np.zeros(4, dtype=np.uint64)
    lib().eagen_table_entry_by_id.argtypes = [C.c_int, C.c_uint8, C.c_size_t, U64P]
    rc = lib().eagen_table_entry_by_id(curve, C.c_uint8(base), idx, _p64(out))
    if rc:
        raise EagenError(rc, lib().eagen_status_string(rc).decode())
    return out


def circuit_sizes(num_pts, base):
    """(a_size, b_size) of the circuit's coefficient columns (reference: src/config.rs:641-642)"""
    a, b = C.c_size_t(), C.c_size_t()
    rc = lib().eagen_circuit_sizes(num_pts, C.c_uint8(base), C.byref(a), C.byref(b))
    if rc:
        raise EagenError(rc, lib().eagen_status_string(rc).decode())
    return a.value, b.value


def _challenge(fn, name, curve, arr, with_flag=False):
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1)
    out = np.zeros(4, dtype=np.uint64)
    flag = C.c_int(-1)
    rc = fn(curve, _p64(a), _p64(out), C.byref(flag)) if with_flag else fn(curve, _p64(a), _p64(out))
    if rc:
        raise EagenError(rc, name + ": " + lib().eagen_last_error(None).decode())
    return (out, flag.value) if with_flag else out


def to_curve_x(curve, c):
    """reference: src/config.rs:165-176 (EagenError E_DOMAIN where the reference's loop would never end)"""
    return _challenge(lib().eagen_to_curve_x, "to_curve_x", curve, c)


def y_from_x(curve, x):
    """reference: src/config.rs:178-183 -> (y, is_square)"""
    return _challenge(lib().eagen_y_from_x, "y_from_x", curve, x, True)


def slope(curve, x, y):
    """reference: src/config.rs:185-188"""
    return _challenge(lib().eagen_slope, "slope", curve, np.concatenate([np.asarray(x, np.uint64).reshape(4), np.asarray(y, np.uint64).reshape(4)]))


def omega_pow(curve, exp2):
    return fft_precomp(curve, 0, exp2)


def omega_pow_inv(curve, exp2):
    return fft_precomp(curve, 1, exp2)


def half_pow(curve, exp):
    return fft_precomp(curve, 2, exp)


class RegularFunction:
    """a(x) + y*b(x): coefficient arrays (len, 4) uint64 Montgomery, low degree first
    (reference: src/regular_functions_utils.rs:220-225)"""

    def __init__(self, a, b):
        self.a, self.b = a, b


class WitnessResult:
    """Owner of an eagen_result handle."""

    def __init__(self, ctx, handle, n):
        self._ctx, self._h, self.n = ctx, handle, n
        L = lib()
        self.d = L.eagen_result_num_digits(handle)
        self.num_functions = L.eagen_result_num_functions(handle)
        self.device_ms = L.eagen_result_device_ms(handle)
        self.first_function = L.eagen_result_first_function(handle)   # digit position of slot 0 (non-zero for a rank's share)

    def poly(self, k, which):
        L = lib()
        ln = L.eagen_result_poly_len(self._h, k, which)
        out = np.zeros((max(ln, 1), 4), dtype=np.uint64)
        self._ctx._chk(L.eagen_result_poly_copy(self._h, k, which, _p64(out)))
        return out[:ln]

    def function(self, k):
        return RegularFunction(self.poly(k, POLY_A), self.poly(k, POLY_B))

    def functions(self):
        return [self.function(k) for k in range(self.num_functions)]

    @property
    def carry(self):
        out = np.zeros(8, dtype=np.uint64)
        self._ctx._chk(lib().eagen_result_carry(self._h, _p64(out)))
        return out

    @property
    def carries(self):
        out = np.zeros((self.d, 8), dtype=np.uint64)
        self._ctx._chk(lib().eagen_result_carries(self._h, _p64(out)))
        return out

    @property
    def digits(self):
        out = np.zeros((self.n, self.d), dtype=np.uint8)
        self._ctx._chk(lib().eagen_result_digits(self._h, out.ctypes.data_as(U8P)))
        return out

    def total_bytes(self):
        return lib().eagen_result_total_bytes(self._h)

    def padded(self, num_pts, base):
        """(a, b) as the circuit's fixed-size rows: (nf, a_size, 4) and (nf, b_size, 4), zero padded (src/config.rs:641-642)"""
        a_size, b_size = circuit_sizes(num_pts, base)
        a = np.zeros((self.num_functions, a_size, 4), dtype=np.uint64)
        b = np.zeros((self.num_functions, b_size, 4), dtype=np.uint64)
        self._ctx._chk(lib().eagen_result_copy_padded(self._h, num_pts, C.c_uint8(base), _p64(a), _p64(b)))
        return a, b

    def ev(self, pts):
        """RegularFunction::ev of every function at the Jacobian points pts -> (nf, m, 4); coefficients stay on the device"""
        p = _arr(pts, 12)
        out = np.zeros((self.num_functions, len(p), 4), dtype=np.uint64)
        self._ctx._chk(lib().eagen_result_eval(self._ctx._h, self._h, _p64(p), len(p), _p64(out)))
        return out

    def copy_all_into(self, host_ptr, nbytes):
        w = C.c_size_t()
        self._ctx._chk(lib().eagen_result_copy_all(self._h, C.c_void_p(host_ptr), C.c_size_t(nbytes), C.byref(w)))
        return w.value

    def device_view(self):
        da, db, sa, sb = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._ctx._chk(lib().eagen_result_device_view(self._h, C.byref(da), C.byref(sa), C.byref(db), C.byref(sb)))
        return da.value, sa.value, db.value, sb.value

    def free(self):
        if self._h is not None:
            lib().eagen_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One eagen_ctx: one curve on one CUDA device."""

    def __init__(self, curve="pallas", device=0):
        self.curve = CURVE_IDS[curve] if isinstance(curve, str) else int(curve)
        h = C.c_void_p()
        rc = lib().eagen_ctx_create(self.curve, int(device), C.byref(h))
        if rc:
            raise EagenError(rc, lib().eagen_last_error(None).decode() or lib().eagen_status_string(rc).decode())
        self._h = h

    def _chk(self, rc):
        if rc:
            raise EagenError(rc, lib().eagen_last_error(self._h).decode() or lib().eagen_status_string(rc).decode())

    def close(self):
        if self._h is not None:
            lib().eagen_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self):
        return lib().eagen_launch_count(self._h)

    def fallback_count(self):
        """groups of divisor trees rebuilt on an isomorphic curve after a domain collision (x of an output point on the domain)"""
        return lib().eagen_fallback_count(self._h)

    def microbench(self, which):
        v = C.c_double()
        self._chk(lib().eagen_microbench(self._h, which, C.byref(v)))
        return v.value

    # ---- multi-GPU (one context per rank; see eagen_lhs_witness_sharded in include/eagen_msm.h) ----------------------
    def comm_init(self, nranks, rank, unique_id):
        """join the NCCL communicator described by the 128-byte id every rank received from rank 0 (comm_unique_id())"""
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._chk(lib().eagen_comm_init(self._h, nranks, rank, buf))

    def comm_destroy(self):
        self._chk(lib().eagen_comm_destroy(self._h))

    def comm_size(self):
        return lib().eagen_comm_size(self._h)

    def comm_rank(self):
        return lib().eagen_comm_rank(self._h)

    def set_stream_split(self, percents):
        arr = (C.c_uint32 * len(percents))(*percents)
        self._chk(lib().eagen_ctx_set_stream_split(self._h, arr, len(percents)))

    def sharded_layout(self, n_total, base, rank, nranks):
        a, b, t = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._chk(lib().eagen_lhs_witness_sharded_layout(self.curve, n_total, C.c_uint8(base), rank, nranks, C.byref(a), C.byref(b), C.byref(t)))
        return a.value, b.value, t.value

    def lhs_witness_sharded_ptr(self, scalars_ptr, pts_ptr, n_local, base, flags=CANONICAL, device=False, out_ptr=None, out_bytes=0):
        """this rank's share of compute_lhs_witness over all ranks' points (raw pointers: host, or device with device=True);
        out_ptr: host buffer the rank's functions are streamed into (host-input form only)"""
        h = C.c_void_p()
        if device:
            self._chk(lib().eagen_dev_lhs_witness_sharded(self._h, C.c_void_p(scalars_ptr), C.c_void_p(pts_ptr), n_local, C.c_uint8(base), flags, C.byref(h)))
        else:
            self._chk(lib().eagen_lhs_witness_sharded(self._h, C.c_void_p(scalars_ptr), C.c_void_p(pts_ptr), n_local, C.c_uint8(base), flags,
                                                      C.c_void_p(out_ptr) if out_ptr else None, out_bytes, C.byref(h)))
        return WitnessResult(self, h, n_local * max(self.comm_size(), 1))

    def lhs_witness_sharded(self, scalars, pts, base, flags=CANONICAL):
        s, p = _arr(scalars, 4), _arr(pts, 12)
        if len(s) != len(p):
            raise EagenError(E_LEN, "incompatible amount of coefficients")
        return self.lhs_witness_sharded_ptr(s.ctypes.data, p.ctypes.data, len(p), base, flags)

    def set_profiling(self, on=True):
        self._chk(lib().eagen_set_profiling(self._h, int(on)))

    def profile_reset(self):
        self._chk(lib().eagen_profile_reset(self._h))

    def profile(self):
        """list of dicts: kernel group, launches, ms (CUDA events), algorithmic bytes and modmul counts"""
        import json
        buf = C.create_string_buffer(1 << 16)
        self._chk(lib().eagen_profile_json(self._h, buf, len(buf)))
        return json.loads(buf.value.decode())

    # ---- the path ----------------------------------------------------------------------------------------
    def negbase_decompose(self, scalars, base):
        """(n,4) Montgomery scalars -> (n,d) uint8 digits, MSD first."""
        s = _arr(scalars, 4)
        d = num_digits(self.curve, base)
        out = np.zeros((len(s), d), dtype=np.uint8)
        self._chk(lib().eagen_negbase_decompose(self._h, _p64(s), len(s), C.c_uint8(base), out.ctypes.data_as(U8P)))
        return out

    def precompute_multiplicities(self, pts, base):
        """(n,12) Jacobian points -> (n, base-1, 8) affine multiples."""
        p = _arr(pts, 12)
        out = np.zeros((len(p), base - 1, 8), dtype=np.uint64)
        self._chk(lib().eagen_precompute_multiplicities(self._h, _p64(p), len(p), C.c_uint8(base), _p64(out)))
        return out

    def compute_lhs_witness(self, scalars, pts, base, flags=CANONICAL):
        s, p = _arr(scalars, 4), _arr(pts, 12)
        if len(s) != len(p):
            raise EagenError(E_LEN, "incompatible amount of coefficients")  # reference: src/argument_witness_calc.rs:88
        h = C.c_void_p()
        self._chk(lib().eagen_lhs_witness(self._h, _p64(s), _p64(p), len(p), C.c_uint8(base), flags, C.byref(h)))
        return WitnessResult(self, h, len(p))

    def compute_lhs_witness_ptr(self, scalars_ptr, pts_ptr, n, base, flags=CANONICAL, device=False):
        """raw-pointer variant: host pointers (pinned or not) or, with device=True, CUDA device pointers"""
        h = C.c_void_p()
        if device:
            self._chk(lib().eagen_dev_lhs_witness(self._h, C.c_void_p(scalars_ptr), C.c_void_p(pts_ptr), n, C.c_uint8(base), flags, C.byref(h)))
        else:
            self._chk(lib().eagen_lhs_witness(self._h, C.cast(C.c_void_p(scalars_ptr), U64P), C.cast(C.c_void_p(pts_ptr), U64P), n,
                                              C.c_uint8(base), flags, C.byref(h)))
        return WitnessResult(self, h, n)

    def stream_layout(self, n, base):
        """(a_stride, b_stride, total_bytes) of the streamed result layout"""
        a, b, t = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._chk(lib().eagen_lhs_witness_stream_layout(self.curve, n, C.c_uint8(base), C.byref(a), C.byref(b), C.byref(t)))
        return a.value, b.value, t.value

    def compute_lhs_witness_stream(self, scalars_ptr, pts_ptr, n, base, out_ptr, out_bytes, flags=CANONICAL):
        """host pointers in, functions streamed into the host buffer at out_ptr (see eagen_lhs_witness_stream)"""
        h = C.c_void_p()
        self._chk(lib().eagen_lhs_witness_stream(self._h, C.c_void_p(scalars_ptr), C.c_void_p(pts_ptr), n, C.c_uint8(base), flags,
                                                 C.c_void_p(out_ptr), out_bytes, C.byref(h)))
        return WitnessResult(self, h, n)

    def compute_divisor_witness_partial(self, pts, flags=CANONICAL):
        p = _arr(pts, 12)
        h = C.c_void_p()
        out_pt = np.zeros(8, dtype=np.uint64)
        self._chk(lib().eagen_divisor_witness(self._h, _p64(p), len(p), flags | PARTIAL, _p64(out_pt), C.byref(h)))
        r = WitnessResult(self, h, len(p))
        return r.function(0), out_pt

    def compute_divisor_witness(self, pts, flags=CANONICAL):
        p = _arr(pts, 12)
        h = C.c_void_p()
        self._chk(lib().eagen_divisor_witness(self._h, _p64(p), len(p), flags & ~PARTIAL, None, C.byref(h)))
        return WitnessResult(self, h, len(p)).function(0)

    def best_multiexp(self, scalars, pts, with_time=False):
        """sum s_j P_j for full-width scalars: (8,) affine point (reference tests' cross-check, halo2 best_multiexp)"""
        s, p = _arr(scalars, 4), _arr(pts, 12)
        if len(s) != len(p):
            raise EagenError(E_LEN, "incompatible amount of coefficients")
        out, ms = np.zeros(8, dtype=np.uint64), C.c_double()
        self._chk(lib().eagen_msm(self._h, _p64(s), _p64(p), len(p), _p64(out), C.byref(ms)))
        return (out, ms.value) if with_time else out

    # ---- helpers ---------------------------------------------------------------------------------------------
    def poly_mul(self, a, b):
        a, b = _arr(a, 4), _arr(b, 4)
        n = len(a) + len(b) - 1 if len(a) + len(b) else 0
        out = np.zeros((max(n, 1), 4), dtype=np.uint64)
        self._chk(lib().eagen_poly_mul(self._h, _p64(a), len(a), _p64(b), len(b), _p64(out)))
        return out[:n]

    def ntt(self, data, inverse=False):
        a = _arr(data, 4).copy()
        log_n = (len(a) - 1).bit_length()
        if 1 << log_n != len(a):
            raise ValueError("length must be a power of two")
        self._chk(lib().eagen_ntt(self._h, _p64(a), log_n, int(inverse)))
        return a

    def batch_invert(self, elems):
        a = _arr(elems, 4).copy()
        self._chk(lib().eagen_batch_invert(self._h, _p64(a), len(a)))
        return a

    def eval_function(self, f, pts):
        a, b, p = _arr(f.a, 4), _arr(f.b, 4), _arr(pts, 12)
        out = np.zeros((len(p), 4), dtype=np.uint64)
        self._chk(lib().eagen_eval_function(self._h, _p64(a), len(a), _p64(b), len(b), _p64(p), len(p), _p64(out)))
        return out

    def prepare_scalar_witness(self, scalars, base, num_digits, logtable, mode=PSW_FAITHFUL):
        """reference: src/negbase_utils.rs:79-124 for every scalar.  Returns a structured array [n][base][num_limbs+1] with fields
        lo, hi (two's complement i128 halves), mask, kind (0 Scalar, 1 Bucket, 2 Limb)"""
        s = _arr(scalars, 4)
        num_limbs = (num_digits + logtable - 1) // logtable
        out = np.zeros((len(s), base, num_limbs + 1), dtype=PSW_DTYPE)
        self._chk(lib().eagen_prepare_scalar_witness(self._h, _p64(s), len(s), C.c_uint8(base), num_digits, logtable, mode,
                                                     out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def compute_divisor_witness_naive(self, pts):
        """reference: src/regular_functions_utils.rs:483-551.  Returns (pos, neg): arrays (k, 3, 4) of lines lx | ly | lz"""
        p = _arr(pts, 12)
        n = len(p)
        pos, neg = np.zeros((max(n, 1), 3, 4), dtype=np.uint64), np.zeros((max(n, 1), 3, 4), dtype=np.uint64)
        npos, nneg = C.c_size_t(len(pos)), C.c_size_t(len(neg))
        self._chk(lib().eagen_divisor_witness_naive(self._h, _p64(p), n, _p64(pos), C.byref(npos), _p64(neg), C.byref(nneg)))
        return pos[: npos.value], neg[: nneg.value]

    def synth_inputs(self, seed, n):
        """deterministic synthetic (scalars (n,4), Jacobian points (n,12)) generated on the device"""
        sc, pt = np.zeros((n, 4), dtype=np.uint64), np.zeros((n, 12), dtype=np.uint64)
        self._chk(lib().eagen_synth_inputs(self._h, C.c_uint64(seed), n, _p64(sc), _p64(pt)))
        return sc, pt

    def dev_synth_inputs(self, seed, n, d_scalars, d_pts):
        self._chk(lib().eagen_dev_synth_inputs(self._h, C.c_uint64(seed), n, d_scalars, d_pts))

    # ---- device-resident stages (pointers are CUDA device addresses, e.g. torch tensors' data_ptr()) -------------
    def dev_shard_sums(self, d_scalars, d_pts, n, base, d_planes, d_table, d_sums):
        self._chk(lib().eagen_dev_shard_sums(self._h, d_scalars, d_pts, n, C.c_uint8(base), d_planes, d_table, d_sums))

    def dev_negbase(self, d_scalars, n, base, d_planes, d_rows=None):
        ms = C.c_double()
        self._chk(lib().eagen_dev_negbase(self._h, d_scalars, n, C.c_uint8(base), d_planes, d_rows, C.byref(ms)))
        return ms.value

    def dev_ntt(self, d_data, log_n, batch, inverse=False):
        ms = C.c_double()
        self._chk(lib().eagen_dev_ntt(self._h, d_data, log_n, batch, int(inverse), C.byref(ms)))
        return ms.value

    def dev_carry_chain(self, d_sums, nparts, base, d_carries):
        self._chk(lib().eagen_dev_carry_chain(self._h, d_sums, nparts, C.c_uint8(base), d_carries))

    def dev_trees(self, d_planes, d_table, d_carries, n, base, pos_begin, pos_end, flags=CANONICAL):
        h = C.c_void_p()
        self._chk(lib().eagen_dev_trees(self._h, d_planes, d_table, d_carries, n, C.c_uint8(base), pos_begin, pos_end, flags, C.byref(h)))
        return WitnessResult(self, h, n)


# ---- host self-test hooks (same HD arithmetic source as the kernels, run on the CPU) ----------------------------
def selftest_field(field, op, a, b=None):
    a = _arr(a)
    out = np.zeros(4, dtype=np.uint64)
    bb = None if b is None else _arr(b)
    rc = lib().eagen_selftest_field(field, op, _p64(a), None if bb is None else _p64(bb), _p64(out))
    if rc:
        raise EagenError(rc, "selftest_field")
    return out


def selftest_curve(curve, op, p, q=None, k=0):
    p = _arr(p)
    out = np.zeros(8, dtype=np.uint64)
    qq = None if q is None else _arr(q)
    rc = lib().eagen_selftest_curve(curve, op, _p64(p), None if qq is None else _p64(qq), k, _p64(out))
    if rc:
        raise EagenError(rc, "selftest_curve")
    return out


def selftest_negbase_params(curve, base):
    d, group, words = C.c_uint32(), C.c_uint32(), C.c_uint32()
    limbs = (C.c_uint32 * 32)()
    rc = lib().eagen_selftest_negbase_params(curve, C.c_uint8(base), C.byref(d), C.byref(group), C.byref(words), limbs)
    if rc:
        raise EagenError(rc, "selftest_negbase_params")
    to_int = lambda ws: sum(int(w) << (32 * i) for i, w in enumerate(ws))
    return dict(d=d.value, group=group.value, words=words.value, sq=to_int(limbs[0:8]), K=to_int(limbs[8:16]), bd=to_int(limbs[16:24]),
                inv=to_int(limbs[24:32]))


def selftest_negbase_digits(curve, base, scalar):
    """K1's per-scalar arithmetic on the host (same source as the kernel): (digits MSD first, kernel error flag)"""
    s = _arr(scalar)
    d = num_digits(curve, base)
    out = np.zeros(d, dtype=np.uint8)
    kerr = C.c_int()
    rc = lib().eagen_selftest_negbase_digits(curve, C.c_uint8(base), _p64(s), out.ctypes.data_as(U8P), C.byref(kerr))
    if rc:
        raise EagenError(rc, "selftest_negbase_digits")
    return out, kerr.value


def selftest_ntt_plan(t):
    pairs = (C.c_int * 16)()
    n = lib().eagen_selftest_ntt_plan(t, pairs)
    return [(pairs[2 * i], pairs[2 * i + 1]) for i in range(n)]
