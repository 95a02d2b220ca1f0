/* eagen_msm.h -- C ABI of the B200-native Liam-Eagen MSM witness engine (libeagen_msm.so).
 *
 * This is the drop-in boundary for the witness hot path of levs57/halo2-liam-eagen-msm.  The reference has no
 * FFI today: its "operator API" is the set of public Rust functions below; each entry point here is what a thin
 * Rust shim binds to keep those signatures (see INTEGRATION.md and halo2-liam-eagen-msm_b200/rust/).
 *
 * Data layout (identical to the in-memory form of halo2curves / pasta_curves types):
 *   field element : 32 bytes, little endian, MONTGOMERY residue (R = 2^256) = Rust `[u64; 4]`
 *                   (reference reinterprets the same bytes: src/precomputed_fft_data.rs:72)
 *   Jacobian point: x | y | z, 3 field elements (96 bytes); identity has z = 0
 *                   (what CurveExt::jacobian_coordinates() returns, src/regular_functions_utils.rs:229,427)
 *   affine point  : x | y, 2 field elements (64 bytes); the identity is encoded as (0, 0)
 *   digits        : u8, n x d row-major, most significant digit first
 *                   (digits_by_scalar after the reverse, src/argument_witness_calc.rs:99-101)
 *
 * All functions return EAGEN_OK (0) or a negative error code; nothing here ever falls back to a CPU path.
 * A context is bound to one CUDA device and is not thread-safe (use one context per thread).
 *
 * Preconditions every entry point relies on:
 *   - curves with a = 0 only (y^2 = x^3 + b).  The reference is generic in C::a() (subst_y2 = [B, A, 0, 1],
 *     src/regular_functions_utils.rs:270); the three curve ids below all have a = 0 and the kernels assume it.  There is no
 *     way to pass another curve: an unknown curve id is EAGEN_E_ARG.
 *   - stream ordering of the eagen_dev_* entry points: the library launches on its own non-blocking streams.  Device buffers
 *     passed in must be COMPLETE (the producing stream synchronised, or an event the caller has already waited on) before the
 *     call; every call returns only after its own device work has finished, so outputs may be used from any stream afterwards.
 */
#ifndef EAGEN_MSM_H
#define EAGEN_MSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eagen_ctx eagen_ctx;
typedef struct eagen_result eagen_result;

enum eagen_curve {
    EAGEN_CURVE_PALLAS = 0,   /* base Fp, scalars Fq, y^2 = x^3 + 5 */
    EAGEN_CURVE_VESTA = 1,    /* base Fq, scalars Fp, y^2 = x^3 + 5 */
    EAGEN_CURVE_GRUMPKIN = 2  /* base bn256::Fr, y^2 = x^3 - 17: the curve the reference's tests instantiate */
};

enum eagen_status {
    EAGEN_OK = 0,
    EAGEN_E_ARG = -1,           /* null pointer, unknown curve, base < 2 ...                                      */
    EAGEN_E_LEN = -2,           /* "incompatible amount of coefficients"   src/argument_witness_calc.rs:88       */
    EAGEN_E_RANGE = -3,         /* scalar >= isqrt(order)+2                src/argument_witness_calc.rs:97       */
    EAGEN_E_SUM_NONZERO = -4,   /* points do not sum to the identity       src/regular_functions_utils.rs:478    */
    EAGEN_E_NTT_TOO_LARGE = -5, /* F::S < loglength                        src/regular_functions_utils.rs:110    */
    EAGEN_E_CUDA = -6,          /* CUDA runtime error (see eagen_last_error)                                     */
    EAGEN_E_NCCL = -7,          /* NCCL failure, libnccl.so.2 not loadable, or a sharded call without a communicator     */
    EAGEN_E_DIGITS = -8,        /* negbase expansion longer than d digits (the reference truncates silently, :99) */
    EAGEN_E_DOMAIN = -9,        /* an intermediate point's x lies on the power-of-two evaluation domain even after the
                                   trees were rebuilt on four isomorphic curves (see eagen_fallback_count)          */
    EAGEN_E_NO_DEVICE = -10,    /* no CUDA device: there is deliberately no CPU fallback                         */
    EAGEN_E_EMPTY = -11         /* group_merge of an empty list            src/regular_functions_utils.rs:382    */
};

/* flags for eagen_lhs_witness / eagen_divisor_witness */
enum eagen_flags {
    EAGEN_CANONICAL = 0,        /* default: trailing zeros trimmed, function monic in its highest-pole-order term */
    EAGEN_RAW_TREE = 1,         /* reference tree order, every line built from z = 1 points, no final scaling;
                                   trailing zero coefficients trimmed                                            */
    EAGEN_PARTIAL = 2,          /* compute_divisor_witness_partial: do not require the points to sum to zero     */
    EAGEN_NO_FUNCTIONS = 4,     /* digits + carries only (skips the divisor witnesses)                           */
    EAGEN_KEEP_DIGITS = 8       /* also materialise the n x d digit matrix in the result                         */
};

/* which polynomial of a function a(x) + y b(x) */
enum eagen_which { EAGEN_POLY_A = 0, EAGEN_POLY_B = 1 };

/* ---- context ------------------------------------------------------------------------------------------ */
int eagen_ctx_create(int curve, int device, eagen_ctx** out);
void eagen_ctx_destroy(eagen_ctx* ctx);
/* streamed output (eagen_lhs_witness_stream / _sharded): per cent of the digit positions per group, e.g. {70, 30} (the default) or
 * {60, 30, 10} on a slow host link; every group's copy hides behind the next group's kernels.  n = 0 restores the default. */
int eagen_ctx_set_stream_split(eagen_ctx* ctx, const uint32_t* percent, int n);
const char* eagen_last_error(const eagen_ctx* ctx);   /* message of the last failing call (never NULL)        */
const char* eagen_status_string(int status);
/* number of kernels this context has launched so far (bench.py reports the per-step delta as gpu_launches) */
uint64_t eagen_launch_count(const eagen_ctx* ctx);
/* how many times this context rebuilt a group of divisor trees on an isomorphic curve because an output point's x-coordinate
 * lay on the evaluation domain; both the canonical and the raw form come back exactly as a direct computation would give them */
uint64_t eagen_fallback_count(const eagen_ctx* ctx);

/* per-kernel-group profiling (CUDA events on the launching stream + exact byte / modmul counts from the launch
 * parameters).  eagen_profile_json writes a JSON array [{"kernel", "launches", "scopes", "ms", "bytes", "modmul"}, ...].
 * on = 0 off, 1 one entry per kernel group, 2 additionally split by tree level ("group@L07": the merge that builds level 8). */
int eagen_set_profiling(eagen_ctx* ctx, int on);
/* integer-pipe roofline denominators measured on this device: which = 0 -> dependent-free 32-bit IMAD per second,
 * which = 1 -> base-field Montgomery products per second in a register-resident loop (ceiling of the field code),
 * which = 2 -> IMAD.WIDE (32x32+64 multiply-accumulate) per second: divided by the 87 IMAD.WIDE of one product it is the multiplier-pipe
 * ceiling every field kernel is measured against */
int eagen_microbench(eagen_ctx* ctx, int which, double* ops_per_second);
int eagen_profile_reset(eagen_ctx* ctx);
int eagen_profile_json(eagen_ctx* ctx, char* buf, size_t cap);

/* ---- host-side scalars of the path ---------------------------------------------------------------------
 * order / isqrt / logb_ceil: d = logb_ceil(isqrt(order)+2, base) + 1      src/argument_witness_calc.rs:32-40,54-56,89-91 */
int eagen_num_digits(int curve, uint8_t base, uint32_t* d);

/* table_entry_by_id over the curve's BASE field (host side, no device needed)      src/negbase_utils.rs:58-77 */
int eagen_table_entry_by_id(int curve, uint8_t base, size_t id, uint64_t* out);

/* ---- the hot path, host buffers in / host buffers out ------------------------------------------------- */

/* negbase_decompose + pad + reverse for n scalars            src/negbase_utils.rs:20-36, argument_witness_calc.rs:93-101
 * scalars: n x 32 B Montgomery elements of the SCALAR field; digits: n x d bytes, MSD first.                */
int eagen_negbase_decompose(eagen_ctx* ctx, const uint64_t* scalars, size_t n, uint8_t base, uint8_t* digits);

/* precompute_multiplicities for n points                      src/argument_witness_calc.rs:43-51,103
 * pts: n Jacobian points; out: n x (base-1) affine points, out[j][k-1] = k * P_j.                          */
int eagen_precompute_multiplicities(eagen_ctx* ctx, const uint64_t* pts, size_t n, uint8_t base, uint64_t* out);

/* compute_lhs_witness                                         src/argument_witness_calc.rs:87-136
 * Returns a result handle holding the carry (= sum s_j P_j, affine), the d per-iteration carries and d functions;
 * function k belongs to digit position k (the reference's ret after the reverse, :132).                    */
int eagen_lhs_witness(eagen_ctx* ctx, const uint64_t* scalars, const uint64_t* pts, size_t n, uint8_t base,
                      uint32_t flags, eagen_result** out);

/* compute_lhs_witness with the functions STREAMED into a caller buffer (pinned host memory recommended) while later digit
 * positions are still being computed, so the device-to-host copy of the result (1.5 GB at 2^20) overlaps the kernels.
 * Layout: function k lives in slot k of (a_stride + b_stride) elements: a_k at out + k*(a_stride+b_stride)*32, b_k right after
 * the a_stride elements; valid lengths come from eagen_result_poly_len.  eagen_lhs_witness_stream_layout gives the strides
 * and the buffer size for (n, base) before the call.                                                                */
int eagen_lhs_witness_stream_layout(int curve, size_t n, uint8_t base, size_t* a_stride, size_t* b_stride, size_t* total_bytes);
int eagen_lhs_witness_stream(eagen_ctx* ctx, const uint64_t* scalars, const uint64_t* pts, size_t n, uint8_t base,
                             uint32_t flags, void* out, size_t out_bytes, eagen_result** res);

/* compute_divisor_witness / compute_divisor_witness_partial   src/regular_functions_utils.rs:453-480
 * pts: n Jacobian points.  out_point (may be NULL): affine output point (-sum), 64 bytes.                   */
int eagen_divisor_witness(eagen_ctx* ctx, const uint64_t* pts, size_t n, uint32_t flags, uint64_t* out_point,
                          eagen_result** out);

/* ---- result handle ------------------------------------------------------------------------------------ */
uint32_t eagen_result_num_digits(const eagen_result* r);                    /* d                              */
size_t eagen_result_num_functions(const eagen_result* r);
size_t eagen_result_poly_len(const eagen_result* r, size_t k, int which);    /* coefficients in a_k or b_k     */
int eagen_result_poly_copy(eagen_result* r, size_t k, int which, uint64_t* out);  /* len x 4 u64, low degree first */
int eagen_result_carry(eagen_result* r, uint64_t* out_affine);              /* 64 bytes                       */
int eagen_result_carries(eagen_result* r, uint64_t* out_affine);            /* d x 64 bytes, iteration order  */
int eagen_result_digits(eagen_result* r, uint8_t* out);                     /* n x d (needs EAGEN_KEEP_DIGITS) */
/* copy every function into one caller buffer: for k = 0..nf-1: a_k then b_k, tightly packed; returns bytes */
int eagen_result_copy_all(eagen_result* r, uint64_t* out, size_t out_bytes, size_t* written);
size_t eagen_result_total_bytes(const eagen_result* r);
/* device time of the compute part of the call that produced this result (ms, CUDA events) */
double eagen_result_device_ms(const eagen_result* r);
void eagen_result_free(eagen_result* r);

/* best_multiexp(coeffs, bases): sum s_j P_j for FULL-WIDTH scalars (windowed bucket method on the device).  The reference only
 * uses it in its tests, as the cross-check of the witness carry        src/argument_witness_calc.rs:144, regular_functions_utils.rs:655
 * scalars: n x 32 B Montgomery (scalar field); pts: n Jacobian points; out: affine (64 B).  device_ms may be NULL. */
int eagen_msm(eagen_ctx* ctx, const uint64_t* scalars, const uint64_t* pts, size_t n, uint64_t* out_affine, double* device_ms);

/* prepare_scalar_witness for n scalars     src/negbase_utils.rs:39-43,79-124
 * scalars: n x 32 B Montgomery (scalar field).  out: n x base x (num_limbs + 1) entries, num_limbs = ceil(num_digits / logtable),
 * row-major [scalar][row][slot]; one entry = 32 bytes:
 *     u64 value_lo | u64 value_hi (two's complement i128; Entry::Scalar: the canonical scalar) | u32 bitmask | u32 kind | 8 zero bytes
 *     kind: 0 = Entry::Scalar (row 0, slot 0), 1 = Entry::Bucket (row r >= 1, slot 0), 2 = Entry::Limb (slot >= 1)
 * mode EAGEN_PSW_FAITHFUL follows the reference as written (limb slot i % logtable + 1, :98-101; EAGEN_E_ARG where the reference
 * indexes out of bounds), EAGEN_PSW_INTENDED uses slot i / logtable + 1.  i128 sums wrap (release-build semantics of the reference).
 * Errors: EAGEN_E_RANGE (scalar >= isqrt(order)+2), EAGEN_E_DIGITS (more than num_digits digits, the assert at :81). */
enum eagen_psw_mode { EAGEN_PSW_FAITHFUL = 0, EAGEN_PSW_INTENDED = 1 };
int eagen_prepare_scalar_witness(eagen_ctx* ctx, const uint64_t* scalars, size_t n, uint8_t base, uint32_t num_digits,
                                 uint32_t logtable, int mode, void* out, size_t out_bytes);

/* compute_divisor_witness_naive: the witness as an Arrangement of numerator (pos) and denominator (neg) lines
 *                                 src/regular_functions_utils.rs:483-551
 * pts: n Jacobian points summing to the identity (EAGEN_E_SUM_NONZERO otherwise).  A line is lx | ly | lz (96 bytes):
 * RegularFunction::from_line(lx, ly, lz), i.e. a = [lz, lx], b = [ly]; lines appear in the reference's push order.
 * *n_pos / *n_neg: capacity of the buffers in lines on entry (n each always suffices), number of lines on return. */
int eagen_divisor_witness_naive(eagen_ctx* ctx, const uint64_t* pts, size_t n, uint64_t* pos_lines, size_t* n_pos,
                                uint64_t* neg_lines, size_t* n_neg);

/* ---- circuit-facing layouts ---------------------------------------------- */

/* the fixed column sizes the circuit gives the coefficients of every f_k        src/config.rs:641-642
 * b_size = (num_pts + base + 1) / 2, a_size = (num_pts + base + 2) / 2 */
int eagen_circuit_sizes(size_t num_pts, uint8_t base, size_t* a_size, size_t* b_size);

/* all functions of a result as fixed-size rows, zero padded: a_out = nf x a_size elements, b_out = nf x b_size elements
 * (row k = digit position k).  EAGEN_E_LEN when a function does not fit its row (for an even list length the canonical a_k has
 * a_size + 1 coefficients: its leading 1 is the monic normalisation the circuit's sizes leave implicit). */
int eagen_result_copy_padded(eagen_result* r, size_t num_pts, uint8_t base, uint64_t* a_out, uint64_t* b_out);

/* RegularFunction::ev of EVERY function of a device-resident result at m points (the circuit evaluates each f_k at the
 * challenge point)                                            src/regular_functions_utils.rs:228-237
 * pts: m Jacobian points; out: nf x m elements, function-major.  The coefficients stay on the device (power table of x,
 * chunked dot products).  Like eagen_eval_function, the value at the identity is 0 (the reference panics there). */
int eagen_result_eval(eagen_ctx* ctx, eagen_result* r, const uint64_t* pts, size_t m, uint64_t* out);

/* challenge post-processors over the curve's BASE field (host side, no device needed)      src/config.rs:163-187
 * eagen_to_curve_x: returns c when c^3 + b is a square; EAGEN_E_DOMAIN otherwise (the reference's loop never terminates there).
 * eagen_y_from_x : sqrt_alt(x^3 + b): *is_square = 1 and y = the root, or 0 and y = sqrt(ROOT_OF_UNITY * (x^3 + b)).
 *                  The root is the one Tonelli-Shanks yields from z = F::ROOT_OF_UNITY (ff::helpers::sqrt_tonelli_shanks).
 * eagen_slope    : xy = x | y (64 bytes) -> (3 x^2 + a) / (2 y); EAGEN_E_DOMAIN when y = 0 (reference: unwrap() panic). */
int eagen_to_curve_x(int curve, const uint64_t* c, uint64_t* x_out);
int eagen_y_from_x(int curve, const uint64_t* x, uint64_t* y_out, int* is_square);
int eagen_slope(int curve, const uint64_t* xy, uint64_t* slope_out);

/* ---- helper API of regular_functions_utils ------------------------------------------------------------ */

/* &Polynomial * &Polynomial                                   src/regular_functions_utils.rs:209-216
 * out must hold la + lb - 1 elements (0 when both are empty).  Exact arithmetic: bits equal mul_naive/mul_fft. */
int eagen_poly_mul(eagen_ctx* ctx, const uint64_t* a, size_t la, const uint64_t* b, size_t lb, uint64_t* out);

/* best_fft(a, omega_pow(S - log_n) or its inverse, log_n): natural order in and out, unscaled
 *                                                             src/regular_functions_utils.rs:119-124        */
int eagen_ntt(eagen_ctx* ctx, uint64_t* data, uint32_t log_n, int inverse);

/* FftPrecomp::omega_pow / omega_pow_inv / half_pow            src/regular_functions_utils.rs:17-24
 * (the Pasta tables the reference lacks: src/precomputed_fft_data.rs only covers bn256::Fr)                 */
int eagen_fft_precomp(int curve, int which /*0 omega_pow, 1 omega_pow_inv, 2 half_pow*/, uint64_t exp, uint64_t* out);

/* batched field inversion, in place (0 stays 0)               z.invert() at src/regular_functions_utils.rs:351-352 */
int eagen_batch_invert(eagen_ctx* ctx, uint64_t* elems, size_t n);

/* RegularFunction::ev at many points                          src/regular_functions_utils.rs:228-237
 * pts: n Jacobian points; out: n field elements (0 for identity points)                                     */
int eagen_eval_function(eagen_ctx* ctx, const uint64_t* a, size_t la, const uint64_t* b, size_t lb,
                        const uint64_t* pts, size_t n, uint64_t* out);

/* ---- multi-GPU: compute_lhs_witness sharded over the GPUs of one box ----------------------
 * The reference API is ONE call (src/argument_witness_calc.rs:87); here every rank (one context per GPU; one host thread or process
 * per context) makes the same call with ITS contiguous share of the points and gets back the functions of ITS share of the digit
 * positions.  Inside the call: K1-K3 on the local point range, ncclAllGather of the d partial digit sums / the digit planes / the
 * multiples table on the library's communication stream, the replicated carry chain (overlapping the gathers), then the trees of
 * positions eagen_position_range(rank).  libnccl.so.2 is resolved at run time; EAGEN_E_NCCL if it is missing or a collective fails.
 *
 *   multi-process : rank 0 calls eagen_comm_unique_id, ships the 128 bytes to the other ranks (any side channel), every rank calls
 *                   eagen_comm_init(ctx, nranks, rank, id).
 *   one process   : eagen_comm_init_all(ctxs, n) (ncclCommInitAll over the contexts' devices); the sharded call must then be issued
 *                   from n host threads, one per context.
 * Every rank must pass the same n_local (pad the last share with zero scalars: a zero scalar contributes nothing); EAGEN_E_LEN
 * otherwise.  Result: function slot s of the returned handle is digit position eagen_result_first_function(res) + s of the whole
 * witness; eagen_result_carry / _carries are the global ones on every rank. */
#define EAGEN_COMM_ID_BYTES 128
int eagen_comm_unique_id(void* id_out /* EAGEN_COMM_ID_BYTES */);
int eagen_comm_init(eagen_ctx* ctx, int nranks, int rank, const void* unique_id);
int eagen_comm_init_all(eagen_ctx** ctxs, int n);
int eagen_comm_destroy(eagen_ctx* ctx);
int eagen_comm_size(const eagen_ctx* ctx);
int eagen_comm_rank(const eagen_ctx* ctx);
/* digit positions (of the MSD-first iteration order) whose trees rank `rank` of `nranks` builds: [begin, end) */
int eagen_position_range(int rank, int nranks, uint32_t d, uint32_t* begin, uint32_t* end);
/* host buffers in; with out != NULL this rank's functions are streamed into `out` (layout of eagen_lhs_witness_stream over the local
 * slots, sizes from eagen_lhs_witness_sharded_layout) while later positions are still being computed */
int eagen_lhs_witness_sharded_layout(int curve, size_t n_total, uint8_t base, int rank, int nranks, size_t* a_stride, size_t* b_stride, size_t* total_bytes);
int eagen_lhs_witness_sharded(eagen_ctx* ctx, const uint64_t* scalars_local, const uint64_t* pts_local, size_t n_local, uint8_t base,
                              uint32_t flags, void* out, size_t out_bytes, eagen_result** res);
/* device-resident inputs */
int eagen_dev_lhs_witness_sharded(eagen_ctx* ctx, const void* d_scalars_local, const void* d_pts_local, size_t n_local, uint8_t base,
                                  uint32_t flags, eagen_result** res);
size_t eagen_result_first_function(const eagen_result* r);

/* ---- synthetic inputs (tests / bench.py) ------------------------------------------------
 * n scalars uniform in [0, 2^k), 2^k <= isqrt(order) (k = 127 on Pasta, 126 on Grumpkin; Montgomery, scalar field) and n distinct points (a + j*b)*G as Jacobian triples
 * with non-trivial z, both functions of `seed` only.  Host-buffer and device-pointer variants.               */
int eagen_synth_inputs(eagen_ctx* ctx, uint64_t seed, size_t n, uint64_t* scalars, uint64_t* pts);
int eagen_dev_synth_inputs(eagen_ctx* ctx, uint64_t seed, size_t n, void* d_scalars, void* d_pts);

/* ---- device-resident entry points (inputs/outputs are CUDA device pointers on the context's device) ------
 * Used by bench.py (inputs already in HBM) and by the multi-GPU driver, which shards the point range across
 * ranks, all-gathers the per-rank partial digit sums and assigns digit positions (trees) to ranks.
 */
/* stage A on a point shard: digits planes (d x n, u8), multiples table (n x (base-1) affine) and the d partial
 * digit sums of this shard as homogeneous projective points (d x 96 B, X|Y|Z).                              */
int eagen_dev_shard_sums(eagen_ctx* ctx, const void* d_scalars, const void* d_pts, size_t n, uint8_t base,
                         void* d_planes, void* d_table, void* d_partial_sums);
/* stand-alone kernels on device buffers (BASELINE config 5 sweeps): K1 digits of n scalars into d x n planes (and n x d rows
 * when d_rows != NULL); K6 `batch` in-place transforms of 2^log_n elements (forward: natural -> bit-reversed order, inverse:
 * bit-reversed -> natural, unscaled).  device_ms (may be NULL) receives the CUDA-event time on the launching stream. */
int eagen_dev_negbase(eagen_ctx* ctx, const void* d_scalars, size_t n, uint8_t base, void* d_planes, void* d_rows, double* device_ms);
int eagen_dev_ntt(eagen_ctx* ctx, void* d_data, uint32_t log_n, size_t batch, int inverse, double* device_ms);
/* stage B: carry chain over `nparts` gathered partial sums (nparts x d projective) -> d affine carries      */
int eagen_dev_carry_chain(eagen_ctx* ctx, const void* d_partial_sums, int nparts, uint8_t base, void* d_carries);
/* stage C: divisor witnesses for the digit positions [pos_begin, pos_end) of the MSD-first iteration order,
 * from the full planes / table of all n points.                                                             */
int eagen_dev_trees(eagen_ctx* ctx, const void* d_planes, const void* d_table, const void* d_carries, size_t n,
                    uint8_t base, uint32_t pos_begin, uint32_t pos_end, uint32_t flags, eagen_result** out);
/* whole path with device-resident inputs */
int eagen_dev_lhs_witness(eagen_ctx* ctx, const void* d_scalars, const void* d_pts, size_t n, uint8_t base,
                          uint32_t flags, eagen_result** out);
/* device pointers of a result's packed functions (layout: function k at a + k*a_stride elements ...) */
int eagen_result_device_view(eagen_result* r, const void** d_a, size_t* a_stride, const void** d_b, size_t* b_stride);

#ifdef __cplusplus
}
#endif
#endif /* EAGEN_MSM_H */
