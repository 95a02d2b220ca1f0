//! Raw bindings to include/eagen_msm.h (hand-written; one declaration per C entry point the shim uses).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct eagen_ctx { _private: [u8; 0] }
#[repr(C)] pub struct eagen_result { _private: [u8; 0] }

pub const EAGEN_CURVE_PALLAS: c_int = 0;
pub const EAGEN_CURVE_VESTA: c_int = 1;
pub const EAGEN_CURVE_GRUMPKIN: c_int = 2;

pub const EAGEN_OK: c_int = 0;
pub const EAGEN_CANONICAL: u32 = 0;
pub const EAGEN_RAW_TREE: u32 = 1;
pub const EAGEN_PARTIAL: u32 = 2;
pub const EAGEN_POLY_A: c_int = 0;
pub const EAGEN_POLY_B: c_int = 1;

extern "C" {
    pub fn eagen_ctx_create(curve: c_int, device: c_int, out: *mut *mut eagen_ctx) -> c_int;
    pub fn eagen_ctx_destroy(ctx: *mut eagen_ctx);
    pub fn eagen_last_error(ctx: *const eagen_ctx) -> *const c_char;
    pub fn eagen_fallback_count(ctx: *const eagen_ctx) -> u64;
    pub fn eagen_num_digits(curve: c_int, base: u8, d: *mut u32) -> c_int;
    pub fn eagen_negbase_decompose(ctx: *mut eagen_ctx, scalars: *const u64, n: usize, base: u8, digits: *mut u8) -> c_int;
    pub fn eagen_precompute_multiplicities(ctx: *mut eagen_ctx, pts: *const u64, n: usize, base: u8, out: *mut u64) -> c_int;
    pub fn eagen_lhs_witness(ctx: *mut eagen_ctx, scalars: *const u64, pts: *const u64, n: usize, base: u8, flags: u32,
                             out: *mut *mut eagen_result) -> c_int;
    pub fn eagen_divisor_witness(ctx: *mut eagen_ctx, pts: *const u64, n: usize, flags: u32, out_point: *mut u64,
                                 out: *mut *mut eagen_result) -> c_int;
    pub fn eagen_result_num_functions(r: *const eagen_result) -> usize;
    pub fn eagen_result_poly_len(r: *const eagen_result, k: usize, which: c_int) -> usize;
    pub fn eagen_result_poly_copy(r: *mut eagen_result, k: usize, which: c_int, out: *mut u64) -> c_int;
    pub fn eagen_result_carry(r: *mut eagen_result, out_affine: *mut u64) -> c_int;
    pub fn eagen_result_free(r: *mut eagen_result);
    pub fn eagen_poly_mul(ctx: *mut eagen_ctx, a: *const u64, la: usize, b: *const u64, lb: usize, out: *mut u64) -> c_int;
    pub fn eagen_ntt(ctx: *mut eagen_ctx, data: *mut u64, log_n: u32, inverse: c_int) -> c_int;
    pub fn eagen_fft_precomp(curve: c_int, which: c_int, exp: u64, out: *mut u64) -> c_int;
    pub fn eagen_msm(ctx: *mut eagen_ctx, scalars: *const u64, pts: *const u64, n: usize, out_affine: *mut u64, device_ms: *mut f64) -> c_int;
    pub fn eagen_table_entry_by_id(curve: c_int, base: u8, id: usize, out: *mut u64) -> c_int;
    pub fn eagen_prepare_scalar_witness(ctx: *mut eagen_ctx, scalars: *const u64, n: usize, base: u8, num_digits: u32, logtable: u32,
                                        mode: c_int, out: *mut PswEntry, out_bytes: usize) -> c_int;
    pub fn eagen_divisor_witness_naive(ctx: *mut eagen_ctx, pts: *const u64, n: usize, pos_lines: *mut u64, n_pos: *mut usize,
                                       neg_lines: *mut u64, n_neg: *mut usize) -> c_int;
    pub fn eagen_circuit_sizes(num_pts: usize, base: u8, a_size: *mut usize, b_size: *mut usize) -> c_int;
    pub fn eagen_result_copy_padded(r: *mut eagen_result, num_pts: usize, base: u8, a_out: *mut u64, b_out: *mut u64) -> c_int;
    pub fn eagen_result_eval(ctx: *mut eagen_ctx, r: *mut eagen_result, pts: *const u64, m: usize, out: *mut u64) -> c_int;
    pub fn eagen_to_curve_x(curve: c_int, c: *const u64, x_out: *mut u64) -> c_int;
    pub fn eagen_y_from_x(curve: c_int, x: *const u64, y_out: *mut u64, is_square: *mut c_int) -> c_int;
    pub fn eagen_slope(curve: c_int, xy: *const u64, slope_out: *mut u64) -> c_int;
    // multi-GPU (include/eagen_msm.h, "multi-GPU" section)
    pub fn eagen_comm_unique_id(id_out: *mut u8) -> c_int;
    pub fn eagen_comm_init(ctx: *mut eagen_ctx, nranks: c_int, rank: c_int, unique_id: *const u8) -> c_int;
    pub fn eagen_comm_init_all(ctxs: *mut *mut eagen_ctx, n: c_int) -> c_int;
    pub fn eagen_comm_destroy(ctx: *mut eagen_ctx) -> c_int;
    pub fn eagen_position_range(rank: c_int, nranks: c_int, d: u32, begin: *mut u32, end: *mut u32) -> c_int;
    pub fn eagen_lhs_witness_sharded(ctx: *mut eagen_ctx, scalars: *const u64, pts: *const u64, n_local: usize, base: u8, flags: u32,
                                     out: *mut c_void, out_bytes: usize, res: *mut *mut eagen_result) -> c_int;
    pub fn eagen_result_first_function(r: *const eagen_result) -> usize;
    pub fn eagen_ctx_set_stream_split(ctx: *mut eagen_ctx, percent: *const u32, n: c_int) -> c_int;
}

/// one entry of eagen_prepare_scalar_witness (32 bytes, see include/eagen_msm.h)
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct PswEntry { pub lo: u64, pub hi: u64, pub mask: u32, pub kind: u32, pub zero: u64 }
pub const EAGEN_PSW_FAITHFUL: c_int = 0;
pub const EAGEN_PSW_INTENDED: c_int = 1;
