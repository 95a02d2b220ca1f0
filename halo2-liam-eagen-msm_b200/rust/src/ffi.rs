//! Raw bindings to include/eagen_msm.h (hand-written; one declaration per C entry point the shim uses).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

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
}
