//! Context handling and marshalling shared by the three API modules.
use crate::ffi::*;
use ff::PrimeField;
use halo2curves::CurveExt;
use std::cell::RefCell;
use std::ffi::CStr;

/// Curves the CUDA library is instantiated for.  The reference is generic over `C: CurveExt`; a curve outside this
/// list has no kernels and panics (there is deliberately no CPU fallback).
pub trait GpuCurve: CurveExt {
    const CURVE_ID: i32;
}
impl GpuCurve for halo2curves::pasta::pallas::Point { const CURVE_ID: i32 = EAGEN_CURVE_PALLAS; }
impl GpuCurve for halo2curves::pasta::vesta::Point { const CURVE_ID: i32 = EAGEN_CURVE_VESTA; }
impl GpuCurve for halo2curves::grumpkin::G1 { const CURVE_ID: i32 = EAGEN_CURVE_GRUMPKIN; }

pub struct Ctx(pub *mut eagen_ctx);
impl Drop for Ctx { fn drop(&mut self) { unsafe { eagen_ctx_destroy(self.0) } } }

thread_local! { static CTXS: RefCell<Vec<(i32, Ctx)>> = RefCell::new(vec![]); }

/// One context per (thread, curve); device chosen by EAGEN_DEVICE (default 0).
pub fn with_ctx<R>(curve: i32, f: impl FnOnce(*mut eagen_ctx) -> R) -> R {
    CTXS.with(|c| {
        let mut v = c.borrow_mut();
        if !v.iter().any(|(id, _)| *id == curve) {
            let dev = std::env::var("EAGEN_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            let mut h = std::ptr::null_mut();
            let rc = unsafe { eagen_ctx_create(curve, dev, &mut h) };
            assert!(rc == EAGEN_OK, "eagen_ctx_create failed with status {} (no CPU fallback)", rc);
            v.push((curve, Ctx(h)));
        }
        let h = v.iter().find(|(id, _)| *id == curve).unwrap().1 .0;
        f(h)
    })
}

/// Panics with the library's message: the reference panics (assert!/panic!/unwrap) in the same situations.
pub fn check(ctx: *mut eagen_ctx, rc: i32) {
    if rc != EAGEN_OK {
        let msg = unsafe { CStr::from_ptr(eagen_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("{}", msg);
    }
}

// The marshalling below reinterprets field elements as `[u64; 4]`: checked at compile time for the instantiated fields (a field
// type of another size or alignment fails the build instead of corrupting memory).  Whether the four words ARE the Montgomery
// residue is a property of halo2curves / pasta_curves that the reference itself relies on (`from_raw_bytes_unchecked`,
// reference: src/precomputed_fft_data.rs:72); pin it with one vector produced by the real crate once a toolchain exists.
const _: () = {
    assert!(std::mem::size_of::<halo2curves::pasta::Fp>() == 32 && std::mem::align_of::<halo2curves::pasta::Fp>() <= 8);
    assert!(std::mem::size_of::<halo2curves::pasta::Fq>() == 32 && std::mem::align_of::<halo2curves::pasta::Fq>() <= 8);
    assert!(std::mem::size_of::<halo2curves::bn256::Fr>() == 32 && std::mem::align_of::<halo2curves::bn256::Fr>() <= 8);
    assert!(std::mem::size_of::<halo2curves::bn256::Fq>() == 32 && std::mem::align_of::<halo2curves::bn256::Fq>() <= 8);
};

/// Montgomery limbs of a field element: both halo2curves and pasta_curves store `[u64; 4]` Montgomery residues and
/// the reference itself reinterprets raw bytes that way (reference: src/precomputed_fft_data.rs:72).
pub fn felt_to_limbs<F: PrimeField>(x: &F) -> [u64; 4] {
    assert!(std::mem::size_of::<F>() == 32);
    unsafe { std::mem::transmute_copy::<F, [u64; 4]>(x) }
}
pub fn felt_from_limbs<F: PrimeField>(l: &[u64]) -> F {
    assert!(std::mem::size_of::<F>() == 32 && l.len() == 4);
    let a: [u64; 4] = [l[0], l[1], l[2], l[3]];
    unsafe { std::mem::transmute_copy::<[u64; 4], F>(&a) }
}
/// x | y | z of `jacobian_coordinates()` (reference: src/regular_functions_utils.rs:229,427)
pub fn pack_points<C: CurveExt>(pts: &[C]) -> Vec<u64> where C::Base: PrimeField {
    let mut v = Vec::with_capacity(pts.len() * 12);
    for p in pts {
        let (x, y, z) = p.jacobian_coordinates();
        v.extend_from_slice(&felt_to_limbs(&x));
        v.extend_from_slice(&felt_to_limbs(&y));
        v.extend_from_slice(&felt_to_limbs(&z));
    }
    v
}
/// affine (x, y) with (0,0) = identity  ->  curve point (z = 1)
pub fn point_from_affine<C: CurveExt>(l: &[u64]) -> C where C::Base: PrimeField {
    if l.iter().all(|w| *w == 0) { return C::identity(); }
    C::new_jacobian(felt_from_limbs(&l[0..4]), felt_from_limbs(&l[4..8]), C::Base::ONE).unwrap()
}
