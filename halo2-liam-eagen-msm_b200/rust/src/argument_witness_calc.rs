//! reference: src/argument_witness_calc.rs -- `compute_lhs_witness` keeps its signature (plus the `GpuCurve` bound
//! that names the instantiated curves); `logb_ceil`, `order` are the reference's host helpers, kept verbatim.
use crate::ffi::*;
use crate::gpu::*;
use crate::regular_functions_utils::{function_from_result, FftPrecomp, RegularFunction};
use ff::PrimeField;
use num_bigint::{BigInt, BigUint, Sign};
use num_traits::{One, Zero};

/// reference: :32-40
pub fn logb_ceil(x: &BigUint, base: u8) -> u8 {
    let mut x = x.clone();
    let mut i = 0;
    while x > BigUint::zero() { x /= base; i += 1; }
    i
}
/// reference: :54-56
pub fn order<Fz: PrimeField>() -> BigInt { BigInt::from_bytes_le(Sign::Plus, (-Fz::ONE).to_repr().as_ref()) + BigInt::one() }

/// reference: :43-51 -- affine-normalised multiples (the group elements are the same; Jacobian representatives differ)
pub fn precompute_multiplicities<C: GpuCurve>(pt: &C, base: u8) -> Vec<C> where C::Base: PrimeField {
    let packed = pack_points(std::slice::from_ref(pt));
    let mut out = vec![0u64; (base as usize - 1) * 8];
    with_ctx(C::CURVE_ID, |ctx| unsafe { check(ctx, eagen_precompute_multiplicities(ctx, packed.as_ptr(), 1, base, out.as_mut_ptr())) });
    out.chunks(8).map(point_from_affine::<C>).collect()
}

/// reference: :87-136.  Same inputs, same outputs: (sum s_j P_j, one RegularFunction per digit position, index k for
/// the coefficient of (-base)^k).  Functions come back in canonical form (trailing zeros trimmed, monic in the term of
/// highest pole order) because the reference's raw scaling depends on Jacobian representatives.
pub fn compute_lhs_witness<C: GpuCurve>(scalars: &[C::ScalarExt], pts: &[C], base: u8) -> (C, Vec<RegularFunction<C>>)
where C::Base: FftPrecomp + PrimeField, C::ScalarExt: PrimeField {
    assert!(scalars.len() == pts.len(), "incompatible amount of coefficients"); // :88
    let sc: Vec<u64> = scalars.iter().flat_map(|s| felt_to_limbs(s)).collect();
    let pp = pack_points(pts);
    with_ctx(C::CURVE_ID, |ctx| unsafe {
        let mut res = std::ptr::null_mut();
        check(ctx, eagen_lhs_witness(ctx, sc.as_ptr(), pp.as_ptr(), pts.len(), base, EAGEN_CANONICAL, &mut res));
        let nf = eagen_result_num_functions(res);
        let fns = (0..nf).map(|k| function_from_result::<C>(res, k)).collect();
        let mut carry = [0u64; 8];
        assert!(eagen_result_carry(res, carry.as_mut_ptr()) == EAGEN_OK);
        eagen_result_free(res);
        (point_from_affine::<C>(&carry), fns)
    })
}
