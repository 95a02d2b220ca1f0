//! reference: src/negbase_utils.rs.  The scalar loop of `negbase_decompose` is a host-side helper in the reference
//! too; the batched GPU entry point is what `compute_lhs_witness` uses.  `range_check`, `id_by_digit`, `digit_by_id`,
//! `table_entry_by_id` and the single-scalar `prepare_scalar_witness` are pure host code in the reference and are kept verbatim
//! by the integrating crate (not reproduced here); the batched form below fills the same `Entry` rows for all scalars at once.
use crate::ffi::*;
use crate::gpu::*;
use ff::PrimeField;
use num_bigint::{BigInt, Sign};
use num_traits::Zero;

/// reference: src/negbase_utils.rs:20-36 (unchanged semantics: LSD first, no padding, empty for 0)
pub fn negbase_decompose(x: &BigInt, base: u8) -> Vec<u8> {
    let mut x = x.clone();
    let mut acc = vec![];
    while x != BigInt::zero() {
        let mut digit = x.clone() % base;
        if digit.sign() == Sign::Minus { digit += base; }
        let mut tmp = digit.clone().to_u64_digits().1;
        tmp.push(0);
        acc.push(tmp[0] as u8);
        x = -((x - digit) / base);
    }
    acc
}

/// Batched form used by the path: `n x d` digits, MSD first (what argument_witness_calc.rs:99-101 builds).
pub fn negbase_decompose_batch<C: GpuCurve>(scalars: &[C::ScalarExt], base: u8) -> (Vec<u8>, usize) where C::ScalarExt: PrimeField {
    let mut d = 0u32;
    let rc = unsafe { eagen_num_digits(C::CURVE_ID, base, &mut d) };
    assert!(rc == EAGEN_OK);
    let limbs: Vec<u64> = scalars.iter().flat_map(|s| felt_to_limbs(s)).collect();
    let mut digits = vec![0u8; scalars.len() * d as usize];
    with_ctx(C::CURVE_ID, |ctx| unsafe { check(ctx, eagen_negbase_decompose(ctx, limbs.as_ptr(), scalars.len(), base, digits.as_mut_ptr())) });
    (digits, d as usize)
}

/// reference: src/negbase_utils.rs:39-43
pub enum Entry { Scalar(BigInt), Bucket(i128), Limb(i128, u32) }

/// `prepare_scalar_witness` (reference: src/negbase_utils.rs:79-124) for every scalar in one device call.
/// `intended = false` reproduces the reference as written (limb slot `i % logtable + 1`), `true` uses `i / logtable + 1`.
pub fn prepare_scalar_witness_batch<C: GpuCurve>(scalars: &[C::ScalarExt], base: u8, num_digits: usize, logtable: usize, intended: bool)
    -> Vec<Vec<Vec<Entry>>> where C::ScalarExt: PrimeField {
    let num_limbs = (num_digits + logtable - 1) / logtable;
    let per = base as usize * (num_limbs + 1);
    let limbs: Vec<u64> = scalars.iter().flat_map(|s| felt_to_limbs(s)).collect();
    let mut raw = vec![PswEntry::default(); scalars.len() * per];
    with_ctx(C::CURVE_ID, |ctx| unsafe {
        check(ctx, eagen_prepare_scalar_witness(ctx, limbs.as_ptr(), scalars.len(), base, num_digits as u32, logtable as u32,
                                                if intended { EAGEN_PSW_INTENDED } else { EAGEN_PSW_FAITHFUL }, raw.as_mut_ptr(), raw.len() * 32))
    });
    raw.chunks(per).map(|sc| sc.chunks(num_limbs + 1).map(|row| row.iter().map(|e| {
        let v = ((e.hi as u128) << 64 | e.lo as u128) as i128;
        match e.kind { 0 => Entry::Scalar(BigInt::from(v as u128)), 1 => Entry::Bucket(v), _ => Entry::Limb(v, e.mask) }
    }).collect()).collect()).collect()
}
