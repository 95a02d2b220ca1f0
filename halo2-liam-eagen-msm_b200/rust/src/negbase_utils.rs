//! reference: src/negbase_utils.rs.  Every public name of the reference module is here with its signature, so `config.rs`
//! (`use crate::negbase_utils::{self, digit_by_id, table_entry_by_id}`, reference: src/config.rs:13,342,486) compiles against
//! this module unchanged.  The scalar loop of `negbase_decompose` and the index helpers are host code in the reference too;
//! the batched GPU entry points are what `compute_lhs_witness` and the circuit's column b use.
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

/// reference: :11-15
pub fn range_check(x: &BigInt) {
    let threshold = BigInt::from(1u8) << 127;
    assert!(x < &threshold);
    assert!(x > &-threshold);
}

/// reference: :46-51 -- index of a digit in the multiples table, None for digit 0
pub fn id_by_digit(digit: u8) -> Option<usize> { if digit == 0 { None } else { Some(digit as usize - 1) } }
/// reference: :54-56
pub fn digit_by_id(id: usize) -> u8 { u8::try_from(id + 1).unwrap() }

/// reference: :58-77 -- Horner over the bits of `id`, most significant first, in base (-base), with one trailing factor:
/// sum_k bit_k (-base)^(k+1).  Generic host code like the reference's (config.rs instantiates it with bn256::Fr);
/// `eagen_table_entry_by_id` gives the same value for the library's base fields and is what the parity tests pin.
pub fn table_entry_by_id<F: PrimeField>(base: u8, id: usize) -> F {
    let nb = -F::from(base as u64);
    let width = usize::BITS - id.leading_zeros();
    (0..width).rev().fold(F::ZERO, |acc, k| (acc + if (id >> k) & 1 == 1 { F::ONE } else { F::ZERO }) * nb)
}

/// reference: :79-124 -- one scalar; the rows come from the device kernel through the batched call (Pallas scalar field: the
/// admissible scalars are < 2^127 + 2 and fit every instantiated field).  Faithful limb indexing, as the reference is written.
pub fn prepare_scalar_witness(sc: &BigInt, base: u8, num_digits: usize, logtable: usize) -> Vec<Vec<Entry>> {
    use halo2curves::pasta::{pallas, Fq};
    let (sign, bytes) = sc.to_bytes_le();
    assert!(sign != Sign::Minus, "prepare_scalar_witness: negative scalar");
    let mut repr = <Fq as PrimeField>::Repr::default();
    repr.as_mut()[..bytes.len()].copy_from_slice(&bytes);
    let s = Fq::from_repr(repr).unwrap();
    prepare_scalar_witness_batch::<pallas::Point>(&[s], base, num_digits, logtable, false).pop().unwrap()
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
