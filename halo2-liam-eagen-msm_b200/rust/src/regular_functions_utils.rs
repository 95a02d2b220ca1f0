//! reference: src/regular_functions_utils.rs -- same public names and signatures for the path; the bodies marshal to
//! the C ABI.  `Polynomial.poly` stays `pub`, `RegularFunction.{a,b}` stay private (reference: :28, :223-224).
use crate::ffi::*;
use crate::gpu::*;
use ff::PrimeField;
use halo2curves::CurveExt;
use std::ops::Mul;

/// reference: :17-24.  The reference only implements this for bn256::Fr (src/precomputed_fft_data.rs); the library
/// derives the Pasta tables from ROOT_OF_UNITY with the same recipe (src/scripts.rs:44-70).
pub trait FftPrecomp {
    fn omega_pow(exp2: u32) -> Self;
    fn omega_pow_inv(exp2: u32) -> Self;
    fn half_pow(exp: u64) -> Self;
}
macro_rules! impl_precomp {
    ($f:ty, $curve:expr) => {
        impl FftPrecomp for $f {
            fn omega_pow(e: u32) -> Self { let mut o = [0u64; 4]; unsafe { eagen_fft_precomp($curve, 0, e as u64, o.as_mut_ptr()) }; felt_from_limbs(&o) }
            fn omega_pow_inv(e: u32) -> Self { let mut o = [0u64; 4]; unsafe { eagen_fft_precomp($curve, 1, e as u64, o.as_mut_ptr()) }; felt_from_limbs(&o) }
            fn half_pow(e: u64) -> Self { let mut o = [0u64; 4]; unsafe { eagen_fft_precomp($curve, 2, e, o.as_mut_ptr()) }; felt_from_limbs(&o) }
        }
    };
}
impl_precomp!(halo2curves::pasta::Fp, EAGEN_CURVE_PALLAS);
impl_precomp!(halo2curves::pasta::Fq, EAGEN_CURVE_VESTA);
impl_precomp!(halo2curves::bn256::Fr, EAGEN_CURVE_GRUMPKIN);

#[derive(Clone)]
pub struct Polynomial<F: PrimeField + FftPrecomp> { pub poly: Vec<F> }

impl<F: PrimeField + FftPrecomp> Polynomial<F> {
    pub fn new(poly: Vec<F>) -> Self { Polynomial { poly } }
    /// reference: :41-43 (Horner; host side, as in the reference)
    pub fn ev(&self, x: F) -> F { self.poly.iter().rev().fold(F::ZERO, |acc, c| acc * x + c) }
    /// reference: :45-47 (kate_division semantics: quotient by (x - b), remainder dropped)
    pub fn kate_div(&self, b: F) -> Self {
        let mut q = vec![F::ZERO; self.poly.len() - 1];
        let mut tmp = F::ZERO;
        for i in (0..q.len()).rev() { q[i] = self.poly[i + 1] + tmp; tmp = q[i] * b; }
        Polynomial::new(q)
    }
    pub fn scale(&self, sc: F) -> Self { Polynomial::new(self.poly.iter().map(|x| *x * sc).collect()) }
}

/// `&Polynomial * &Polynomial` (reference: :209-216) -> eagen_poly_mul.  `curve_of::<F>()` picks the context whose BASE
/// field is F.
impl<F: PrimeField + FftPrecomp + BaseFieldOf> Mul for &Polynomial<F> {
    type Output = Polynomial<F>;
    fn mul(self, other: Self) -> Self::Output {
        let (la, lb) = (self.poly.len(), other.poly.len());
        let a: Vec<u64> = self.poly.iter().flat_map(|x| felt_to_limbs(x)).collect();
        let b: Vec<u64> = other.poly.iter().flat_map(|x| felt_to_limbs(x)).collect();
        let len = la + lb - 1; // same usize arithmetic as the reference (:55): panics on two empty operands
        let mut out = vec![0u64; len * 4];
        with_ctx(F::CURVE_ID, |ctx| unsafe { check(ctx, eagen_poly_mul(ctx, a.as_ptr(), la, b.as_ptr(), lb, out.as_mut_ptr())) });
        Polynomial::new(out.chunks(4).map(felt_from_limbs).collect())
    }
}
pub trait BaseFieldOf { const CURVE_ID: i32; }
impl BaseFieldOf for halo2curves::pasta::Fp { const CURVE_ID: i32 = EAGEN_CURVE_PALLAS; }
impl BaseFieldOf for halo2curves::pasta::Fq { const CURVE_ID: i32 = EAGEN_CURVE_VESTA; }
impl BaseFieldOf for halo2curves::bn256::Fr { const CURVE_ID: i32 = EAGEN_CURVE_GRUMPKIN; }

/// A function of the form a(x) + y*b(x) on a curve (reference: :220-225)
#[derive(Clone)]
pub struct RegularFunction<C: CurveExt> where C::Base: FftPrecomp { a: Polynomial<C::Base>, b: Polynomial<C::Base> }

impl<C: CurveExt> RegularFunction<C> where C::Base: FftPrecomp {
    pub fn new(a: Polynomial<C::Base>, b: Polynomial<C::Base>) -> Self { RegularFunction { a, b } }
    /// reference: :228-237
    pub fn ev(&self, pt: C) -> C::Base {
        let (x, y, z) = pt.jacobian_coordinates();
        let zinv = z.invert().unwrap();
        let zinvsq = zinv * zinv;
        self.ev_unchecked(x * zinvsq, y * zinvsq * zinv)
    }
    pub fn ev_unchecked(&self, x: C::Base, y: C::Base) -> C::Base { self.a.ev(x) + self.b.ev(x) * y }
    pub fn scale(&self, sc: C::Base) -> Self { RegularFunction { a: self.a.scale(sc), b: self.b.scale(sc) } }
}

pub(crate) unsafe fn function_from_result<C: CurveExt>(res: *mut eagen_result, k: usize) -> RegularFunction<C> where C::Base: FftPrecomp + PrimeField {
    let mut polys = vec![];
    for which in [EAGEN_POLY_A, EAGEN_POLY_B] {
        let len = eagen_result_poly_len(res, k, which);
        let mut buf = vec![0u64; len * 4];
        if len > 0 { assert!(eagen_result_poly_copy(res, k, which, buf.as_mut_ptr()) == EAGEN_OK); }
        polys.push(Polynomial::new(buf.chunks(4).map(felt_from_limbs::<C::Base>).collect()));
    }
    let b = polys.pop().unwrap();
    RegularFunction::new(polys.pop().unwrap(), b)
}

/// reference: :453-467
pub fn compute_divisor_witness_partial<C: GpuCurve>(pts: Vec<C>) -> (RegularFunction<C>, C) where C::Base: FftPrecomp + PrimeField {
    let packed = pack_points(&pts);
    let mut out_pt = [0u64; 8];
    with_ctx(C::CURVE_ID, |ctx| unsafe {
        let mut res = std::ptr::null_mut();
        check(ctx, eagen_divisor_witness(ctx, packed.as_ptr(), pts.len(), EAGEN_CANONICAL | EAGEN_PARTIAL, out_pt.as_mut_ptr(), &mut res));
        let f = function_from_result::<C>(res, 0);
        eagen_result_free(res);
        (f, point_from_affine::<C>(&out_pt))
    })
}

/// reference: :476-480 (panics when the points do not sum to the identity: EAGEN_E_SUM_NONZERO -> panic!)
pub fn compute_divisor_witness<C: GpuCurve>(pts: Vec<C>) -> RegularFunction<C> where C::Base: FftPrecomp + PrimeField {
    let packed = pack_points(&pts);
    with_ctx(C::CURVE_ID, |ctx| unsafe {
        let mut res = std::ptr::null_mut();
        check(ctx, eagen_divisor_witness(ctx, packed.as_ptr(), pts.len(), EAGEN_CANONICAL, std::ptr::null_mut(), &mut res));
        let f = function_from_result::<C>(res, 0);
        eagen_result_free(res);
        f
    })
}

/// reference: :483-495
pub struct Arrangement<C: GpuCurve> where C::Base: FftPrecomp { pub pos: Vec<RegularFunction<C>>, pub neg: Vec<RegularFunction<C>> }

/// reference: :502-551 (lines of the numerator and of the denominator, in the reference's push order)
pub fn compute_divisor_witness_naive<C: GpuCurve>(pts: Vec<C>) -> Arrangement<C> where C::Base: FftPrecomp + PrimeField {
    let packed = pack_points(&pts);
    let cap = pts.len().max(1);
    let (mut pos, mut neg) = (vec![0u64; cap * 12], vec![0u64; cap * 12]);
    let (mut np, mut nn) = (cap, cap);
    with_ctx(C::CURVE_ID, |ctx| unsafe {
        check(ctx, eagen_divisor_witness_naive(ctx, packed.as_ptr(), pts.len(), pos.as_mut_ptr(), &mut np, neg.as_mut_ptr(), &mut nn))
    });
    let lines = |v: &[u64], k: usize| (0..k).map(|i| {
        let l = &v[12 * i..12 * i + 12];   // lx | ly | lz  ->  from_line(lx, ly, lz): a = [lz, lx], b = [ly]
        RegularFunction::new(Polynomial::new(vec![felt_from_limbs(&l[8..12]), felt_from_limbs(&l[0..4])]), Polynomial::new(vec![felt_from_limbs(&l[4..8])]))
    }).collect::<Vec<_>>();
    Arrangement { pos: lines(&pos, np), neg: lines(&neg, nn) }
}
