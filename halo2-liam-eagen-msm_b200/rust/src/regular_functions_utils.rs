//! reference: src/regular_functions_utils.rs -- same public names and signatures for the path; the bodies marshal to
//! the C ABI.  `Polynomial.poly` stays `pub`, `RegularFunction.{a,b}` stay private (reference: :28, :223-224).
use crate::ffi::*;
use crate::gpu::*;
use ff::PrimeField;
use halo2curves::CurveExt;
use std::ops::{Add, Mul, Shr};

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
    /// reference: :54-62 (schoolbook; the reference underflows `usize` on two empty operands, kept)
    pub fn mul_naive(a: &Self, b: &Self) -> Self {
        let mut out = vec![F::ZERO; a.poly.len() + b.poly.len() - 1];
        for (i, x) in a.poly.iter().enumerate() { for (j, y) in b.poly.iter().enumerate() { out[i + j] += *x * y; } }
        Polynomial::new(out)
    }
}
impl<F: PrimeField + FftPrecomp + BaseFieldOf> Polynomial<F> {
    /// reference: :102-129 -- same result (exact arithmetic), computed by the device transform (eagen_poly_mul)
    pub fn mul_fft(&self, other: &Self) -> Self { self * other }
}
/// reference: :31-33
pub fn poly<T: IntoIterator>(it: T) -> Polynomial<T::Item> where T::Item: PrimeField + FftPrecomp { Polynomial::new(it.into_iter().collect()) }

/// reference: :167-175 -- multiplication by x^k
impl<F: PrimeField + FftPrecomp> Shr<usize> for &Polynomial<F> {
    type Output = Polynomial<F>;
    fn shr(self, k: usize) -> Polynomial<F> { Polynomial::new(std::iter::repeat(F::ZERO).take(k).chain(self.poly.iter().cloned()).collect()) }
}
/// reference: :178-195 -- coefficient-wise sum, length of the longer operand
impl<F: PrimeField + FftPrecomp> Add for &Polynomial<F> {
    type Output = Polynomial<F>;
    fn add(self, other: Self) -> Polynomial<F> {
        let (long, short) = if self.poly.len() >= other.poly.len() { (self, other) } else { (other, self) };
        let mut out = long.poly.clone();
        for (o, s) in out.iter_mut().zip(short.poly.iter()) { *o += s; }
        Polynomial::new(out)
    }
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
    /// reference: :239-242
    pub fn from_const(x: C::Base) -> Self { RegularFunction { a: Polynomial::new(vec![x]), b: Polynomial::new(vec![]) } }
    /// reference: :244-246 -- the line a*x + b*y + c
    pub fn from_line(a: C::Base, b: C::Base, c: C::Base) -> Self { RegularFunction { a: Polynomial::new(vec![c, a]), b: Polynomial::new(vec![b]) } }
}
/// reference: :257-264
impl<C: CurveExt> Add for &RegularFunction<C> where C::Base: FftPrecomp {
    type Output = RegularFunction<C>;
    fn add(self, other: Self) -> RegularFunction<C> { RegularFunction { a: &self.a + &other.a, b: &self.b + &other.b } }
}
/// reference: :266-273 -- (a + y b)(a' + y b') with y^2 -> x^3 + A x + B
impl<C: CurveExt> Mul for &RegularFunction<C> where C::Base: FftPrecomp + BaseFieldOf {
    type Output = RegularFunction<C>;
    fn mul(self, other: Self) -> RegularFunction<C> {
        let y2 = Polynomial::new(vec![C::b(), C::a(), C::Base::ZERO, C::Base::ONE]);
        let bb = &(&self.b * &other.b) * &y2;
        RegularFunction { a: &(&self.a * &other.a) + &bb, b: &(&self.a * &other.b) + &(&self.b * &other.a) }
    }
}

/// reference: :426-431 -- (X Z, Y, Z^3) of the Jacobian triple: homogeneous coordinates of the same point
pub fn projective_coords<C: CurveExt>(pt: &C) -> (C::Base, C::Base, C::Base) {
    let (x, y, z) = pt.jacobian_coordinates();
    (x * z, y, z.square() * z)
}
/// reference: :285-303 -- the line through a and b as the cross product of their homogeneous triples; when that vanishes
/// (a == b) the tangent, i.e. the line through a and -(a + b)
pub fn linefunc<C: CurveExt>(a: &C, b: &C) -> RegularFunction<C> where C::Base: FftPrecomp {
    let cross = |p: (C::Base, C::Base, C::Base), q: (C::Base, C::Base, C::Base)| (p.1 * q.2 - p.2 * q.1, p.2 * q.0 - p.0 * q.2, p.0 * q.1 - p.1 * q.0);
    let (pa, pb) = (projective_coords(a), projective_coords(b));
    let mut l = cross(pa, pb);
    if l.0 == C::Base::ZERO && l.1 == C::Base::ZERO && l.2 == C::Base::ZERO {
        let c = -(*a + *b);
        l = cross(pa, projective_coords(&c));
    }
    RegularFunction::from_line(l.0, l.1, l.2)
}

/// reference: :305-408 -- a partial witness: the function with divisor sum[inputs] + [output] - (n+1)[inf].  The tree the
/// reference builds with group_merge is what eagen_divisor_witness runs on the device; these host forms keep the public type
/// for callers that compose witnesses by hand.  `merge` follows :333-360 (identity shortcut, division by the two vertical lines).
#[derive(Clone)]
pub struct Propagation<C: CurveExt> where C::Base: FftPrecomp { pub inputs: Vec<C>, pub output: C, pub wtns: RegularFunction<C> }
impl<C: CurveExt> Propagation<C> where C::Base: FftPrecomp + BaseFieldOf {
    pub fn empty() -> Self { Propagation { inputs: vec![], output: C::identity(), wtns: RegularFunction::from_const(C::Base::ONE) } }
    pub fn from_point(pt: C) -> Self {
        if bool::from(pt.is_identity()) { return Self::empty(); }
        Propagation { inputs: vec![pt], output: -pt, wtns: linefunc(&pt, &-pt) }
    }
    pub fn from_pair(p: C, q: C) -> Self {
        if bool::from(p.is_identity()) { return Self::from_point(q); }
        Propagation { inputs: vec![p, q], output: -(p + q), wtns: linefunc(&p, &q) }
    }
    pub fn merge(a: Self, b: Self) -> Self {
        let output = a.output + b.output;
        let inputs = a.inputs.iter().chain(b.inputs.iter()).cloned().collect();
        if bool::from(a.output.is_identity()) || bool::from(b.output.is_identity()) {
            return Propagation { inputs, output, wtns: &a.wtns * &b.wtns };
        }
        let num = &a.wtns * &(&b.wtns * &linefunc(&-a.output, &-b.output));
        let affine_x = |p: &C| { let (x, _, z) = p.jacobian_coordinates(); x * z.square().invert().unwrap() };
        let (xa, xb) = (affine_x(&a.output), affine_x(&b.output));
        let wtns = RegularFunction::new(num.a.kate_div(xa).kate_div(xb), num.b.kate_div(xa).kate_div(xb));
        Propagation { inputs, output, wtns }
    }
    /// reference: :380-405 -- pair neighbours level by level, an odd tail passes through
    pub fn group_merge(arr: Vec<Self>) -> Self {
        assert!(!arr.is_empty());
        let mut level = arr;
        while level.len() > 1 {
            let mut next = Vec::with_capacity((level.len() + 1) / 2);
            let mut it = level.into_iter();
            while let Some(x) = it.next() { next.push(match it.next() { Some(y) => Self::merge(x, y), None => x }); }
            level = next;
        }
        level.pop().unwrap()
    }
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
