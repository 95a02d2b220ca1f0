// Curve arithmetic for y^2 = x^3 + b (a = 0: Pallas, Vesta, Grumpkin) on the device.
//
// Accumulation uses the complete homogeneous-projective formulas of Renes-Costello-Batina (2016,
// algorithms 7-9, a = 0): one branch-free code path covers P+Q, P+P, P+(-P) and the identity, which
// is what a warp wants (no divergence on the degenerate geometry the reference's tests are made of:
// repeated points and P/-P pairs, reference: src/argument_witness_calc.rs:141-142,
// src/regular_functions_utils.rs:668).  Results leave the device as affine (x, y), identity = (0, 0),
// because Jacobian / projective triples are representation dependent.
#pragma once
#include "field.cuh"

namespace eagen {

template <class FP>
struct Affine {  // identity encoded as (0,0), which is never on y^2 = x^3 + b with b != 0
    Fe<FP> x, y;
    EAGEN_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
    static EAGEN_HD Affine identity() { Affine a; a.x = Fe<FP>::zero(); a.y = Fe<FP>::zero(); return a; }
};

template <class FP>
struct Proj {  // homogeneous (X : Y : Z), identity = (0 : 1 : 0)
    Fe<FP> x, y, z;
    static EAGEN_HD Proj identity() { Proj p; p.x = Fe<FP>::zero(); p.y = Fe<FP>::one(); p.z = Fe<FP>::zero(); return p; }
    EAGEN_HD bool is_identity() const { return z.is_zero(); }
};

#define EAGEN_DEFINE_CURVE(NAME, BASEF, SCALARF, PREFIX)                                                         \
    struct NAME {                                                                                                \
        typedef BASEF Base;                                                                                      \
        typedef SCALARF Scalar;                                                                                  \
        static EAGEN_HD Fe<BASEF> b() { constexpr uint32_t t[8] = PREFIX##_B_MONT; Fe<BASEF> r;                  \
            for (int i = 0; i < 8; ++i) r.v[i] = t[i]; return r; }                                               \
        static EAGEN_HD Fe<BASEF> b3() { constexpr uint32_t t[8] = PREFIX##_B3_MONT; Fe<BASEF> r;                \
            for (int i = 0; i < 8; ++i) r.v[i] = t[i]; return r; }                                               \
        static EAGEN_HD Fe<BASEF> gx() { constexpr uint32_t t[8] = PREFIX##_GX_MONT; Fe<BASEF> r;                \
            for (int i = 0; i < 8; ++i) r.v[i] = t[i]; return r; }                                               \
        static EAGEN_HD Fe<BASEF> gy() { constexpr uint32_t t[8] = PREFIX##_GY_MONT; Fe<BASEF> r;                \
            for (int i = 0; i < 8; ++i) r.v[i] = t[i]; return r; }                                               \
    };

EAGEN_DEFINE_CURVE(Pallas, PallasFp, PallasFq, EAGEN_PALLAS)
EAGEN_DEFINE_CURVE(Vesta, PallasFq, PallasFp, EAGEN_VESTA)
EAGEN_DEFINE_CURVE(Grumpkin, Bn256Fr, Bn256Fq, EAGEN_GRUMPKIN)

template <class CC>
EAGEN_HD Fe<typename CC::Base> mul_b3(const Fe<typename CC::Base>& a) { return mul(a, CC::b3()); }

// Jacobian (X, Y, Z) with affine (X/Z^2, Y/Z^3)  ->  homogeneous (XZ : Y : Z^3)
// (the same map the reference calls projective_coords, src/regular_functions_utils.rs:426-431)
template <class CC>
EAGEN_HD Proj<typename CC::Base> jacobian_to_proj(const Fe<typename CC::Base>& X, const Fe<typename CC::Base>& Y,
                                                  const Fe<typename CC::Base>& Z) {
    typedef typename CC::Base F;
    if (Z.is_zero()) return Proj<F>::identity();
    Proj<F> p;
    Fe<F> zz = sqr(Z);
    p.x = mul(X, Z);
    p.y = Y;
    p.z = mul(zz, Z);
    return p;
}

// RCB16 algorithm 7 (complete addition, a = 0): 12 M + 2 m_3b
template <class CC>
EAGEN_HD Proj<typename CC::Base> padd(const Proj<typename CC::Base>& p, const Proj<typename CC::Base>& q) {
    typedef typename CC::Base F;
    Fe<F> t0 = mul(p.x, q.x), t1 = mul(p.y, q.y), t2 = mul(p.z, q.z);
    Fe<F> t3 = mul(add(p.x, p.y), add(q.x, q.y));
    t3 = sub(t3, add(t0, t1));
    Fe<F> t4 = mul(add(p.y, p.z), add(q.y, q.z));
    t4 = sub(t4, add(t1, t2));
    Fe<F> y3 = mul(add(p.x, p.z), add(q.x, q.z));
    y3 = sub(y3, add(t0, t2));
    Fe<F> x3 = dbl(t0);
    t0 = add(x3, t0);
    t2 = mul_b3<CC>(t2);
    Fe<F> z3 = add(t1, t2);
    t1 = sub(t1, t2);
    y3 = mul_b3<CC>(y3);
    x3 = mul(t4, y3);
    t2 = mul(t3, t1);
    Proj<F> r;
    r.x = sub(t2, x3);
    y3 = mul(y3, t0);
    t1 = mul(t1, z3);
    r.y = add(t1, y3);
    t0 = mul(t0, t3);
    z3 = mul(z3, t4);
    r.z = add(z3, t0);
    return r;
}

// RCB16 algorithm 8 (complete mixed addition, a = 0, q affine and NOT the identity): 11 M + 2 m_3b
template <class CC>
EAGEN_HD Proj<typename CC::Base> padd_mixed(const Proj<typename CC::Base>& p, const Affine<typename CC::Base>& q) {
    typedef typename CC::Base F;
    Fe<F> t0 = mul(p.x, q.x), t1 = mul(p.y, q.y);
    Fe<F> t3 = mul(add(q.x, q.y), add(p.x, p.y));
    t3 = sub(t3, add(t0, t1));
    Fe<F> t4 = add(mul(q.y, p.z), p.y);
    Fe<F> y3 = add(mul(q.x, p.z), p.x);
    Fe<F> x3 = dbl(t0);
    t0 = add(x3, t0);
    Fe<F> t2 = mul_b3<CC>(p.z);
    Fe<F> z3 = add(t1, t2);
    t1 = sub(t1, t2);
    y3 = mul_b3<CC>(y3);
    x3 = mul(t4, y3);
    t2 = mul(t3, t1);
    Proj<F> r;
    r.x = sub(t2, x3);
    y3 = mul(y3, t0);
    t1 = mul(t1, z3);
    r.y = add(t1, y3);
    t0 = mul(t0, t3);
    z3 = mul(z3, t4);
    r.z = add(z3, t0);
    return r;
}

// RCB16 algorithm 9 (doubling, a = 0): 6 M + 2 S + 1 m_3b
template <class CC>
EAGEN_HD Proj<typename CC::Base> pdbl(const Proj<typename CC::Base>& p) {
    typedef typename CC::Base F;
    Fe<F> t0 = sqr(p.y);
    Fe<F> z3 = dbl(dbl(dbl(t0)));
    Fe<F> t1 = mul(p.y, p.z);
    Fe<F> t2 = mul_b3<CC>(sqr(p.z));
    Fe<F> x3 = mul(t2, z3);
    Fe<F> y3 = add(t0, t2);
    z3 = mul(t1, z3);
    t1 = dbl(t2);
    t2 = add(t1, t2);
    t0 = sub(t0, t2);
    y3 = mul(t0, y3);
    Proj<F> r;
    r.y = add(x3, y3);
    t1 = mul(p.x, p.y);
    x3 = mul(t0, t1);
    r.x = dbl(x3);
    r.z = z3;
    return r;
}

template <class CC>
EAGEN_HD Proj<typename CC::Base> pneg(const Proj<typename CC::Base>& p) {
    Proj<typename CC::Base> r = p;
    r.y = neg(p.y);
    return r;
}

template <class FP>
EAGEN_HD Affine<FP> aneg(const Affine<FP>& p) {
    Affine<FP> r = p;
    r.y = neg(p.y);  // identity (0,0) stays (0,0)
    return r;
}

// small multiple k*P by double-and-add on complete formulas
template <class CC>
EAGEN_HD Proj<typename CC::Base> pmul_small(const Proj<typename CC::Base>& p, uint32_t k) {
    Proj<typename CC::Base> acc = Proj<typename CC::Base>::identity();
    for (int i = 31; i >= 0; --i) {
        acc = pdbl<CC>(acc);
        if ((k >> i) & 1) acc = padd<CC>(acc, p);
    }
    return acc;
}

// (X : Y : Z) -> affine given zinv = 1/Z (identity when Z == 0)
template <class FP>
EAGEN_HD Affine<FP> proj_to_affine(const Proj<FP>& p, const Fe<FP>& zinv) {
    if (p.z.is_zero()) return Affine<FP>::identity();
    Affine<FP> a;
    a.x = mul(p.x, zinv);
    a.y = mul(p.y, zinv);
    return a;
}

}  // namespace eagen
