// ORACLE — test infrastructure only (see field.hpp header).  PARITY UNPINNED.
//
// Short-Weierstrass curve y^2 = x^3 + b (a = 0: Pallas, Vesta, Grumpkin) in Jacobian coordinates,
// the group law the reference obtains from halo2curves `CurveExt` (+, -, ==, identity,
// jacobian_coordinates; reference call sites: src/argument_witness_calc.rs:48,114,118,123 and
// src/regular_functions_utils.rs:229,298,335,427).  Jacobian triples are representation dependent;
// parity is defined on affine coordinates.
#pragma once
#include "field.hpp"

namespace oracle {

template <class C>
struct Point {
    typedef typename C::BaseP BP;
    typedef Fe<BP> F;
    F x, y, z;

    static Point identity() { Point p; p.x = F::zero(); p.y = F::zero(); p.z = F::zero(); return p; }
    static Point from_affine(const F& ax, const F& ay) { Point p; p.x = ax; p.y = ay; p.z = F::one(); return p; }
    bool is_identity() const { return z.is_zero(); }
    Point operator-() const { Point p = *this; p.y = -p.y; return p; }

    Point dbl() const {
        if (is_identity()) return identity();
        // dbl-2009-l (a = 0)
        F a = x.square(), b = y.square(), c = b.square();
        F d = (x + b).square() - a - c; d = d + d;
        F e = a + a + a, f = e.square();
        Point r;
        r.x = f - (d + d);
        F c8 = c + c; c8 = c8 + c8; c8 = c8 + c8;
        r.y = e * (d - r.x) - c8;
        r.z = y * z; r.z = r.z + r.z;
        return r;
    }

    Point operator+(const Point& o) const {
        if (is_identity()) return o;
        if (o.is_identity()) return *this;
        F z1z1 = z.square(), z2z2 = o.z.square();
        F u1 = x * z2z2, u2 = o.x * z1z1;
        F s1 = y * z2z2 * o.z, s2 = o.y * z1z1 * z;
        if (u1 == u2) {
            if (s1 == s2) return dbl();
            return identity();
        }
        F h = u2 - u1, r = s2 - s1;
        F hh = h.square(), hhh = hh * h, v = u1 * hh;
        Point p;
        p.x = r.square() - hhh - (v + v);
        p.y = r * (v - p.x) - s1 * hhh;
        p.z = z * o.z * h;
        return p;
    }

    bool operator==(const Point& o) const {
        if (is_identity() || o.is_identity()) return is_identity() && o.is_identity();
        F z1z1 = z.square(), z2z2 = o.z.square();
        return x * z2z2 == o.x * z1z1 && y * z2z2 * o.z == o.y * z1z1 * z;
    }
    bool operator!=(const Point& o) const { return !(*this == o); }

    // affine (x/z^2, y/z^3); identity -> (0,0) with return value false
    bool to_affine(F& ax, F& ay) const {
        if (is_identity()) { ax = F::zero(); ay = F::zero(); return false; }
        F zi = z.invert(), zi2 = zi.square();
        ax = x * zi2; ay = y * zi2 * zi;
        return true;
    }
    // same point with z = 1 (identity stays (0,0,0)) -- the oracle's convention for every
    // point fed to linefunc (SURVEY.md section 8c "RAW_TREE")
    Point normalized() const {
        F ax, ay;
        if (!to_affine(ax, ay)) return identity();
        return from_affine(ax, ay);
    }
    bool on_curve() const {
        if (is_identity()) return true;
        F z2 = z.square(), z6 = z2.square() * z2;
        return y.square() == x.square() * x + C::b() * z6;
    }
    // small scalar multiple, double-and-add (Mul<Scalar> of the reference at
    // src/argument_witness_calc.rs:118 restricted to the base-sized scalars it is used with)
    Point mul_small(u64 k) const {
        Point acc = identity();
        for (int i = 63; i >= 0; --i) {
            acc = acc.dbl();
            if ((k >> i) & 1) acc = acc + *this;
        }
        return acc;
    }
    // full-width scalar multiple (canonical little-endian limbs); used for independent MSM checks
    Point mul_limbs(const u64* k, int nlimbs) const {
        Point acc = identity();
        for (int i = nlimbs * 64 - 1; i >= 0; --i) {
            acc = acc.dbl();
            if ((k[i / 64] >> (i % 64)) & 1) acc = acc + *this;
        }
        return acc;
    }
};

struct Pallas {
    typedef PallasFp BaseP; typedef PallasFq ScalarP;
    static Fe<BaseP> a() { return Fe<BaseP>::zero(); }
    static Fe<BaseP> b() { return Fe<BaseP>::from_i64(5); }
    // published generator of the curve (synthetic inputs only; pinned in tests/test_oracle_properties.py)
    static Fe<BaseP> gx() { static constexpr u64 t[4] = EAGEN_PALLAS_GX_MONT; return Fe<BaseP>::from_raw(t); }
    static Fe<BaseP> gy() { static constexpr u64 t[4] = EAGEN_PALLAS_GY_MONT; return Fe<BaseP>::from_raw(t); }
};
struct Vesta {
    typedef PallasFq BaseP; typedef PallasFp ScalarP;
    static Fe<BaseP> a() { return Fe<BaseP>::zero(); }
    static Fe<BaseP> b() { return Fe<BaseP>::from_i64(5); }
    // published generator of the curve (synthetic inputs only; pinned in tests/test_oracle_properties.py)
    static Fe<BaseP> gx() { static constexpr u64 t[4] = EAGEN_VESTA_GX_MONT; return Fe<BaseP>::from_raw(t); }
    static Fe<BaseP> gy() { static constexpr u64 t[4] = EAGEN_VESTA_GY_MONT; return Fe<BaseP>::from_raw(t); }
};
struct Grumpkin {
    typedef Bn256Fr BaseP; typedef Bn256Fq ScalarP;
    static Fe<BaseP> a() { return Fe<BaseP>::zero(); }
    static Fe<BaseP> b() { return Fe<BaseP>::from_i64(-17); }
    // published generator of the curve (synthetic inputs only; pinned in tests/test_oracle_properties.py)
    static Fe<BaseP> gx() { static constexpr u64 t[4] = EAGEN_GRUMPKIN_GX_MONT; return Fe<BaseP>::from_raw(t); }
    static Fe<BaseP> gy() { static constexpr u64 t[4] = EAGEN_GRUMPKIN_GY_MONT; return Fe<BaseP>::from_raw(t); }
};

}  // namespace oracle
