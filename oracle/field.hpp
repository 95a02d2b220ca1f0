// ORACLE — test infrastructure only. Nothing under oracle/ is linked into, imported by or
// executed from the product path (halo2-liam-eagen-msm_b200/); only tests/, the smoke check and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker.
//
// PARITY UNPINNED: the reference (Rust, un-vendored git dependencies, no toolchain here) cannot be
// executed in this environment and ships no known-answer vectors for this path.  The only constant
// data it holds -- the bn256::Fr FFT tables, reference: src/precomputed_fft_data.rs:3-216 -- IS
// reproduced bit-for-bit by this field code (tests/test_oracle_golden.py).
//
// 4 x 64-bit Montgomery prime-field arithmetic (R = 2^256), the in-memory form used by
// halo2curves / pasta_curves field types that the reference manipulates through ff::PrimeField
// (reference: src/regular_functions_utils.rs:26-29, src/precomputed_fft_data.rs:72).
#pragma once
#include <cstdint>
#include <cstring>
#include "oracle_params.h"

namespace oracle {

typedef unsigned __int128 u128;
typedef uint64_t u64;

template <class P>
struct Fe {
    u64 v[4];

    static Fe zero() { Fe r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0; return r; }
    static Fe one() { Fe r; std::memcpy(r.v, P::ONE, 32); return r; }
    static Fe from_raw(const u64* p) { Fe r; std::memcpy(r.v, p, 32); return r; }  // Montgomery limbs
    // canonical integer (must be < p) -> Montgomery
    static Fe from_canonical(const u64* c) {
        Fe t; std::memcpy(t.v, c, 32);
        Fe r2; std::memcpy(r2.v, P::R2, 32);
        return t * r2;
    }
    static Fe from_u64(u64 x) { u64 c[4] = {x, 0, 0, 0}; return from_canonical(c); }
    static Fe from_i64(int64_t x) { return x >= 0 ? from_u64((u64)x) : -from_u64((u64)(-x)); }
    // Montgomery -> canonical little-endian limbs (PrimeField::to_repr)
    void to_canonical(u64* out) const {
        Fe o; o.v[0] = 1; o.v[1] = o.v[2] = o.v[3] = 0;
        Fe r = (*this) * o;
        std::memcpy(out, r.v, 32);
    }

    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
    bool operator==(const Fe& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    bool operator!=(const Fe& o) const { return !(*this == o); }

    static bool geq_mod(const u64* a) {
        for (int i = 3; i >= 0; --i) {
            if (a[i] > P::MOD[i]) return true;
            if (a[i] < P::MOD[i]) return false;
        }
        return true;
    }

    Fe operator+(const Fe& o) const {
        Fe r; u128 c = 0;
        for (int i = 0; i < 4; ++i) { c += (u128)v[i] + o.v[i]; r.v[i] = (u64)c; c >>= 64; }
        if (c || geq_mod(r.v)) {
            u128 b = 0;
            for (int i = 0; i < 4; ++i) { u128 t = (u128)r.v[i] - P::MOD[i] - (u64)b; r.v[i] = (u64)t; b = (t >> 64) & 1; }
        }
        return r;
    }
    Fe operator-(const Fe& o) const {
        Fe r; u128 b = 0;
        for (int i = 0; i < 4; ++i) { u128 t = (u128)v[i] - o.v[i] - (u64)b; r.v[i] = (u64)t; b = (t >> 64) & 1; }
        if (b) {
            u128 c = 0;
            for (int i = 0; i < 4; ++i) { c += (u128)r.v[i] + P::MOD[i]; r.v[i] = (u64)c; c >>= 64; }
        }
        return r;
    }
    Fe operator-() const { return zero() - *this; }

    // CIOS Montgomery product
    Fe operator*(const Fe& o) const {
        u64 t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) { c += (u128)v[j] * o.v[i] + t[j]; t[j] = (u64)c; c >>= 64; }
            c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
            u64 m = t[0] * P::INV;
            c = (u128)m * P::MOD[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; ++j) { c += (u128)m * P::MOD[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
            c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
        }
        Fe r; std::memcpy(r.v, t, 32);
        if (t[4] || geq_mod(r.v)) {
            u128 b = 0;
            for (int i = 0; i < 4; ++i) { u128 s = (u128)r.v[i] - P::MOD[i] - (u64)b; r.v[i] = (u64)s; b = (s >> 64) & 1; }
        }
        return r;
    }
    Fe& operator+=(const Fe& o) { *this = *this + o; return *this; }
    Fe& operator-=(const Fe& o) { *this = *this - o; return *this; }
    Fe& operator*=(const Fe& o) { *this = *this * o; return *this; }
    Fe square() const { return (*this) * (*this); }

    Fe pow(const u64* e, int nlimbs) const {
        Fe r = one();
        for (int i = nlimbs * 64 - 1; i >= 0; --i) {
            r = r.square();
            if ((e[i / 64] >> (i % 64)) & 1) r = r * (*this);
        }
        return r;
    }
    // Field::invert (Fermat); zero maps to zero (callers check)
    Fe invert() const {
        u64 e[4]; std::memcpy(e, P::MOD, 32);
        e[0] -= 2;  // all moduli here end in ...01 or ...47, no borrow
        return pow(e, 4);
    }
};

#define ORACLE_DEFINE_FIELD(NAME, PREFIX)                                   \
    struct NAME {                                                           \
        static constexpr u64 MOD[4] = PREFIX##_MOD;                         \
        static constexpr u64 ONE[4] = PREFIX##_ONE;                         \
        static constexpr u64 R2[4] = PREFIX##_R2;                           \
        static constexpr u64 ROOT_MONT[4] = PREFIX##_ROOT_MONT;             \
        static constexpr u64 ROOT_INV_MONT[4] = PREFIX##_ROOT_INV_MONT;     \
        static constexpr u64 TWO_INV_MONT[4] = PREFIX##_TWO_INV_MONT;       \
        static constexpr u64 INV = PREFIX##_INV64;                          \
        static constexpr unsigned S = PREFIX##_S;                           \
    };

ORACLE_DEFINE_FIELD(PallasFp, EAGEN_PALLAS_FP)
ORACLE_DEFINE_FIELD(PallasFq, EAGEN_PALLAS_FQ)
ORACLE_DEFINE_FIELD(Bn256Fr, EAGEN_BN256_FR)
ORACLE_DEFINE_FIELD(Bn256Fq, EAGEN_BN256_FQ)

// FftPrecomp restated (reference: src/regular_functions_utils.rs:17-24 and the recipe in
// src/scripts.rs:44-70): omega_pow(k) = ROOT_OF_UNITY^(2^k), omega_pow_inv(k) likewise for the
// inverse root, half_pow(k) = 2^-k; computed by repeated squaring instead of a 64-entry table.
template <class P> Fe<P> omega_pow(unsigned exp2) {
    Fe<P> w = Fe<P>::from_raw(P::ROOT_MONT);
    for (unsigned i = 0; i < exp2; ++i) w = w.square();
    return w;
}
template <class P> Fe<P> omega_pow_inv(unsigned exp2) {
    Fe<P> w = Fe<P>::from_raw(P::ROOT_INV_MONT);
    for (unsigned i = 0; i < exp2; ++i) w = w.square();
    return w;
}
template <class P> Fe<P> half_pow(u64 exp) {
    Fe<P> h = Fe<P>::from_raw(P::TWO_INV_MONT), r = Fe<P>::one();
    for (u64 i = 0; i < exp; ++i) r = r * h;
    return r;
}

}  // namespace oracle
