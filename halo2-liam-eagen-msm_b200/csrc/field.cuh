// 256-bit prime-field arithmetic in 8 x 32-bit limbs, Montgomery form (R = 2^256), for the
// sm_100a integer pipe (IMAD / IADD3 on the fma and alu pipes; no tensor cores: this is carry-chain
// arithmetic, not a dense contraction).
//
// Memory form of an element = 32 little-endian bytes of the Montgomery residue, i.e. exactly the
// `[u64; 4]` that halo2curves / pasta_curves field types hold and that the reference reinterprets with
// from_raw_bytes_unchecked (reference: src/precomputed_fft_data.rs:72).  Every value is kept fully
// reduced in [0, p) so byte equality == field equality.
//
// The same source compiles for the host (portable path) so tests/ can check it against the oracle
// without a GPU; the device path below swaps in PTX carry chains.
#pragma once
#include <cstdint>
#include "eagen_params.h"

#if defined(__CUDACC__)
#define EAGEN_HD __host__ __device__ __forceinline__
#define EAGEN_D __device__ __forceinline__
#else
#define EAGEN_HD inline
#define EAGEN_D inline
#endif

namespace eagen {

#define EAGEN_DEFINE_FIELD(NAME, PREFIX)                                                                    \
    struct NAME {                                                                                           \
        static EAGEN_HD constexpr uint32_t mod(int i) { constexpr uint32_t t[8] = PREFIX##_MOD; return t[i]; }        \
        static EAGEN_HD constexpr uint32_t one(int i) { constexpr uint32_t t[8] = PREFIX##_ONE; return t[i]; }        \
        static EAGEN_HD constexpr uint32_t r2(int i) { constexpr uint32_t t[8] = PREFIX##_R2; return t[i]; }          \
        static EAGEN_HD constexpr uint32_t root(int i) { constexpr uint32_t t[8] = PREFIX##_ROOT_MONT; return t[i]; } \
        static EAGEN_HD constexpr uint32_t root_inv(int i) { constexpr uint32_t t[8] = PREFIX##_ROOT_INV_MONT; return t[i]; } \
        static EAGEN_HD constexpr uint32_t two_inv(int i) { constexpr uint32_t t[8] = PREFIX##_TWO_INV_MONT; return t[i]; }   \
        static constexpr uint32_t INV = PREFIX##_INV32;                                                     \
        static constexpr unsigned S = PREFIX##_S;                                                           \
    };

EAGEN_DEFINE_FIELD(PallasFp, EAGEN_PALLAS_FP)
EAGEN_DEFINE_FIELD(PallasFq, EAGEN_PALLAS_FQ)
EAGEN_DEFINE_FIELD(Bn256Fr, EAGEN_BN256_FR)
EAGEN_DEFINE_FIELD(Bn256Fq, EAGEN_BN256_FQ)

template <class FP>
struct alignas(16) Fe {
    uint32_t v[8];

    static EAGEN_HD Fe zero() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = 0;
        return r;
    }
    static EAGEN_HD Fe one() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = FP::one(i);
        return r;
    }
    static EAGEN_HD Fe r2() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = FP::r2(i);
        return r;
    }
    static EAGEN_HD Fe root_of_unity() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = FP::root(i);
        return r;
    }
    static EAGEN_HD Fe root_of_unity_inv() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = FP::root_inv(i);
        return r;
    }
    static EAGEN_HD Fe two_inv() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = FP::two_inv(i);
        return r;
    }
    EAGEN_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= v[i];
        return o == 0;
    }
    EAGEN_HD bool operator==(const Fe& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    EAGEN_HD bool operator!=(const Fe& b) const { return !(*this == b); }
};

// r = a - p if a >= p (a < 2p, `top` = carry-out bit above limb 7)
template <class FP>
EAGEN_HD void reduce_once(uint32_t* a, uint32_t top) {
    uint32_t s[8];
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t t = (uint64_t)a[i] - FP::mod(i) - br;
        s[i] = (uint32_t)t;
        br = (t >> 32) & 1;
    }
    // a >= p  <=>  no final borrow, or the carry-out covers it
    bool ge = top || !br;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ge ? s[i] : a[i];
}

template <class FP>
EAGEN_HD Fe<FP> add(const Fe<FP>& a, const Fe<FP>& b) {
    Fe<FP> r;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    reduce_once<FP>(r.v, (uint32_t)c);
    return r;
}

template <class FP>
EAGEN_HD Fe<FP> sub(const Fe<FP>& a, const Fe<FP>& b) {
    Fe<FP> r;
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t t = (uint64_t)a.v[i] - b.v[i] - br;
        r.v[i] = (uint32_t)t;
        br = (t >> 32) & 1;
    }
    uint32_t mask = br ? 0xffffffffu : 0u;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)r.v[i] + (FP::mod(i) & mask);
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    return r;
}

template <class FP>
EAGEN_HD Fe<FP> neg(const Fe<FP>& a) { return sub(Fe<FP>::zero(), a); }
template <class FP>
EAGEN_HD Fe<FP> dbl(const Fe<FP>& a) { return add(a, a); }

// Montgomery product, CIOS over 32-bit limbs.
template <class FP>
EAGEN_HD Fe<FP> mul(const Fe<FP>& a, const Fe<FP>& b) {
    uint32_t t[8];
    uint32_t t8 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c += (uint64_t)a.v[j] * b.v[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t8;
        t8 = (uint32_t)c;
        uint32_t t9 = (uint32_t)(c >> 32);
        uint32_t m = t[0] * FP::INV;
        c = (uint64_t)m * FP::mod(0) + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            c += (uint64_t)m * FP::mod(j) + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t8;
        t[7] = (uint32_t)c;
        t8 = t9 + (uint32_t)(c >> 32);
    }
    Fe<FP> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = t[i];
    reduce_once<FP>(r.v, t8);
    return r;
}

template <class FP>
EAGEN_HD Fe<FP> sqr(const Fe<FP>& a) { return mul(a, a); }

template <class FP>
EAGEN_HD Fe<FP> from_u32(uint32_t x) {  // small integer -> Montgomery
    Fe<FP> t = Fe<FP>::zero();
    t.v[0] = x;
    return mul(t, Fe<FP>::r2());
}
template <class FP>
EAGEN_HD Fe<FP> to_canonical(const Fe<FP>& a) {  // Montgomery -> canonical integer limbs
    Fe<FP> o = Fe<FP>::zero();
    o.v[0] = 1;
    return mul(a, o);
}
template <class FP>
EAGEN_HD Fe<FP> from_canonical(const Fe<FP>& a) { return mul(a, Fe<FP>::r2()); }

// a^(p-2); 0 -> 0.  Plain square-and-multiply over the constant exponent (only used once per batch).
template <class FP>
EAGEN_HD Fe<FP> inv(const Fe<FP>& a) {
    Fe<FP> r = Fe<FP>::one();
    uint32_t ex[8];  // p - 2 with borrow (the Pasta moduli are 1 mod 2^32)
    uint32_t br = 2;
    for (int i = 0; i < 8; ++i) {
        uint32_t m = FP::mod(i);
        ex[i] = m - br;
        br = m < br ? 1u : 0u;
    }
    for (int i = 7; i >= 0; --i) {
        uint32_t e = ex[i];
        for (int bit = 31; bit >= 0; --bit) {
            r = sqr(r);
            if ((e >> bit) & 1) r = mul(r, a);
        }
    }
    return r;
}

// w^(2^k)
template <class FP>
EAGEN_HD Fe<FP> pow2k(Fe<FP> w, unsigned k) {
    for (unsigned i = 0; i < k; ++i) w = sqr(w);
    return w;
}
// primitive 2^log_n-th root of unity (forward) or its inverse: FftPrecomp::omega_pow(S - log_n)
// (reference: src/regular_functions_utils.rs:17-24,111-113)
template <class FP>
EAGEN_HD Fe<FP> omega_for(unsigned log_n, bool inverse) {
    return pow2k(inverse ? Fe<FP>::root_of_unity_inv() : Fe<FP>::root_of_unity(), FP::S - log_n);
}

}  // namespace eagen
