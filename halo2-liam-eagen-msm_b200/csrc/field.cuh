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
EAGEN_HD Fe<FP> add_portable(const Fe<FP>& a, const Fe<FP>& b) {
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
EAGEN_HD Fe<FP> sub_portable(const Fe<FP>& a, const Fe<FP>& b) {
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


// Portable Montgomery product, CIOS over 32-bit limbs with 64-bit temporaries (host path; reference for the PTX path).
template <class FP>
EAGEN_HD Fe<FP> mul_portable(const Fe<FP>& a, const Fe<FP>& b) {
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

// ------------------------------------------------------------------------------------------------
// Carry-chain primitives.  Device: PTX mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32(.X)
// with the carry in a predicate, so a 32x32+64 multiply-accumulate with carry costs one issue slot.
// Host: the same operations with an explicit carry flag, so the algorithm below can be tested on the CPU.
// ------------------------------------------------------------------------------------------------
namespace cc {
#if defined(__CUDA_ARCH__)
EAGEN_D void mul_lo(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
EAGEN_D void mul_hi(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
EAGEN_D void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
EAGEN_D void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
EAGEN_D void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
EAGEN_D void madc_hi(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
EAGEN_D void add_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
EAGEN_D void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
EAGEN_D void addc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
EAGEN_D void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
EAGEN_D void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
EAGEN_D void subc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
#else
static thread_local uint32_t CF = 0;  // host emulation of the condition-code carry / borrow flag
inline void mul_lo(uint32_t& d, uint32_t a, uint32_t b) { d = a * b; }
inline void mul_hi(uint32_t& d, uint32_t a, uint32_t b) { d = (uint32_t)(((uint64_t)a * b) >> 32); }
inline void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; d = (uint32_t)t; CF = (uint32_t)(t >> 32); }
inline void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + CF; d = (uint32_t)t; CF = (uint32_t)(t >> 32); }
inline void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + CF; d = (uint32_t)t; CF = (uint32_t)(t >> 32); }
inline void madc_hi(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + CF; d = (uint32_t)t; }
inline void add_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; d = (uint32_t)t; CF = (uint32_t)(t >> 32); }
inline void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + CF; d = (uint32_t)t; CF = (uint32_t)(t >> 32); }
inline void addc(uint32_t& d, uint32_t a, uint32_t b) { d = a + b + CF; }
inline void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; d = (uint32_t)t; CF = (uint32_t)((t >> 32) & 1); }
inline void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - CF; d = (uint32_t)t; CF = (uint32_t)((t >> 32) & 1); }
inline void subc(uint32_t& d, uint32_t a, uint32_t b) { d = a - b - CF; }
#endif
#if defined(__CUDA_ARCH__)
EAGEN_D void ripple_cc(uint32_t& x) { asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(x)); }  // x += carry, chain continues
EAGEN_D void neg32(uint32_t& d, uint32_t a) { asm volatile("sub.u32 %0, 0, %1;" : "=r"(d) : "r"(a)); }
// a constant moved through a register so that ptxas keeps the mad.lo.cc/madc.hi.cc pair fusable into one IMAD.WIDE
EAGEN_D uint32_t opaque(uint32_t c) { uint32_t r; asm("mov.b32 %0, %1;" : "=r"(r) : "r"(c)); return r; }  // not volatile: hoistable / CSE-able
#else
inline void ripple_cc(uint32_t& x) { addc_cc(x, x, 0); }
inline void neg32(uint32_t& d, uint32_t a) { d = 0u - a; }
inline uint32_t opaque(uint32_t c) { return c; }
#endif
}  // namespace cc

// final a < 2p -> [0, p) with a borrow chain and selects
template <class FP>
EAGEN_HD void reduce_once_cc(uint32_t* r) {
    uint32_t s[8], brw;
    cc::sub_cc(s[0], r[0], FP::mod(0));
#pragma unroll
    for (int i = 1; i < 8; ++i) cc::subc_cc(s[i], r[i], FP::mod(i));
    cc::subc(brw, 0, 0);  // 0 - 0 - borrow: all ones when r < p
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = brw ? r[i] : s[i];
}

// Field addition / subtraction on carry chains (device) -- the 64-bit C form above compiles to ~40 / ~45 instructions per
// operation (borrow extraction, sign shifts), the chains to 25 / 22: add.cc x8, trial subtraction sub.cc x8, one borrow, 8 selects.
// Every modulus here is below 2^255, so a + b never carries out of limb 7.
template <class FP>
EAGEN_HD Fe<FP> add_chain(const Fe<FP>& a, const Fe<FP>& b) {
    static_assert((FP::mod(7) >> 31) == 0, "add_chain assumes p < 2^255");
    Fe<FP> r;
    cc::add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) cc::addc_cc(r.v[i], a.v[i], b.v[i]);
    cc::addc(r.v[7], a.v[7], b.v[7]);
    reduce_once_cc<FP>(r.v);
    return r;
}
template <class FP>
EAGEN_HD Fe<FP> sub_chain(const Fe<FP>& a, const Fe<FP>& b) {
    Fe<FP> r;
    uint32_t mask;
    cc::sub_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) cc::subc_cc(r.v[i], a.v[i], b.v[i]);
    cc::subc(mask, 0, 0);   // all ones when a < b
    cc::add_cc(r.v[0], r.v[0], FP::mod(0) & mask);
#pragma unroll
    for (int i = 1; i < 7; ++i) cc::addc_cc(r.v[i], r.v[i], FP::mod(i) & mask);
    cc::addc(r.v[7], r.v[7], FP::mod(7) & mask);
    return r;
}
template <class FP>
EAGEN_HD Fe<FP> add(const Fe<FP>& a, const Fe<FP>& b) {
#if defined(__CUDA_ARCH__)
    return add_chain(a, b);
#else
    return add_portable(a, b);
#endif
}
template <class FP>
EAGEN_HD Fe<FP> sub(const Fe<FP>& a, const Fe<FP>& b) {
#if defined(__CUDA_ARCH__)
    return sub_chain(a, b);
#else
    return sub_portable(a, b);
#endif
}
template <class FP>
EAGEN_HD Fe<FP> neg(const Fe<FP>& a) { return sub(Fe<FP>::zero(), a); }
template <class FP>
EAGEN_HD Fe<FP> dbl(const Fe<FP>& a) { return add(a, a); }

// (lo, hi) += c * m inside a carry chain, c a modulus limb (a compile-time constant once the caller's loop is unrolled).
// Measured on B200 (tools/probe/pipe_probe.cu): IMAD.WIDE / IMAD.HI issue at 32 lanes/clk/SM, 32-bit IMAD at 64, IADD3 with
// carry at 128 -- so limbs equal to 1 or a power of two are done with adds and shifts on the ALU pipe, zero limbs only
// propagate the carry, and only the remaining limbs pay for a multiplier slot.
EAGEN_HD void mad_const_pair(uint32_t& lo, uint32_t& hi, uint32_t c, uint32_t m, bool first, bool last) {
    if (c == 0) {
        cc::ripple_cc(lo);
        if (last) cc::addc(hi, hi, 0); else cc::ripple_cc(hi);
#ifndef EAGEN_MUL_NO_ONE_LIMB
    } else if (c == 1) {
        if (first) cc::add_cc(lo, lo, m); else cc::addc_cc(lo, lo, m);
        if (last) cc::addc(hi, hi, 0); else cc::addc_cc(hi, hi, 0);
#endif
#ifndef EAGEN_MUL_NO_POW2_LIMB
    } else if ((c & (c - 1)) == 0 && c != 1) {
        int k = 0;
        while ((c >> k) != 1) ++k;
        uint32_t l = m << k, h = m >> (32 - k);
        if (first) cc::add_cc(lo, lo, l); else cc::addc_cc(lo, lo, l);
        if (last) cc::addc(hi, hi, h); else cc::addc_cc(hi, hi, h);
#endif
    } else {
        uint32_t cr = cc::opaque(c);
        if (first) cc::mad_lo_cc(lo, cr, m, lo); else cc::madc_lo_cc(lo, cr, m, lo);
        if (last) cc::madc_hi(hi, cr, m, hi); else cc::madc_hi_cc(hi, cr, m, hi);
    }
}

// One row of the interleaved (CIOS) Montgomery product on split accumulators.
// The running sum is T = E + O*2^32 with E[0] == 0 on entry (E "even role": limb k = column k; O "odd role": limb k = column k+1).
// The row computes T' = T/2^32 + a*bi + m*p with column 0 cleared; on exit the roles are swapped: O holds the even role
// (O[0] == 0) and E the odd role.  Writing the new odd-role limb k from the old E[k+2] performs the shift for free.
template <class FP, bool FIRST>
EAGEN_HD void mont_row(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
    if (FIRST) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            cc::mul_lo(E[j], a[j + 1], bi); cc::mul_hi(E[j + 1], a[j + 1], bi);
            cc::mul_lo(O[j], a[j], bi); cc::mul_hi(O[j + 1], a[j], bi);
        }
    } else {
        cc::add_cc(O[0], O[0], E[1]);
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            cc::madc_lo_cc(E[j], a[j + 1], bi, E[j + 2]);
            cc::madc_hi_cc(E[j + 1], a[j + 1], bi, E[j + 3]);
        }
        cc::madc_lo_cc(E[6], a[7], bi, 0);
        cc::madc_hi(E[7], a[7], bi, 0);
        cc::mad_lo_cc(O[0], a[0], bi, O[0]);
        cc::madc_hi_cc(O[1], a[0], bi, O[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            cc::madc_lo_cc(O[j], a[j], bi, O[j]);
            cc::madc_hi_cc(O[j + 1], a[j], bi, O[j + 1]);
        }
        cc::addc(E[7], E[7], 0);
    }
    // m = -T0 * p^-1 mod 2^32; the Pasta moduli are 1 mod 2^32, so m = -T0 (one ALU op instead of a multiply)
    uint32_t m;
    if (FP::INV == 0xffffffffu) cc::neg32(m, O[0]); else m = O[0] * FP::INV;
    // odd limbs of p onto the odd-role array
    bool started = false;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        if (FP::mod(j + 1) != 0 || started) {
            mad_const_pair(E[j], E[j + 1], FP::mod(j + 1), m, !started, j == 6);
            started = true;
        }
    }
    // even limbs of p onto the even-role array (p[0] is odd, hence never zero); the carry out lands in column 8 = E[7]
#pragma unroll
    for (int j = 0; j < 8; j += 2) mad_const_pair(O[j], O[j + 1], FP::mod(j), m, j == 0, false);
    cc::addc(E[7], E[7], 0);
}

// Montgomery product on carry chains (device fast path; host-emulated for tests)
template <class FP>
EAGEN_HD Fe<FP> mul_chain(const Fe<FP>& a, const Fe<FP>& b) {
    uint32_t X[8], Y[8];
    mont_row<FP, true>(X, Y, a.v, b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        mont_row<FP, false>(Y, X, a.v, b.v[i]);
        if (i + 1 < 8) mont_row<FP, false>(X, Y, a.v, b.v[i + 1]);
    }
    // roles now: X even (X[0] == 0), Y odd:  result limb k = X[k+1] + Y[k]
    Fe<FP> r;
    cc::add_cc(r.v[0], X[1], Y[0]);
#pragma unroll
    for (int k = 1; k < 7; ++k) cc::addc_cc(r.v[k], X[k + 1], Y[k]);
    cc::addc(r.v[7], Y[7], 0);
    reduce_once_cc<FP>(r.v);
    return r;
}

// ------------------------------------------------------------------------------------------------
// Lazily reduced arithmetic for the transform's butterflies: values live in [0, 2p) between stages (2p < 2^256 for every modulus
// here), twiddles stay canonical.  Congruent to the canonical operations mod p; normalise_lazy brings a value back to [0, p), so
// a transform that normalises its outputs produces the same bits as one on canonical values.
//   mul_lazy(a, w): a < 2p, w < p.  The interleaved reduction leaves T = (a w + M p) / R with M < R, so T < p (2p/R + 1) < 2p
//                   because p < R/2: the final conditional subtraction of the canonical product is simply not needed.
//   add_lazy / sub_lazy: one conditional correction by 2p (a + b may carry out of limb 7: 4p > 2^256 for the Pasta moduli).
// ------------------------------------------------------------------------------------------------
template <class FP>
EAGEN_HD constexpr uint32_t mod2(int i) { return (FP::mod(i) << 1) | (i ? (FP::mod(i - 1) >> 31) : 0u); }   // limb i of 2p

template <class FP>
EAGEN_HD Fe<FP> mul_lazy(const Fe<FP>& a, const Fe<FP>& b) {
    static_assert((FP::mod(7) >> 31) == 0, "lazy reduction assumes p < 2^255");
    uint32_t X[8], Y[8];   // the same chains on the host (emulated carry flag), so the CPU suite runs this very algorithm
    mont_row<FP, true>(X, Y, a.v, b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        mont_row<FP, false>(Y, X, a.v, b.v[i]);
        if (i + 1 < 8) mont_row<FP, false>(X, Y, a.v, b.v[i + 1]);
    }
    Fe<FP> r;
    cc::add_cc(r.v[0], X[1], Y[0]);
#pragma unroll
    for (int k = 1; k < 7; ++k) cc::addc_cc(r.v[k], X[k + 1], Y[k]);
    cc::addc(r.v[7], Y[7], 0);
    return r;
}
template <class FP>
EAGEN_HD Fe<FP> add_lazy(const Fe<FP>& a, const Fe<FP>& b) {
    Fe<FP> r;
    uint32_t c8, s[8], t;
    cc::add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) cc::addc_cc(r.v[i], a.v[i], b.v[i]);
    cc::addc(c8, 0, 0);                                   // bit 256 of a + b
    cc::sub_cc(s[0], r.v[0], mod2<FP>(0));
#pragma unroll
    for (int i = 1; i < 8; ++i) cc::subc_cc(s[i], r.v[i], mod2<FP>(i));
    cc::subc(t, c8, 0);                                   // c8 - borrow: all ones exactly when a + b < 2p
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = (t >> 31) ? r.v[i] : s[i];
    return r;
}
template <class FP>
EAGEN_HD Fe<FP> sub_lazy(const Fe<FP>& a, const Fe<FP>& b) {
    Fe<FP> r;
    uint32_t mask;
    cc::sub_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) cc::subc_cc(r.v[i], a.v[i], b.v[i]);
    cc::subc(mask, 0, 0);   // all ones when a < b
    cc::add_cc(r.v[0], r.v[0], mod2<FP>(0) & mask);
#pragma unroll
    for (int i = 1; i < 7; ++i) cc::addc_cc(r.v[i], r.v[i], mod2<FP>(i) & mask);
    cc::addc(r.v[7], r.v[7], mod2<FP>(7) & mask);
    return r;
}
template <class FP>
EAGEN_HD Fe<FP> normalise_lazy(Fe<FP> a) {   // [0, 2p) -> [0, p)
    reduce_once_cc<FP>(a.v);
    return a;
}

template <class FP>
EAGEN_HD Fe<FP> mul(const Fe<FP>& a, const Fe<FP>& b) {
#if defined(__CUDA_ARCH__)
    return mul_chain(a, b);
#else
    return mul_portable(a, b);
#endif
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
