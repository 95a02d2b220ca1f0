// ORACLE — test infrastructure only (see field.hpp header).  PARITY UNPINNED.
//
// CPU restatement of the witness entry point and the negbase helpers
// (reference: src/argument_witness_calc.rs:32-136, src/negbase_utils.rs:20-77).
#pragma once
#include <utility>
#include "poly.hpp"

namespace oracle {

// ---- tiny unsigned 256-bit helpers (stand-in for num-bigint on the sizes the path uses) ----------
struct U256 {
    u64 w[4];
    static U256 from_u64(u64 x) { U256 r; r.w[0] = x; r.w[1] = r.w[2] = r.w[3] = 0; return r; }
    bool is_zero() const { return (w[0] | w[1] | w[2] | w[3]) == 0; }
    int cmp(const U256& o) const {
        for (int i = 3; i >= 0; --i) { if (w[i] != o.w[i]) return w[i] < o.w[i] ? -1 : 1; }
        return 0;
    }
    // divide in place by a small divisor, return remainder
    u64 divmod_small(u64 d) {
        u128 rem = 0;
        for (int i = 3; i >= 0; --i) { u128 cur = (rem << 64) | w[i]; w[i] = (u64)(cur / d); rem = cur % d; }
        return (u64)rem;
    }
    void add_small(u64 x) { u128 c = x; for (int i = 0; i < 4 && c; ++i) { c += w[i]; w[i] = (u64)c; c >>= 64; } }
    void sub_small(u64 x) { u64 b = x; for (int i = 0; i < 4 && b; ++i) { u64 o = w[i]; w[i] = o - b; b = o < b ? 1 : 0; } }
};

// order::<F>()  (reference: src/argument_witness_calc.rs:54-56)
template <class P> U256 order() { U256 r; std::memcpy(r.w, P::MOD, 32); return r; }

// floor(sqrt(p)) for a 256-bit p
inline U256 isqrt(const U256& p) {
    u128 r = 0;
    for (int bit = 127; bit >= 0; --bit) {
        u128 cand = r | ((u128)1 << bit);
        // cand^2 as 256 bit
        u64 c0 = (u64)cand, c1 = (u64)(cand >> 64);
        u128 p00 = (u128)c0 * c0, p01 = (u128)c0 * c1, p11 = (u128)c1 * c1;
        U256 sq;
        sq.w[0] = (u64)p00;
        u128 mid = (p00 >> 64) + (u64)p01 + (u64)p01;
        sq.w[1] = (u64)mid;
        u128 hi = (mid >> 64) + (p01 >> 64) + (p01 >> 64) + (u64)p11;
        sq.w[2] = (u64)hi;
        sq.w[3] = (u64)((hi >> 64) + (p11 >> 64));
        if (sq.cmp(p) <= 0) r = cand;
    }
    U256 out; out.w[0] = (u64)r; out.w[1] = (u64)(r >> 64); out.w[2] = out.w[3] = 0;
    return out;
}

// logb_ceil  (reference: src/argument_witness_calc.rs:32-40): number of base-b digits of x
inline unsigned logb_ceil(U256 x, uint8_t base) {
    unsigned i = 0;
    while (!x.is_zero()) { x.divmod_small(base); ++i; }
    return i;
}

// negbase_decompose  (reference: src/negbase_utils.rs:20-36) on sign + 256-bit magnitude.
// Digits in [0, base), least significant first, x = sum d_i (-base)^i; empty for 0.
inline std::vector<uint8_t> negbase_decompose(U256 mag, bool negative, uint8_t base) {
    if (base < 2) throw std::runtime_error("negbase_decompose: base < 2");
    std::vector<uint8_t> acc;
    while (!mag.is_zero()) {
        U256 q = mag;
        u64 r = q.divmod_small(base);  // |x| = q*base + r
        u64 digit;
        if (!negative) {  // x >= 0: digit = r, x <- -((x - r)/base) = -q
            digit = r; mag = q;
        } else {          // x < 0: digit = (base - r) % base, x <- (|x| + digit)/base
            digit = r ? base - r : 0;
            mag = q; if (r) mag.add_small(1);
        }
        acc.push_back((uint8_t)digit);
        negative = !negative;
    }
    return acc;
}

// table_entry_by_id  (reference: src/negbase_utils.rs:58-77)
template <class P>
Fe<P> table_entry_by_id(uint8_t base, size_t id) {
    typedef Fe<P> F;
    if (id == 0) return F::zero();
    F b = -F::from_u64(base), acc = F::zero();
    std::vector<int> bits;
    while (id > 0) { bits.push_back(id & 1); id >>= 1; }
    size_t l = bits.size();
    for (size_t i = 0; i < l; ++i) {
        if (bits[l - i - 1] == 1) acc += F::one();
        acc *= b;
    }
    return acc;
}

// prepare_scalar_witness  (reference: src/negbase_utils.rs:39-43,79-124)
// Returns base x (num_limbs + 1) entries, row-major.  Entry (0,0) is Entry::Scalar, (i,0) for i >= 1 Entry::Bucket(i128),
// (i,j) for j >= 1 Entry::Limb(i128, u32).  The reference accumulates in i128 / u32; (-base)^i overflows i128 for long
// expansions of small bases (debug build: panic, release build: wrap) -- the oracle wraps (two's complement, mod 2^128).
//   mode 0 = faithful: limb slot i % logtable + 1 and exponent i % logtable exactly as written (:98-101); a slot beyond
//            num_limbs is the reference's out-of-bounds panic and throws here
//   mode 1 = intended: limb slot i / logtable + 1, exponent i % logtable
struct PswEntry {
    unsigned __int128 value = 0;   // two's complement i128 (Scalar: the canonical scalar)
    uint32_t mask = 0;
    uint32_t kind = 0;             // 0 Scalar, 1 Bucket, 2 Limb
};
inline std::vector<PswEntry> prepare_scalar_witness(const U256& sc, uint8_t base, size_t num_digits, size_t logtable, int mode) {
    typedef unsigned __int128 u128w;
    if (base < 2 || logtable == 0) throw std::runtime_error("prepare_scalar_witness: bad base / logtable");
    std::vector<uint8_t> digits = negbase_decompose(sc, false, base);
    if (digits.size() > num_digits) throw std::runtime_error("prepare_scalar_witness: more than num_digits digits");  // :81
    size_t num_limbs = (num_digits + logtable - 1) / logtable;  // :82
    size_t cols = num_limbs + 1;
    std::vector<PswEntry> ret((size_t)base * cols);
    auto negpow = [&](size_t e) { u128w r = 1; for (size_t k = 0; k < e; ++k) r = (u128w)0 - r * base; return r; };  // (-base)^e
    for (size_t i = 0; i < digits.size(); ++i) {  // :93-105
        if (digits[i] == 0) continue;             // id_by_digit -> None
        size_t row = digits[i];                   // id + 1
        size_t slot = (mode == 0 ? i % logtable : i / logtable) + 1, e = i % logtable;
        if (slot > num_limbs) throw std::runtime_error("prepare_scalar_witness: limb slot out of bounds (reference panics)");
        ret[row * cols + 0].value += negpow(i);
        ret[row * cols + slot].value += negpow(e);
        ret[row * cols + slot].mask += (uint32_t)1 << e;
        ret[0 * cols + slot].value += negpow(e);
        ret[0 * cols + slot].mask += (uint32_t)1 << e;
    }
    for (size_t i = 0; i < base; ++i)  // :109-121
        for (size_t j = 0; j < cols; ++j) {
            PswEntry& en = ret[i * cols + j];
            if (i == 0 && j == 0) { en.kind = 0; en.value = ((u128w)sc.w[1] << 64) | sc.w[0]; en.mask = 0; }
            else if (j == 0) { en.kind = 1; en.mask = 0; }
            else en.kind = 2;
        }
    return ret;
}

template <class C>
struct LhsWitness {
    unsigned d = 0;
    std::vector<uint8_t> digits;               // N x d, MSD first (digits_by_scalar after the reverse)
    std::vector<Point<C>> carries;             // carry after each of the d iterations (z = 1 / identity)
    Point<C> carry;                            // sum s_j P_j
    std::vector<RegularFunction<C>> fns;       // fns[k] belongs to digit position k (after ret.reverse())
};

// number of digits d for a curve and base  (reference: src/argument_witness_calc.rs:89-91)
template <class C>
unsigned num_digits(uint8_t base, U256* sq_out = nullptr) {
    U256 sq = isqrt(order<typename C::ScalarP>());
    sq.add_small(2);
    if (sq_out) *sq_out = sq;
    return logb_ceil(sq, base) + 1;
}

// compute_lhs_witness  (reference: src/argument_witness_calc.rs:87-136)
// with_functions=false skips the divisor witnesses (carry / digits only; used for large-N checks)
template <class C>
LhsWitness<C> compute_lhs_witness(const std::vector<Fe<typename C::ScalarP>>& scalars, const std::vector<Point<C>>& pts,
                                  uint8_t base, bool with_functions = true) {
    if (scalars.size() != pts.size()) throw std::runtime_error("incompatible amount of coefficients");  // :88
    if (base < 2) throw std::runtime_error("base < 2");
    LhsWitness<C> out;
    U256 sq_p;
    unsigned d = num_digits<C>(base, &sq_p);  // :89-91
    out.d = d;
    size_t n = scalars.size();
    out.digits.assign(n * d, 0);
    for (size_t j = 0; j < n; ++j) {  // :93-101
        U256 x; scalars[j].to_canonical(x.w);
        if (x.cmp(sq_p) >= 0) throw std::runtime_error("scalar out of range (>= sqrt(p)+2)");  // :97
        std::vector<uint8_t> dg = negbase_decompose(x, false, base);
        // the reference silently truncates with .take(d) (:99); the oracle insists the digits fit
        if (dg.size() > d) throw std::runtime_error("negbase expansion longer than d digits");
        for (size_t i = 0; i < dg.size(); ++i) out.digits[j * d + (d - 1 - i)] = dg[i];  // pad + reverse
    }
    // precompute_multiplicities (:43-51,103): [P, 2P, ..., (b-1)P], kept z = 1
    std::vector<Point<C>> mult(n * (size_t)(base - 1));
    Pool::instance().parallel_for(n, [&](size_t lo, size_t hi) {
        for (size_t j = lo; j < hi; ++j) {
            Point<C> acc = pts[j];
            for (unsigned k = 1; k < base; ++k) { mult[j * (base - 1) + (k - 1)] = acc.normalized(); acc = acc + pts[j]; }
        }
    });
    Point<C> carry = Point<C>::identity();  // :105
    std::vector<RegularFunction<C>> ret;
    for (unsigned i = 0; i < d; ++i) {  // :108
        std::vector<Point<C>> tmp;
        if (!carry.is_identity()) {  // :112-116
            Point<C> nc = (-carry).normalized();
            for (unsigned k = 0; k < base; ++k) tmp.push_back(nc);
        }
        carry = (-carry).mul_small(base);  // :118
        for (size_t j = 0; j < n; ++j) {  // :120-125  (id_by_digit: digit k -> index k-1)
            uint8_t dg = out.digits[j * d + i];
            if (dg != 0) { const Point<C>& m = mult[j * (base - 1) + (dg - 1)]; tmp.push_back(m); carry = carry + m; }
        }
        carry = carry.normalized();
        tmp.push_back(-carry);  // :127
        out.carries.push_back(carry);
        if (with_functions) ret.push_back(compute_divisor_witness<C>(tmp));  // :129
    }
    std::reverse(ret.begin(), ret.end());  // :132
    out.carry = carry;
    out.fns = std::move(ret);
    return out;
}

}  // namespace oracle
