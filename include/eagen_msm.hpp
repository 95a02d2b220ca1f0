// eagen_msm.hpp -- C++17 host-side mirror of the reference crate's public API for the witness path, header-only over
// the C ABI in eagen_msm.h.  The reference is Rust and no Rust toolchain exists in the build image, so this is the
// compiled-language host layer a caller links against (the Rust shim in halo2-liam-eagen-msm_b200/rust/ has the same
// shape).  Names, argument meaning and error behaviour follow the reference:
//
//   eagen::argument_witness_calc::compute_lhs_witness      src/argument_witness_calc.rs:87-136
//   eagen::argument_witness_calc::precompute_multiplicities  :43-51
//   eagen::argument_witness_calc::num_digits               :89-91 (order / isqrt / logb_ceil)
//   eagen::negbase_utils::negbase_decompose                src/negbase_utils.rs:20-36 (+ pad/reverse, batched)
//   eagen::negbase_utils::id_by_digit / digit_by_id        :46-56
//   eagen::regular_functions_utils::{Polynomial, RegularFunction, compute_divisor_witness(_partial), FftPrecomp}
//                                                          src/regular_functions_utils.rs:17-47,209-273,453-480
//   eagen::negbase_utils::{Entry, prepare_scalar_witness}   src/negbase_utils.rs:39-43,79-124 (batched over scalars)
//   eagen::regular_functions_utils::{Arrangement, compute_divisor_witness_naive}   src/regular_functions_utils.rs:483-551
//   eagen::config::{to_curve_x, y_from_x, slope, circuit_sizes}   src/config.rs:163-187,641-642
//
// Where the reference panics, these functions throw eagen::Error carrying the ABI status.
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "eagen_msm.h"

namespace eagen {

using Felt = std::array<uint64_t, 4>;            // Montgomery limbs, little endian
using JacobianPoint = std::array<uint64_t, 12>;  // x | y | z
using AffinePoint = std::array<uint64_t, 8>;     // x | y, identity = zeros

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

class Context {
public:
    explicit Context(eagen_curve curve, int device = 0) : curve_(curve) {
        int rc = eagen_ctx_create(curve, device, &ctx_);
        if (rc != EAGEN_OK) throw Error(rc, std::string("eagen_ctx_create: ") + eagen_status_string(rc));
    }
    ~Context() { eagen_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    eagen_ctx* raw() const { return ctx_; }
    eagen_curve curve() const { return curve_; }
    void check(int rc) const { if (rc != EAGEN_OK) throw Error(rc, eagen_last_error(ctx_)); }

private:
    eagen_ctx* ctx_ = nullptr;
    eagen_curve curve_;
};

namespace regular_functions_utils {

// reference: trait FftPrecomp, src/regular_functions_utils.rs:17-24
struct FftPrecomp {
    static Felt omega_pow(eagen_curve c, uint32_t exp2) { Felt o; eagen_fft_precomp(c, 0, exp2, o.data()); return o; }
    static Felt omega_pow_inv(eagen_curve c, uint32_t exp2) { Felt o; eagen_fft_precomp(c, 1, exp2, o.data()); return o; }
    static Felt half_pow(eagen_curve c, uint64_t exp) { Felt o; eagen_fft_precomp(c, 2, exp, o.data()); return o; }
};

// reference: struct Polynomial { pub poly: Vec<F> }, :26-29; coefficients low degree first
struct Polynomial {
    std::vector<Felt> poly;
    Polynomial() {}
    explicit Polynomial(std::vector<Felt> p) : poly(std::move(p)) {}
    // &Polynomial * &Polynomial, :209-216
    Polynomial mul(const Context& ctx, const Polynomial& o) const {
        if (poly.empty() && o.poly.empty()) return Polynomial();
        std::vector<Felt> out(poly.size() + o.poly.size() - 1);
        ctx.check(eagen_poly_mul(ctx.raw(), poly.empty() ? nullptr : poly[0].data(), poly.size(),
                                 o.poly.empty() ? nullptr : o.poly[0].data(), o.poly.size(), out.empty() ? nullptr : out[0].data()));
        return Polynomial(std::move(out));
    }
};

// reference: struct RegularFunction { a, b } = a(x) + y b(x), :220-225
struct RegularFunction {
    Polynomial a, b;
    // reference: RegularFunction::ev, :228-237 (batched over points; identity points evaluate to 0)
    std::vector<Felt> ev(const Context& ctx, const std::vector<JacobianPoint>& pts) const {
        std::vector<Felt> out(pts.size());
        if (pts.empty()) return out;
        ctx.check(eagen_eval_function(ctx.raw(), a.poly.empty() ? nullptr : a.poly[0].data(), a.poly.size(),
                                      b.poly.empty() ? nullptr : b.poly[0].data(), b.poly.size(), pts[0].data(), pts.size(), out[0].data()));
        return out;
    }
};

inline RegularFunction function_from_result(eagen_result* r, size_t k) {
    RegularFunction f;
    f.a.poly.resize(eagen_result_poly_len(r, k, EAGEN_POLY_A));
    f.b.poly.resize(eagen_result_poly_len(r, k, EAGEN_POLY_B));
    if (!f.a.poly.empty()) eagen_result_poly_copy(r, k, EAGEN_POLY_A, f.a.poly[0].data());
    if (!f.b.poly.empty()) eagen_result_poly_copy(r, k, EAGEN_POLY_B, f.b.poly[0].data());
    return f;
}

// reference: compute_divisor_witness_partial, :453-467
inline std::pair<RegularFunction, AffinePoint> compute_divisor_witness_partial(const Context& ctx, const std::vector<JacobianPoint>& pts,
                                                                               uint32_t flags = EAGEN_CANONICAL) {
    eagen_result* r = nullptr;
    AffinePoint out{};
    ctx.check(eagen_divisor_witness(ctx.raw(), pts.empty() ? nullptr : pts[0].data(), pts.size(), flags | EAGEN_PARTIAL, out.data(), &r));
    RegularFunction f = function_from_result(r, 0);
    eagen_result_free(r);
    return {std::move(f), out};
}
// reference: compute_divisor_witness, :476-480 (throws EAGEN_E_SUM_NONZERO where the reference panics)
inline RegularFunction compute_divisor_witness(const Context& ctx, const std::vector<JacobianPoint>& pts, uint32_t flags = EAGEN_CANONICAL) {
    eagen_result* r = nullptr;
    ctx.check(eagen_divisor_witness(ctx.raw(), pts.empty() ? nullptr : pts[0].data(), pts.size(), flags & ~(uint32_t)EAGEN_PARTIAL, nullptr, &r));
    RegularFunction f = function_from_result(r, 0);
    eagen_result_free(r);
    return f;
}

// reference: struct Arrangement { pos, neg }, :483-495; a line is RegularFunction::from_line(lx, ly, lz): a = [lz, lx], b = [ly]
struct Arrangement { std::vector<RegularFunction> pos, neg; };

// reference: compute_divisor_witness_naive, :502-551
inline Arrangement compute_divisor_witness_naive(const Context& ctx, const std::vector<JacobianPoint>& pts) {
    size_t cap = pts.empty() ? 1 : pts.size(), np = cap, nn = cap;
    std::vector<uint64_t> pos(cap * 12), neg(cap * 12);
    ctx.check(eagen_divisor_witness_naive(ctx.raw(), pts.empty() ? nullptr : pts[0].data(), pts.size(), pos.data(), &np, neg.data(), &nn));
    auto lines = [](const std::vector<uint64_t>& v, size_t k) {
        std::vector<RegularFunction> out(k);
        for (size_t i = 0; i < k; ++i) {
            Felt lx, ly, lz;
            for (int w = 0; w < 4; ++w) { lx[w] = v[12 * i + w]; ly[w] = v[12 * i + 4 + w]; lz[w] = v[12 * i + 8 + w]; }
            out[i].a.poly = {lz, lx};
            out[i].b.poly = {ly};
        }
        return out;
    };
    return Arrangement{lines(pos, np), lines(neg, nn)};
}

}  // namespace regular_functions_utils

namespace config {

// reference: a_size / b_size in LiamMSMCircuit::synthesize, src/config.rs:641-642 -> {a_size, b_size}
inline std::pair<size_t, size_t> circuit_sizes(size_t num_pts, uint8_t base) {
    size_t a = 0, b = 0;
    int rc = eagen_circuit_sizes(num_pts, base, &a, &b);
    if (rc != EAGEN_OK) throw Error(rc, eagen_status_string(rc));
    return {a, b};
}
// reference: to_curve_x, y_from_x, slope, src/config.rs:163-187 (Error EAGEN_E_DOMAIN where the reference hangs or panics)
inline Felt to_curve_x(eagen_curve c, const Felt& ch) { Felt o; int rc = eagen_to_curve_x(c, ch.data(), o.data()); if (rc != EAGEN_OK) throw Error(rc, eagen_last_error(nullptr)); return o; }
inline Felt y_from_x(eagen_curve c, const Felt& x) { Felt o; int rc = eagen_y_from_x(c, x.data(), o.data(), nullptr); if (rc != EAGEN_OK) throw Error(rc, eagen_last_error(nullptr)); return o; }
inline Felt slope(eagen_curve c, const Felt& x, const Felt& y) {
    uint64_t xy[8]; Felt o;
    for (int w = 0; w < 4; ++w) { xy[w] = x[w]; xy[4 + w] = y[w]; }
    int rc = eagen_slope(c, xy, o.data());
    if (rc != EAGEN_OK) throw Error(rc, eagen_last_error(nullptr));
    return o;
}

}  // namespace config

namespace negbase_utils {

// reference: enum Entry { Scalar(BigInt), Bucket(i128), Limb(i128, u32) }, src/negbase_utils.rs:39-43 (one 32-byte ABI entry)
struct Entry {
    enum Kind : uint32_t { Scalar = 0, Bucket = 1, Limb = 2 };
    uint64_t lo, hi;   // two's complement i128 (Scalar: the canonical scalar)
    uint32_t mask, kind;
    uint64_t zero;
};
static_assert(sizeof(Entry) == 32, "ABI entry is 32 bytes");

// reference: prepare_scalar_witness, :79-124, for n scalars: out[(j * base + row) * (num_limbs + 1) + slot]
inline std::vector<Entry> prepare_scalar_witness(const Context& ctx, const std::vector<Felt>& scalars, uint8_t base, uint32_t num_digits,
                                                 uint32_t logtable, eagen_psw_mode mode = EAGEN_PSW_FAITHFUL) {
    size_t num_limbs = ((size_t)num_digits + logtable - 1) / logtable;
    std::vector<Entry> out(scalars.size() * (size_t)base * (num_limbs + 1));
    ctx.check(eagen_prepare_scalar_witness(ctx.raw(), scalars.empty() ? nullptr : scalars[0].data(), scalars.size(), base, num_digits, logtable,
                                           mode, out.data(), out.size() * sizeof(Entry)));
    return out;
}


// reference: id_by_digit / digit_by_id, src/negbase_utils.rs:46-56
inline std::optional<size_t> id_by_digit(uint8_t digit) { if (digit == 0) return std::nullopt; return (size_t)(digit - 1); }
inline uint8_t digit_by_id(size_t id) { return (uint8_t)(id + 1); }

// reference: table_entry_by_id::<F>, src/negbase_utils.rs:58-77 (F = the curve's base field)
inline Felt table_entry_by_id(eagen_curve curve, uint8_t base, size_t id) { Felt o; eagen_table_entry_by_id(curve, base, id, o.data()); return o; }

// reference: negbase_decompose + pad + reverse for n scalars (src/negbase_utils.rs:20-36, argument_witness_calc.rs:99-101):
// n x d digits, most significant first
inline std::vector<uint8_t> negbase_decompose(const Context& ctx, const std::vector<Felt>& scalars, uint8_t base, uint32_t* d_out = nullptr) {
    uint32_t d = 0;
    ctx.check(eagen_num_digits(ctx.curve(), base, &d));
    std::vector<uint8_t> digits(scalars.size() * d);
    ctx.check(eagen_negbase_decompose(ctx.raw(), scalars.empty() ? nullptr : scalars[0].data(), scalars.size(), base, digits.data()));
    if (d_out) *d_out = d;
    return digits;
}

}  // namespace negbase_utils

namespace argument_witness_calc {

// d = logb_ceil(isqrt(order)+2, base) + 1, reference: :32-40,54-56,89-91
inline uint32_t num_digits(eagen_curve curve, uint8_t base) {
    uint32_t d = 0;
    int rc = eagen_num_digits(curve, base, &d);
    if (rc != EAGEN_OK) throw Error(rc, eagen_status_string(rc));
    return d;
}

// reference: precompute_multiplicities, :43-51 (n points at once; out[j*(base-1) + k-1] = k * P_j, affine)
inline std::vector<AffinePoint> precompute_multiplicities(const Context& ctx, const std::vector<JacobianPoint>& pts, uint8_t base) {
    std::vector<AffinePoint> out(pts.size() * (size_t)(base - 1));
    if (!pts.empty()) ctx.check(eagen_precompute_multiplicities(ctx.raw(), pts[0].data(), pts.size(), base, out[0].data()));
    return out;
}

// halo2 best_multiexp(coeffs, bases) as the reference's tests use it (src/argument_witness_calc.rs:144)
inline AffinePoint best_multiexp(const Context& ctx, const std::vector<Felt>& scalars, const std::vector<JacobianPoint>& pts) {
    if (scalars.size() != pts.size()) throw Error(EAGEN_E_LEN, "incompatible amount of coefficients");
    AffinePoint out{};
    ctx.check(eagen_msm(ctx.raw(), scalars.empty() ? nullptr : scalars[0].data(), pts.empty() ? nullptr : pts[0].data(), pts.size(), out.data(), nullptr));
    return out;
}

struct LhsWitness {
    AffinePoint carry;                                              // sum s_j P_j
    std::vector<regular_functions_utils::RegularFunction> functions;  // index k <-> coefficient of (-base)^k
};

// reference: compute_lhs_witness, :87-136
inline LhsWitness compute_lhs_witness(const Context& ctx, const std::vector<Felt>& scalars, const std::vector<JacobianPoint>& pts, uint8_t base,
                                      uint32_t flags = EAGEN_CANONICAL) {
    if (scalars.size() != pts.size()) throw Error(EAGEN_E_LEN, "incompatible amount of coefficients");  // :88
    eagen_result* r = nullptr;
    ctx.check(eagen_lhs_witness(ctx.raw(), scalars.empty() ? nullptr : scalars[0].data(), pts.empty() ? nullptr : pts[0].data(), pts.size(),
                                base, flags, &r));
    LhsWitness w;
    eagen_result_carry(r, w.carry.data());
    size_t nf = eagen_result_num_functions(r);
    for (size_t k = 0; k < nf; ++k) w.functions.push_back(regular_functions_utils::function_from_result(r, k));
    eagen_result_free(r);
    return w;
}

// The same call over several GPUs: every rank passes its contiguous share of the points and gets the functions
// of its share of the digit positions; `first_function` is the index k of functions[0] in the whole witness.  The communicator comes
// from join() (multi-process: the 128-byte id of unique_id() reaches the other ranks over any side channel) or from
// eagen_comm_init_all (one process, one thread per context).
struct ShardedLhsWitness {
    AffinePoint carry;                                              // sum over ALL ranks' points (the same on every rank)
    size_t first_function = 0;
    std::vector<regular_functions_utils::RegularFunction> functions;
};
inline std::array<unsigned char, EAGEN_COMM_ID_BYTES> unique_id() {
    std::array<unsigned char, EAGEN_COMM_ID_BYTES> id{};
    int rc = eagen_comm_unique_id(id.data());
    if (rc != EAGEN_OK) throw Error(rc, eagen_last_error(nullptr));
    return id;
}
inline void join(const Context& ctx, int nranks, int rank, const std::array<unsigned char, EAGEN_COMM_ID_BYTES>& id) {
    ctx.check(eagen_comm_init(ctx.raw(), nranks, rank, id.data()));
}
inline ShardedLhsWitness compute_lhs_witness_sharded(const Context& ctx, const std::vector<Felt>& my_scalars, const std::vector<JacobianPoint>& my_pts,
                                                     uint8_t base, uint32_t flags = EAGEN_CANONICAL) {
    if (my_scalars.size() != my_pts.size()) throw Error(EAGEN_E_LEN, "incompatible amount of coefficients");
    eagen_result* r = nullptr;
    ctx.check(eagen_lhs_witness_sharded(ctx.raw(), my_scalars.empty() ? nullptr : my_scalars[0].data(), my_pts.empty() ? nullptr : my_pts[0].data(),
                                        my_pts.size(), base, flags, nullptr, 0, &r));
    ShardedLhsWitness w;
    eagen_result_carry(r, w.carry.data());
    w.first_function = eagen_result_first_function(r);
    size_t nf = eagen_result_num_functions(r);
    for (size_t k = 0; k < nf; ++k) w.functions.push_back(regular_functions_utils::function_from_result(r, k));
    eagen_result_free(r);
    return w;
}

}  // namespace argument_witness_calc
}  // namespace eagen
