// ORACLE — test infrastructure only (see field.hpp header).  PARITY UNPINNED.
//
// extern "C" surface of the CPU oracle, loaded with ctypes by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.  Field elements cross this boundary as 4 x u64
// little-endian Montgomery limbs, points as Jacobian x|y|z (12 x u64), the same layout the product's
// C ABI uses (include/eagen_msm.h).
#include <chrono>
#include <cstdio>
#include <string>
#include "witness.hpp"

using namespace oracle;

namespace {
thread_local std::string g_err;

struct ResultBase {
    virtual ~ResultBase() {}
    unsigned d = 0;
    size_t n = 0;
    std::vector<uint8_t> digits;
    std::vector<u64> carries;                  // (d) x 8 affine x|y, identity = zeros
    u64 carry[8];
    std::vector<std::vector<u64>> a, b;        // raw coefficients (reference lengths)
    std::vector<std::vector<u64>> ca, cb;      // canonical form
    double seconds = 0;
};

template <class C> void pack_affine(const Point<C>& p, u64* out) {
    typename Point<C>::F x, y;
    p.to_affine(x, y);
    std::memcpy(out, x.v, 32); std::memcpy(out + 4, y.v, 32);
}
template <class P> std::vector<u64> pack_poly(const Polynomial<P>& p) {
    std::vector<u64> v(p.poly.size() * 4);
    for (size_t i = 0; i < p.poly.size(); ++i) std::memcpy(&v[4 * i], p.poly[i].v, 32);
    return v;
}
template <class C> std::vector<Point<C>> unpack_points(const u64* pts, size_t n) {
    std::vector<Point<C>> v(n);
    for (size_t i = 0; i < n; ++i) {
        v[i].x = Point<C>::F::from_raw(pts + 12 * i);
        v[i].y = Point<C>::F::from_raw(pts + 12 * i + 4);
        v[i].z = Point<C>::F::from_raw(pts + 12 * i + 8);
    }
    return v;
}

template <class C>
ResultBase* run_lhs(const u64* scalars, const u64* pts, size_t n, uint8_t base, int with_functions) {
    typedef Fe<typename C::ScalarP> S;
    std::vector<S> sc(n);
    for (size_t i = 0; i < n; ++i) sc[i] = S::from_raw(scalars + 4 * i);
    std::vector<Point<C>> p = unpack_points<C>(pts, n);
    auto t0 = std::chrono::steady_clock::now();
    LhsWitness<C> w = compute_lhs_witness<C>(sc, p, base, with_functions != 0);
    auto t1 = std::chrono::steady_clock::now();
    ResultBase* r = new ResultBase();
    r->seconds = std::chrono::duration<double>(t1 - t0).count();
    r->d = w.d; r->n = n; r->digits = std::move(w.digits);
    r->carries.assign((size_t)w.d * 8, 0);
    for (unsigned i = 0; i < w.d; ++i) pack_affine<C>(w.carries[i], &r->carries[8 * i]);
    pack_affine<C>(w.carry, r->carry);
    for (auto& f : w.fns) {
        r->a.push_back(pack_poly(f.a)); r->b.push_back(pack_poly(f.b));
        RegularFunction<C> c = canonicalize<C>(f);
        r->ca.push_back(pack_poly(c.a)); r->cb.push_back(pack_poly(c.b));
    }
    return r;
}

template <class C>
ResultBase* run_divisor(const u64* pts, size_t n, int partial, u64* out_point) {
    std::vector<Point<C>> p = unpack_points<C>(pts, n);
    auto t0 = std::chrono::steady_clock::now();
    auto res = compute_divisor_witness_partial<C>(p);
    auto t1 = std::chrono::steady_clock::now();
    if (!partial && !res.second.is_identity()) throw std::runtime_error("compute_divisor_witness: points do not sum to identity");
    if (out_point) pack_affine<C>(res.second, out_point);
    ResultBase* r = new ResultBase();
    r->seconds = std::chrono::duration<double>(t1 - t0).count();
    r->n = n;
    r->a.push_back(pack_poly(res.first.a)); r->b.push_back(pack_poly(res.first.b));
    RegularFunction<C> c = canonicalize<C>(res.first);
    r->ca.push_back(pack_poly(c.a)); r->cb.push_back(pack_poly(c.b));
    return r;
}

template <class P> void field_op(int op, const u64* a, const u64* b, u64* out) {
    typedef Fe<P> F;
    F x = F::from_raw(a), y = b ? F::from_raw(b) : F::zero(), r;
    switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x.invert(); break;
        case 4: r = F::from_canonical(a); break;        // canonical -> Montgomery
        case 5: x.to_canonical(out); return;            // Montgomery -> canonical
        case 6: r = omega_pow<P>((unsigned)b[0]); break;
        case 7: r = omega_pow_inv<P>((unsigned)b[0]); break;
        case 8: r = half_pow<P>(b[0]); break;
        default: throw std::runtime_error("bad field op");
    }
    std::memcpy(out, r.v, 32);
}

template <class P> void poly_mul(const u64* a, size_t la, const u64* b, size_t lb, u64* out, int mode) {
    Polynomial<P> pa, pb;
    pa.poly.resize(la); pb.poly.resize(lb);
    for (size_t i = 0; i < la; ++i) pa.poly[i] = Fe<P>::from_raw(a + 4 * i);
    for (size_t i = 0; i < lb; ++i) pb.poly[i] = Fe<P>::from_raw(b + 4 * i);
    Polynomial<P> r = mode == 1 ? Polynomial<P>::mul_naive(pa, pb) : mode == 2 ? pa.mul_fft(pb) : pa * pb;
    for (size_t i = 0; i < r.poly.size(); ++i) std::memcpy(out + 4 * i, r.poly[i].v, 32);
}

template <class P> void fft(u64* a, unsigned log_n, int inverse) {
    std::vector<Fe<P>> v(size_t(1) << log_n);
    for (size_t i = 0; i < v.size(); ++i) v[i] = Fe<P>::from_raw(a + 4 * i);
    Fe<P> w = inverse ? omega_pow_inv<P>(P::S - log_n) : omega_pow<P>(P::S - log_n);
    best_fft<P>(v, w, log_n);
    for (size_t i = 0; i < v.size(); ++i) std::memcpy(a + 4 * i, v[i].v, 32);
}

template <class C> void msm_naive(const u64* scalars, const u64* pts, size_t n, u64* out) {
    typedef Fe<typename C::ScalarP> S;
    std::vector<Point<C>> p = unpack_points<C>(pts, n);
    std::vector<Point<C>> part(Pool::instance().threads() + 1, Point<C>::identity());
    std::atomic<int> slot{0};
    Pool::instance().parallel_for(n, [&](size_t lo, size_t hi) {
        Point<C> acc = Point<C>::identity();
        for (size_t i = lo; i < hi; ++i) { u64 k[4]; S::from_raw(scalars + 4 * i).to_canonical(k); acc = acc + p[i].mul_limbs(k, 4); }
        part[slot.fetch_add(1)] = acc;
    });
    Point<C> acc = Point<C>::identity();
    for (auto& q : part) acc = acc + q;
    pack_affine<C>(acc, out);
}

template <class C> void curve_op(int op, const u64* a, const u64* b, u64 k, u64* out) {
    std::vector<Point<C>> pa = unpack_points<C>(a, 1);
    Point<C> r;
    switch (op) {
        case 0: r = pa[0] + unpack_points<C>(b, 1)[0]; break;
        case 1: r = pa[0].dbl(); break;
        case 2: r = -pa[0]; break;
        case 3: r = pa[0].mul_small(k); break;
        case 4: out[0] = pa[0].on_curve() ? 1 : 0; return;
        default: throw std::runtime_error("bad curve op");
    }
    pack_affine<C>(r, out);
}

template <class C> int eval_fn(const u64* a, size_t la, const u64* b, size_t lb, const u64* pt, u64* out) {
    RegularFunction<C> f;
    f.a.poly.resize(la); f.b.poly.resize(lb);
    for (size_t i = 0; i < la; ++i) f.a.poly[i] = Point<C>::F::from_raw(a + 4 * i);
    for (size_t i = 0; i < lb; ++i) f.b.poly[i] = Point<C>::F::from_raw(b + 4 * i);
    Point<C> p = unpack_points<C>(pt, 1)[0];
    if (p.is_identity()) return 1;
    typename Point<C>::F v = f.ev(p);
    std::memcpy(out, v.v, 32);
    return 0;
}

// Synthetic inputs of SURVEY.md section 8d as the product's k_synth_inputs defines them (halo2-liam-eagen-msm_b200/csrc/kernels.cuh):
// scalar_j = SplitMix64 draws below 2^bits (2^bits <= isqrt(order)), P_j = (a + j*b) * G for seed-derived odd 64-bit a, b.
// The oracle emits the points affine (z = 1); every consumer of the path is invariant under the Jacobian representative
// (SURVEY.md section 8c), so the GPU's triples with non-trivial z describe the same inputs (pinned by a -m gpu test).
inline u64 splitmix64(u64& s) {
    s += 0x9E3779B97F4A7C15ull;
    u64 z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <class C> void synth_inputs(u64 seed, size_t n, u64* scalars, u64* pts) {
    typedef Fe<typename C::ScalarP> S;
    typedef Point<C> Pt;
    U256 sq = isqrt(order<typename C::ScalarP>());
    int bits = 255;
    while (bits > 0 && !((sq.w[bits / 64] >> (bits % 64)) & 1)) --bits;   // 2^bits <= isqrt(order): 127 on Pasta, 126 on Grumpkin
    for (size_t j = 0; j < n; ++j) {
        u64 st = seed ^ (0xD1B54A32D192ED03ull * (u64)(j + 1));
        u64 c[4] = {0, 0, 0, 0};
        c[0] = splitmix64(st);
        c[1] = splitmix64(st) >> (128 - bits);
        S m = S::from_canonical(c);
        std::memcpy(scalars + 4 * j, m.v, 32);
    }
    u64 s2 = seed;
    const u64 a = splitmix64(s2) | 1, b = splitmix64(s2) | 1;
    const Pt G = Pt::from_affine(C::gx(), C::gy());
    const Pt D = G.mul_limbs(&b, 1);
    Pool::instance().parallel_for(n, [&](size_t lo, size_t hi) {
        // k = a + lo*b as a 128-bit integer, then one addition of D per point
        unsigned __int128 k = (unsigned __int128)a + (unsigned __int128)lo * b;
        u64 kl[2] = {(u64)k, (u64)(k >> 64)};
        Pt acc = G.mul_limbs(kl, 2);
        for (size_t j = lo; j < hi; ++j) {
            Pt q = acc.normalized();
            std::memcpy(pts + 12 * j, q.x.v, 32); std::memcpy(pts + 12 * j + 4, q.y.v, 32); std::memcpy(pts + 12 * j + 8, q.z.v, 32);
            acc = acc + D;
        }
    });
}
}  // namespace

#define DISPATCH_CURVE(curve, EXPR)                                        \
    switch (curve) {                                                       \
        case 0: { typedef Pallas C; EXPR; break; }                         \
        case 1: { typedef Vesta C; EXPR; break; }                          \
        case 2: { typedef Grumpkin C; EXPR; break; }                       \
        default: throw std::runtime_error("unknown curve id");             \
    }
// field ids: 0 pallas_fp, 1 pallas_fq, 2 bn256_fr, 3 bn256_fq
#define DISPATCH_FIELD(field, EXPR)                                        \
    switch (field) {                                                       \
        case 0: { typedef PallasFp P; EXPR; break; }                       \
        case 1: { typedef PallasFq P; EXPR; break; }                       \
        case 2: { typedef Bn256Fr P; EXPR; break; }                        \
        case 3: { typedef Bn256Fq P; EXPR; break; }                        \
        default: throw std::runtime_error("unknown field id");             \
    }
#define GUARD(BODY) try { BODY; return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }

// compute_divisor_witness_naive: lines as lx | ly | lz (12 x u64 each); *n_pos / *n_neg carry the capacities in, counts out
template <class C> void run_naive(const u64* pts, size_t n, u64* pos, size_t* n_pos, u64* neg, size_t* n_neg) {
    Arrangement<C> a = compute_divisor_witness_naive<C>(unpack_points<C>(pts, n));
    if (a.pos.size() > *n_pos || a.neg.size() > *n_neg) throw std::runtime_error("line buffers too small");
    auto put = [](const std::vector<typename Arrangement<C>::Line>& v, u64* out) {
        for (size_t i = 0; i < v.size(); ++i) {
            std::memcpy(out + 12 * i, v[i].lx.v, 32); std::memcpy(out + 12 * i + 4, v[i].ly.v, 32); std::memcpy(out + 12 * i + 8, v[i].lz.v, 32);
        }
    };
    put(a.pos, pos); put(a.neg, neg);
    *n_pos = a.pos.size(); *n_neg = a.neg.size();
}

extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }
int oracle_set_threads(int n) { GUARD(Pool::instance().set_threads(n)) }
int oracle_get_threads() { return Pool::instance().threads(); }

int oracle_num_digits(int curve, uint8_t base, unsigned* d) { GUARD(DISPATCH_CURVE(curve, *d = num_digits<C>(base))) }

// sign + 4 x u64 magnitude -> digits (LSD first); *len gets the count, out must hold >= 260 bytes
int oracle_negbase_decompose(const u64* mag, int negative, uint8_t base, uint8_t* out, size_t* len) {
    GUARD({
        U256 m; std::memcpy(m.w, mag, 32);
        std::vector<uint8_t> dg = negbase_decompose(m, negative != 0, base);
        *len = dg.size();
        std::memcpy(out, dg.data(), dg.size());
    })
}

int oracle_table_entry_by_id(int field, uint8_t base, size_t id, u64* out) {
    GUARD(DISPATCH_FIELD(field, { Fe<P> r = table_entry_by_id<P>(base, id); std::memcpy(out, r.v, 32); }))
}

int oracle_field_op(int field, int op, const u64* a, const u64* b, u64* out) { GUARD(DISPATCH_FIELD(field, field_op<P>(op, a, b, out))) }
int oracle_poly_mul(int field, const u64* a, size_t la, const u64* b, size_t lb, u64* out, int mode) {
    GUARD(DISPATCH_FIELD(field, poly_mul<P>(a, la, b, lb, out, mode)))
}
int oracle_fft(int field, u64* a, unsigned log_n, int inverse) { GUARD(DISPATCH_FIELD(field, fft<P>(a, log_n, inverse))) }
int oracle_curve_op(int curve, int op, const u64* a, const u64* b, u64 k, u64* out) { GUARD(DISPATCH_CURVE(curve, curve_op<C>(op, a, b, k, out))) }
int oracle_msm_naive(int curve, const u64* scalars, const u64* pts, size_t n, u64* out) { GUARD(DISPATCH_CURVE(curve, msm_naive<C>(scalars, pts, n, out))) }
// returns 0 and the value, or 1 when the point is the identity
int oracle_eval_function(int curve, const u64* a, size_t la, const u64* b, size_t lb, const u64* pt, u64* out) {
    try { int r = 0; DISPATCH_CURVE(curve, r = eval_fn<C>(a, la, b, lb, pt, out)); return r; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}

int oracle_synth_inputs(int curve, u64 seed, size_t n, u64* scalars, u64* pts) { GUARD(DISPATCH_CURVE(curve, synth_inputs<C>(seed, n, scalars, pts))) }

int oracle_lhs_witness(int curve, const u64* scalars, const u64* pts, size_t n, uint8_t base, int with_functions, void** handle) {
    GUARD(DISPATCH_CURVE(curve, *handle = run_lhs<C>(scalars, pts, n, base, with_functions)))
}
int oracle_divisor_witness(int curve, const u64* pts, size_t n, int partial, u64* out_point, void** handle) {
    GUARD(DISPATCH_CURVE(curve, *handle = run_divisor<C>(pts, n, partial, out_point)))
}
// prepare_scalar_witness: `mag` = canonical scalar (4 x u64); out = base * (num_limbs+1) entries of 32 bytes:
// value (u64 lo, u64 hi: two's complement i128) | u32 mask | u32 kind (0 Scalar, 1 Bucket, 2 Limb) | 8 zero bytes
int oracle_prepare_scalar_witness(const u64* mag, uint8_t base, size_t num_digits, size_t logtable, int mode, u64* out) {
    GUARD({
        U256 m; std::memcpy(m.w, mag, 32);
        std::vector<PswEntry> v = prepare_scalar_witness(m, base, num_digits, logtable, mode);
        for (size_t i = 0; i < v.size(); ++i) {
            out[4 * i] = (u64)v[i].value; out[4 * i + 1] = (u64)(v[i].value >> 64);
            out[4 * i + 2] = (u64)v[i].mask | ((u64)v[i].kind << 32); out[4 * i + 3] = 0;
        }
    })
}
int oracle_divisor_witness_naive(int curve, const u64* pts, size_t n, u64* pos, size_t* n_pos, u64* neg, size_t* n_neg) {
    GUARD(DISPATCH_CURVE(curve, run_naive<C>(pts, n, pos, n_pos, neg, n_neg)))
}
void oracle_result_free(void* h) { delete (ResultBase*)h; }
unsigned oracle_result_d(void* h) { return ((ResultBase*)h)->d; }
double oracle_result_seconds(void* h) { return ((ResultBase*)h)->seconds; }
size_t oracle_result_num_functions(void* h) { return ((ResultBase*)h)->a.size(); }
void oracle_result_digits(void* h, uint8_t* out) { auto* r = (ResultBase*)h; std::memcpy(out, r->digits.data(), r->digits.size()); }
void oracle_result_carry(void* h, u64* out) { std::memcpy(out, ((ResultBase*)h)->carry, 64); }
void oracle_result_carries(void* h, u64* out) { auto* r = (ResultBase*)h; std::memcpy(out, r->carries.data(), r->carries.size() * 8); }
// which: 0 = a raw, 1 = b raw, 2 = a canonical, 3 = b canonical
static std::vector<u64>& sel(ResultBase* r, size_t k, int which) {
    return which == 0 ? r->a[k] : which == 1 ? r->b[k] : which == 2 ? r->ca[k] : r->cb[k];
}
size_t oracle_result_poly_len(void* h, size_t k, int which) { return sel((ResultBase*)h, k, which).size() / 4; }
void oracle_result_poly_copy(void* h, size_t k, int which, u64* out) {
    auto& v = sel((ResultBase*)h, k, which);
    std::memcpy(out, v.data(), v.size() * 8);
}

}  // extern "C"
