// ORACLE — test infrastructure only (see field.hpp header).  PARITY UNPINNED.
//
// CPU restatement of the reference's polynomial / regular-function / divisor-witness code
// (reference: src/regular_functions_utils.rs).  Each function cites the lines it follows.
// Third-party semantics restated from their published behaviour (halo2_proofs::arithmetic,
// un-vendored, un-pinned git dependency -- reference: Cargo.toml:10):
//   best_fft(a, omega, log_n)  in-place radix-2 DFT, natural order in and out
//   kate_division(a, b)        quotient of a(x) by (x - b), remainder dropped, len-1 coefficients
//   eval_polynomial            Horner
//   parallelize(v, f)          contiguous chunks on a thread pool
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>
#include "curve.hpp"

namespace oracle {

// ---------------------------------------------------------------------------------------------
// parallelize(): contiguous chunks over a fixed pool; nested calls run inline (the reference gets
// the same effect from rayon's scoped pool).  reference: src/regular_functions_utils.rs:391
// ---------------------------------------------------------------------------------------------
class Pool {
public:
    static Pool& instance() { static Pool p; return p; }
    void set_threads(int n) {
        stop_workers();
        nthreads_ = std::max(1, n);
        start_workers();
    }
    int threads() const { return nthreads_; }
    // f(lo, hi) over [0, n)
    void parallel_for(size_t n, const std::function<void(size_t, size_t)>& f) {
        if (n == 0) return;
        if (nthreads_ <= 1 || in_worker_ || n == 1 || busy_.exchange(true)) { f(0, n); return; }
        size_t chunks = std::min<size_t>(nthreads_, n);
        size_t per = (n + chunks - 1) / chunks;
        {
            std::unique_lock<std::mutex> lk(m_);
            fn_ = &f; n_ = n; per_ = per; next_ = 1; pending_ = (int)chunks - 1; ++gen_;
            chunks_ = chunks;
        }
        cv_.notify_all();
        in_worker_ = true;
        f(0, std::min(per, n));
        in_worker_ = false;
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
        busy_ = false;
    }
    ~Pool() { stop_workers(); }

private:
    Pool() { nthreads_ = (int)std::max(1u, std::thread::hardware_concurrency()); start_workers(); }
    void start_workers() {
        quit_ = false;
        for (int i = 1; i < nthreads_; ++i) workers_.emplace_back([this] { worker(); });
    }
    void stop_workers() {
        { std::unique_lock<std::mutex> lk(m_); quit_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
    }
    void worker() {
        in_worker_ = true;
        uint64_t seen = 0;
        for (;;) {
            size_t c;
            const std::function<void(size_t, size_t)>* f;
            size_t n, per;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return quit_ || (gen_ != seen && fn_ && next_ < chunks_); });
                if (quit_) return;
                c = next_++;
                if (next_ >= chunks_) seen = gen_;
                f = fn_; n = n_; per = per_;
            }
            size_t lo = c * per, hi = std::min(n, lo + per);
            if (lo < hi) (*f)(lo, hi);
            std::unique_lock<std::mutex> lk(m_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    int nthreads_ = 1;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t, size_t)>* fn_ = nullptr;
    size_t n_ = 0, per_ = 0, next_ = 0, chunks_ = 0;
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool quit_ = false;
    std::atomic<bool> busy_{false};
    static thread_local bool in_worker_;
};
inline thread_local bool Pool::in_worker_ = false;

inline unsigned log2_floor(size_t num) {  // reference: src/regular_functions_utils.rs:197-207
    if (num == 0) throw std::runtime_error("log2_floor(0)");
    unsigned p = 0;
    while ((size_t(1) << (p + 1)) <= num) ++p;
    return p;
}

// best_fft restated: natural order in -> natural order out, a[k] <- sum_j a[j] omega^(jk)
template <class P>
void best_fft(std::vector<Fe<P>>& a, const Fe<P>& omega, unsigned log_n) {
    typedef Fe<P> F;
    size_t n = size_t(1) << log_n;
    if (a.size() != n) throw std::runtime_error("best_fft: size");
    for (size_t k = 0; k < n; ++k) {
        size_t rk = 0;
        for (unsigned b = 0; b < log_n; ++b) rk |= ((k >> b) & 1) << (log_n - 1 - b);
        if (k < rk) std::swap(a[k], a[rk]);
    }
    std::vector<F> tw(n / 2 ? n / 2 : 1);
    tw[0] = F::one();
    for (size_t i = 1; i < n / 2; ++i) tw[i] = tw[i - 1] * omega;
    for (unsigned s = 0; s < log_n; ++s) {
        size_t half = size_t(1) << s, step = n >> (s + 1);
        auto body = [&](size_t lo, size_t hi) {
            for (size_t t = lo; t < hi; ++t) {
                size_t grp = t >> s, j = t & (half - 1);
                size_t i0 = (grp << (s + 1)) + j, i1 = i0 + half;
                F u = a[i0], v = a[i1] * tw[j * step];
                a[i0] = u + v; a[i1] = u - v;
            }
        };
        if (n >= 4096) Pool::instance().parallel_for(n / 2, body); else body(0, n / 2);
    }
}

// ---------------------------------------------------------------------------------------------
// Polynomial  (reference: src/regular_functions_utils.rs:26-216), coefficients low degree first
// ---------------------------------------------------------------------------------------------
template <class P>
struct Polynomial {
    typedef Fe<P> F;
    std::vector<F> poly;
    Polynomial() {}
    explicit Polynomial(std::vector<F> v) : poly(std::move(v)) {}

    F ev(const F& x) const {  // :41-43  eval_polynomial = Horner
        F acc = F::zero();
        for (size_t i = poly.size(); i-- > 0;) acc = acc * x + poly[i];
        return acc;
    }
    Polynomial kate_div(const F& b) const {  // :45-47  kate_division
        if (poly.empty()) throw std::runtime_error("kate_div of empty polynomial");
        std::vector<F> q(poly.size() - 1);
        F tmp = F::zero();
        for (size_t i = poly.size() - 1; i-- > 0;) {  // q[i] = a[i+1] + b*q[i+1]
            F lead = poly[i + 1] + tmp;
            q[i] = lead;
            tmp = lead * b;
        }
        return Polynomial(std::move(q));
    }
    Polynomial scale(const F& sc) const {  // :49-51
        std::vector<F> r(poly.size());
        for (size_t i = 0; i < poly.size(); ++i) r[i] = poly[i] * sc;
        return Polynomial(std::move(r));
    }
    static Polynomial mul_naive(const Polynomial& a, const Polynomial& b) {  // :54-62
        size_t la = a.poly.size(), lb = b.poly.size();
        // The reference computes la+lb-1 in usize and panics (debug) when both are empty; that
        // combination is defined here as the empty product (documented deviation).
        if (la + lb == 0) return Polynomial();
        std::vector<F> r(la + lb - 1, F::zero());
        for (size_t i = 0; i < la; ++i)
            for (size_t j = 0; j < lb; ++j) r[i + j] += a.poly[i] * b.poly[j];
        return Polynomial(std::move(r));
    }
    Polynomial mul_fft(const Polynomial& o) const {  // :102-129
        size_t length = poly.size() + o.poly.size() - 1;
        unsigned loglength = log2_floor(length) + 1;
        size_t padded = size_t(1) << loglength;
        std::vector<F> a(padded, F::zero()), b(padded, F::zero());
        std::copy(poly.begin(), poly.end(), a.begin());
        std::copy(o.poly.begin(), o.poly.end(), b.begin());
        if (P::S < loglength) throw std::runtime_error("mul_fft: F::S < loglength");  // :110
        F omega = omega_pow<P>(P::S - loglength), omega_inv = omega_pow_inv<P>(P::S - loglength);
        F scaling = half_pow<P>(loglength);
        best_fft<P>(a, omega, loglength);
        best_fft<P>(b, omega, loglength);
        std::vector<F> prod(padded);
        for (size_t i = 0; i < padded; ++i) prod[i] = a[i] * b[i] * scaling;
        best_fft<P>(prod, omega_inv, loglength);
        prod.resize(length);
        return Polynomial(std::move(prod));
    }
    Polynomial shr(size_t k) const {  // :167-176  (multiply by x^k)
        std::vector<F> r(k, F::zero());
        r.insert(r.end(), poly.begin(), poly.end());
        return Polynomial(std::move(r));
    }
    Polynomial operator+(const Polynomial& o) const {  // :178-195
        size_t n = std::max(poly.size(), o.poly.size());
        std::vector<F> r(n);
        for (size_t i = 0; i < n; ++i)
            r[i] = (i < poly.size() ? poly[i] : F::zero()) + (i < o.poly.size() ? o.poly[i] : F::zero());
        return Polynomial(std::move(r));
    }
    Polynomial operator*(const Polynomial& o) const {  // :209-216
        if (poly.size() < 32 || o.poly.size() < 32) return mul_naive(*this, o);
        return mul_fft(o);
    }
};

// ---------------------------------------------------------------------------------------------
// RegularFunction a(x) + y b(x)  (reference: src/regular_functions_utils.rs:220-273)
// ---------------------------------------------------------------------------------------------
template <class C>
struct RegularFunction {
    typedef typename C::BaseP BP;
    typedef Fe<BP> F;
    Polynomial<BP> a, b;
    RegularFunction() {}
    RegularFunction(Polynomial<BP> a_, Polynomial<BP> b_) : a(std::move(a_)), b(std::move(b_)) {}

    static RegularFunction from_const(const F& x) { return RegularFunction(Polynomial<BP>({x}), Polynomial<BP>()); }  // :239
    static RegularFunction from_line(const F& ca, const F& cb, const F& cc) {  // :244  a*x + b*y + c
        return RegularFunction(Polynomial<BP>({cc, ca}), Polynomial<BP>({cb}));
    }
    F ev_unchecked(const F& x, const F& y) const { return a.ev(x) + b.ev(x) * y; }  // :235
    F ev(const Point<C>& pt) const {  // :228-233
        F zinv = pt.z.invert(), zinvsq = zinv * zinv;
        return ev_unchecked(pt.x * zinvsq, pt.y * zinvsq * zinv);
    }
    RegularFunction scale(const F& sc) const { return RegularFunction(a.scale(sc), b.scale(sc)); }  // :252
    RegularFunction operator+(const RegularFunction& o) const { return RegularFunction(a + o.a, b + o.b); }  // :257
    RegularFunction operator*(const RegularFunction& o) const {  // :266-273
        Polynomial<BP> subst_y2({C::b(), C::a(), F::zero(), F::one()});
        return RegularFunction((a * o.a) + ((b * o.b) * subst_y2), (a * o.b) + (b * o.a));
    }
};

// projective (xz, y, z^3) from Jacobian  (reference: :426-431)
template <class C>
void projective_coords(const Point<C>& p, typename Point<C>::F& px, typename Point<C>::F& py, typename Point<C>::F& pz) {
    typename Point<C>::F zsq = p.z * p.z;
    px = p.x * p.z; py = p.y; pz = p.z * zsq;
}

// line through two points, tangent fallback via -(a+b)  (reference: :285-303).
// Oracle convention: both arguments are z = 1 normalised (identity = (0,0,0)).
template <class C>
RegularFunction<C> linefunc(const Point<C>& a, const Point<C>& b) {
    typedef typename Point<C>::F F;
    F ax, ay, az, bx, by, bz;
    projective_coords(a, ax, ay, az);
    projective_coords(b, bx, by, bz);
    F lz = ax * by - ay * bx, lx = ay * bz - az * by, ly = az * bx - ax * bz;
    if (!lx.is_zero() || !ly.is_zero() || !lz.is_zero()) return RegularFunction<C>::from_line(lx, ly, lz);
    Point<C> c = (-(a + b)).normalized();
    F cx, cy, cz;
    projective_coords(c, cx, cy, cz);
    return RegularFunction<C>::from_line(ay * cz - az * cy, az * cx - ax * cz, ax * cy - ay * cx);
}

// Propagation (reference: :305-408).  The `inputs` vector the reference carries is never read on
// the path, so it is not materialised.  `output` is kept z = 1 normalised.
template <class C>
struct Propagation {
    typedef typename Point<C>::F F;
    Point<C> output;
    RegularFunction<C> wtns;

    static Propagation empty() {  // :324-326
        Propagation p; p.output = Point<C>::identity();
        p.wtns = RegularFunction<C>::from_const(F::one());
        return p;
    }
    static Propagation from_point(const Point<C>& pt) {  // :319-322
        if (pt.is_identity()) return empty();
        Propagation p; Point<C> n = pt.normalized();
        p.output = -n; p.wtns = linefunc<C>(n, -n);
        return p;
    }
    static Propagation from_pair(const Point<C>& pt1, const Point<C>& pt2) {  // :328-331
        if (pt1.is_identity()) return from_point(pt2);
        Propagation p; Point<C> n1 = pt1.normalized(), n2 = pt2.normalized();
        p.output = (-(n1 + n2)).normalized(); p.wtns = linefunc<C>(n1, n2);
        return p;
    }
    static Propagation merge(const Propagation& a, const Propagation& b) {  // :333-360
        Propagation r;
        r.output = (a.output + b.output).normalized();
        if (a.output.is_identity() || b.output.is_identity()) {  // :340-342
            r.wtns = a.wtns * b.wtns;
            return r;
        }
        RegularFunction<C> numerator = a.wtns * (b.wtns * linefunc<C>(-a.output, -b.output));  // :344
        F ax = a.output.x, bx = b.output.x;  // outputs are z = 1, so x/z^2 = x  (:351-355)
        r.wtns = RegularFunction<C>(numerator.a.kate_div(ax).kate_div(bx), numerator.b.kate_div(ax).kate_div(bx));  // :357
        return r;
    }
    // level-by-level sequential pairing, odd tail passes through  (:370-405)
    static Propagation group_merge(std::vector<Propagation> arr) {
        if (arr.empty()) throw std::runtime_error("group_merge of empty list");  // :382
        while (arr.size() > 1) {
            size_t pairs = (arr.size() + 1) / 2;
            std::vector<Propagation> next(pairs);
            auto body = [&](size_t lo, size_t hi) {
                for (size_t k = lo; k < hi; ++k)
                    next[k] = (2 * k + 1 < arr.size()) ? merge(arr[2 * k], arr[2 * k + 1]) : arr[2 * k];
            };
            // reference parallelises across the pairs of one level (:391); the FFTs inside a merge
            // are threaded by best_fft when the level has too few pairs to fill the pool
            if (pairs >= (size_t)Pool::instance().threads()) Pool::instance().parallel_for(pairs, body);
            else body(0, pairs);
            arr.swap(next);
        }
        return arr[0];
    }
};

// reference: :453-467
template <class C>
std::pair<RegularFunction<C>, Point<C>> compute_divisor_witness_partial(const std::vector<Point<C>>& pts) {
    typedef typename Point<C>::F F;
    if (pts.empty()) return {RegularFunction<C>::from_const(F::one()), Point<C>::identity()};
    std::vector<Propagation<C>> tmp((pts.size() + 1) / 2);
    auto body = [&](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k)
            tmp[k] = (2 * k + 1 < pts.size()) ? Propagation<C>::from_pair(pts[2 * k], pts[2 * k + 1])
                                              : Propagation<C>::from_point(pts[2 * k]);
    };
    Pool::instance().parallel_for(tmp.size(), body);
    Propagation<C> ret = Propagation<C>::group_merge(std::move(tmp));
    return {ret.wtns, ret.output};
}

// reference: :476-480  (panics when the points do not sum to the identity)
template <class C>
RegularFunction<C> compute_divisor_witness(const std::vector<Point<C>>& pts) {
    auto tmp = compute_divisor_witness_partial<C>(pts);
    if (!tmp.second.is_identity()) throw std::runtime_error("compute_divisor_witness: points do not sum to identity");
    return tmp.first;
}

// compute_divisor_witness_naive  (reference: :483-551): the witness as an arrangement of numerator and denominator lines.
// Lines are kept as (lx, ly, lz) triples (RegularFunction::from_line(lx, ly, lz): a = [lz, lx], b = [ly]).  Oracle convention
// as everywhere: every point handed to linefunc is z = 1 normalised.
template <class C>
struct Arrangement {
    typedef typename Point<C>::F F;
    struct Line { F lx, ly, lz; };
    std::vector<Line> pos, neg;
};
template <class C>
Arrangement<C> compute_divisor_witness_naive(const std::vector<Point<C>>& pts) {
    typedef typename Arrangement<C>::Line Line;
    std::vector<Point<C>> pos(pts.size()), neg;
    for (size_t i = 0; i < pts.size(); ++i) pos[i] = pts[i].normalized();
    Arrangement<C> ret;
    struct Glue { Point<C> a, b, sum; Line line; };
    std::vector<Glue> tmp;
    auto f = [&](std::vector<Glue>& v) {  // :497-504  q = a + b; Out(linefunc(a, b), -q)
        Pool::instance().parallel_for(v.size(), [&](size_t lo, size_t hi) {
            for (size_t k = lo; k < hi; ++k) {
                RegularFunction<C> l = linefunc<C>(v[k].a, v[k].b);
                v[k].line.lz = l.a.poly[0]; v[k].line.lx = l.a.poly[1]; v[k].line.ly = l.b.poly[0];
                v[k].sum = (-(v[k].a + v[k].b)).normalized();
            }
        });
    };
    auto drain = [&](std::vector<Point<C>>& from) {  // :515-520 / :531-536
        while (from.size() > 1) {
            Point<C> inc1 = from.back(); from.pop_back();
            if (!inc1.is_identity()) { Glue g; g.a = inc1; g.b = from.back(); from.pop_back(); tmp.push_back(g); }
        }
    };
    while (pos.size() > 1 || neg.size() > 1) {  // :513
        drain(pos);
        f(tmp);
        while (!tmp.empty()) { ret.pos.push_back(tmp.back().line); neg.push_back(tmp.back().sum); tmp.pop_back(); }  // :523-529
        drain(neg);
        f(tmp);
        while (!tmp.empty()) { ret.neg.push_back(tmp.back().line); pos.push_back(tmp.back().sum); tmp.pop_back(); }  // :539-545
    }
    bool ok = (pos.empty() && neg.empty()) || (pos.size() == 1 && neg.empty() && pos[0].is_identity()) ||
              (pos.empty() && neg.size() == 1 && neg[0].is_identity()) ||
              (pos.size() == 1 && neg.size() == 1 && pos[0] == neg[0]);  // :549-553
    if (!ok) throw std::runtime_error("compute_divisor_witness_naive: points do not sum to identity");
    return ret;
}

// Canonical form: trailing zeros stripped, then both polynomials divided by
// the coefficient of the term of highest pole order (x^i has order 2i, y x^i has order 2i+3).
template <class C>
RegularFunction<C> canonicalize(const RegularFunction<C>& f) {
    typedef typename Point<C>::F F;
    std::vector<F> a = f.a.poly, b = f.b.poly;
    while (!a.empty() && a.back().is_zero()) a.pop_back();
    while (!b.empty() && b.back().is_zero()) b.pop_back();
    if (a.empty() && b.empty()) return RegularFunction<C>();
    long oa = a.empty() ? -1 : 2 * (long)(a.size() - 1), ob = b.empty() ? -1 : 2 * (long)(b.size() - 1) + 3;
    F lead = oa > ob ? a.back() : b.back();
    F inv = lead.invert();
    for (auto& c : a) c = c * inv;
    for (auto& c : b) c = c * inv;
    return RegularFunction<C>(Polynomial<typename C::BaseP>(a), Polynomial<typename C::BaseP>(b));
}

}  // namespace oracle
