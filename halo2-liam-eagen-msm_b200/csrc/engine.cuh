// Host-side orchestration of the witness path on one B200: memory plan, launch order, error mapping.
// Templated on the curve; instantiated once per curve in engine_<curve>.cu, used through IEngine by capi.cu.
//
// Data flow (all device resident, one stream):
//   scalars --K1--> digit planes (d x n)                                        HBM bound
//   points  --K2--> multiples table (n x (b-1) affine)      one batched inversion
//   planes+table --K3--> d partial sums --K4--> d carries (affine)
//   planes+table+carries --scatter--> T_i for a group of digit positions (tmp of the reference)
//   T_i --K5--> leaves --[K5 pair sums, K6 NTT, K7 merge, K9 inversions] x levels--> root (a, b) --K10--> canonical
#pragma once
#include <algorithm>
#include <chrono>
#include <memory>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/eagen_msm.h"
#include "kernels.cuh"
#include "comm.cuh"

namespace eagen {

struct CudaError {
    std::string msg;
    int code;
};

#define EAGEN_CUDA(call)                                                                                        \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            throw CudaError{std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")", EAGEN_E_CUDA}; \
    } while (0)

struct StatusError {
    int code;
    std::string msg;
};

#define EAGEN_NCCL(call)                                                                                        \
    do {                                                                                                        \
        ncclResult_t r_ = (call);                                                                               \
        if (r_ != ncclSuccess)                                                                                  \
            throw StatusError{EAGEN_E_NCCL, std::string(#call) + ": " + (NcclApi::get().GetErrorString ? NcclApi::get().GetErrorString(r_) : "NCCL error")}; \
    } while (0)

// contiguous, balanced split of the d digit positions over the ranks (56 = 8 x 7 at base 5)
inline void position_range(int rank, int nranks, uint32_t d, uint32_t* begin, uint32_t* end) {
    uint32_t q = d / (uint32_t)nranks, r = d % (uint32_t)nranks;
    *begin = (uint32_t)rank * q + std::min<uint32_t>((uint32_t)rank, r);
    *end = *begin + q + ((uint32_t)rank < r ? 1u : 0u);
}

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* ensure(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFree(p);
            p = nullptr; cap = 0;
            size_t want = bytes + (bytes >> 3) + 256;
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess) { p = nullptr; throw CudaError{std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e), EAGEN_E_CUDA}; }
            cap = want;
        }
        return p;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    ~DevBuf() { release(); }
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; } return *this; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// Result buffers are recycled through a pool shared by the context and its results: cudaFree of a ~2 GB buffer was measured
// at up to 0.5 s on the B200 boxes (and is device-synchronising), which would dominate a 0.4 s witness call.
struct BufPool {
    std::mutex m;
    std::vector<DevBuf> bufs;
    DevBuf take(size_t bytes) {
        std::lock_guard<std::mutex> lk(m);
        int best = -1;
        for (size_t i = 0; i < bufs.size(); ++i)
            if (bufs[i].cap >= bytes && (best < 0 || bufs[i].cap < bufs[(size_t)best].cap)) best = (int)i;
        DevBuf b;
        if (best >= 0) { b = std::move(bufs[(size_t)best]); bufs.erase(bufs.begin() + best); }
        else {
            if (bufs.size() >= 8) bufs.erase(bufs.begin());  // bounded: drop the oldest (frees it)
            b.ensure(bytes);
        }
        return b;
    }
    void give(DevBuf&& b) {
        if (!b.p) return;
        std::lock_guard<std::mutex> lk(m);
        bufs.push_back(std::move(b));
    }
};

// Pinned host staging ring for the small per-call parameter tables (level counts, position maps, error flags): an asynchronous
// copy from PAGEABLE memory makes the driver synchronise the stream first, a copy from pinned memory does not.  Every public call
// ends with a stream synchronisation and stages far less than the ring holds, so a wrapped allocation never overlaps data in flight.
struct PinnedRing {
    char* p = nullptr;
    size_t cap = 0, off = 0;
    void init(size_t bytes) {
        if (cudaHostAlloc((void**)&p, bytes, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; throw CudaError{"cudaHostAlloc of the staging ring failed", EAGEN_E_CUDA}; }
        cap = bytes;
    }
    void* take(size_t bytes) {
        bytes = (bytes + 63) & ~(size_t)63;
        if (bytes > cap) throw CudaError{"staging ring too small", EAGEN_E_CUDA};
        if (off + bytes > cap) off = 0;
        void* r = p + off;
        off += bytes;
        return r;
    }
    ~PinnedRing() { if (p) cudaFreeHost(p); }
};

// ---- result handle -------------------------------------------------------------------------------------
struct ResultImpl {
    std::shared_ptr<BufPool> pool;       // buffers go back to the owning context's pool when the result is freed
    ~ResultImpl() { if (pool) { pool->give(std::move(A)); pool->give(std::move(B)); pool->give(std::move(digits)); } }
    void alloc(DevBuf& b, size_t bytes) {
        if (b.cap >= bytes) return;
        if (pool) { pool->give(std::move(b)); b = pool->take(bytes); } else b.ensure(bytes);
    }
    int device = 0;
    uint32_t d = 0;
    size_t n = 0;
    size_t nf = 0;                       // functions held
    size_t k0 = 0;                       // digit position (function index of the whole witness) of slot 0: non-zero for a rank's share
    size_t a_stride = 0, b_stride = 0;   // elements
    DevBuf A, B;                         // nf x stride field elements
    std::vector<int> la, lb;             // trimmed lengths
    DevBuf digits;                       // n x d (optional)
    bool has_digits = false;
    std::vector<uint64_t> carries;       // d x 8 (host copy, tiny)
    uint64_t carry[8] = {0};
    double device_ms = 0;
};

// host destination of a streamed result: function k at out + k*(a_stride + b_stride)*32 (a first, then b)
struct StreamOut {
    uint8_t* out = nullptr;
    size_t a_stride = 0, b_stride = 0;  // elements
};

// slot sizes of the streamed layout from an upper bound of the list length (n + base + 1 points per digit position)
inline void stream_slot_elems(size_t n, uint8_t base, size_t* a_stride, size_t* b_stride) {
    size_t nmax = n + base + 1, lc = (nmax + 1) / 2;
    int L = 0;
    while (((size_t)1 << L) < lc) ++L;
    *a_stride = ((size_t)1 << L) + 1;
    *b_stride = std::max<size_t>((size_t)1 << L, 1);
}

struct IEngine {
    virtual ~IEngine() {}
    virtual ResultImpl* lhs_stream_host(const uint64_t* scalars, const uint64_t* pts, size_t n, uint8_t base, uint32_t flags, void* out, size_t out_bytes) = 0;
    virtual int device() const = 0;
    virtual uint64_t launches() const = 0;
    virtual uint64_t iso_fallbacks() const = 0;
    virtual void negbase_host(const uint64_t* scalars, size_t n, uint8_t base, uint8_t* digits) = 0;
    virtual void multiples_host(const uint64_t* pts, size_t n, uint8_t base, uint64_t* out) = 0;
    virtual ResultImpl* lhs_host(const uint64_t* scalars, const uint64_t* pts, size_t n, uint8_t base, uint32_t flags) = 0;
    virtual ResultImpl* lhs_dev(const void* d_scalars, const void* d_pts, size_t n, uint8_t base, uint32_t flags) = 0;
    virtual ResultImpl* divisor_host(const uint64_t* pts, size_t n, uint32_t flags, uint64_t* out_point) = 0;
    virtual void poly_mul_host(const uint64_t* a, size_t la, const uint64_t* b, size_t lb, uint64_t* out) = 0;
    virtual void ntt_host(uint64_t* data, uint32_t log_n, int inverse) = 0;
    virtual void batch_invert_host(uint64_t* elems, size_t n) = 0;
    virtual void eval_host(const uint64_t* a, size_t la, const uint64_t* b, size_t lb, const uint64_t* pts, size_t n, uint64_t* out) = 0;
    virtual void shard_sums_dev(const void* d_scalars, const void* d_pts, size_t n, uint8_t base, void* d_planes, void* d_table, void* d_sums) = 0;
    virtual void carry_chain_dev(const void* d_sums, int nparts, uint8_t base, void* d_carries) = 0;
    virtual double negbase_dev(const void* d_scalars, size_t n, uint8_t base, void* d_planes, void* d_rows) = 0;
    virtual double ntt_dev(void* d_data, uint32_t log_n, size_t batch, int inverse) = 0;
    virtual double msm_host(const uint64_t* scalars, const uint64_t* pts, size_t n, uint64_t* out_affine) = 0;
    virtual void scalar_witness_host(const uint64_t* scalars, size_t n, uint8_t base, uint32_t num_digits, uint32_t logtable, int mode, void* out) = 0;
    virtual void naive_host(const uint64_t* pts, size_t n, uint64_t* pos, size_t* n_pos, uint64_t* neg, size_t* n_neg) = 0;
    virtual void result_eval_host(ResultImpl* r, const uint64_t* pts, size_t m, uint64_t* out) = 0;
    virtual double microbench(int which) = 0;
    virtual void comm_attach(void* nccl_comm, int nranks, int rank) = 0;
    virtual void comm_init_rank(int nranks, int rank, const void* unique_id) = 0;
    virtual void comm_destroy() = 0;
    virtual int comm_size() const = 0;
    virtual int comm_rank() const = 0;
    virtual ResultImpl* lhs_sharded(const void* scalars, const void* pts, bool device_inputs, size_t n_local, uint8_t base, uint32_t flags,
                                    void* out, size_t out_bytes) = 0;
    virtual void set_stream_split(const uint32_t* pct, int n) = 0;
    virtual void set_profiling(int mode) = 0;
    virtual std::string profile_json() = 0;
    virtual void profile_reset() = 0;
    virtual void synth_dev(uint64_t seed, size_t n, void* d_scalars, void* d_pts) = 0;
    virtual void synth_host(uint64_t seed, size_t n, uint64_t* scalars, uint64_t* pts) = 0;
    virtual ResultImpl* trees_dev(const void* d_planes, const void* d_table, const void* d_carries, size_t n, uint8_t base,
                                  uint32_t pos_begin, uint32_t pos_end, uint32_t flags) = 0;
};

// ---- small host big-integer helpers (order / isqrt / logb_ceil of the reference, host side) ---------------
struct HostU256 {
    uint32_t w[8];
    static HostU256 zero() { HostU256 r; std::memset(r.w, 0, 32); return r; }
    bool is_zero() const { for (int i = 0; i < 8; ++i) if (w[i]) return false; return true; }
    uint32_t divmod_small(uint32_t dv) {
        uint64_t rem = 0;
        for (int i = 7; i >= 0; --i) { uint64_t cur = (rem << 32) | w[i]; w[i] = (uint32_t)(cur / dv); rem = cur % dv; }
        return (uint32_t)rem;
    }
    bool mul_small_add(uint32_t m, uint32_t addv) {  // this = this*m + addv ; false on overflow
        uint64_t c = addv;
        for (int i = 0; i < 8; ++i) { c += (uint64_t)w[i] * m; w[i] = (uint32_t)c; c >>= 32; }
        return c == 0;
    }
    int cmp(const HostU256& o) const {
        for (int i = 7; i >= 0; --i) if (w[i] != o.w[i]) return w[i] < o.w[i] ? -1 : 1;
        return 0;
    }
};

// floor(sqrt(p)), bit-by-bit on a 128-bit candidate
inline HostU256 host_isqrt(const HostU256& p) {
    HostU256 r = HostU256::zero();
    for (int bit = 127; bit >= 0; --bit) {
        HostU256 c = r;
        c.w[bit / 32] |= 1u << (bit % 32);
        uint32_t sq[9] = {0};  // c is 4 limbs -> c^2 is 8 limbs
        for (int i = 0; i < 4; ++i) {
            uint64_t carry = 0;
            for (int j = 0; j < 4; ++j) { carry += (uint64_t)c.w[i] * c.w[j] + sq[i + j]; sq[i + j] = (uint32_t)carry; carry >>= 32; }
            sq[i + 4] += (uint32_t)carry;
        }
        HostU256 s; std::memcpy(s.w, sq, 32);
        if (s.cmp(p) <= 0) r = c;
    }
    return r;
}

template <class FS>
inline HostU256 host_order() { HostU256 r; for (int i = 0; i < 8; ++i) r.w[i] = FS::mod(i); return r; }

// d and the K1 constants for (scalar field, base)      reference: src/argument_witness_calc.rs:89-91
template <class FS>
inline NegbaseParams make_negbase_params(uint8_t base) {
    NegbaseParams p;
    HostU256 sq = host_isqrt(host_order<FS>());
    sq.mul_small_add(1, 2);
    std::memcpy(p.sq, sq.w, 32);
    uint32_t d = 0;
    for (HostU256 x = sq; !x.is_zero(); x.divmod_small(base)) ++d;
    p.d = d + 1;
    p.base = base;
    HostU256 K = HostU256::zero(), pw = HostU256::zero();
    pw.w[0] = 1;  // base^i
    for (uint32_t i = 0; i < p.d; ++i) {
        if (i & 1) {  // K += (base-1) * base^i
            uint64_t c = 0;
            for (int l = 0; l < 8; ++l) { c += (uint64_t)pw.w[l] * (base - 1) + K.w[l]; K.w[l] = (uint32_t)c; c >>= 32; }
        }
        pw.mul_small_add(base, 0);
    }
    std::memcpy(p.K, K.w, 32);
    std::memcpy(p.bd, pw.w, 32);
    // b^0 .. b^4 and the table group size
    uint64_t pwv = 1;
    for (int k = 0; k < 5; ++k) { p.pw[k] = (uint32_t)pwv; pwv *= base; }
    p.g = base <= 5 ? 4 : (base <= 31 ? 2 : 1);
    p.lut_n = p.g == 1 ? 0 : p.pw[p.g];
    p.nw = (p.d + 3) / 4;
    p.pad = 4 * p.nw - p.d;
    if (p.nw > (uint32_t)NEGBASE_MAX_WORDS) throw StatusError{EAGEN_E_ARG, "negbase: more than 144 digits"};
    // The 160-bit fixed-point image of y / b^d is exact for every digit iff 2^-160 + b^d 2^-288 <= b^-d; b^d < 2^143 suffices.
    for (int l = 5; l < 8; ++l) if (pw.w[l]) throw StatusError{EAGEN_E_ARG, "negbase: b^d does not fit 143 bits"};
    if (pw.w[4] >> 15) throw StatusError{EAGEN_E_ARG, "negbase: b^d does not fit 143 bits"};
    // inv = ceil(2^288 / b^d): d nested floor divisions by b (floor(floor(x/a)/b) = floor(x/(ab))), +1 unless all were exact
    uint32_t q[10] = {0};
    q[9] = 1;  // 2^288
    bool exact = true;
    for (uint32_t i = 0; i < p.d; ++i) {
        uint64_t rem = 0;
        for (int l = 9; l >= 0; --l) { uint64_t cur = (rem << 32) | q[l]; q[l] = (uint32_t)(cur / base); rem = cur % base; }
        if (rem) exact = false;
    }
    if (!exact) { for (int l = 0; l < 10; ++l) { if (++q[l] != 0) break; } }
    for (int l = 6; l < 10; ++l) if (q[l]) throw StatusError{EAGEN_E_ARG, "negbase: reciprocal does not fit 192 bits"};
    if (q[5] > 1 || (q[5] == 1 && (q[0] | q[1] | q[2] | q[3] | q[4]))) throw StatusError{EAGEN_E_ARG, "negbase: b^d < 2^128"};
    std::memcpy(p.inv, q, 24);
    return p;
}

inline int ceil_log2(size_t x) { int l = 0; while (((size_t)1 << l) < x) ++l; return l; }

// forward order (descending stages) of shared-memory passes for a 2^t transform
inline std::vector<std::pair<int, int>> ntt_plan(int t) {
    std::vector<std::pair<int, int>> v;
    if (t <= NTT_TILE_LOG) { v.push_back({t - 1, 0}); return v; }
    int r = t - NTT_TILE_LOG, np = (r + 7) / 8, each = r / np, extra = r % np, s = t - 1;
    for (int p = 0; p < np; ++p) { int k = each + (p < extra ? 1 : 0); v.push_back({s, s - k + 1}); s -= k; }
    v.push_back({NTT_TILE_LOG - 1, 0});
    return v;
}

template <class CC>
class Engine : public IEngine {
public:
    typedef typename CC::Base FB;
    typedef typename CC::Scalar FS;
    typedef Fe<FB> F;
    typedef Affine<FB> Aff;
    typedef Proj<FB> Prj;

    explicit Engine(int dev) : dev_(dev) {
        EAGEN_CUDA(cudaSetDevice(dev_));
        EAGEN_CUDA(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, dev_));
        EAGEN_CUDA(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
        EAGEN_CUDA(cudaStreamCreateWithFlags(&cst_, cudaStreamNonBlocking));
        {   // the side stream produces what the main stream waits for (descriptors, inverse denominators): highest priority, so its
            // blocks are placed as soon as SM resources free up instead of queueing behind the main stream's multi-millisecond grids
            int lo = 0, hi = 0;
            EAGEN_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            EAGEN_CUDA(cudaStreamCreateWithPriority(&pst_, cudaStreamNonBlocking, hi));
        }
        ls_ = st_;
        ring_.init((size_t)8 << 20);
        EAGEN_CUDA(cudaMalloc(&d_err_, sizeof(int)));
        EAGEN_CUDA(cudaMemsetAsync(d_err_, 0, sizeof(int), st_));
        EAGEN_CUDA(cudaEventCreate(&ev0_));
        EAGEN_CUDA(cudaEventCreate(&ev1_));
        int one = 1;
        EAGEN_CUDA(cudaMalloc(&d_one_, sizeof(int)));
        EAGEN_CUDA(cudaMemcpyAsync(d_one_, &one, sizeof(int), cudaMemcpyHostToDevice, st_));
        EAGEN_CUDA(cudaStreamSynchronize(st_));
    }
    ~Engine() override {
        cudaSetDevice(dev_);
        cudaStreamSynchronize(st_);
        if (comm_ && comm_owned_ && NcclApi::get().ok) NcclApi::get().CommDestroy(comm_);
        if (nst_) cudaStreamDestroy(nst_);
        cudaFree(d_err_); cudaFree(d_one_);
        cudaEventDestroy(ev0_); cudaEventDestroy(ev1_);
        for (cudaEvent_t e : sync_events_) cudaEventDestroy(e);
        cudaStreamDestroy(cst_);
        cudaStreamDestroy(pst_);
        cudaStreamDestroy(st_);
    }
    int device() const override { return dev_; }
    uint64_t launches() const override { return launches_; }
    uint64_t iso_fallbacks() const override { return iso_fallbacks_; }

    // ---- per-kernel-group profiling: CUDA events on the launching stream + exact work counts from launch parameters
    void set_profiling(int mode) override { prof_on_ = mode != 0; prof_detail_ = mode == 2; }
    void profile_reset() override { for (auto& e : prof_) { e.launches = 0; e.scopes = 0; e.ms = 0; e.bytes = 0; e.modmul = 0; } }
    std::string profile_json() override {
        prof_collect();
        std::string o = "[";
        bool first = true;
        for (auto& e : prof_) {
            if (!e.scopes) continue;
            char buf[512];
            snprintf(buf, sizeof buf, "%s{\"kernel\": \"%s\", \"launches\": %llu, \"scopes\": %llu, \"ms\": %.6f, \"bytes\": %.0f, \"modmul\": %.0f}",
                     first ? "" : ", ", e.name.c_str(), (unsigned long long)e.launches, (unsigned long long)e.scopes, e.ms, e.bytes, e.modmul);
            o += buf; first = false;
        }
        return o + "]";
    }

    // ------------------------------------------------------------------------------------------------
    // public operations
    // ------------------------------------------------------------------------------------------------
    void negbase_host(const uint64_t* scalars, size_t n, uint8_t base, uint8_t* digits) override {
        use();
        NegbaseParams prm = make_negbase_params<FS>(base);
        if (n == 0) return;
        Fe<FS>* ds = (Fe<FS>*)in_scalars_.ensure(n * 32);
        uint8_t* planes = (uint8_t*)planes_.ensure(n * prm.d);
        uint8_t* rows = (uint8_t*)rows_.ensure(n * prm.d);
        EAGEN_CUDA(cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, st_));
        run_negbase(ds, n, prm, planes, rows);
        EAGEN_CUDA(cudaMemcpyAsync(digits, rows, n * prm.d, cudaMemcpyDeviceToHost, st_));
        sync_check();
    }

    void multiples_host(const uint64_t* pts, size_t n, uint8_t base, uint64_t* out) override {
        use();
        if (n == 0) return;
        F* dp = (F*)in_points_.ensure(n * 96);
        Aff* tab = (Aff*)table_.ensure(n * (size_t)(base - 1) * 64);
        EAGEN_CUDA(cudaMemcpyAsync(dp, pts, n * 96, cudaMemcpyHostToDevice, st_));
        run_multiples(dp, n, base, tab);
        EAGEN_CUDA(cudaMemcpyAsync(out, tab, n * (size_t)(base - 1) * 64, cudaMemcpyDeviceToHost, st_));
        sync_check();
    }

    ResultImpl* lhs_host(const uint64_t* scalars, const uint64_t* pts, size_t n, uint8_t base, uint32_t flags) override {
        use();
        Fe<FS>* ds = (Fe<FS>*)in_scalars_.ensure(std::max<size_t>(n, 1) * 32);
        F* dp = (F*)in_points_.ensure(std::max<size_t>(n, 1) * 96);
        if (n) {
            EAGEN_CUDA(cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, st_));
            EAGEN_CUDA(cudaMemcpyAsync(dp, pts, n * 96, cudaMemcpyHostToDevice, st_));
        }
        return lhs_dev(ds, dp, n, base, flags);
    }

    // compute_lhs_witness from host buffers with the functions copied into `out` while later digit positions are still being
    // computed (the D2H of 1.5 GB at 2^20 otherwise adds ~25 % to the call)
    ResultImpl* lhs_stream_host(const uint64_t* scalars, const uint64_t* pts, size_t n, uint8_t base, uint32_t flags, void* out, size_t out_bytes) override {
        use();
        NegbaseParams prm = make_negbase_params<FS>(base);
        StreamOut so;
        so.out = (uint8_t*)out;
        stream_slot_elems(n, base, &so.a_stride, &so.b_stride);
        if (out_bytes < (size_t)prm.d * (so.a_stride + so.b_stride) * 32) throw StatusError{EAGEN_E_LEN, "streamed output buffer too small"};
        Fe<FS>* ds = (Fe<FS>*)in_scalars_.ensure(std::max<size_t>(n, 1) * 32);
        F* dp = (F*)in_points_.ensure(std::max<size_t>(n, 1) * 96);
        if (n) {
            EAGEN_CUDA(cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, st_));
            EAGEN_CUDA(cudaMemcpyAsync(dp, pts, n * 96, cudaMemcpyHostToDevice, st_));
        }
        return lhs_dev_impl(ds, dp, n, base, flags, &so);
    }

    ResultImpl* lhs_dev(const void* d_scalars, const void* d_pts, size_t n, uint8_t base, uint32_t flags) override {
        return lhs_dev_impl(d_scalars, d_pts, n, base, flags, nullptr);
    }

    ResultImpl* lhs_dev_impl(const void* d_scalars, const void* d_pts, size_t n, uint8_t base, uint32_t flags, const StreamOut* so) {
        use();
        NegbaseParams prm = make_negbase_params<FS>(base);
        const uint32_t d = prm.d;
        EAGEN_CUDA(cudaEventRecord(ev0_, st_));
        size_t nn = std::max<size_t>(n, 1);
        uint8_t* planes = (uint8_t*)planes_.ensure(nn * d);
        Aff* tab = (Aff*)table_.ensure(nn * (size_t)(base - 1) * 64);
        Prj* sums = (Prj*)sums_.ensure((size_t)d * sizeof(Prj));
        Aff* carries = (Aff*)carries_.ensure((size_t)d * sizeof(Aff));
        uint8_t* rows = nullptr;
        std::unique_ptr<ResultImpl> res(new ResultImpl());
        res->device = dev_; res->pool = pool_; res->d = d; res->n = n;
        if (flags & EAGEN_KEEP_DIGITS) { res->alloc(res->digits, nn * d); rows = (uint8_t*)res->digits.p; res->has_digits = true; }
        run_shard_sums((const Fe<FS>*)d_scalars, (const F*)d_pts, n, prm, planes, rows, tab, sums);
        run_carry_chain(sums, 1, d, base, carries);
        res->carries.assign((size_t)d * 8, 0);
        uint64_t* hcar = (uint64_t*)ring_.take((size_t)d * 64);   // pinned staging: a copy into pageable memory would block the host here
        EAGEN_CUDA(cudaMemcpyAsync(hcar, carries, (size_t)d * 64, cudaMemcpyDeviceToHost, st_));
        if (!(flags & EAGEN_NO_FUNCTIONS)) run_position_trees(planes, tab, carries, n, base, d, 0, d, flags, res.get(), so);
        EAGEN_CUDA(cudaEventRecord(ev1_, st_));
        sync_check();
        std::memcpy(res->carries.data(), hcar, (size_t)d * 64);
        std::memcpy(res->carry, &res->carries[(size_t)(d - 1) * 8], 64);
        float ms = 0; EAGEN_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
        res->device_ms = ms;
        return res.release();
    }

    void shard_sums_dev(const void* d_scalars, const void* d_pts, size_t n, uint8_t base, void* d_planes, void* d_table, void* d_sums) override {
        use();
        NegbaseParams prm = make_negbase_params<FS>(base);
        run_shard_sums((const Fe<FS>*)d_scalars, (const F*)d_pts, n, prm, (uint8_t*)d_planes, nullptr, (Aff*)d_table, (Prj*)d_sums);
        sync_check();
    }
    // stand-alone K1 on device buffers; returns the device milliseconds (CUDA events on the launching stream)
    double negbase_dev(const void* d_scalars, size_t n, uint8_t base, void* d_planes, void* d_rows) override {
        use();
        NegbaseParams prm = make_negbase_params<FS>(base);
        EAGEN_CUDA(cudaEventRecord(ev0_, st_));
        if (n) run_negbase((const Fe<FS>*)d_scalars, n, prm, (uint8_t*)d_planes, (uint8_t*)d_rows);
        EAGEN_CUDA(cudaEventRecord(ev1_, st_));
        sync_check();
        float ms = 0; EAGEN_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
        return ms;
    }
    // stand-alone K6: `batch` transforms of 2^log_n elements in place (forward: natural -> bit-reversed, inverse: bit-reversed ->
    // natural, unscaled); returns device milliseconds
    double ntt_dev(void* d_data, uint32_t log_n, size_t batch, int inverse) override {
        use();
        if (log_n == 0 || batch == 0) return 0.0;
        if (log_n > FB::S) throw StatusError{EAGEN_E_NTT_TOO_LARGE, "log_n exceeds the field's two-adicity"};
        ensure_twiddles((int)log_n);
        int* cnt = (int*)tops_.ensure(sizeof(int));
        int one_batch = (int)batch;
        EAGEN_CUDA(cudaMemcpyAsync(cnt, &one_batch, sizeof(int), cudaMemcpyHostToDevice, st_));
        EAGEN_CUDA(cudaStreamSynchronize(st_));
        EAGEN_CUDA(cudaEventRecord(ev0_, st_));
        ntt(inverse != 0, (F*)d_data, nullptr, 0, 0, nullptr, 0, 0, (int)log_n, batch, cnt, (int)batch);
        EAGEN_CUDA(cudaEventRecord(ev1_, st_));
        sync_check();
        float ms = 0; EAGEN_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
        return ms;
    }
    void carry_chain_dev(const void* d_sums, int nparts, uint8_t base, void* d_carries) override {
        use();
        NegbaseParams prm = make_negbase_params<FS>(base);
        run_carry_chain((const Prj*)d_sums, nparts, prm.d, base, (Aff*)d_carries);
        sync_check();
    }
    ResultImpl* trees_dev(const void* d_planes, const void* d_table, const void* d_carries, size_t n, uint8_t base,
                          uint32_t pos_begin, uint32_t pos_end, uint32_t flags) override {
        use();
        NegbaseParams prm = make_negbase_params<FS>(base);
        if (pos_begin > pos_end || pos_end > prm.d) throw StatusError{EAGEN_E_ARG, "digit position range out of bounds"};
        std::unique_ptr<ResultImpl> res(new ResultImpl());
        res->device = dev_; res->pool = pool_; res->d = prm.d; res->n = n;
        EAGEN_CUDA(cudaEventRecord(ev0_, st_));
        res->carries.assign((size_t)prm.d * 8, 0);
        EAGEN_CUDA(cudaMemcpyAsync(res->carries.data(), d_carries, (size_t)prm.d * 64, cudaMemcpyDeviceToHost, st_));
        run_position_trees((const uint8_t*)d_planes, (const Aff*)d_table, (const Aff*)d_carries, n, base, prm.d, pos_begin, pos_end, flags, res.get());
        EAGEN_CUDA(cudaEventRecord(ev1_, st_));
        sync_check();
        std::memcpy(res->carry, &res->carries[(size_t)(prm.d - 1) * 8], 64);
        float ms = 0; EAGEN_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
        res->device_ms = ms;
        return res.release();
    }

    // best_multiexp equivalent: sum s_j P_j for full-width scalars (Montgomery, scalar field) and Jacobian points; returns the
    // device milliseconds of the compute part
    double msm_host(const uint64_t* scalars, const uint64_t* pts, size_t n, uint64_t* out_affine) override {
        use();
        if (n == 0) { std::memset(out_affine, 0, 64); return 0.0; }
        if (n >= ((size_t)1 << 32)) throw StatusError{EAGEN_E_ARG, "eagen_msm: more than 2^32 points"};
        Fe<FS>* ds = (Fe<FS>*)in_scalars_.ensure(n * 32);
        F* dp = (F*)in_points_.ensure(n * 96);
        EAGEN_CUDA(cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, st_));
        EAGEN_CUDA(cudaMemcpyAsync(dp, pts, n * 96, cudaMemcpyHostToDevice, st_));
        EAGEN_CUDA(cudaEventRecord(ev0_, st_));
        Aff* A = (Aff*)tpts_.ensure(n * sizeof(Aff));
        F* zs = (F*)den_.ensure(std::max<size_t>(n, 64) * 32);
        launch(k_jac_z<FB>, n, 256, (const F*)dp, n, zs);
        batch_invert(zs, n);
        launch(k_jac_to_affine<FB>, n, 256, (const F*)dp, (const F*)zs, n, A);
        int chunks = (int)((n + MSM_CHUNK - 1) / MSM_CHUNK);
        uint8_t* dig = (uint8_t*)planes_.ensure((size_t)MSM_WINDOWS * n);
        uint32_t* hist = (uint32_t*)cnt_.ensure((size_t)MSM_WINDOWS * chunks * 256 * sizeof(uint32_t));
        uint32_t* bstart = (uint32_t*)tree_n_.ensure((size_t)MSM_WINDOWS * 257 * sizeof(uint32_t));
        uint32_t* sorted = (uint32_t*)ea_.ensure((size_t)MSM_WINDOWS * n * sizeof(uint32_t));
        Prj* buckets = (Prj*)partials_.ensure((size_t)MSM_WINDOWS * 256 * sizeof(Prj));
        Prj* wsum = (Prj*)sums_.ensure((size_t)(MSM_WINDOWS + 1) * sizeof(Prj));
        {
            Scope ps(this, "msm", (double)n * (32.0 + 96.0 + 64.0 + 32.0 * (1 + 4 + 4 + 64)), (double)n * (1.0 + 5.0 + 32.0 * 13.0));
            launch(k_msm_digits<FS>, n, 128, (const Fe<FS>*)ds, n, dig);
            launch2d(k_msm_hist, dim3(chunks, MSM_WINDOWS), 256, (const uint8_t*)dig, n, chunks, hist);
            launch2d(k_msm_scan, dim3(MSM_WINDOWS, 1), 256, hist, chunks, bstart);
            launch2d(k_msm_scatter, dim3(chunks, MSM_WINDOWS), 256, (const uint8_t*)dig, n, chunks, (const uint32_t*)hist, sorted);
            launch(k_msm_buckets<CC>, (size_t)MSM_WINDOWS * 255 * 32, 128, (const Aff*)A, n, (const uint32_t*)sorted, (const uint32_t*)bstart, buckets);
            launch(k_msm_window_sums<CC>, MSM_WINDOWS, 32, (const Prj*)buckets, wsum);
            launch(k_msm_combine<CC>, 1, 32, (const Prj*)wsum, wsum + MSM_WINDOWS, zs);
        }
        batch_invert(zs, 1);
        Aff* res = (Aff*)carries_.ensure(sizeof(Aff) * 64);
        launch(k_proj_to_affine<FB>, 1, 32, (const Prj*)(wsum + MSM_WINDOWS), (const F*)zs, (size_t)1, res);
        EAGEN_CUDA(cudaEventRecord(ev1_, st_));
        EAGEN_CUDA(cudaMemcpyAsync(out_affine, res, 64, cudaMemcpyDeviceToHost, st_));
        sync_check();
        float ms = 0; EAGEN_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
        return ms;
    }

    // ------------------------------------------------------------------------------------------------
    // multi-GPU: one context per rank / device, NCCL communicator owned by the context
    // ------------------------------------------------------------------------------------------------
    void comm_attach(void* nccl_comm, int nranks, int rank) override {   // communicator created by the caller (ncclCommInitAll): not destroyed here
        comm_destroy();
        comm_ = (ncclComm_t)nccl_comm; nranks_ = nranks; rank_ = rank; comm_owned_ = false;
        ensure_comm_stream();
    }
    void comm_init_rank(int nranks, int rank, const void* unique_id) override {
        use();
        NcclApi& nc = NcclApi::get();
        if (!nc.ok) throw StatusError{EAGEN_E_NCCL, nc.load_error};
        if (nranks < 1 || rank < 0 || rank >= nranks || !unique_id) throw StatusError{EAGEN_E_ARG, "eagen_comm_init: bad rank / size / id"};
        comm_destroy();
        ncclUniqueId id;
        std::memcpy(&id, unique_id, sizeof id);
        EAGEN_NCCL(nc.CommInitRank(&comm_, nranks, id, rank));
        nranks_ = nranks; rank_ = rank; comm_owned_ = true;
        ensure_comm_stream();
    }
    void comm_destroy() override {
        if (comm_ && comm_owned_) { use(); cudaStreamSynchronize(st_); if (nst_) cudaStreamSynchronize(nst_); NcclApi::get().CommDestroy(comm_); }
        comm_ = nullptr; nranks_ = 1; rank_ = 0; comm_owned_ = false;
    }
    int comm_size() const override { return nranks_; }
    int comm_rank() const override { return rank_; }
    void set_stream_split(const uint32_t* pct, int n) override {
        stream_split_.clear();
        for (int i = 0; i < n; ++i) if (pct[i] > 0) stream_split_.push_back(pct[i]);
        if (stream_split_.empty()) { stream_split_.push_back(75); stream_split_.push_back(25); }
    }

    // compute_lhs_witness over the point ranges of ALL ranks; this rank passes its own n_local scalars / points (every rank the same
    // n_local) and receives the functions of its share of the digit positions (position_range).  Stages:
    //   K1-K3 on the local range  ->  all-gather of the d x 96-byte partial digit sums, the digit planes (one grouped all-gather,
    //   landing position-major over the global point range) and the multiples table (the one real exchange: N x (b-1) x 64 B), all on
    //   the communication stream  ->  carry chain (replicated; overlaps the plane / table gathers)  ->  this rank's trees.
    // out != nullptr: host buffer the functions are streamed into (layout of eagen_lhs_witness_stream over the LOCAL slots).
    ResultImpl* lhs_sharded(const void* scalars, const void* pts, bool device_inputs, size_t n_local, uint8_t base, uint32_t flags,
                            void* out, size_t out_bytes) override {
        use();
        if (!comm_) throw StatusError{EAGEN_E_NCCL, "eagen_lhs_witness_sharded: no communicator (call eagen_comm_init first)"};
        NcclApi& nc = NcclApi::get();
        NegbaseParams prm = make_negbase_params<FS>(base);
        const uint32_t d = prm.d;
        const size_t W = (size_t)nranks_, n_total = n_local * W, nn = std::max<size_t>(n_local, 1);
        uint32_t pos0, pos1;
        position_range(rank_, nranks_, d, &pos0, &pos1);
        StreamOut so;
        if (out) {
            so.out = (uint8_t*)out;
            stream_slot_elems(n_total, base, &so.a_stride, &so.b_stride);
            if (out_bytes < (size_t)(pos1 - pos0) * (so.a_stride + so.b_stride) * 32) throw StatusError{EAGEN_E_LEN, "streamed output buffer too small"};
        }
        const Fe<FS>* ds = (const Fe<FS>*)scalars;
        const F* dp = (const F*)pts;
        if (!device_inputs) {
            Fe<FS>* hs = (Fe<FS>*)in_scalars_.ensure(nn * 32);
            F* hp = (F*)in_points_.ensure(nn * 96);
            if (n_local) {
                EAGEN_CUDA(cudaMemcpyAsync(hs, scalars, n_local * 32, cudaMemcpyHostToDevice, st_));
                EAGEN_CUDA(cudaMemcpyAsync(hp, pts, n_local * 96, cudaMemcpyHostToDevice, st_));
            }
            ds = hs; dp = hp;
        }
        EAGEN_CUDA(cudaEventRecord(ev0_, st_));
        uint8_t* pl_local = (uint8_t*)sh_planes_.ensure(nn * d);
        Aff* tab_local = (Aff*)sh_table_.ensure(nn * (size_t)(base - 1) * 64);
        // [0, d): this rank's partial sums; [d, d + W*d): all ranks'; the 8-byte slots after them carry n_local of every rank
        Prj* sums = (Prj*)sums_.ensure((size_t)(d + W * d) * sizeof(Prj) + (W + 1) * sizeof(unsigned long long));
        Prj* all_sums = sums + d;
        unsigned long long* nl = (unsigned long long*)(all_sums + W * d);   // nl[0]: mine, nl[1..W]: gathered
        uint8_t* planes = (uint8_t*)planes_.ensure(std::max<size_t>(n_total, 1) * d);
        Aff* tab = (Aff*)table_.ensure(std::max<size_t>(n_total, 1) * (size_t)(base - 1) * 64);
        Aff* carries = (Aff*)carries_.ensure((size_t)d * sizeof(Aff));
        std::unique_ptr<ResultImpl> res(new ResultImpl());
        res->device = dev_; res->pool = pool_; res->d = d; res->n = n_total; res->k0 = d - pos1;
        {
            unsigned long long* stage = (unsigned long long*)ring_.take(sizeof(unsigned long long));
            *stage = (unsigned long long)n_local;
            EAGEN_CUDA(cudaMemcpyAsync(nl, stage, sizeof(unsigned long long), cudaMemcpyHostToDevice, st_));
        }
        run_shard_sums(ds, dp, n_local, prm, pl_local, nullptr, tab_local, sums);
        // ---- the exchange, on the communication stream
        EAGEN_CUDA(cudaEventRecord(sync_event(100), st_));
        EAGEN_CUDA(cudaStreamWaitEvent(nst_, sync_event(100), 0));
        unsigned long long* hnl = (unsigned long long*)ring_.take((W + 1) * sizeof(unsigned long long));
        EAGEN_NCCL(nc.AllGather(nl, nl + 1, sizeof(unsigned long long), ncclUint8, comm_, nst_));
        EAGEN_CUDA(cudaMemcpyAsync(hnl, nl, (W + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, nst_));
        EAGEN_NCCL(nc.AllGather(sums, all_sums, (size_t)d * sizeof(Prj), ncclUint8, comm_, nst_));
        EAGEN_CUDA(cudaEventRecord(sync_event(101), nst_));
        if (n_local) {
            // one all-gather per digit position, grouped into a single NCCL launch: row `pos` of every rank lands at its place in the
            // position-major planes over the global point range (no transpose pass afterwards)
            EAGEN_NCCL(nc.GroupStart());
            for (uint32_t pos = 0; pos < d; ++pos)
                EAGEN_NCCL(nc.AllGather(pl_local + (size_t)pos * n_local, planes + (size_t)pos * n_total, n_local, ncclUint8, comm_, nst_));
            EAGEN_NCCL(nc.GroupEnd());
            EAGEN_NCCL(nc.AllGather(tab_local, tab, n_local * (size_t)(base - 1) * 64, ncclUint8, comm_, nst_));
        }
        EAGEN_CUDA(cudaEventRecord(sync_event(102), nst_));
        // ---- replicated carry chain while the planes / table are still in flight
        EAGEN_CUDA(cudaStreamWaitEvent(st_, sync_event(101), 0));
        run_carry_chain(all_sums, nranks_, d, base, carries);
        res->carries.assign((size_t)d * 8, 0);
        uint64_t* hcar = (uint64_t*)ring_.take((size_t)d * 64);
        EAGEN_CUDA(cudaMemcpyAsync(hcar, carries, (size_t)d * 64, cudaMemcpyDeviceToHost, st_));
        EAGEN_CUDA(cudaStreamWaitEvent(st_, sync_event(102), 0));
        EAGEN_CUDA(cudaStreamSynchronize(nst_));   // n_local of every rank is on the host now
        for (size_t r = 0; r < W; ++r)
            if (hnl[1 + r] != (unsigned long long)n_local)
                throw StatusError{EAGEN_E_LEN, "eagen_lhs_witness_sharded: every rank must pass the same number of points (pad with zero scalars)"};
        if (!(flags & EAGEN_NO_FUNCTIONS)) run_position_trees(planes, tab, carries, n_total, base, d, pos0, pos1, flags, res.get(), out ? &so : nullptr);
        EAGEN_CUDA(cudaEventRecord(ev1_, st_));
        sync_check();
        std::memcpy(res->carries.data(), hcar, (size_t)d * 64);
        std::memcpy(res->carry, &res->carries[(size_t)(d - 1) * 8], 64);
        float ms = 0; EAGEN_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
        res->device_ms = ms;
        return res.release();
    }

    // which = 0: 32-bit IMAD per second; 1: base-field Montgomery products per second (register-resident chains); 2: IMAD.WIDE per second
    double microbench(int which) override {
        use();
        int sms = 0;
        EAGEN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_));
        const int blocks = sms * 8, threads = 256;
        void* buf = den_.ensure((size_t)blocks * threads * 32);
        double best = 0;
        for (int rep = 0; rep < 4; ++rep) {
            const int iters = which == 1 ? 512 : 4096;
            EAGEN_CUDA(cudaEventRecord(ev0_, st_));
            if (which == 0) k_imad_peak<<<blocks, threads, 0, st_>>>((uint32_t*)buf, iters, 12345u + rep);
            else if (which == 2) k_imad_wide_peak<<<blocks, threads, 0, st_>>>((uint32_t*)buf, iters, 12345u + rep);
            else k_modmul_peak<FB><<<blocks, threads, 0, st_>>>((F*)buf, iters);
            ++launches_;
            EAGEN_CUDA(cudaGetLastError());
            EAGEN_CUDA(cudaEventRecord(ev1_, st_));
            EAGEN_CUDA(cudaStreamSynchronize(st_));
            float ms = 0; EAGEN_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
            double ops = (double)blocks * threads * (double)iters * (which == 0 ? 128.0 : which == 2 ? 64.0 : 4.0);
            best = std::max(best, ops / (ms * 1e-3));
        }
        return best;
    }

    // largest power-of-two range below isqrt(order)+2: 127 bits on the Pasta curves, 126 on Grumpkin
    static int synth_bits() {
        HostU256 sq = host_isqrt(host_order<FS>());
        int top = 255;
        while (top > 0 && !((sq.w[top / 32] >> (top % 32)) & 1)) --top;
        return top;  // 2^top <= isqrt(order)
    }
    void synth_dev(uint64_t seed, size_t n, void* d_scalars, void* d_pts) override {
        use();
        launch(k_synth_inputs<CC>, n, 128, seed, n, synth_bits(), (Fe<FS>*)d_scalars, (F*)d_pts);
        sync_check();
    }
    void synth_host(uint64_t seed, size_t n, uint64_t* scalars, uint64_t* pts) override {
        use();
        if (!n) return;
        Fe<FS>* ds = (Fe<FS>*)in_scalars_.ensure(n * 32);
        F* dp = (F*)in_points_.ensure(n * 96);
        launch(k_synth_inputs<CC>, n, 128, seed, n, synth_bits(), ds, dp);
        EAGEN_CUDA(cudaMemcpyAsync(scalars, ds, n * 32, cudaMemcpyDeviceToHost, st_));
        EAGEN_CUDA(cudaMemcpyAsync(pts, dp, n * 96, cudaMemcpyDeviceToHost, st_));
        sync_check();
    }

    ResultImpl* divisor_host(const uint64_t* pts, size_t n, uint32_t flags, uint64_t* out_point) override {
        use();
        std::unique_ptr<ResultImpl> res(new ResultImpl());
        res->device = dev_; res->pool = pool_; res->n = n;
        if (n == 0) {  // reference: :455 -> (from_const(1), identity)
            res->nf = 1; res->a_stride = 1; res->b_stride = 1;
            F one = F::one();
            res->alloc(res->A, 32); res->alloc(res->B, 32);
            EAGEN_CUDA(cudaMemcpyAsync(res->A.p, &one, 32, cudaMemcpyHostToDevice, st_));
            res->la = {1}; res->lb = {0};
            if (out_point) std::memset(out_point, 0, 64);
            sync_check();
            return res.release();
        }
        F* dp = (F*)in_points_.ensure(n * 96);
        EAGEN_CUDA(cudaMemcpyAsync(dp, pts, n * 96, cudaMemcpyHostToDevice, st_));
        EAGEN_CUDA(cudaEventRecord(ev0_, st_));
        Aff* T = (Aff*)tpts_.ensure(n * sizeof(Aff));
        F* zs = (F*)den_.ensure(n * 32);
        launch(k_jac_z<FB>, n, 256, dp, n, zs);
        batch_invert(zs, n);
        launch(k_jac_to_affine<FB>, n, 256, (const F*)dp, (const F*)zs, n, T);
        std::vector<int> cnt{(int)n};
        Aff root;
        std::vector<Aff> roots(1);
        run_trees_safe(T, n, cnt, flags, res.get(), 0, 1, roots.data());
        EAGEN_CUDA(cudaEventRecord(ev1_, st_));
        sync_check();
        root = roots[0];
        if (out_point) std::memcpy(out_point, &root, 64);
        if (!(flags & EAGEN_PARTIAL) && !root.is_identity())
            throw StatusError{EAGEN_E_SUM_NONZERO, "compute_divisor_witness: points do not sum to the identity"};
        float ms = 0; EAGEN_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
        res->device_ms = ms;
        return res.release();
    }

    void poly_mul_host(const uint64_t* a, size_t la, const uint64_t* b, size_t lb, uint64_t* out) override {
        use();
        if (la + lb == 0) return;
        size_t len = la + lb - 1;
        if (la == 0 || lb == 0) { std::memset(out, 0, len * 32); return; }  // mul_naive with one empty operand (:54-62)
        int t = std::max(1, ceil_log2(len));
        if ((unsigned)t > FB::S) throw StatusError{EAGEN_E_NTT_TOO_LARGE, "polynomial product longer than 2^S"};
        size_t T = (size_t)1 << t;
        F* da = (F*)ea_.ensure(T * 32); F* db = (F*)eb_.ensure(T * 32);
        F* ca = (F*)oa_.ensure(std::max(la, len) * 32); F* cb = (F*)ob_.ensure(lb * 32);
        EAGEN_CUDA(cudaMemcpyAsync(ca, a, la * 32, cudaMemcpyHostToDevice, st_));
        EAGEN_CUDA(cudaMemcpyAsync(cb, b, lb * 32, cudaMemcpyHostToDevice, st_));
        ensure_twiddles(t);
        ntt(false, da, ca, la, (int)la, nullptr, 0, 0, t, 1, d_one_, 1);
        ntt(false, db, cb, lb, (int)lb, nullptr, 0, 0, t, 1, d_one_, 1);
        launch(k_mul_scale<FB>, T, 256, da, (const F*)db, T, half_pow(t));
        ntt(true, da, nullptr, 0, 0, ca, len, (int)len, t, 1, d_one_, 1);
        EAGEN_CUDA(cudaMemcpyAsync(out, ca, len * 32, cudaMemcpyDeviceToHost, st_));
        sync_check();
    }

    void ntt_host(uint64_t* data, uint32_t log_n, int inverse) override {
        use();
        if (log_n > FB::S) throw StatusError{EAGEN_E_NTT_TOO_LARGE, "log_n exceeds the field's two-adicity"};
        size_t T = (size_t)1 << log_n;
        if (log_n == 0) return;
        // best_fft is natural order in and out; the kernels are DIF (natural->bit-reversed) / DIT (bit-reversed->natural),
        // so the helper permutes on the host (this entry point is an API helper, not on the hot path)
        std::vector<uint64_t> tmp(T * 4);
        auto brev = [&](size_t i) { size_t r = 0; for (uint32_t b = 0; b < log_n; ++b) r |= ((i >> b) & 1) << (log_n - 1 - b); return r; };
        F* dd = (F*)ea_.ensure(T * 32);
        ensure_twiddles((int)log_n);
        if (!inverse) {
            EAGEN_CUDA(cudaMemcpyAsync(dd, data, T * 32, cudaMemcpyHostToDevice, st_));
            ntt(false, dd, nullptr, 0, 0, nullptr, 0, 0, (int)log_n, 1, d_one_, 1);
            EAGEN_CUDA(cudaMemcpyAsync(tmp.data(), dd, T * 32, cudaMemcpyDeviceToHost, st_));
            sync_check();
            for (size_t i = 0; i < T; ++i) std::memcpy(data + 4 * brev(i), &tmp[4 * i], 32);
        } else {
            for (size_t i = 0; i < T; ++i) std::memcpy(&tmp[4 * i], data + 4 * brev(i), 32);
            EAGEN_CUDA(cudaMemcpyAsync(dd, tmp.data(), T * 32, cudaMemcpyHostToDevice, st_));
            ntt(true, dd, nullptr, 0, 0, nullptr, 0, 0, (int)log_n, 1, d_one_, 1);
            EAGEN_CUDA(cudaMemcpyAsync(data, dd, T * 32, cudaMemcpyDeviceToHost, st_));
            sync_check();
        }
    }

    void batch_invert_host(uint64_t* elems, size_t n) override {
        use();
        if (!n) return;
        F* dd = (F*)den_.ensure(n * 32);
        EAGEN_CUDA(cudaMemcpyAsync(dd, elems, n * 32, cudaMemcpyHostToDevice, st_));
        batch_invert(dd, n);
        EAGEN_CUDA(cudaMemcpyAsync(elems, dd, n * 32, cudaMemcpyDeviceToHost, st_));
        sync_check();
    }

    void eval_host(const uint64_t* a, size_t la, const uint64_t* b, size_t lb, const uint64_t* pts, size_t n, uint64_t* out) override {
        use();
        if (!n) return;
        F* da = (F*)oa_.ensure(std::max<size_t>(la, 1) * 32); F* db = (F*)ob_.ensure(std::max<size_t>(lb, 1) * 32);
        F* dp = (F*)in_points_.ensure(n * 96);
        Aff* T = (Aff*)tpts_.ensure(n * sizeof(Aff));
        F* zs = (F*)den_.ensure(n * 32);
        if (la) EAGEN_CUDA(cudaMemcpyAsync(da, a, la * 32, cudaMemcpyHostToDevice, st_));
        if (lb) EAGEN_CUDA(cudaMemcpyAsync(db, b, lb * 32, cudaMemcpyHostToDevice, st_));
        EAGEN_CUDA(cudaMemcpyAsync(dp, pts, n * 96, cudaMemcpyHostToDevice, st_));
        launch(k_jac_z<FB>, n, 256, (const F*)dp, n, zs);
        batch_invert(zs, n);
        launch(k_jac_to_affine<FB>, n, 256, (const F*)dp, (const F*)zs, n, T);
        launch(k_eval_function<FB>, n, 128, (const F*)da, (int)la, (const F*)db, (int)lb, (const Aff*)T, n, zs);
        EAGEN_CUDA(cudaMemcpyAsync(out, zs, n * 32, cudaMemcpyDeviceToHost, st_));
        sync_check();
    }

    // prepare_scalar_witness for n scalars (reference: src/negbase_utils.rs:79-124); out: n x base x (num_limbs+1) entries of 32 B
    void scalar_witness_host(const uint64_t* scalars, size_t n, uint8_t base, uint32_t num_digits, uint32_t logtable, int mode, void* out) override {
        use();
        if (logtable < 1 || logtable > 24) throw StatusError{EAGEN_E_ARG, "prepare_scalar_witness: logtable must be in [1, 24]"};
        if (num_digits < 1 || num_digits > 4096) throw StatusError{EAGEN_E_ARG, "prepare_scalar_witness: num_digits must be in [1, 4096]"};
        if (mode != 0 && mode != 1) throw StatusError{EAGEN_E_ARG, "prepare_scalar_witness: mode must be 0 (faithful) or 1 (intended)"};
        if (n == 0) return;
        NegbaseParams prm = make_negbase_params<FS>(base);
        const uint32_t num_limbs = (num_digits + logtable - 1) / logtable;
        const size_t bytes = n * (size_t)base * (num_limbs + 1) * sizeof(PswEntry);
        Fe<FS>* ds = (Fe<FS>*)in_scalars_.ensure(n * 32);
        uint8_t* planes = (uint8_t*)planes_.ensure(n * prm.d);
        PswEntry* dout = (PswEntry*)table_.ensure(bytes);
        EAGEN_CUDA(cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, st_));
        run_negbase(ds, n, prm, planes, nullptr);
        {
            Scope ps(this, "scalar_witness", (double)n * (32.0 + (double)base * prm.d) + (double)bytes, 0.0);
            launch2d(k_scalar_witness<FS>, dim3((unsigned)((n + 127) / 128), base), 128, (const uint8_t*)planes, n, prm.d, (uint32_t)base, num_digits,
                     logtable, num_limbs, mode, (const Fe<FS>*)ds, dout, d_err_);
        }
        EAGEN_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, st_));
        sync_check();
    }

    // compute_divisor_witness_naive (reference: src/regular_functions_utils.rs:483-551).  The pairing of a round depends on which
    // points of the list are the identity (`if inc1 != C::identity()`, :517,:533), so every round reads the identity flags back
    // (one byte per point) and the host replays the reference's pop order into index pairs; sums, lines and the batched
    // inversion of the slopes' denominators run on the device.  *n_pos / *n_neg: capacity in (lines), count out.
    void naive_host(const uint64_t* pts, size_t n, uint64_t* pos, size_t* n_pos, uint64_t* neg, size_t* n_neg) override {
        use();
        const size_t cap_pos = *n_pos, cap_neg = *n_neg;
        *n_pos = 0; *n_neg = 0;
        if (n == 0) return;
        typedef LineTriple<FB> Line;
        F* dp = (F*)in_points_.ensure(n * 96);
        EAGEN_CUDA(cudaMemcpyAsync(dp, pts, n * 96, cudaMemcpyHostToDevice, st_));
        Aff* lists = (Aff*)tpts_.ensure(2 * (n + 2) * sizeof(Aff));
        Aff* L[2] = {lists, lists + (n + 2)};          // pos, neg
        size_t len[2] = {n, 0};
        F* zs = (F*)den_.ensure(n * 32);
        launch(k_jac_z<FB>, n, 256, (const F*)dp, n, zs);
        batch_invert(zs, n);
        launch(k_jac_to_affine<FB>, n, 256, (const F*)dp, (const F*)zs, n, L[0]);
        Line* lines[2] = {(Line*)oa_.ensure((n + 1) * sizeof(Line)), (Line*)ob_.ensure((n + 1) * sizeof(Line))};
        size_t nl[2] = {0, 0};
        uint8_t* dflags = (uint8_t*)rows_.ensure(n + 2);
        int2* dpairs = (int2*)cnt_.ensure((n / 2 + 1) * sizeof(int2));
        std::vector<uint8_t> flags;
        std::vector<int2> pairs;
        auto round = [&](int from) {
            const int to = 1 - from;
            const size_t m = len[from];
            if (m <= 1) return;
            flags.resize(m);
            launch(k_identity_flags<FB>, m, 256, (const Aff*)L[from], m, dflags);
            EAGEN_CUDA(cudaMemcpyAsync(flags.data(), dflags, m, cudaMemcpyDeviceToHost, st_));
            EAGEN_CUDA(cudaStreamSynchronize(st_));
            pairs.clear();
            size_t top = m;
            while (top > 1) {   // :515-520
                const size_t i1 = --top;
                if (!flags[i1]) { const size_t i2 = --top; pairs.push_back(make_int2((int)i1, (int)i2)); }
            }
            len[from] = top;
            const size_t np = pairs.size();
            if (np == 0) return;
            if (nl[from] + np > n + 1) throw StatusError{EAGEN_E_ARG, "compute_divisor_witness_naive: internal line buffer overflow"};
            EAGEN_CUDA(cudaMemcpyAsync(dpairs, pairs.data(), np * sizeof(int2), cudaMemcpyHostToDevice, st_));
            launch(k_naive_den<FB>, np, 256, (const Aff*)L[from], (const int2*)dpairs, np, zs);
            batch_invert(zs, np);
            launch(k_naive_finish<FB>, np, 128, (const Aff*)L[from], (const int2*)dpairs, np, (const F*)zs, L[to] + len[to], lines[from] + nl[from]);
            EAGEN_CUDA(cudaStreamSynchronize(st_));   // `pairs` is reused by the next round
            len[to] += np;
            nl[from] += np;
        };
        while (len[0] > 1 || len[1] > 1) { round(0); round(1); }   // :513
        Aff left[2];
        std::memset(left, 0, sizeof left);
        for (int s = 0; s < 2; ++s) if (len[s]) EAGEN_CUDA(cudaMemcpyAsync(&left[s], L[s], sizeof(Aff), cudaMemcpyDeviceToHost, st_));
        sync_check();
        const bool ok = (len[0] == 0 && len[1] == 0) || (len[0] == 1 && len[1] == 0 && left[0].is_identity()) ||
                        (len[0] == 0 && len[1] == 1 && left[1].is_identity()) ||
                        (len[0] == 1 && len[1] == 1 && left[0].x == left[1].x && left[0].y == left[1].y);   // :549-553
        if (!ok) throw StatusError{EAGEN_E_SUM_NONZERO, "compute_divisor_witness_naive: points do not sum to the identity"};
        if (nl[0] > cap_pos || nl[1] > cap_neg) throw StatusError{EAGEN_E_LEN, "compute_divisor_witness_naive: line buffers too small"};
        if (nl[0]) EAGEN_CUDA(cudaMemcpyAsync(pos, lines[0], nl[0] * sizeof(Line), cudaMemcpyDeviceToHost, st_));
        if (nl[1]) EAGEN_CUDA(cudaMemcpyAsync(neg, lines[1], nl[1] * sizeof(Line), cudaMemcpyDeviceToHost, st_));
        EAGEN_CUDA(cudaStreamSynchronize(st_));
        *n_pos = nl[0]; *n_neg = nl[1];
    }

    // RegularFunction::ev of every function held by a device-resident result at m Jacobian points (reference:
    // src/regular_functions_utils.rs:228-237); out: nf x m elements, function-major.  The coefficients never leave the device.
    void result_eval_host(ResultImpl* r, const uint64_t* pts, size_t m, uint64_t* out) override {
        use();
        if (r->device != dev_) throw StatusError{EAGEN_E_ARG, "eagen_result_eval: the result lives on another device"};
        const size_t nf = r->nf;
        if (m == 0 || nf == 0) return;
        int lmax = 1;
        for (size_t k = 0; k < nf; ++k) lmax = std::max(lmax, std::max(r->la[k], r->lb[k]));
        const size_t len = (size_t)lmax;
        const int nchunks = (int)((len + EVAL_CHUNK - 1) / EVAL_CHUNK);
        std::vector<int> lens(2 * nf);
        for (size_t k = 0; k < nf; ++k) { lens[k] = r->la[k]; lens[nf + k] = r->lb[k]; }
        int* dl = (int*)cnt_.ensure(2 * nf * sizeof(int));
        EAGEN_CUDA(cudaMemcpyAsync(dl, lens.data(), 2 * nf * sizeof(int), cudaMemcpyHostToDevice, st_));
        const size_t batch = 64;
        F* dp = (F*)in_points_.ensure(std::min(m, batch) * 96);
        Aff* T = (Aff*)tpts_.ensure(std::min(m, batch) * sizeof(Aff));
        F* zs = (F*)den_.ensure(std::min(m, batch) * 32);
        F* X = (F*)wk_[0].ensure(std::min(m, batch) * len * 32);
        F* part = (F*)binv_.ensure(std::min(m, batch) * nf * (size_t)nchunks * 64);
        F* dout = (F*)oa_.ensure(nf * std::min(m, batch) * 32);
        std::vector<uint64_t> tmp(nf * std::min(m, batch) * 4);
        for (size_t j0 = 0; j0 < m; j0 += batch) {
            const size_t mb = std::min(batch, m - j0);
            EAGEN_CUDA(cudaMemcpyAsync(dp, pts + j0 * 12, mb * 96, cudaMemcpyHostToDevice, st_));
            launch(k_jac_z<FB>, mb, 256, (const F*)dp, mb, zs);
            batch_invert(zs, mb);
            launch(k_jac_to_affine<FB>, mb, 256, (const F*)dp, (const F*)zs, mb, T);
            {
            Scope ps(this, "result_eval", (double)mb * (double)r_total_elems(r) * 32.0, (double)mb * ((double)r_total_elems(r) + (double)len));
            launch(k_pow_init<FB>, mb, 64, (const Aff*)T, mb, len, X);
            for (size_t half = 1; half + 1 < len; half *= 2)
                launch2d(k_pow_step<FB>, dim3((unsigned)((half + 255) / 256), (unsigned)mb), 256, X, len, half, len);
            launch2d(k_eval_chunks<FB>, dim3((unsigned)nchunks, (unsigned)nf, (unsigned)mb), EVAL_THREADS, (const F*)r->A.p, r->a_stride,
                     (const F*)r->B.p, r->b_stride, (const int*)dl, (const int*)(dl + nf), (const F*)X, len, nchunks, part);
            launch(k_eval_finish<FB>, nf * mb, 128, (const F*)part, nchunks, (int)nf, (const Aff*)T, mb, dout);
            }
            EAGEN_CUDA(cudaMemcpyAsync(tmp.data(), dout, nf * mb * 32, cudaMemcpyDeviceToHost, st_));
            sync_check();
            for (size_t k = 0; k < nf; ++k) std::memcpy(out + (k * m + j0) * 4, tmp.data() + k * mb * 4, mb * 32);
        }
    }
    static size_t r_total_elems(const ResultImpl* r) {
        size_t t = 0;
        for (size_t k = 0; k < r->nf; ++k) t += (size_t)r->la[k] + (size_t)r->lb[k];
        return t;
    }

private:
    struct ProfEntry { std::string name; uint64_t launches = 0, scopes = 0; double ms = 0, bytes = 0, modmul = 0; };
    struct ProfPending { int tag; cudaEvent_t a, b; uint64_t l0, l1; cudaStream_t st; };
    bool prof_on_ = false;
    bool prof_detail_ = false;   // mode 2: one entry per (kernel group, tree level): "name@L<level>"
    int prof_level_ = -1;        // tree level the launches being issued belong to (-1: outside the level loop)
    std::vector<ProfEntry> prof_;
    std::vector<ProfPending> pending_;
    std::vector<cudaEvent_t> evpool_;
    int prof_tag(const char* name) {
        for (size_t i = 0; i < prof_.size(); ++i) if (prof_[i].name == name) return (int)i;
        ProfEntry e; e.name = name; prof_.push_back(e);
        return (int)prof_.size() - 1;
    }
    cudaEvent_t get_event() {
        if (!evpool_.empty()) { cudaEvent_t e = evpool_.back(); evpool_.pop_back(); return e; }
        cudaEvent_t e; EAGEN_CUDA(cudaEventCreate(&e)); return e;
    }
    // RAII scope: events around the launches issued while it is alive
    struct Scope {
        Engine* eng; bool live;
        Scope(Engine* e, const char* name, double bytes, double modmul) : eng(e), live(e->prof_on_) {
            if (!live) return;
            int tag;
            char buf[112];
            // launches issued to the side stream overlap the main stream's kernels: their event times are not additive with the
            // rest, so they are booked under their own name ("pair_points~side")
            const char* side = eng->ls_ != eng->st_ ? "~side" : "";
            if (eng->prof_detail_ && eng->prof_level_ >= 0) snprintf(buf, sizeof buf, "%s%s@L%02d", name, side, eng->prof_level_);
            else snprintf(buf, sizeof buf, "%s%s", name, side);
            tag = eng->prof_tag(buf);
            eng->prof_[tag].bytes += bytes; eng->prof_[tag].modmul += modmul; eng->prof_[tag].scopes += 1;
            ProfPending p; p.tag = tag; p.a = eng->get_event(); p.b = eng->get_event(); p.l0 = eng->launches_; p.l1 = 0;
            p.st = eng->ls_;
            cudaEventRecord(p.a, p.st);
            eng->pending_.push_back(p);
            idx = eng->pending_.size() - 1;
        }
        ~Scope() {
            if (!live) return;
            ProfPending& p = eng->pending_[idx];
            cudaEventRecord(p.b, p.st);
            p.l1 = eng->launches_;
        }
        size_t idx = 0;
    };
    // host-side milliseconds spent while the GPU waits for the host (booked as "host~<what>"; profiling only)
    void prof_host(const char* what, std::chrono::steady_clock::time_point t0) {
        if (!prof_on_) return;
        int t = prof_tag(what);
        prof_[t].ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        prof_[t].scopes += 1;
    }
    void prof_collect() {
        if (pending_.empty()) return;
        cudaStreamSynchronize(st_);
        cudaStreamSynchronize(pst_);
        // idle time of the main stream between consecutive scopes (waits for the side stream, launch gaps): booked as "idle~gaps"
        // (mode 2: one entry per scope that FOLLOWS the gap, "idle~before:<scope>")
        const ProfPending* prev = nullptr;
        std::vector<std::pair<std::string, double>> gaps;
        for (auto& p : pending_) {
            if (p.st != st_) continue;
            float g = 0;
            if (prev && cudaEventElapsedTime(&g, prev->b, p.a) == cudaSuccess && g > 0)
                gaps.push_back({prof_detail_ ? "idle~before:" + prof_[p.tag].name : std::string("idle~gaps"), (double)g});
            prev = &p;
        }
        for (auto& g : gaps) { int t = prof_tag(g.first.c_str()); prof_[t].ms += g.second; prof_[t].scopes += 1; }
        for (auto& p : pending_) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) prof_[p.tag].ms += ms;
            prof_[p.tag].launches += p.l1 - p.l0;
            evpool_.push_back(p.a); evpool_.push_back(p.b);
        }
        pending_.clear();
    }

    int dev_;
    int sm_count_ = 148;
    std::shared_ptr<BufPool> pool_ = std::make_shared<BufPool>();
    cudaStream_t st_ = nullptr, cst_ = nullptr;  // compute stream, copy stream (streamed results)
    cudaStream_t pst_ = nullptr;                 // side stream: the point pyramid (output points, lines, descriptors) runs ahead of the polynomial levels
    cudaStream_t ls_ = nullptr;                  // stream launch() / Scope / batch_invert currently issue to (st_ unless inside OnStream)
    std::vector<cudaEvent_t> sync_events_;       // timing-disabled events for cross-stream ordering, reused across calls
    cudaEvent_t sync_event(size_t i) {
        while (sync_events_.size() <= i) { cudaEvent_t e; EAGEN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); sync_events_.push_back(e); }
        return sync_events_[i];
    }
    struct OnStream {   // RAII: route launches to another stream
        Engine* eng; cudaStream_t prev;
        OnStream(Engine* e, cudaStream_t s) : eng(e), prev(e->ls_) { e->ls_ = s; }
        ~OnStream() { eng->ls_ = prev; }
    };
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    int* d_err_ = nullptr;
    int last_tree_err_ = 0;
    uint32_t plan_group_ = 0, plan_npos_ = 0;    // group plan of the last call that had to ask the driver for the free memory
    size_t plan_per_tree_ = 0;
    ncclComm_t comm_ = nullptr;                  // multi-GPU: this rank's communicator (null: single GPU)
    int nranks_ = 1, rank_ = 0;
    bool comm_owned_ = false;
    cudaStream_t nst_ = nullptr;                 // communication stream (collectives overlap the carry chain)
    std::vector<uint32_t> stream_split_{75, 25}; // streamed output: per cent of the positions per group (eagen_ctx_set_stream_split)
    DevBuf sh_planes_, sh_table_;
    void ensure_comm_stream() { if (!nst_) { use(); EAGEN_CUDA(cudaStreamCreateWithFlags(&nst_, cudaStreamNonBlocking)); } }
    PinnedRing ring_;
    int* d_one_ = nullptr;
    uint64_t launches_ = 0;
    uint64_t iso_fallbacks_ = 0;   // trees rebuilt on an isomorphic curve after a domain collision
    int tw_max_ = 0;
    DevBuf tw_fwd_, tw_inv_, xt_, gt_;
    DevBuf in_scalars_, in_points_, planes_, rows_, table_, sums_, partials_, carries_, carries_proj_;
    DevBuf cnt_, tree_n_, tree_of_pos_, tpts_, isodeg_;
    std::unique_ptr<Scope> prof_scope_;
    DevBuf pden_, pbinv_, ptlev_, lvlcnt_, a_[2], b_[2], ea_, eb_, oa_, ob_, wk_[2], top_, twist_, den_, binv_, desc_, tops_, lead_;

    void use() { EAGEN_CUDA(cudaSetDevice(dev_)); }

    template <class K, class... Args>
    void launch(K kernel, size_t work, int threads, Args... args) {
        if (work == 0) return;
        size_t blocks = (work + threads - 1) / threads;
        kernel<<<(unsigned)blocks, threads, 0, ls_>>>(args...);
        ++launches_;
        EAGEN_CUDA(cudaGetLastError());
    }
    template <class K, class... Args>
    void launch2d(K kernel, dim3 grid, int threads, Args... args) {
        if (grid.x == 0 || grid.y == 0) return;
        kernel<<<grid, threads, 0, ls_>>>(args...);
        ++launches_;
        EAGEN_CUDA(cudaGetLastError());
    }

    void sync_check() {
        int* herr = (int*)ring_.take(sizeof(int));
        EAGEN_CUDA(cudaMemcpyAsync(herr, d_err_, sizeof(int), cudaMemcpyDeviceToHost, st_));
        EAGEN_CUDA(cudaStreamSynchronize(st_));
        prof_collect();
        const int e = *herr;
        if (e) {
            EAGEN_CUDA(cudaMemset(d_err_, 0, sizeof(int)));
            if (e & KERR_RANGE) throw StatusError{EAGEN_E_RANGE, "scalar out of range: must be < isqrt(order)+2"};
            if (e & KERR_DIGITS) throw StatusError{EAGEN_E_DIGITS, "negbase expansion does not fit in d digits"};
            if (e & KERR_PSW_SLOT) throw StatusError{EAGEN_E_ARG, "prepare_scalar_witness (faithful mode): limb slot i % logtable + 1 exceeds num_limbs (the reference indexes out of bounds here)"};
            throw StatusError{EAGEN_E_DOMAIN, "an intermediate point's x-coordinate lies on the evaluation domain"};
        }
    }

    static F half_pow(int t) {
        F r = F::one(), h = F::two_inv();
        for (int i = 0; i < t; ++i) r = mul(r, h);
        return r;
    }

    void ensure_twiddles(int t) {
        if (t <= tw_max_) return;
        if ((unsigned)t > FB::S) throw StatusError{EAGEN_E_NTT_TOO_LARGE, "transform size exceeds the field's two-adicity"};
        size_t cnt = (size_t)1 << t;
        F* f = (F*)tw_fwd_.ensure(cnt * 32);
        F* i = (F*)tw_inv_.ensure(cnt * 32);
        launch(k_gen_twiddles<FB>, cnt, 256, f, t, 0);
        launch(k_gen_twiddles<FB>, cnt, 256, i, t, 1);
        F* xt = (F*)xt_.ensure(2 * cnt * 32);
        F* gt = (F*)gt_.ensure(2 * cnt * 32);
        launch(k_gen_points<CC>, 2 * cnt, 256, (const F*)f, t, xt, gt);
        F* tws = (F*)twist_.ensure(cnt * 32);
        launch(k_gen_twist_all<FB>, cnt, 256, (const F*)f, t, tws);
        tw_max_ = t;
    }
    const F* xtab(int t) { return xt_.as<F>() + ((size_t)1 << t); }
    const F* gtab(int t) { return gt_.as<F>() + ((size_t)1 << t); }
    const F* twist_tab(int l) { return twist_.as<F>() + ((size_t)1 << l); }   // coset pre-multipliers of the level with m = 2^l
    const F* tw(bool inverse, int t) { return (inverse ? tw_inv_.as<F>() : tw_fwd_.as<F>()) + ((size_t)1 << (t - 1)); }

    // in-place batched inversion of M elements (zeros stay zero); launches go to the current launch stream, `scratch` must not be
    // shared with a batch in flight on another stream
    void batch_invert(F* x, size_t M) { batch_invert(x, M, binv_); }
    void batch_invert(F* x, size_t M, DevBuf& scratch_buf) {
        if (M == 0) return;
        Scope ps(this, "batch_invert", 160.0 * M * (1.0 + 1.0 / (BINV_G - 1)), 3.0 * M * (1.0 + 1.0 / (BINV_G - 1)));
        // level sizes
        std::vector<size_t> sz{M};
        while (sz.back() > 1024) sz.push_back((sz.back() + BINV_G - 1) / BINV_G);
        size_t scratch = 0;
        for (size_t l = 0; l + 1 < sz.size(); ++l) scratch += sz[l] + sz[l + 1];
        F* s = (F*)scratch_buf.ensure(std::max<size_t>(scratch, 1) * 32);
        std::vector<F*> xs{x}, prefs;
        F* cur = s;
        for (size_t l = 0; l + 1 < sz.size(); ++l) { prefs.push_back(cur); cur += sz[l]; xs.push_back(cur); cur += sz[l + 1]; }
        for (size_t l = 0; l + 1 < sz.size(); ++l)
            launch(k_binv_up<FB>, sz[l + 1], 128, (const F*)xs[l], prefs[l], xs[l + 1], sz[l], sz[l + 1]);
        launch(k_binv_base<FB>, sz.back(), 64, xs.back(), sz.back());
        for (size_t l = sz.size() - 1; l-- > 0;)
            launch(k_binv_down<FB>, sz[l + 1], 128, xs[l], (const F*)prefs[l], (const F*)xs[l + 1], sz[l], sz[l + 1]);
    }

    // One polynomial family of a batched transform: n_tr arrays of size 2^t in `data` (the in-place workspace of the middle
    // passes), optionally gathered from compact slots by the first pass (`src`, zero padded beyond src_len) and scattered back
    // into compact slots by the last one (`dst`).
    struct NttJob {
        F* data; const F* src; size_t src_stride; int src_len; F* dst; size_t dst_stride; int dst_len; const F* sub_top;
    };
    // batched transform of up to two families (the a and b polynomials of a level) per launch: blockIdx.y selects the family
    void ntt(bool inverse, const NttJob* jobs, int njobs, int t, size_t n_tr, const int* counts, int node_max, size_t n_present = (size_t)-1,
             const F* twist = nullptr, size_t dst_off = 0) {
        if (n_present == (size_t)-1) n_present = n_tr;
        std::vector<std::pair<int, int>> plan = ntt_plan(t);
        {   // algorithmic work: every present element is read and written once per pass, t/2 modmul per element in total
            double el = (double)n_present * (double)((size_t)1 << t);
            double bytes = 0;
            for (int j = 0; j < njobs; ++j) {
                bytes += 64.0 * el * plan.size();
                if (jobs[j].src) bytes -= 32.0 * el - 32.0 * (double)n_present * jobs[j].src_len;
                if (jobs[j].dst) bytes -= 32.0 * el - 32.0 * (double)n_present * jobs[j].dst_len;
            }
            prof_scope_.reset(new Scope(this, inverse ? "ntt_inverse" : "ntt_forward", bytes,
                                        njobs * (el * t / 2.0 - (t >= 2 ? 0.75 * el : 0.5 * el) + (twist ? el : 0.0))));  // stages 0/1 have w = 1
        }
        struct Closer { std::unique_ptr<Scope>& s; ~Closer() { s.reset(); } } closer{prof_scope_};
        NttPass<FB> a;
        a.counts = counts; a.node_max = node_max;
        a.total = n_tr << t; a.t = t;
        if (inverse) std::reverse(plan.begin(), plan.end());
        size_t tiles = (a.total + NTT_TILE - 1) / NTT_TILE;
        for (size_t p = 0; p < plan.size(); ++p) {
            a.s_hi = plan[p].first; a.s_lo = plan[p].second;
            // w_T^(j 2^(t-1-s)) = w_{2^(s+1)}^j: the last (contiguous) pass only needs the 2^k-point table, which stays in L1
            a.tw_t = a.s_lo == 0 ? a.s_hi + 1 : t;
            a.tw = tw(inverse, a.tw_t);
            a.twist = (p == 0) ? twist : nullptr;
            a.dst_off = dst_off;
            a.final_pass = (p + 1 == plan.size()) ? 1 : 0;
            for (int j = 0; j < 2; ++j) {
                const NttJob& jb = jobs[j < njobs ? j : 0];
                NttSide<FB>& sd = a.side[j];
                sd.data = jb.data;
                sd.src = (p == 0) ? jb.src : nullptr; sd.src_stride = jb.src_stride; sd.src_len = jb.src_len;
                sd.dst = (p + 1 == plan.size()) ? jb.dst : nullptr; sd.dst_stride = jb.dst_stride; sd.dst_len = jb.dst_len;
                sd.sub_top = (p + 1 == plan.size()) ? jb.sub_top : nullptr;
            }
            dim3 grid((unsigned)tiles, (unsigned)njobs);
            if (inverse) k_ntt_pass<FB, true><<<grid, NTT_THREADS, 0, st_>>>(a);
            else k_ntt_pass<FB, false><<<grid, NTT_THREADS, 0, st_>>>(a);
            ++launches_;
            EAGEN_CUDA(cudaGetLastError());
        }
    }
    // single-family convenience form
    void ntt(bool inverse, F* data, const F* src, size_t src_stride, int src_len, F* dst, size_t dst_stride, int dst_len,
             int t, size_t n_tr, const int* counts, int node_max) {
        NttJob jb{data, src, src_stride, src_len, dst, dst_stride, dst_len, nullptr};
        ntt(inverse, &jb, 1, t, n_tr, counts, node_max);
    }

    void run_negbase(const Fe<FS>* ds, size_t n, const NegbaseParams& prm, uint8_t* planes, uint8_t* rows) {
        Scope ps(this, "negbase", (double)n * (32.0 + prm.d * (rows ? 2.0 : 1.0)), (double)n);
        if (n == 0) return;
        const size_t smem = ((size_t)prm.nw * NEGBASE_THREADS + prm.lut_n) * sizeof(uint32_t);
        const int vec = (n % 4 == 0) && (((uintptr_t)planes & 3) == 0);
        const size_t chunks = (n + NEGBASE_THREADS - 1) / NEGBASE_THREADS;
        const unsigned grid = (unsigned)std::min<size_t>(chunks, (size_t)sm_count_ * 12);   // persistent: a multiple of the SM count
        k_negbase<FS><<<grid, NEGBASE_THREADS, smem, st_>>>(ds, n, prm, planes, rows, vec, d_err_);
        ++launches_;
        EAGEN_CUDA(cudaGetLastError());
    }

    void run_multiples(const F* dp, size_t n, uint8_t base, Aff* tab) {
        size_t m = n * (size_t)(base - 1);
        if (m == 0) return;
        F* zs = (F*)den_.ensure(m * 32);
        {
            Scope ps(this, "multiples", (double)n * 96.0 + (double)m * 96.0, (double)n * (3.0 + 14.0 * (base - 2)));
            launch(k_multiples_proj<CC>, n, 128, dp, n, (uint32_t)base, tab, zs);
        }
        batch_invert(zs, m);
        Scope ps(this, "multiples_finish", (double)m * 160.0, (double)m * 2.0);
        launch(k_scale_by_zinv<FB>, m, 256, tab, (const F*)zs, m);
    }

    void run_shard_sums(const Fe<FS>* ds, const F* dp, size_t n, const NegbaseParams& prm, uint8_t* planes, uint8_t* rows, Aff* tab, Prj* sums) {
        const uint32_t d = prm.d;
        if (n) {
            run_negbase(ds, n, prm, planes, rows);
            run_multiples(dp, n, (uint8_t)prm.base, tab);
        }
        // points per thread: every block ends with a 7-step tree reduce of complete additions behind barriers, so large inputs
        // amortise it over 128 points per thread (measured at 2^20 Pallas, tools/scope_ms.py: 32 -> 15.3 ms, 64 -> 13.5, 128 -> 12.6);
        // smaller inputs keep 32 so that the grid still fills the chip
        int per_thread = n >= ((size_t)1 << 19) ? 128 : 32;
        size_t chunk = (size_t)SUMS_THREADS * per_thread;
        int chunks = (int)std::max<size_t>(1, (n + chunk - 1) / chunk);
        Prj* partials = (Prj*)partials_.ensure((size_t)d * chunks * sizeof(Prj));
        // work: one table point (64 B) + one digit per (position, point) with a non-zero digit (~(b-1)/b of them), 13 modmul per mixed add
        Scope ps(this, "digit_sums", (double)n * d * (1.0 + 64.0 * (prm.base - 1) / prm.base), (double)n * d * 13.0 * (prm.base - 1) / prm.base);
        launch2d(k_digit_sums<CC>, dim3(chunks, d), SUMS_THREADS, (const uint8_t*)planes, (const Aff*)tab, n, prm.base, per_thread, partials);
        launch2d(k_reduce_partials<CC>, dim3(d, 1), SUMS_THREADS, (const Prj*)partials, chunks, sums);
    }

    void run_carry_chain(const Prj* sums, int nparts, uint32_t d, uint8_t base, Aff* carries) {
        Prj* cp = (Prj*)carries_proj_.ensure((size_t)d * sizeof(Prj));
        F* zs = (F*)lead_.ensure((size_t)d * 32);
        Scope ps(this, "carry_chain", (double)d * 96.0 * (nparts + 1), (double)d * (14.0 * nparts + 14.0 * 4));
        if (nparts > 1) {   // fold the ranks' partial sums per position in parallel first: the serial chain stays d steps of one addition
            Prj* folded = (Prj*)partials_.ensure((size_t)d * sizeof(Prj));
            launch(k_sum_parts<CC>, d, 32, sums, d, nparts, folded);
            sums = folded; nparts = 1;
        }
        launch(k_carry_chain<CC>, 1, 32, sums, d, (uint32_t)base, nparts, cp, zs);
        batch_invert(zs, d);
        launch(k_proj_to_affine<FB>, d, 64, (const Prj*)cp, (const F*)zs, (size_t)d, carries);
    }

    // divisor witnesses for iteration positions [pos_begin, pos_end); function index k = d-1-pos (ret.reverse(), :132)
    void run_position_trees(const uint8_t* planes, const Aff* tab, const Aff* carries, size_t n, uint8_t base, uint32_t d,
                            uint32_t pos_begin, uint32_t pos_end, uint32_t flags, ResultImpl* res, const StreamOut* so = nullptr) {
        const uint32_t npos = pos_end - pos_begin;
        if (npos == 0) return;
        int chunks = (int)std::max<size_t>(1, (n + CHUNK_PTS - 1) / CHUNK_PTS);
        int* cnt = (int*)cnt_.ensure((size_t)d * chunks * sizeof(int));
        int* tree_n = (int*)tree_n_.ensure((size_t)d * sizeof(int));
        {
            Scope ps(this, "count_scan", (double)n * d, 0.0);
            launch2d(k_count_nonzero, dim3(chunks, d), 256, planes, n, chunks, cnt);
            launch(k_scan_chunks<FB>, d, 64, cnt, chunks, d, (uint32_t)base, carries, tree_n);
        }
        // list lengths (they size everything below) and, in the same round trip, the error flag of K1: a scalar out of range fails
        // here, before any tree is built or streamed (the reference asserts before any work, src/argument_witness_calc.rs:97)
        int* hstage = (int*)ring_.take((size_t)(d + 1) * sizeof(int));
        EAGEN_CUDA(cudaMemcpyAsync(hstage, tree_n, (size_t)d * sizeof(int), cudaMemcpyDeviceToHost, st_));
        EAGEN_CUDA(cudaMemcpyAsync(hstage + d, d_err_, sizeof(int), cudaMemcpyDeviceToHost, st_));
        EAGEN_CUDA(cudaStreamSynchronize(st_));
        auto t_host = std::chrono::steady_clock::now();
        if (hstage[d] & (KERR_RANGE | KERR_DIGITS)) sync_check();   // throws with the proper status
        std::vector<int> hn(hstage, hstage + d);
        size_t nmax = 1;
        for (uint32_t p = pos_begin; p < pos_end; ++p) nmax = std::max<size_t>(nmax, (size_t)hn[p]);
        // result slots sized for the deepest tree of the range
        int Lmax = ceil_log2((nmax + 1) / 2);
        res->nf = npos;
        res->a_stride = ((size_t)1 << Lmax) + 1;
        res->b_stride = std::max<size_t>((size_t)1 << Lmax, 1);
        res->alloc(res->A, res->nf * res->a_stride * 32);
        res->alloc(res->B, res->nf * res->b_stride * 32);
        res->la.assign(npos, 0); res->lb.assign(npos, 0);
        prof_host("host~result_alloc", t_host);
        t_host = std::chrono::steady_clock::now();
        // group positions so that one group's working set fits the memory budget.  cudaMemGetInfo was measured at 0.2 - 27 ms per call
        // on the B200 boxes once the process holds tens of GB (it sits between two kernels with the GPU idle, and was the whole
        // process-to-process spread of the step time: profiles/r02_experiments.md), so the plan of the previous call is reused
        // whenever this call needs no more than that one did: the working buffers are grow-only and already large enough.
        size_t per_tree = tree_bytes(nmax);
        uint32_t group;
        if (plan_group_ && per_tree <= plan_per_tree_ && npos <= plan_npos_) {
            group = std::min<uint32_t>(npos, plan_group_);
        } else {
            size_t free_b = 0, total_b = 0;
            EAGEN_CUDA(cudaMemGetInfo(&free_b, &total_b));
            size_t budget = (size_t)((double)(free_b + pooled_bytes()) * 0.80);
            group = (uint32_t)std::max<size_t>(1, std::min<size_t>(npos, budget / std::max<size_t>(per_tree, 1)));
            plan_group_ = group; plan_per_tree_ = per_tree; plan_npos_ = npos;
        }
        prof_host("host~group_plan", t_host);
        // group sizes: as large as the memory budget allows; for streamed output a decreasing schedule (per cent of the positions,
        // eagen_ctx_set_stream_split, default 75,25) so that every group's copy hides behind the next group's compute and only the
        // small last group's copy is exposed.  More, smaller groups shorten the exposed copy but add ~1300 launches each:
        // measured (tools/e2e_groups.py, 2^20 Pallas, ms end to end): one group 328, 70/30 316, 75/25 315.5, 80/20 314-321, 60/30/10 318-320;
        // device-resident results use one group per memory budget (two or more groups measured 289 / 292 / 295 ms against 286.5)
        std::vector<uint32_t> sizes;
        if (so) {
            const std::vector<uint32_t>& pct = stream_split_;
            uint32_t left = npos;
            for (size_t i = 0; i < pct.size() && left; ++i) {
                uint32_t want = i + 1 == pct.size() ? left : std::min<uint32_t>(left, std::max<uint32_t>(1, (npos * pct[i] + 50) / 100));
                while (want) {   // never more than the memory budget per group
                    uint32_t g = std::min(want, group);
                    sizes.push_back(g); want -= g; left -= g;
                }
            }
            while (left) { uint32_t g = std::min(left, group); sizes.push_back(g); left -= g; }
        } else {
            for (uint32_t left = npos; left;) { uint32_t g = std::min(left, group); sizes.push_back(g); left -= g; }
        }
        int* tree_of_pos = (int*)tree_of_pos_.ensure((size_t)d * sizeof(int));
        std::vector<Aff> roots(npos);
        uint32_t g0 = pos_begin;
        for (size_t gi = 0; gi < sizes.size(); g0 += sizes[gi], ++gi) {
            uint32_t g1 = std::min(pos_end, g0 + sizes[gi]), nt = g1 - g0;
            std::vector<int> map(d, -1), cnts(nt);
            for (uint32_t p = g0; p < g1; ++p) { map[p] = (int)(p - g0); cnts[p - g0] = hn[p]; }
            {
                int* stage = (int*)ring_.take((size_t)d * sizeof(int));
                std::memcpy(stage, map.data(), (size_t)d * sizeof(int));
                EAGEN_CUDA(cudaMemcpyAsync(tree_of_pos, stage, (size_t)d * sizeof(int), cudaMemcpyHostToDevice, st_));
            }
            Aff* T = (Aff*)tpts_.ensure((size_t)nt * nmax * sizeof(Aff));
            {
                Scope ps(this, "scatter_points", (double)nt * ((double)n + 128.0 * (double)nmax), 0.0);
                launch2d(k_scatter_points<FB>, dim3(chunks, d), 256, planes, tab, n, (uint32_t)base, (const int*)cnt, chunks, carries,
                         (const int*)tree_n, (const int*)tree_of_pos, T, nmax);
            }
            // result slot of position p is k = d-1-p; within this result handle slots are relative to the range:
            // slot = (pos_end-1-p), so slot 0 is the highest position of the range (= lowest k)
            run_trees_safe(T, nmax, cnts, flags, res, /*first slot*/ pos_end - g1, /*reverse*/ -1, roots.data() + (g0 - pos_begin));
            if (so) {  // run_trees has synchronised the compute stream: this group's functions are final, copy them on the copy stream
                for (size_t slot = pos_end - g1; slot < (size_t)(pos_end - g0); ++slot) {
                    uint8_t* dst = so->out + slot * (so->a_stride + so->b_stride) * 32;
                    if (res->la[slot]) EAGEN_CUDA(cudaMemcpyAsync(dst, res->A.as<F>() + slot * res->a_stride, (size_t)res->la[slot] * 32, cudaMemcpyDeviceToHost, cst_));
                    if (res->lb[slot]) EAGEN_CUDA(cudaMemcpyAsync(dst + so->a_stride * 32, res->B.as<F>() + slot * res->b_stride, (size_t)res->lb[slot] * 32, cudaMemcpyDeviceToHost, cst_));
                }
            }
        }
        if (so) EAGEN_CUDA(cudaStreamSynchronize(cst_));
        sync_check();
        for (uint32_t i = 0; i < npos; ++i)
            if (!roots[i].is_identity()) throw StatusError{EAGEN_E_SUM_NONZERO, "internal: a digit position's points do not sum to the identity"};
    }

    size_t pooled_bytes() const {
        return tpts_.cap + ptlev_.cap + pden_.cap + pbinv_.cap + a_[0].cap + a_[1].cap + b_[0].cap + b_[1].cap + ea_.cap + eb_.cap + oa_.cap + ob_.cap + wk_[0].cap + wk_[1].cap + den_.cap + binv_.cap + desc_.cap;
    }
    // working-set estimate per tree of n points (bytes)
    static size_t tree_bytes(size_t n) {
        size_t lc = (n + 1) / 2;
        size_t L = (size_t)ceil_log2(lc);
        size_t pad = lc + ((size_t)1 << L) + 8;  // slack for the +1 slots and the top levels
        return n * 64 + 2 * lc * 64 + pad * 32 * (4 + 8 + 2 + 3) + pad * 32 * (L + 1) + lc * (sizeof(MergeDesc<FB>) + 32 + 32 * 3 / 2) + 4096;
    }

    static F iso_gshift(uint32_t u) {   // (u^6 - 1) b
        const F uu = from_u32<FB>(u), u2 = sqr(uu), u6 = mul(sqr(u2), u2);
        return mul(sub(u6, F::one()), CC::b());
    }
    // run_trees with the collision fallback: when an output point's x-coordinate hits the evaluation domain (k_den raises
    // KERR_COLLISION) the lists are mapped to an isomorphic curve and the trees rebuilt; both the canonical and the raw form are
    // recovered exactly.
    // T is modified in place; roots come back on the original curve.
    void run_trees_safe(Aff* T, size_t cap, const std::vector<int>& cnts, uint32_t flags, ResultImpl* res, size_t first_slot, int dir, Aff* roots) {
        run_trees(T, cap, cnts, flags, res, first_slot, dir, roots, 0);
        const int nt = (int)cnts.size();
        uint32_t total_u = 1;
        static const uint32_t steps[4] = {2, 3, 5, 7};
        for (int attempt = 0; attempt < 4; ++attempt) {
            int e = last_tree_err_;   // the device flag as run_trees read it back with its lengths (no extra round trip)
            if (!(e & KERR_COLLISION)) break;
            e &= ~KERR_COLLISION;
            EAGEN_CUDA(cudaMemcpy(d_err_, &e, sizeof(int), cudaMemcpyHostToDevice));
            const F s = from_u32<FB>(steps[attempt]), s2 = sqr(s), s3 = mul(s2, s);
            total_u *= steps[attempt];
            // row 0 of the level-count table the previous run_trees call uploaded is the list length of every tree
            launch(k_iso_points<FB>, cap * (size_t)nt, 256, T, cap, (const int*)lvlcnt_.p, nt, s2, s3);
            EAGEN_CUDA(cudaStreamSynchronize(st_));
            ++iso_fallbacks_;
            run_trees(T, cap, cnts, flags, res, first_slot, dir, roots, total_u);
        }
        if (total_u != 1) {   // roots back to the original curve: (x / u^2, y / u^3)
            const F u = from_u32<FB>(total_u), iu = inv(u), iu2 = sqr(iu), iu3 = mul(iu2, iu);
            for (int tr = 0; tr < nt; ++tr)
                if (!roots[tr].is_identity()) { roots[tr].x = mul(roots[tr].x, iu2); roots[tr].y = mul(roots[tr].y, iu3); }
        }
    }

    // Divisor witnesses of `nt` point lists (T + tree*cap, counts cnts[tree]) -> res slots.
    // slot(tree) = first_slot + (dir > 0 ? tree : nt-1-tree).  roots[tree] receives the root output point.
    // iso_u != 0: T already holds the points mapped to y^2 = x^3 + u^6 b (see run_trees_safe and k_iso_points)
    void run_trees(const Aff* T, size_t cap, const std::vector<int>& cnts, uint32_t flags, ResultImpl* res, size_t first_slot, int dir, Aff* roots,
                   uint32_t iso_u = 0) {
        const int nt = (int)cnts.size();
        size_t nmax = 0;
        for (int c : cnts) nmax = std::max<size_t>(nmax, (size_t)c);
        if (nmax == 0) throw StatusError{EAGEN_E_EMPTY, "divisor witness of an empty point list"};
        if (res->nf == 0) {  // stand-alone call: size the result here
            int Lm = ceil_log2((nmax + 1) / 2);
            res->nf = nt; res->a_stride = ((size_t)1 << Lm) + 1; res->b_stride = std::max<size_t>((size_t)1 << Lm, 1);
            res->alloc(res->A, res->nf * res->a_stride * 32); res->alloc(res->B, res->nf * res->b_stride * 32);
            res->la.assign(nt, 0); res->lb.assign(nt, 0);
        }
        const size_t lc_max = (nmax + 1) / 2;
        const int L = ceil_log2(lc_max);
        if ((unsigned)L + 1 > FB::S) throw StatusError{EAGEN_E_NTT_TOO_LARGE, "tree too deep for the field's two-adicity"};
        std::vector<size_t> node_max(L + 1);
        node_max[0] = lc_max;
        for (int l = 1; l <= L; ++l) node_max[l] = (node_max[l - 1] + 1) / 2;
        // per-level, per-tree node counts; row 0 of the table is the point count itself
        std::vector<int> lv((size_t)(L + 2) * nt);
        for (int tr = 0; tr < nt; ++tr) {
            lv[tr] = cnts[tr];
            int c = (cnts[tr] + 1) / 2;
            for (int l = 0; l <= L; ++l) { lv[(size_t)(l + 1) * nt + tr] = c; c = (c + 1) / 2; }
        }
        int* dlv = (int*)lvlcnt_.ensure(lv.size() * sizeof(int));
        {
            int* stage = (int*)ring_.take(lv.size() * sizeof(int));
            std::memcpy(stage, lv.data(), lv.size() * sizeof(int));
            EAGEN_CUDA(cudaMemcpyAsync(dlv, stage, lv.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
        }
        auto cnt_of = [&](int level) { return (const int*)(dlv + (size_t)(level + 1) * nt); };

        // memory plan
        size_t pt_total = 0, szA = 0, szB = 0, szE = 0, szO = 0;
        std::vector<size_t> pt_off(L + 1);
        for (int l = 0; l <= L; ++l) {
            pt_off[l] = pt_total; pt_total += (size_t)nt * node_max[l];
            size_t m = (size_t)1 << l;
            szA = std::max(szA, (size_t)nt * node_max[l] * (m + 1));
            szB = std::max(szB, (size_t)nt * node_max[l] * m);
            if (l < L) {
                szE = std::max(szE, (size_t)nt * node_max[l] * 2 * m);
                szO = std::max(szO, (size_t)nt * node_max[l + 1] * 2 * m);
            }
        }
        Aff* PT = (Aff*)ptlev_.ensure(pt_total * sizeof(Aff));
        F* A[2] = {(F*)a_[0].ensure(szA * 32), (F*)a_[1].ensure(szA * 32)};
        F* B[2] = {(F*)b_[0].ensure(szB * 32), (F*)b_[1].ensure(szB * 32)};
        // evaluation buffers are double-buffered: the merge of level l writes the parents' evaluations straight into the
        // first half of level l+1's buffers; W is the in-place workspace of the coset and inverse transforms
        F* EA[2] = {(F*)ea_.ensure(std::max<size_t>(szE, 1) * 32), (F*)oa_.ensure(std::max<size_t>(szE, 1) * 32)};
        F* EB[2] = {(F*)eb_.ensure(std::max<size_t>(szE, 1) * 32), (F*)ob_.ensure(std::max<size_t>(szE, 1) * 32)};
        F* W[2] = {(F*)wk_[0].ensure(std::max<size_t>(szO, 1) * 32), (F*)wk_[1].ensure(std::max<size_t>(szO, 1) * 32)};
        F* TOP = (F*)top_.ensure(std::max<size_t>((size_t)nt * (L ? node_max[1] : 1), 1) * 32);
        // 1 / ((x - alpha)(x - beta)) of every evaluation point of every level: the denominators depend on the point pyramid only, so
        // the side stream computes and inverts them ahead of the polynomial levels (the latency-bound tails of 19 inversion batches and
        // their serial base inversions leave the main stream)
        std::vector<size_t> den_off(L + 1, 0);
        for (int l = 0; l < L; ++l) den_off[l + 1] = den_off[l] + (((size_t)nt * node_max[l + 1]) << (l + 1));
        F* den_all = (F*)den_.ensure(std::max<size_t>(den_off[L], 1) * 32);
        std::vector<size_t> desc_off(L + 1, 0);   // descriptors of every level are kept: the point pyramid runs ahead on its own stream
        for (int l = 0; l < L; ++l) desc_off[l + 1] = desc_off[l] + (size_t)nt * node_max[l + 1];
        MergeDesc<FB>* desc_all = (MergeDesc<FB>*)desc_.ensure(std::max<size_t>(desc_off[L], 1) * sizeof(MergeDesc<FB>));
        F* pden = (F*)pden_.ensure(std::max<size_t>((size_t)nt * node_max[0], 1) * 32);
        if (L) ensure_twiddles(L);

        // present nodes per level (exact work counts for the profiler)
        std::vector<double> present(L + 1, 0.0);
        for (int l = 0; l <= L; ++l) for (int tr = 0; tr < nt; ++tr) present[l] += lv[(size_t)(l + 1) * nt + tr];

        // raw function on the isomorphic curve: per-tree homogeneity degree k (see k_iso_points), counted by the leaf / descriptor kernels
        int* iso_deg = nullptr;
        if (iso_u && (flags & EAGEN_RAW_TREE)) {
            iso_deg = (int*)isodeg_.ensure((size_t)nt * sizeof(int));
            EAGEN_CUDA(cudaMemsetAsync(iso_deg, 0, (size_t)nt * sizeof(int), st_));
        }
        // The point pyramid -- output points -(P+Q) of the leaves, A+B of every merge, the merge descriptors (division roots, lines) --
        // depends on the lists only, never on the polynomials, and its upper levels are latency bound (a few nodes per tree, one
        // serial field inversion per batch): it runs on the side stream, ahead of the polynomial levels, with its own scratch.
        // Event 0: leaves' output points ready; event l + 1: descriptors of the merge that builds level l + 1 ready.
        const size_t w0 = (size_t)nt * node_max[0];
        {
            EAGEN_CUDA(cudaEventRecord(sync_event(0), st_));          // the lists (and the level counts) are complete on the main stream
            EAGEN_CUDA(cudaStreamWaitEvent(pst_, sync_event(0), 0));
            OnStream side(this, pst_);
            {
                Scope ps(this, "pair_points", present[0] * (128.0 + 32.0), 0.0);
                launch(k_pair_den<FB>, w0, 256, T, cap, (const int*)dlv, node_max[0], nt, pden);
            }
            batch_invert(pden, w0, pbinv_);
            {
                Scope ps(this, "pair_points", present[0] * (128.0 + 32.0 + 64.0), present[0] * 4.0);
                launch(k_pair_finish<FB>, w0, 256, T, cap, (const int*)dlv, node_max[0], nt, (const F*)pden, 1, PT + pt_off[0]);
            }
            EAGEN_CUDA(cudaEventRecord(sync_event(1), pst_));
            for (int l = 0; l < L; ++l) {
                const size_t nodes = node_max[l], merges = node_max[l + 1], wm = (size_t)nt * merges;
                Aff* Pc = PT + pt_off[l];
                Aff* Pp = PT + pt_off[l + 1];
                {
                    Scope ps(this, "pair_points", present[l + 1] * (128.0 + 32.0), 0.0);
                    launch(k_pair_den<FB>, wm, 256, (const Aff*)Pc, nodes, cnt_of(l), merges, nt, pden);
                }
                batch_invert(pden, wm, pbinv_);
                {
                    Scope ps(this, "pair_points", present[l + 1] * (128.0 + 32.0 + 64.0 + 192.0 + (double)sizeof(MergeDesc<FB>)), present[l + 1] * (4.0 + 2.0));
                    launch(k_pair_finish<FB>, wm, 256, (const Aff*)Pc, nodes, cnt_of(l), merges, nt, (const F*)pden, 0, Pp);
                    launch(k_merge_desc<FB>, wm, 128, (const Aff*)Pc, nodes, cnt_of(l), (const Aff*)Pp, merges, nt, desc_all + desc_off[l], iso_deg);
                }
                EAGEN_CUDA(cudaEventRecord(sync_event((size_t)l + 2), pst_));
                // ... and straight away the inverse denominators of that merge (event L + 2 + l), so level l of the polynomial loop
                // never waits for more of the pyramid than it needs
                const int t = l + 1;
                const double pts = present[l + 1] * (double)((size_t)1 << t);
                prof_level_ = l;
                {
                    Scope ps(this, "merge_den", pts * 64.0, pts * 1.0);
                    launch(k_den<FB>, wm << t, 256, (const MergeDesc<FB>*)(desc_all + desc_off[l]), wm, t, xtab(t), den_all + den_off[l], d_err_);
                }
                batch_invert(den_all + den_off[l], wm << t, pbinv_);
                prof_level_ = -1;
                EAGEN_CUDA(cudaEventRecord(sync_event((size_t)L + 2 + l), pst_));
            }
        }
        // level 0 on the main stream: the line functions of the leaves
        EAGEN_CUDA(cudaStreamWaitEvent(st_, sync_event(1), 0));
        {
            Scope ps(this, "pair_points", present[0] * (128.0 + 64.0 + 96.0), present[0] * 2.0);
            // ... and, with them, the leaves' evaluations on the 2-point domain (no separate 2-point transform)
            launch(k_leaf_lines<FB>, w0, 256, T, cap, (const int*)dlv, (const Aff*)(PT + pt_off[0]), node_max[0], nt, A[0], B[0],
                   L > 0 ? EA[0] : (F*)nullptr, L > 0 ? EB[0] : (F*)nullptr, iso_deg);
        }

        int cur = 0, e = 0;
        for (int l = 0; l < L; ++l) {
            prof_level_ = l;
            struct LevelReset { int& v; ~LevelReset() { v = -1; } } level_reset{prof_level_};
            const int t = l + 1;
            const size_t m = (size_t)1 << l, Tn = 2 * m;
            const size_t nodes = node_max[l], merges = node_max[l + 1];
            const size_t wm = (size_t)nt * merges;
            const double pts = present[l + 1] * (double)Tn;  // evaluation points of this level
            const MergeDesc<FB>* desc = desc_all + desc_off[l];
            // Children in the evaluation domain.  Positions [0, m) of each child's 2m-point buffer already hold its values on the
            // m-point domain (written by the previous level's merge); only the odd coset w_T * w_m^k is transformed here:
            // a(x) = sum_{i<m} c_i x^i + c_m x^m and x^m = -1 on the coset, so values = NTT_m(c_i w_T^i) - c_m.
            // Stored coefficients carry the factor s = m of the unscaled inverse transform; the twist table removes it.
            if (l > 0) {
                NttJob jf[2] = {{W[0], A[cur], m + 1, (int)m, EA[e], Tn, (int)m, (const F*)TOP}, {W[1], B[cur], m, (int)m, EB[e], Tn, (int)m, nullptr}};
                ntt(false, jf, 2, l, (size_t)nt * nodes, cnt_of(l), (int)nodes, (size_t)present[l], twist_tab(l), m);
            }
            // pointwise merge with exact division; parents' evaluations go to the next level's buffers (stride 2T)
            EAGEN_CUDA(cudaStreamWaitEvent(st_, sync_event((size_t)L + 2 + l), 0));   // this level's descriptors and inverse denominators (side stream)
            {
                const F* den = den_all + den_off[l];
                Scope ps(this, "merge_pointwise", pts * 288.0, pts * 11.0);
                if (iso_u)
                    launch(k_pointwise<CC, true>, wm << t, 128, (const MergeDesc<FB>*)desc, wm, t, xtab(t), gtab(t), (const F*)EA[e], (const F*)EB[e],
                           den, merges, nodes, EA[e ^ 1], EB[e ^ 1], 2 * Tn, iso_gshift(iso_u));
                else
                    launch(k_pointwise<CC, false>, wm << t, 128, (const MergeDesc<FB>*)desc, wm, t, xtab(t), gtab(t), (const F*)EA[e], (const F*)EB[e],
                           den, merges, nodes, EA[e ^ 1], EB[e ^ 1], 2 * Tn, F::zero());
            }
            // back to coefficients (unscaled: stored = T * true), compact parent slots
            {
                NttJob ji[2] = {{W[0], EA[e ^ 1], 2 * Tn, (int)Tn, A[cur ^ 1], Tn + 1, (int)Tn, nullptr}, {W[1], EB[e ^ 1], 2 * Tn, (int)Tn, B[cur ^ 1], Tn, (int)Tn, nullptr}};
                ntt(true, ji, 2, t, wm, cnt_of(l + 1), (int)merges, (size_t)present[l + 1]);
            }
            {
                Scope ps(this, "merge_fixup", present[l + 1] * (6 * 32.0 + 96.0), present[l + 1] * 7.0);
                F kc = l == 0 ? F::one() : half_pow(2 * l);
                launch(k_fixup<FB>, wm, 128, (const MergeDesc<FB>*)desc, wm, t, merges, nodes, (const F*)A[cur], (const F*)B[cur], A[cur ^ 1],
                       kc, from_u32<FB>((uint32_t)Tn), TOP);
            }
            cur ^= 1; e ^= 1;
        }
        // roots: one node per tree at level L (node_max[L] == 1)
        const size_t ra = ((size_t)1 << L) + 1, rb = (size_t)1 << L;
        int* tops = (int*)tops_.ensure((size_t)2 * nt * sizeof(int));
        EAGEN_CUDA(cudaMemsetAsync(tops, 0, (size_t)2 * nt * sizeof(int), st_));
        Scope ps_can(this, "canonical_form", (double)nt * (double)(ra + rb) * 32.0 * ((flags & EAGEN_RAW_TREE) ? 1.0 : 3.0),
                     (flags & EAGEN_RAW_TREE) ? 0.0 : (double)nt * (double)(ra + rb));
        if ((flags & EAGEN_RAW_TREE) && L > 0) {  // stored root = 2^L * true coefficients
            launch2d(k_scale_const<FB>, dim3((unsigned)((ra + 255) / 256), nt), 256, A[cur], ra, (int)ra, half_pow(L));
            launch2d(k_scale_const<FB>, dim3((unsigned)((rb + 255) / 256), nt), 256, B[cur], rb, (int)rb, half_pow(L));
        }
        if (iso_u) {  // back to the original curve: a_i = a'_i u^(2i), b_i = b'_i u^(2i+3) (any common factor disappears in the monic form)
            const F u = from_u32<FB>(iso_u), u2 = sqr(u), u3 = mul(u2, u);
            const F* per_tree = nullptr;
            if (iso_deg) {   // raw form: divide tree tr by u^k_tr
                std::vector<int> hk(nt);
                EAGEN_CUDA(cudaMemcpyAsync(hk.data(), iso_deg, (size_t)nt * sizeof(int), cudaMemcpyDeviceToHost, st_));
                EAGEN_CUDA(cudaStreamSynchronize(st_));
                const F iu = inv(u);
                std::vector<F> fac(nt);
                for (int tr = 0; tr < nt; ++tr) {
                    F r = F::one(), w = iu;
                    for (uint32_t e = (uint32_t)hk[tr]; e; e >>= 1) { if (e & 1) r = mul(r, w); w = sqr(w); }
                    fac[tr] = r;
                }
                F* dfac = (F*)lead_.ensure((size_t)std::max(nt, 1) * 32);
                EAGEN_CUDA(cudaMemcpyAsync(dfac, fac.data(), (size_t)nt * 32, cudaMemcpyHostToDevice, st_));
                EAGEN_CUDA(cudaStreamSynchronize(st_));   // `fac` is a stack temporary
                per_tree = dfac;
            }
            launch2d(k_iso_unscale<FB>, dim3((unsigned)((ra + 255) / 256), nt), 256, A[cur], ra, (int)ra, u2, F::one(), per_tree);
            launch2d(k_iso_unscale<FB>, dim3((unsigned)((rb + 255) / 256), nt), 256, B[cur], rb, (int)rb, u2, u3, per_tree);
        }
        launch2d(k_find_top<FB>, dim3((unsigned)((ra + 255) / 256), nt), 256, (const F*)A[cur], ra, (int)ra, tops);
        launch2d(k_find_top<FB>, dim3((unsigned)((rb + 255) / 256), nt), 256, (const F*)B[cur], rb, (int)rb, tops + nt);
        if (!(flags & EAGEN_RAW_TREE)) {
            F* lead = (F*)lead_.ensure((size_t)std::max(nt, 1) * 32);
            launch(k_lead<FB>, nt, 64, (const F*)A[cur], ra, (const int*)tops, (const F*)B[cur], rb, (const int*)(tops + nt), nt, lead);
            batch_invert(lead, nt);
            launch2d(k_scale<FB>, dim3((unsigned)((ra + 255) / 256), nt), 256, (const F*)A[cur], ra, (const int*)tops, (const F*)lead,
                     res->A.as<F>(), res->a_stride, first_slot, dir);
            launch2d(k_scale<FB>, dim3((unsigned)((rb + 255) / 256), nt), 256, (const F*)B[cur], rb, (const int*)(tops + nt), (const F*)lead,
                     res->B.as<F>(), res->b_stride, first_slot, dir);
        }
        // scatter into the result slots
        std::vector<int> htops((size_t)2 * nt);
        EAGEN_CUDA(cudaMemcpyAsync(htops.data(), tops, htops.size() * sizeof(int), cudaMemcpyDeviceToHost, st_));
        EAGEN_CUDA(cudaMemcpyAsync(roots, PT + pt_off[L], (size_t)nt * sizeof(Aff), cudaMemcpyDeviceToHost, st_));
        if (flags & EAGEN_RAW_TREE) {   // canonical form: k_scale has already written the slots
            launch2d(k_copy_strided<FB>, dim3((unsigned)((ra + 255) / 256), nt), 256, (const F*)A[cur], ra, res->A.as<F>(), res->a_stride, (int)ra, first_slot, dir);
            launch2d(k_copy_strided<FB>, dim3((unsigned)((rb + 255) / 256), nt), 256, (const F*)B[cur], rb, res->B.as<F>(), res->b_stride, (int)rb, first_slot, dir);
        }
        int* herr = (int*)ring_.take(sizeof(int));
        EAGEN_CUDA(cudaMemcpyAsync(herr, d_err_, sizeof(int), cudaMemcpyDeviceToHost, st_));
        EAGEN_CUDA(cudaStreamSynchronize(st_));
        last_tree_err_ = *herr;
        for (int tr = 0; tr < nt; ++tr) {
            size_t slot = first_slot + (dir > 0 ? (size_t)tr : (size_t)(nt - 1 - tr));
            res->la[slot] = htops[tr]; res->lb[slot] = htops[nt + tr];
        }
    }
};

IEngine* make_engine_pallas(int device);
IEngine* make_engine_vesta(int device);
IEngine* make_engine_grumpkin(int device);

}  // namespace eagen
