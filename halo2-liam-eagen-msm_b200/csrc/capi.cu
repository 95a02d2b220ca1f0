// extern "C" surface of libeagen_msm.so (declared in include/eagen_msm.h).  Plain pointers and sizes only.
// There is no CPU fallback anywhere below: without a CUDA device eagen_ctx_create fails with EAGEN_E_NO_DEVICE.
#include "engine.cuh"

using namespace eagen;

struct eagen_ctx {
    int curve;
    IEngine* eng;
    std::string err;
    bool comm_from_init_all = false;   // communicator created by eagen_comm_init_all: destroyed by eagen_comm_destroy / ctx_destroy
    void* raw_comm = nullptr;
};
struct eagen_result {
    ResultImpl* r;
};

namespace {
thread_local std::string g_global_err;

template <class Fn>
int guarded(eagen_ctx* ctx, Fn fn) {
    try {
        fn();
        return EAGEN_OK;
    } catch (const StatusError& e) {
        if (ctx) ctx->err = e.msg; else g_global_err = e.msg;
        return e.code;
    } catch (const CudaError& e) {
        if (ctx) ctx->err = e.msg; else g_global_err = e.msg;
        cudaGetLastError();
        return e.code;
    } catch (const std::exception& e) {
        if (ctx) ctx->err = e.what(); else g_global_err = e.what();
        return EAGEN_E_ARG;
    }
}
void need(bool ok, const char* what) { if (!ok) throw StatusError{EAGEN_E_ARG, what}; }

template <class FS> uint32_t digits_for(uint8_t base) { return make_negbase_params<FS>(base).d; }
uint32_t digits_of_curve(int curve, uint8_t base) {
    switch (curve) {
        case EAGEN_CURVE_PALLAS: return digits_for<Pallas::Scalar>(base);
        case EAGEN_CURVE_VESTA: return digits_for<Vesta::Scalar>(base);
        case EAGEN_CURVE_GRUMPKIN: return digits_for<Grumpkin::Scalar>(base);
    }
    throw StatusError{EAGEN_E_ARG, "unknown curve id"};
}
template <class FP> void precomp(int which, uint64_t e, uint64_t* out) {
    Fe<FP> r;
    if (which == 0 || which == 1) {
        // omega_pow(k) = ROOT_OF_UNITY^(2^k): the identity from k = S on, so larger exponents need no more squarings
        unsigned k = (unsigned)std::min<uint64_t>(e, FP::S);
        r = pow2k(which == 0 ? Fe<FP>::root_of_unity() : Fe<FP>::root_of_unity_inv(), k);
    } else {   // (1/2)^e by square-and-multiply (a 64-bit exponent must not become a 2^64-step loop)
        r = Fe<FP>::one();
        Fe<FP> b = Fe<FP>::two_inv();
        for (uint64_t x = e; x; x >>= 1) { if (x & 1) r = mul(r, b); b = sqr(b); }
    }
    std::memcpy(out, r.v, 32);
}
// table_entry_by_id restated on the host with the kernels' field code (reference: src/negbase_utils.rs:58-77)
template <class FP> void table_entry(uint8_t base, size_t id, uint64_t* out) {
    Fe<FP> acc = Fe<FP>::zero();
    if (id != 0) {
        Fe<FP> b = neg(from_u32<FP>(base));
        int l = 0;
        while ((id >> l) != 0) ++l;
        for (int i = l - 1; i >= 0; --i) {
            if ((id >> i) & 1) acc = add(acc, Fe<FP>::one());
            acc = mul(acc, b);
        }
    }
    std::memcpy(out, acc.v, 32);
}
// ---- challenge-point helpers of the circuit (reference: src/config.rs:163-187), host side like the reference's ------------
// a^e for a 256-bit exponent given as limbs
template <class FP> Fe<FP> pow_limbs(const Fe<FP>& a, const uint32_t* e) {
    Fe<FP> r = Fe<FP>::one();
    for (int i = 7; i >= 0; --i)
        for (int bit = 31; bit >= 0; --bit) {
            r = sqr(r);
            if ((e[i] >> bit) & 1) r = mul(r, a);
        }
    return r;
}
// Tonelli-Shanks with z = F::ROOT_OF_UNITY (order 2^S): the root x = a^((t+1)/2) * z^e that ff::helpers::sqrt_tonelli_shanks
// computes (p - 1 = 2^S t).  Returns false for a non-residue.
template <class FP> bool sqrt_ts(const Fe<FP>& a, Fe<FP>& root) {
    if (a.is_zero()) { root = a; return true; }
    uint32_t t[8], pm1[8];
    for (int i = 0; i < 8; ++i) pm1[i] = FP::mod(i);
    pm1[0] -= 1;   // p is odd
    const unsigned S = FP::S, ws = S / 32, bs = S % 32;
    for (int i = 0; i < 8; ++i) {
        uint64_t lo = ((unsigned)i + ws < 8) ? pm1[i + ws] : 0, hi = ((unsigned)i + ws + 1 < 8) ? pm1[i + ws + 1] : 0;
        t[i] = bs ? (uint32_t)((lo >> bs) | (hi << (32 - bs))) : (uint32_t)lo;
    }
    uint32_t e[8];   // (t - 1) / 2 = t >> 1 (t is odd)
    for (int i = 0; i < 8; ++i) e[i] = (t[i] >> 1) | (i + 1 < 8 ? (t[i + 1] << 31) : 0);
    Fe<FP> w = pow_limbs(a, e);
    Fe<FP> x = mul(a, w), b = mul(x, w), z = Fe<FP>::root_of_unity();
    unsigned v = S;
    const Fe<FP> one = Fe<FP>::one();
    while (!(b == one)) {
        unsigned k = 0;
        Fe<FP> b2 = b;
        while (!(b2 == one)) { b2 = sqr(b2); ++k; if (k == v) return false; }
        Fe<FP> zz = z;
        for (unsigned i = 0; i + k + 1 < v; ++i) zz = sqr(zz);
        x = mul(x, zz);
        z = sqr(zz);
        b = mul(b, z);
        v = k;
    }
    root = x;
    return true;
}
// x^3 + a x + b of the curve (a = 0 for all three curves)
template <class CC> Fe<typename CC::Base> curve_rhs(const Fe<typename CC::Base>& x) { return add(mul(sqr(x), x), CC::b()); }
// which: 0 to_curve_x, 1 y_from_x, 2 slope
template <class CC> int challenge_op(int which, const uint64_t* in, uint64_t* out, int* flag) {
    typedef Fe<typename CC::Base> F;
    F x, y;
    std::memcpy(x.v, in, 32);
    if (which == 2) {
        std::memcpy(y.v, in + 4, 32);
        if (y.is_zero()) throw StatusError{EAGEN_E_DOMAIN, "slope: y = 0 (the reference's invert().unwrap() panics)"};
        F three_x2 = mul(from_u32<typename CC::Base>(3), sqr(x));
        F r = mul(three_x2, inv(dbl(y)));
        std::memcpy(out, r.v, 32);
        return 0;
    }
    F rhs = curve_rhs<CC>(x), root;
    bool sq = sqrt_ts(rhs, root);
    if (flag) *flag = sq ? 1 : 0;
    if (which == 0) {
        // the reference loops forever when c is not the x of a curve point (the loop never changes x, :170-173)
        if (!sq) throw StatusError{EAGEN_E_DOMAIN, "to_curve_x: x^3 + b is not a square (the reference never returns here)"};
        std::memcpy(out, x.v, 32);
        return 0;
    }
    if (!sq) {   // sqrt_alt of a non-square: (false, sqrt(ROOT_OF_UNITY * value))
        if (!sqrt_ts(mul(rhs, F::root_of_unity()), root)) throw StatusError{EAGEN_E_ARG, "sqrt_alt: internal error"};
    }
    std::memcpy(out, root.v, 32);
    return 0;
}
int challenge_dispatch(int curve, int which, const uint64_t* in, uint64_t* out, int* flag) {
    switch (curve) {
        case EAGEN_CURVE_PALLAS: return challenge_op<Pallas>(which, in, out, flag);
        case EAGEN_CURVE_VESTA: return challenge_op<Vesta>(which, in, out, flag);
        case EAGEN_CURVE_GRUMPKIN: return challenge_op<Grumpkin>(which, in, out, flag);
    }
    throw StatusError{EAGEN_E_ARG, "unknown curve id"};
}
}  // namespace

extern "C" {

const char* eagen_status_string(int s) {
    switch (s) {
        case EAGEN_OK: return "ok";
        case EAGEN_E_ARG: return "invalid argument";
        case EAGEN_E_LEN: return "incompatible amount of coefficients";
        case EAGEN_E_RANGE: return "scalar out of range";
        case EAGEN_E_SUM_NONZERO: return "points do not sum to the identity";
        case EAGEN_E_NTT_TOO_LARGE: return "transform larger than the field's two-adicity";
        case EAGEN_E_CUDA: return "CUDA error";
        case EAGEN_E_NCCL: return "NCCL error";
        case EAGEN_E_DIGITS: return "negbase expansion longer than d digits";
        case EAGEN_E_DOMAIN: return "point on the evaluation domain";
        case EAGEN_E_NO_DEVICE: return "no CUDA device (no CPU fallback)";
        case EAGEN_E_EMPTY: return "empty point list";
    }
    return "unknown status";
}

int eagen_ctx_create(int curve, int device, eagen_ctx** out) {
    if (!out) return EAGEN_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); g_global_err = "no CUDA device visible"; return EAGEN_E_NO_DEVICE; }
    if (device < 0 || device >= ndev) { g_global_err = "device index out of range"; return EAGEN_E_ARG; }
    eagen_ctx* c = new eagen_ctx();
    c->curve = curve; c->eng = nullptr;
    int rc = guarded(nullptr, [&] {
        switch (curve) {
            case EAGEN_CURVE_PALLAS: c->eng = make_engine_pallas(device); break;
            case EAGEN_CURVE_VESTA: c->eng = make_engine_vesta(device); break;
            case EAGEN_CURVE_GRUMPKIN: c->eng = make_engine_grumpkin(device); break;
            default: throw StatusError{EAGEN_E_ARG, "unknown curve id"};
        }
    });
    if (rc != EAGEN_OK) { delete c; return rc; }
    *out = c;
    return EAGEN_OK;
}

void eagen_ctx_destroy(eagen_ctx* ctx) {
    if (!ctx) return;
    eagen_comm_destroy(ctx);
    delete ctx->eng;
    delete ctx;
}

const char* eagen_last_error(const eagen_ctx* ctx) { return ctx ? ctx->err.c_str() : g_global_err.c_str(); }
uint64_t eagen_launch_count(const eagen_ctx* ctx) { return ctx ? ctx->eng->launches() : 0; }
uint64_t eagen_fallback_count(const eagen_ctx* ctx) { return ctx ? ctx->eng->iso_fallbacks() : 0; }

int eagen_microbench(eagen_ctx* ctx, int which, double* ops_per_second) {
    if (!ctx || !ops_per_second || which < 0 || which > 2) return EAGEN_E_ARG;
    return guarded(ctx, [&] { *ops_per_second = ctx->eng->microbench(which); });
}
int eagen_set_profiling(eagen_ctx* ctx, int on) {
    if (!ctx) return EAGEN_E_ARG;
    ctx->eng->set_profiling(on);
    return EAGEN_OK;
}
int eagen_profile_reset(eagen_ctx* ctx) {
    if (!ctx) return EAGEN_E_ARG;
    ctx->eng->profile_reset();
    return EAGEN_OK;
}
int eagen_profile_json(eagen_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return EAGEN_E_ARG;
    std::string j = ctx->eng->profile_json();
    if (j.size() + 1 > cap) return EAGEN_E_LEN;
    std::memcpy(buf, j.c_str(), j.size() + 1);
    return EAGEN_OK;
}

int eagen_num_digits(int curve, uint8_t base, uint32_t* d) {
    return guarded(nullptr, [&] { need(d && base >= 2, "eagen_num_digits: null output or base < 2"); *d = digits_of_curve(curve, base); });
}

int eagen_table_entry_by_id(int curve, uint8_t base, size_t id, uint64_t* out) {
    return guarded(nullptr, [&] {
        need(out != nullptr, "eagen_table_entry_by_id: null output");
        switch (curve) {
            case EAGEN_CURVE_PALLAS: table_entry<Pallas::Base>(base, id, out); break;
            case EAGEN_CURVE_VESTA: table_entry<Vesta::Base>(base, id, out); break;
            case EAGEN_CURVE_GRUMPKIN: table_entry<Grumpkin::Base>(base, id, out); break;
            default: throw StatusError{EAGEN_E_ARG, "unknown curve id"};
        }
    });
}

int eagen_fft_precomp(int curve, int which, uint64_t exp, uint64_t* out) {
    return guarded(nullptr, [&] {
        need(out && which >= 0 && which <= 2, "eagen_fft_precomp: bad arguments");
        switch (curve) {
            case EAGEN_CURVE_PALLAS: precomp<Pallas::Base>(which, exp, out); break;
            case EAGEN_CURVE_VESTA: precomp<Vesta::Base>(which, exp, out); break;
            case EAGEN_CURVE_GRUMPKIN: precomp<Grumpkin::Base>(which, exp, out); break;
            default: throw StatusError{EAGEN_E_ARG, "unknown curve id"};
        }
    });
}

int eagen_negbase_decompose(eagen_ctx* ctx, const uint64_t* scalars, size_t n, uint8_t base, uint8_t* digits) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] { need(base >= 2 && (n == 0 || (scalars && digits)), "eagen_negbase_decompose: bad arguments"); ctx->eng->negbase_host(scalars, n, base, digits); });
}

int eagen_precompute_multiplicities(eagen_ctx* ctx, const uint64_t* pts, size_t n, uint8_t base, uint64_t* out) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] { need(base >= 2 && (n == 0 || (pts && out)), "eagen_precompute_multiplicities: bad arguments"); ctx->eng->multiples_host(pts, n, base, out); });
}

int eagen_lhs_witness(eagen_ctx* ctx, const uint64_t* scalars, const uint64_t* pts, size_t n, uint8_t base, uint32_t flags, eagen_result** out) {
    if (!ctx || !out) return EAGEN_E_ARG;
    *out = nullptr;
    return guarded(ctx, [&] {
        need(base >= 2 && (n == 0 || (scalars && pts)), "eagen_lhs_witness: bad arguments");
        *out = new eagen_result{ctx->eng->lhs_host(scalars, pts, n, base, flags)};
    });
}

int eagen_lhs_witness_stream_layout(int curve, size_t n, uint8_t base, size_t* a_stride, size_t* b_stride, size_t* total_bytes) {
    return guarded(nullptr, [&] {
        need(a_stride && b_stride && base >= 2, "eagen_lhs_witness_stream_layout: bad arguments");
        stream_slot_elems(n, base, a_stride, b_stride);
        if (total_bytes) *total_bytes = (size_t)digits_of_curve(curve, base) * (*a_stride + *b_stride) * 32;
    });
}
int eagen_lhs_witness_stream(eagen_ctx* ctx, const uint64_t* scalars, const uint64_t* pts, size_t n, uint8_t base, uint32_t flags,
                             void* out, size_t out_bytes, eagen_result** res) {
    if (!ctx || !res) return EAGEN_E_ARG;
    *res = nullptr;
    return guarded(ctx, [&] {
        need(base >= 2 && out && (n == 0 || (scalars && pts)), "eagen_lhs_witness_stream: bad arguments");
        *res = new eagen_result{ctx->eng->lhs_stream_host(scalars, pts, n, base, flags, out, out_bytes)};
    });
}

int eagen_dev_lhs_witness(eagen_ctx* ctx, const void* d_scalars, const void* d_pts, size_t n, uint8_t base, uint32_t flags, eagen_result** out) {
    if (!ctx || !out) return EAGEN_E_ARG;
    *out = nullptr;
    return guarded(ctx, [&] {
        need(base >= 2 && (n == 0 || (d_scalars && d_pts)), "eagen_dev_lhs_witness: bad arguments");
        *out = new eagen_result{ctx->eng->lhs_dev(d_scalars, d_pts, n, base, flags)};
    });
}

int eagen_divisor_witness(eagen_ctx* ctx, const uint64_t* pts, size_t n, uint32_t flags, uint64_t* out_point, eagen_result** out) {
    if (!ctx || !out) return EAGEN_E_ARG;
    *out = nullptr;
    return guarded(ctx, [&] {
        need(n == 0 || pts, "eagen_divisor_witness: null points");
        *out = new eagen_result{ctx->eng->divisor_host(pts, n, flags, out_point)};
    });
}

int eagen_dev_shard_sums(eagen_ctx* ctx, const void* d_scalars, const void* d_pts, size_t n, uint8_t base, void* d_planes, void* d_table, void* d_partial_sums) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        need(base >= 2 && d_partial_sums && (n == 0 || (d_scalars && d_pts && d_planes && d_table)), "eagen_dev_shard_sums: bad arguments");
        ctx->eng->shard_sums_dev(d_scalars, d_pts, n, base, d_planes, d_table, d_partial_sums);
    });
}
int eagen_dev_negbase(eagen_ctx* ctx, const void* d_scalars, size_t n, uint8_t base, void* d_planes, void* d_rows, double* device_ms) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        need(base >= 2 && (n == 0 || (d_scalars && d_planes)), "eagen_dev_negbase: bad arguments");
        double ms = ctx->eng->negbase_dev(d_scalars, n, base, d_planes, d_rows);
        if (device_ms) *device_ms = ms;
    });
}
int eagen_dev_ntt(eagen_ctx* ctx, void* d_data, uint32_t log_n, size_t batch, int inverse, double* device_ms) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        need(d_data != nullptr, "eagen_dev_ntt: null buffer");
        double ms = ctx->eng->ntt_dev(d_data, log_n, batch, inverse);
        if (device_ms) *device_ms = ms;
    });
}
int eagen_dev_carry_chain(eagen_ctx* ctx, const void* d_partial_sums, int nparts, uint8_t base, void* d_carries) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        need(base >= 2 && d_partial_sums && d_carries && nparts >= 1, "eagen_dev_carry_chain: bad arguments");
        ctx->eng->carry_chain_dev(d_partial_sums, nparts, base, d_carries);
    });
}
int eagen_dev_trees(eagen_ctx* ctx, const void* d_planes, const void* d_table, const void* d_carries, size_t n, uint8_t base,
                    uint32_t pos_begin, uint32_t pos_end, uint32_t flags, eagen_result** out) {
    if (!ctx || !out) return EAGEN_E_ARG;
    *out = nullptr;
    return guarded(ctx, [&] {
        need(base >= 2 && d_carries && (n == 0 || (d_planes && d_table)), "eagen_dev_trees: bad arguments");
        *out = new eagen_result{ctx->eng->trees_dev(d_planes, d_table, d_carries, n, base, pos_begin, pos_end, flags)};
    });
}

// ---- multi-GPU behind the boundary ---------------------------------------------------
int eagen_comm_unique_id(void* id_out) {
    return guarded(nullptr, [&] {
        need(id_out != nullptr, "eagen_comm_unique_id: null output");
        NcclApi& nc = NcclApi::get();
        if (!nc.ok) throw StatusError{EAGEN_E_NCCL, nc.load_error};
        ncclUniqueId id;
        ncclResult_t r = nc.GetUniqueId(&id);
        if (r != ncclSuccess) throw StatusError{EAGEN_E_NCCL, std::string("ncclGetUniqueId: ") + nc.GetErrorString(r)};
        static_assert(sizeof(ncclUniqueId) == EAGEN_COMM_ID_BYTES, "unique id size");
        std::memcpy(id_out, &id, sizeof id);
    });
}
int eagen_comm_init(eagen_ctx* ctx, int nranks, int rank, const void* unique_id) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] { ctx->eng->comm_init_rank(nranks, rank, unique_id); });
}
int eagen_comm_init_all(eagen_ctx** ctxs, int n) {
    if (!ctxs || n < 1) return EAGEN_E_ARG;
    for (int i = 0; i < n; ++i) if (!ctxs[i]) return EAGEN_E_ARG;
    return guarded(ctxs[0], [&] {
        NcclApi& nc = NcclApi::get();
        if (!nc.ok) throw StatusError{EAGEN_E_NCCL, nc.load_error};
        std::vector<int> devs(n);
        for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->eng->device();
        std::vector<ncclComm_t> comms(n);
        ncclResult_t r = nc.CommInitAll(comms.data(), n, devs.data());
        if (r != ncclSuccess) throw StatusError{EAGEN_E_NCCL, std::string("ncclCommInitAll: ") + nc.GetErrorString(r)};
        for (int i = 0; i < n; ++i) { ctxs[i]->eng->comm_attach(comms[i], n, i); ctxs[i]->comm_from_init_all = true; ctxs[i]->raw_comm = comms[i]; }
    });
}
int eagen_comm_destroy(eagen_ctx* ctx) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        ctx->eng->comm_destroy();
        if (ctx->comm_from_init_all && ctx->raw_comm) { NcclApi::get().CommDestroy((ncclComm_t)ctx->raw_comm); }
        ctx->comm_from_init_all = false; ctx->raw_comm = nullptr;
    });
}
int eagen_comm_size(const eagen_ctx* ctx) { return ctx ? ctx->eng->comm_size() : 0; }
int eagen_comm_rank(const eagen_ctx* ctx) { return ctx ? ctx->eng->comm_rank() : -1; }
int eagen_position_range(int rank, int nranks, uint32_t d, uint32_t* begin, uint32_t* end) {
    if (!begin || !end || nranks < 1 || rank < 0 || rank >= nranks) return EAGEN_E_ARG;
    position_range(rank, nranks, d, begin, end);
    return EAGEN_OK;
}
int eagen_lhs_witness_sharded(eagen_ctx* ctx, const uint64_t* scalars_local, const uint64_t* pts_local, size_t n_local, uint8_t base, uint32_t flags,
                              void* out, size_t out_bytes, eagen_result** res) {
    if (!ctx || !res) return EAGEN_E_ARG;
    *res = nullptr;
    return guarded(ctx, [&] {
        need(base >= 2 && (n_local == 0 || (scalars_local && pts_local)), "eagen_lhs_witness_sharded: bad arguments");
        *res = new eagen_result{ctx->eng->lhs_sharded(scalars_local, pts_local, false, n_local, base, flags, out, out_bytes)};
    });
}
int eagen_dev_lhs_witness_sharded(eagen_ctx* ctx, const void* d_scalars_local, const void* d_pts_local, size_t n_local, uint8_t base, uint32_t flags,
                                  eagen_result** res) {
    if (!ctx || !res) return EAGEN_E_ARG;
    *res = nullptr;
    return guarded(ctx, [&] {
        need(base >= 2 && (n_local == 0 || (d_scalars_local && d_pts_local)), "eagen_dev_lhs_witness_sharded: bad arguments");
        *res = new eagen_result{ctx->eng->lhs_sharded(d_scalars_local, d_pts_local, true, n_local, base, flags, nullptr, 0)};
    });
}
int eagen_lhs_witness_sharded_layout(int curve, size_t n_total, uint8_t base, int rank, int nranks, size_t* a_stride, size_t* b_stride, size_t* total_bytes) {
    return guarded(nullptr, [&] {
        need(a_stride && b_stride && base >= 2 && nranks >= 1 && rank >= 0 && rank < nranks, "eagen_lhs_witness_sharded_layout: bad arguments");
        stream_slot_elems(n_total, base, a_stride, b_stride);
        uint32_t b0, b1;
        position_range(rank, nranks, digits_of_curve(curve, base), &b0, &b1);
        if (total_bytes) *total_bytes = (size_t)(b1 - b0) * (*a_stride + *b_stride) * 32;
    });
}
int eagen_ctx_set_stream_split(eagen_ctx* ctx, const uint32_t* percent, int n) {
    if (!ctx || n < 0 || (n > 0 && !percent)) return EAGEN_E_ARG;
    ctx->eng->set_stream_split(percent, n);
    return EAGEN_OK;
}
size_t eagen_result_first_function(const eagen_result* r) { return r ? r->r->k0 : 0; }

int eagen_synth_inputs(eagen_ctx* ctx, uint64_t seed, size_t n, uint64_t* scalars, uint64_t* pts) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] { need(n == 0 || (scalars && pts), "eagen_synth_inputs: null buffer"); ctx->eng->synth_host(seed, n, scalars, pts); });
}
int eagen_dev_synth_inputs(eagen_ctx* ctx, uint64_t seed, size_t n, void* d_scalars, void* d_pts) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] { need(n == 0 || (d_scalars && d_pts), "eagen_dev_synth_inputs: null buffer"); ctx->eng->synth_dev(seed, n, d_scalars, d_pts); });
}

int eagen_msm(eagen_ctx* ctx, const uint64_t* scalars, const uint64_t* pts, size_t n, uint64_t* out_affine, double* device_ms) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        need(out_affine && (n == 0 || (scalars && pts)), "eagen_msm: null buffer");
        double ms = ctx->eng->msm_host(scalars, pts, n, out_affine);
        if (device_ms) *device_ms = ms;
    });
}

int eagen_prepare_scalar_witness(eagen_ctx* ctx, const uint64_t* scalars, size_t n, uint8_t base, uint32_t num_digits, uint32_t logtable,
                                 int mode, void* out, size_t out_bytes) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        need(base >= 2 && logtable >= 1 && (n == 0 || (scalars && out)), "eagen_prepare_scalar_witness: bad arguments");
        size_t num_limbs = ((size_t)num_digits + logtable - 1) / logtable;
        if (out_bytes < n * (size_t)base * (num_limbs + 1) * 32) throw StatusError{EAGEN_E_LEN, "eagen_prepare_scalar_witness: output buffer too small"};
        ctx->eng->scalar_witness_host(scalars, n, base, num_digits, logtable, mode, out);
    });
}
int eagen_divisor_witness_naive(eagen_ctx* ctx, const uint64_t* pts, size_t n, uint64_t* pos_lines, size_t* n_pos, uint64_t* neg_lines, size_t* n_neg) {
    if (!ctx || !n_pos || !n_neg) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        need(n == 0 || (pts && pos_lines && neg_lines), "eagen_divisor_witness_naive: null buffer");
        ctx->eng->naive_host(pts, n, pos_lines, n_pos, neg_lines, n_neg);
    });
}

// ---- circuit-facing layouts ---------------------------------------------------
int eagen_circuit_sizes(size_t num_pts, uint8_t base, size_t* a_size, size_t* b_size) {
    if (!a_size || !b_size || base < 2) return EAGEN_E_ARG;
    *b_size = (num_pts + base + 1) / 2;   // src/config.rs:641
    *a_size = (num_pts + base + 2) / 2;   // src/config.rs:642
    return EAGEN_OK;
}
int eagen_result_copy_padded(eagen_result* r, size_t num_pts, uint8_t base, uint64_t* a_out, uint64_t* b_out) {
    if (!r || base < 2) return EAGEN_E_ARG;
    ResultImpl* x = r->r;
    size_t a_size, b_size;
    eagen_circuit_sizes(num_pts, base, &a_size, &b_size);
    if (x->nf == 0) return EAGEN_OK;
    if (!a_out || !b_out) return EAGEN_E_ARG;
    for (size_t k = 0; k < x->nf; ++k)
        if ((size_t)x->la[k] > a_size || (size_t)x->lb[k] > b_size) return EAGEN_E_LEN;
    if (!x->A.p || !x->B.p) return EAGEN_E_ARG;
    cudaSetDevice(x->device);
    std::memset(a_out, 0, x->nf * a_size * 32);
    std::memset(b_out, 0, x->nf * b_size * 32);
    for (size_t k = 0; k < x->nf; ++k) {
        size_t la = (size_t)x->la[k] * 32, lb = (size_t)x->lb[k] * 32;
        if (la && cudaMemcpyAsync((char*)a_out + k * a_size * 32, (const char*)x->A.p + k * x->a_stride * 32, la, cudaMemcpyDeviceToHost, 0) != cudaSuccess) return EAGEN_E_CUDA;
        if (lb && cudaMemcpyAsync((char*)b_out + k * b_size * 32, (const char*)x->B.p + k * x->b_stride * 32, lb, cudaMemcpyDeviceToHost, 0) != cudaSuccess) return EAGEN_E_CUDA;
    }
    return cudaStreamSynchronize(0) == cudaSuccess ? EAGEN_OK : EAGEN_E_CUDA;
}
int eagen_result_eval(eagen_ctx* ctx, eagen_result* r, const uint64_t* pts, size_t m, uint64_t* out) {
    if (!ctx || !r) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        need(m == 0 || (pts && out), "eagen_result_eval: null buffer");
        need(r->r->nf == 0 || (r->r->A.p && r->r->B.p), "eagen_result_eval: the result holds no device-resident functions");
        ctx->eng->result_eval_host(r->r, pts, m, out);
    });
}
int eagen_to_curve_x(int curve, const uint64_t* c, uint64_t* x_out) {
    return guarded(nullptr, [&] { need(c && x_out, "eagen_to_curve_x: null buffer"); challenge_dispatch(curve, 0, c, x_out, nullptr); });
}
int eagen_y_from_x(int curve, const uint64_t* x, uint64_t* y_out, int* is_square) {
    return guarded(nullptr, [&] { need(x && y_out, "eagen_y_from_x: null buffer"); challenge_dispatch(curve, 1, x, y_out, is_square); });
}
int eagen_slope(int curve, const uint64_t* xy, uint64_t* slope_out) {
    return guarded(nullptr, [&] { need(xy && slope_out, "eagen_slope: null buffer"); challenge_dispatch(curve, 2, xy, slope_out, nullptr); });
}

int eagen_poly_mul(eagen_ctx* ctx, const uint64_t* a, size_t la, const uint64_t* b, size_t lb, uint64_t* out) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] { need((la == 0 || a) && (lb == 0 || b) && (la + lb <= 1 || out), "eagen_poly_mul: null buffer"); ctx->eng->poly_mul_host(a, la, b, lb, out); });
}
int eagen_ntt(eagen_ctx* ctx, uint64_t* data, uint32_t log_n, int inverse) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] { need(data != nullptr, "eagen_ntt: null buffer"); ctx->eng->ntt_host(data, log_n, inverse); });
}
int eagen_batch_invert(eagen_ctx* ctx, uint64_t* elems, size_t n) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] { need(n == 0 || elems, "eagen_batch_invert: null buffer"); ctx->eng->batch_invert_host(elems, n); });
}
int eagen_eval_function(eagen_ctx* ctx, const uint64_t* a, size_t la, const uint64_t* b, size_t lb, const uint64_t* pts, size_t n, uint64_t* out) {
    if (!ctx) return EAGEN_E_ARG;
    return guarded(ctx, [&] {
        need((la == 0 || a) && (lb == 0 || b) && (n == 0 || (pts && out)), "eagen_eval_function: null buffer");
        ctx->eng->eval_host(a, la, b, lb, pts, n, out);
    });
}

// ---- result accessors ------------------------------------------------------------------------------------
uint32_t eagen_result_num_digits(const eagen_result* r) { return r ? r->r->d : 0; }
size_t eagen_result_num_functions(const eagen_result* r) { return r ? r->r->nf : 0; }
size_t eagen_result_poly_len(const eagen_result* r, size_t k, int which) {
    if (!r || k >= r->r->nf) return 0;
    return (size_t)(which == EAGEN_POLY_A ? r->r->la[k] : r->r->lb[k]);
}
int eagen_result_poly_copy(eagen_result* r, size_t k, int which, uint64_t* out) {
    if (!r || k >= r->r->nf || (which != EAGEN_POLY_A && which != EAGEN_POLY_B)) return EAGEN_E_ARG;
    ResultImpl* x = r->r;
    size_t len = (size_t)(which == EAGEN_POLY_A ? x->la[k] : x->lb[k]);
    if (len == 0) return EAGEN_OK;
    if (!out) return EAGEN_E_ARG;
    cudaSetDevice(x->device);
    const char* src = (const char*)(which == EAGEN_POLY_A ? x->A.p : x->B.p) + k * (which == EAGEN_POLY_A ? x->a_stride : x->b_stride) * 32;
    return cudaMemcpy(out, src, len * 32, cudaMemcpyDeviceToHost) == cudaSuccess ? EAGEN_OK : EAGEN_E_CUDA;
}
size_t eagen_result_total_bytes(const eagen_result* r) {
    if (!r) return 0;
    size_t t = 0;
    for (size_t k = 0; k < r->r->nf; ++k) t += ((size_t)r->r->la[k] + (size_t)r->r->lb[k]) * 32;
    return t;
}
int eagen_result_copy_all(eagen_result* r, uint64_t* out, size_t out_bytes, size_t* written) {
    if (!r || !out) return EAGEN_E_ARG;
    ResultImpl* x = r->r;
    if (out_bytes < eagen_result_total_bytes(r)) return EAGEN_E_LEN;
    cudaSetDevice(x->device);
    char* dst = (char*)out;
    for (size_t k = 0; k < x->nf; ++k) {
        size_t la = (size_t)x->la[k] * 32, lb = (size_t)x->lb[k] * 32;
        if (la && cudaMemcpyAsync(dst, (const char*)x->A.p + k * x->a_stride * 32, la, cudaMemcpyDeviceToHost, 0) != cudaSuccess) return EAGEN_E_CUDA;
        dst += la;
        if (lb && cudaMemcpyAsync(dst, (const char*)x->B.p + k * x->b_stride * 32, lb, cudaMemcpyDeviceToHost, 0) != cudaSuccess) return EAGEN_E_CUDA;
        dst += lb;
    }
    if (cudaStreamSynchronize(0) != cudaSuccess) return EAGEN_E_CUDA;
    if (written) *written = (size_t)(dst - (char*)out);
    return EAGEN_OK;
}
int eagen_result_carry(eagen_result* r, uint64_t* out) {
    if (!r || !out) return EAGEN_E_ARG;
    std::memcpy(out, r->r->carry, 64);
    return EAGEN_OK;
}
int eagen_result_carries(eagen_result* r, uint64_t* out) {
    if (!r || !out) return EAGEN_E_ARG;
    std::memcpy(out, r->r->carries.data(), r->r->carries.size() * 8);
    return EAGEN_OK;
}
int eagen_result_digits(eagen_result* r, uint8_t* out) {
    if (!r || !out || !r->r->has_digits) return EAGEN_E_ARG;
    cudaSetDevice(r->r->device);
    size_t bytes = r->r->n * r->r->d;
    if (bytes == 0) return EAGEN_OK;
    return cudaMemcpy(out, r->r->digits.p, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? EAGEN_OK : EAGEN_E_CUDA;
}
double eagen_result_device_ms(const eagen_result* r) { return r ? r->r->device_ms : 0.0; }
int eagen_result_device_view(eagen_result* r, const void** d_a, size_t* a_stride, const void** d_b, size_t* b_stride) {
    if (!r) return EAGEN_E_ARG;
    if (d_a) *d_a = r->r->A.p;
    if (a_stride) *a_stride = r->r->a_stride;
    if (d_b) *d_b = r->r->B.p;
    if (b_stride) *b_stride = r->r->b_stride;
    return EAGEN_OK;
}
void eagen_result_free(eagen_result* r) {
    if (!r) return;
    cudaSetDevice(r->r->device);
    delete r->r;
    delete r;
}

}  // extern "C"

// ---- host-side self-test hooks (include/eagen_msm_selftest.h): the HD arithmetic run on the CPU ------------
#include "../../include/eagen_msm_selftest.h"
namespace {
template <class FP> void st_field(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    Fe<FP> x, y = Fe<FP>::zero(), r;
    std::memcpy(x.v, a, 32);
    if (b) std::memcpy(y.v, b, 32);
    switch (op) {
        case 0: r = add(x, y); break;
        case 1: r = sub(x, y); break;
        case 2: r = mul(x, y); break;
        case 3: r = inv(x); break;
        case 4: r = from_canonical(x); break;
        case 5: r = to_canonical(x); break;
        case 6: r = mul_chain(x, y); break;  // the device algorithm with a host-emulated carry flag
        // lazily reduced arithmetic of the transform (operands in [0, 2p), twiddle y < p for the product); results are returned as
        // they are, i.e. in [0, 2p), except op 10
        case 7: r = mul_lazy(x, y); break;
        case 8: r = add_lazy(x, y); break;
        case 9: r = sub_lazy(x, y); break;
        case 10: r = normalise_lazy(x); break;
        default: throw StatusError{EAGEN_E_ARG, "bad op"};
    }
    std::memcpy(out, r.v, 32);
}
template <class CC> void st_curve(int op, const uint64_t* p, const uint64_t* q, uint32_t k, uint64_t* out) {
    typedef typename CC::Base F;
    auto load = [](const uint64_t* s) { Fe<F> x, y, z; std::memcpy(x.v, s, 32); std::memcpy(y.v, s + 4, 32); std::memcpy(z.v, s + 8, 32); return jacobian_to_proj<CC>(x, y, z); };
    Proj<F> a = load(p), r;
    switch (op) {
        case 0: r = padd<CC>(a, load(q)); break;
        case 1: r = pdbl<CC>(a); break;
        case 2: { Affine<F> qa; std::memcpy(qa.x.v, q, 32); std::memcpy(qa.y.v, q + 4, 32); r = padd_mixed<CC>(a, qa); break; }
        case 3: r = pmul_small<CC>(a, k); break;
        default: throw StatusError{EAGEN_E_ARG, "bad op"};
    }
    Affine<F> o = proj_to_affine(r, inv(r.z));
    std::memcpy(out, &o, 64);
}
template <class FS> void st_nb(uint8_t base, uint32_t* d, uint32_t* group, uint32_t* words, uint32_t* limbs) {
    NegbaseParams p = make_negbase_params<FS>(base);
    *d = p.d; *group = p.g; *words = p.nw;
    std::memcpy(limbs, p.sq, 32); std::memcpy(limbs + 8, p.K, 32); std::memcpy(limbs + 16, p.bd, 32);
    std::memset(limbs + 24, 0, 32); std::memcpy(limbs + 24, p.inv, 24);
}
// K1's per-scalar arithmetic (from_mont + negbase_words + the table) run on the host from the kernel's own source
template <class FS> int st_nb_digits(uint8_t base, const uint64_t* scalar, uint8_t* digits) {
    NegbaseParams p = make_negbase_params<FS>(base);
    std::vector<uint32_t> lut(p.lut_n);
    for (uint32_t v = 0; v < p.lut_n; ++v) lut[v] = negbase_lut_entry(v, p.base, p.g);
    Fe<FS> xm; std::memcpy(xm.v, scalar, 32);
    uint32_t x[8], words[NEGBASE_MAX_WORDS];
    from_mont<FS>(xm.v, x);
    int e = negbase_words(x, p, lut.data(), words, 1);
    for (uint32_t w = 0; w < p.nw; ++w)
        for (int k = 0; k < 4; ++k) { int pos = (int)(4 * w + k) - (int)p.pad; if (pos >= 0) digits[pos] = (uint8_t)(words[w] >> (8 * k)); }
    return e;
}
}  // namespace

extern "C" {
int eagen_selftest_field(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    return guarded(nullptr, [&] {
        need(a && out, "null");
        switch (field) {
            case 0: st_field<PallasFp>(op, a, b, out); break;
            case 1: st_field<PallasFq>(op, a, b, out); break;
            case 2: st_field<Bn256Fr>(op, a, b, out); break;
            case 3: st_field<Bn256Fq>(op, a, b, out); break;
            default: throw StatusError{EAGEN_E_ARG, "unknown field id"};
        }
    });
}
int eagen_selftest_curve(int curve, int op, const uint64_t* p, const uint64_t* q, uint32_t k, uint64_t* out) {
    return guarded(nullptr, [&] {
        need(p && out, "null");
        switch (curve) {
            case EAGEN_CURVE_PALLAS: st_curve<Pallas>(op, p, q, k, out); break;
            case EAGEN_CURVE_VESTA: st_curve<Vesta>(op, p, q, k, out); break;
            case EAGEN_CURVE_GRUMPKIN: st_curve<Grumpkin>(op, p, q, k, out); break;
            default: throw StatusError{EAGEN_E_ARG, "unknown curve id"};
        }
    });
}
int eagen_selftest_negbase_params(int curve, uint8_t base, uint32_t* d, uint32_t* chunk, uint32_t* cd, uint32_t* limbs) {
    return guarded(nullptr, [&] {
        need(d && chunk && cd && limbs && base >= 2, "bad arguments");
        switch (curve) {
            case EAGEN_CURVE_PALLAS: st_nb<Pallas::Scalar>(base, d, chunk, cd, limbs); break;
            case EAGEN_CURVE_VESTA: st_nb<Vesta::Scalar>(base, d, chunk, cd, limbs); break;
            case EAGEN_CURVE_GRUMPKIN: st_nb<Grumpkin::Scalar>(base, d, chunk, cd, limbs); break;
            default: throw StatusError{EAGEN_E_ARG, "unknown curve id"};
        }
    });
}
int eagen_selftest_negbase_digits(int curve, uint8_t base, const uint64_t* scalar, uint8_t* digits, int* kerr) {
    return guarded(nullptr, [&] {
        need(scalar && digits && kerr && base >= 2, "bad arguments");
        switch (curve) {
            case EAGEN_CURVE_PALLAS: *kerr = st_nb_digits<Pallas::Scalar>(base, scalar, digits); break;
            case EAGEN_CURVE_VESTA: *kerr = st_nb_digits<Vesta::Scalar>(base, scalar, digits); break;
            case EAGEN_CURVE_GRUMPKIN: *kerr = st_nb_digits<Grumpkin::Scalar>(base, scalar, digits); break;
            default: throw StatusError{EAGEN_E_ARG, "unknown curve id"};
        }
    });
}
int eagen_selftest_ntt_plan(int t, int* pairs) {
    if (!pairs || t < 1 || t > 40) return EAGEN_E_ARG;
    auto v = ntt_plan(t);
    for (size_t i = 0; i < v.size() && i < 8; ++i) { pairs[2 * i] = v[i].first; pairs[2 * i + 1] = v[i].second; }
    return (int)v.size();
}
}
