// Hand-written sm_100a kernels for the Liam-Eagen MSM witness hot path.
//
// Kernel map (K-numbers as in SURVEY.md section 2.2), reference semantics each one replaces:
//   K1  k_negbase            negbase_decompose + pad + reverse     src/negbase_utils.rs:20-36, src/argument_witness_calc.rs:93-101
//   K2  k_multiples_*        precompute_multiplicities (affine)    src/argument_witness_calc.rs:43-51,103
//   K3  k_digit_sums, k_reduce_partials   the carry += mult[j][digit] loop   src/argument_witness_calc.rs:120-125
//   K4  k_sum_parts, k_carry_chain   carry = (-carry)*base + S_i   src/argument_witness_calc.rs:105-127
//                            (k_sum_parts folds the ranks' partial sums of a multi-GPU call per position first)
//   K5  k_pair_den/finish, k_leaf_lines, k_merge_desc   from_pair / from_point / linefunc / output points
//                                                                    src/regular_functions_utils.rs:285-331,335
//                            (the whole point pyramid and the inverse denominators run on the engine's side stream)
//   K6  k_ntt_pass           Polynomial::mul_fft -> best_fft        src/regular_functions_utils.rs:102-129
//                            (both polynomial families of a level per launch, lazily reduced butterflies, absent tiles skipped;
//                             k_gen_twiddles / k_gen_points / k_gen_twist_all build the per-context tables)
//   K7  k_den, k_pointwise, k_fixup   RegularFunction::mul, Propagation::merge, kate_div   :266-273,333-360,45-47
//   K9  k_binv_*             z.invert() (batched, Montgomery trick) :351-352
//   K10 k_find_top, k_lead, k_scale   trim + monic canonical form    SURVEY.md section 8c
//
// All kernels are integer-pipe (IMAD/IADD3) or HBM bound; none is a dense contraction, so no tcgen05.
#pragma once
#include <cuda_runtime.h>
#include "curve.cuh"

namespace eagen {

// ------------------------------------------------------------------------------------------------
// 128-bit vectorised element access
// ------------------------------------------------------------------------------------------------
template <class FP>
EAGEN_D Fe<FP> ldg(const Fe<FP>* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fe<FP> r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class FP>
EAGEN_D void stg(Fe<FP>* p, const Fe<FP>& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <class FP>
EAGEN_D Affine<FP> ldg_aff(const Affine<FP>* p) {
    Affine<FP> a;
    a.x = ldg(&p->x);
    a.y = ldg(&p->y);
    return a;
}
template <class FP>
EAGEN_D void stg_aff(Affine<FP>* p, const Affine<FP>& a) {
    stg(&p->x, a.x);
    stg(&p->y, a.y);
}

// error flag bits written by kernels (device int, OR-ed)
enum : int {
    KERR_RANGE = 1,        // scalar >= isqrt(order)+2            (reference: src/argument_witness_calc.rs:97)
    KERR_DIGITS = 2,       // negbase expansion needs more than d digits (reference truncates silently, :99)
    KERR_COLLISION = 4,    // an output point's x-coordinate is a 2^k-th root of unity of the evaluation domain
};

// ------------------------------------------------------------------------------------------------
// K9  batched inversion (Montgomery trick), hierarchical: each thread owns G strided elements
// ------------------------------------------------------------------------------------------------
#ifndef EAGEN_BINV_G
#define EAGEN_BINV_G 16
#endif
constexpr int BINV_G = EAGEN_BINV_G;

template <class FP>
__global__ void k_binv_up(const Fe<FP>* __restrict__ x, Fe<FP>* __restrict__ pref, Fe<FP>* __restrict__ tot, size_t M, size_t Tn) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tn) return;
    Fe<FP> acc = Fe<FP>::one();
#pragma unroll 1
    for (int j = 0; j < BINV_G; ++j) {
        size_t idx = t + (size_t)j * Tn;
        if (idx >= M) break;
        Fe<FP> v = ldg(x + idx);
        if (!v.is_zero()) acc = mul(acc, v);
        stg(pref + idx, acc);
    }
    stg(tot + t, acc);
}

template <class FP>
__global__ void k_binv_down(Fe<FP>* __restrict__ x, const Fe<FP>* __restrict__ pref, const Fe<FP>* __restrict__ tot, size_t M, size_t Tn) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tn) return;
    Fe<FP> it = ldg(tot + t);
#pragma unroll 1
    for (int j = BINV_G - 1; j >= 0; --j) {
        size_t idx = t + (size_t)j * Tn;
        if (idx >= M) continue;
        Fe<FP> v = ldg(x + idx);
        if (v.is_zero()) continue;
        Fe<FP> prev = j > 0 ? ldg(pref + idx - Tn) : Fe<FP>::one();
        stg(x + idx, mul(it, prev));
        it = mul(it, v);
    }
}

template <class FP>
__global__ void k_binv_base(Fe<FP>* x, size_t M) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    stg(x + t, inv(ldg(x + t)));
}

// ------------------------------------------------------------------------------------------------
// K1  negabase digits.  x = sum d_i (-b)^i  <=>  the ordinary base-b digits e_i of y = x + K with
// K = sum_{odd i<d} (b-1) b^i satisfy d_i = e_i (i even), d_i = b-1-e_i (i odd).
//
// The base-b digits of y are produced WITHOUT any division, most significant first (the order the planes are stored in):
// G = floor(y * ceil(2^288 / b^d) / 2^128) + 1 is a 160-bit fixed-point image of y / b^d whose error lies in (0, b^-d), so the
// integer part of G * b^k is exactly floor(y / b^(d-k)) for every k; every step multiplies the fraction by b^g (five
// 32x32->64 products) and the overflow limb is the value of the next g digits.  Groups of g = 4 (b <= 5), 2 (b <= 31) digits are
// split with a shared-memory table that already holds the complemented bytes; larger bases pop one digit per step.
// Four digit positions are packed per 32-bit word, the block transposes the words in shared memory and writes the
// position-major planes with 32-bit stores of four neighbouring scalars (a warp covers 128 contiguous bytes of four rows).
// ------------------------------------------------------------------------------------------------
struct NegbaseParams {
    uint32_t sq[8];      // isqrt(order)+2 (canonical limbs)
    uint32_t K[8];       // offset constant
    uint32_t bd[8];      // b^d
    uint32_t inv[6];     // ceil(2^288 / b^d)  (<= 2^160 because b^d >= 2^128)
    uint32_t pw[5];      // b^0 .. b^4
    uint32_t base, d;
    uint32_t g;          // digits per table group: 4, 2, or 1 (no table)
    uint32_t lut_n;      // b^g table entries (0 when g == 1)
    uint32_t nw;         // words of four digit positions: 4*nw = d + pad
    uint32_t pad;        // leading positions of word 0 that do not exist (their digits are zero)
};

constexpr int NEGBASE_THREADS = 128;
constexpr int NEGBASE_MAX_WORDS = 36;   // d <= 144

EAGEN_HD bool lt8(const uint32_t* a, const uint32_t* b) {
    for (int i = 7; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] < b[i];
    }
    return false;
}

// Montgomery residue -> canonical integer: eight reduction rows and no products by the operand (a * 1), so the zero and
// power-of-two limbs of the Pasta moduli fold away at compile time.
template <class FP>
EAGEN_HD void from_mont(const uint32_t* a, uint32_t* t) {
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t m = t[0] * FP::INV;
        uint64_t c = (uint64_t)m * FP::mod(0) + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            c += (uint64_t)m * FP::mod(j) + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        t[7] = (uint32_t)c;
    }
    reduce_once<FP>(t, 0);
}

// table entry for the group value v < b^g: g bytes, most significant digit in byte 0; inside a word of four positions the
// bytes at even offsets belong to odd digit indices (4 | d + pad) and are stored complemented
EAGEN_HD uint32_t negbase_lut_entry(uint32_t v, uint32_t base, uint32_t g) {
    uint32_t w = 0;
    for (uint32_t k = g; k-- > 0;) {
        uint32_t e = v % base;
        v /= base;
        uint32_t dg = (k & 1) ? e : (base - 1 - e);
        w |= dg << (8 * k);
    }
    return w;
}

// The NW words (four digit positions each, MSD first, byte 0 = first position) of one scalar's canonical value x.
// Word w goes to out[w * stride]; `lut` may be null when prm.g == 1.  Returns 0 or a KERR_* flag (words are zeroed then).
EAGEN_HD int negbase_words(const uint32_t* x, const NegbaseParams& prm, const uint32_t* lut, uint32_t* out, uint32_t stride) {
    if (!lt8(x, prm.sq)) { for (uint32_t w = 0; w < prm.nw; ++w) out[w * stride] = 0u; return KERR_RANGE; }
    uint32_t y[8];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c += (uint64_t)x[i] + prm.K[i]; y[i] = (uint32_t)c; c >>= 32; }
    if (!lt8(y, prm.bd)) { for (uint32_t w = 0; w < prm.nw; ++w) out[w * stride] = 0u; return KERR_DIGITS; }
    // y < b^d < 2^143: five limbs.  P = y * inv < 2^288 (limbs 9 and 10 end up zero), G = P[4..8] + 1
    uint32_t P[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) P[i] = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        uint64_t cy = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            cy += (uint64_t)y[i] * prm.inv[j] + P[i + j];
            P[i + j] = (uint32_t)cy;
            cy >>= 32;
        }
        P[i + 6] = (uint32_t)cy;
    }
    uint32_t G[5];
    c = 1;
#pragma unroll
    for (int i = 0; i < 5; ++i) { c += P[4 + i]; G[i] = (uint32_t)c; c >>= 32; }
    const uint32_t g = prm.g, base = prm.base, slots = 4 / g;
    for (uint32_t w = 0; w < prm.nw; ++w) {
        uint32_t word = 0;
        for (uint32_t s = 0; s < slots; ++s) {
            // positions 4w + s*g .. + g-1 of the padded expansion; r of them exist
            int r = (int)(4 * w + (s + 1) * g) - (int)prm.pad;
            r = r < 0 ? 0 : (r > (int)g ? (int)g : r);
            uint32_t v = 0;
            if (r > 0) {
                const uint32_t m = prm.pw[r];
                uint64_t cy = 0;
#pragma unroll
                for (int l = 0; l < 5; ++l) { cy += (uint64_t)G[l] * m; G[l] = (uint32_t)cy; cy >>= 32; }
                v = (uint32_t)cy;
            }
            uint32_t bytes;
            if (g == 1) bytes = (s & 1) ? v : (base - 1 - v);
            else bytes = lut[v];
            word |= bytes << (8 * g * s);
        }
        out[w * stride] = word;
    }
    return 0;
}

// One thread per scalar for the arithmetic, then a block-wide transpose through shared memory (see the header above).
// vec != 0: n % 4 == 0 and the planes are 4-byte aligned, so the 32-bit plane stores are legal; otherwise bytes are stored.
template <class FS>
__global__ void __launch_bounds__(NEGBASE_THREADS)
k_negbase(const Fe<FS>* __restrict__ scalars, size_t n, NegbaseParams prm, uint8_t* __restrict__ planes /* d x n */,
          uint8_t* __restrict__ rows /* n x d or null */, int vec, int* err) {
    extern __shared__ uint32_t nb_sm[];
    uint32_t* W = nb_sm;                                   // nw x NEGBASE_THREADS words
    uint32_t* lut = nb_sm + prm.nw * NEGBASE_THREADS;      // lut_n entries
    for (uint32_t v = threadIdx.x; v < prm.lut_n; v += NEGBASE_THREADS) lut[v] = negbase_lut_entry(v, prm.base, prm.g);
    __syncthreads();
    // persistent blocks: the table (b^g entries, built with real divisions) is paid once per block, not once per 128 scalars
    const size_t nchunks = (n + NEGBASE_THREADS - 1) / NEGBASE_THREADS;
    for (size_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const size_t j0 = chunk * NEGBASE_THREADS;
    const size_t j = j0 + threadIdx.x;
    if (j < n) {
        Fe<FS> xm = ldg(scalars + j);
        uint32_t x[8];
        from_mont<FS>(xm.v, x);
        int e = negbase_words(x, prm, lut, W + threadIdx.x, NEGBASE_THREADS);
        if (e) atomicOr(err, e);
    } else {
        for (uint32_t w = 0; w < prm.nw; ++w) W[w * NEGBASE_THREADS + threadIdx.x] = 0;
    }
    __syncthreads();
    const uint32_t d = prm.d, pad = prm.pad;
    if (vec) {
        const uint32_t quads = NEGBASE_THREADS / 4;
        for (uint32_t it = threadIdx.x; it < prm.nw * quads; it += NEGBASE_THREADS) {
            const uint32_t w = it / quads, q = it - w * quads;
            const size_t jq = j0 + 4 * q;
            if (jq >= n) continue;
            const uint4 xw = *reinterpret_cast<const uint4*>(W + w * NEGBASE_THREADS + 4 * q);
            // 4 x 4 byte transpose: o[k] = byte k of the four scalars' words
            const uint32_t t0 = __byte_perm(xw.x, xw.y, 0x5140), t1 = __byte_perm(xw.z, xw.w, 0x5140);
            const uint32_t t2 = __byte_perm(xw.x, xw.y, 0x7362), t3 = __byte_perm(xw.z, xw.w, 0x7362);
            const uint32_t o[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632), __byte_perm(t2, t3, 0x5410), __byte_perm(t2, t3, 0x7632)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int pos = (int)(4 * w + k) - (int)pad;
                if (pos >= 0) *reinterpret_cast<uint32_t*>(planes + (size_t)pos * n + jq) = o[k];
            }
        }
    } else {
        for (uint32_t it = threadIdx.x; it < prm.nw * NEGBASE_THREADS; it += NEGBASE_THREADS) {
            const uint32_t w = it / NEGBASE_THREADS, t = it - w * NEGBASE_THREADS;
            if (j0 + t >= n) continue;
            const uint32_t word = W[w * NEGBASE_THREADS + t];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int pos = (int)(4 * w + k) - (int)pad;
                if (pos >= 0) planes[(size_t)pos * n + j0 + t] = (uint8_t)(word >> (8 * k));
            }
        }
    }
    if (rows && j < n) {
        for (uint32_t w = 0; w < prm.nw; ++w) {
            const uint32_t word = W[w * NEGBASE_THREADS + threadIdx.x];
            for (int k = 0; k < 4; ++k) {
                const int pos = (int)(4 * w + k) - (int)pad;
                if (pos >= 0) rows[j * d + pos] = (uint8_t)(word >> (8 * k));
            }
        }
    }
    __syncthreads();   // W is rewritten by the next chunk
    }
}

// ------------------------------------------------------------------------------------------------
// K2  small multiples [P, 2P, ..., (b-1)P], affine after one batched inversion of the Z's
// ------------------------------------------------------------------------------------------------
template <class CC>
__global__ void k_multiples_proj(const Fe<typename CC::Base>* __restrict__ jac /* n x 3 */, size_t n, uint32_t base,
                                 Affine<typename CC::Base>* __restrict__ table /* n x (b-1): X,Y for now */,
                                 Fe<typename CC::Base>* __restrict__ zs) {
    typedef typename CC::Base F;
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Proj<F> p = jacobian_to_proj<CC>(ldg(jac + 3 * j), ldg(jac + 3 * j + 1), ldg(jac + 3 * j + 2));
    Proj<F> acc = p;
    for (uint32_t k = 1; k < base; ++k) {
        size_t o = j * (base - 1) + (k - 1);
        stg(&table[o].x, acc.x);
        stg(&table[o].y, acc.y);
        stg(zs + o, acc.z);
        if (k + 1 < base) acc = padd<CC>(acc, p);
    }
}

// (X, Y) *= 1/Z ; Z == 0 -> identity (0,0).  zs holds the inverses (0 where Z was 0).
template <class FP>
__global__ void k_scale_by_zinv(Affine<FP>* __restrict__ pts, const Fe<FP>* __restrict__ zinv, size_t m) {
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= m) return;
    Fe<FP> zi = ldg(zinv + o);
    Affine<FP> a;
    if (zi.is_zero()) a = Affine<FP>::identity();
    else { a.x = mul(ldg(&pts[o].x), zi); a.y = mul(ldg(&pts[o].y), zi); }
    stg_aff(pts + o, a);
}

// Jacobian (X,Y,Z) -> affine (X/Z^2, Y/Z^3), two kernels around one batched inversion
template <class FP>
__global__ void k_jac_z(const Fe<FP>* __restrict__ jac, size_t n, Fe<FP>* __restrict__ zs) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) stg(zs + j, ldg(jac + 3 * j + 2));
}
template <class FP>
__global__ void k_jac_to_affine(const Fe<FP>* __restrict__ jac, const Fe<FP>* __restrict__ zinv, size_t n, Affine<FP>* __restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Fe<FP> zi = ldg(zinv + j);
    Affine<FP> a;
    if (zi.is_zero()) a = Affine<FP>::identity();
    else {
        Fe<FP> zi2 = sqr(zi);
        a.x = mul(ldg(jac + 3 * j), zi2);
        a.y = mul(ldg(jac + 3 * j + 1), mul(zi2, zi));
    }
    stg_aff(out + j, a);
}

// ------------------------------------------------------------------------------------------------
// K3  per digit position i:  S_i = sum_j table[j][D[i][j]-1]   (complete mixed additions)
// grid = (chunks, d); each thread walks `per_thread` points strided by blockDim
// ------------------------------------------------------------------------------------------------
constexpr int SUMS_THREADS = 128;

#ifndef EAGEN_SUMS_MINBLOCKS
#define EAGEN_SUMS_MINBLOCKS 5
#endif
template <class CC>
__global__ void __launch_bounds__(SUMS_THREADS, EAGEN_SUMS_MINBLOCKS)
k_digit_sums(const uint8_t* __restrict__ planes, const Affine<typename CC::Base>* __restrict__ table, size_t n, uint32_t base,
             int per_thread, Proj<typename CC::Base>* __restrict__ partials /* d x chunks */) {
    typedef typename CC::Base F;
    __shared__ Proj<F> sm[SUMS_THREADS];
    const uint32_t pos = blockIdx.y;
    const size_t chunk0 = (size_t)blockIdx.x * SUMS_THREADS * per_thread;
    Proj<F> acc = Proj<F>::identity();
    const uint8_t* dg = planes + (size_t)pos * n;
    for (int k = 0; k < per_thread; ++k) {
        size_t j = chunk0 + (size_t)k * SUMS_THREADS + threadIdx.x;
        if (j >= n) break;
        uint32_t dv = dg[j];
        if (dv == 0) continue;
        Affine<F> q = ldg_aff(table + j * (base - 1) + (dv - 1));
        if (q.is_identity()) continue;
        acc = padd_mixed<CC>(acc, q);
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = SUMS_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = padd<CC>(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[(size_t)pos * gridDim.x + blockIdx.x] = sm[0];
}

// one block per digit position: reduce `count` partial sums
template <class CC>
__global__ void __launch_bounds__(SUMS_THREADS)
k_reduce_partials(const Proj<typename CC::Base>* __restrict__ partials, int count, Proj<typename CC::Base>* __restrict__ out) {
    typedef typename CC::Base F;
    __shared__ Proj<F> sm[SUMS_THREADS];
    const uint32_t pos = blockIdx.x;
    Proj<F> acc = Proj<F>::identity();
    for (int k = threadIdx.x; k < count; k += SUMS_THREADS) acc = padd<CC>(acc, partials[(size_t)pos * count + k]);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = SUMS_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = padd<CC>(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[pos] = sm[0];
}

// per-rank partial digit sums (multi-GPU all-gather: sums[part*d + i]) -> one sum per position, all positions in parallel, so the
// serial chain below adds ONE point per step whatever the number of ranks
template <class CC>
__global__ void k_sum_parts(const Proj<typename CC::Base>* __restrict__ sums, uint32_t d, int nparts, Proj<typename CC::Base>* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    Proj<typename CC::Base> acc = sums[i];
    for (int part = 1; part < nparts; ++part) acc = padd<CC>(acc, sums[(size_t)part * d + i]);
    out[i] = acc;
}

// K4  Horner chain in base (-b): carry_i = b * (-carry_{i-1}) + S_i ; serial in i (d steps)
template <class CC>
__global__ void k_carry_chain(const Proj<typename CC::Base>* __restrict__ sums, uint32_t d, uint32_t base, int nparts,
                              Proj<typename CC::Base>* __restrict__ carries /* d */, Fe<typename CC::Base>* __restrict__ zs /* d */) {
    typedef typename CC::Base F;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Proj<F> carry = Proj<F>::identity();
    int hb = 31 - __clz(base);
    for (uint32_t i = 0; i < d; ++i) {
        Proj<F> nc = pneg<CC>(carry);
        Proj<F> acc = nc;  // top bit of base
        for (int bit = hb - 1; bit >= 0; --bit) {
            acc = pdbl<CC>(acc);
            if ((base >> bit) & 1) acc = padd<CC>(acc, nc);
        }
        // sums may come as several per-rank partials (multi-GPU all-gather): sums[part*d + i]
        for (int part = 0; part < nparts; ++part) acc = padd<CC>(acc, sums[(size_t)part * d + i]);
        carry = acc;
        carries[i] = carry;
        stg(zs + i, carry.z);
    }
}

template <class FP>
__global__ void k_proj_to_affine(const Proj<FP>* __restrict__ p, const Fe<FP>* __restrict__ zinv, size_t n, Affine<FP>* __restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Proj<FP> q = p[j];
    stg_aff(out + j, proj_to_affine(q, ldg(zinv + j)));
}

// ------------------------------------------------------------------------------------------------
// Building tmp_i (reference: src/argument_witness_calc.rs:110-127) for every digit position at once:
//   T_i = [-carry_{i-1}] x base (if carry_{i-1} != O)  ++  [table[j][D[i][j]-1] : D[i][j] != 0, j ascending]  ++  [-carry_i]
// stream compaction = count per 1024-point chunk, scan of chunk counts, scatter.
// ------------------------------------------------------------------------------------------------
constexpr int CHUNK_PTS = 1024;

static __global__ void k_count_nonzero(const uint8_t* __restrict__ planes, size_t n, int chunks, int* __restrict__ cnt /* d x chunks */) {
    const uint32_t pos = blockIdx.y;
    const size_t j0 = (size_t)blockIdx.x * CHUNK_PTS;
    int c = 0;
    for (int k = threadIdx.x; k < CHUNK_PTS; k += blockDim.x) {
        size_t j = j0 + k;
        if (j < n && planes[(size_t)pos * n + j] != 0) ++c;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int sm[32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w];
        cnt[(size_t)pos * chunks + blockIdx.x] = s;
    }
}

// one thread per digit position: exclusive scan of its chunk counts (chunks is small: n/1024)
template <class FP>
__global__ void k_scan_chunks(int* __restrict__ cnt, int chunks, uint32_t d, uint32_t base, const Affine<FP>* __restrict__ carries,
                              int* __restrict__ tree_n /* d */) {
    uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= d) return;
    int off = 0;
    if (pos > 0 && !ldg_aff(carries + pos - 1).is_identity()) off = (int)base;
    int* c = cnt + (size_t)pos * chunks;
    for (int k = 0; k < chunks; ++k) { int v = c[k]; c[k] = off; off += v; }
    tree_n[pos] = off + 1;  // + the closing -carry_i
}

template <class FP>
__global__ void k_scatter_points(const uint8_t* __restrict__ planes, const Affine<FP>* __restrict__ table, size_t n, uint32_t base,
                                 const int* __restrict__ offs, int chunks, const Affine<FP>* __restrict__ carries,
                                 const int* __restrict__ tree_n, const int* __restrict__ tree_of_pos /* d: slot or -1 */,
                                 Affine<FP>* __restrict__ T, size_t cap) {
    const uint32_t pos = blockIdx.y;
    const int slot = tree_of_pos[pos];
    if (slot < 0) return;
    Affine<FP>* out = T + (size_t)slot * cap;
    __shared__ int warp_tot[32];
    const size_t j0 = (size_t)blockIdx.x * CHUNK_PTS;
    int run = offs[(size_t)pos * chunks + blockIdx.x];
    // CHUNK_PTS / blockDim rounds, each an intra-block exclusive scan of the keep flags
    for (int k0 = 0; k0 < CHUNK_PTS; k0 += blockDim.x) {
        size_t j = j0 + k0 + threadIdx.x;
        uint32_t dv = j < n ? planes[(size_t)pos * n + j] : 0;
        unsigned keep = dv != 0;
        unsigned bal = __ballot_sync(0xffffffffu, keep);
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int before = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int wbase = 0, tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { int v = warp_tot[w]; if (w < wid) wbase += v; tot += v; }
        if (keep) stg_aff(out + run + wbase + before, ldg_aff(table + j * (base - 1) + (dv - 1)));
        run += tot;
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x < base + 1) {
        if (threadIdx.x < base) {
            if (pos > 0) {
                Affine<FP> c = ldg_aff(carries + pos - 1);
                if (!c.is_identity()) stg_aff(out + threadIdx.x, aneg(c));
            }
        } else {
            stg_aff(out + tree_n[pos] - 1, aneg(ldg_aff(carries + pos)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5  pairwise affine sums with one batched inversion per level.
//   mode 0 (leaves): in = T (points), pair (2k, 2k+1), result = -(P+Q)
//   mode 1 (merges): in = outputs of the level below, result = A + B
// A missing / identity partner needs no inversion (den = 0 is skipped by the batch inverter).
// ------------------------------------------------------------------------------------------------
template <class FP>
EAGEN_D Fe<FP> pair_den(const Affine<FP>& p, const Affine<FP>& q) {
    if (p.is_identity() || q.is_identity()) return Fe<FP>::zero();
    if (p.x == q.x) {
        if (p.y == q.y) return dbl(p.y);  // tangent (y != 0 on a prime-order curve)
        return Fe<FP>::zero();            // P + (-P)
    }
    return sub(q.x, p.x);
}
template <class FP>
EAGEN_D Affine<FP> pair_sum(const Affine<FP>& p, const Affine<FP>& q, const Fe<FP>& dinv) {
    if (p.is_identity()) return q;
    if (q.is_identity()) return p;
    Fe<FP> lam;
    if (p.x == q.x) {
        if (!(p.y == q.y)) return Affine<FP>::identity();
        Fe<FP> xx = sqr(p.x);
        lam = mul(add(dbl(xx), xx), dinv);
    } else {
        lam = mul(sub(q.y, p.y), dinv);
    }
    Affine<FP> r;
    r.x = sub(sub(sqr(lam), p.x), q.x);
    r.y = sub(mul(lam, sub(p.x, r.x)), p.y);
    return r;
}

// tree-batched addressing: element (tree, idx) of a level lives at tree*stride + idx; idx valid below cnt[tree]
template <class FP>
__global__ void k_pair_den(const Affine<FP>* __restrict__ in, size_t in_stride, const int* __restrict__ in_cnt,
                           size_t out_stride, int ntrees, Fe<FP>* __restrict__ den) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= out_stride * ntrees) return;
    int tree = (int)((uint32_t)g / (uint32_t)out_stride);   // indices < 2^32
    size_t k = (uint32_t)g - (uint32_t)tree * (uint32_t)out_stride;
    int c = in_cnt[tree];
    Fe<FP> dv = Fe<FP>::zero();
    if ((long long)(2 * k + 1) < c) dv = pair_den(ldg_aff(in + tree * in_stride + 2 * k), ldg_aff(in + tree * in_stride + 2 * k + 1));
    stg(den + g, dv);
}

template <class FP>
__global__ void k_pair_finish(const Affine<FP>* __restrict__ in, size_t in_stride, const int* __restrict__ in_cnt,
                              size_t out_stride, int ntrees, const Fe<FP>* __restrict__ dinv, int negate,
                              Affine<FP>* __restrict__ out) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= out_stride * ntrees) return;
    int tree = (int)((uint32_t)g / (uint32_t)out_stride);   // indices < 2^32
    size_t k = (uint32_t)g - (uint32_t)tree * (uint32_t)out_stride;
    int c = in_cnt[tree];
    if ((long long)(2 * k) >= c) return;
    Affine<FP> p = ldg_aff(in + tree * in_stride + 2 * k);
    Affine<FP> q = (long long)(2 * k + 1) < c ? ldg_aff(in + tree * in_stride + 2 * k + 1) : Affine<FP>::identity();
    Affine<FP> s = pair_sum(p, q, ldg(dinv + g));
    stg_aff(out + g, negate ? aneg(s) : s);
}

// line through affine a, b as (lx, ly, lz) = cross((ax,ay,1),(bx,by,1)); tangent fallback through c = -(a+b)
// (reference: src/regular_functions_utils.rs:285-303 with z = 1).  Precondition: neither a nor b is the identity -- every caller
// substitutes from_point's partner -a (or skips the merge) before it gets here.
template <class FP>
EAGEN_D void line_coeffs(const Affine<FP>& a, const Affine<FP>& b, const Affine<FP>& c, Fe<FP>& lx, Fe<FP>& ly, Fe<FP>& lz) {
    lz = sub(mul(a.x, b.y), mul(a.y, b.x));   // ax*by - ay*bx
    lx = sub(a.y, b.y);                       // ay*bz - az*by
    ly = sub(b.x, a.x);                       // az*bx - ax*bz
    if (!lx.is_zero() || !ly.is_zero() || !lz.is_zero()) return;
    // a == b: the cross product vanishes and the reference takes the line through a and c = -(a + b), the tangent (:296-302)
    const bool ci = c.is_identity();
    lz = sub(mul(a.x, c.y), mul(a.y, c.x));
    lx = ci ? a.y : sub(a.y, c.y);
    ly = ci ? Fe<FP>::zero() : sub(c.x, a.x);
}

// level-0 functions: a = [lz, lx], b = [ly]; both points identity -> the constant 1
// (reference: Propagation::from_pair / from_point / empty, src/regular_functions_utils.rs:319-331)
template <class FP>
__global__ void k_leaf_lines(const Affine<FP>* __restrict__ T, size_t t_stride, const int* __restrict__ t_cnt,
                             const Affine<FP>* __restrict__ outp, size_t out_stride, int ntrees,
                             Fe<FP>* __restrict__ A /* stride 2 */, Fe<FP>* __restrict__ B /* stride 1 */,
                             Fe<FP>* __restrict__ EA /* optional, stride 2 */, Fe<FP>* __restrict__ EB /* optional, stride 2 */,
                             int* __restrict__ iso_deg /* optional: per tree, += 5 for every leaf that is a line */) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= out_stride * ntrees) return;
    int tree = (int)((uint32_t)g / (uint32_t)out_stride);   // indices < 2^32
    size_t k = (uint32_t)g - (uint32_t)tree * (uint32_t)out_stride;
    int c = t_cnt[tree];
    if ((long long)(2 * k) >= c) return;
    Affine<FP> p = ldg_aff(T + tree * t_stride + 2 * k);
    Affine<FP> q = (long long)(2 * k + 1) < c ? ldg_aff(T + tree * t_stride + 2 * k + 1) : Affine<FP>::identity();
    Fe<FP> lx, ly, lz;
    if (p.is_identity() && q.is_identity()) {
        lz = Fe<FP>::one(); lx = Fe<FP>::zero(); ly = Fe<FP>::zero();
    } else {
        if (p.is_identity()) { p = q; q = Affine<FP>::identity(); }  // from_pair(O, Q) = from_point(Q)
        if (q.is_identity()) q = aneg(p);                             // from_point(P) = line(P, -P)
        line_coeffs(p, q, ldg_aff(outp + g), lx, ly, lz);
        if (iso_deg) atomicAdd(iso_deg + tree, 5);
    }
    stg(A + 2 * g, lz);
    stg(A + 2 * g + 1, lx);
    stg(B + g, ly);
    if (EA) {   // the leaf's values on the 2-point domain {1, -1} (what a 2-point forward transform of [lz, lx] and [ly] gives)
        stg(EA + 2 * g, add(lz, lx)); stg(EA + 2 * g + 1, sub(lz, lx));
        stg(EB + 2 * g, ly); stg(EB + 2 * g + 1, ly);
    }
}

// per-merge descriptor for the level that joins children (2j, 2j+1)
enum : uint32_t { MERGE_ABSENT = 0, MERGE_PASS = 1, MERGE_SHORTCUT = 2, MERGE_GENERIC = 3 };

template <class FP>
struct MergeDesc {
    Fe<FP> alpha, beta;       // x of the children's outputs (division roots)
    Fe<FP> lz, lx, ly;        // line(-A, -B)   (l0 + l1 x + l2 y = lz + lx x + ly y)
    uint32_t mode, pad[3];
};

template <class FP>
__global__ void k_merge_desc(const Affine<FP>* __restrict__ child, size_t child_stride, const int* __restrict__ child_cnt,
                             const Affine<FP>* __restrict__ parent, size_t parent_stride, int ntrees,
                             MergeDesc<FP>* __restrict__ desc, int* __restrict__ iso_deg /* optional: per tree, += 1 per generic merge */) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= parent_stride * ntrees) return;
    int tree = (int)((uint32_t)g / (uint32_t)parent_stride);
    size_t j = (uint32_t)g - (uint32_t)tree * (uint32_t)parent_stride;
    int c = child_cnt[tree];
    MergeDesc<FP> dsc;
    dsc.pad[0] = dsc.pad[1] = dsc.pad[2] = 0;
    dsc.alpha = dsc.beta = dsc.lz = dsc.lx = dsc.ly = Fe<FP>::zero();
    if ((long long)(2 * j) >= c) dsc.mode = MERGE_ABSENT;
    else if ((long long)(2 * j + 1) >= c) dsc.mode = MERGE_PASS;
    else {
        Affine<FP> a = ldg_aff(child + tree * child_stride + 2 * j), b = ldg_aff(child + tree * child_stride + 2 * j + 1);
        if (a.is_identity() || b.is_identity()) dsc.mode = MERGE_SHORTCUT;
        else {
            dsc.mode = MERGE_GENERIC;
            if (iso_deg) atomicAdd(iso_deg + tree, 1);
            dsc.alpha = a.x; dsc.beta = b.x;
            line_coeffs(aneg(a), aneg(b), ldg_aff(parent + g), dsc.lx, dsc.ly, dsc.lz);
        }
    }
    desc[g] = dsc;
}

// ------------------------------------------------------------------------------------------------
// K6  batched radix-2 NTT, shared-memory passes of up to 10 stages over 1024-element tiles.
//   forward : decimation in frequency, natural order in -> bit-reversed order out
//   inverse : decimation in time, bit-reversed in -> natural order out, unscaled (1/T is folded into K7)
// The whole batch is one flat array of n_transforms * T elements; a pass covers stages [s_lo, s_hi].
// The first forward pass can gather from compact coefficient slots (zero padding), the last inverse
// pass can scatter back into compact slots, so no separate pad / copy kernels touch HBM.
// ------------------------------------------------------------------------------------------------
constexpr int NTT_TILE_LOG = 10;
constexpr int NTT_TILE = 1 << NTT_TILE_LOG;
constexpr int NTT_THREADS = 256;

// one polynomial family of a pass (blockIdx.y selects it: the a and the b polynomials of a level are transformed by ONE launch)
template <class FP>
struct NttSide {
    Fe<FP>* data;            // workspace, n_transforms * T
    const Fe<FP>* src;       // optional compact source (first pass)
    Fe<FP>* dst;             // optional compact destination (last pass)
    const Fe<FP>* sub_top;   // optional: sub_top[transform] is subtracted from every scattered output
    size_t src_stride, dst_stride;
    int src_len, dst_len;
};
template <class FP>
struct NttPass {
    NttSide<FP> side[2];
    const Fe<FP>* tw;        // tw[e] = w_T^e (or w_T^-e), e < T/2
    const Fe<FP>* twist;     // optional: element i of the gathered source is multiplied by twist[i] (coset transform)
    const int* counts;       // transforms present per tree
    size_t total;            // n_transforms * T
    size_t dst_off;
    int node_max;            // transforms per tree (stride of the tree index)
    int t, s_hi, s_lo;
    int tw_t;                // log2 of the transform size the twiddle table `tw` belongs to (the contiguous pass uses the compact 2^k table)
    int final_pass;          // last pass of the transform: stored values are normalised to [0, p)
};

EAGEN_D uint32_t insert_zero_bit(uint32_t v, int pos) {
    uint32_t lo = v & ((1u << pos) - 1);
    return ((v >> pos) << (pos + 1)) | lo;
}

// Shared-memory tile access with an XOR swizzle of the 16-byte chunk index.  A 32-byte element is two chunks; without the
// swizzle every power-of-two element stride maps a quarter-warp's LDS.128/STS.128 onto 2-8 times fewer bank groups than
// lanes (measured: 58 % of the shared wavefronts were conflict replays, and the butterfly phase was bound by them, not by the
// multiplier).  Folding three higher 3-bit groups into the low three chunk bits makes all strides 1..512 conflict free.
EAGEN_D uint32_t ntt_swz(uint32_t e) {
    uint32_t c = 2u * e;
    return c ^ (((c >> 3) ^ (c >> 6) ^ (c >> 9)) & 7u);
}
template <class FP>
EAGEN_D Fe<FP> ntt_lds(const uint4* sm, uint32_t e) {
    uint32_t c = ntt_swz(e);
    uint4 a = sm[c], b = sm[c ^ 1u];
    Fe<FP> r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class FP>
EAGEN_D void ntt_sts(uint4* sm, uint32_t e, const Fe<FP>& r) {
    uint32_t c = ntt_swz(e);
    sm[c] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    sm[c ^ 1u] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

#ifndef EAGEN_NTT_MINBLOCKS
#define EAGEN_NTT_MINBLOCKS 4
#endif
// EAGEN_NTT_DIAG_NOMUL (diagnostic builds only, results are wrong): butterflies add the twiddle instead of multiplying by it,
// to measure what the pass costs without its products
// EAGEN_NTT_DIAG_NOTW (diagnostic): every twiddle index is folded into the first 8 table entries (no gather traffic)
#ifdef EAGEN_NTT_DIAG_NOTW
#define NTT_TWI(x) ((size_t)7 & (size_t)(x))
#else
#define NTT_TWI(x) (x)
#endif
#ifdef EAGEN_NTT_DIAG_NOMUL
#define NTT_MUL(a, b) add(a, b)
#elif defined(EAGEN_NTT_CANONICAL)   // round-1 butterflies on canonical values, kept for A/B timing (tools/variant.sh)
#define NTT_MUL(a, b) mul(a, b)
#define NTT_ADD(a, b) add(a, b)
#define NTT_SUB(a, b) sub(a, b)
#define NTT_NORM(a) (a)
#else
// Butterflies on lazily reduced values (field.cuh): everything in shared memory and between the passes of one transform lies in
// [0, 2p); the product by a canonical twiddle needs no final subtraction.  The LAST pass normalises what it stores, so a
// transform's outputs are canonical and bit-identical to the fully reduced formulation.
#define NTT_MUL(a, b) mul_lazy(a, b)
#define NTT_ADD(a, b) add_lazy(a, b)
#define NTT_SUB(a, b) sub_lazy(a, b)
#define NTT_NORM(a) normalise_lazy(a)
#endif
#ifdef EAGEN_NTT_DIAG_NOMUL
#define NTT_ADD(a, b) add(a, b)
#define NTT_SUB(a, b) sub(a, b)
#define NTT_NORM(a) (a)
#endif
template <class FP, bool INVERSE>
__global__ void __launch_bounds__(NTT_THREADS, EAGEN_NTT_MINBLOCKS)
k_ntt_pass(NttPass<FP> a) {
    __shared__ uint4 sm[NTT_TILE * 2];
    const int k = a.s_hi - a.s_lo + 1;
    const int lw = NTT_TILE_LOG - k;
    const size_t tile = blockIdx.x;
    const size_t Tmask = ((size_t)1 << a.t) - 1;
    const NttSide<FP>& sd = a.side[blockIdx.y];
    {   // Tiles whose transforms all belong to absent nodes (trees shorter than the longest one of the batch: ~15 % of the tiles of a
        // 2^20-point witness) do no work at all.  A strided pass touches one transform per tile, the contiguous pass 2^(10-k)
        // consecutive ones; a tile that straddles two trees is treated as present.
        const size_t first = a.s_lo == 0 ? tile << NTT_TILE_LOG : (((tile << lw) >> a.s_lo) << (a.s_hi + 1));
        const size_t last = a.s_lo == 0 ? first + NTT_TILE - 1 : first;
        const uint32_t tr0 = (uint32_t)(first >> a.t), tr1 = (uint32_t)(last >> a.t);
        const uint32_t tree0 = tr0 / (uint32_t)a.node_max, tree1 = tr1 / (uint32_t)a.node_max;
        if (tree0 == tree1 && (int)(tr0 - tree0 * (uint32_t)a.node_max) >= a.counts[tree0]) return;
    }

    size_t lin[4];
    bool live[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t q = r * NTT_THREADS + threadIdx.x;
        uint32_t E, wl;
        if (a.s_lo == 0) { E = q & ((1u << k) - 1); wl = q >> k; }
        else { wl = q & ((1u << lw) - 1); E = q >> lw; }
        size_t w = (tile << lw) + wl;
        size_t L = w & (((size_t)1 << a.s_lo) - 1), H = w >> a.s_lo;
        lin[r] = (H << (a.s_hi + 1)) | ((size_t)E << a.s_lo) | L;
        live[r] = lin[r] < a.total;
        Fe<FP> v = Fe<FP>::zero();
        if (live[r]) {
            size_t tr = lin[r] >> a.t;
            const uint32_t tr32 = (uint32_t)tr;   // transform index < 2^32: 32-bit division instead of a 64-bit div/mod pair
            int tree = (int)(tr32 / (uint32_t)a.node_max), node = (int)(tr32 - (uint32_t)tree * (uint32_t)a.node_max);
            live[r] = node < a.counts[tree];
            if (live[r]) {
                size_t i = lin[r] & Tmask;
                if (sd.src) {
                    if ((int)i < sd.src_len) {
                        v = ldg(sd.src + tr * sd.src_stride + i);
                        if (a.twist) v = NTT_MUL(v, ldg(a.twist + i));
                    }
                }
                else v = ldg(sd.data + lin[r]);
            }
        }
        ntt_sts(sm, (E << lw) | wl, v);
    }
    __syncthreads();

    // Radix-4 rounds: two stages per shared-memory round trip, four elements per thread in registers (one barrier,
    // one set of index computations and three twiddle loads for four butterflies; the two products of each half-round are
    // independent, which doubles the instruction-level parallelism of the carry chains).
    int st = 0;
    for (; st + 1 < k; st += 2) {
        const int slo = INVERSE ? st : (k - 2 - st);     // lower local stage of the pair (the other one is slo + 1)
        const int bitpos = slo + lw;
        const int sgl = a.s_lo + slo, sgh = sgl + 1;     // global stages
        const uint32_t bq = threadIdx.x;
        const uint32_t i0 = ((bq >> bitpos) << (bitpos + 2)) | (bq & ((1u << bitpos) - 1));
        const uint32_t d = 1u << bitpos;
        const uint32_t E = i0 >> lw, wl = i0 & ((1u << lw) - 1);
        const size_t w = (tile << lw) + wl;
        const size_t L = w & (((size_t)1 << a.s_lo) - 1);
        const size_t jl = ((size_t)(E & ((1u << slo) - 1)) << a.s_lo) | L;   // index inside the 2^sgl group (same for all four)
        Fe<FP> x0 = ntt_lds<FP>(sm, i0), x1 = ntt_lds<FP>(sm, i0 + d), x2 = ntt_lds<FP>(sm, i0 + 2 * d), x3 = ntt_lds<FP>(sm, i0 + 3 * d);
        if (sgl == 0) {
            // Global stages 1 and 0: the twiddles are 1, 1 and w_4 for every thread (jl = 0), so three of the four products
            // vanish (uniform branch).  Over a whole tree 2/t of all butterflies of a 2^t-point transform have w = 1; this
            // round alone carries three quarters of them.  Multiplying by the Montgomery 1 would give the same bits.
            const Fe<FP> W1b = ldg(a.tw + NTT_TWI((size_t)1 << (a.tw_t - 2)));   // w_4 (or its inverse)
            if (INVERSE) {
                Fe<FP> y0 = NTT_ADD(x0, x1), y1 = NTT_SUB(x0, x1), y2 = NTT_ADD(x2, x3), y3 = NTT_SUB(x2, x3);
                Fe<FP> u3 = NTT_MUL(y3, W1b);
                ntt_sts(sm, i0, NTT_ADD(y0, y2)); ntt_sts(sm, i0 + 2 * d, NTT_SUB(y0, y2));
                ntt_sts(sm, i0 + d, NTT_ADD(y1, u3)); ntt_sts(sm, i0 + 3 * d, NTT_SUB(y1, u3));
            } else {
                Fe<FP> y0 = NTT_ADD(x0, x2), y2 = NTT_SUB(x0, x2);
                Fe<FP> y1 = NTT_ADD(x1, x3), y3 = NTT_MUL(NTT_SUB(x1, x3), W1b);
                ntt_sts(sm, i0, NTT_ADD(y0, y1)); ntt_sts(sm, i0 + d, NTT_SUB(y0, y1));
                ntt_sts(sm, i0 + 2 * d, NTT_ADD(y2, y3)); ntt_sts(sm, i0 + 3 * d, NTT_SUB(y2, y3));
            }
        } else {
            const Fe<FP> W0 = ldg(a.tw + NTT_TWI(jl << (a.tw_t - 1 - sgl)));
            const Fe<FP> W1a = ldg(a.tw + NTT_TWI(jl << (a.tw_t - 1 - sgh)));
            const Fe<FP> W1b = ldg(a.tw + NTT_TWI((jl + ((size_t)1 << sgl)) << (a.tw_t - 1 - sgh)));
            if (INVERSE) {   // decimation in time: stage slo (distance d), then stage slo+1 (distance 2d)
                Fe<FP> v1 = NTT_MUL(x1, W0), v3 = NTT_MUL(x3, W0);
                Fe<FP> y0 = NTT_ADD(x0, v1), y1 = NTT_SUB(x0, v1), y2 = NTT_ADD(x2, v3), y3 = NTT_SUB(x2, v3);
                Fe<FP> u2 = NTT_MUL(y2, W1a), u3 = NTT_MUL(y3, W1b);
                ntt_sts(sm, i0, NTT_ADD(y0, u2)); ntt_sts(sm, i0 + 2 * d, NTT_SUB(y0, u2));
                ntt_sts(sm, i0 + d, NTT_ADD(y1, u3)); ntt_sts(sm, i0 + 3 * d, NTT_SUB(y1, u3));
            } else {         // decimation in frequency: stage slo+1 (distance 2d), then stage slo (distance d)
                Fe<FP> y0 = NTT_ADD(x0, x2), y2 = NTT_MUL(NTT_SUB(x0, x2), W1a);
                Fe<FP> y1 = NTT_ADD(x1, x3), y3 = NTT_MUL(NTT_SUB(x1, x3), W1b);
                ntt_sts(sm, i0, NTT_ADD(y0, y1)); ntt_sts(sm, i0 + d, NTT_MUL(NTT_SUB(y0, y1), W0));
                ntt_sts(sm, i0 + 2 * d, NTT_ADD(y2, y3)); ntt_sts(sm, i0 + 3 * d, NTT_MUL(NTT_SUB(y2, y3), W0));
            }
        }
        __syncthreads();
    }
    for (; st < k; ++st) {   // odd stage count: one radix-2 stage is left
        const int sigma = INVERSE ? st : (k - 1 - st);   // local stage
        const int sg = a.s_lo + sigma;                   // global stage
        const int bitpos = sigma + lw;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            uint32_t bq = r * NTT_THREADS + threadIdx.x;
            uint32_t i0 = insert_zero_bit(bq, bitpos), i1 = i0 | (1u << bitpos);
            uint32_t E = i0 >> lw, wl = i0 & ((1u << lw) - 1);
            size_t w = (tile << lw) + wl;
            size_t L = w & (((size_t)1 << a.s_lo) - 1);
            size_t j = ((size_t)(E & ((1u << sigma) - 1)) << a.s_lo) | L;
            Fe<FP> u = ntt_lds<FP>(sm, i0), v = ntt_lds<FP>(sm, i1);
            if (sg == 0) {   // w = 1 for every butterfly of global stage 0
                ntt_sts(sm, i0, NTT_ADD(u, v));
                ntt_sts(sm, i1, NTT_SUB(u, v));
                continue;
            }
            Fe<FP> wj = ldg(a.tw + NTT_TWI(j << (a.tw_t - 1 - sg)));
            if (INVERSE) {
                v = NTT_MUL(v, wj);
                ntt_sts(sm, i0, NTT_ADD(u, v));
                ntt_sts(sm, i1, NTT_SUB(u, v));
            } else {
                ntt_sts(sm, i0, NTT_ADD(u, v));
                ntt_sts(sm, i1, NTT_MUL(NTT_SUB(u, v), wj));
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (!live[r]) continue;
        uint32_t q = r * NTT_THREADS + threadIdx.x;
        uint32_t E, wl;
        if (a.s_lo == 0) { E = q & ((1u << k) - 1); wl = q >> k; }
        else { wl = q & ((1u << lw) - 1); E = q >> lw; }
        Fe<FP> v = ntt_lds<FP>(sm, (E << lw) | wl);
        if (a.final_pass) v = NTT_NORM(v);   // the transform's outputs leave canonical
        if (sd.dst) {
            size_t tr = lin[r] >> a.t, i = lin[r] & Tmask;
            if ((int)i < sd.dst_len) {
                if (sd.sub_top) v = sub(v, ldg(sd.sub_top + tr));
                stg(sd.dst + tr * sd.dst_stride + a.dst_off + i, v);
            }
        } else {
            stg(sd.data + lin[r], v);
        }
    }
}

// w_T^e tables for every size up to 2^tmax:  tab[2^(t-1) + e] = w_{2^t}^e, e < 2^(t-1)
template <class FP>
__global__ void k_gen_twiddles(Fe<FP>* __restrict__ tab, int tmax, int inverse) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0 || idx >= ((size_t)1 << tmax)) return;
    int t = 64 - __clzll((unsigned long long)idx);   // idx in [2^(t-1), 2^t)
    size_t e = idx - ((size_t)1 << (t - 1));
    Fe<FP> w = omega_for<FP>((unsigned)t, inverse != 0);
    Fe<FP> r = Fe<FP>::one();
    for (int bit = t - 2; bit >= 0; --bit) {
        r = sqr(r);
        if ((e >> bit) & 1) r = mul(r, w);
    }
    stg(tab + idx, r);
}

// evaluation point stored at position p of a forward transform of size 2^t
template <class FP>
EAGEN_D Fe<FP> eval_point(const Fe<FP>* __restrict__ tw, int t, uint32_t p) {
    uint32_t kx = t ? (__brev(p) >> (32 - t)) : 0;
    uint32_t half = 1u << (t - 1);
    if (kx < half) return ldg(tw + kx);
    return neg(ldg(tw + (kx - half)));
}

// Evaluation-point tables for every transform size up to 2^tmax: for size T = 2^t, xt[T + p] is the point stored at position p
// of the bit-reversed forward transform and gt[T + p] = x^3 + b there (y^2 of the curve), so the merge kernels load them
// instead of decoding the position and spending two products per point.
template <class CC>
__global__ void k_gen_points(const Fe<typename CC::Base>* __restrict__ tw_all, int tmax, Fe<typename CC::Base>* __restrict__ xt,
                             Fe<typename CC::Base>* __restrict__ gt) {
    typedef typename CC::Base F;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2 || idx >= ((size_t)2 << tmax)) return;
    int t = 63 - __clzll((unsigned long long)idx);
    uint32_t p = (uint32_t)(idx - ((size_t)1 << t));
    Fe<F> x = eval_point(tw_all + ((size_t)1 << (t - 1)), t, p);
    stg(xt + idx, x);
    stg(gt + idx, add(mul(sqr(x), x), CC::b()));
}

// ------------------------------------------------------------------------------------------------
// K7  merge in the evaluation domain.  Children (a1 + y b1), (a2 + y b2), line l = lz + lx x + ly y,
// g(x) = x^3 + b:
//   (a2 + y b2) * l      = U + y V,   U = a2*lam + b2*ly*g,  V = a2*ly + b2*lam,  lam = lz + lx x
//   (a1 + y b1)(U + y V) = (a1 U + b1 V g) + y (a1 V + b1 U)
// then / ((x - alpha)(x - beta)) pointwise (reference: src/regular_functions_utils.rs:266-273,344-357).
// ------------------------------------------------------------------------------------------------
template <class FP>
__global__ void k_den(const MergeDesc<FP>* __restrict__ desc, size_t nmerges, int t, const Fe<FP>* __restrict__ xt /* x at position p */,
                      Fe<FP>* __restrict__ den, int* err) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (nmerges << t)) return;
    size_t m = g >> t;
    uint32_t p = (uint32_t)(g & (((size_t)1 << t) - 1));
    Fe<FP> dv = Fe<FP>::zero();
    if (desc[m].mode == MERGE_GENERIC) {
        Fe<FP> x = ldg(xt + p);
        dv = mul(sub(x, ldg(&desc[m].alpha)), sub(x, ldg(&desc[m].beta)));
        if (dv.is_zero()) atomicOr(err, KERR_COLLISION);
    }
    stg(den + g, dv);
}

// Writes the parents' TRUE evaluations on the T-point domain, in the layout of the next level's evaluation buffers
// (parent m at m*out_stride, out_stride = 2T): positions [0, T) of a 2T-point bit-reversed transform are exactly the
// T-point domain in bit-reversed order, so the next level only has to add the odd coset (see Engine::run_trees).
#ifndef EAGEN_PW_MINBLOCKS
#define EAGEN_PW_MINBLOCKS 6
#endif
// ISO: the tree is being built on the isomorphic curve y^2 = x^3 + u^6 b (collision fallback, see Engine::run_trees_safe), whose
// right-hand side is the tabulated x^3 + b plus the constant gshift = (u^6 - 1) b.
// one evaluation point g = (merge m, position p) of the merge; `di` = 1 / ((x - alpha)(x - beta)) there (read only by generic merges)
template <class CC, bool ISO>
EAGEN_D void pointwise_at(const MergeDesc<typename CC::Base>* __restrict__ desc, size_t g, int t, uint32_t mode,
                          const Fe<typename CC::Base>* __restrict__ xt, const Fe<typename CC::Base>* __restrict__ gt,
                          const Fe<typename CC::Base>* __restrict__ EA, const Fe<typename CC::Base>* __restrict__ EB,
                          const Fe<typename CC::Base>& di, size_t merges_per_tree, size_t nodes_per_tree,
                          Fe<typename CC::Base>* __restrict__ OA, Fe<typename CC::Base>* __restrict__ OB, size_t out_stride,
                          const Fe<typename CC::Base>& gshift) {
    typedef typename CC::Base F;
    size_t m = g >> t;
    uint32_t p = (uint32_t)(g & (((size_t)1 << t) - 1));
    const uint32_t tree32 = (uint32_t)m / (uint32_t)merges_per_tree;   // merge index < 2^32
    size_t tree = tree32, j = (uint32_t)m - tree32 * (uint32_t)merges_per_tree;
    size_t c1 = ((tree * nodes_per_tree + 2 * j) << t) + p, c2 = c1 + ((size_t)1 << t);
    Fe<F> a1 = ldg(EA + c1), b1 = ldg(EB + c1);
    Fe<F> ra, rb;
    if (mode == MERGE_PASS) {
        ra = a1; rb = b1;
    } else {
        Fe<F> a2 = ldg(EA + c2), b2 = ldg(EB + c2);
        Fe<F> gx = ldg(gt + p);
        if (ISO) gx = add(gx, gshift);
        // products of the form (u + y v)(u' + y v') = (u u' + v v' g) + y (u v' + v u') with three multiplications for the
        // cross term (Karatsuba): 4 instead of 5 field products each
        if (mode == MERGE_SHORTCUT) {
            Fe<F> q1 = mul(a1, a2), q2 = mul(b1, b2), q3 = mul(add(a1, b1), add(a2, b2));
            ra = add(q1, mul(q2, gx));
            rb = sub(sub(q3, q1), q2);
        } else {
            Fe<F> x = ldg(xt + p);
            Fe<F> lam = add(ldg(&desc[m].lz), mul(ldg(&desc[m].lx), x));
            Fe<F> ly = ldg(&desc[m].ly);
            Fe<F> p1 = mul(a2, lam), p2 = mul(b2, ly), p3 = mul(add(a2, b2), add(lam, ly));
            Fe<F> U = add(p1, mul(p2, gx));
            Fe<F> V = sub(sub(p3, p1), p2);
            Fe<F> q1 = mul(a1, U), q2 = mul(b1, V), q3 = mul(add(a1, b1), add(U, V));
            ra = mul(add(q1, mul(q2, gx)), di);
            rb = mul(sub(sub(q3, q1), q2), di);
        }
    }
    stg(OA + m * out_stride + p, ra);
    stg(OB + m * out_stride + p, rb);
}

template <class CC, bool ISO>
__global__ void __launch_bounds__(128, EAGEN_PW_MINBLOCKS)
k_pointwise(const MergeDesc<typename CC::Base>* __restrict__ desc, size_t nmerges, int t,
                            const Fe<typename CC::Base>* __restrict__ xt /* x at position p */,
                            const Fe<typename CC::Base>* __restrict__ gt /* x^3 + b at position p */,
                            const Fe<typename CC::Base>* __restrict__ EA, const Fe<typename CC::Base>* __restrict__ EB,
                            const Fe<typename CC::Base>* __restrict__ dinv, size_t merges_per_tree, size_t nodes_per_tree,
                            Fe<typename CC::Base>* __restrict__ OA, Fe<typename CC::Base>* __restrict__ OB, size_t out_stride,
                            Fe<typename CC::Base> gshift) {
    typedef typename CC::Base F;
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (nmerges << t)) return;
    uint32_t mode = desc[g >> t].mode;
    if (mode == MERGE_ABSENT) return;
    Fe<F> di = Fe<F>::zero();
    if (mode == MERGE_GENERIC) di = ldg(dinv + g);
    pointwise_at<CC, ISO>(desc, g, t, mode, xt, gt, EA, EB, di, merges_per_tree, nodes_per_tree, OA, OB, out_stride, gshift);
}

// The parent's a has T+1 coefficients but the transform has T points: the top coefficient q_T aliases
// onto coefficient 0.  q_T has a closed form in the children's top coefficients, so compute it, subtract it
// from coefficient 0 and store it at index T.
// Scaling: the inverse transforms are unscaled, so the stored coefficients of a level are s * (true coefficients)
// with s = (size of the transform that produced them) (1 for the leaves).  kc = 1/s_children^2 brings the closed form
// back to the true q_T (written to `top` for the next level's coset transform); the stored copy is T * q_T.
template <class FP>
__global__ void k_fixup(const MergeDesc<FP>* __restrict__ desc, size_t nmerges, int t, size_t merges_per_tree, size_t nodes_per_tree,
                        const Fe<FP>* __restrict__ A, const Fe<FP>* __restrict__ B /* children, slots m+1 / m */,
                        Fe<FP>* __restrict__ PA /* parent a, slot T+1 */, Fe<FP> kc, Fe<FP> kT, Fe<FP>* __restrict__ top) {
    size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmerges) return;
    uint32_t mode = desc[m].mode;
    if (mode == MERGE_ABSENT) return;
    const size_t T = (size_t)1 << t, h = T >> 1;  // children: a has h+1 coefficients, b has h
    Fe<FP>* pa = PA + m * (T + 1);
    Fe<FP> q = Fe<FP>::zero();
    if (mode != MERGE_PASS) {
        const uint32_t tree32 = (uint32_t)m / (uint32_t)merges_per_tree;
        size_t tree = tree32, j = (uint32_t)m - tree32 * (uint32_t)merges_per_tree;
        size_t n1 = tree * nodes_per_tree + 2 * j, n2 = n1 + 1;
        const Fe<FP>* a1 = A + n1 * (h + 1); const Fe<FP>* a2 = A + n2 * (h + 1);
        const Fe<FP>* b1 = B + n1 * h; const Fe<FP>* b2 = B + n2 * h;
        Fe<FP> a1t = ldg(a1 + h), a2t = ldg(a2 + h), b1t = ldg(b1 + h - 1), b2t = ldg(b2 + h - 1);
        if (mode == MERGE_GENERIC) {
            // x^(T+2) coefficient of the numerator's a-part = b1t b2t lx + (a1t b2t + b1t a2t) ly
            q = add(mul(mul(b1t, b2t), ldg(&desc[m].lx)), mul(add(mul(a1t, b2t), mul(b1t, a2t)), ldg(&desc[m].ly)));
        } else {
            // x^T coefficient of a1 a2 + b1 b2 (x^3 + b)
            q = mul(a1t, a2t);
            if (h >= 2) q = add(q, add(mul(b1t, ldg(b2 + h - 2)), mul(ldg(b1 + h - 2), b2t)));
        }
        q = mul(q, kc);
    } else {
        // pass-through: the parent IS the child, whose true top coefficient (index h of T+1, zero above) needs no alias fix
        q = Fe<FP>::zero();
    }
    Fe<FP> qs = mul(q, kT);
    stg(pa, sub(ldg(pa), qs));
    stg(pa + T, qs);
    stg(top + m, q);
}

// Odd-coset pre-multipliers of every level in one table laid out like the twiddle table: for the merge that transforms children
// of m = 2^l coefficients, twist[m + i] = w_{2m}^i * 2^-l, i < m (w_{2m}^i is tw_all[m + i]); the factor 2^-l undoes the scaling the
// unscaled inverse transforms leave in the stored coefficients.
template <class FP>
__global__ void k_gen_twist_all(const Fe<FP>* __restrict__ tw_all, int tmax, Fe<FP>* __restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0 || idx >= ((size_t)1 << tmax)) return;
    int l = 63 - __clzll((unsigned long long)idx);
    Fe<FP> sc = Fe<FP>::one(), h = Fe<FP>::two_inv();
    for (int i = 0; i < l; ++i) sc = mul(sc, h);
    stg(out + idx, mul(ldg(tw_all + idx), sc));
}

// out[i] *= c (root rescale in raw mode)
template <class FP>
__global__ void k_scale_const(Fe<FP>* __restrict__ coef, size_t stride, int len, Fe<FP> c) {
    int tree = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    Fe<FP>* q = coef + (size_t)tree * stride + i;
    stg(q, mul(ldg(q), c));
}

// ------------------------------------------------------------------------------------------------
// K10  canonical form: trim trailing zeros, divide by the coefficient of highest pole order
// ------------------------------------------------------------------------------------------------
template <class FP>
__global__ void k_find_top(const Fe<FP>* __restrict__ coef, size_t stride, int len, int* __restrict__ top /* per tree, init 0 */) {
    int tree = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    if (!ldg(coef + (size_t)tree * stride + i).is_zero()) atomicMax(top + tree, i + 1);
}

template <class FP>
__global__ void k_lead(const Fe<FP>* __restrict__ A, size_t sa, const int* __restrict__ topa, const Fe<FP>* __restrict__ B, size_t sb,
                       const int* __restrict__ topb, int ntrees, Fe<FP>* __restrict__ lead) {
    int tree = blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= ntrees) return;
    int la = topa[tree], lb = topb[tree];
    long oa = la ? 2L * (la - 1) : -1, ob = lb ? 2L * (lb - 1) + 3 : -1;
    Fe<FP> l = Fe<FP>::zero();
    if (la || lb) l = oa > ob ? ldg(A + (size_t)tree * sa + la - 1) : ldg(B + (size_t)tree * sb + lb - 1);
    stg(lead + tree, l);
}

// monic scaling of the root functions, written straight into the result's slots (tree tr -> slot first_slot + (dir > 0 ? tr : ntrees-1-tr))
template <class FP>
__global__ void k_scale(const Fe<FP>* __restrict__ coef, size_t stride, const int* __restrict__ top, const Fe<FP>* __restrict__ linv,
                        Fe<FP>* __restrict__ dst, size_t dst_stride, size_t first_slot, int dir) {
    int tree = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= top[tree]) return;
    size_t slot = first_slot + (dir > 0 ? (size_t)tree : (size_t)(gridDim.y - 1 - tree));
    stg(dst + slot * dst_stride + i, mul(ldg(coef + (size_t)tree * stride + i), ldg(linv + tree)));
}

// root functions of the trees of a group into the result's slots: tree tr -> slot first_slot + (dir > 0 ? tr : ntrees - 1 - tr)
template <class FP>
__global__ void k_copy_strided(const Fe<FP>* __restrict__ src, size_t src_stride, Fe<FP>* __restrict__ dst, size_t dst_stride, int len,
                               size_t first_slot, int dir) {
    int tree = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    size_t slot = first_slot + (dir > 0 ? (size_t)tree : (size_t)(gridDim.y - 1 - tree));
    stg(dst + slot * dst_stride + i, ldg(src + (size_t)tree * src_stride + i));
}

// pointwise product with scaling for the stand-alone polynomial product (Polynomial::mul_fft, :119-127)
template <class FP>
__global__ void k_mul_scale(Fe<FP>* __restrict__ a, const Fe<FP>* __restrict__ b, size_t n, Fe<FP> sc) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) stg(a + i, mul(mul(ldg(a + i), ldg(b + i)), sc));
}

// evaluate a(x) + y b(x) at many affine points: one thread per point, Horner (RegularFunction::ev, :228-237)
template <class FP>
__global__ void k_eval_function(const Fe<FP>* __restrict__ A, int la, const Fe<FP>* __restrict__ B, int lb,
                                const Affine<FP>* __restrict__ pts, size_t n, Fe<FP>* __restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Affine<FP> p = ldg_aff(pts + j);
    if (p.is_identity()) { stg(out + j, Fe<FP>::zero()); return; }
    Fe<FP> va = Fe<FP>::zero(), vb = Fe<FP>::zero();
    for (int i = la - 1; i >= 0; --i) va = add(mul(va, p.x), ldg(A + i));
    for (int i = lb - 1; i >= 0; --i) vb = add(mul(vb, p.x), ldg(B + i));
    stg(out + j, add(va, mul(vb, p.y)));
}

// ------------------------------------------------------------------------------------------------
// prepare_scalar_witness on the digit planes (reference: src/negbase_utils.rs:79-124): per scalar, base rows of
// (num_limbs + 1) entries -- Entry::Scalar at (0,0), Entry::Bucket(sum of (-b)^i over the positions holding digit r) at (r,0),
// Entry::Limb(value, bitmask) elsewhere.  i128 arithmetic wraps like a release build of the reference.
//   mode 0 (faithful): a non-zero digit at position i lands in slot i % logtable + 1 with exponent i % logtable (:98-101)
//   mode 1 (intended): slot i / logtable + 1, exponent i % logtable
// One thread per (scalar, row); the digit planes are read coalesced across scalars; HBM-bound, no field arithmetic.
// ------------------------------------------------------------------------------------------------
struct PswEntry { uint64_t lo, hi; uint32_t mask, kind; uint64_t zero; };   // 32 bytes; kind: 0 Scalar, 1 Bucket, 2 Limb
enum : int { KERR_PSW_SLOT = 8 };   // faithful mode: limb slot beyond num_limbs (the reference's out-of-bounds panic)

struct W128 { uint64_t lo, hi; };
EAGEN_HD W128 w128_add(W128 a, W128 b) { W128 r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi + (r.lo < a.lo ? 1u : 0u); return r; }
EAGEN_HD W128 w128_mul_small(W128 a, uint32_t b) {
    const uint64_t M = 0xffffffffull;
    uint64_t p0 = (a.lo & M) * b, p1 = (a.lo >> 32) * b + (p0 >> 32);
    W128 r; r.lo = (p1 << 32) | (p0 & M); r.hi = a.hi * b + (p1 >> 32);
    return r;
}
EAGEN_HD W128 w128_neg(W128 a) { W128 r; r.lo = ~a.lo + 1; r.hi = ~a.hi + (r.lo == 0 ? 1u : 0u); return r; }

template <class FS>
__global__ void k_scalar_witness(const uint8_t* __restrict__ planes /* d x n, MSD first */, size_t n, uint32_t d, uint32_t base,
                                 uint32_t num_digits, uint32_t logtable, uint32_t num_limbs, int mode,
                                 const Fe<FS>* __restrict__ scalars, PswEntry* __restrict__ out, int* err) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = blockIdx.y;
    if (j >= n) return;
    const uint32_t cols = num_limbs + 1;
    PswEntry* o = out + ((size_t)j * base + r) * cols;
    auto digit = [&](uint32_t i) -> uint32_t { return planes[(size_t)(d - 1 - i) * n + j]; };
    auto match = [&](uint32_t dg) -> bool { return r == 0 ? dg != 0 : dg == r; };
    PswEntry e0;
    e0.zero = 0; e0.mask = 0;
    if (r == 0) {
        uint32_t x[8];
        Fe<FS> xm = ldg(scalars + j);
        from_mont<FS>(xm.v, x);
        e0.lo = (uint64_t)x[0] | ((uint64_t)x[1] << 32); e0.hi = (uint64_t)x[2] | ((uint64_t)x[3] << 32); e0.kind = 0;
        int bad = 0;
        for (uint32_t i = 0; i < d; ++i) {
            if (digit(i) == 0) continue;
            if (i >= num_digits) bad |= KERR_DIGITS;                               // assert!(digits.len() <= num_digits)  :81
            if (mode == 0 && (i % logtable) >= num_limbs) bad |= KERR_PSW_SLOT;    // ret[..][i % logtable + 1] out of bounds
        }
        if (bad) atomicOr(err, bad);
    } else {
        W128 acc = {0, 0}, pw = {1, 0};
        for (uint32_t i = 0; i < d; ++i) {
            if (digit(i) == r) acc = w128_add(acc, pw);
            pw = w128_neg(w128_mul_small(pw, base));
        }
        e0.lo = acc.lo; e0.hi = acc.hi; e0.kind = 1;
    }
    o[0] = e0;
    for (uint32_t e = 1; e <= num_limbs; ++e) {
        W128 acc = {0, 0};
        uint32_t mask = 0;
        if (mode == 0) {
            const uint32_t k0 = e - 1;   // the only exponent that reaches slot e
            if (k0 < logtable) {
                uint32_t cnt = 0;
                for (uint32_t i = k0; i < d; i += logtable) cnt += match(digit(i)) ? 1u : 0u;
                W128 pw = {1, 0};
                for (uint32_t k = 0; k < k0; ++k) pw = w128_neg(w128_mul_small(pw, base));
                acc = w128_mul_small(pw, cnt);
                mask = cnt << k0;
            }
        } else {
            W128 pw = {1, 0};
            for (uint32_t k = 0; k < logtable; ++k) {
                const uint32_t i = (e - 1) * logtable + k;
                if (i < d && match(digit(i))) { acc = w128_add(acc, pw); mask += 1u << k; }
                pw = w128_neg(w128_mul_small(pw, base));
            }
        }
        PswEntry en;
        en.lo = acc.lo; en.hi = acc.hi; en.mask = mask; en.kind = 2; en.zero = 0;
        o[e] = en;
    }
}

// ------------------------------------------------------------------------------------------------
// compute_divisor_witness_naive (reference: src/regular_functions_utils.rs:483-551): one round joins the pairs the host
// planned from the identity flags of the current list; pair t of the round (in the reference's push order) produces the line
// through (a, b) and the point -(a + b), both stored at the reference's pop order np - 1 - t.
// ------------------------------------------------------------------------------------------------
template <class FP>
__global__ void k_identity_flags(const Affine<FP>* __restrict__ pts, size_t m, uint8_t* __restrict__ flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) flags[i] = ldg_aff(pts + i).is_identity() ? 1 : 0;
}
template <class FP>
__global__ void k_naive_den(const Affine<FP>* __restrict__ list, const int2* __restrict__ pairs, size_t np, Fe<FP>* __restrict__ den) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np) return;
    int2 pr = pairs[t];
    stg(den + t, pair_den(ldg_aff(list + pr.x), ldg_aff(list + pr.y)));
}
template <class FP>
struct LineTriple { Fe<FP> lx, ly, lz; };
template <class FP>
__global__ void k_naive_finish(const Affine<FP>* __restrict__ list, const int2* __restrict__ pairs, size_t np, const Fe<FP>* __restrict__ dinv,
                               Affine<FP>* __restrict__ next /* np slots */, LineTriple<FP>* __restrict__ lines /* np slots */) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np) return;
    int2 pr = pairs[t];
    Affine<FP> a = ldg_aff(list + pr.x), b = ldg_aff(list + pr.y);
    Affine<FP> c = aneg(pair_sum(a, b, ldg(dinv + t)));
    LineTriple<FP> l;
    // a is never the identity (the reference skips it, :517); b may be: the cross product then vanishes and linefunc falls back
    // to the line through a and -(a + O) = -a (:296-302)
    line_coeffs(a, b.is_identity() ? aneg(a) : b, c, l.lx, l.ly, l.lz);
    const size_t o = np - 1 - t;
    stg_aff(next + o, c);
    stg(&lines[o].lx, l.lx); stg(&lines[o].ly, l.ly); stg(&lines[o].lz, l.lz);
}

// ------------------------------------------------------------------------------------------------
// Collision fallback (canonical form only).  The pointwise division of a merge is undefined when an output point's
// x-coordinate lies on the power-of-two evaluation domain -- which happens for natural inputs: x = -1 (the Pasta generator)
// and x = 1 (the Grumpkin generator) are on EVERY domain.  The divisor witness is unique up to a constant, and
// (x, y) -> (u^2 x, u^3 y) is an isomorphism onto y^2 = x^3 + u^6 b, so the tree is rebuilt there (x-coordinates move off the
// domain) and mapped back:  f(x, y) = f'(u^2 x, u^3 y), i.e. a_i = a'_i u^(2i), b_i = b'_i u^(2i+3); the canonical (monic) form of
// f is the one the direct computation would have produced.  The RAW function is recovered exactly as well: every line is
// homogeneous of degree 5 in u and every generic merge divides by u^4 (x - alpha)(x - beta), so the tree built on the isomorphic
// curve is u^k times the reference's raw function with k = 5 * (leaves that are lines) + (generic merges), counted per tree.
// ------------------------------------------------------------------------------------------------
template <class FP>
__global__ void k_iso_points(Affine<FP>* __restrict__ T, size_t cap, const int* __restrict__ cnt, int nt, Fe<FP> u2, Fe<FP> u3) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= cap * (size_t)nt) return;
    const int tree = (int)(g / cap);
    if ((long long)(g - (size_t)tree * cap) >= cnt[tree]) return;
    Affine<FP> p = ldg_aff(T + g);
    if (p.is_identity()) return;
    p.x = mul(p.x, u2); p.y = mul(p.y, u3);
    stg_aff(T + g, p);
}
// c_i *= u2^i * extra (* per_tree[tree]) for the `len` coefficients of each of the gridDim.y polynomials (stride elements apart)
template <class FP>
__global__ void k_iso_unscale(Fe<FP>* __restrict__ C, size_t stride, int len, Fe<FP> u2, Fe<FP> extra, const Fe<FP>* __restrict__ per_tree) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)len) return;
    Fe<FP> s = per_tree ? mul(extra, ldg(per_tree + blockIdx.y)) : extra;
    Fe<FP> w = u2;
    for (uint32_t e = (uint32_t)i; e; e >>= 1) {
        if (e & 1) s = mul(s, w);
        w = sqr(w);
    }
    Fe<FP>* c = C + (size_t)blockIdx.y * stride + i;
    stg(c, mul(ldg(c), s));
}

// ------------------------------------------------------------------------------------------------
// RegularFunction::ev for every function of a device-resident result at m points (reference:
// src/regular_functions_utils.rs:228-237; the circuit evaluates each f_k at the challenge point, src/config.rs:166-187).
// Horner over 2^19 coefficients is serial, so the evaluation is a dot product with a table of powers of x:
//   k_pow_step   X[j][2^s + i] = X[j][i] * X[j][2^s], 1 <= i <= 2^s   (log2(len) launches build x^0 .. x^(len-1) for all m points)
//   k_eval_chunks   per (chunk of 2048 coefficients, function, point): partial sums of a_k . X_j and b_k . X_j
//   k_eval_finish   out[k][j] = sum of a-partials + y_j * sum of b-partials
// Field addition is exact, so the order of the partial sums does not change a bit of the result.
// ------------------------------------------------------------------------------------------------
template <class FP>
__global__ void k_pow_init(const Affine<FP>* __restrict__ pts, size_t m, size_t stride, Fe<FP>* __restrict__ X) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    stg(X + j * stride, Fe<FP>::one());
    if (stride > 1) stg(X + j * stride + 1, ldg_aff(pts + j).x);
}
template <class FP>
__global__ void k_pow_step(Fe<FP>* __restrict__ X, size_t stride, size_t half /* 2^s: x^0 .. x^half are known */, size_t len) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + 1;   // 1 .. half
    Fe<FP>* row = X + (size_t)blockIdx.y * stride;
    if (i > half || half + i >= len) return;
    stg(row + half + i, mul(ldg(row + i), ldg(row + half)));
}
constexpr int EVAL_THREADS = 256;
constexpr int EVAL_PER_THREAD = 8;
constexpr int EVAL_CHUNK = EVAL_THREADS * EVAL_PER_THREAD;
template <class FP>
EAGEN_D Fe<FP> block_sum(Fe<FP> v, Fe<FP>* sm) {
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int h = EVAL_THREADS / 2; h > 0; h >>= 1) {
        if ((int)threadIdx.x < h) sm[threadIdx.x] = add(sm[threadIdx.x], sm[threadIdx.x + h]);
        __syncthreads();
    }
    Fe<FP> r = sm[0];
    __syncthreads();
    return r;
}
template <class FP>
__global__ void __launch_bounds__(EVAL_THREADS)
k_eval_chunks(const Fe<FP>* __restrict__ A, size_t a_stride, const Fe<FP>* __restrict__ B, size_t b_stride, const int* __restrict__ la,
              const int* __restrict__ lb, const Fe<FP>* __restrict__ X, size_t x_stride, int nchunks,
              Fe<FP>* __restrict__ part /* [point][function][chunk][2] */) {
    __shared__ Fe<FP> sm[EVAL_THREADS];
    const int c = blockIdx.x, k = blockIdx.y, j = blockIdx.z, nf = gridDim.y;
    const Fe<FP>* xr = X + (size_t)j * x_stride;
    const int na = la[k], nb = lb[k];
    Fe<FP> sa = Fe<FP>::zero(), sb = Fe<FP>::zero();
    if (c * EVAL_CHUNK < na || c * EVAL_CHUNK < nb) {
#pragma unroll 2
        for (int r = 0; r < EVAL_PER_THREAD; ++r) {
            const int i = c * EVAL_CHUNK + r * EVAL_THREADS + (int)threadIdx.x;
            if (i >= na && i >= nb) break;
            const Fe<FP> xp = ldg(xr + i);
            if (i < na) sa = add(sa, mul(ldg(A + (size_t)k * a_stride + i), xp));
            if (i < nb) sb = add(sb, mul(ldg(B + (size_t)k * b_stride + i), xp));
        }
    }
    sa = block_sum(sa, sm);
    sb = block_sum(sb, sm);
    if (threadIdx.x == 0) {
        Fe<FP>* o = part + (((size_t)j * nf + k) * nchunks + c) * 2;
        stg(o, sa); stg(o + 1, sb);
    }
}
template <class FP>
__global__ void k_eval_finish(const Fe<FP>* __restrict__ part, int nchunks, int nf, const Affine<FP>* __restrict__ pts, size_t m,
                              Fe<FP>* __restrict__ out /* [function][point] */) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)nf * m) return;
    const size_t k = g / m, j = g - k * m;
    const Fe<FP>* p = part + ((j * nf + k) * (size_t)nchunks) * 2;
    Fe<FP> sa = Fe<FP>::zero(), sb = Fe<FP>::zero();
    for (int c = 0; c < nchunks; ++c) { sa = add(sa, ldg(p + 2 * c)); sb = add(sb, ldg(p + 2 * c + 1)); }
    const Affine<FP> pt = ldg_aff(pts + j);
    stg(out + g, pt.is_identity() ? Fe<FP>::zero() : add(sa, mul(sb, pt.y)));   // like k_eval_function: 0 at the identity
}

// ------------------------------------------------------------------------------------------------
// Synthetic inputs for tests and bench.py: scalars uniform in [0, 2^127) from SplitMix64,
// points P_j = (a + j*b) * G for seed-derived 64-bit a, b (distinct points, no structure a kernel could exploit),
// emitted as Jacobian triples with non-trivial z.
// ------------------------------------------------------------------------------------------------
EAGEN_HD uint64_t splitmix64(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <class CC>
__global__ void k_synth_inputs(uint64_t seed, size_t n, int scalar_bits /* <= 127: below isqrt(order)+2 */,
                               Fe<typename CC::Scalar>* __restrict__ scalars, Fe<typename CC::Base>* __restrict__ jac) {
    typedef typename CC::Base F;
    typedef typename CC::Scalar S;
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t st = seed ^ (0xD1B54A32D192ED03ull * (j + 1));
    uint64_t lo = splitmix64(st), hi = splitmix64(st) >> (128 - scalar_bits);  // scalar_bits in (64, 127]
    Fe<S> sc = Fe<S>::zero();
    sc.v[0] = (uint32_t)lo; sc.v[1] = (uint32_t)(lo >> 32); sc.v[2] = (uint32_t)hi; sc.v[3] = (uint32_t)(hi >> 32);
    stg(scalars + j, from_canonical(sc));
    uint64_t s2 = seed;
    uint64_t a = splitmix64(s2) | 1, b = splitmix64(s2) | 1;
    // k = a + j*b as a 128-bit integer
    unsigned long long klo = a + (unsigned long long)j * b;
    unsigned long long khi = __umul64hi((unsigned long long)j, b) + (klo < a ? 1ull : 0ull);
    Affine<F> g; g.x = CC::gx(); g.y = CC::gy();
    Proj<F> acc = Proj<F>::identity();
    for (int bit = 127; bit >= 0; --bit) {
        acc = pdbl<CC>(acc);
        unsigned long long w = bit >= 64 ? khi : klo;
        if ((w >> (bit & 63)) & 1) acc = padd_mixed<CC>(acc, g);
    }
    // homogeneous (x : y : z) -> Jacobian (x z, y z^2, z)
    Fe<F> zz = sqr(acc.z);
    stg(jac + 3 * j, mul(acc.x, acc.z));
    stg(jac + 3 * j + 1, mul(acc.y, zz));
    stg(jac + 3 * j + 2, acc.z);
}

// ------------------------------------------------------------------------------------------------
// Windowed bucket MSM over full-width scalars (the `best_multiexp` the reference's tests compare the carry with:
// src/argument_witness_calc.rs:144, src/regular_functions_utils.rs:655).  c = 8 bit windows = the bytes of the canonical
// scalar; per window a counting sort of the point indices by digit, one warp per bucket (complete mixed additions, shared
// memory tree reduce), a running-sum pass per window and a Horner combine.  SURVEY.md section 8f rank 1.
// ------------------------------------------------------------------------------------------------
constexpr int MSM_WINDOWS = 32;
constexpr int MSM_CHUNK = 1024;

template <class FS>
__global__ void k_msm_digits(const Fe<FS>* __restrict__ scalars, size_t n, uint8_t* __restrict__ digits /* 32 x n */) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Fe<FS> x = to_canonical(ldg(scalars + j));
#pragma unroll
    for (int w = 0; w < MSM_WINDOWS; ++w) digits[(size_t)w * n + j] = (uint8_t)(x.v[w >> 2] >> (8 * (w & 3)));
}

static __global__ void k_msm_hist(const uint8_t* __restrict__ digits, size_t n, int chunks, uint32_t* __restrict__ hist /* 32 x chunks x 256 */) {
    __shared__ uint32_t cnt[256];
    const int w = blockIdx.y;
    cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < MSM_CHUNK; k += blockDim.x) {
        size_t j = (size_t)blockIdx.x * MSM_CHUNK + k;
        if (j < n) atomicAdd(&cnt[digits[(size_t)w * n + j]], 1u);
    }
    __syncthreads();
    hist[((size_t)w * chunks + blockIdx.x) * 256 + threadIdx.x] = cnt[threadIdx.x];
}

// one block of 256 threads per window: bucket starts, then per-(chunk, bucket) write offsets in place of the counts
static __global__ void k_msm_scan(uint32_t* __restrict__ hist, int chunks, uint32_t* __restrict__ bstart /* 32 x 257 */) {
    __shared__ uint32_t tot[256];
    const int w = blockIdx.x, b = threadIdx.x;
    uint32_t* h = hist + (size_t)w * chunks * 256;
    uint32_t t = 0;
    for (int c = 0; c < chunks; ++c) t += h[(size_t)c * 256 + b];
    tot[b] = t;
    __syncthreads();
    if (b == 0) {
        uint32_t off = 0;
        for (int i = 0; i < 256; ++i) { uint32_t v = tot[i]; tot[i] = off; bstart[w * 257 + i] = off; off += v; }
        bstart[w * 257 + 256] = off;
    }
    __syncthreads();
    uint32_t off = tot[b];
    for (int c = 0; c < chunks; ++c) { uint32_t v = h[(size_t)c * 256 + b]; h[(size_t)c * 256 + b] = off; off += v; }
}

static __global__ void k_msm_scatter(const uint8_t* __restrict__ digits, size_t n, int chunks, const uint32_t* __restrict__ hist,
                                     uint32_t* __restrict__ sorted /* 32 x n */) {
    __shared__ uint32_t cur[256];
    const int w = blockIdx.y;
    cur[threadIdx.x] = hist[((size_t)w * chunks + blockIdx.x) * 256 + threadIdx.x];
    __syncthreads();
    for (int k = threadIdx.x; k < MSM_CHUNK; k += blockDim.x) {
        size_t j = (size_t)blockIdx.x * MSM_CHUNK + k;
        if (j < n) {
            uint32_t pos = atomicAdd(&cur[digits[(size_t)w * n + j]], 1u);   // order inside a bucket is irrelevant for a sum
            sorted[(size_t)w * n + pos] = (uint32_t)j;
        }
    }
}

// one warp per (window, bucket >= 1)
template <class CC>
__global__ void __launch_bounds__(128)
k_msm_buckets(const Affine<typename CC::Base>* __restrict__ pts, size_t n, const uint32_t* __restrict__ sorted,
              const uint32_t* __restrict__ bstart, Proj<typename CC::Base>* __restrict__ buckets /* 32 x 256 */) {
    typedef typename CC::Base F;
    __shared__ Proj<F> sm[128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * 4 + wid;          // global warp = w * 255 + (bucket - 1)
    const int w = gw / 255, b = gw % 255 + 1;
    Proj<F> acc = Proj<F>::identity();
    if (w < MSM_WINDOWS) {
        const uint32_t lo = bstart[w * 257 + b], hi = bstart[w * 257 + b + 1];
        for (uint32_t i = lo + lane; i < hi; i += 32) {
            Affine<F> q = ldg_aff(pts + sorted[(size_t)w * n + i]);
            if (!q.is_identity()) acc = padd_mixed<CC>(acc, q);
        }
    }
    sm[threadIdx.x] = acc;
    __syncwarp();
    for (int s = 16; s > 0; s >>= 1) {
        if (lane < s) sm[threadIdx.x] = padd<CC>(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncwarp();
    }
    if (lane == 0 && w < MSM_WINDOWS) buckets[w * 256 + b] = sm[threadIdx.x];
}

// S_w = sum_b b * B_b by running sums; one thread per window
template <class CC>
__global__ void k_msm_window_sums(const Proj<typename CC::Base>* __restrict__ buckets, Proj<typename CC::Base>* __restrict__ wsum) {
    typedef typename CC::Base F;
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= MSM_WINDOWS) return;
    Proj<F> run = Proj<F>::identity(), sum = Proj<F>::identity();
    for (int b = 255; b >= 1; --b) {
        run = padd<CC>(run, buckets[w * 256 + b]);
        sum = padd<CC>(sum, run);
    }
    wsum[w] = sum;
}

template <class CC>
__global__ void k_msm_combine(const Proj<typename CC::Base>* __restrict__ wsum, Proj<typename CC::Base>* __restrict__ out, Fe<typename CC::Base>* __restrict__ z) {
    typedef typename CC::Base F;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Proj<F> r = Proj<F>::identity();
    for (int w = MSM_WINDOWS - 1; w >= 0; --w) {
        for (int k = 0; k < 8; ++k) r = pdbl<CC>(r);
        r = padd<CC>(r, wsum[w]);
    }
    out[0] = r;
    stg(z, r.z);
}

// ------------------------------------------------------------------------------------------------
// Integer-pipe micro-benchmarks: the roofline denominators MEASURED_PEAKS.json does not carry.
//   k_imad_peak   : 16 independent 32-bit IMAD chains per thread, no memory traffic  -> IMAD/s of the chip
//   k_modmul_peak : 4 independent Montgomery-product chains per thread               -> modmul/s ceiling of mul()
// ------------------------------------------------------------------------------------------------
static __global__ void k_imad_peak(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 16 + i;
    uint32_t m = seed | 1, c = blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] * m + c;
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// 16 independent carry-chained 32x32+64 multiply-accumulates per thread (mad.lo.cc / madc.hi pairs, the form the Montgomery
// product issues: ptxas fuses each pair into one IMAD.WIDE.U32) -> IMAD.WIDE per second: the multiplier-pipe ceiling
static __global__ void k_imad_wide_peak(uint32_t* out, int iters, uint32_t seed) {
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo[i] = seed + threadIdx.x * 16 + i; hi[i] = seed * 3 + i; }
    uint32_t m = seed | 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                unsigned long long t = (unsigned long long)lo[i] * m + (((unsigned long long)hi[i] << 32) | lo[i]);
                lo[i] = (uint32_t)t; hi[i] = (uint32_t)(t >> 32);
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= lo[i] ^ hi[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <class FP>
__global__ void k_modmul_peak(Fe<FP>* out, int iters) {
    Fe<FP> a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = Fe<FP>::one(); a[i].v[0] += threadIdx.x * 4 + i; a[i].v[1] = blockIdx.x; }
    Fe<FP> m = ldg(out);  // runtime multiplier (a constant would be folded into immediates and is not what kernels see)
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = mul(a[i], m);
    }
    Fe<FP> x = add(add(a[0], a[1]), add(a[2], a[3]));
    stg(out + (size_t)blockIdx.x * blockDim.x + threadIdx.x, x);
}

}  // namespace eagen
