/* eagen_msm_selftest.h -- host-side self-test hooks of libeagen_msm.so.
 *
 * The field and curve arithmetic in csrc/field.cuh and csrc/curve.cuh is written once as
 * __host__ __device__ code; these entry points run that same source on the HOST so that the CPU test
 * suite (no GPU in the build container) can check it against the oracle.  They are diagnostics only:
 * no product entry point of eagen_msm.h routes through them, and they compute single operations, not the path.
 */
#ifndef EAGEN_MSM_SELFTEST_H
#define EAGEN_MSM_SELFTEST_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* field ids: 0 pallas_fp, 1 pallas_fq, 2 bn256_fr, 3 bn256_fq
 * op: 0 add, 1 sub, 2 mul, 3 inv(a), 4 from_canonical(a), 5 to_canonical(a), 6 mul by the device's carry-chain
 * algorithm with the carry flag emulated on the host    (Montgomery 32-byte elements);
 * 7 mul_lazy, 8 add_lazy, 9 sub_lazy, 10 normalise_lazy: the transform's lazily reduced arithmetic (csrc/field.cuh) run on the host
 * from the same source -- operands are raw 256-bit values in [0, 2p) (b < p for op 7), results of 7-9 come back in [0, 2p) */
int eagen_selftest_field(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out);
/* op: 0 complete add (p, q Jacobian), 1 double p, 2 mixed add (q must be z = 1), 3 small multiple k*p.
 * Inputs are Jacobian (96 B); out is affine (64 B), identity = zeros. */
int eagen_selftest_curve(int curve, int op, const uint64_t* p, const uint64_t* q, uint32_t k, uint64_t* out_affine);
/* K1 constants for (curve, base): d, digits per table group, words of four positions, and the limbs
 * sq | K | b^d | ceil(2^288 / b^d) (4 x 32 bytes) */
int eagen_selftest_negbase_params(int curve, uint8_t base, uint32_t* d, uint32_t* group, uint32_t* words, uint32_t* limbs32);
/* K1's per-scalar arithmetic (Montgomery -> canonical, division-free digit extraction, complement table) on the host:
 * d digits, most significant first, of one Montgomery scalar; *kerr = 0, 1 (range) or 2 (more than d digits) */
int eagen_selftest_negbase_digits(int curve, uint8_t base, const uint64_t* scalar, uint8_t* digits, int* kerr);
/* the pass plan of a 2^t transform: writes up to 8 (s_hi, s_lo) pairs, returns the count */
int eagen_selftest_ntt_plan(int t, int* pairs16);
#ifdef __cplusplus
}
#endif
#endif
