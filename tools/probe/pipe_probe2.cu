// throwaway pipe co-issue probes (not part of the product): which instruction classes share the multiplier pipe with
// IMAD.WIDE on sm_100a, and what the FP64 pipe could add.  Build ON the GPU box:
//   nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe2 pipe_probe2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// per iteration and array slot: W wide ops, then X "other" ops of kind MODE
template <int MODE, int NW, int NX>
__global__ void probe(uint32_t* out, int iters, uint32_t seed, double dseed) {
    uint32_t a[8], b[8], c[8];
    double f[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 8 + i; b[i] = seed * 3 + i; c[i] = seed * 7 + i; f[i] = dseed + i; }
    uint32_t m = seed | 1;
    double fm = dseed * 1.0000001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    unsigned long long t = (unsigned long long)a[i] * m + (((unsigned long long)b[i] << 32) | a[i]);
                    a[i] = (uint32_t)t; b[i] = (uint32_t)(t >> 32);
                }
#pragma unroll
                for (int x = 0; x < NX; ++x) {
                    if (MODE == 0) c[i] = c[i] * m + b[(i + 1) & 7];                     // IMAD lo
                    if (MODE == 1) c[i] = c[i] + b[(i + 1) & 7] + m;                     // IADD3
                    if (MODE == 2) f[i] = fma(f[i], fm, f[(i + 1) & 7]);                 // DFMA
                    if (MODE == 3) asm volatile("mov.b32 %0, %1;" : "=r"(c[i]) : "r"(c[(i + 1) & 7]));   // MOV
                    if (MODE == 4) c[i] = __umulhi(c[i], m) + b[(i + 1) & 7];            // IMAD.HI
                    if (MODE == 5) c[i] = (c[i] << 3) ^ (c[(i + 1) & 7] >> 5);           // SHF/LOP3
                }
            }
        }
    }
    uint32_t x = 0;
    for (int i = 0; i < 8; ++i) x ^= a[i] ^ b[i] ^ c[i] ^ (uint32_t)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int MODE, int NW, int NX> void run(const char* name) {
    uint32_t* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2048; float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); probe<MODE, NW, NX><<<148 * 8, 256>>>(d, iters, 12345 + rep, 1.000001 + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double slots = 148.0 * 8 * 256 * iters * 32;   // (array slot, round) pairs executed by all threads
    double clk_per_slot_per_smsp = best * 1e-3 * 1.965e9 / (slots / 32 / (148 * 4));   // cycles an SMSP spends per warp-level slot
    printf("%-40s wide %d + other %d per slot: %.2f clk per warp-slot per SMSP  (wide alone would be %d x 4 = %d)\n", name, NW, NX, clk_per_slot_per_smsp, NW, NW * 4);
    cudaFree(d);
}
int main() {
    run<0, 1, 0>("IMAD.WIDE only");
    run<0, 0, 1>("IMAD lo only");
    run<1, 0, 1>("IADD3 only");
    run<2, 0, 1>("DFMA only");
    run<4, 0, 1>("IMAD.HI only");
    run<0, 1, 1>("IMAD.WIDE + IMAD lo");
    run<0, 1, 2>("IMAD.WIDE + 2 IMAD lo");
    run<1, 1, 2>("IMAD.WIDE + 2 IADD3");
    run<1, 1, 3>("IMAD.WIDE + 3 IADD3");
    run<2, 1, 1>("IMAD.WIDE + DFMA");
    run<2, 1, 2>("IMAD.WIDE + 2 DFMA");
    run<2, 1, 4>("IMAD.WIDE + 4 DFMA");
    run<3, 1, 2>("IMAD.WIDE + 2 MOV");
    run<5, 1, 2>("IMAD.WIDE + 2 SHF/LOP3 pairs");
    run<4, 1, 1>("IMAD.WIDE + IMAD.HI");
    return 0;
}
