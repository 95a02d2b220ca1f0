// throwaway pipe-throughput probes (not part of the product)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void probe(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 8 + i; b[i] = seed * 3 + i; }
    uint32_t m = seed | 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i] = a[i] * m + b[i]; }                                  // IMAD lo
                if (MODE == 1) { unsigned long long t = (unsigned long long)a[i] * m + (((unsigned long long)b[i] << 32) | a[i]); a[i] = (uint32_t)t; b[i] = (uint32_t)(t >> 32); }  // IMAD.WIDE
                if (MODE == 2) { a[i] = __umulhi(a[i], m) + b[i]; }                          // IMAD.HI
                if (MODE == 3) { asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(a[(i+1)&7]), "r"(m)); } // chained wide with carry
                if (MODE == 4) { a[i] = a[i] + b[i] + m; }                                   // IADD3
                if (MODE == 5) { asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }  // IADD3 with carry out
            }
        }
    }
    uint32_t x = 0;
    for (int i = 0; i < 8; ++i) x ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int MODE> void run(const char* name) {
    uint32_t* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4096; float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); probe<MODE><<<148 * 8, 256>>>(d, iters, 12345 + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = 148.0 * 8 * 256 * iters * 64;
    printf("%-28s %.2f Tops/s  (%.1f per clk per SM at 1.965 GHz)\n", name, ops / (best * 1e-3) / 1e12, ops / (best * 1e-3) / 148 / 1.965e9);
}
int main() {
    run<0>("IMAD (lo)"); run<1>("IMAD.WIDE 64-bit addend"); run<2>("IMAD.HI"); run<3>("mad.lo.cc+madc.hi.cc pair"); run<4>("IADD3"); run<5>("add.cc");
    return 0;
}
