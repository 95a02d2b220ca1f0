// throwaway probe: Montgomery-product throughput as a function of independent chains per thread and warps per SM
#include <cstdio>
#include <cuda_runtime.h>
#include "../../halo2-liam-eagen-msm_b200/csrc/field.cuh"
using namespace eagen;
typedef Fe<PallasFp> F;
template <int CH>
__global__ void __launch_bounds__(256) probe(F* out, int iters) {
    F a[CH];
    for (int i = 0; i < CH; ++i) { a[i] = F::one(); a[i].v[0] += threadIdx.x * CH + i; a[i].v[1] = blockIdx.x; }
    F m = out[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = mul(a[i], m);
    }
    F x = a[0];
    for (int i = 1; i < CH; ++i) x = add(x, a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int CH> void run(int blocks_per_sm) {
    F* d; cudaMalloc(&d, 148 * 8 * 256 * 32); cudaMemset(d, 1, 148 * 8 * 256 * 32);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2048 / CH; float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); probe<CH><<<148 * blocks_per_sm, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = 148.0 * blocks_per_sm * 256 * iters * CH;
    printf("chains/thread %d, blocks/SM %d (warps/SM %d): %.1f G modmul/s\n", CH, blocks_per_sm, blocks_per_sm * 8, ops / (best * 1e-3) / 1e9);
    cudaFree(d);
}
int main() {
    for (int b : {1, 2, 4, 8}) run<1>(b);
    for (int b : {1, 2, 4, 8}) run<2>(b);
    for (int b : {1, 2, 4}) run<4>(b);
    return 0;
}
