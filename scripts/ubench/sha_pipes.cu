// Micro-benchmark: cycles per SHA-256 compression (blobstreamx_b200/csrc/sha256.cuh) per SM sub-partition for a given
// BSX_SHA_FMA_MASK (which rotation groups run as mul.hi + mad.lo on the FMA pipe instead of funnel shifts on the ALU pipe).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DBSX_SHA_FMA_MASK=<m> -o sha_pipes_<m> sha_pipes.cu
// NEGATIVE RESULT (profiles/r01l_ubench_sha_pipes.txt): no mask beats 0.  Replacing the 96 plain shifts by IMAD.HI (mask 16)
// removes 7.5 % of the ALU-pipe instructions and changes nothing (129.9 vs 128.7 ms); moving Sigma0+Sigma1 (mask 3) is 9 %
// slower.  IMAD.HI costs the FMA pipe 4 cycles like IMAD.WIDE and the two integer pipes do not overlap the way their
// separate utilisation counters suggest, so sha256.cuh keeps funnel shifts.  To re-run this file, re-apply the rotr32g /
// shr32g variant switch (see git history of this commit) to a copy of sha256.cuh.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../blobstreamx_b200/csrc/sha256.cuh"
using namespace bsx;
#define ITERS 2000
__global__ void __launch_bounds__(256) k(uint32_t *out, long long *cyc, uint32_t seed) {
    uint32_t st[8], w[16];
    sha256_init(st);
    for (int i = 0; i < 16; i++) w[i] = seed * (i + 1) + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        uint32_t ww[16];
        for (int i = 0; i < 16; i++) ww[i] = w[i] ^ st[i & 7];
        sha256_compress(st, ww);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= st[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out; cudaMalloc(&out, 4ull * 4096 * 1024);
    long long *cyc; cudaMallocManaged(&cyc, 8);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, 0);
    uint32_t ref = 0;
    const int blocks = n_sm * 24;              // several full waves at any occupancy
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<blocks, 256>>>(out, cyc, 77);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<blocks, 256>>>(out, cyc, 77);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&ref, out, 4, cudaMemcpyDeviceToHost);
    const double comp = (double)blocks * 256 * ITERS;
    printf("mask=%2d resident blocks/SM=%d  %.3f ms  %.2f G compressions/s  = %.1f GB/s of message blocks   check=%08x\n", BSX_SHA_FMA_MASK, occ, ms,
           comp / ms * 1e-6, comp * 64 / ms * 1e-6, ref);
    return 0;
}
