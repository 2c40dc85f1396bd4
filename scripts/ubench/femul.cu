// Micro-benchmark: cycles per fe_mul / fe_sq (blobstreamx_b200/csrc/ed25519.cuh) as a function of warps per SM
// sub-partition, for the variant selected at compile time (-DBSX_FE_SCHOOLBOOK or the default pair-Karatsuba form).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a [-DBSX_FE_SCHOOLBOOK] -o femul femul.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../blobstreamx_b200/csrc/ed25519.cuh"
using namespace bsx::ed;

#define ITERS 2000

// MODE 0: chain of fe_mul calls; 1: chain of fe_sq calls; 2: two independent fe_mul chains (calls);
// 3: point doubling + p1p1->p3 (the inner loop of h*A)
template <int MODE>
__global__ void k(int32_t *out, long long *cyc, int32_t seed) {
    fe x, y, z;
    for (int i = 0; i < 10; i++) { x.v[i] = seed + threadIdx.x + i; y.v[i] = seed * 3 + i * threadIdx.x; z.v[i] = seed - i; }
    ge_p3 p; p.X = x; p.Y = y; p.Z = z; p.T = x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) x = fe_mul(x, y);
        if (MODE == 1) x = fe_sq(x);
        if (MODE == 2) { x = fe_mul(x, y); z = fe_mul(z, y); }
        if (MODE == 3) p = ge_p1p1_to_p3(ge_dbl(p), false);
    }
    long long t1 = clock64();
    int32_t s = 0;
    for (int i = 0; i < 10; i++) s += x.v[i] + z.v[i] + p.X.v[i] + p.Y.v[i] + p.Z.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps_per_sm, int n_sm, double ops) {
    int32_t *out; cudaMalloc(&out, 4ull * 2048 * 1024);
    long long *cyc; cudaMallocManaged(&cyc, 8);
    k<MODE><<<n_sm, 32 * warps_per_sm>>>(out, cyc, 12345);
    cudaDeviceSynchronize();
    k<MODE><<<n_sm, 32 * warps_per_sm>>>(out, cyc, 12345);
    cudaDeviceSynchronize();
    double per_warp = (double)*cyc / (ITERS * ops);
    printf("%-22s warps/SMSP=%4.1f  cycles/op/warp=%7.1f   cycles/op/SMSP (throughput)=%7.1f\n", name, warps_per_sm / 4.0, per_warp,
           per_warp / (warps_per_sm / 4.0));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
#ifdef BSX_FE_SCHOOLBOOK
    printf("variant: schoolbook 10x10\n");
#else
    printf("variant: pair Karatsuba\n");
#endif
    for (int w : {4, 8, 12, 16, 24, 32}) {
        run<0>("fe_mul chain", w, n_sm, 1);
        run<1>("fe_sq chain", w, n_sm, 1);
        run<2>("2 fe_mul chains", w, n_sm, 2);
        run<3>("dbl+p3 (4S+3M)", w, n_sm, 1);
    }
    return 0;
}
