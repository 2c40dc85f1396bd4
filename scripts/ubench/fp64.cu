// Micro-benchmark (B200): issue rate of the FP64 pipe against the integer pipes, and cycles per field multiplication /
// squaring of the FP64-limb form (blobstreamx_b200/csrc/fe51d.cuh) beside the IMAD.WIDE form (ed25519.cuh).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64 fp64.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../blobstreamx_b200/csrc/ed25519.cuh"
#include "../../blobstreamx_b200/csrc/fe51d.cuh"
using namespace bsx::ed;
using namespace bsx::edd;

#define ITERS 2000

// MODE 0: 8 independent DFMA chains; 1: 8 independent IMAD.WIDE chains; 2: 8 DFMA + 8 IMAD.WIDE; 3: 8 DFMA + 16 IADD3/LOP3;
// 4: 8 DFMA + 8 IMAD.WIDE + 16 ALU
template <int MODE>
__global__ void pipes(double *out, long long *cyc, double seed) {
    double a[8]; uint64_t m[8]; uint32_t u[16];
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 0.5 + i; m[i] = (uint64_t)(seed * 77) + threadIdx.x + i; }
    for (int i = 0; i < 16; i++) u[i] = threadIdx.x * 3 + i;
    const double b = seed * 1e-9, c = seed * 1e-3;
    const uint32_t k = (uint32_t)seed | 1;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (MODE == 0 || MODE >= 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = __fma_rz(a[i], b, c);
            }
            if (MODE == 1 || MODE == 2 || MODE == 4) {
#pragma unroll
                for (int i = 0; i < 8; i++) m[i] = (uint64_t)(uint32_t)m[i] * k + m[i];
            }
            if (MODE == 3 || MODE == 4) {
#pragma unroll
                for (int i = 0; i < 16; i++) u[i] = (u[i] ^ k) + (u[(i + 1) & 15] & 0x55555555u);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + (double)m[i];
    for (int i = 0; i < 16; i++) s += u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// MODE 0: fed_mul chain (call); 1: fed_sq chain (call); 2: fe_mul (int) chain; 3: fe_sq (int) chain;
// 4: 4 independent fed_mul inlined; 5: 4 independent fed_sq inlined
template <int MODE>
__global__ void field(double *out, long long *cyc, int32_t seed) {
    fed x[4], y;
    fe xi, yi;
    for (int i = 0; i < 5; i++) {
        for (int q = 0; q < 4; q++) x[q].v[i] = (double)((int64_t)(seed + threadIdx.x + i + q) << 20);
        y.v[i] = (double)((int64_t)(seed * 3 + i * threadIdx.x + 7) << 18);
    }
    for (int i = 0; i < 10; i++) { xi.v[i] = seed + threadIdx.x + i; yi.v[i] = seed * 3 + i * threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) x[0] = fed_mul(x[0], y);
        if (MODE == 1) x[0] = fed_sq(x[0]);
        if (MODE == 2) xi = fe_mul(xi, yi);
        if (MODE == 3) xi = fe_sq(xi);
        if (MODE == 4) { for (int q = 0; q < 4; q++) x[q] = fed_mul_inl(x[q], y); }
        if (MODE == 5) { for (int q = 0; q < 4; q++) x[q] = fed_sq_impl<false>(x[q]); }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 5; i++) for (int q = 0; q < 4; q++) s += x[q].v[i];
    for (int i = 0; i < 10; i++) s += xi.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename K>
void run(K kern, const char *name, int warps_per_sm, int n_sm, double ops, bool is_field) {
    double *out; cudaMalloc(&out, 8ull * 2048 * 1024);
    long long *cyc; cudaMallocManaged(&cyc, 8);
    for (int rep = 0; rep < 2; rep++) {
        if (is_field) ((void (*)(double *, long long *, int32_t))kern)<<<n_sm, 32 * warps_per_sm>>>(out, cyc, 12345);
        else ((void (*)(double *, long long *, double))kern)<<<n_sm, 32 * warps_per_sm>>>(out, cyc, 12345.0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    double per_warp = (double)*cyc / (ITERS * ops);
    printf("%-34s warps/SMSP=%4.1f  cycles/op/warp=%8.2f   cycles/op/SMSP=%8.2f\n", name, warps_per_sm / 4.0, per_warp,
           per_warp / (warps_per_sm / 4.0));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs=%d  (op = one group: 32 DFMA / 32 IMAD.WIDE / 64 ALU per loop pass as listed)\n", n_sm);
    for (int w : {4, 8, 16, 32}) {
        run((void *)pipes<0>, "32 DFMA", w, n_sm, 1, false);
        run((void *)pipes<1>, "32 IMAD.WIDE", w, n_sm, 1, false);
        run((void *)pipes<2>, "32 DFMA + 32 IMAD.WIDE", w, n_sm, 1, false);
        run((void *)pipes<3>, "32 DFMA + 128 ALU ops", w, n_sm, 1, false);
        run((void *)pipes<4>, "32 DFMA + 32 IMAD.WIDE + 128 ALU", w, n_sm, 1, false);
    }
    for (int w : {4, 8, 12, 16, 24, 32}) {
        run((void *)field<0>, "fed_mul chain (FP64 limbs)", w, n_sm, 1, true);
        run((void *)field<1>, "fed_sq chain (FP64 limbs)", w, n_sm, 1, true);
        run((void *)field<4>, "4 fed_mul inlined", w, n_sm, 4, true);
        run((void *)field<5>, "4 fed_sq inlined", w, n_sm, 4, true);
        run((void *)field<2>, "fe_mul chain (IMAD.WIDE limbs)", w, n_sm, 1, true);
        run((void *)field<3>, "fe_sq chain (IMAD.WIDE limbs)", w, n_sm, 1, true);
    }
    return 0;
}
