// Micro-benchmark: issue rate of IMAD.WIDE / IMAD / DFMA / IADD3 on sm_100a, alone and mixed.
// Cycles are counted on the SM (clock64), so the result is independent of the clock the GPU happens to run at.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 8192
#define ILP 8

// MODE bit 0: IMAD.WIDE, bit 1: IMAD (32), bit 2: DFMA, bit 3: LOP3 (ALU pipe), bit 4: IADD3 64-bit pair (ALU)
template <int MODE>
__global__ void k(uint64_t *out, long long *cyc, int32_t a0, int32_t b0, double d0) {
    int64_t acc[ILP];
    double dac[ILP];
    int32_t lo[ILP], al[ILP], xs[ILP];
    int32_t a = a0 + threadIdx.x, b = b0 + threadIdx.x;
    double da = d0 + threadIdx.x, db = d0 * 0.5;
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = i; dac[i] = i; lo[i] = i; al[i] = i; xs[i] = a * (i + 3); }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE & 1) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(xs[i]), "r"(b));
            if (MODE & 2) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(a), "r"(b));
            if (MODE & 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dac[i]) : "d"(da), "d"(db));
            if (MODE & 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(al[i]) : "r"(a), "r"(b));
            if (MODE & 16) asm volatile("add.s64 %0, %0, %1;" : "+l"(acc[i]) : "l"((int64_t)a));
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += (uint64_t)acc[i] + (uint64_t)dac[i] + (uint64_t)lo[i] + (uint64_t)al[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps_per_sm, int n_sm) {
    uint64_t *out; cudaMalloc(&out, 8ull * 2048 * 1024);
    long long *cyc; cudaMallocManaged(&cyc, 8);
    int threads = 32 * warps_per_sm;
    k<MODE><<<n_sm, threads>>>(out, cyc, 3, 5, 1.0000001);
    cudaDeviceSynchronize();
    k<MODE><<<n_sm, threads>>>(out, cyc, 3, 5, 1.0000001);
    cudaDeviceSynchronize();
    double per_kind = (double)ITERS * ILP * warps_per_sm;   // warp-instructions of EACH enabled kind per SM
    printf("%-26s warps/SM=%2d  cycles=%9lld  cycles per warp-inst of each kind per SM = %.3f  (-> %.2f lanes/clk/SM each)\n", name,
           warps_per_sm, *cyc, *cyc / per_kind, 32.0 * per_kind / *cyc);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {4, 8, 16, 32}) {
        run<1>("IMAD.WIDE", w, n_sm);
        run<2>("IMAD", w, n_sm);
        run<4>("DFMA", w, n_sm);
        run<8>("LOP3", w, n_sm);
        run<16>("ADD64 (2 IADD3)", w, n_sm);
        run<1 | 4>("IMAD.WIDE+DFMA", w, n_sm);
        run<1 | 8>("IMAD.WIDE+LOP3", w, n_sm);
        run<1 | 2>("IMAD.WIDE+IMAD", w, n_sm);
        run<1 | 16>("IMAD.WIDE+ADD64", w, n_sm);
        run<1 | 8 | 4>("IMAD.WIDE+LOP3+DFMA", w, n_sm);
    }
    return 0;
}
