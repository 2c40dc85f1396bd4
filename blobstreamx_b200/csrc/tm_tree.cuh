// Fixed-shape Tendermint Merkle tree evaluation by one CTA (PX/frontend/merkle/tendermint.rs:124-204).
// Every pair of every layer is hashed and written out (each inner digest is a witness value);
// the value that moves up is select(both_enabled, inner, left).  Because the enabled mask is a
// prefix (running AND of i != nb_enabled, tendermint.rs:184-194), node i of layer s is enabled
// iff (i << s) < nb_enabled, so no mask array is needed.
// Wide layers go through shared memory (ping-pong, one __syncthreads per layer); once a layer
// has <= 64 nodes, warp 0 finishes alone with register shuffles.
#pragma once
#include "sha256.cuh"

namespace bsx {

// A: P digests (8 big-endian words each) in shared memory, Bf: scratch for P/2 digests.
// inner: global, (P-1)*32 bytes layer-major, may be nullptr.  All threads of the CTA must call.
// Returns the root in `root` for every lane of warp 0.
__device__ __forceinline__ void tm_tree_cta(uint32_t *A, uint32_t *Bf, uint32_t P, uint64_t nb, uint8_t *inner,
                                            uint32_t root[8]) {
    const uint32_t tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    uint32_t *src = A, *dst = Bf;
    uint32_t len = P, off = 0, shift = 0;
    while (len > 64) {
        for (uint32_t i = tid; i < len / 2; i += nthr) {
            uint32_t l[8], r[8], p[8];
            const uint4 *s = reinterpret_cast<const uint4 *>(src + 16 * i);
            uint4 a = s[0], b = s[1], c = s[2], d = s[3];
            l[0] = a.x; l[1] = a.y; l[2] = a.z; l[3] = a.w; l[4] = b.x; l[5] = b.y; l[6] = b.z; l[7] = b.w;
            r[0] = c.x; r[1] = c.y; r[2] = c.z; r[3] = c.w; r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = d.w;
            tm_inner_hash(l, r, p);
            if (inner) store_digest_be(inner + 32 * (size_t)(off + i), p);
            bool both = ((uint64_t)(2 * i + 1) << shift) < nb;
            uint4 *o = reinterpret_cast<uint4 *>(dst + 8 * i);
            o[0] = both ? make_uint4(p[0], p[1], p[2], p[3]) : a;
            o[1] = both ? make_uint4(p[4], p[5], p[6], p[7]) : b;
        }
        off += len / 2;
        len /= 2;
        shift++;
        uint32_t *t = src; src = dst; dst = t;
        __syncthreads();
    }
    if (tid >= 32) return;
    uint32_t L[8], R[8], sel[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { L[k] = 0; R[k] = 0; sel[k] = 0; }
    if (len == 1) {
#pragma unroll
        for (int k = 0; k < 8; k++) sel[k] = src[k];
    } else if (lane < len / 2) {
#pragma unroll
        for (int k = 0; k < 8; k++) { L[k] = src[16 * lane + k]; R[k] = src[16 * lane + 8 + k]; }
    }
    while (len > 1) {
        if (lane < len / 2) {
            uint32_t p[8];
            tm_inner_hash(L, R, p);
            if (inner) store_digest_be(inner + 32 * (size_t)(off + lane), p);
            bool both = ((uint64_t)(2 * lane + 1) << shift) < nb;
#pragma unroll
            for (int k = 0; k < 8; k++) sel[k] = both ? p[k] : L[k];
        }
        off += len / 2;
        len /= 2;
        shift++;
        if (len > 1) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t lo = __shfl_sync(0xffffffffu, sel[k], (2 * lane) & 31);
                uint32_t hi = __shfl_sync(0xffffffffu, sel[k], (2 * lane + 1) & 31);
                L[k] = lo;
                R[k] = hi;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) root[k] = __shfl_sync(0xffffffffu, sel[k], 0);
}

}  // namespace bsx
