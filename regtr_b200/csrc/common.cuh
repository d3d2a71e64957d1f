// Shared helpers for the regtr_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/regtr_b200.h"

#define REGTR_NUM_SMS 148

#define REGTR_CHECK_LAUNCH()                                         \
    do {                                                             \
        cudaError_t e__ = cudaGetLastError();                        \
        if (e__ != cudaSuccess) return -(1000 + (int)e__);           \
    } while (0)

static inline int regtr_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t regtr_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// 64-bit spatial key: [cloud:16][x:16][y:16][z:16], coordinates biased by 32768.
__device__ __forceinline__ unsigned long long regtr_pack_key(int cloud, int x, int y, int z) {
    return ((unsigned long long)(unsigned)cloud << 48) | ((unsigned long long)(unsigned)(x + 32768) << 32) |
           ((unsigned long long)(unsigned)(y + 32768) << 16) | (unsigned long long)(unsigned)(z + 32768);
}

// floor(p / cell) with IEEE division -- the pinned voxel rule.
__device__ __forceinline__ int regtr_cell_of(float p, float cell) { return (int)floorf(__fdiv_rn(p, cell)); }

// Index of the cloud that owns packed row i: largest c with offs[c] <= i.
__device__ __forceinline__ int regtr_cloud_of(const int32_t* __restrict__ offs, int n_clouds, int i) {
    int lo = 0, hi = n_clouds;  // invariant: offs[lo] <= i < offs[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (offs[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
