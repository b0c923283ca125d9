// gp_common.cuh -- shared helpers for libgrandplus_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/grandplus_b200.h"

#include <nvtx3/nvToolsExt.h>   // header-only (dlopens the injection library when a profiler is attached)

// RAII NVTX range around a C-ABI entry point: `ncu --nvtx --nvtx-include "gp_gfpush/"` or an nsys timeline then show the
// call the kernels belong to (the reference has no tracing hooks at all).
struct GpRange {
    explicit GpRange(const char *name) { nvtxRangePushA(name); }
    ~GpRange() { nvtxRangePop(); }
    GpRange(const GpRange &) = delete;
    GpRange &operator=(const GpRange &) = delete;
};

#ifndef GP_NUM_SMS_FALLBACK
#define GP_NUM_SMS_FALLBACK 148  // B200: 2 dies x 74 SMs
#endif

// ---- error plumbing (thread-local message, integer status across the C ABI) -------------
void gp_set_error(const char *fmt, ...);

#define GP_CUDA_TRY(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            gp_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                         cudaGetErrorString(_e));                                           \
            return (_e == cudaErrorMemoryAllocation) ? GP_ERR_NOMEM : GP_ERR_CUDA;          \
        }                                                                                   \
    } while (0)

#define GP_REQUIRE(cond, ...)                                                               \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            gp_set_error(__VA_ARGS__);                                                      \
            return GP_ERR_INVALID;                                                          \
        }                                                                                   \
    } while (0)

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ int gp_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ int gp_warp() { return threadIdx.x >> 5; }

// Exclusive block scan of one unsigned per thread.  s_warp must hold BLOCK/32 + 1 words.
// Contains two __syncthreads(); every thread of the CTA must call it.
template <int BLOCK>
__device__ __forceinline__ unsigned gp_block_exclusive_scan(unsigned x, unsigned *s_warp, unsigned &total) {
    constexpr int NW = BLOCK / 32;
    const int lane = gp_lane(), wid = gp_warp();
    unsigned incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned w = (lane < NW) ? s_warp[lane] : 0u;
        unsigned wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        if (lane < NW) s_warp[lane] = wi - w;
        if (lane == NW - 1) s_warp[NW] = wi;
    }
    __syncthreads();
    total = s_warp[NW];
    return incl - x + s_warp[wid];
}

// 128-bit read-only streaming load (LDG.E.128.CONSTANT); rows are touched once per CTA.
__device__ __forceinline__ float4 gp_ldg_f4(const float4 *p) { return __ldg(p); }

// Philox4x32-10 (Salmon et al. 2011), counter = (c0,c1,c2,c3), key = (k0,k1).
__host__ __device__ __forceinline__ void gp_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// DropNode keep decisions of entry j for up to four augmentations from ONE Philox block: counter = (j lo, j hi, 0,
// offset lo), key = seed ^ (offset hi folded); augmentation a keeps the entry iff word a >= p * 2^32
// (P[keep] = 1-p to within 2^-32).  Bit a of the result = keep in augmentation a.
__host__ __device__ __forceinline__ uint32_t gp_keep_threshold(float p) {
    double t = (double)p * 4294967296.0;
    if (t <= 0.0) return 0u;
    if (t >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)t;
}
__host__ __device__ __forceinline__ uint32_t gp_dropnode_keep4(uint64_t j, uint64_t seed, uint64_t offset, uint32_t thresh) {
    uint32_t r[4];
    gp_philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), 0u, (uint32_t)offset,
                     (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32), r);
    return (r[0] >= thresh ? 1u : 0u) | (r[1] >= thresh ? 2u : 0u) | (r[2] >= thresh ? 4u : 0u) | (r[3] >= thresh ? 8u : 0u);
}
__host__ __device__ __forceinline__ bool gp_dropnode_keep(uint64_t j, uint32_t a, uint64_t seed, uint64_t offset,
                                                          uint32_t thresh) {
    return (gp_dropnode_keep4(j, seed, offset, thresh) >> (a & 3u)) & 1u;
}
