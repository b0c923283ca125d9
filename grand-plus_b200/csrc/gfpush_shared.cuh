// gfpush_shared.cuh -- pieces shared by the GFPush kernels (gfpush.cu: per-CTA sources; gfpush_cluster.cu: one source
// per thread-block cluster): limits, the edge-owner search of the balanced expansion and the radix-select bin search.
#pragma once
#include "gp_common.cuh"

namespace gpp {

constexpr int kHistBins = 2048;     // 11-bit radix digits
constexpr int kBucketCap = 512;     // boundary bucket resolved in shared memory
constexpr int kMaxK = kBucketCap;   // K above this is refused
constexpr int kMaxLevels = 256;

enum : unsigned long long { kErrOverflow = 1ull, kErrBadSource = 2ull };

// Largest t in [0, BLOCK) with off[t] <= e (off is a non-decreasing exclusive scan, off[0] == 0).
template <int BLOCK>
__device__ __forceinline__ int owner_of_edge(const unsigned *off, unsigned e) {
    int lo = 0;
#pragma unroll
    for (int step = BLOCK / 2; step >= 1; step >>= 1) {
        const int mid = lo + step;
        if (off[mid] <= e) lo = mid;   // mid <= BLOCK-1 always: lo + step never exceeds BLOCK-1
    }
    return lo;
}

// Finds the radix bin holding the kk-th largest among `hist` (bins ordered ascending by key).
// Results in *sel_bin / *sel_above (count in strictly higher bins) / *sel_inbin; returns total.
// Every thread of the CTA must call it (block scan inside).
template <int BLOCK>
__device__ __forceinline__ unsigned select_bin_generic(const unsigned *hist, unsigned *warp_scan, int nbins, int kk,
                                                       bool kk_is_cap, int *sel_bin, int *sel_above, int *sel_inbin) {
    // thread t owns bins [hi - per + 1, hi], hi = nbins-1 - t*per, walking from the top
    const int per = (nbins + BLOCK - 1) / BLOCK;
    const int tid = threadIdx.x;
    unsigned local = 0;
    const int hi = nbins - 1 - tid * per;
#pragma unroll 4
    for (int i = 0; i < per; i++) {
        int b = hi - i;
        if (b >= 0) local += hist[b];
    }
    unsigned total;
    unsigned above = gp_block_exclusive_scan<BLOCK>(local, warp_scan, total);
    unsigned want = kk_is_cap ? min((unsigned)kk, total) : (unsigned)kk;
    if (want > 0 && above < want && want <= above + local) {
        unsigned acc = above;
        for (int i = 0; i < per; i++) {
            int b = hi - i;
            if (b < 0) break;
            unsigned h = hist[b];
            if (acc + h >= want) { *sel_bin = b; *sel_above = (int)acc; *sel_inbin = (int)h; break; }
            acc += h;
        }
    }
    if (want == 0 && tid == 0) { *sel_bin = -1; *sel_above = 0; *sel_inbin = 0; }
    __syncthreads();
    return total;
}

}  // namespace gpp
