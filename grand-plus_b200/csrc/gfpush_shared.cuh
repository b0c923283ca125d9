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

constexpr int kRankCap = 64;     // the radix select refines until the boundary bucket is this small, then rank-counts it

// `Smem` must provide: sel.hist[kHistBins], sel.bkey[kBucketCap], sel.bid[kBucketCap], warp_scan[CB/32+1], n_out, n_bucket,
// sel_bin, sel_above, sel_inbin.
// The K largest of the items `each` enumerates (each(f) calls f(value > 0, id) for the calling thread's items; it is
// invoked once per radix pass).  emit_fn(position, id, value) receives them in arbitrary order; returns their number.
// MSD radix select on the fp64 bit pattern as in gfpush.cu.  Every thread of the CTA must call it.
template <int CB, class Smem, class Each, class Emit>
__device__ __forceinline__ int block_topk(Smem &sm, const int K, const bool small, Each each, Emit emit_fn) {
    const int tid = threadIdx.x;
    if (tid == 0) { sm.n_out = 0; sm.n_bucket = 0; }
    int want_bucket = K;
    if (small) {
        // at most kBucketCap items in all (uniform): rank-count them directly
        __syncthreads();
        each([&](double x, int id) {
            const int pos = atomicAdd(&sm.n_bucket, 1);
            sm.sel.bkey[pos] = (unsigned long long)__double_as_longlong(x); sm.sel.bid[pos] = id;
        });
    } else {
        for (int i = tid; i < kHistBins; i += CB) sm.sel.hist[i] = 0;
        __syncthreads();
        each([&](double x, int) { atomicAdd(&sm.sel.hist[(unsigned)((unsigned long long)__double_as_longlong(x) >> 52)], 1u); });
        __syncthreads();
        int shift = 52, bits = 11;
        unsigned long long prefix = 0;   // value of key >> (shift + bits) shared by the boundary bucket
        int kk = K;
        bool first = true;
        unsigned long long Tkey = 0;
        for (;;) {
            const unsigned total = select_bin_generic<CB>(sm.sel.hist, sm.warp_scan, 1 << bits, kk, first, &sm.sel_bin,
                                                          &sm.sel_above, &sm.sel_inbin);
            if (first) kk = min(kk, (int)total);   // k = min(K, #positive): graph.h:113 + the v > 0 filter of :121
            if (kk == 0) { want_bucket = 0; Tkey = ~0ull; break; }
            const int bin = sm.sel_bin, above = sm.sel_above, inbin = sm.sel_inbin;
            Tkey = (prefix << bits) | (unsigned long long)bin;
            want_bucket = kk - above;
            if (inbin <= kRankCap || shift == 0) break;
            kk = want_bucket; first = false; prefix = Tkey;
            __syncthreads();
            for (int i = tid; i < kHistBins; i += CB) sm.sel.hist[i] = 0;
            __syncthreads();
            const int nshift = shift >= 11 ? shift - 11 : 0;
            const int nbits = shift >= 11 ? 11 : shift;
            each([&](double x, int) {
                const unsigned long long key = (unsigned long long)__double_as_longlong(x);
                if ((key >> shift) == prefix) atomicAdd(&sm.sel.hist[(unsigned)((key >> nshift) & ((1ull << nbits) - 1ull))], 1u);
            });
            shift = nshift; bits = nbits;
            __syncthreads();
        }
        each([&](double x, int id) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(x);
            const unsigned long long t = key >> shift;
            if (t > Tkey) {
                emit_fn(atomicAdd(&sm.n_out, 1), id, x);
            } else if (t == Tkey) {
                const int pos = atomicAdd(&sm.n_bucket, 1);
                if (pos < kBucketCap) { sm.sel.bkey[pos] = key; sm.sel.bid[pos] = id; }
            }
        });
    }
    __syncthreads();
    {
        // rank-count the boundary bucket: keep its `want_bucket` largest (ties: lower position first)
        const int nb = min(sm.n_bucket, kBucketCap);
        for (int i = tid; i < nb; i += CB) {
            const unsigned long long ki = sm.sel.bkey[i];
            int rank = 0;
            for (int q = 0; q < nb; q++) {
                const unsigned long long kq = sm.sel.bkey[q];
                rank += (kq > ki) || (kq == ki && q < i);
            }
            if (rank < want_bucket) emit_fn(atomicAdd(&sm.n_out, 1), sm.sel.bid[i], __longlong_as_double((long long)ki));
        }
    }
    __syncthreads();
    return sm.n_out;
}

// A LOWER BOUND of the K-th largest of the threads' values (one per thread, `vbits` = bit pattern of a non-negative double
// <= 1, 0 = none), within 1/32 of it: ONE histogram pass over 6 exponent bits + 5 mantissa bits (2^-63 .. 1; smaller
// values share bin 0, which gives no bound) instead of the full radix select -- for thresholds, where any lower bound is
// correct and a tight one only saves work.  Returns 0.0 when fewer than K values are positive.  Every thread of the CTA
// must call it; uses sm.sel.hist, sm.warp_scan, sm.sel_bin / sel_above / sel_inbin.
template <int CB, class Smem>
__device__ __forceinline__ double block_kth_lower_bound(Smem &sm, const int K, const long long vbits) {
    constexpr long long kBinBase = (1023ll - 63ll) << 5;
    const int tid = threadIdx.x;
    for (int i = tid; i < kHistBins; i += CB) sm.sel.hist[i] = 0;
    __syncthreads();
    if (vbits > 0) {
        long long t = (vbits >> 47) - kBinBase;
        t = t < 0 ? 0 : (t > kHistBins - 1 ? kHistBins - 1 : t);
        atomicAdd(&sm.sel.hist[(int)t], 1u);
    }
    __syncthreads();
    const unsigned total = select_bin_generic<CB>(sm.sel.hist, sm.warp_scan, kHistBins, K, false, &sm.sel_bin, &sm.sel_above, &sm.sel_inbin);
    if (total < (unsigned)K) return 0.0;
    const int bin = sm.sel_bin;
    if (bin <= 0) return 0.0;
    return __longlong_as_double(((long long)bin + kBinBase) << 47);   // the lower edge of the K-th largest value's bin
}


}  // namespace gpp
