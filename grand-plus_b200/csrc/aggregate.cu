// aggregate.cu -- fused gather - mask - scale - reduce Pi.X aggregation on B200 (sm_100a).
//
// Replaces, in ONE kernel and one pass over the feature rows:
//   * the host gather `features[neighbor_idx]` + H2D copy      (/root/reference/model.py:314)
//   * F.dropout on the Pi scores ("DropNode")                   (/root/reference/model.py:82)
//   * feats * scores[:,None]  -- an [nz,F] temporary            (/root/reference/model.py:83)
//   * two torch_scatter.scatter(..., reduce='sum') calls        (/root/reference/model.py:83-86)
//   * the final divide                                          (/root/reference/model.py:87)
// and, with eps = 1e-10 and the embedding table as `table`, MLP.emb (/root/reference/model_mag.py:48-55).
//
// Shape of the work: HBM-bound row gather.  One warp owns one (output row, column tile); it reads
// the row's entry metadata coalesced (one entry per lane), draws the DropNode decisions from
// counter-based Philox BEFORE touching the table so dropped rows cost no bandwidth, then streams
// the kept table rows with 128-bit read-only loads, several rows in flight per lane, accumulating
// in fp32 registers.  Rows are owned, so there are no atomics and no [nz,F] temporary; all n_aug
// augmentations of model.py:321 share one read of every table row.  No tensor cores: this is a
// segmented AXPY, not a dense contraction.
#include "gp_common.cuh"

#include <algorithm>
#include <string>

extern int g_push_bucket, g_push_bucket_block, g_push_bucket_fill, g_push_bucket_nb, g_push_bucket_merge, g_push_cluster, g_push_cluster_probe, g_push_hub_deg, g_push_max_clusters, g_push_smem_hash, g_push_smem_probe, g_push_max_ctas,
    g_push_tuning_gen;  // gfpush.cu

namespace {

constexpr int kAggBlock = 256;  // 8 warps per CTA
constexpr int kMaxAug = 4;

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ void vec_load(float (&x)[VEC], const float *p) {
    if constexpr (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else if constexpr (VEC == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        x[0] = v.x; x[1] = v.y;
    } else {
        x[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void vec_store(float *p, const float (&x)[VEC]) {
    if constexpr (VEC == 4) *reinterpret_cast<float4 *>(p) = make_float4(x[0], x[1], x[2], x[3]);
    else if constexpr (VEC == 2) *reinterpret_cast<float2 *>(p) = make_float2(x[0], x[1]);
    else *p = x[0];
}

struct AggParams {
    const float *table;
    long long ld_table;
    int F;
    const int *row_ptr;
    const int *slot_rows;
    int slot_K;
    const int *nbr;
    const float *score;
    long long B;
    long long n_entries;
    int use_mask;       // training && p > 0
    float scale;        // 1/(1-p)
    unsigned thresh;    // keep iff philox >= thresh
    int drop_all;       // p >= 1
    unsigned long long seed, offset;
    const unsigned char *mask_in;
    unsigned char *mask_out;
    float eps;
    float *out;
    long long ld_out;
    float *denom_out;
    int n_ctile;        // column tiles per row
    int tile_cols;      // columns per tile = NCHUNK*32*VEC
};

// Entry range of output row b.
__device__ __forceinline__ void row_range(const AggParams &P, long long b, long long &j0, long long &j1) {
    if (P.row_ptr) { j0 = P.row_ptr[b]; j1 = P.row_ptr[b + 1]; }
    else {
        const long long r = P.slot_rows ? (long long)P.slot_rows[b] : b;
        if (r < 0) { j0 = j1 = 0; return; }  // not a GFPush source: empty row -> zeros
        j0 = r * P.slot_K; j1 = j0 + P.slot_K;
    }
}

// Keep bits of entry jj for all augmentations (bit a = kept in augmentation a): read from the caller's mask or drawn
// from ONE Philox block per entry.
__device__ __forceinline__ unsigned entry_keep_bits(const AggParams &P, long long jj, int n_aug) {
    if (!P.use_mask) return 0xFu;
    if (P.mask_in) {
        unsigned bits = 0;
        for (int a = 0; a < n_aug; a++) bits |= (P.mask_in[(long long)a * P.n_entries + jj] != 0 ? 1u : 0u) << a;
        return bits;
    }
    return P.drop_all ? 0u : gp_dropnode_keep4((unsigned long long)jj, P.seed, P.offset, P.thresh);
}
// Weight of entry jj in augmentation a (0 when dropped / pad / out of range).
__device__ __forceinline__ float entry_weight(const AggParams &P, unsigned bits, int a, float s, bool &keep) {
    keep = (bits >> a) & 1u;
    if (P.use_mask) return keep ? s * P.scale : 0.0f;
    return s;
}

template <int VEC, int NCHUNK, int NAUG, int UNROLL>
__global__ void __launch_bounds__(kAggBlock) aggregate_fwd_kernel(AggParams P) {
    const int lane = gp_lane();
    // Persistent warps: warp w takes (row, column tile) items w, w + W, ... and loads the entry metadata (neighbour id,
    // score) of its NEXT item before it gathers the rows of the current one, so the metadata round trip of an item is
    // hidden behind the row traffic of the previous item instead of sitting in front of its own.
    const long long n_items = P.B * P.n_ctile;
    const long long n_warps = ((long long)gridDim.x * kAggBlock) >> 5;
    long long item = ((long long)blockIdx.x * kAggBlock + threadIdx.x) >> 5;
    if (item >= n_items) return;  // warp-uniform

    // metadata of the first 32-entry block of an item: entry range, this lane's neighbour id and score
    long long nj0 = 0, nj1 = 0;
    int n_nbr = 0;
    float n_s = 0.0f;
    auto fetch = [&](long long it) {
        row_range(P, it / P.n_ctile, nj0, nj1);
        const long long jj = nj0 + lane;
        n_nbr = 0; n_s = 0.0f;
        if (jj < nj1) {
            n_s = P.score[jj];
            n_nbr = P.nbr ? P.nbr[jj] : (int)jj;
        }
    };
    fetch(item);
    for (; item < n_items; item += n_warps) {
        const long long b = item / P.n_ctile;
        const int ctile = (int)(item - b * P.n_ctile);
        const int col0 = ctile * P.tile_cols + lane * VEC;  // this lane's first column in chunk 0
        const long long j0 = nj0, j1 = nj1;
        int first_nbr = n_nbr;
        float first_s = n_s;
        if (item + n_warps < n_items) fetch(item + n_warps);   // in flight while this item's rows stream

        float acc[NAUG][NCHUNK][VEC];
        float wsum[NAUG];
#pragma unroll
        for (int a = 0; a < NAUG; a++) {
            wsum[a] = 0.0f;
#pragma unroll
            for (int c = 0; c < NCHUNK; c++)
#pragma unroll
                for (int k = 0; k < VEC; k++) acc[a][c][k] = 0.0f;
        }

        for (long long j = j0; j < j1; j += 32) {
            const long long jj = j + lane;
            const bool valid = jj < j1;
            int my_nbr = 0;
            float my_m[NAUG];
            bool any = false;
            {
                float s = 0.0f;
                if (j == j0) { s = first_s; my_nbr = first_nbr; }
                else if (valid) {
                    s = P.score[jj];
                    my_nbr = P.nbr ? P.nbr[jj] : (int)jj;
                }
                const bool live = valid && (P.row_ptr != nullptr || s > 0.0f);  // slot layout: skip zero pads
                const unsigned keep_bits = live ? entry_keep_bits(P, jj, NAUG) : 0u;   // one Philox block for all augmentations
#pragma unroll
                for (int a = 0; a < NAUG; a++) {
                    bool keep = false;
                    my_m[a] = live ? entry_weight(P, keep_bits, a, s, keep) : 0.0f;
                    if (valid && P.mask_out && ctile == 0) P.mask_out[(long long)a * P.n_entries + jj] = (live && keep) ? 1 : 0;
                    any |= (my_m[a] != 0.0f);
                }
            }
            unsigned kept = __ballot_sync(0xffffffffu, any);  // entries some augmentation keeps
            while (kept) {
                int src_lane[UNROLL];
                int nb[UNROLL];
                int cnt = 0;
#pragma unroll
                for (int q = 0; q < UNROLL; q++) {
                    src_lane[q] = 0;
                    if (kept) { src_lane[q] = __ffs(kept) - 1; kept &= kept - 1; cnt = q + 1; }
                    nb[q] = __shfl_sync(0xffffffffu, my_nbr, src_lane[q]);
                }
                float x[UNROLL][NCHUNK][VEC];
#pragma unroll
                for (int q = 0; q < UNROLL; q++) {
                    const float *row = P.table + (long long)nb[q] * P.ld_table;
#pragma unroll
                    for (int c = 0; c < NCHUNK; c++) {
                        const int col = col0 + c * 32 * VEC;
                        if (q < cnt && col < P.F) vec_load<VEC>(x[q][c], row + col);
                        else {
#pragma unroll
                            for (int k = 0; k < VEC; k++) x[q][c][k] = 0.0f;
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < UNROLL; q++) {
#pragma unroll
                    for (int a = 0; a < NAUG; a++) {
                        float m = __shfl_sync(0xffffffffu, my_m[a], src_lane[q]);
                        if (q >= cnt) m = 0.0f;
                        wsum[a] += m;
#pragma unroll
                        for (int c = 0; c < NCHUNK; c++)
#pragma unroll
                            for (int k = 0; k < VEC; k++) acc[a][c][k] = fmaf(m, x[q][c][k], acc[a][c][k]);
                    }
                }
            }
        }
#pragma unroll
        for (int a = 0; a < NAUG; a++) {
            const float den = wsum[a] + P.eps;  // model.py:87 / model_mag.py:54
            float *orow = P.out + ((long long)a * P.B + b) * P.ld_out;
#pragma unroll
            for (int c = 0; c < NCHUNK; c++) {
                const int col = col0 + c * 32 * VEC;
                if (col < P.F) {
                    float y[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; k++) y[k] = acc[a][c][k] / den;
                    vec_store<VEC>(orow + col, y);
                }
            }
            if (P.denom_out && ctile == 0 && lane == 0) P.denom_out[(long long)a * P.B + b] = den;
        }
    }
}


// ---------------------------------------------------------------------------------- TMA-staged forward
// Same math as aggregate_fwd_kernel, but the kept table rows are staged through shared memory by the
// TMA unit: one elected lane issues one `cp.async.bulk` (SASS: UBLKCP) per row tile, completion is
// signalled on a per-slot mbarrier, and every warp runs its own ring of `nbuf` slots.  Bytes in
// flight are then bounded by shared memory (~190 KB per SM) instead of by registers, and address
// generation for a 2.4 KB row costs one instruction instead of 5 x 32 LDG.128.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int NCHUNK, int NAUG>
__global__ void __launch_bounds__(kAggBlock) aggregate_fwd_bulk_kernel(AggParams P, int nbuf, int buf_stride, int warp_stride) {
    extern __shared__ __align__(128) unsigned char agg_smem[];
    const int lane = gp_lane();
    const long long warp = ((long long)blockIdx.x * kAggBlock + threadIdx.x) >> 5;
    const long long b = warp / P.n_ctile;
    if (b >= P.B) return;  // warp-uniform; no CTA-wide barrier is used below
    const int ctile = (int)(warp - b * P.n_ctile);
    unsigned char *my = agg_smem + (size_t)(threadIdx.x >> 5) * warp_stride;
    const unsigned buf0 = smem_u32(my);
    const unsigned bar0 = smem_u32(my + (size_t)nbuf * buf_stride);
    if (lane == 0) {
        for (int i = 0; i < nbuf; i++) mbar_init(bar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int Fpad = (P.F + 3) & ~3;
    const int tile_c0 = ctile * P.tile_cols;
    const int tile_valid = min(P.tile_cols, Fpad - tile_c0);  // columns of this tile, multiple of 4
    const unsigned tile_bytes = (unsigned)tile_valid * 4u;
    const int col_lane = lane * 4;

    long long j0, j1;
    row_range(P, b, j0, j1);

    float acc[NAUG][NCHUNK][4];
    float wsum[NAUG];
#pragma unroll
    for (int a = 0; a < NAUG; a++) {
        wsum[a] = 0.0f;
#pragma unroll
        for (int c = 0; c < NCHUNK; c++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc[a][c][k] = 0.0f;
    }
    int seq_issue = 0, seq_cons = 0;  // running copy numbers of this warp: slot = seq % nbuf, parity = (seq / nbuf) & 1

    for (long long j = j0; j < j1; j += 32) {
        const long long jj = j + lane;
        const bool valid = jj < j1;
        int my_nbr = 0;
        float my_m[NAUG];
        bool any = false;
        {
            float s = 0.0f;
            if (valid) {
                s = P.score[jj];
                my_nbr = P.nbr ? P.nbr[jj] : (int)jj;
            }
            const bool live = valid && (P.row_ptr != nullptr || s > 0.0f);
            const unsigned keep_bits = live ? entry_keep_bits(P, jj, NAUG) : 0u;   // one Philox block for all augmentations
#pragma unroll
            for (int a = 0; a < NAUG; a++) {
                bool keep = false;
                my_m[a] = live ? entry_weight(P, keep_bits, a, s, keep) : 0.0f;
                if (valid && P.mask_out && ctile == 0) P.mask_out[(long long)a * P.n_entries + jj] = (live && keep) ? 1 : 0;
                any |= (my_m[a] != 0.0f);
            }
        }
        const unsigned kept = __ballot_sync(0xffffffffu, any);
        const int n = __popc(kept);
        int issued = 0;
        for (int i = 0; i < n; i++) {
            // keep up to nbuf row copies in flight
            while (issued < n && issued - i < nbuf) {
                const int el = __fns(kept, 0, issued + 1);
                const int nb = __shfl_sync(0xffffffffu, my_nbr, el);
                if (lane == 0) {
                    const int slot = seq_issue % nbuf;
                    mbar_expect_tx(bar0 + 8 * slot, tile_bytes);
                    bulk_g2s(buf0 + slot * buf_stride, P.table + (long long)nb * P.ld_table + tile_c0, tile_bytes,
                             bar0 + 8 * slot);
                }
                seq_issue++; issued++;
            }
            const int el = __fns(kept, 0, i + 1);
            const int slot = seq_cons % nbuf;
            mbar_wait(bar0 + 8 * slot, (unsigned)((seq_cons / nbuf) & 1));
            const float *row = reinterpret_cast<const float *>(my + (size_t)slot * buf_stride);
            float x[NCHUNK][4];
#pragma unroll
            for (int c = 0; c < NCHUNK; c++) {
                const int col = col_lane + c * 128;
                if (col < tile_valid) {
                    const float4 v = *reinterpret_cast<const float4 *>(row + col);
                    x[c][0] = v.x; x[c][1] = v.y; x[c][2] = v.z; x[c][3] = v.w;
                } else { x[c][0] = x[c][1] = x[c][2] = x[c][3] = 0.0f; }
            }
#pragma unroll
            for (int a = 0; a < NAUG; a++) {
                const float m = __shfl_sync(0xffffffffu, my_m[a], el);
                wsum[a] += m;
#pragma unroll
                for (int c = 0; c < NCHUNK; c++)
#pragma unroll
                    for (int k = 0; k < 4; k++) acc[a][c][k] = fmaf(m, x[c][k], acc[a][c][k]);
            }
            seq_cons++;
            __syncwarp();  // every lane has consumed the slot before lane 0 may refill it
        }
    }
    const bool out_vec = ((reinterpret_cast<uintptr_t>(P.out) | (uintptr_t)(P.ld_out * 4)) & 15) == 0;
#pragma unroll
    for (int a = 0; a < NAUG; a++) {
        const float den = wsum[a] + P.eps;
        float *orow = P.out + ((long long)a * P.B + b) * P.ld_out + tile_c0;
#pragma unroll
        for (int c = 0; c < NCHUNK; c++) {
            const int col = col_lane + c * 128;
            if (col < tile_valid) {
                float y[4];
#pragma unroll
                for (int k = 0; k < 4; k++) y[k] = acc[a][c][k] / den;
                if (out_vec && tile_c0 + col + 4 <= P.ld_out) vec_store<4>(orow + col, y);
                else {
#pragma unroll
                    for (int k = 0; k < 4; k++) if (tile_c0 + col + k < P.F) orow[col + k] = y[k];
                }
            }
        }
        if (P.denom_out && ctile == 0 && lane == 0) P.denom_out[(long long)a * P.B + b] = den;
    }
}

// ---------------------------------------------------------------------------------- backward
struct AggBwdParams {
    const float *grad_out;
    long long ld_grad_out;
    const float *denom;
    const int *row_ptr;
    const int *nbr;
    const float *score;
    long long B, n_entries;
    int F;
    int use_mask;
    float scale;
    const unsigned char *mask_in;
    float *grad_table;
    long long ld_grad_table;
    int n_ctile, tile_cols;
};

template <int VEC>
__device__ __forceinline__ void vec_atomic_add(float *p, const float (&x)[VEC]) {
    if constexpr (VEC == 4) atomicAdd(reinterpret_cast<float4 *>(p), make_float4(x[0], x[1], x[2], x[3]));
    else if constexpr (VEC == 2) atomicAdd(reinterpret_cast<float2 *>(p), make_float2(x[0], x[1]));
    else atomicAdd(p, x[0]);
}

template <int VEC, int NCHUNK, int NAUG>
__global__ void __launch_bounds__(kAggBlock) aggregate_bwd_kernel(AggBwdParams P) {
    const int lane = gp_lane();
    const long long warp = ((long long)blockIdx.x * kAggBlock + threadIdx.x) >> 5;
    const long long b = warp / P.n_ctile;
    if (b >= P.B) return;
    const int ctile = (int)(warp - b * P.n_ctile);
    const int col0 = ctile * P.tile_cols + lane * VEC;
    const long long j0 = P.row_ptr[b], j1 = P.row_ptr[b + 1];

    float g[NAUG][NCHUNK][VEC];
#pragma unroll
    for (int a = 0; a < NAUG; a++) {
        const float inv = 1.0f / P.denom[(long long)a * P.B + b];
        const float *grow = P.grad_out + ((long long)a * P.B + b) * P.ld_grad_out;
#pragma unroll
        for (int c = 0; c < NCHUNK; c++) {
            const int col = col0 + c * 32 * VEC;
            if (col < P.F) {
                vec_load<VEC>(g[a][c], grow + col);
#pragma unroll
                for (int k = 0; k < VEC; k++) g[a][c][k] *= inv;
            } else {
#pragma unroll
                for (int k = 0; k < VEC; k++) g[a][c][k] = 0.0f;
            }
        }
    }
    for (long long j = j0; j < j1; j += 32) {
        const long long jj = j + lane;
        const bool valid = jj < j1;
        int my_nbr = 0;
        float my_m[NAUG];
        {
            float s = 0.0f;
            if (valid) { s = P.score[jj]; my_nbr = P.nbr ? P.nbr[jj] : 0; }
#pragma unroll
            for (int a = 0; a < NAUG; a++) {
                float m = s;
                if (P.use_mask) m = (valid && P.mask_in[(long long)a * P.n_entries + jj]) ? s * P.scale : 0.0f;
                my_m[a] = valid ? m : 0.0f;
            }
        }
        const int cnt = (int)min((long long)32, j1 - j);
        for (int t = 0; t < cnt; t++) {
            float m[NAUG];
            bool any = false;
#pragma unroll
            for (int a = 0; a < NAUG; a++) { m[a] = __shfl_sync(0xffffffffu, my_m[a], t); any |= (m[a] != 0.0f); }
            const int nb = __shfl_sync(0xffffffffu, my_nbr, t);
            if (P.nbr && !any) continue;  // accumulate mode: nothing to add
            float *drow = P.grad_table + (P.nbr ? (long long)nb : (j + t)) * P.ld_grad_table;
#pragma unroll
            for (int c = 0; c < NCHUNK; c++) {
                const int col = col0 + c * 32 * VEC;
                if (col < P.F) {
                    float y[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; k++) {
                        float v = 0.0f;
#pragma unroll
                        for (int a = 0; a < NAUG; a++) v = fmaf(m[a], g[a][c][k], v);
                        y[k] = v;
                    }
                    if (P.nbr) vec_atomic_add<VEC>(drow + col, y);
                    else vec_store<VEC>(drow + col, y);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------- small utilities
__global__ void segments_kernel(const long long *idx, long long n, long long B, int *row_ptr, int *flags) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += stride) {
        const long long prev = (j == 0) ? -1 : idx[j - 1];
        const long long cur = (j == n) ? B : idx[j];
        if (j < n && (cur < 0 || cur >= B)) { atomicOr(flags, 1); continue; }
        if (cur < prev) { atomicOr(flags, 1); continue; }
        for (long long r = max(prev, -1ll) + 1; r <= min(cur, B); r++) row_ptr[r] = (int)j;  // rows prev+1..cur start at j
    }
}

__global__ void narrow_kernel(const long long *idx, long long n, long long n_rows, int *out, int *flags) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const long long v = idx[j];
        if (v < 0 || v >= n_rows) atomicOr(flags, 1);
        out[j] = (int)v;
    }
}

__global__ void mask_kernel(long long n, int n_aug, unsigned thresh, int drop_all, unsigned long long seed,
                            unsigned long long offset, unsigned char *mask) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        for (int a = 0; a < n_aug; a++)
            mask[(long long)a * n + j] = (!drop_all && gp_dropnode_keep((unsigned long long)j, (unsigned)a, seed, offset, thresh)) ? 1 : 0;
}

// ---------------------------------------------------------------------------------- dispatch
extern int g_agg_max_vec, g_agg_max_chunk, g_agg_waves;
struct Tiling { int vec, nchunk, n_ctile, tile_cols; };

bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Widest vector the row starts allow, then the fewest column tiles of <= 4 chunks per lane.
Tiling choose_tiling(int F, long long B, const void *p0, long long ld0, const void *p1, long long ld1) {
    int vec = g_agg_max_vec >= 4 ? 4 : g_agg_max_vec >= 2 ? 2 : 1;
    while (vec > 1) {
        const size_t bytes = (size_t)vec * 4;
        if (aligned_to(p0, bytes) && aligned_to(p1, bytes) && ld0 % vec == 0 && ld1 % vec == 0) break;
        vec >>= 1;
    }
    while (vec > 1 && 32 * (vec / 2) >= F) vec >>= 1;  // narrow rows: keep all 32 lanes busy
    const int chunks = (F + 32 * vec - 1) / (32 * vec);
    int maxc = std::max(1, std::min(4, g_agg_max_chunk));
    // small batches (the reference trains with B = 150..250 rows): more column tiles so the launch
    // still covers the 148 SMs with several warps each
    while (maxc > 1 && B * ((chunks + maxc - 1) / maxc) < (long long)GP_NUM_SMS_FALLBACK * 16) maxc--;
    const int n_ctile = (chunks + maxc - 1) / maxc;
    const int nchunk = (chunks + n_ctile - 1) / n_ctile;
    return Tiling{vec, nchunk, n_ctile, nchunk * 32 * vec};
}

template <int VEC, int NCHUNK, int NAUG>
int launch_fwd_t(const AggParams &P, cudaStream_t stream) {
    constexpr int UNROLL = NCHUNK == 1 ? 8 : NCHUNK == 2 ? 4 : 2;
    const long long warps = P.B * P.n_ctile;
    long long blocks = (warps * 32 + kAggBlock - 1) / kAggBlock;
    // persistent warps over the (row, column tile) items: a grid of the resident CTAs ("agg_waves" x occupancy x SMs),
    // every warp walks several items and prefetches the next item's metadata
    static int resident = 0, sms = 0;
    if (resident == 0) {
        int dev = 0;
        GP_CUDA_TRY(cudaGetDevice(&dev));
        GP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        GP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, aggregate_fwd_kernel<VEC, NCHUNK, NAUG, UNROLL>, kAggBlock, 0));
        resident = std::max(resident, 1);
    }
    if (g_agg_waves > 0) blocks = std::min<long long>(blocks, (long long)sms * resident * g_agg_waves);
    aggregate_fwd_kernel<VEC, NCHUNK, NAUG, UNROLL><<<(unsigned)blocks, kAggBlock, 0, stream>>>(P);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

template <int VEC, int NCHUNK>
int launch_fwd_a(const AggParams &P, int n_aug, cudaStream_t s) {
    switch (n_aug) {
        case 1: return launch_fwd_t<VEC, NCHUNK, 1>(P, s);
        case 2: return launch_fwd_t<VEC, NCHUNK, 2>(P, s);
        case 3: return launch_fwd_t<VEC, NCHUNK, 3>(P, s);
        default: return launch_fwd_t<VEC, NCHUNK, 4>(P, s);
    }
}

template <int VEC>
int launch_fwd_c(const AggParams &P, int nchunk, int n_aug, cudaStream_t s) {
    switch (nchunk) {
        case 1: return launch_fwd_a<VEC, 1>(P, n_aug, s);
        case 2: return launch_fwd_a<VEC, 2>(P, n_aug, s);
        case 3: return launch_fwd_a<VEC, 3>(P, n_aug, s);
        default: return launch_fwd_a<VEC, 4>(P, n_aug, s);
    }
}

template <int VEC, int NCHUNK, int NAUG>
int launch_bwd_t(const AggBwdParams &P, cudaStream_t stream) {
    const long long warps = P.B * P.n_ctile;
    const long long blocks = (warps * 32 + kAggBlock - 1) / kAggBlock;
    GP_REQUIRE(blocks < (1ll << 31), "batch too large for one launch");
    aggregate_bwd_kernel<VEC, NCHUNK, NAUG><<<(unsigned)blocks, kAggBlock, 0, stream>>>(P);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

template <int VEC, int NCHUNK>
int launch_bwd_a(const AggBwdParams &P, int n_aug, cudaStream_t s) {
    switch (n_aug) {
        case 1: return launch_bwd_t<VEC, NCHUNK, 1>(P, s);
        case 2: return launch_bwd_t<VEC, NCHUNK, 2>(P, s);
        case 3: return launch_bwd_t<VEC, NCHUNK, 3>(P, s);
        default: return launch_bwd_t<VEC, NCHUNK, 4>(P, s);
    }
}

template <int VEC>
int launch_bwd_c(const AggBwdParams &P, int nchunk, int n_aug, cudaStream_t s) {
    switch (nchunk) {
        case 1: return launch_bwd_a<VEC, 1>(P, n_aug, s);
        case 2: return launch_bwd_a<VEC, 2>(P, n_aug, s);
        case 3: return launch_bwd_a<VEC, 3>(P, n_aug, s);
        default: return launch_bwd_a<VEC, 4>(P, n_aug, s);
    }
}


// ---- tuning knobs (gp_set_tuning): defaults are the measured best, knobs exist for the sweeps in profiles/
// Measured (profiles/r01_aggregate_sweep.txt): the register-staged kernel with 64-bit loads wins or ties
// everywhere -- 128-bit loads cost registers (80 vs 58 -> 24 vs 32 resident warps) and the TMA-staged
// kernel only ties it on 2.4 KB rows and loses badly on 400-byte rows (one lane issues every copy).
int g_agg_kernel = 0;     // 0 auto (= 1), 1 register-staged LDG kernel, 2 TMA-staged bulk kernel
int g_agg_nbuf = 0;       // 0 auto
int g_agg_max_vec = 2;
int g_agg_max_chunk = 4;
int g_agg_smem_kb = 96;   // dynamic shared memory per CTA for the bulk kernel (2 CTAs per SM)
int g_agg_waves = 1;      // "agg_waves": grid of the register-staged kernel = this many resident waves (0 = one warp per item)

template <int NCHUNK, int NAUG>
int launch_bulk_t(AggParams P, int nbuf, int buf_stride, int warp_stride, size_t smem, cudaStream_t stream) {
    const long long warps = P.B * P.n_ctile;
    const long long blocks = (warps * 32 + kAggBlock - 1) / kAggBlock;
    GP_REQUIRE(blocks < (1ll << 31), "batch too large for one launch");
    static size_t configured = 0;
    if (smem > configured) {
        GP_CUDA_TRY(cudaFuncSetAttribute(aggregate_fwd_bulk_kernel<NCHUNK, NAUG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    aggregate_fwd_bulk_kernel<NCHUNK, NAUG><<<(unsigned)blocks, kAggBlock, smem, stream>>>(P, nbuf, buf_stride, warp_stride);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

template <int NCHUNK>
int launch_bulk_a(const AggParams &P, int n_aug, int nbuf, int bs, int ws, size_t smem, cudaStream_t s) {
    switch (n_aug) {
        case 1: return launch_bulk_t<NCHUNK, 1>(P, nbuf, bs, ws, smem, s);
        case 2: return launch_bulk_t<NCHUNK, 2>(P, nbuf, bs, ws, smem, s);
        case 3: return launch_bulk_t<NCHUNK, 3>(P, nbuf, bs, ws, smem, s);
        default: return launch_bulk_t<NCHUNK, 4>(P, nbuf, bs, ws, smem, s);
    }
}

// Returns 1 when the bulk kernel was launched, 0 when it does not apply, negative on error.
int try_launch_bulk(AggParams P, int n_aug, cudaStream_t s) {
    const int Fpad = (P.F + 3) & ~3;
    if (!aligned_to(P.table, 16) || P.ld_table % 4 != 0 || P.ld_table < Fpad) return 0;
    const int chunks = (Fpad + 127) / 128;
    const int n_ctile = (chunks + 7) / 8;
    const int nchunk = (chunks + n_ctile - 1) / n_ctile;
    P.n_ctile = n_ctile; P.tile_cols = nchunk * 128;
    const int tile_bytes = std::min(P.tile_cols, Fpad) * 4;
    const int buf_stride = (tile_bytes + 127) / 128 * 128;
    const int warps_per_cta = kAggBlock / 32;
    const int budget = g_agg_smem_kb * 1024 / warps_per_cta;
    int nbuf = g_agg_nbuf > 0 ? g_agg_nbuf : std::min(16, (budget - 128) / buf_stride);
    if (nbuf < 2) return 0;
    const int warp_stride = (nbuf * buf_stride + nbuf * 8 + 127) / 128 * 128;
    const size_t smem = (size_t)warp_stride * warps_per_cta;
    if (smem > 200 * 1024) return 0;
    int rc;
    switch (nchunk) {
        case 1: rc = launch_bulk_a<1>(P, n_aug, nbuf, buf_stride, warp_stride, smem, s); break;
        case 2: rc = launch_bulk_a<2>(P, n_aug, nbuf, buf_stride, warp_stride, smem, s); break;
        case 3: rc = launch_bulk_a<3>(P, n_aug, nbuf, buf_stride, warp_stride, smem, s); break;
        case 4: rc = launch_bulk_a<4>(P, n_aug, nbuf, buf_stride, warp_stride, smem, s); break;
        case 5: rc = launch_bulk_a<5>(P, n_aug, nbuf, buf_stride, warp_stride, smem, s); break;
        case 6: rc = launch_bulk_a<6>(P, n_aug, nbuf, buf_stride, warp_stride, smem, s); break;
        case 7: rc = launch_bulk_a<7>(P, n_aug, nbuf, buf_stride, warp_stride, smem, s); break;
        default: rc = launch_bulk_a<8>(P, n_aug, nbuf, buf_stride, warp_stride, smem, s); break;
    }
    return rc == GP_OK ? 1 : rc;
}

void dropout_consts(double p, int training, int *use_mask, float *scale, unsigned *thresh, int *drop_all) {
    *use_mask = (training && p > 0.0) ? 1 : 0;
    *drop_all = (p >= 1.0) ? 1 : 0;
    // ATen: noise = bernoulli(1-p) / (1-p) with the python-float p narrowed to fp32 (Dropout.cpp)
    *scale = *drop_all ? 0.0f : 1.0f / (float)(1.0 - p);
    *thresh = gp_keep_threshold((float)p);
}

}  // namespace

extern "C" {

int gp_aggregate_fwd(const gp_aggregate_args *A, void *stream) {
    GpRange nvtx_range("gp_aggregate_fwd");
    GP_REQUIRE(A != nullptr, "args is null");
    GP_REQUIRE(A->B >= 0 && A->n_entries >= 0, "negative size");
    GP_REQUIRE(A->F >= 1, "F must be >= 1");
    GP_REQUIRE(A->n_aug >= 1 && A->n_aug <= kMaxAug, "n_aug must be in 1..%d", kMaxAug);
    GP_REQUIRE(A->p >= 0.0 && A->p <= 1.0, "dropnode rate must be in [0,1]");
    if (A->B == 0) return GP_OK;
    GP_REQUIRE(A->table && A->score && A->out, "null device buffer");
    GP_REQUIRE(A->ld_table >= A->F && A->ld_out >= A->F, "row stride smaller than F");
    GP_REQUIRE(A->row_ptr != nullptr || A->slot_K >= 1, "slot layout needs slot_K >= 1");
    GP_REQUIRE(A->n_entries < (1ll << 31), "entry count exceeds int32 (split the batch)");
    GP_REQUIRE(gp_device_count() > 0, "no CUDA device: this library has no CPU fallback");
    AggParams P{};
    P.table = A->table; P.ld_table = A->ld_table; P.F = A->F;
    P.row_ptr = A->row_ptr; P.slot_rows = A->slot_rows; P.slot_K = A->slot_K;
    P.nbr = A->nbr; P.score = A->score; P.B = A->B; P.n_entries = A->n_entries;
    dropout_consts(A->p, A->training, &P.use_mask, &P.scale, &P.thresh, &P.drop_all);
    P.seed = A->seed; P.offset = A->offset;
    P.mask_in = A->mask_in; P.mask_out = A->mask_out; P.eps = A->eps;
    P.out = A->out; P.ld_out = A->ld_out; P.denom_out = A->denom_out;
    cudaStream_t s = (cudaStream_t)stream;
    if (g_agg_kernel == 2) {
        const int rc = try_launch_bulk(P, A->n_aug, s);
        if (rc != 0) return rc < 0 ? rc : GP_OK;
        GP_REQUIRE(g_agg_kernel != 2, "bulk kernel forced but the table is not 16-byte aligned / padded");
    }
    const Tiling t = choose_tiling(A->F, A->B, A->table, A->ld_table, A->out, A->ld_out);
    P.n_ctile = t.n_ctile; P.tile_cols = t.tile_cols;
    switch (t.vec) {
        case 4: return launch_fwd_c<4>(P, t.nchunk, A->n_aug, s);
        case 2: return launch_fwd_c<2>(P, t.nchunk, A->n_aug, s);
        default: return launch_fwd_c<1>(P, t.nchunk, A->n_aug, s);
    }
}

int gp_aggregate_bwd(const gp_aggregate_bwd_args *A, void *stream) {
    GpRange nvtx_range("gp_aggregate_bwd");
    GP_REQUIRE(A != nullptr, "args is null");
    GP_REQUIRE(A->B >= 0 && A->n_entries >= 0, "negative size");
    GP_REQUIRE(A->F >= 1, "F must be >= 1");
    GP_REQUIRE(A->n_aug >= 1 && A->n_aug <= kMaxAug, "n_aug must be in 1..%d", kMaxAug);
    if (A->B == 0) return GP_OK;
    GP_REQUIRE(A->grad_out && A->denom && A->row_ptr && A->score && A->grad_table, "null device buffer");
    GP_REQUIRE(A->ld_grad_out >= A->F && A->ld_grad_table >= A->F, "row stride smaller than F");
    GP_REQUIRE(gp_device_count() > 0, "no CUDA device: this library has no CPU fallback");
    AggBwdParams P{};
    P.grad_out = A->grad_out; P.ld_grad_out = A->ld_grad_out; P.denom = A->denom;
    P.row_ptr = A->row_ptr; P.nbr = A->nbr; P.score = A->score; P.B = A->B; P.n_entries = A->n_entries; P.F = A->F;
    unsigned thresh; int drop_all;
    dropout_consts(A->p, A->training, &P.use_mask, &P.scale, &thresh, &drop_all);
    GP_REQUIRE(!P.use_mask || A->mask_in, "training-mode backward needs the forward pass's mask");
    P.mask_in = A->mask_in; P.grad_table = A->grad_table; P.ld_grad_table = A->ld_grad_table;
    const Tiling t = choose_tiling(A->F, A->B, A->grad_out, A->ld_grad_out, A->grad_table, A->ld_grad_table);
    P.n_ctile = t.n_ctile; P.tile_cols = t.tile_cols;
    cudaStream_t s = (cudaStream_t)stream;
    switch (t.vec) {
        case 4: return launch_bwd_c<4>(P, t.nchunk, A->n_aug, s);
        case 2: return launch_bwd_c<2>(P, t.nchunk, A->n_aug, s);
        default: return launch_bwd_c<1>(P, t.nchunk, A->n_aug, s);
    }
}

int gp_segments_from_sorted_index(const int64_t *d_idx, int64_t n, int64_t B, int32_t *d_row_ptr, int32_t *d_flags,
                                  void *stream) {
    GP_REQUIRE(n >= 0 && B >= 0 && n < (1ll << 31), "bad sizes");
    GP_REQUIRE(d_row_ptr && d_flags && (d_idx || n == 0), "null device buffer");
    const long long blocks = std::min<long long>((n + 1 + 255) / 256, 148 * 8);
    segments_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const long long *)d_idx, n, B, d_row_ptr, d_flags);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

int gp_narrow_index(const int64_t *d_idx, int64_t n, int64_t n_rows, int32_t *d_out, int32_t *d_flags, void *stream) {
    GP_REQUIRE(n >= 0 && n_rows >= 0 && n_rows < (1ll << 31), "bad sizes");
    if (n == 0) return GP_OK;
    GP_REQUIRE(d_idx && d_out && d_flags, "null device buffer");
    const long long blocks = std::min<long long>((n + 255) / 256, 148 * 8);
    narrow_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const long long *)d_idx, n, n_rows, d_out, d_flags);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

int gp_set_tuning(const char *key, int64_t value) {
    GP_REQUIRE(key != nullptr, "key is null");
    const std::string k(key);
    if (k == "agg_kernel") g_agg_kernel = (int)value;
    else if (k == "agg_nbuf") g_agg_nbuf = (int)value;
    else if (k == "agg_max_vec") g_agg_max_vec = (int)value;
    else if (k == "agg_max_chunk") g_agg_max_chunk = (int)value;
    else if (k == "agg_smem_kb") g_agg_smem_kb = (int)value;
    else if (k == "agg_waves") { GP_REQUIRE(value >= 0, "agg_waves must be >= 0"); g_agg_waves = (int)value; }
    else if (k == "push_cluster") {
        GP_REQUIRE(value == -1 || value == 0 || value == 1 || value == 2 || value == 4 || value == 8 || value == 16,
                   "push_cluster must be 0 (off), 1 (auto), -1 (one CTA per source), 2, 4, 8 or 16");
        g_push_cluster = (int)value; g_push_tuning_gen++;
    }
    else if (k == "push_bucket_merge") { GP_REQUIRE(value == 0 || value == 1, "push_bucket_merge must be 0 (top-k candidates) or 1 (whole reserve)"); g_push_bucket_merge = (int)value; g_push_tuning_gen++; }
    else if (k == "push_bucket_block") { GP_REQUIRE(value == 0 || value == 256 || value == 512 || value == 1024, "push_bucket_block must be 0 (default), 256, 512 or 1024"); g_push_bucket_block = (int)value; g_push_tuning_gen++; }
    else if (k == "push_bucket_fill") { GP_REQUIRE(value >= 3 && value <= 7, "push_bucket_fill must be in 3..7 (eighths of the table)"); g_push_bucket_fill = (int)value; }
    else if (k == "push_bucket_nb") { GP_REQUIRE(value >= 0 && value <= 256, "push_bucket_nb must be in 0..256 (0 = automatic)"); g_push_bucket_nb = (int)value; g_push_tuning_gen++; }
    else if (k == "push_bucket") { GP_REQUIRE(value >= 0 && value <= 2, "push_bucket must be 0 (off), 1 (auto) or 2 (always)"); g_push_bucket = (int)value; g_push_tuning_gen++; }
    else if (k == "push_cluster_probe") { GP_REQUIRE(value >= 1 && value <= 4096, "push_cluster_probe must be in [1, 4096]"); g_push_cluster_probe = (int)value; }
    else if (k == "push_hub_deg") { GP_REQUIRE(value >= 0, "push_hub_deg must be >= 0"); g_push_hub_deg = (int)value; g_push_tuning_gen++; }
    else if (k == "push_max_clusters") { GP_REQUIRE(value >= 0, "push_max_clusters must be >= 0"); g_push_max_clusters = (int)value; g_push_tuning_gen++; }
    else if (k == "push_smem_hash") { g_push_smem_hash = (int)value; g_push_tuning_gen++; }
    else if (k == "push_smem_probe") { GP_REQUIRE(value >= 1 && value <= 1024, "push_smem_probe must be in [1, 1024]"); g_push_smem_probe = (int)value; }
    else if (k == "push_max_ctas") { GP_REQUIRE(value >= 0, "push_max_ctas must be >= 0"); g_push_max_ctas = (int)value; }
    else { gp_set_error("unknown tuning key '%s'", key); return GP_ERR_INVALID; }
    return GP_OK;
}

int gp_dropnode_mask(int64_t n_entries, int32_t n_aug, double p, uint64_t seed, uint64_t offset, uint8_t *d_mask,
                     void *stream) {
    GP_REQUIRE(n_entries >= 0 && n_aug >= 1 && n_aug <= kMaxAug, "bad sizes");
    GP_REQUIRE(p >= 0.0 && p <= 1.0, "dropnode rate must be in [0,1]");
    if (n_entries == 0) return GP_OK;
    GP_REQUIRE(d_mask != nullptr, "null device buffer");
    const long long blocks = std::min<long long>((n_entries + 255) / 256, 148 * 8);
    mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n_entries, n_aug, gp_keep_threshold((float)p),
                                                                   p >= 1.0 ? 1 : 0, seed, offset, d_mask);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

}  // extern "C"
