// gfpush.cu -- GFPush + top-k on B200 (sm_100a).  Replaces Graph::gfpush_omp,
// /root/reference/precompute/graph.h:53-131 (bound at precompute/propagation.cpp:8-12).
//
// Algorithm (what the reference computes, restated in SURVEY.md 3.2 / oracle/gfpush_oracle.c):
//   per source s:  r^0 = e_s;  for i < L-1:  reserve += coef[i] r^i ;
//                  r^{i+1}[v] = sum_{u in N^-1(v), r^i[u] >= rmax deg(u)} r^i[u]/deg(u)  (+ dangling -> s)
//                  reserve += coef[L-1] r^{L-1};  keep the top-K positive reserve entries.
//
// This file: the host side of GFPush (handles, planning, launches, stats, the C ABI) and gfpush_kernel, which is
//   MODE 1 -- the DEFAULT for graphs whose dense residue array fits shared memory (Cora / Citeseer / Pubmed-sized),
//   MODE 0 -- the slab kernel that takes the hand-overs of the first-pass kernels (and everything with push_bucket=0),
//   MODE 2 -- the shared-memory-table tier, the default for table-sized supports until the hash-bucket kernel
//             (gfpush_bucket.cu, two sources per SM) measured 23 - 35 % faster; kept behind push_bucket=0 / push_smem_hash=2.
// Every graph beyond the dense mode runs gfpush_bucket.cu first (plan_bucket below).
//
// B200 design of gfpush_kernel (not the reference's per-thread unordered_maps):
//   * persistent CTAs (SM count x resident CTAs), sources handed out through one global atomic
//     counter -- the schedule(dynamic) of graph.h:73, so hub-heavy sources do not stall a wave;
//   * the next-level residues of a source live ON CHIP whenever its support allows: a dense fp64 array in shared
//     memory for small graphs (MODE 1), a 16 384-slot open-addressed {key, residue} table in the 227 KB of shared
//     memory for supports of that order (MODE 2); what does not fit -- single nodes in MODE 2, everything for
//     supports of 10^5 nodes (MODE 0) -- goes to per-CTA DIRECT-ADDRESSED slabs in HBM (16-byte slots
//     {residue, epoch, reserve position}: a perfect hash node id -> slot, no probing, never reset);
//   * the reserve of on-chip residents is APPENDED as (slot, coef * residue) pairs to a coalesced log and summed by
//     slot into the (then all-zero) residue array after the last level; slab residents keep compact support arrays;
//   * frontiers are compact lists; a node joins the next frontier when its atomic add returns 0.0 (first touch),
//     so tables are cleared by walking lists, never by memset;
//   * edge-balanced expansion: a CTA tile of BLOCK frontier nodes is prefix-summed by degree and
//     the tile's edges are dealt to threads by rank, so a 240K-degree hub is expanded by the whole
//     CTA with coalesced `indices` reads and degree-1 leaves do not idle a warp each;
//   * top-k is an MSD radix select on the fp64 bit pattern (11-bit digits, exponent first),
//     finished by rank-counting the single boundary bucket in shared memory.
//   Residues stay fp64 end to end (the threshold test r >= rmax*deg is a hard comparison).
//   Measured history, rooflines and the designs that lost: DESIGN.md 4.1, profiles/r01_hash_tier.md.
#include "gp_common.cuh"
#include "gfpush_shared.cuh"
#include "gfpush_cluster.h"
#include "gfpush_bucket.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <new>
#include <vector>

// tuning knobs (gp_set_tuning)
int g_push_smem_hash = 1;     // "push_smem_hash": shared-memory residue table in front of the slabs (HBM mode, MODE 2 kernel):
                              // 0 off, 1 auto (on when the expected support is of the order of the table), 2 always on
int g_push_smem_probe = 2;    // "push_smem_probe": 4-key buckets tried before a node is sent to the slab
int g_push_cluster = 0;       // "push_cluster": the cluster kernel (gfpush_cluster.cu) for graphs beyond the dense shared-memory mode:
                              // 0 off (default: measured slower than the per-CTA kernels on every BASELINE shape,
                              // profiles/r02_gfpush.md), 1 auto (cluster size from the expected support), 2/4/8/16 = that
                              // cluster size, -1 = one CTA per source
int g_push_cluster_probe = 128;   // "push_cluster_probe": 4-key buckets tried before a source is handed to the slab kernel
                                  // (at the loads the planner aims at the longest probe sequence is a few buckets)
int g_push_hub_deg = 0;       // "push_hub_deg": entries of at least this degree are expanded by the whole cluster (0 = 64 x G)
int g_push_bucket = 1;        // "push_bucket": the hash-bucket kernel (gfpush_bucket.cu), the default for every graph beyond the dense
                              // shared-memory mode: 0 off (the shared-memory table tier / the slabs run), 1 auto (on unless
                              // push_smem_hash = 2 asks for the table tier), 2 always
int g_push_bucket_nb = 0;     // "push_bucket_nb": buckets per source (rounded up to a power of two); 0 = from the expected support
int g_push_bucket_block = 0;  // "push_bucket_block": threads per CTA of the hash-bucket kernel: 1024 (one CTA per SM, 16 384-slot table), 512 (two per
                              // SM, 8 192 slots each), 256 (three per SM, 4 096 slots); 0 = from the expected support (plan_bucket)
int g_push_bucket_fill = 5;   // "push_bucket_fill": eighths of the table one visit of the bucket kernel may fill with pushed edges (3..7)
int g_push_bucket_merge = 0;  // "push_bucket_merge": 0 = merge only the top-k candidates (the support is not counted), 1 = merge the whole reserve
int g_push_max_clusters = 0;  // "push_max_clusters": cap on the resident clusters (0 = all the device schedules); scaling experiments
int g_push_max_ctas = 0;      // "push_max_ctas": cap on the persistent CTAs of gfpush_kernel (0 = all SMs); scaling experiments
int g_push_tuning_gen = 0;    // bumped by gp_set_tuning so that handles re-plan

using namespace gpp;

namespace {

constexpr int kEdgeUnroll = 4;
constexpr int kSettleUnroll = 4;
constexpr int kSmallList = kHistBins;   // frontier-list entries kept in shared memory (the rest go to HBM)

// One direct-addressed table slot.  The reserve itself is NOT in the table: a slot only remembers
// where the node sits in the compact (sup_id, sup_val) arrays, and an epoch tag (= source number)
// says whether that position belongs to the current source, so nothing is ever reset.
struct __align__(16) Slot {
    double nxt;
    int epoch;
    int pos;
};

struct PushParams {
    const int *indptr;
    const int2 *node_rec;  // [n] {indptr[v], degree}: the pair settle needs, as ONE aligned 8-byte load
    const int *indices;
    const int *packed;     // MODE 2 (nullable): indices with min(degree, cap) in the bits above idbits (gpc_pack_indices)
    int idbits;            // 32 = no degree code
    int n;
    const int *node_idx;
    long long S;
    const double *coef;  // device [L]
    int L;
    double rmax;
    int K;
    int *out_row;
    int *out_col;
    double *out_val;
    float *out_val32;  // nullable
    // per-CTA scratch
    Slot *tab;         // HBM mode: [ctas][n] 16-byte slots {next residue, epoch, support position}
    int epoch_base;    // slot.epoch == epoch_base + it + 1  <=>  node already in source `it`'s reserve
    int *push_start;   // [ctas][capF]  frontier nodes that passed the threshold: CSR offset (-1 = dangling -> source)
    int *push_deg;     // [ctas][capF]
    double *push_val;  // [ctas][capF]  residue r (r/deg is taken when the entry is expanded)
    int *nxt_id;       // [ctas][capF]  ids of the next frontier (first-touch order)
    int *sup_id;       // [ctas][capS]  reserve support: node ids ...
    double *sup_val;   // [ctas][capS]  ... and reserve values, compact (first-touch order)
    long long capF, capS;
    unsigned long long *queue;  // [1] next source
    unsigned long long *stats;  // [0] edges [1] frontier [2] support [3] error flags
    unsigned long long *cum;    // [0] edges [1] frontier [2] support [3] sources; never reset by a call
    // redo mode (second pass after gfpush_cluster_kernel): the queue indexes redo[0 .. *redo_count)
    const int *redo;
    const unsigned long long *redo_count;
    unsigned long long *max_support;  // largest reserve support of any source of this launch (pilot statistics)
    unsigned long long *phase;        // SM cycles per phase, summed over CTAs (see gp_gfpush_phase_cycles)
    // MODE 2 (per-level residue table in shared memory in front of the slabs)
    int *log_id;       // [ctas][capLog]  reserve log: one (node, coef * residue) entry per settled table resident,
    double *log_val;   // [ctas][capLog]  appended coalesced, merged in shared memory before the top-k
    long long capLog;
    int hslots;        // table slots, power of two
    int max_probe;     // a node that finds no slot within this many 4-key buckets lives on the slab for this source
};

template <int BLOCK>
struct PushSmem {
    unsigned off[BLOCK + 1];   // exclusive scan of the tile's degrees, off[BLOCK] = total (sentinel of the owner walk)
    int start[BLOCK];
    double val[BLOCK];
    unsigned warp_scan[BLOCK / 32 + 1];
    double wtau[BLOCK / 32];   // top-k pre-filter: per-warp lower bounds of the K-th largest reserve
    union {
        unsigned hist[kHistBins];   // top-k: radix histogram
        int fl[kHistBins];          // levels: the first kSmallList entries of the frontier list (dead before the top-k)
    };
    unsigned long long bkey[kBucketCap];
    int bid[kBucketCap];
    long long it;
    int n_push, n_nxt, n_sup, n_out, n_bucket;
    int n_list;    // top-k: reserves above the pre-filter threshold, compacted into val[] / start[]
    int n_log;     // MODE 2: entries in the reserve log
    int n_tab;     // MODE 2: distinct nodes in the shared-memory table after the merge
    int table_on;  // MODE 2: this CTA currently uses the table (off when most edges spill: supports far beyond it)
    unsigned spill_edges, all_edges, trial;
    long long ph[8], t_prev, wide_expand, wide_settle;   // thread 0: cycles per phase (gp_gfpush_phase_cycles)
    unsigned wide_E;
    int sel_bin, sel_above, sel_inbin;
};

// Append ids flagged by `is_new` to list[0..cap) through one shared counter per warp; returns the
// position (or -1).  Must be called by all 32 lanes.
__device__ __forceinline__ long long warp_append_pos(bool is_new, long long cap, int *s_count,
                                                     unsigned long long *err) {
    const unsigned m = __ballot_sync(0xffffffffu, is_new);
    if (m == 0) return -1;
    const int lane = gp_lane();
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(s_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!is_new) return -1;
    const long long pos = base + __popc(m & ((1u << lane) - 1u));
    if (pos >= cap) { atomicOr(err, kErrOverflow); return -1; }
    return pos;
}

// The same for U flags per lane with ONE atomic per warp: same-address shared-memory atomics from the 32 warps
// of a CTA serialise, and the append counters were the hottest addresses of both phases.  Positions are assigned
// flag-major (all lanes' flag 0 first).  Must be called by all 32 lanes.
template <int U>
__device__ __forceinline__ void warp_append_multi(const bool (&is_new)[U], long long cap, int *s_count,
                                                  unsigned long long *err, long long (&pos)[U]) {
#pragma unroll
    for (int q = 0; q < U; q++) pos[q] = -1;
    bool any = false;
#pragma unroll
    for (int q = 0; q < U; q++) any |= is_new[q];
    if (!__any_sync(0xffffffffu, any)) return;   // nothing to append in this warp: one vote instead of U ballots
    unsigned m[U];
    int total = 0;
#pragma unroll
    for (int q = 0; q < U; q++) { m[q] = __ballot_sync(0xffffffffu, is_new[q]); total += __popc(m[q]); }
    const int lane = gp_lane();
    int base = 0;
    if (lane == 0) base = atomicAdd(s_count, total);
    base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
    for (int q = 0; q < U; q++) {
        if (is_new[q]) {
            const long long p = base + __popc(m[q] & ((1u << lane) - 1u));
            if (p >= cap) atomicOr(err, kErrOverflow); else pos[q] = p;
        }
        base += __popc(m[q]);
    }
}

template <int BLOCK>
__device__ __forceinline__ unsigned select_bin(PushSmem<BLOCK> &sm, int nbins, int kk, bool kk_is_cap) {
    return select_bin_generic<BLOCK>(sm.hist, sm.warp_scan, nbins, kk, kk_is_cap, &sm.sel_bin, &sm.sel_above, &sm.sel_inbin);
}

template <class Params>
__device__ __forceinline__ void emit(const Params &P, long long it, int src, int slot, int col, double v) {
    long long o = it * P.K + slot;
    P.out_row[o] = src;
    P.out_col[o] = col;
    P.out_val[o] = v;
    if (P.out_val32) P.out_val32[o] = (float)v;
}

// L2 eviction-priority hints (experiment): the slab sectors touched by a level's expand are read again by its
// settle; keep them (evict_last) and let the streams (CSR indices, lists) go first (evict_first / .cs).
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double atom_add_f64_hint(double *p, double x, unsigned long long pol) {
    double old;
    asm volatile("atom.global.add.L2::cache_hint.f64 %0, [%1], %2, %3;" : "=d"(old) : "l"(p), "d"(x), "l"(pol) : "memory");
    return old;
}
__device__ __forceinline__ int4 ldcg_int4_hint(const void *p, unsigned long long pol) {
    int4 r;
    asm volatile("ld.global.cg.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ void st_int4_hint(void *p, int4 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

// Per-CTA view of the direct-addressed slab (MODE 0, and the slab residents of MODE 2).
template <bool UNUSED>
struct Tables {
    Slot *tab;
    unsigned long long pol;  // L2 evict_last policy for the slab sectors
    // next[v] += x; true when v had no residue yet (first touch at this level)
    __device__ __forceinline__ bool add_next(int v, double x) const {
        return atom_add_f64_hint(&tab[v].nxt, x, pol) == 0.0;  // graph.h:98
    }
    // Takes next[v] (the caller clears it with put).  pos >= 0: v already has a reserve entry at that position.
    __device__ __forceinline__ double take(int v, int epoch, int &pos) const {
        // one 16-byte L2 read: the residue half was produced by atomics, so bypass L1
        const int4 raw = ldcg_int4_hint(tab + v, pol);
        pos = (raw.z == epoch) ? raw.w : -1;
        return __hiloint2double(raw.y, raw.x);
    }
    // Clears next[v] and records (epoch, pos): one 16-byte write.
    __device__ __forceinline__ void put(int v, int epoch, int pos, bool changed) const {
        (void)changed;
        st_int4_hint(tab + v, make_int4(0, 0, epoch, pos), pol);
    }
};

// MODE 0: direct-addressed slabs in HBM.
// MODE 1: dense next-residue array double[n] in shared memory (small graphs): MODE 2's machinery with slot == node id
//   (no keys, no probes, no slab residents, a bitmap of reached nodes for the support count).
// MODE 2: an open-addressed {key, next residue} table in SHARED MEMORY in front of the slabs.  A node that finds a
//   slot within max_probe 4-key buckets when the source first touches it lives in the table until the source is done,
//   every other node lives on the slab (both decisions are stable: entries are never removed while a source is live).
//   expand: table residents cost a probe and a shared-memory atomic (measured 0.6 edges/clk/SM incl. probing at a
//           table load of 0.8, profiles/r01_smem_hash_microbench.txt) instead of a DRAM round trip;
//   settle: the frontier list holds ~slot for residents (first touch = the shared-memory add returned 0); a resident's
//           coef * residue is APPENDED to a reserve log as (slot, value) -- coalesced, no read-modify-write in global memory; the only
//           scattered access left per node is its indptr pair;
//   after the last level the log is summed into the (then all-zero) residue array by slot, the top-k reads shared
//   memory, and the slab residents' reserve is in the support arrays exactly as in MODE 0.
//   A CTA whose sources spill most of their edges (supports far beyond the table) switches the table off and
//   re-tries 32 sources later.
template <int BLOCK, int MODE>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) gfpush_kernel(PushParams P) {
    constexpr bool DENSE = MODE == 1;          // MODE 1 = the same machinery as MODE 2 with slot == node id: no keys, no probes,
    constexpr bool SHASH = MODE != 0;          //          no slab residents; a bitmap tells which nodes the source has reached
    __shared__ PushSmem<BLOCK> sm;
    extern __shared__ double s_nxt_dyn[];
    int *s_keys = reinterpret_cast<int *>(s_nxt_dyn + (SHASH ? P.hslots : 0));  // MODE 2: key[hslots]; MODE 1: bitmap[(n+31)/32]
    unsigned *s_seen = reinterpret_cast<unsigned *>(s_keys);
    const unsigned hmask = (unsigned)(P.hslots - 1);
    int *log_id = SHASH ? P.log_id + (long long)blockIdx.x * P.capLog : nullptr;
    double *log_val = SHASH ? P.log_val + (long long)blockIdx.x * P.capLog : nullptr;

    const int tid = threadIdx.x;
    const int lane = gp_lane();
    const long long cta = blockIdx.x;
    // MODE 0 / 2 read the CSR copy whose entries carry the neighbour's degree code (packed entries are non-negative): the
    // table key and the frontier-list entry of a slab resident are the packed entry, and settle only fetches a node's
    // {start, degree} record when the code says it may push (r >= rmax * code)
    const bool has_code = MODE != 1 && P.packed != nullptr && P.idbits < 32;
    const unsigned idmask = has_code ? (1u << P.idbits) - 1u : 0xFFFFFFFFu;
    const int *csr = (MODE != 1 && P.packed != nullptr) ? P.packed : P.indices;
    Tables<false> T;
    T.tab = DENSE ? nullptr : P.tab + cta * (long long)P.n;
    T.pol = l2_policy_evict_last();
    int *push_start = P.push_start + cta * P.capF;
    int *push_deg = P.push_deg + cta * P.capF;
    double *push_val = P.push_val + cta * P.capF;
    int *nxt_id = P.nxt_id + cta * P.capS;   // (allocated with capS entries per CTA)
    int *sup_id = P.sup_id + cta * P.capS;
    double *sup_val = P.sup_val + cta * P.capS;
    unsigned long long *err = P.stats + 3;

    if (SHASH) {
        for (int i = tid; i < P.hslots; i += BLOCK) { s_nxt_dyn[i] = 0.0; if (!DENSE) s_keys[i] = -1; }
        if (DENSE) for (int i = tid; i < (P.hslots + 31) / 32; i += BLOCK) s_seen[i] = 0u;
        if (tid == 0) { sm.table_on = 1; sm.trial = 0; }
    }
    bool table_on = false;   // this level / merge uses the table
    // SHASH: slot of node v in the shared-memory table (claiming one if v is new), or -1 = no slot within max_probe
    auto s_find = [&](int v, bool &claimed) -> int {
        claimed = false;
        if (DENSE) return v;
        if (!table_on) return -1;
        // buckets of four keys (one 16-byte shared-memory read per probe); max_probe buckets are tried
        unsigned b = ((((unsigned)v & idmask) * 2654435761u) >> 9) & (hmask >> 2);
        for (int probe = 0; probe < P.max_probe; probe++, b = (b + 1) & (hmask >> 2)) {
            const int4 k4 = *reinterpret_cast<const int4 *>(s_keys + 4 * b);
            const int kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int k = kk[i];
                if (k == -1) {   // an observed empty is only acted on through the CAS; observed keys are final
                    k = atomicCAS(s_keys + 4 * b + i, -1, v);
                    if (k == -1) { claimed = true; return (int)(4 * b + i); }
                }
                if (k == v) return (int)(4 * b + i);
            }
        }
        return -1;
    };
    unsigned long long st_edges = 0, st_frontier = 0, st_support = 0, st_sources = 0, st_maxsup = 0;  // thread 0 only
    const long long t_begin = clock64();
    if (tid == 0) { for (int i = 0; i < 8; i++) sm.ph[i] = 0; sm.t_prev = t_begin; }
#define GP_PHASE(i) do { if (tid == 0) { const long long t_now = clock64(); sm.ph[i] += t_now - sm.t_prev; sm.t_prev = t_now; } } while (0)

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sm.it = (long long)atomicAdd(P.queue, 1ull);
            sm.n_push = 0; sm.n_nxt = 0; sm.n_sup = 0; sm.n_out = 0; sm.n_bucket = 0; sm.n_tab = 0; sm.n_log = 0;
            sm.spill_edges = 0; sm.all_edges = 0;
        }
        __syncthreads();
        long long it = sm.it;
        if (P.redo) {
            if (it >= (long long)*P.redo_count) break;
            it = P.redo[it];
        } else if (it >= P.S) break;
        const int src = P.node_idx[it];
        if (src < 0 || src >= P.n) {  // refuse instead of reading out of bounds
            if (tid == 0) atomicOr(err, kErrBadSource);
            for (int i = tid; i < P.K; i += BLOCK) emit(P, it, 0, i, 0, 0.0);
            continue;
        }
        const int epoch = P.epoch_base + (int)it + 1;
        int src_key = src;   // the source as a table key / CSR entry
        if (has_code) {
            const unsigned cap = (1u << (31 - P.idbits)) - 1u;
            src_key = (int)((unsigned)src | (min((unsigned)(P.indptr[src + 1] - P.indptr[src]), cap) << P.idbits));
        }
        table_on = SHASH && sm.table_on != 0;   // fixed for the whole source: residency must not change between levels
        // level 0: residue = {src: 1}, reserve = {src: 0} (graph.h:80-81); settle it right away
        if (tid == 0) {
            st_sources++; st_frontier++;
            if (SHASH && sm.table_on) {   // the table is empty: the source claims its home slot; reserve[src] = coef[0]
                bool claimed;
                const int h = s_find(src_key, claimed);
                if (DENSE) s_seen[src >> 5] |= 1u << (src & 31);
                log_id[0] = h; log_val[0] = P.coef[0]; sm.n_log = 1; sm.n_tab = 1;
            } else {
                T.put(src, epoch, 0, true);
                sup_id[0] = src; sup_val[0] = P.coef[0]; sm.n_sup = 1;
            }
            if (P.L > 1) {
                const int a = P.indptr[src], b = P.indptr[src + 1];
                const unsigned d = (unsigned)(b - a);
                // push-list entries [0, BLOCK) live in the tile arrays themselves (off = degree until expand scans it)
                if (d == 0) { sm.start[0] = -1; sm.off[0] = 1; sm.val[0] = 1.0; sm.n_push = 1; }
                else if (1.0 >= P.rmax * (double)d) { sm.start[0] = a; sm.off[0] = d; sm.val[0] = 1.0; sm.n_push = 1; }
            }
        }
        __syncthreads();
        GP_PHASE(0);
        if (tid == 0) { sm.wide_expand = 0; sm.wide_settle = 0; sm.wide_E = 0; }

        for (int level = 0; level < P.L - 1; level++) {  // graph.h:83
            // ---------------------------------------------------------------- expand (graph.h:94-100)
            // Only nodes that passed r >= rmax*deg are in the push list.  A tile of BLOCK of them is
            // prefix-summed by degree and the tile's edges are dealt to threads by rank.
            const int n_push = sm.n_push;
            int n_spill = 0, n_claims = 0;
            unsigned lvl_E = 0;

            for (int base = 0; base < n_push; base += BLOCK) {
                const int j = base + tid;
                unsigned d_push = 0;
                int start = 0;
                double val = 0.0;
                if (j < n_push) {
                    if (base == 0) { d_push = sm.off[tid]; start = sm.start[tid]; val = sm.val[tid]; }   // written by settle
                    else { d_push = (unsigned)push_deg[j]; start = push_start[j]; val = push_val[j]; }
                    // the push list carries the residue; r/deg (graph.h:95) is taken here, one fp64 division per ENTRY with every
                    // lane busy, instead of in settle where a few pushing lanes made whole warps walk the division
                    val = val / (double)d_push;
                }
                unsigned total;
                const unsigned excl = gp_block_exclusive_scan<BLOCK>(d_push, sm.warp_scan, total);
                sm.off[tid] = excl; sm.start[tid] = start; sm.val[tid] = val;
                if (tid == 0) sm.off[BLOCK] = total;
                __syncthreads();
                if (tid == 0) { st_edges += total; lvl_E += total; }
                // edge e of the tile goes to thread e % BLOCK: every warp gets work as soon as the tile has
                // BLOCK edges, and a warp's 32 lanes read 32 consecutive `indices` entries.  (A contiguous share per warp
                // with a walking owner search was measured slower: the walk chains the unrolled edges together.)
                for (unsigned e0 = (unsigned)(tid & ~31); e0 < total; e0 += BLOCK * kEdgeUnroll) {
                    int v[kEdgeUnroll];
                    double add[kEdgeUnroll];
                    bool ok[kEdgeUnroll];
#pragma unroll
                    for (int q = 0; q < kEdgeUnroll; q++) {
                        const unsigned e = e0 + q * BLOCK + lane;
                        ok[q] = e < total;
                        v[q] = src_key; add[q] = 0.0;
                        if (ok[q]) {
                            const int t = owner_of_edge<BLOCK>(sm.off, e);
                            const int st = sm.start[t];
                            add[q] = sm.val[t];
                            if (st >= 0) v[q] = __ldcs(csr + st + (e - sm.off[t]));  // graph.h:96-97 (streaming: evict first)
                        }
                    }
                    bool fresh[kEdgeUnroll];
                    if (SHASH) {
                        int h[kEdgeUnroll];
#pragma unroll
                        for (int q = 0; q < kEdgeUnroll; q++) {
                            h[q] = -1;
                            bool claimed = false;
                            if (ok[q]) h[q] = s_find(v[q], claimed);
                            n_claims += claimed ? 1 : 0;
                        }
#pragma unroll
                        for (int q = 0; q < kEdgeUnroll; q++) {
                            fresh[q] = false;
                            if (ok[q]) {
                                if (h[q] >= 0) { fresh[q] = atomicAdd(s_nxt_dyn + h[q], add[q]) == 0.0; v[q] = ~h[q]; }  // table resident: list entry ~slot
                                else { fresh[q] = T.add_next((int)((unsigned)v[q] & idmask), add[q]); n_spill++; }   // slab resident: list entry = packed v
                            }
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < kEdgeUnroll; q++) fresh[q] = ok[q] && T.add_next((int)((unsigned)v[q] & idmask), add[q]);
                    }
                    long long lpos[kEdgeUnroll];
                    warp_append_multi<kEdgeUnroll>(fresh, P.capS, &sm.n_nxt, err, lpos);
#pragma unroll
                    for (int q = 0; q < kEdgeUnroll; q++)
                        if (lpos[q] >= 0) { if (lpos[q] < kSmallList) sm.fl[lpos[q]] = v[q]; else nxt_id[lpos[q]] = v[q]; }
                }
                __syncthreads();
            }
            if (SHASH) {
                n_spill = (int)__reduce_add_sync(0xffffffffu, (unsigned)n_spill);
                if (lane == 0 && n_spill) atomicAdd(&sm.spill_edges, (unsigned)n_spill);
                n_claims = (int)__reduce_add_sync(0xffffffffu, (unsigned)n_claims);
                if (lane == 0 && n_claims) atomicAdd(&sm.n_tab, n_claims);
                if (tid == 0) sm.all_edges += lvl_E;
            }
            if (n_push == 0) __syncthreads();
            long long t_e = 0;
            if (tid == 0) t_e = clock64() - sm.t_prev;
            GP_PHASE(1);
            // ---------------------------------------------------------------- settle (graph.h:85-93,102)
            // Every node of the new frontier, independently: take its residue, credit the reserve,
            // and decide now whether it will push at the next level.
            const int n_nxt = min((long long)sm.n_nxt, P.capF);
            if (tid == 0 && (long long)sm.n_nxt > P.capF) atomicOr(err, kErrOverflow);   // never silent
            const int next_level = level + 1;
            const bool will_push = next_level < P.L - 1;
            const double c = P.coef[next_level];
            const long long log_base = sm.n_log;   // (SHASH) this level's slice of the reserve log starts here
            if (tid == 0) { st_frontier += n_nxt; sm.n_push = 0; }
            __syncthreads();
            int n_new = 0;   // MODE 1: nodes reached for the first time
            for (int base = 0; base < n_nxt; base += BLOCK * kSettleUnroll) {
                int v[kSettleUnroll];
                bool ok[kSettleUnroll];
                int a[kSettleUnroll], b[kSettleUnroll];
                bool may[kSettleUnroll];
#pragma unroll
                for (int q = 0; q < kSettleUnroll; q++) {
                    const int j = base + q * BLOCK + tid;
                    ok[q] = j < n_nxt;
                    v[q] = ok[q] ? (j < kSmallList ? sm.fl[j] : nxt_id[j]) : 0;
                }
                double x[kSettleUnroll];
                int pos[kSettleUnroll], hs[kSettleUnroll];
                unsigned code[kSettleUnroll];
#pragma unroll
                for (int q = 0; q < kSettleUnroll; q++) {
                    pos[q] = 0; x[q] = 0.0; hs[q] = -1; code[q] = 0u;
                    if (ok[q]) {
                        if (SHASH && v[q] < 0) {   // table resident: residue and key in shared memory
                            hs[q] = ~v[q];
                            x[q] = s_nxt_dyn[hs[q]]; s_nxt_dyn[hs[q]] = 0.0;
                            if (DENSE) {   // slot == node; first time this source reaches it?
                                v[q] = hs[q];
                                const unsigned bit = 1u << (v[q] & 31);
                                if (!(atomicOr(&s_seen[v[q] >> 5], bit) & bit)) n_new++;
                            } else {
                                const unsigned key = (unsigned)s_keys[hs[q]];
                                v[q] = (int)(key & idmask);
                                if (has_code) code[q] = key >> P.idbits;   // min(deg, cap): a lower bound of the degree
                            }
                        } else {
                            if (has_code) code[q] = (unsigned)v[q] >> P.idbits;
                            v[q] = (int)((unsigned)v[q] & idmask);
                            x[q] = T.take(v[q], epoch, pos[q]);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < kSettleUnroll; q++) {
                    a[q] = 0; b[q] = 0;
                    // necessary for graph.h:94: r >= rmax * code (code = 0 when unknown); the exact test on the fetched degree follows
                    may[q] = ok[q] && will_push && x[q] >= P.rmax * (double)code[q];
                    if (may[q]) { const int2 nr = __ldg(P.node_rec + v[q]); a[q] = nr.x; b[q] = nr.x + nr.y; }
                }
                // reserve[v] += coef * r (graph.h:90): logged by slot for table residents (summed after the last level),
                // in the compact support arrays for slab residents
                bool tbl[kSettleUnroll], first[kSettleUnroll], push[kSettleUnroll];
                int st[kSettleUnroll], dg[kSettleUnroll];
                double val[kSettleUnroll];
#pragma unroll
                for (int q = 0; q < kSettleUnroll; q++) {
                    tbl[q] = SHASH && hs[q] >= 0;
                    first[q] = ok[q] && !tbl[q] && pos[q] < 0;
                    push[q] = false; st[q] = -1; dg[q] = 1; val[q] = x[q];
                    if (may[q]) {
                        const unsigned d = (unsigned)(b[q] - a[q]);
                        if (d == 0) push[q] = true;                                   // graph.h:91-93: back to the source
                        else if (x[q] >= P.rmax * (double)d) {                        // graph.h:94
                            push[q] = true; st[q] = a[q]; dg[q] = (int)d;   // val stays the residue: r/deg (graph.h:95) is taken in expand
                        }
                    }
                }
                // The reserve log is written at the frontier-list position (nearly every frontier node is a table resident, so
                // compacting the log would cost a ballot round per batch for nothing): entry log_base + j, -1 = not a resident.
                long long ps[kSettleUnroll], pp[kSettleUnroll], pl[kSettleUnroll];
                if (SHASH) {
#pragma unroll
                    for (int q = 0; q < kSettleUnroll; q++) {
                        pl[q] = ok[q] ? log_base + base + q * BLOCK + tid : -1;
                        if (pl[q] >= P.capLog) { pl[q] = -1; atomicOr(err, kErrOverflow); }
                    }
                }
                warp_append_multi<kSettleUnroll>(first, P.capS, &sm.n_sup, err, ps);
                warp_append_multi<kSettleUnroll>(push, P.capF, &sm.n_push, err, pp);
#pragma unroll
                for (int q = 0; q < kSettleUnroll; q++) {
                    if (SHASH && !tbl[q] && pl[q] >= 0) log_id[pl[q]] = -1;
                    if (tbl[q]) {
                        if (pl[q] >= 0) { log_id[pl[q]] = hs[q]; log_val[pl[q]] = c * x[q]; }
                    } else if (first[q]) {
                        if (ps[q] >= 0) { sup_id[ps[q]] = v[q]; sup_val[ps[q]] = c * x[q]; }
                        T.put(v[q], epoch, (int)max(ps[q], 0ll), true);
                    } else if (ok[q]) {
                        sup_val[pos[q]] += c * x[q];
                        T.put(v[q], epoch, pos[q], false);
                    }
                    if (pp[q] >= 0) {
                        if (pp[q] < BLOCK) { sm.start[pp[q]] = st[q]; sm.off[pp[q]] = (unsigned)dg[q]; sm.val[pp[q]] = val[q]; }
                        else { push_start[pp[q]] = st[q]; push_deg[pp[q]] = dg[q]; push_val[pp[q]] = val[q]; }
                    }
                }
            }
            if (DENSE) {
                n_new = (int)__reduce_add_sync(0xffffffffu, (unsigned)n_new);
                if (lane == 0 && n_new) atomicAdd(&sm.n_tab, n_new);
            }
            __syncthreads();
            if (tid == 0) {
                sm.n_nxt = 0;
                if (SHASH) sm.n_log += n_nxt;
                if (lvl_E >= sm.wide_E) { sm.wide_E = lvl_E; sm.wide_expand = t_e; sm.wide_settle = clock64() - sm.t_prev; }
            }
            GP_PHASE(2);
            __syncthreads();
        }
        if (tid == 0) { sm.ph[5] += sm.wide_expand; sm.ph[6] += sm.wide_settle; }
        for (int i = tid; i < kHistBins; i += BLOCK) sm.hist[i] = 0;
        const bool merged = SHASH && table_on;
        if (merged) {
            // the residue array is all zero after the last settle: sum the reserve log into it, by slot
            __syncthreads();
            const int n_log = min((long long)sm.n_log, P.capLog);
            for (int base = 0; base < n_log; base += BLOCK * 4) {   // four log entries in flight per thread
                int id[4];
                double lv[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int j = base + q * BLOCK + tid;
                    id[q] = j < n_log ? log_id[j] : -1;
                    lv[q] = j < n_log ? log_val[j] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (id[q] >= 0) atomicAdd(s_nxt_dyn + id[q], lv[q]);
            }
            __syncthreads();
            GP_PHASE(3);   // reserve merge
        }
        if (SHASH && !DENSE && tid == 0) {
            // adapt: a CTA whose sources spill most of their edges leaves the table off and tries again later
            if (sm.table_on) {
                if (sm.all_edges > 4096 && 2 * sm.spill_edges > sm.all_edges) { sm.table_on = 0; sm.trial = 0; }
            } else if (++sm.trial >= 32) {
                sm.table_on = 1;
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------ top-k, graph.h:111-126
        const int n_sup = min((long long)sm.n_sup, P.capS);
        if (tid == 0) {
            const unsigned long long sup_all = (unsigned long long)n_sup + (SHASH ? (unsigned)sm.n_tab : 0u);
            st_support += sup_all; st_maxsup = max(st_maxsup, sup_all);
        }
        const int n_tslots = merged ? P.hslots : 0;
        // Pass A -- a threshold below which nothing can be among the K largest: v_r = the r-th largest (distinct) of a
        // warp's 32 lane maxima, r = ceil(K / warps); every warp holds at least r reserves >= its v_r, so at least K
        // reserves are >= tau = min over the warps.  Reads are batched four to a thread (the support of an Amazon2M-shape
        // source is 152 K values in HBM: one load in flight per thread made every pass ~150 dependent round trips).
        double tau = 0.0;
        {
            const int per_warp = (P.K + BLOCK / 32 - 1) / (BLOCK / 32);
            if (per_warp <= 8) {
                long long m1 = 0;   // non-negative doubles order like their bit patterns
                for (int j0 = tid; j0 < n_sup; j0 += 4 * BLOCK) {
                    long long x[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) x[q] = j0 + q * BLOCK < n_sup ? __double_as_longlong(sup_val[j0 + q * BLOCK]) : 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) m1 = max(m1, x[q]);
                }
                for (int j = tid; j < n_tslots; j += BLOCK) m1 = max(m1, __double_as_longlong(s_nxt_dyn[j]));
                long long v = 0x7fffffffffffffffll;
                for (int t = 0; t < per_warp; t++) {   // next distinct lane maximum below v
                    long long w = m1 < v ? m1 : 0;
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
                    v = w;
                }
                if (lane == 0) sm.wtau[tid >> 5] = __longlong_as_double(v);
                __syncthreads();
                tau = sm.wtau[0];
                for (int i = 1; i < BLOCK / 32; i++) tau = fmin(tau, sm.wtau[i]);
            }
        }
        // Pass B -- the survivors (a few dozen) are compacted into shared memory (the tile arrays of the expansion are
        // free now); the table is emptied for the next source on the way.  Every later pass of the select reads the list.
        if (tid == 0) sm.n_list = 0;
        __syncthreads();
        const bool filtered = tau > 0.0;
        auto keep = [&](double x, int id) {
            const int pos = atomicAdd(&sm.n_list, 1);
            if (pos < BLOCK) { sm.val[pos] = x; sm.start[pos] = id; }
        };
        if (filtered) {
            for (int j0 = tid; j0 < n_sup; j0 += 4 * BLOCK) {
                double x[4];
#pragma unroll
                for (int q = 0; q < 4; q++) x[q] = j0 + q * BLOCK < n_sup ? sup_val[j0 + q * BLOCK] : 0.0;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (x[q] >= tau) keep(x[q], sup_id[j0 + q * BLOCK]);
            }
        }
        int n_list = 0;
        bool listed = false;
        if (filtered) {
            for (int j = tid; j < n_tslots; j += BLOCK) {
                const double x = s_nxt_dyn[j];
                if (x >= tau) keep(x, DENSE ? j : (int)((unsigned)s_keys[j] & idmask));
            }
            __syncthreads();
            n_list = sm.n_list;
            listed = n_list <= BLOCK;   // (uniform) more survivors than the list holds: select over the full arrays
        }
        // each(f): f(value > 0, node) for the calling thread's share of the values the select runs over
        auto each = [&](auto f) {
            if (listed) {
                for (int i = tid; i < n_list; i += BLOCK) f(sm.val[i], sm.start[i]);
            } else {
                for (int j = tid; j < n_sup; j += BLOCK) {
                    const double x = sup_val[j];
                    if (x > 0.0) f(x, sup_id[j]);
                }
                for (int j = tid; j < n_tslots; j += BLOCK) {
                    const double x = s_nxt_dyn[j];
                    if (x > 0.0) f(x, DENSE ? j : (int)((unsigned)s_keys[j] & idmask));
                }
            }
        };
        // radix select (11-bit digits, exponent first) down to a boundary bucket of <= kBucketCap, then rank counting
        each([&](double x, int) { atomicAdd(&sm.hist[(unsigned)((unsigned long long)__double_as_longlong(x) >> 52)], 1u); });
        __syncthreads();
        int shift = 52, bits = 11;
        unsigned long long prefix = 0;  // value of key >> (shift+bits) shared by the boundary bucket
        int kk = P.K;
        bool first = true;
        unsigned long long Tkey = 0;
        int want_bucket = 0;
        for (;;) {
            const unsigned total = select_bin<BLOCK>(sm, 1 << bits, kk, first);
            if (first) kk = min(kk, (int)total);  // k = min(K, #positive): graph.h:113 + the v>0 filter of :121
            if (kk == 0) { want_bucket = 0; Tkey = ~0ull; break; }
            const int bin = sm.sel_bin, above = sm.sel_above, inbin = sm.sel_inbin;
            Tkey = (prefix << bits) | (unsigned long long)bin;
            want_bucket = kk - above;
            if (inbin <= kBucketCap || shift == 0) break;
            // refine inside the boundary bucket on the next digit
            kk = want_bucket; first = false; prefix = Tkey;
            __syncthreads();
            for (int i = tid; i < kHistBins; i += BLOCK) sm.hist[i] = 0;
            __syncthreads();
            const int nshift = shift >= 11 ? shift - 11 : 0;
            const int nbits = shift >= 11 ? 11 : shift;
            each([&](double x, int) {
                const unsigned long long key = (unsigned long long)__double_as_longlong(x);
                if ((key >> shift) == prefix) atomicAdd(&sm.hist[(unsigned)((key >> nshift) & ((1ull << nbits) - 1ull))], 1u);
            });
            shift = nshift; bits = nbits;
            __syncthreads();
        }
        // final pass: everything above the boundary bucket is selected; the bucket goes to smem
        each([&](double x, int id) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(x);
            const unsigned long long t = key >> shift;
            if (t > Tkey) {
                emit(P, it, src, atomicAdd(&sm.n_out, 1), id, x);
            } else if (t == Tkey) {
                const int pos = atomicAdd(&sm.n_bucket, 1);
                if (pos < kBucketCap) { sm.bkey[pos] = key; sm.bid[pos] = id; }
            }
        });
        __syncthreads();
        // the table is empty again for the next source
        for (int j = tid; j < n_tslots; j += BLOCK) {
            if (!DENSE) s_keys[j] = -1;
            s_nxt_dyn[j] = 0.0;
        }
        if (DENSE && merged) for (int i = tid; i < (P.hslots + 31) / 32; i += BLOCK) s_seen[i] = 0u;
        {
            // rank-count the boundary bucket: keep its `want_bucket` largest (ties: lower slot first)
            const int nb = min(sm.n_bucket, kBucketCap);
            for (int i = tid; i < nb; i += BLOCK) {
                const unsigned long long ki = sm.bkey[i];
                int rank = 0;
                for (int q = 0; q < nb; q++) {
                    const unsigned long long kq = sm.bkey[q];
                    rank += (kq > ki) || (kq == ki && q < i);
                }
                if (rank < want_bucket)
                    emit(P, it, src, atomicAdd(&sm.n_out, 1), sm.bid[i], __longlong_as_double((long long)ki));
            }
        }
        __syncthreads();
        // unfilled slots read (0, 0, 0.0): what graph.h:117-126 leaves in the caller-zeroed arrays
        for (int i = sm.n_out + tid; i < P.K; i += BLOCK) emit(P, it, 0, i, 0, 0.0);
        GP_PHASE(4);
    }
    if (tid == 0) {
        sm.ph[7] = clock64() - t_begin;
        for (int i = 0; i < 8; i++) atomicAdd(P.phase + i, (unsigned long long)sm.ph[i]);
        atomicAdd(P.stats + 0, st_edges);
        atomicAdd(P.stats + 1, st_frontier);
        atomicAdd(P.stats + 2, st_support);
        atomicMax(P.max_support, st_maxsup);
        atomicAdd(P.cum + 0, st_edges);
        atomicAdd(P.cum + 1, st_frontier);
        atomicAdd(P.cum + 2, st_support);
        atomicAdd(P.cum + 3, st_sources);
    }
}

#undef GP_PHASE
// Table invariant between sources: next residue 0; epoch 0 never matches a source (epochs start at 1).
__global__ void init_tables_kernel(int4 *tab16, int2 *meta8, long long n_slots) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        if (tab16) tab16[i] = make_int4(0, 0, 0, 0);
        else meta8[i] = make_int2(0, 0);
    }
}

__global__ void node_rec_kernel(const int *indptr, long long n, int2 *rec) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        rec[i] = make_int2(indptr[i], indptr[i + 1] - indptr[i]);
}

// CSR sanity: indptr[0]==0, non-decreasing, indptr[n]==nnz, 0 <= indices < n.
__global__ void validate_csr_kernel(const int *indptr, long long n, const int *indices, long long nnz, int *flag) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    for (long long i = i0; i < n; i += stride) bad |= (indptr[i + 1] < indptr[i]);
    for (long long i = i0; i < nnz; i += stride) bad |= (indices[i] < 0 || indices[i] >= n);
    if (i0 == 0) bad |= (indptr[0] != 0) | ((long long)indptr[n] != nnz);
    if (bad) atomicOr(flag, 1);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
struct gp_graph {
    int device = 0;
    long long n = 0, nnz = 0;
    int *d_indptr = nullptr;
    int *d_indices = nullptr;
    int2 *d_node_rec = nullptr;
    bool owns_csr = false;
    int num_sms = GP_NUM_SMS_FALLBACK;
    size_t smem_optin = 0;
    gp_push_config cfg{};
    cudaStream_t stream = nullptr;
    // scratch (lazily sized for the most demanding call so far)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    size_t budget_cached = 0;   // half of the free device memory when it was last asked for (make_plan)
    long long scratch_ctas = 0, scratch_capF = 0, scratch_capS = 0;
    int scratch_mode = 0;
    int scratch_hslots = 0;
    long long scratch_capLog = 0;
    long long epoch_base = 0;              // sources pushed since the tables were last initialised
    double *d_coef = nullptr;              // [kMaxLevels]
    // [0] queue [1..3] stats [4] error flags (sticky until read) [6] redo count [7] slab-kernel queue [8] max support
    // | [16..21] cumulative | [24..31] phase cycles
    unsigned long long *d_ctrl = nullptr;
    // cluster kernel (gfpush_cluster.cu): CSR entries with the degree code, per-CTA / per-cluster scratch
    int *d_packed = nullptr;
    int idbits = 32;
    void *cscratch = nullptr;
    size_t cscratch_bytes = 0;
    void *bscratch = nullptr;   // bucket kernel: per-CTA pair streams, reserve logs, push list, merged reserve
    size_t bscratch_bytes = 0;
    // largest support of a source measured on this handle, and the (L, rmax) it was measured for: sizes the hash buckets
    long long support_hint = 0;
    int hint_L = 0;
    double hint_rmax = 0.0;
    int cur_L = 0;              // parameters of the last push (collect_stats files its measurement under them)
    double cur_rmax = 0.0;
    int max_clusters[5] = {-1, -1, -1, -1, -1};   // resident clusters of 1, 2, 4, 8, 16 CTAs (-1 = not asked yet)
    // orders a push after the previous one on this handle when the two run on different streams
    cudaEvent_t ev_done = nullptr;
    bool ev_recorded = false;
    int *d_redo = nullptr;
    size_t d_redo_cap = 0;
    // staging for the host-buffer entry point
    int *d_node = nullptr;
    size_t d_node_cap = 0;
    void *d_out = nullptr;
    size_t d_out_cap = 0;
    gp_push_stats last{};
    std::mutex mu;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <int BLOCK>
size_t static_smem_bytes() { return sizeof(PushSmem<BLOCK>); }

struct Plan {
    int block;
    int mode;  // GP_SCRATCH_SMEM / GP_SCRATCH_HBM
    long long ctas, capF, capS;
    size_t dyn_smem;
    size_t bytes;
    size_t off_tab, off_push_start, off_push_deg, off_push_val, off_nxt_id, off_sup_id, off_cand, off_log_id, off_log_val;
    long long capLog;
    int hslots;  // > 0: MODE 2 (shared-memory hash in front of the slabs)
};

int make_plan(gp_graph *g, long long S, int L, double rmax, bool redo_only, Plan *pl) {
    const long long n = g->n;
    // Measured on B200 (profiles/r01_block_sweep.md): one 1024-thread CTA per SM wins whenever a source
    // carries thousands of frontier nodes (fewer concurrent sources -> their table lines stay in L2);
    // tiny graphs (Cora-sized) prefer four 256-thread CTAs per SM because their levels are sync-bound.
    int block = g->cfg.block_threads ? g->cfg.block_threads : (n <= 4096 ? 256 : 1024);
    GP_REQUIRE(block == 256 || block == 512 || block == 1024, "block_threads must be 256, 512 or 1024 (got %d)", block);
    const size_t stat = block == 256 ? static_smem_bytes<256>() : block == 512 ? static_smem_bytes<512>() : static_smem_bytes<1024>();
    const size_t smem_need = stat + (size_t)n * sizeof(double) + ((size_t)n + 31) / 32 * 4 + 1024;
    int mode = g->cfg.scratch_mode;
    if (mode == GP_SCRATCH_AUTO) mode = (smem_need <= g->smem_optin) ? GP_SCRATCH_SMEM : GP_SCRATCH_HBM;
    GP_REQUIRE(mode == GP_SCRATCH_SMEM || mode == GP_SCRATCH_HBM, "unknown scratch_mode %d", mode);
    GP_REQUIRE(mode != GP_SCRATCH_SMEM || smem_need <= g->smem_optin,
               "GP_SCRATCH_SMEM needs %zu B of shared memory, device offers %zu", smem_need, g->smem_optin);
    // SURVEY 7 work bound: each pushed edge carries >= rmax of a level's <= 1 total mass, so a
    // level creates at most 1/rmax frontier entries (+1 for the dangling return to the source).
    long long capF = n;
    if (rmax > 0.0) {
        double b = std::ceil(1.0 / rmax * 1.0001) + 16.0;
        if (b < (double)n) capF = (long long)b;
    }
    long long capS = n;
    {
        double b = 1.0 + (double)std::max(L - 1, 0) * (double)capF + 16.0;
        if (b < (double)n) capS = (long long)b;
    }
    capF = std::max<long long>(capF, 1);
    capS = std::max<long long>(capS, 1);
    int per_sm = g->cfg.ctas_per_sm;
    if (per_sm <= 0) per_sm = std::max(1, 2048 / block / 2);  // half the thread slots: 2 x 512
    if (mode == GP_SCRATCH_SMEM) {
        const int fit = std::max(1, (int)(((size_t)228 * 1024) / smem_need));  // 228 KB of smem per SM
        per_sm = std::min(per_sm, fit);
    }
    long long ctas = (long long)g->num_sms * per_sm;  // scratch is sized for a full grid; small calls launch fewer
    // behind the cluster kernel the slabs only take the few sources it hands over: a quarter of the SMs, no table
    if (redo_only) ctas = std::max<long long>(1, ctas / 4);
    // MODE 2: the largest power-of-two table {int key, double residue} that fits beside the static shared memory
    int hslots = 0;
    // The table pays when a source's support is of the order of the table (measured on B200, profiles/r01_hash_tier.md);
    // supports of the BASELINE shapes are 0.13 - 0.19 / rmax (Reddit 12.9 K @1e-5, MAG 18.7 K @1e-5, Amazon2M 152 K
    // @1e-6), so "auto" (1) keeps the plain slabs when 0.15 / rmax is beyond twice the largest table; 2 forces it on.
    if (mode == GP_SCRATCH_SMEM) hslots = (int)n;   // MODE 1: slot == node id
    bool want_table = mode == GP_SCRATCH_HBM && g_push_smem_hash != 0 && !redo_only;
    if (want_table && g_push_smem_hash == 1 && rmax > 0.0 && std::min(0.15 / rmax, (double)n) > 2.0 * 16384.0) want_table = false;
    if (want_table && g_push_smem_hash == 1 && rmax <= 0.0 && n > 2 * 16384) want_table = false;
    if (want_table) {
        const size_t avail = g->smem_optin / (size_t)per_sm;
        if (avail > stat + 2048) {
            hslots = 1;
            while ((size_t)hslots * 2 * 12 + stat + 1024 <= avail) hslots *= 2;
            if (hslots < 1024) hslots = 0;
        }
    }
    // reserve log of MODE 2: at most one entry per table slot per level
    // (one entry per frontier node and level, at the node's frontier-list position)
    const long long capLog = hslots ? 1 + (long long)std::max(L - 1, 0) * capF + 8192 : 0;
    auto bytes_for = [&](long long c, Plan *p) {
        size_t o = 0;
        p->off_tab = o; o += align_up((size_t)c * n * (mode == GP_SCRATCH_HBM ? 16 : 0), 256);
        p->off_push_start = o; o += align_up((size_t)c * capF * 4, 256);
        p->off_push_deg = o; o += align_up((size_t)c * capF * 4, 256);
        p->off_push_val = o; o += align_up((size_t)c * capF * 8, 256);
        p->off_nxt_id = o; o += align_up((size_t)c * capS * 4, 256);
        p->off_sup_id = o; o += align_up((size_t)c * capS * 4, 256);
        p->off_cand = o; o += align_up((size_t)c * capS * 8, 256);
        p->off_log_id = o; o += align_up((size_t)c * capLog * 4, 256);
        p->off_log_val = o; o += align_up((size_t)c * capLog * 8, 256);
        return o;
    };
    size_t budget = (size_t)g->cfg.max_scratch_bytes;
    Plan tmp{};
    if (budget == 0) {
        // cudaMemGetInfo takes the driver's allocation lock (0.2 - 6 ms per call measured inside a running pipeline): only
        // ask when the scratch the handle already holds does not cover this call
        if (g->scratch && bytes_for(ctas, &tmp) <= g->scratch_bytes) budget = g->scratch_bytes;
        else if (g->budget_cached) budget = g->budget_cached;   // (asked once per allocation: ensure_scratch invalidates it)
        else {
            size_t free_b = 0, total_b = 0;
            GP_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            budget = (free_b + g->scratch_bytes) / 2;
            g->budget_cached = budget;
        }
    }
    while (ctas > 1 && bytes_for(ctas, &tmp) > budget) ctas = std::max<long long>(1, ctas * 3 / 4);
    pl->block = block; pl->mode = mode; pl->ctas = ctas; pl->capF = capF; pl->capS = capS;
    pl->hslots = hslots; pl->capLog = capLog;
    pl->dyn_smem = mode == GP_SCRATCH_SMEM ? (size_t)n * sizeof(double) + ((size_t)n + 31) / 32 * 4 : (size_t)hslots * 12;
    pl->bytes = bytes_for(ctas, pl);
    GP_REQUIRE(pl->bytes <= budget || ctas == 1, "scratch does not fit the budget");
    return GP_OK;
}

int ensure_scratch(gp_graph *g, const Plan &pl, long long S, cudaStream_t stream) {
    const bool same = g->scratch && g->scratch_bytes >= pl.bytes && g->scratch_ctas == pl.ctas &&
                      g->scratch_capF == pl.capF && g->scratch_capS == pl.capS && g->scratch_mode == pl.mode &&
                      g->scratch_hslots == pl.hslots && g->scratch_capLog == pl.capLog &&
                      g->epoch_base + S < (1ll << 31) - 2;  // epoch tags are int32: re-initialise before they wrap
    if (same) return GP_OK;
    if (g->scratch && g->scratch_bytes < pl.bytes) {
        GP_CUDA_TRY(cudaStreamSynchronize(stream));
        GP_CUDA_TRY(cudaFree(g->scratch));
        g->scratch = nullptr; g->scratch_bytes = 0;
    }
    if (!g->scratch) {
        GP_CUDA_TRY(cudaMalloc(&g->scratch, pl.bytes));
        g->scratch_bytes = pl.bytes;
        g->budget_cached = 0;   // the next plan that needs a budget asks the driver again
    }
    // table invariants between sources: next residue == 0 everywhere, reserve == kUnseen everywhere
    char *base = (char *)g->scratch;
    if (pl.mode == GP_SCRATCH_HBM) {   // MODE 1 keeps no per-node state in HBM
        init_tables_kernel<<<g->num_sms * 8, 256, 0, stream>>>((int4 *)(base + pl.off_tab), nullptr, pl.ctas * g->n);
        GP_CUDA_TRY(cudaGetLastError());
    }
    g->epoch_base = 0;
    g->scratch_hslots = pl.hslots; g->scratch_capLog = pl.capLog;
    g->scratch_ctas = pl.ctas; g->scratch_capF = pl.capF; g->scratch_capS = pl.capS; g->scratch_mode = pl.mode;
    return GP_OK;
}

template <int BLOCK>
int launch_push(const PushParams &P, const Plan &pl, cudaStream_t stream) {
    unsigned grid = (unsigned)std::min<long long>(pl.ctas, P.S);
    if (g_push_max_ctas > 0) grid = std::min(grid, (unsigned)g_push_max_ctas);
    if (pl.mode == GP_SCRATCH_SMEM) {
        GP_CUDA_TRY(cudaFuncSetAttribute(gfpush_kernel<BLOCK, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)pl.dyn_smem));
        GP_CUDA_TRY(cudaFuncSetAttribute(gfpush_kernel<BLOCK, 1>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        gfpush_kernel<BLOCK, 1><<<grid, BLOCK, pl.dyn_smem, stream>>>(P);
    } else if (pl.hslots > 0) {
        GP_CUDA_TRY(cudaFuncSetAttribute(gfpush_kernel<BLOCK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)pl.dyn_smem));
        GP_CUDA_TRY(cudaFuncSetAttribute(gfpush_kernel<BLOCK, 2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        gfpush_kernel<BLOCK, 2><<<grid, BLOCK, pl.dyn_smem, stream>>>(P);
    } else {
        gfpush_kernel<BLOCK, 0><<<grid, BLOCK, 0, stream>>>(P);
    }
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

int launch_slab(const PushParams &P, const Plan &pl, cudaStream_t stream) {
    return pl.block == 256 ? launch_push<256>(P, pl, stream)
           : pl.block == 512 ? launch_push<512>(P, pl, stream)
                             : launch_push<1024>(P, pl, stream);
}

// ---- cluster kernel: planning ------------------------------------------------------------------------
struct ClusterPlan {
    int G = 0, clusters = 0, hub_min_deg = 0;
    long long capP = 0, capX = 0;
    int capHub = 0;
    size_t bytes = 0;
    size_t off_ps = 0, off_pl = 0, off_pa = 0, off_xi = 0, off_xv = 0, off_ci = 0, off_cv = 0, off_hs = 0, off_hd = 0, off_ha = 0;
};

int cluster_slot_of(int G) { return G == 1 ? 0 : G == 2 ? 1 : G == 4 ? 2 : G == 8 ? 3 : 4; }

// Chooses the cluster size for this call (0 = use the per-CTA kernels) and sizes its scratch.
int plan_cluster(gp_graph *g, long long S, int L, double rmax, int K, long long capF, ClusterPlan *cp) {
    *cp = ClusterPlan{};
    if (g_push_cluster == 0) return GP_OK;
    const long long n = g->n;
    // supports of the BASELINE shapes are 0.13 - 0.19 / rmax (Reddit 13.2 K @1e-5, MAG 18.8 K @1e-5, Amazon2M 152 K @1e-6);
    // a CTA's 16 384-slot table is kept below a load of ~0.6 for the typical source, outliers go to the slab kernel
    const double est = rmax > 0.0 ? std::min(0.15 / rmax, (double)n) : (double)n;
    int G = 0;
    if (g_push_cluster == 1) {
        for (int c = 1; c <= kClusterMaxG; c *= 2)
            if (est / c <= 0.6 * kClusterSlots) { G = c; break; }
        if (G == 0) return GP_OK;   // beyond 16 CTAs: the slabs take it
    } else {
        G = g_push_cluster < 0 ? 1 : g_push_cluster;
    }
    int &mc = g->max_clusters[cluster_slot_of(G)];
    if (mc < 0) {
        int rc = gpc_max_clusters(G, g->num_sms, &mc);
        if (rc != GP_OK) return rc;
    }
    if (mc <= 0) {
        GP_REQUIRE(g_push_cluster == 1, "clusters of %d CTAs cannot be scheduled on this device", G);
        return GP_OK;
    }
    int clusters = mc;
    if (g_push_max_clusters > 0) clusters = std::min(clusters, g_push_max_clusters);
    clusters = (int)std::min<long long>(clusters, S);
    cp->G = G; cp->clusters = clusters;
    cp->hub_min_deg = g_push_hub_deg > 0 ? g_push_hub_deg : 64 * G;
    cp->capP = capF;
    // a level pushes at most min(nnz + n, 1/rmax) edges (every pushed edge carries >= rmax of a level's <= 1 total mass; a
    // dangling node returns its residue over one pseudo-edge); hashed ownership spreads them evenly over the G x G streams
    long long capE = g->nnz + n;
    if (rmax > 0.0) capE = (long long)std::min<double>((double)capE, std::ceil(1.0 / rmax * 1.0001) + 16.0);
    cp->capX = G == 1 ? 1 : std::min<long long>(capE, std::max<long long>(4096, 8 * capE / ((long long)G * G)));
    // a pushing node of degree d carries r >= rmax * d, and a level's residues sum to <= 1
    double hub = rmax > 0.0 ? std::ceil(1.0 / (rmax * cp->hub_min_deg)) + 16.0 : (double)n;
    cp->capHub = G == 1 ? 1 : (int)std::min<double>(hub, (double)std::min<long long>(n, 1ll << 24));
    const size_t ctas = (size_t)clusters * G;
    size_t o = 0;
    cp->off_ps = o; o += align_up(ctas * cp->capP * 4, 256);
    cp->off_pl = o; o += align_up(ctas * cp->capP * 4, 256);
    cp->off_pa = o; o += align_up(ctas * cp->capP * 8, 256);
    cp->off_xi = o; o += align_up(ctas * 2 * G * cp->capX * 4, 256);   // two generations (level parity)
    cp->off_xv = o; o += align_up(ctas * 2 * G * cp->capX * 8, 256);
    cp->off_ci = o; o += align_up(ctas * K * 4, 256);
    cp->off_cv = o; o += align_up(ctas * K * 8, 256);
    cp->off_hs = o; o += align_up((size_t)clusters * cp->capHub * 4, 256);
    cp->off_hd = o; o += align_up((size_t)clusters * cp->capHub * 4, 256);
    cp->off_ha = o; o += align_up((size_t)clusters * cp->capHub * 8, 256);
    cp->bytes = o;
    return GP_OK;
}

// ---- bucket kernel: planning -------------------------------------------------------------------------
constexpr long long kPilotMinSources = 2048;
constexpr int kBucketDefaultBlock = 1024;   // (measured: profiles/r02_gfpush.md 5)   // calls with fewer sources are not worth the pilot's synchronisation
struct BucketPlan {
    int nb = 0, log_nb = 0, block = 0;
    long long ctas = 0, capPair = 0, capLog = 0, capP = 0, capS = 0;
    size_t bytes = 0;
    size_t off_pi = 0, off_pv = 0, off_li = 0, off_lv = 0, off_ps = 0, off_pl = 0, off_pa = 0, off_si = 0, off_sv = 0;
};

// The bucket kernel replaces the slab kernel as the first pass where the slabs would run (supports far beyond the
// shared-memory table) and the bucket counters fit shared memory; the slabs then only take what it hands over.
int plan_bucket(gp_graph *g, long long S, int L, double rmax, const Plan &pl, long long support_hint, BucketPlan *bp) {
    *bp = BucketPlan{};
    if (g_push_bucket == 0) return GP_OK;
    if (g_push_bucket == 1 && pl.hslots > 0 && g_push_smem_hash == 2) return GP_OK;   // the shared-memory table tier was asked for
    const long long n = g->n;
    // Geometry (measured, profiles/r02_gfpush.md 6): supports of the order of one table (the regime the shared-memory table tier
    // was built for: Reddit-, MAG-shape) run two sources per SM in 512-thread CTAs with 8 192-slot tables -- the barriers and
    // dependent shared-memory round trips of one source overlap the other's; supports far beyond the table (Amazon2M-shape)
    // keep one 1 024-thread CTA per SM with the 16 384-slot table (half the bucket visits per level).
    const double est_support = rmax > 0.0 ? std::min(0.15 / rmax, (double)n) : (double)n;
    // (tiny supports -- rmax 1e-3 on the Reddit-shape graph: a few hundred nodes -- are pure latency chains: three sources per SM in
    // 256-thread CTAs measured 21.4 M rows/s against 16.2 M at 512 threads and 6.8 M for the shared-memory-table kernel)
    const int block = g_push_bucket_block ? g_push_bucket_block : est_support <= 512.0 ? 256 : est_support <= 2.0 * 16384.0 ? 512 : 1024;
    const long long slots = gpb_slots(block);
    // a level pushes at most min(nnz + n, 1/rmax) edges
    long long capE = g->nnz + n;
    if (rmax > 0.0) capE = (long long)std::min<double>((double)capE, std::ceil(1.0 / rmax * 1.0001) + 16.0);
    // buckets: a power of two >= 2 such that the LARGEST support seen on this handle for these parameters (a pilot launch
    // measures it, push_device_locked) loads the table to <= 0.9; without a measurement the a-priori bound min(n, 1.5 x the
    // per-level edge bound) stands in, which is several times too large on the BASELINE shapes: more and smaller visits
    // than necessary, but never an overflow.  A source that outgrows the table anyway is handed to the slab kernel.
    long long nb = 2;
    if (g_push_bucket_nb > 0) {
        while (nb < g_push_bucket_nb && nb < kBucketMaxBuckets) nb *= 2;
    } else {
        // (the probe sequence only runs out near a load of 1: the largest support seen may load the table to 0.9)
        if (support_hint > 0) while (nb < kBucketMaxBuckets && support_hint > nb * (slots * 9ll / 10)) nb *= 2;
        else while (nb < kBucketMaxBuckets && std::min<long long>(n, capE + capE / 2) > nb * (slots * 4ll / 5)) nb *= 2;
    }
    int log_nb = 0;
    while ((1ll << log_nb) < nb) log_nb++;
    bp->nb = (int)nb; bp->log_nb = log_nb; bp->block = block;
    int per_sm = 1;
    int rc = gpb_ctas_per_sm((int)nb, block, &per_sm);
    if (rc != GP_OK) return rc;
    bp->ctas = (long long)g->num_sms * per_sm;
    // a bucket stream holds four times its even share of a level (and at least 4096 pairs)
    bp->capPair = std::min<long long>(capE, std::max<long long>(4096, 4 * capE / nb));
    // the reserve log holds one entry per frontier node and level: at most 1 + (L-1) * capF in all
    bp->capLog = std::min<long long>(1 + (long long)std::max(L - 1, 0) * pl.capF + 16, std::max<long long>(4096, 4 * capE / nb));
    bp->capP = pl.capF * 2 + 4096;   // push entries are cut into 128-edge chunks: at most capF nodes + capE / 128 chunks
    // the merged reserve: one slot per node of the support
    bp->capS = std::min<long long>(n, 1 + (long long)std::max(L - 1, 0) * pl.capF) + 16;
    // (the kernel addresses a CTA's streams with 32-bit offsets)
    if (nb * bp->capPair >= (1ll << 31) || nb * bp->capLog >= (1ll << 31)) { *bp = BucketPlan{}; return GP_OK; }
    const size_t c = (size_t)bp->ctas;
    size_t o = 0;
    bp->off_pi = o; o += align_up(c * nb * bp->capPair * 4, 256);
    bp->off_pv = o; o += align_up(c * nb * bp->capPair * 8, 256);
    bp->off_li = o; o += align_up(c * nb * bp->capLog * 4, 256);
    bp->off_lv = o; o += align_up(c * nb * bp->capLog * 8, 256);
    bp->off_ps = o; o += align_up(c * bp->capP * 4, 256);
    bp->off_pl = o; o += align_up(c * bp->capP * 4, 256);
    bp->off_pa = o; o += align_up(c * bp->capP * 8, 256);
    bp->off_si = o; o += align_up(c * bp->capS * 4, 256);
    bp->off_sv = o; o += align_up(c * bp->capS * 8, 256);
    bp->bytes = o;
    return GP_OK;
}

// CSR entries with the degree code in the spare high bits (built once per handle, on first use).
int ensure_packed(gp_graph *g, cudaStream_t stream) {
    if (g->d_packed) return GP_OK;
    int idbits = 1;
    while (idbits < 32 && (1ll << idbits) < g->n) idbits++;
    if (31 - idbits < 3) idbits = 32;   // fewer than 3 spare bits below the sign bit: no code, the node record is always fetched
    cudaError_t e = cudaMalloc(&g->d_packed, sizeof(int) * (size_t)std::max<long long>(g->nnz, 1));
    if (e != cudaSuccess) { gp_set_error("cudaMalloc of the packed CSR failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return GP_ERR_NOMEM; }
    g->idbits = idbits;
    if (idbits == 32) {
        GP_CUDA_TRY(cudaMemcpyAsync(g->d_packed, g->d_indices, sizeof(int) * (size_t)g->nnz, cudaMemcpyDeviceToDevice, stream));
        return GP_OK;
    }
    return gpc_pack_indices(g->d_node_rec, g->d_indices, g->nnz, idbits, g->d_packed, g->num_sms, stream);
}

int push_device_locked(gp_graph *g, const int *d_node_idx, long long S, const double *coef, int L, double rmax,
                       int K, int *d_row, int *d_col, double *d_val, float *d_val32, cudaStream_t stream) {
    GP_REQUIRE(S >= 0, "negative source count");
    GP_REQUIRE(L >= 1 && L <= kMaxLevels, "coef must have 1..%d entries (got %d)", kMaxLevels, L);
    GP_REQUIRE(K >= 1 && K <= kMaxK, "top_k must be in 1..%d (got %d)", kMaxK, K);
    GP_REQUIRE(coef != nullptr, "coef is null");
    GP_REQUIRE(!(rmax != rmax), "rmax is NaN");
    g->last = gp_push_stats{};
    g->cur_L = L; g->cur_rmax = rmax;
    if (S == 0) return GP_OK;
    GP_REQUIRE(d_node_idx && d_row && d_col && d_val, "null device buffer");
    // Every push on a handle shares the control block, the coefficient array and the scratch: a push on another stream
    // (gp_gfpush runs on the handle's own stream, gp_gfpush_device on the caller's) waits for the previous one.
    if (g->ev_recorded) GP_CUDA_TRY(cudaStreamWaitEvent(stream, g->ev_done, 0));
    Plan pl{};
    int rc = make_plan(g, S, L, rmax, false, &pl);
    if (rc != GP_OK) return rc;
    ClusterPlan cp{};
    BucketPlan bp{};
    bool pilot = false;
    if (pl.mode == GP_SCRATCH_HBM) {
        rc = plan_cluster(g, S, L, rmax, K, pl.capF, &cp);
        if (rc != GP_OK) return rc;
        const bool have_hint = g->support_hint > 0 && g->hint_L == L && g->hint_rmax == rmax;
        if (cp.G == 0) {
            rc = plan_bucket(g, S, L, rmax, pl, have_hint ? g->support_hint : 0, &bp);
            if (rc != GP_OK) return rc;
        }
        // no measurement yet and enough sources to pay for one: the first sources run as a pilot (below)
        pilot = bp.nb > 0 && !have_hint && g_push_bucket_nb == 0 && S >= kPilotMinSources;
        if (cp.G > 0 || bp.nb > 0) {   // the slabs only back the first-pass kernel up
            rc = make_plan(g, S, L, rmax, true, &pl);
            if (rc != GP_OK) return rc;
        }
    }
    rc = ensure_scratch(g, pl, S, stream);
    if (rc != GP_OK) return rc;
    GP_CUDA_TRY(cudaMemcpyAsync(g->d_coef, coef, sizeof(double) * L, cudaMemcpyHostToDevice, stream));
    // the error word [4] is sticky until collect_stats reads it
    GP_CUDA_TRY(cudaMemsetAsync(g->d_ctrl, 0, sizeof(unsigned long long) * 4, stream));
    GP_CUDA_TRY(cudaMemsetAsync(g->d_ctrl + 5, 0, sizeof(unsigned long long) * 11, stream));
    char *base = (char *)g->scratch;
    PushParams P{};
    P.indptr = g->d_indptr; P.node_rec = g->d_node_rec; P.indices = g->d_indices; P.n = (int)g->n;
    P.packed = nullptr; P.idbits = 32;
    if (pl.mode == GP_SCRATCH_HBM) {   // MODE 0 / 2 read the CSR copy with the degree codes
        rc = ensure_packed(g, stream);
        if (rc != GP_OK) return rc;
        P.packed = g->d_packed; P.idbits = g->idbits;
    }
    P.node_idx = d_node_idx; P.S = S; P.coef = g->d_coef; P.L = L; P.rmax = rmax; P.K = K;
    P.out_row = d_row; P.out_col = d_col; P.out_val = d_val; P.out_val32 = d_val32;
    P.tab = (Slot *)(base + pl.off_tab);
    P.epoch_base = (int)g->epoch_base;
    P.push_start = (int *)(base + pl.off_push_start);
    P.push_deg = (int *)(base + pl.off_push_deg);
    P.push_val = (double *)(base + pl.off_push_val);
    P.nxt_id = (int *)(base + pl.off_nxt_id);
    P.sup_id = (int *)(base + pl.off_sup_id);
    P.sup_val = (double *)(base + pl.off_cand);
    P.capF = pl.capF; P.capS = pl.capS;
    P.queue = g->d_ctrl + 7; P.stats = g->d_ctrl + 1; P.cum = g->d_ctrl + 16;
    P.max_support = g->d_ctrl + 8; P.phase = g->d_ctrl + 24;
    P.log_id = (int *)(base + pl.off_log_id); P.log_val = (double *)(base + pl.off_log_val); P.capLog = pl.capLog;
    P.hslots = pl.hslots; P.max_probe = std::max(g_push_smem_probe, 1);
    P.redo = nullptr; P.redo_count = nullptr;
    int launches = 0;
    if (cp.G > 0) {
        rc = ensure_packed(g, stream);
        if (rc != GP_OK) return rc;
        if (g->cscratch_bytes < cp.bytes) {
            GP_CUDA_TRY(cudaStreamSynchronize(stream));
            if (g->ev_recorded) GP_CUDA_TRY(cudaEventSynchronize(g->ev_done));
            cudaFree(g->cscratch); g->cscratch = nullptr; g->cscratch_bytes = 0;
            GP_CUDA_TRY(cudaMalloc(&g->cscratch, cp.bytes));
            g->cscratch_bytes = cp.bytes;
        }
        if (g->d_redo_cap < (size_t)S) {
            GP_CUDA_TRY(cudaStreamSynchronize(stream));
            if (g->ev_recorded) GP_CUDA_TRY(cudaEventSynchronize(g->ev_done));
            cudaFree(g->d_redo); g->d_redo = nullptr; g->d_redo_cap = 0;
            GP_CUDA_TRY(cudaMalloc(&g->d_redo, sizeof(int) * (size_t)S));
            g->d_redo_cap = (size_t)S;
        }
        char *cb = (char *)g->cscratch;
        ClusterPushParams C{};
        C.node_rec = g->d_node_rec; C.packed = g->d_packed; C.n = (int)g->n; C.idbits = g->idbits;
        C.node_idx = d_node_idx; C.S = S; C.coef = g->d_coef; C.L = L; C.rmax = rmax; C.K = K;
        C.out_row = d_row; C.out_col = d_col; C.out_val = d_val; C.out_val32 = d_val32;
        C.push_start = (int *)(cb + cp.off_ps); C.push_len = (int *)(cb + cp.off_pl); C.push_add = (double *)(cb + cp.off_pa);
        C.capP = cp.capP;
        C.x_id = (int *)(cb + cp.off_xi); C.x_val = (double *)(cb + cp.off_xv); C.capX = cp.capX;
        C.cand_id = (int *)(cb + cp.off_ci); C.cand_val = (double *)(cb + cp.off_cv);
        C.hub_start = (int *)(cb + cp.off_hs); C.hub_deg = (int *)(cb + cp.off_hd); C.hub_add = (double *)(cb + cp.off_ha);
        C.capHub = cp.capHub; C.hub_min_deg = cp.hub_min_deg;
        C.max_probe = std::max(g_push_cluster_probe, 1);
        C.queue = g->d_ctrl; C.stats = g->d_ctrl + 1; C.cum = g->d_ctrl + 16; C.phase = g->d_ctrl + 24;
        C.redo = g->d_redo; C.redo_count = g->d_ctrl + 6;
        rc = gpc_launch(C, cp.G, cp.clusters, stream);
        if (rc != GP_OK) return rc;
        launches++;
        // second pass: what outgrew the tables or the streams, on the slabs (exits at once when the list is empty)
        P.redo = g->d_redo; P.redo_count = g->d_ctrl + 6;
    } else if (bp.nb > 0) {
        rc = ensure_packed(g, stream);
        if (rc != GP_OK) return rc;
        if (g->d_redo_cap < (size_t)S) {
            GP_CUDA_TRY(cudaStreamSynchronize(stream));
            if (g->ev_recorded) GP_CUDA_TRY(cudaEventSynchronize(g->ev_done));
            cudaFree(g->d_redo); g->d_redo = nullptr; g->d_redo_cap = 0;
            GP_CUDA_TRY(cudaMalloc(&g->d_redo, sizeof(int) * (size_t)S));
            g->d_redo_cap = (size_t)S;
        }
        auto launch_bucket = [&](const BucketPlan &b, long long first, long long last, bool full_merge) -> int {
            if (g->bscratch_bytes < b.bytes) {
                GP_CUDA_TRY(cudaStreamSynchronize(stream));
                if (g->ev_recorded) GP_CUDA_TRY(cudaEventSynchronize(g->ev_done));
                cudaFree(g->bscratch); g->bscratch = nullptr; g->bscratch_bytes = 0;
                GP_CUDA_TRY(cudaMalloc(&g->bscratch, b.bytes));
                g->bscratch_bytes = b.bytes;
            }
            char *bb = (char *)g->bscratch;
            BucketPushParams B{};
            B.node_rec = g->d_node_rec; B.packed = g->d_packed; B.n = (int)g->n; B.idbits = g->idbits; B.nb = b.nb; B.log_nb = b.log_nb; B.block = b.block;
            B.max_probe = std::max(g_push_cluster_probe, 1);
            B.full_merge = full_merge ? 1 : 0;
            B.node_idx = d_node_idx; B.S = last; B.it_base = first; B.coef = g->d_coef; B.L = L; B.rmax = rmax; B.K = K;
            B.out_row = d_row; B.out_col = d_col; B.out_val = d_val; B.out_val32 = d_val32;
            B.pair_id = (int *)(bb + b.off_pi); B.pair_val = (double *)(bb + b.off_pv); B.capPair = b.capPair;
            B.pair_stride = (long long)b.nb * b.capPair; B.log_stride = (long long)b.nb * b.capLog;
            B.group_pairs = gpb_slots(b.block) * g_push_bucket_fill / 8;
            B.log_id = (int *)(bb + b.off_li); B.log_val = (double *)(bb + b.off_lv); B.capLog = b.capLog;
            B.push_start = (int *)(bb + b.off_ps); B.push_len = (int *)(bb + b.off_pl); B.push_add = (double *)(bb + b.off_pa);
            B.capP = b.capP;
            B.sup_id = (int *)(bb + b.off_si); B.sup_val = (double *)(bb + b.off_sv); B.capS = b.capS;
            B.queue = g->d_ctrl; B.stats = g->d_ctrl + 1; B.cum = g->d_ctrl + 16; B.phase = g->d_ctrl + 24;
            B.max_support = g->d_ctrl + 8;
            B.redo = g->d_redo; B.redo_count = g->d_ctrl + 6;
            launches++;
            return gpb_launch(B, (int)std::min<long long>(b.ctas, last - first), stream);
        };
        long long first = 0;
        if (pilot) {
            // Pilot: two sources per SM with the a-priori bucket count, then read the largest support back (the one
            // synchronisation of the first call on a handle) and size the buckets of the rest -- and of later calls -- from it.
            first = std::min<long long>(S, 2ll * g->num_sms);
            rc = launch_bucket(bp, 0, first, true);   // (the full merge counts the support)
            if (rc != GP_OK) return rc;
            unsigned long long h_max = 0;
            GP_CUDA_TRY(cudaMemcpyAsync(&h_max, g->d_ctrl + 8, sizeof h_max, cudaMemcpyDeviceToHost, stream));
            GP_CUDA_TRY(cudaStreamSynchronize(stream));
            if (h_max > 0) {
                g->support_hint = (long long)h_max; g->hint_L = L; g->hint_rmax = rmax;
                rc = plan_bucket(g, S, L, rmax, pl, g->support_hint, &bp);
                if (rc != GP_OK) return rc;
            }
            GP_CUDA_TRY(cudaMemsetAsync(g->d_ctrl, 0, sizeof(unsigned long long), stream));   // the queue restarts at `first`
        }
        if (first < S) {
            rc = launch_bucket(bp, first, S, g_push_bucket_merge != 0);
            if (rc != GP_OK) return rc;
        }
        P.redo = g->d_redo; P.redo_count = g->d_ctrl + 6;
    }
    rc = launch_slab(P, pl, stream);
    if (rc != GP_OK) return rc;
    launches++;
    GP_CUDA_TRY(cudaEventRecord(g->ev_done, stream));
    g->ev_recorded = true;
    g->epoch_base += S;
    g->last.sources = S;
    g->last.ctas = cp.G > 0 ? (long long)cp.clusters * cp.G : bp.nb > 0 ? std::min<long long>(bp.ctas, S) : std::min<long long>(pl.ctas, S);
    g->last.scratch_bytes = (int64_t)(g->scratch_bytes + g->cscratch_bytes + g->bscratch_bytes);
    g->last.scratch_mode = pl.mode; g->last.kernel_launches = launches; g->last.cluster_size = cp.G;
    g->last.table_slots = cp.G > 0 ? kClusterSlots : bp.nb > 0 ? gpb_slots(bp.block) : pl.hslots;
    g->last.bucket_count = bp.nb;
    return GP_OK;
}

// Reads the control block back (synchronises `stream`) and turns device-side flags into errors.
int collect_stats(gp_graph *g, cudaStream_t stream) {
    unsigned long long h[9];
    GP_CUDA_TRY(cudaMemcpyAsync(h, g->d_ctrl, sizeof h, cudaMemcpyDeviceToHost, stream));
    GP_CUDA_TRY(cudaStreamSynchronize(stream));
    if (g->last.bucket_count > 0 && h[8] > 0) {
        // The buckets are sized from the largest support the PILOT saw.  A later call raises the measurement only when more
        // than 2 % of its sources outgrew the table and were handed over: one hub source in ten thousand is cheaper on the
        // slabs than twice the bucket visits for every source of every later call.
        const bool same = g->support_hint > 0 && g->hint_L == g->cur_L && g->hint_rmax == g->cur_rmax;
        const bool many_redone = (long long)h[6] * 50 > g->last.sources;
        if (!same) g->support_hint = (long long)h[8];
        else if (many_redone) g->support_hint = std::max<long long>(g->support_hint, (long long)h[8]);
        g->hint_L = g->cur_L; g->hint_rmax = g->cur_rmax;
    }
    g->last.edges_pushed = (int64_t)h[1];
    g->last.frontier_total = (int64_t)h[2];
    g->last.support_total = (int64_t)h[3];
    if (h[4]) {
        // the error word is sticky across calls until it is read here
        GP_CUDA_TRY(cudaMemsetAsync(g->d_ctrl + 4, 0, sizeof(unsigned long long), stream));
        GP_CUDA_TRY(cudaStreamSynchronize(stream));
    }
    if (h[4] & kErrOverflow) {
        // a truncated list leaves residues behind on the slabs: have the next call re-initialise them
        g->scratch_mode = -1;
        gp_set_error("frontier/support list outgrew its bound (rmax too small for the budget?)");
        return GP_ERR_OVERFLOW;
    }
    if (h[4] & kErrBadSource) { gp_set_error("node_idx contains an id outside [0, %lld)", g->n); return GP_ERR_INVALID; }
    return GP_OK;
}

int graph_finish_create(gp_graph *g) {
    cudaDeviceProp prop;
    GP_CUDA_TRY(cudaGetDeviceProperties(&prop, g->device));
    g->num_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : GP_NUM_SMS_FALLBACK;
    g->smem_optin = prop.sharedMemPerBlockOptin;
    GP_CUDA_TRY(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    GP_CUDA_TRY(cudaEventCreateWithFlags(&g->ev_done, cudaEventDisableTiming));
    GP_CUDA_TRY(cudaMalloc(&g->d_coef, sizeof(double) * kMaxLevels));
    GP_CUDA_TRY(cudaMalloc(&g->d_ctrl, sizeof(unsigned long long) * 32));
    // validate the CSR once, on the device (the reference validates nothing)
    int *d_flag = (int *)(g->d_ctrl);
    GP_CUDA_TRY(cudaMemsetAsync(g->d_ctrl, 0, sizeof(unsigned long long) * 32, g->stream));
    validate_csr_kernel<<<g->num_sms * 4, 256, 0, g->stream>>>(g->d_indptr, g->n, g->d_indices, g->nnz, d_flag);
    GP_CUDA_TRY(cudaGetLastError());
    int flag = 0;
    GP_CUDA_TRY(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    GP_CUDA_TRY(cudaStreamSynchronize(g->stream));
    GP_CUDA_TRY(cudaMalloc(&g->d_node_rec, sizeof(int2) * (size_t)g->n));
    node_rec_kernel<<<g->num_sms * 4, 256, 0, g->stream>>>(g->d_indptr, g->n, g->d_node_rec);
    GP_CUDA_TRY(cudaGetLastError());
    GP_CUDA_TRY(cudaStreamSynchronize(g->stream));
    GP_REQUIRE(flag == 0, "malformed CSR: indptr must start at 0, be non-decreasing and end at nnz; indices must lie in [0, n)");
    return GP_OK;
}

}  // namespace

extern "C" {

int gp_graph_create(const int32_t *indptr, int64_t n_nodes, const int32_t *indices, int64_t nnz, int32_t seed,
                    int device, gp_graph **out) {
    GpRange nvtx_range("gp_graph_create");
    (void)seed;  // stored and never read by the reference either (graph.h:30,40)
    GP_REQUIRE(out != nullptr, "out is null");
    *out = nullptr;
    GP_REQUIRE(indptr != nullptr && (indices != nullptr || nnz == 0), "null CSR array");
    GP_REQUIRE(n_nodes >= 1 && n_nodes < (1ll << 31) - 1, "n_nodes out of range: %lld", (long long)n_nodes);
    GP_REQUIRE(nnz >= 0 && nnz < (1ll << 31), "nnz out of range for int32 CSR: %lld", (long long)nnz);
    GP_REQUIRE(gp_device_count() > 0, "no CUDA device: this library has no CPU fallback");
    DeviceGuard guard(device);
    GP_REQUIRE(guard.ok, "cannot select CUDA device %d", device);
    gp_graph *g = new (std::nothrow) gp_graph();
    if (!g) { gp_set_error("out of host memory"); return GP_ERR_NOMEM; }
    g->device = device; g->n = n_nodes; g->nnz = nnz; g->owns_csr = true;
    auto fail = [&](int rc) { gp_graph_destroy(g); return rc; };
    cudaError_t e = cudaMalloc(&g->d_indptr, sizeof(int) * (size_t)(n_nodes + 1));
    if (e == cudaSuccess) e = cudaMalloc(&g->d_indices, sizeof(int) * (size_t)std::max<int64_t>(nnz, 1));
    if (e == cudaSuccess) e = cudaMemcpy(g->d_indptr, indptr, sizeof(int) * (size_t)(n_nodes + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nnz) e = cudaMemcpy(g->d_indices, indices, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { gp_set_error("CSR upload failed: %s", cudaGetErrorString(e)); return fail(e == cudaErrorMemoryAllocation ? GP_ERR_NOMEM : GP_ERR_CUDA); }
    int rc = graph_finish_create(g);
    if (rc != GP_OK) return fail(rc);
    *out = g;
    return GP_OK;
}

int gp_graph_create_device(const int32_t *d_indptr, int64_t n_nodes, const int32_t *d_indices, int64_t nnz,
                           int device, gp_graph **out) {
    GpRange nvtx_range("gp_graph_create_device");
    GP_REQUIRE(out != nullptr, "out is null");
    *out = nullptr;
    GP_REQUIRE(d_indptr != nullptr && (d_indices != nullptr || nnz == 0), "null CSR array");
    GP_REQUIRE(n_nodes >= 1 && n_nodes < (1ll << 31) - 1, "n_nodes out of range: %lld", (long long)n_nodes);
    GP_REQUIRE(nnz >= 0 && nnz < (1ll << 31), "nnz out of range for int32 CSR: %lld", (long long)nnz);
    GP_REQUIRE(gp_device_count() > 0, "no CUDA device: this library has no CPU fallback");
    DeviceGuard guard(device);
    GP_REQUIRE(guard.ok, "cannot select CUDA device %d", device);
    gp_graph *g = new (std::nothrow) gp_graph();
    if (!g) { gp_set_error("out of host memory"); return GP_ERR_NOMEM; }
    g->device = device; g->n = n_nodes; g->nnz = nnz; g->owns_csr = false;
    g->d_indptr = const_cast<int *>(d_indptr);
    g->d_indices = const_cast<int *>(d_indices);
    int rc = graph_finish_create(g);
    if (rc != GP_OK) { gp_graph_destroy(g); return rc; }
    *out = g;
    return GP_OK;
}

void gp_graph_destroy(gp_graph *g) {
    if (!g) return;
    DeviceGuard guard(g->device);
    if (g->ev_recorded) cudaEventSynchronize(g->ev_done);
    if (g->stream) cudaStreamSynchronize(g->stream);
    if (g->owns_csr) { cudaFree(g->d_indptr); cudaFree(g->d_indices); }
    cudaFree(g->d_node_rec); cudaFree(g->d_packed); cudaFree(g->scratch); cudaFree(g->cscratch); cudaFree(g->bscratch); cudaFree(g->d_redo); cudaFree(g->d_coef); cudaFree(g->d_ctrl); cudaFree(g->d_node); cudaFree(g->d_out);
    if (g->ev_done) cudaEventDestroy(g->ev_done);
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
}

int64_t gp_graph_num_nodes(const gp_graph *g) { return g ? g->n : 0; }
int64_t gp_graph_num_edges(const gp_graph *g) { return g ? g->nnz : 0; }

int gp_graph_configure(gp_graph *g, const gp_push_config *cfg) {
    GP_REQUIRE(g && cfg, "null argument");
    std::lock_guard<std::mutex> lk(g->mu);
    g->cfg = *cfg;
    return GP_OK;
}

int gp_gfpush_device(gp_graph *g, const int32_t *d_node_idx, int64_t S, const double *coef, int32_t L, double rmax,
                     int32_t K, int32_t *d_row_idx, int32_t *d_col_idx, double *d_value, float *d_value32,
                     void *stream) {
    GpRange nvtx_range("gp_gfpush_device");
    GP_REQUIRE(g != nullptr, "graph handle is null");
    std::lock_guard<std::mutex> lk(g->mu);
    DeviceGuard guard(g->device);
    GP_REQUIRE(guard.ok, "cannot select CUDA device %d", g->device);
    return push_device_locked(g, d_node_idx, S, coef, L, rmax, K, d_row_idx, d_col_idx, d_value, d_value32,
                              (cudaStream_t)stream);
}

int gp_gfpush(gp_graph *g, const int32_t *node_idx, int64_t S, const double *coef, int32_t L, double rmax, int32_t K,
              int32_t *row_idx, int32_t *col_idx, double *value) {
    GpRange nvtx_range("gp_gfpush");
    GP_REQUIRE(g != nullptr, "graph handle is null");
    GP_REQUIRE(S >= 0, "negative source count");
    if (S == 0) return GP_OK;
    GP_REQUIRE(node_idx && row_idx && col_idx && value, "null host buffer");
    GP_REQUIRE(K >= 1 && K <= kMaxK, "top_k must be in 1..%d (got %d)", kMaxK, K);
    std::lock_guard<std::mutex> lk(g->mu);
    DeviceGuard guard(g->device);
    GP_REQUIRE(guard.ok, "cannot select CUDA device %d", g->device);
    const size_t slots = (size_t)S * (size_t)K;
    if (g->d_node_cap < (size_t)S) {
        cudaFree(g->d_node); g->d_node = nullptr; g->d_node_cap = 0;
        GP_CUDA_TRY(cudaMalloc(&g->d_node, sizeof(int) * (size_t)S));
        g->d_node_cap = (size_t)S;
    }
    const size_t out_bytes = slots * 16;
    if (g->d_out_cap < out_bytes) {
        cudaFree(g->d_out); g->d_out = nullptr; g->d_out_cap = 0;
        GP_CUDA_TRY(cudaMalloc(&g->d_out, out_bytes));
        g->d_out_cap = out_bytes;
    }
    double *d_val = (double *)g->d_out;
    int *d_row = (int *)((char *)g->d_out + slots * 8);
    int *d_col = d_row + slots;
    // GP_TRACE=1: host wall time of every stage of the host-buffer entry point (diagnostics for end-to-end numbers)
    static const bool trace = getenv("GP_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto t0 = now();
    GP_CUDA_TRY(cudaMemcpyAsync(g->d_node, node_idx, sizeof(int) * (size_t)S, cudaMemcpyHostToDevice, g->stream));
    int rc = push_device_locked(g, g->d_node, S, coef, L, rmax, K, d_row, d_col, d_val, nullptr, g->stream);
    if (rc != GP_OK) return rc;
    const auto t1 = now();
    if (trace) GP_CUDA_TRY(cudaStreamSynchronize(g->stream));
    const auto t2 = now();
    GP_CUDA_TRY(cudaMemcpyAsync(value, d_val, slots * 8, cudaMemcpyDeviceToHost, g->stream));
    GP_CUDA_TRY(cudaMemcpyAsync(row_idx, d_row, slots * 4, cudaMemcpyDeviceToHost, g->stream));
    GP_CUDA_TRY(cudaMemcpyAsync(col_idx, d_col, slots * 4, cudaMemcpyDeviceToHost, g->stream));
    if (trace) GP_CUDA_TRY(cudaStreamSynchronize(g->stream));
    const auto t3 = now();
    rc = collect_stats(g, g->stream);
    if (trace)
        fprintf(stderr, "[gp_gfpush] S=%lld K=%d: enqueue %.3f ms, kernels %.3f ms, D2H of %.1f MB %.3f ms, flags %.3f ms\n",
                (long long)S, K, ms(t0, t1), ms(t1, t2), (double)slots * 16 / 1e6, ms(t2, t3), ms(t3, now()));
    return rc;
}

int gp_gfpush_cumulative_stats(gp_graph *g, gp_push_stats *out, int reset) {
    GP_REQUIRE(g && out, "null argument");
    std::lock_guard<std::mutex> lk(g->mu);
    DeviceGuard guard(g->device);
    GP_CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long h[6];
    GP_CUDA_TRY(cudaMemcpy(h, g->d_ctrl + 16, sizeof h, cudaMemcpyDeviceToHost));
    if (reset) GP_CUDA_TRY(cudaMemset(g->d_ctrl + 16, 0, sizeof h));
    *out = g->last;
    out->edges_pushed = (int64_t)h[0]; out->frontier_total = (int64_t)h[1];
    out->support_total = (int64_t)h[2]; out->sources = (int64_t)h[3];
    out->cluster_sources = (int64_t)h[4]; out->redo_sources = (int64_t)h[5];
    return GP_OK;
}

int gp_gfpush_phase_cycles(gp_graph *g, uint64_t out[8], int reset) {
    GP_REQUIRE(g && out, "null argument");
    std::lock_guard<std::mutex> lk(g->mu);
    DeviceGuard guard(g->device);
    GP_CUDA_TRY(cudaDeviceSynchronize());
    GP_CUDA_TRY(cudaMemcpy(out, g->d_ctrl + 24, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost));
    if (reset) GP_CUDA_TRY(cudaMemset(g->d_ctrl + 24, 0, sizeof(uint64_t) * 8));
    return GP_OK;
}

int gp_gfpush_last_stats(gp_graph *g, gp_push_stats *out) {
    GP_REQUIRE(g && out, "null argument");
    std::lock_guard<std::mutex> lk(g->mu);
    DeviceGuard guard(g->device);
    if (g->last.sources > 0 && g->last.kernel_launches > 0) {
        // the device-pointer entry point is asynchronous: wait for the push, whatever stream it ran on
        if (g->ev_recorded) GP_CUDA_TRY(cudaEventSynchronize(g->ev_done));
        int rc = collect_stats(g, g->stream);
        if (rc != GP_OK) { *out = g->last; return rc; }
    }
    *out = g->last;
    return GP_OK;
}

}  // extern "C"
