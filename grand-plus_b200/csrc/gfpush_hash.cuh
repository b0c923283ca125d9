// gfpush_hash.cuh -- the L2-resident tier of GFPush (included by gfpush.cu inside its anonymous
// namespace; shares PushSmem's helpers, the radix select and `emit`).
//
// Why: the direct-addressed slabs of gfpush_kernel cost one random DRAM read-modify-write per pushed
// edge (measured roof on B200: 21 G fp64 atomics/s once the footprint exceeds L2, 127 G/s inside it --
// profiles/r01_random_access_microbench.txt).  A source only ever touches its own support (Reddit-shape:
// 13 K of 233 K nodes; Amazon2M-shape: 170 K of 2.4 M), so this tier keeps exactly that in a compact
// open-addressed table and sizes the number of concurrently processed sources so that ALL live tables
// fit in the 126 MB L2:
//   * one source per THREAD-BLOCK CLUSTER of G CTAs (G = 1, 2, 4, 8, 16): large supports get many SMs
//     per source instead of many sources per GPU, the level barrier is the hardware cluster barrier and
//     the per-source counters live in the leader CTA's shared memory, reached through DSMEM;
//   * table = three parallel arrays keys[C] / next residue[C] / reserve[C] (20 B per slot), buckets of
//     four keys (one 16-byte probe), C chosen PER SOURCE AND LEVEL from the bound
//     support + edges-about-to-be-pushed, grown by re-insertion (the support grows geometrically, so
//     re-insertion is cheap);
//   * settle walks the table slots coalesced (no frontier list, no random access) whenever the level
//     is large, and a first-touch list when it is small;
//   * top-k scans the reserve array coalesced and wipes the table in the same pass.
// A source whose bound does not fit C_max is appended to a redo list and finished by gfpush_kernel on
// the direct-addressed slabs afterwards.  graph.h:73-126 line references as in gfpush.cu.
#pragma once

namespace cg = cooperative_groups;  // <cooperative_groups.h> is included by gfpush.cu at file scope

constexpr int kEmptyKey = -1;
constexpr int kHashMinSlots = 1024;   // initial table of every source
constexpr int kScanUnroll = 4;

struct HashParams {
    const int *indptr;
    const int *indices;
    int n;
    const int *node_idx;
    long long S;
    const double *coef;
    int L;
    double rmax;
    int K;
    int *out_row;
    int *out_col;
    double *out_val;
    float *out_val32;
    // per-cluster scratch
    int *keys;         // [clusters][Cmax]
    double *nxt;       // [clusters][Cmax]
    double *rsv;       // [clusters][Cmax]
    int *push_start;   // [clusters][capP]
    int *push_deg;     // [clusters][capP]
    double *push_val;  // [clusters][capP]
    int *nxt_id;       // [clusters][capL]  slots of the next frontier (list mode only)
    int *tmp_key;      // [clusters][Cmax]  staging for table growth
    double *tmp_val;   // [clusters][Cmax]
    int Cmax;          // multiple of 4
    long long capP, capL;
    int load_pct;      // table is sized so that bound/C <= load_pct/100
    int list_div;      // a level with fewer than C/list_div edges settles through the first-touch list
    int *redo;         // [S] positions `it` handed to the slab kernel
    unsigned long long *redo_count;
    unsigned long long *queue;
    unsigned long long *stats;  // [0] edges [1] frontier [2] support [3] error flags
    unsigned long long *cum;    // [0] edges [1] frontier [2] support [3] sources [4] hash sources [5] redo
    unsigned long long *phase;  // SM cycles of the leader CTAs: [0] fetch+init [1] grow [2] expand [3] settle [4] top-k
                                // [5] expand of the widest level [6] its settle [7] total
};

// Per-source state shared by the cluster; lives in the leader CTA's shared memory.
struct ClusterState {
    long long it;
    int pushcnt[2];                 // parity = level & 1: push list being read / being built
    unsigned long long elevel[2];   // edges the push list of that parity will traverse
    int nnxt[2];                    // first-touch list length (list mode)
    int claims[2];                  // nodes inserted by the expands of even / odd levels (support = 1 + both)
    int n_out, n_bucket, n_tmp;
    unsigned long long frontier;
    int sel_bin, sel_above, sel_inbin;
    unsigned sel_total;
    unsigned hist[kHistBins];
    unsigned long long bkey[kBucketCap];
    int bid[kBucketCap];
};

template <int BLOCK>
struct HashSmem {
    unsigned off[BLOCK];
    int start[BLOCK];
    double val[BLOCK];
    unsigned warp_scan[BLOCK / 32 + 1];
    unsigned hist[kHistBins];  // this CTA's share of a radix pass (clusters only)
    // thread 0 only (kept out of registers: every thread would carry them)
    long long ph[8], t_prev, wide_expand, wide_settle;
    unsigned long long wide_E, st_edges, st_frontier, st_support, st_sources, st_redo;
    ClusterState cs;
};

template <bool MULTI>
__device__ __forceinline__ void csync() {
    if (MULTI) cg::this_cluster().sync();
    else __syncthreads();
}

__device__ __forceinline__ unsigned warp_sum(unsigned x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// ---- the table ------------------------------------------------------------------------------------
// Probe order of node v: h = v * 2654435761; buckets (hi32(h * n_buckets) + t) mod n_buckets for t = 0, 1, ...;
// inside a bucket the four positions (p0 + i) & 3, p0 = (h >> 13) & 3.  Entries are never removed while a
// source is live, so v sits at the first position of its order that was empty when it was inserted, and a
// lookup may stop at the first empty position.  Observed non-empty keys are final; an observed EMPTY is only
// ever acted on through atomicCAS.
__device__ __forceinline__ void hash_start(int v, unsigned n_buckets, unsigned &b, unsigned &p0) {
    const unsigned h = (unsigned)v * 2654435761u;
    b = __umulhi(h, n_buckets);
    p0 = (h >> 13) & 3u;
}

// k4[p] without a runtime-indexed local array (which would live in local memory)
__device__ __forceinline__ int sel4(const int4 &k, unsigned p) { return p == 0 ? k.x : p == 1 ? k.y : p == 2 ? k.z : k.w; }

// Generic walk from position index i0 (0..3) of bucket b.  Returns the slot or -1 when the table is full
// (cannot happen: the caller sized it from an upper bound; kept so that a bug cannot hang the GPU).
__device__ __forceinline__ int hash_walk(int *keys, unsigned n_buckets, int v, unsigned b, unsigned p0, int i0, bool &claimed) {
    claimed = false;
    for (unsigned probe = 0; probe <= n_buckets; probe++) {
        const int4 k4 = __ldcg(reinterpret_cast<const int4 *>(keys) + b);
        for (int i = i0; i < 4; i++) {
            const unsigned p = (p0 + (unsigned)i) & 3u;
            int k = sel4(k4, p);
            if (k == kEmptyKey) {
                k = atomicCAS(keys + 4 * b + p, kEmptyKey, v);
                if (k == kEmptyKey) { claimed = true; return (int)(4 * b + p); }
            }
            if (k == v) return (int)(4 * b + p);
        }
        i0 = 0;
        b = (b + 1 == n_buckets) ? 0u : b + 1;
    }
    return -1;
}

__device__ __forceinline__ int hash_find_or_claim(int *keys, unsigned n_buckets, int v, bool &claimed) {
    unsigned b, p0;
    hash_start(v, n_buckets, b, p0);
    return hash_walk(keys, n_buckets, v, b, p0, 0, claimed);
}

// U lookups/inserts at once, written so that the memory operations of the U keys overlap (a per-key probe
// loop would serialise them: one L2 round trip at a time per thread):
//   A. compare-and-swap on the first position of every key's order -- most inserts of a growing support end here;
//   B. the others read their bucket (16 bytes) and pick the first position that holds v or looks empty;
//   C. compare-and-swap on those that looked empty;
//   D. whatever is still unresolved (lost a race, bucket full) walks on with the generic loop.
template <int U>
__device__ __forceinline__ unsigned hash_find_or_claim_batch(int *keys, unsigned n_buckets, const int (&v)[U],
                                                             const bool (&ok)[U], int (&h)[U]) {
    unsigned b[U], p0[U];
    int r[U];
    unsigned claims = 0;
#pragma unroll
    for (int q = 0; q < U; q++) {
        hash_start(v[q], n_buckets, b[q], p0[q]);
        r[q] = v[q];
        if (ok[q]) r[q] = atomicCAS(keys + 4 * b[q] + p0[q], kEmptyKey, v[q]);
    }
    bool pending = false;
#pragma unroll
    for (int q = 0; q < U; q++) {
        h[q] = -1;
        if (ok[q]) {
            if (r[q] == kEmptyKey) { claims++; h[q] = (int)(4 * b[q] + p0[q]); }
            else if (r[q] == v[q]) h[q] = (int)(4 * b[q] + p0[q]);
            else { h[q] = -2; pending = true; }
        }
    }
    if (pending) {
        int4 k4[U];
#pragma unroll
        for (int q = 0; q < U; q++)
            if (h[q] == -2) k4[q] = __ldcg(reinterpret_cast<const int4 *>(keys) + b[q]);
        int cand[U];  // position index 1..3 to CAS, or 4 = bucket exhausted
#pragma unroll
        for (int q = 0; q < U; q++) {
            cand[q] = 4;
            if (h[q] == -2) {
                bool hit = false;
#pragma unroll
                for (int i = 3; i >= 1; i--) {
                    const int k = sel4(k4[q], (p0[q] + (unsigned)i) & 3u);
                    if (k == v[q] || k == kEmptyKey) { cand[q] = i; hit = k == v[q]; }
                }
                if (hit) h[q] = (int)(4 * b[q] + ((p0[q] + (unsigned)cand[q]) & 3u));
            }
        }
#pragma unroll
        for (int q = 0; q < U; q++)
            if (h[q] == -2 && cand[q] < 4) r[q] = atomicCAS(keys + 4 * b[q] + ((p0[q] + (unsigned)cand[q]) & 3u), kEmptyKey, v[q]);
#pragma unroll
        for (int q = 0; q < U; q++) {
            if (h[q] == -2) {
                int next_i = cand[q] + 1;  // where the generic walk resumes if this did not settle it
                unsigned bb = b[q];
                if (cand[q] < 4) {
                    const int slot = (int)(4 * b[q] + ((p0[q] + (unsigned)cand[q]) & 3u));
                    if (r[q] == kEmptyKey) { claims++; h[q] = slot; }
                    else if (r[q] == v[q]) h[q] = slot;
                }
                if (h[q] == -2) {
                    if (next_i >= 4) { next_i = 0; bb = (bb + 1 == n_buckets) ? 0u : bb + 1; }
                    bool claimed;
                    h[q] = hash_walk(keys, n_buckets, v[q], bb, p0[q], next_i, claimed);
                    claims += claimed ? 1u : 0u;
                }
            }
        }
    }
    return claims;
}

// Expands the tile staged in shared memory (off/start/val): edges [e_begin, total) in steps of
// e_stride are this CTA's share.  LIST: atomics return the old residue and first touches are appended
// to nxt_id; otherwise the adds are fire-and-forget (RED) and settle will walk the table.
template <int BLOCK, int U, bool LIST>
__device__ __forceinline__ unsigned expand_edges(const HashParams &P, HashSmem<BLOCK> &sm, ClusterState *cs, int src,
                                                 unsigned total, unsigned e_begin, unsigned e_stride, int *keys,
                                                 double *nxt, unsigned n_buckets, int *nxt_id, int *list_count,
                                                 unsigned long long *err) {
    const int lane = gp_lane();
    unsigned claims = 0;
    for (unsigned e0 = e_begin + (unsigned)(threadIdx.x & ~31); e0 < total; e0 += e_stride) {
        int v[U];
        double add[U];
        bool ok[U];
#pragma unroll
        for (int q = 0; q < U; q++) {
            const unsigned e = e0 + q * BLOCK + lane;
            ok[q] = e < total;
            v[q] = src; add[q] = 0.0;
            if (ok[q]) {
                const int t = owner_of_edge<BLOCK>(sm.off, e);
                const int st = sm.start[t];
                add[q] = sm.val[t];
                if (st >= 0) v[q] = __ldg(P.indices + st + (e - sm.off[t]));  // graph.h:96-97
            }
        }
        int h[U];
        claims += hash_find_or_claim_batch<U>(keys, n_buckets, v, ok, h);
#pragma unroll
        for (int q = 0; q < U; q++) {
            if (ok[q] && h[q] < 0) atomicOr(err, kErrOverflow);
            if (LIST) {
                bool fresh = false;
                if (h[q] >= 0) fresh = atomicAdd(nxt + h[q], add[q]) == 0.0;  // graph.h:98
                const long long pos = warp_append_pos(fresh, P.capL, list_count, err);
                if (pos >= 0) nxt_id[pos] = h[q];
            } else {
                if (h[q] >= 0) atomicAdd(nxt + h[q], add[q]);  // result unused: RED.E.ADD.F64
            }
        }
    }
    return claims;
}

template <int BLOCK, bool MULTI>
__global__ void __launch_bounds__(BLOCK, 1) gfpush_hash_kernel(HashParams P) {
    constexpr int U = 4096 / BLOCK;  // edges in flight per thread: 4096 per CTA whatever the block size
    __shared__ HashSmem<BLOCK> sm;
    const int tid = threadIdx.x;
    const int lane = gp_lane();
    unsigned rank = 0, G = 1;
    ClusterState *cs = &sm.cs;
    if (MULTI) {
        cg::cluster_group cl = cg::this_cluster();
        rank = cl.block_rank();
        G = cl.num_blocks();
        cs = cl.map_shared_rank(&sm.cs, 0);
    }
    const bool leader = rank == 0;
    const long long cluster_id = blockIdx.x / G;
    const unsigned gtid = rank * BLOCK + tid;
    const unsigned GT = G * BLOCK;

    int *keys = P.keys + cluster_id * (long long)P.Cmax;
    double *nxt = P.nxt + cluster_id * (long long)P.Cmax;
    double *rsv = P.rsv + cluster_id * (long long)P.Cmax;
    int *push_start = P.push_start + cluster_id * P.capP;
    int *push_deg = P.push_deg + cluster_id * P.capP;
    double *push_val = P.push_val + cluster_id * P.capP;
    int *nxt_id = P.nxt_id + cluster_id * P.capL;
    int *tmp_key = P.tmp_key + cluster_id * (long long)P.Cmax;
    double *tmp_val = P.tmp_val + cluster_id * (long long)P.Cmax;
    unsigned long long *err = P.stats + 3;

    const long long t_begin = clock64();
    if (tid == 0) {
        for (int i = 0; i < 8; i++) sm.ph[i] = 0;
        sm.t_prev = t_begin;
        sm.st_edges = 0; sm.st_frontier = 0; sm.st_support = 0; sm.st_sources = 0; sm.st_redo = 0;
    }
#define GP_PHASE(i) do { if (tid == 0) { const long long t_now = clock64(); sm.ph[i] += t_now - sm.t_prev; sm.t_prev = t_now; } } while (0)

    for (;;) {
        csync<MULTI>();  // the previous source is finished everywhere (its counters may be reset)
        if (leader && tid == 0) {
            cs->it = (long long)atomicAdd(P.queue, 1ull);
            cs->pushcnt[0] = 0; cs->pushcnt[1] = 0; cs->elevel[0] = 0; cs->elevel[1] = 0;
            cs->nnxt[0] = 0; cs->nnxt[1] = 0;
            cs->claims[0] = 0; cs->claims[1] = 0; cs->n_out = 0; cs->n_bucket = 0; cs->n_tmp = 0; cs->frontier = 0;
        }
        csync<MULTI>();
        const long long it = cs->it;
        if (it >= P.S) break;
        const int src = P.node_idx[it];
        if (src < 0 || src >= P.n) {  // refuse instead of reading out of bounds
            if (leader) {
                if (tid == 0) atomicOr(err, kErrBadSource);
                for (int i = tid; i < P.K; i += BLOCK) emit(P, it, 0, i, 0, 0.0);
            }
            continue;
        }
        int C = min(kHashMinSlots, P.Cmax);
        // level 0: residue = {src: 1}, reserve = {src: 0} (graph.h:80-81); settled right away by one thread
        if (leader && tid == 0) {
            bool claimed;
            const int h = hash_find_or_claim(keys, (unsigned)C / 4, src, claimed);
            rsv[h] = P.coef[0];
            cs->frontier = 1;
            if (P.L > 1) {
                const int a = P.indptr[src], b = P.indptr[src + 1];
                const unsigned d = (unsigned)(b - a);
                if (d == 0) { push_start[0] = -1; push_deg[0] = 1; push_val[0] = 1.0; cs->pushcnt[0] = 1; cs->elevel[0] = 1; }
                else if (1.0 >= P.rmax * (double)d) {
                    push_start[0] = a; push_deg[0] = (int)d; push_val[0] = 1.0 / (double)d; cs->pushcnt[0] = 1; cs->elevel[0] = d;
                }
            }
        }
        csync<MULTI>();

        GP_PHASE(0);
        bool aborted = false;
        unsigned long long att_edges = 0;
        if (tid == 0) { sm.wide_expand = 0; sm.wide_settle = 0; sm.wide_E = 0; }
        int seen0 = 0, seen1 = 0;  // claims[0], claims[1] as last read while stable (see below)
        for (int level = 0; level < P.L - 1; level++) {  // graph.h:83
            const int par = level & 1;
            const int n_push = cs->pushcnt[par];
            const unsigned long long E = cs->elevel[par];
            // claims[par] is being added to by CTAs already expanding this level; claims[par ^ 1] is final
            // until the next level, and seen<par> was read one level ago when IT was final
            if (par) seen0 = cs->claims[0]; else seen1 = cs->claims[1];
            const int n_sup0 = 1 + seen0 + seen1;
            // ------------------------------------------------------------ size the table for this level
            // support after this level <= support now + edges pushed now
            const unsigned long long bound = (unsigned long long)n_sup0 + E;
            if (bound * 100ull > (unsigned long long)P.Cmax * 90ull) { aborted = true; break; }  // cluster-uniform
            // grow only when the bound passes 85 % of the table, then to load_pct of the bound: the hysteresis keeps
            // the later, smaller levels from re-inserting the whole support again
            int Cnew = C;
            if (bound * 100ull > (unsigned long long)C * 85ull) {
                unsigned long long want = (bound * 100ull + P.load_pct - 1) / P.load_pct;
                want = (want + 1023ull) & ~1023ull;
                Cnew = (int)min(want, (unsigned long long)P.Cmax);
            }
            if (Cnew > C) {
                // move the entries out, wipe, re-insert at the new size (next residues are all zero here)
                for (unsigned j = gtid; j < (unsigned)C; j += GT) {
                    const int k = __ldcg(keys + j);
                    const long long pos = warp_append_pos(k != kEmptyKey, P.Cmax, &cs->n_tmp, err);
                    if (pos >= 0) { tmp_key[pos] = k; tmp_val[pos] = __ldcg(rsv + j); keys[j] = kEmptyKey; rsv[j] = 0.0; }
                }
                csync<MULTI>();
                const int n_tmp = cs->n_tmp;
                for (unsigned i = gtid; i < (unsigned)n_tmp; i += GT) {
                    bool claimed;
                    const int h = hash_find_or_claim(keys, (unsigned)Cnew / 4, __ldcg(tmp_key + i), claimed);
                    if (h >= 0) rsv[h] = __ldcg(tmp_val + i); else atomicOr(err, kErrOverflow);
                }
                csync<MULTI>();
                if (leader && tid == 0) cs->n_tmp = 0;
                C = Cnew;
            }
            GP_PHASE(1);
            const unsigned n_buckets = (unsigned)C / 4;
            const bool use_list = E * (unsigned long long)P.list_div < (unsigned long long)C;
            // the counters of the other parity were last read one level ago: reset them for this level's settle
            // (readers of those finished before the last cluster barrier; writers start after the next one)
            if (leader && tid == 0) { cs->pushcnt[par ^ 1] = 0; cs->elevel[par ^ 1] = 0; cs->nnxt[par ^ 1] = 0; }
            // ---------------------------------------------------------------- expand (graph.h:94-100)
            // Tiles of BLOCK push-list entries are prefix-summed by degree and their edges dealt to threads by
            // rank.  Few tiles: every CTA of the cluster scans every tile and takes a share of its edges (a hub
            // is expanded by the whole cluster).  Many tiles: each CTA takes whole tiles.
            unsigned claims = 0;
            const int n_tiles = (n_push + BLOCK - 1) / BLOCK;
            const bool split_tiles = MULTI && (unsigned)n_tiles >= 4 * G;
            for (int tile = split_tiles ? (int)rank : 0; tile < n_tiles; tile += split_tiles ? (int)G : 1) {
                const int j = tile * BLOCK + tid;
                unsigned d_push = 0;
                int start = 0;
                double val = 0.0;
                if (j < n_push) { d_push = (unsigned)push_deg[j]; start = push_start[j]; val = push_val[j]; }
                unsigned total;
                const unsigned excl = gp_block_exclusive_scan<BLOCK>(d_push, sm.warp_scan, total);
                sm.off[tid] = excl; sm.start[tid] = start; sm.val[tid] = val;
                __syncthreads();
                const unsigned e_begin = split_tiles ? 0u : rank * BLOCK * U;
                const unsigned e_stride = (split_tiles ? 1u : G) * BLOCK * U;
                if (use_list)
                    claims += expand_edges<BLOCK, U, true>(P, sm, cs, src, total, e_begin, e_stride, keys, nxt, n_buckets,
                                                        nxt_id, &cs->nnxt[par], err);
                else
                    claims += expand_edges<BLOCK, U, false>(P, sm, cs, src, total, e_begin, e_stride, keys, nxt, n_buckets,
                                                         nxt_id, &cs->nnxt[par], err);
                __syncthreads();
            }
            claims = warp_sum(claims);
            if (lane == 0 && claims) atomicAdd(&cs->claims[par], (int)claims);
            att_edges += E;
            csync<MULTI>();
            long long t_e = 0;
            if (tid == 0) t_e = clock64() - sm.t_prev;
            GP_PHASE(2);
            // ---------------------------------------------------------------- settle (graph.h:85-93,102)
            const int next_level = level + 1;
            const bool will_push = next_level < P.L - 1;
            const double c = P.coef[next_level];
            int *pushcnt = &cs->pushcnt[par ^ 1];
            const unsigned n_items = use_list ? (unsigned)min((long long)cs->nnxt[par], P.capL) : (unsigned)C;
            unsigned n_front = 0;
            unsigned long long e_next = 0;
            for (unsigned base = 0; base < n_items; base += GT * kScanUnroll) {
                int slot[kScanUnroll];
                double x[kScanUnroll];
                bool ok[kScanUnroll];
#pragma unroll
                for (int q = 0; q < kScanUnroll; q++) {
                    const unsigned i = base + q * GT + gtid;
                    ok[q] = i < n_items;
                    slot[q] = (int)i;
                    if (use_list && ok[q]) slot[q] = __ldcg(nxt_id + i);
                }
#pragma unroll
                for (int q = 0; q < kScanUnroll; q++) {
                    x[q] = ok[q] ? __ldcg(nxt + slot[q]) : 0.0;
                    ok[q] = ok[q] && x[q] != 0.0;
                }
                int v[kScanUnroll], a[kScanUnroll], b[kScanUnroll];
                double r0[kScanUnroll];
#pragma unroll
                for (int q = 0; q < kScanUnroll; q++) {
                    v[q] = 0; r0[q] = 0.0;
                    if (ok[q]) { v[q] = __ldcg(keys + slot[q]); r0[q] = __ldcg(rsv + slot[q]); }
                }
#pragma unroll
                for (int q = 0; q < kScanUnroll; q++) {
                    a[q] = 0; b[q] = 0;
                    if (ok[q] && will_push) { a[q] = __ldg(P.indptr + v[q]); b[q] = __ldg(P.indptr + v[q] + 1); }
                }
#pragma unroll
                for (int q = 0; q < kScanUnroll; q++) {
                    if (ok[q]) {
                        n_front++;
                        nxt[slot[q]] = 0.0;
                        rsv[slot[q]] = r0[q] + c * x[q];   // reserve[v] += coef * r (graph.h:90)
                    }
                    bool push = false;
                    int st = -1, dg = 1;
                    double val = x[q];
                    if (ok[q] && will_push) {
                        const unsigned d = (unsigned)(b[q] - a[q]);
                        if (d == 0) push = true;                                   // graph.h:91-93: back to the source
                        else if (x[q] >= P.rmax * (double)d) {                     // graph.h:94
                            push = true; st = a[q]; dg = (int)d; val = x[q] / (double)d;  // graph.h:95
                        }
                    }
                    if (push) e_next += (unsigned)dg;
                    const long long pp = warp_append_pos(push, P.capP, pushcnt, err);
                    if (pp >= 0) { push_start[pp] = st; push_deg[pp] = dg; push_val[pp] = val; }
                }
            }
            n_front = warp_sum(n_front);
            e_next = warp_sum64(e_next);
            if (lane == 0) {
                if (n_front) atomicAdd(&cs->frontier, (unsigned long long)n_front);
                if (e_next) atomicAdd(&cs->elevel[par ^ 1], e_next);
            }
            csync<MULTI>();
            if (tid == 0 && E >= sm.wide_E) { sm.wide_E = E; sm.wide_expand = t_e; sm.wide_settle = clock64() - sm.t_prev; }
            GP_PHASE(3);
        }
        if (tid == 0) { sm.ph[5] += sm.wide_expand; sm.ph[6] += sm.wide_settle; }

        // ------------------------------------------------------------------ aborted: hand over to the slabs
        if (aborted) {
            for (unsigned j = gtid; j < (unsigned)C; j += GT) { keys[j] = kEmptyKey; rsv[j] = 0.0; }
            if (leader && tid == 0) {
                const unsigned long long pos = atomicAdd(P.redo_count, 1ull);
                P.redo[pos] = (int)it;
                sm.st_redo++;
            }
            continue;
        }

        // ------------------------------------------------------------------ top-k, graph.h:111-126
        int shift = 52, bits = 11;
        unsigned long long prefix = 0;
        int kk = P.K;
        bool first = true;
        unsigned long long Tkey = 0;
        int want_bucket = 0;
        for (;;) {
            for (int i = tid; i < kHistBins; i += BLOCK) { sm.hist[i] = 0; if (leader) sm.cs.hist[i] = 0; }
            csync<MULTI>();
            unsigned *hist = MULTI ? sm.hist : sm.cs.hist;
            const int pshift = shift + bits;  // digits above this one must equal `prefix`
            for (unsigned j = gtid; j < (unsigned)C; j += GT) {
                const double x = __ldcg(rsv + j);
                if (x > 0.0) {
                    const unsigned long long key = (unsigned long long)__double_as_longlong(x);
                    if (first || (key >> pshift) == prefix)
                        atomicAdd(&hist[(unsigned)((key >> shift) & ((1ull << bits) - 1ull))], 1u);
                }
            }
            if (MULTI) {
                __syncthreads();
                for (int i = tid; i < kHistBins; i += BLOCK) {
                    const unsigned hcount = sm.hist[i];
                    if (hcount) atomicAdd(&cs->hist[i], hcount);
                }
            }
            csync<MULTI>();
            if (leader) {
                const unsigned total = select_bin_generic<BLOCK>(sm.cs.hist, sm.warp_scan, 1 << bits, kk, first,
                                                                 &sm.cs.sel_bin, &sm.cs.sel_above, &sm.cs.sel_inbin);
                if (tid == 0) sm.cs.sel_total = total;
            }
            csync<MULTI>();
            const unsigned total = cs->sel_total;
            if (first) kk = min(kk, (int)total);  // k = min(K, #positive): graph.h:113 + the v>0 filter of :121
            if (kk == 0) { want_bucket = 0; Tkey = ~0ull; break; }
            const int bin = cs->sel_bin, above = cs->sel_above, inbin = cs->sel_inbin;
            Tkey = (prefix << bits) | (unsigned long long)bin;
            want_bucket = kk - above;
            if (inbin <= kBucketCap || shift == 0) break;
            kk = want_bucket; first = false; prefix = Tkey;
            const int nshift = shift >= 11 ? shift - 11 : 0;
            const int nbits = shift >= 11 ? 11 : shift;
            shift = nshift; bits = nbits;
        }
        // final pass: everything above the boundary bucket is selected, the bucket goes to the leader's
        // shared memory, and the table is wiped on the way (next residues are already zero)
        for (unsigned j = gtid; j < (unsigned)C; j += GT) {
            const double x = __ldcg(rsv + j);
            const int id = __ldcg(keys + j);
            if (id != kEmptyKey) { keys[j] = kEmptyKey; rsv[j] = 0.0; }
            if (x > 0.0) {
                const unsigned long long key = (unsigned long long)__double_as_longlong(x);
                const unsigned long long t = key >> shift;
                if (t > Tkey) {
                    emit(P, it, src, atomicAdd(&cs->n_out, 1), id, x);
                } else if (t == Tkey) {
                    const int pos = atomicAdd(&cs->n_bucket, 1);
                    if (pos < kBucketCap) { cs->bkey[pos] = key; cs->bid[pos] = id; }
                }
            }
        }
        csync<MULTI>();
        if (leader) {
            // rank-count the boundary bucket: keep its `want_bucket` largest (ties: lower slot first)
            const int nb = min(sm.cs.n_bucket, kBucketCap);
            for (int i = tid; i < nb; i += BLOCK) {
                const unsigned long long ki = sm.cs.bkey[i];
                int rk = 0;
                for (int q = 0; q < nb; q++) {
                    const unsigned long long kq = sm.cs.bkey[q];
                    rk += (kq > ki) || (kq == ki && q < i);
                }
                if (rk < want_bucket)
                    emit(P, it, src, atomicAdd(&sm.cs.n_out, 1), sm.cs.bid[i], __longlong_as_double((long long)ki));
            }
            __syncthreads();
            // unfilled slots read (0, 0, 0.0): what graph.h:117-126 leaves in the caller-zeroed arrays
            for (int i = sm.cs.n_out + tid; i < P.K; i += BLOCK) emit(P, it, 0, i, 0, 0.0);
            if (tid == 0) {
                GP_PHASE(4);
                sm.st_sources++; sm.st_edges += att_edges; sm.st_frontier += sm.cs.frontier;
                sm.st_support += (unsigned)(1 + sm.cs.claims[0] + sm.cs.claims[1]);
            }
        }
    }
    if (leader && tid == 0) {
        atomicAdd(P.stats + 0, sm.st_edges);
        atomicAdd(P.stats + 1, sm.st_frontier);
        atomicAdd(P.stats + 2, sm.st_support);
        atomicAdd(P.cum + 0, sm.st_edges);
        atomicAdd(P.cum + 1, sm.st_frontier);
        atomicAdd(P.cum + 2, sm.st_support);
        atomicAdd(P.cum + 3, sm.st_sources);
        atomicAdd(P.cum + 4, sm.st_sources);
        atomicAdd(P.cum + 5, sm.st_redo);
        sm.ph[7] = clock64() - t_begin;
        for (int i = 0; i < 8; i++) atomicAdd(P.phase + i, (unsigned long long)sm.ph[i]);
    }
#undef GP_PHASE
    // a CTA must not exit while cluster peers may still address its shared memory
    csync<MULTI>();
}
