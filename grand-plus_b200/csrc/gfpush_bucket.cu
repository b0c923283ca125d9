// gfpush_bucket.cu -- GFPush + top-k for every graph beyond the dense shared-memory mode: accumulate by HASH BUCKETS, one
// bucket at a time in a shared-memory table, instead of random read-modify-writes in HBM.
// Same computation as gfpush.cu (Graph::gfpush_omp, /root/reference/precompute/graph.h:53-131), one CTA per source.
//
// The slab kernel (gfpush.cu MODE 0) touches one 16-byte slot per pushed edge and per settled node at a random HBM
// address: a 64-byte fetch and a 32-byte write-back for 8 useful bytes, 62 MB of DRAM traffic per Amazon2M-shape source
// (11 x the algorithmic bytes), bound by the DRAM random-access rate (profiles/r02_gfpush.md).  Here
//   * expand APPENDS every pushed edge (packed node, r/deg) to the stream of the node's bucket, bucket = the top bits of
//     hash(node) (nb = 2^k buckets, chosen so that a source's support per bucket fits the table; one shared-memory
//     counter per bucket).  Push-list entries are cut into chunks of 128 edges that one warp expands with four coalesced
//     CSR reads per lane (no owner search); on levels of SHORT entries (low-degree graphs) a warp packs up to 32 entries:
//     a scan of their lengths, then every lane finds the owner of its edge slot by a shuffle binary search;
//   * a level is then settled bucket by bucket: the bucket's pairs are read back coalesced and accumulated into an
//     open-addressed {key, residue} table in shared memory (4-key buckets, four pairs in flight per thread); the threads
//     scan the table, compact the occupied slots into a per-warp queue and settle them 32 at a time: reserve += coef * r
//     is appended to the bucket's reserve log, the push decision uses the degree code carried in the key (gfpush.cu) and
//     fetches {start, degree} only for the few nodes that pass, the slot is emptied for the next visit.  A level whose
//     pushed edges all fit ONE table fill skips the streams: expand accumulates straight into the table;
//   * the reserve is merged for the TOP-K CANDIDATES only: a node can be among the K largest only if one of its <= L
//     contributions is at least (K-th largest reserve) / L; a lower bound of that reserve is taken after every level (one
//     histogram pass over the threads' largest contributions), and after the last level the logs stream once through a
//     table that holds just the candidates.  The K largest are then selected as in gfpush.cu (gfpush_shared.cuh).
// Because the table only ever holds one visit of one level and nodes keep no slot between levels, the table can be SMALL:
// the kernel is a template over the CTA size with 16 table slots per thread -- 1024 threads (one CTA per SM, 16 384
// slots) for supports far beyond shared memory, 512 threads (TWO sources per SM, 8 192 slots each) for supports of the
// order of one table, 256 threads (three per SM) for tiny supports: the kernel is bound by barriers and dependent
// shared-memory round trips, and a second source on the SM fills them (+23 % / +35 % over the shared-memory-table kernel
// on the Reddit- / MAG-shape graphs).
// All streams are written and read coalesced; nothing is read-modify-written in HBM.  A source whose bucket stream or
// table overflows is handed to the slab kernel through the redo list.  Measured history, including the variants that
// lost: profiles/r02_gfpush.md sections 5 and 7.
#include "gfpush_bucket.h"
#include "gfpush_shared.cuh"

#include <algorithm>

namespace gpp {
namespace {

constexpr int kEmpty = -1;
constexpr int kChunk = 128;             // edges per push-list entry
constexpr int kBigLen = 4096;           // longer entries stay whole and are expanded by all warps together

// Geometry of one CTA (template parameter BB = threads): the table has 16 slots per thread, so
//   BB = 1024: one CTA per SM with a 16 384-slot table (192 KB);
//   BB =  512: TWO CTAs per SM with 8 192-slot tables (96 KB each) -- two sources in flight per SM, the barriers and dependent
//              shared-memory round trips of one overlap the other's (the kernel is latency-bound, profiles/r02_gfpush.md 5);
//   BB =  256: three CTAs per SM with 4 096-slot tables.
template <int BB>
struct Geo {
    static constexpr int kSlots = BB * 16;                // slots of the shared-memory table
    static constexpr int kBuckets4 = kSlots / 4;          // 4-key buckets of the table: one 16-byte shared-memory read per probe
    static constexpr int SPT = kSlots / BB;               // table slots per thread in the scan
    static constexpr int kHashBits = BB == 1024 ? 12 : BB == 512 ? 11 : 10;   // log2(kBuckets4)
    static constexpr int kBigCap = BB == 1024 ? 32 : 8;
    static constexpr int kListCap = BB == 1024 ? 1024 : BB * 5 / 8;   // push-list entries / top-k survivors kept in shared memory
    static constexpr int kCandCap = BB;                   // push candidates of one table scan kept in shared memory
    static constexpr int kCandMax = kSlots / 2;           // nodes the candidate merge can hold in one table fill
    static constexpr int kMinCtas = BB == 1024 ? 1 : BB == 512 ? 2 : 3;
};

__device__ __forceinline__ unsigned hash_node(unsigned id) { return id * 2654435761u; }

// Slot of packed node `vp` in the table (claiming one when it is new), or -1 when `max_probe` buckets hold neither it nor an
// empty slot.  Keys are never removed while a bucket is live and empties are taken in index order, so an observed key is final.
template <int kBuckets4>
__device__ __forceinline__ int find_slot(int *keys, unsigned bucket, int vp, int max_probe, bool &claimed) {
    unsigned b = bucket;
    claimed = false;
    for (int probe = 0; probe < max_probe; probe++, b = (b + 1) & (kBuckets4 - 1)) {
        const int4 k4 = *reinterpret_cast<const int4 *>(keys + 4 * b);
        const int kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int k = kk[i];
            if (k == kEmpty) {
                k = atomicCAS(keys + 4 * b + i, kEmpty, vp);
                if (k == kEmpty) { claimed = true; return (int)(4 * b + i); }
            }
            if (k == vp) return (int)(4 * b + i);
        }
    }
    return -1;
}

// Adds four pairs (v_q, a_q) into the table (graph.h:98 / :106); v_q == kEmpty: no pair.  Returns true when a probe sequence
// ran out (the source is then handed to the slab kernel).  The kernel is bound by the latency of dependent shared-memory
// operations, so every step runs for all four pairs before the next one: four key-bucket reads, four claims, four adds are in
// flight together.  The home bucket settles ~9 of 10 pairs; the rest take find_slot.
template <int kBuckets4, int kHashBits>
__device__ __noinline__ bool accum4_fn(int *s_keys, double *s_vals, const unsigned idmask, const int bshift, const int max_probe,
                                       const int v0, const int v1, const int v2, const int v3,
                                       const double a0, const double a1, const double a2, const double a3) {
    const int vp[4] = {v0, v1, v2, v3};
    const double av[4] = {a0, a1, a2, a3};
    bool overflow = false;
    int at[4];     // slot of the key / of the first free slot in the home bucket
    int st[4];     // 0 skip, 1 key found, 2 free slot to claim, 3 bucket holds other keys only
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const unsigned hb = (hash_node((unsigned)vp[q] & idmask) >> (bshift - kHashBits)) & (kBuckets4 - 1);
        const int4 k4 = *reinterpret_cast<const int4 *>(s_keys + 4 * hb);
        const int kk[4] = {k4.x, k4.y, k4.z, k4.w};
        at[q] = 4 * (int)hb; st[q] = 3;
#pragma unroll
        for (int i = 3; i >= 0; i--) {   // (free slots are taken in index order: the first free one ends the search)
            if (kk[i] == vp[q]) { st[q] = 1; at[q] = 4 * (int)hb + i; }
            else if (kk[i] == kEmpty) { st[q] = 2; at[q] = 4 * (int)hb + i; }
        }
        if (vp[q] == kEmpty) st[q] = 0;
    }
    int won[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        won[q] = 0;
        if (st[q] == 2) won[q] = atomicCAS(s_keys + at[q], kEmpty, vp[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        // a home bucket seen FULL of other keys stays that way (keys are final while the visit is live): probe on from
        // the next bucket; after a lost claim the home bucket may still have a free slot: probe it again
        unsigned first = ((unsigned)at[q] >> 2) + (st[q] == 3 ? 1u : 0u);
        if (st[q] == 2 && won[q] != kEmpty && won[q] != vp[q]) st[q] = 3;   // somebody else took the slot for another node
        if (st[q] == 3) {
            bool claimed;
            at[q] = find_slot<kBuckets4>(s_keys, first & (kBuckets4 - 1), vp[q], max_probe, claimed);
            if (at[q] < 0) { overflow = true; st[q] = 0; }
        }
    }
    // (atomicAdd on a shared-memory double is ptxas' ATOMS.CAST.SPIN loop; a hand-written compare-and-swap against
    // 0.0 for freshly claimed slots measured SLOWER than leaving every add to it)
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (st[q] != 0) atomicAdd(s_vals + at[q], av[q]);
    return overflow;
}

// The r-th largest of the lanes' values (bit patterns of non-negative doubles; 0 when fewer than r lanes hold a positive one).
__device__ __forceinline__ long long warp_rth_largest(long long v, int r) {
    const int lane = threadIdx.x & 31;
    long long r_val = 0;
    for (int i = 0; i < r; i++) {
        long long mx = v;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        r_val = mx;
        const unsigned holders = __ballot_sync(0xffffffffu, v == mx);
        if (lane == __ffs(holders) - 1) v = 0;
    }
    return r_val;
}

// Slot of packed node `vp` if the table holds it, else -1 (never claims).
template <int kBuckets4>
__device__ __forceinline__ int lookup_slot(const int *keys, unsigned bucket, int vp) {
    unsigned b = bucket;
    for (int probe = 0; probe < kBuckets4; probe++, b = (b + 1) & (kBuckets4 - 1)) {
        const int4 k4 = *reinterpret_cast<const int4 *>(keys + 4 * b);
        if (k4.x == vp) return (int)(4 * b);
        if (k4.y == vp) return (int)(4 * b + 1);
        if (k4.z == vp) return (int)(4 * b + 2);
        if (k4.w == vp) return (int)(4 * b + 3);
        if (k4.w == kEmpty) return -1;   // (the used slots of a bucket are a prefix)
    }
    return -1;
}

template <int BB>
struct BSmem {
    int start[Geo<BB>::kListCap];   // push list of the level: first kListCap entries {start, len, add}; during the top-k (the
                                    // list is dead then) `add` / `start` hold the survivors of the pre-filter
    unsigned len[Geo<BB>::kListCap];
    double add[Geo<BB>::kListCap];
    unsigned warp_scan[BB / 32 + 1];
    double wtau[BB / 32];
    unsigned short wq[BB / 32][64];   // settle: per-warp queue of occupied table slots
    union {
        struct {   // top-k (block_topk): the histogram is dead once the boundary bucket is collected
            union {
                unsigned hist[kHistBins];
                struct {
                    unsigned long long bkey[kBucketCap];
                    int bid[kBucketCap];
                };
            };
        } sel;
        struct {
            double r[Geo<BB>::kCandCap];
            unsigned key[Geo<BB>::kCandCap];
        } cand;
    };
    long long it;
    unsigned n_edges;          // pushed edges of the level being built (sum of the push-list entries' lengths)
    int n_push, n_sel, n_sup, n_all, n_list, next_item, n_big, n_cand;   // n_sup: reserves written out, n_all: nodes of the support
    int big_st[Geo<BB>::kBigCap];
    unsigned big_len[Geo<BB>::kBigCap];
    double big_add[Geo<BB>::kBigCap];
    int n_out, n_bucket;
    int sel_bin, sel_above, sel_inbin;
    int ovf;
    int full;                  // a probe sequence ran out: the source goes to the slab kernel, stop probing
    long long ph[8], t_prev;
};

template <int BB>
__global__ void __launch_bounds__(BB, Geo<BB>::kMinCtas) gfpush_bucket_kernel(const BucketPushParams P) {
    constexpr int kSlots = Geo<BB>::kSlots, kBuckets4 = Geo<BB>::kBuckets4, SPT = Geo<BB>::SPT, kHashBits = Geo<BB>::kHashBits;
    constexpr int kBigCap = Geo<BB>::kBigCap, kListCap = Geo<BB>::kListCap, kCandCap = Geo<BB>::kCandCap;
    constexpr int kCandMax = Geo<BB>::kCandMax;
    const int kGroupPairs = P.group_pairs;   // buckets are visited together while their pairs stay below this table load
    __shared__ BSmem<BB> sm;
    extern __shared__ __align__(16) double s_vals[];                                  // [kSlots] the bucket's residues / reserves
    int *s_keys = reinterpret_cast<int *>(s_vals + kSlots);             // [kSlots] packed node, kEmpty = free
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_keys + kSlots);    // [nb] pairs per bucket (this level)
    unsigned *s_lcnt = s_cnt + P.nb;                                    // [nb] reserve-log entries per bucket (this source)
    unsigned *s_lmark = s_lcnt + P.nb;                                  // [nb] ... before the current level (deferred candidates)
    const int bshift = 32 - P.log_nb;                                   // hash >> bshift = bucket (log_nb >= 1)

    const int tid = threadIdx.x;
    const int lane = gp_lane();
    const long long cta = blockIdx.x;
    const unsigned idmask = P.idbits >= 32 ? 0xFFFFFFFFu : ((1u << P.idbits) - 1u);
    const bool has_code = P.idbits < 32;
    int *pair_id = P.pair_id + cta * P.pair_stride;
    double *pair_val = P.pair_val + cta * P.pair_stride;
    int *log_id = P.log_id + cta * P.log_stride;
    double *log_val = P.log_val + cta * P.log_stride;
    int *push_start = P.push_start + cta * P.capP;
    int *push_len = P.push_len + cta * P.capP;
    double *push_add = P.push_add + cta * P.capP;
    int *sup_id = P.sup_id + cta * P.capS;
    double *sup_val = P.sup_val + cta * P.capS;
    unsigned long long *err = P.stats + 3;
    const unsigned capPair32 = (unsigned)P.capPair, capLog32 = (unsigned)P.capLog;   // (nb * cap < 2^31: plan_bucket)
    const int capC = (int)min((long long)kCandMax, P.capS);   // the candidate list lives in sup_id until the merge

    for (int i = tid; i < kSlots; i += BB) { s_vals[i] = 0.0; s_keys[i] = kEmpty; }
    for (int i = tid; i < P.nb; i += BB) { s_cnt[i] = 0; s_lcnt[i] = 0; }
    if (tid == 0) {
        sm.n_push = 0; sm.n_edges = 0; sm.n_sel = 0; sm.next_item = 0; sm.n_big = 0; sm.ovf = 0; sm.full = 0;
        for (int i = 0; i < 8; i++) sm.ph[i] = 0;
    }
    unsigned long long st_sources = 0, st_redo = 0;      // thread 0
    unsigned long long st_edges = 0, st_frontier = 0, st_support = 0;   // every thread
    const long long t_begin = clock64();
    if (tid == 0) sm.t_prev = t_begin;
#define GPB_PHASE(i) do { if (tid == 0) { const long long t_now = clock64(); sm.ph[i] += t_now - sm.t_prev; sm.t_prev = t_now; } } while (0)

    for (;;) {
        __syncthreads();
        if (tid == 0) sm.it = P.it_base + (long long)atomicAdd(P.queue, 1ull);
        __syncthreads();
        const long long it = sm.it;
        if (it >= P.S) break;
        const int src = P.node_idx[it];
        if (src < 0 || src >= P.n) {   // refuse instead of reading out of bounds
            if (tid == 0) atomicOr(err, kErrBadSource);
            for (int i = tid; i < P.K; i += BB) {
                const long long o = it * P.K + i;
                P.out_row[o] = 0; P.out_col[o] = 0; P.out_val[o] = 0.0;
                if (P.out_val32) P.out_val32[o] = 0.f;
            }
            continue;
        }
        const int2 src_rec = __ldg(P.node_rec + src);
        int src_key = src;
        if (has_code) {
            const unsigned cap = (1u << (31 - P.idbits)) - 1u;
            src_key = (int)((unsigned)src | (min((unsigned)src_rec.y, cap) << P.idbits));
        }
        unsigned src_front = 0;
        unsigned long long src_edges = 0;
        bool ovf = false;
        double tau_lb = 0.0;   // a lower bound of the source's K-th largest reserve, from the levels settled so far (uniform over the CTA)

        // A push-list entry {start, len, add = r / deg}, cut into chunks of kChunk edges (the unit one warp expands).
        auto add_entry = [&](int e_start, unsigned e_len, double e_add) {
            atomicAdd(&sm.n_edges, e_len);
            const unsigned step = e_len > (unsigned)kBigLen ? e_len : (unsigned)kChunk;
            for (unsigned o = 0; o < e_len; o += step) {
                const int p = atomicAdd(&sm.n_push, 1);
                const unsigned l = min(step, e_len - o);
                if (p < kListCap) { sm.start[p] = e_start + (int)o; sm.len[p] = l; sm.add[p] = e_add; }
                else if (p < P.capP) { push_start[p] = e_start + (int)o; push_len[p] = (int)l; push_add[p] = e_add; }
                else ovf = true;
            }
        };
        // Exact push decision of a node whose degree code allows it (graph.h:91-95).
        auto consider = [&](unsigned key, double r) {
            const int2 rec = __ldg(P.node_rec + (key & idmask));
            const unsigned d = (unsigned)rec.y;
            if (d == 0) add_entry(-1, 1u, r);                                      // graph.h:91-93: back to the source
            else if (r >= P.rmax * (double)d) add_entry(rec.x, d, r / (double)d);   // graph.h:94-95
        };
        auto accum4 = [&](const int (&vp)[4], const double (&av)[4], const int max_probe) {   // (vp == kEmpty: no pair)
            // (ONE copy of the find-or-claim + add code per kernel: inlined at its eight call sites it made a third of an
            // 11 K-instruction kernel that stalls on instruction fetch)
            if (accum4_fn<kBuckets4, kHashBits>(s_keys, s_vals, idmask, bshift, max_probe, vp[0], vp[1], vp[2], vp[3], av[0], av[1], av[2], av[3])) {
                ovf = true; sm.full = 1;
            }
        };
        auto accumulate = [&](const int *ids, const double *vals, const unsigned n, const int max_probe) {
            for (unsigned i0 = tid; i0 < n; i0 += BB * 4) {
                int vp[4];
                double av[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const unsigned i = i0 + BB * q;
                    vp[q] = kEmpty; av[q] = 0.0;   // (kEmpty is not a node: the pair is skipped)
                    if (i < n) { vp[q] = __ldcs(ids + i); av[q] = __ldcs(vals + i); }
                }
                if (*(volatile int *)&sm.full) break;
                accum4(vp, av, max_probe);
            }
        };
        // ------------------------------------------------------------------ level 0 (graph.h:80-82)
        const unsigned deg0 = (unsigned)src_rec.y;
        const bool push0 = P.L > 1 && (deg0 == 0 || 1.0 >= P.rmax * (double)deg0);
        if (tid == 0) {
            st_sources++;
            sm.n_push = 0; sm.n_edges = 0;
            if (push0) {
                if (deg0 == 0) add_entry(-1, 1u, 1.0);
                else add_entry(src_rec.x, deg0, 1.0 / (double)deg0);
            }
        }
        if (tid == 0) {   // reserve[src] = coef[0] * 1: the first entry of the log of the source's bucket
            const unsigned b0 = hash_node((unsigned)src) >> bshift;
            log_id[(long long)b0 * P.capLog] = src_key; log_val[(long long)b0 * P.capLog] = P.coef[0];
            s_lcnt[b0] = 1;
            src_front++;
            sup_id[0] = src_key; sm.n_cand = 1;   // candidate merge (below): the source is a candidate
        }
        __syncthreads();
        GPB_PHASE(0);

        for (int level = 0; level < P.L - 1; level++) {   // graph.h:83
            const int n_items = min((long long)sm.n_push, P.capP);
            if (n_items == 0) break;   // nothing pushes: every later residue is zero (the reserve is complete)
            // A level whose pushed edges all fit ONE table fill (the usual case when a source's support is of the order of the
            // table: Reddit-, MAG-shape) skips the bucket streams: the edges are accumulated straight into the table.
            const unsigned level_edges = sm.n_edges;
            const bool direct = level_edges <= (unsigned)kGroupPairs;
            // ---------------------------------------------------------------- expand (graph.h:94-100): append to the buckets
            // Reserve a position in the node's bucket stream / store the pair there: split so that a batch of edges issues all
            // its counter updates before the first dependent store.
            auto reserve = [&](const int vp, const bool ok, unsigned &b) -> unsigned {
                b = hash_node((unsigned)vp & idmask) >> bshift;
                return ok ? atomicAdd(&s_cnt[b], 1u) : 0u;
            };
            auto store = [&](const int vp, const double av, const bool ok, const unsigned b, const unsigned pos) {
                if (ok) {
                    if (pos < (unsigned)P.capPair) { const unsigned o = b * capPair32 + pos; pair_id[o] = vp; pair_val[o] = av; }
                    else ovf = true;
                }
            };
            auto sink = [&](const int vp, const double av, const bool ok) {
                if (direct) {
                    const int v4[4] = {ok ? vp : kEmpty, kEmpty, kEmpty, kEmpty};
                    const double a4[4] = {av, 0.0, 0.0, 0.0};
                    accum4(v4, a4, P.max_probe);
                    return;
                }
                unsigned b;
                const unsigned pos = reserve(vp, ok, b);
                store(vp, av, ok, b, pos);
            };
            auto load_item = [&](const int i, int &st, unsigned &len, double &add) {
                st = 0; len = 0; add = 0.0;
                if (i < n_items) {
                    if (i < kListCap) { st = sm.start[i]; len = sm.len[i]; add = sm.add[i]; }
                    else { st = push_start[i]; len = (unsigned)push_len[i]; add = push_add[i]; }
                }
            };
            // One pushed-edge batch of four slots per lane: straight into the table (direct levels) or into the bucket streams.
            auto push4 = [&](const int (&v4)[4], const double (&a4)[4]) {   // (v4 == kEmpty: no edge)
                if (direct) { accum4(v4, a4, P.max_probe); return; }
                unsigned bk[4], ps[4];
#pragma unroll
                for (int q = 0; q < 4; q++) ps[q] = reserve(v4[q], v4[q] != kEmpty, bk[q]);
#pragma unroll
                for (int q = 0; q < 4; q++) store(v4[q], a4[q], v4[q] != kEmpty, bk[q], ps[q]);
            };
            // A warp takes `grab` push-list entries at a time, one per lane (few entries in the level: two, so that every warp
            // gets work; thousands of them: 32).  Entries of at least 32 edges (chunks of up to 128: four coalesced CSR reads
            // per lane) are expanded one after the other by the whole warp; SHORTER entries -- the common case on graphs with
            // many low-degree nodes, where one warp per entry would leave most lanes idle -- are PACKED: a warp scan of their
            // lengths, then every lane finds the owner of its edge slot by a shuffle binary search over the scan.
            // (measured: where the entries are long -- Reddit-shape: 119 edges on average -- the plain loop below is 5 % faster than the
            // packing loop; on low-degree graphs the packing loop is 40 % faster)
            if (level_edges >= 16u * (unsigned)n_items) {
                constexpr int kItemBatch = 2;
                for (;;) {
                    int i0 = 0;
                    if (lane == 0) i0 = atomicAdd(&sm.next_item, kItemBatch);
                    i0 = __shfl_sync(0xffffffffu, i0, 0);
                    if (i0 >= n_items) break;
                    int st[kItemBatch];
                    unsigned len[kItemBatch];
                    double add[kItemBatch];
    #pragma unroll
                    for (int k = 0; k < kItemBatch; k++) {
                        load_item(i0 + k, st[k], len[k], add[k]);
                        if (lane == 0) src_edges += len[k];   // (src_edges is summed over the lanes at the end)
                        if (len[k] > (unsigned)kChunk) {   // an uncut (long) entry: all warps expand it together after this pass
                            int bpos = 0;
                            if (lane == 0) bpos = atomicAdd(&sm.n_big, 1);
                            bpos = __shfl_sync(0xffffffffu, bpos, 0);
                            if (bpos < kBigCap) {
                                if (lane == 0) { sm.big_st[bpos] = st[k]; sm.big_len[bpos] = len[k]; sm.big_add[bpos] = add[k]; }
                            } else {
                                for (unsigned base = 0; base < len[k]; base += 32) {
                                    const bool ok = base + lane < len[k];
                                    int v = src_key;
                                    if (ok && st[k] >= 0) v = __ldcs(P.packed + st[k] + base + lane);
                                    sink(v, add[k], ok);
                                }
                            }
                            len[k] = 0;
                        }
                    }
                    int vp[kItemBatch][kChunk / 32];
    #pragma unroll
                    for (int k = 0; k < kItemBatch; k++) {
    #pragma unroll
                        for (int q = 0; q < kChunk / 32; q++) {
                            vp[k][q] = src_key;
                            if (st[k] >= 0 && (unsigned)(32 * q + lane) < len[k]) vp[k][q] = __ldcs(P.packed + st[k] + 32 * q + lane);   // graph.h:96-97
                        }
                    }
                    if (direct) {
    #pragma unroll
                        for (int k = 0; k < kItemBatch; k++) {
                            int v4[4];
                            double a4[4];
    #pragma unroll
                            for (int q = 0; q < 4; q++) { v4[q] = (unsigned)(32 * q + lane) < len[k] ? vp[k][q] : kEmpty; a4[q] = add[k]; }
                            if (len[k]) accum4(v4, a4, P.max_probe);   // (warp-uniform)
                        }
                        continue;
                    }
                    unsigned bk[kItemBatch][kChunk / 32], ps[kItemBatch][kChunk / 32];
    #pragma unroll
                    for (int k = 0; k < kItemBatch; k++) {
    #pragma unroll
                        for (int q = 0; q < kChunk / 32; q++) ps[k][q] = reserve(vp[k][q], (unsigned)(32 * q + lane) < len[k], bk[k][q]);
                    }
    #pragma unroll
                    for (int k = 0; k < kItemBatch; k++) {
    #pragma unroll
                        for (int q = 0; q < kChunk / 32; q++) store(vp[k][q], add[k], (unsigned)(32 * q + lane) < len[k], bk[k][q], ps[k][q]);
                    }
                }
            } else {
                const int grab = min(32, max(2, n_items / (BB / 32 * 2)));
                for (;;) {
                    int i0 = 0;
                    if (lane == 0) i0 = atomicAdd(&sm.next_item, grab);
                    i0 = __shfl_sync(0xffffffffu, i0, 0);
                    if (i0 >= n_items) break;
                    int st;
                    unsigned len;
                    double add;
                    load_item(lane < grab ? i0 + lane : n_items, st, len, add);
                    src_edges += len;
                    unsigned long_m = __ballot_sync(0xffffffffu, len > (unsigned)kChunk);   // uncut (long) entries: all warps expand them together after this pass
                    while (long_m) {
                        const int l = __ffs(long_m) - 1;
                        long_m &= long_m - 1;
                        const int b_st = __shfl_sync(0xffffffffu, st, l);
                        const unsigned b_len = __shfl_sync(0xffffffffu, len, l);
                        const double b_add = __shfl_sync(0xffffffffu, add, l);
                        int bpos = 0;
                        if (lane == 0) bpos = atomicAdd(&sm.n_big, 1);
                        bpos = __shfl_sync(0xffffffffu, bpos, 0);
                        if (bpos < kBigCap) {
                            if (lane == 0) { sm.big_st[bpos] = b_st; sm.big_len[bpos] = b_len; sm.big_add[bpos] = b_add; }
                        } else {
                            for (unsigned base = 0; base < b_len; base += 32) {
                                const bool ok = base + lane < b_len;
                                int v = src_key;
                                if (ok && b_st >= 0) v = __ldcs(P.packed + b_st + base + lane);
                                sink(v, b_add, ok);
                            }
                        }
                        if (lane == l) len = 0;
                    }
                    unsigned chunk_m = __ballot_sync(0xffffffffu, len >= 32u);   // chunks: the whole warp per entry, two entries in flight
                    while (chunk_m) {
                        int c_st[2];
                        unsigned c_len[2];
                        double c_add[2];
    #pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const bool have = chunk_m != 0;
                            const int l = have ? __ffs(chunk_m) - 1 : 0;
                            chunk_m &= chunk_m - 1;
                            c_st[k] = __shfl_sync(0xffffffffu, st, l);
                            c_len[k] = __shfl_sync(0xffffffffu, len, l);
                            if (!have) c_len[k] = 0u;
                            c_add[k] = __shfl_sync(0xffffffffu, add, l);
                        }
                        int vp[2][kChunk / 32];
    #pragma unroll
                        for (int k = 0; k < 2; k++) {
    #pragma unroll
                            for (int q = 0; q < kChunk / 32; q++) {
                                vp[k][q] = kEmpty;
                                if ((unsigned)(32 * q + lane) < c_len[k]) vp[k][q] = c_st[k] >= 0 ? __ldcs(P.packed + c_st[k] + 32 * q + lane) : src_key;   // graph.h:96-97
                            }
                        }
    #pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const double a4[4] = {c_add[k], c_add[k], c_add[k], c_add[k]};
                            if (c_len[k]) push4(vp[k], a4);   // (warp-uniform)
                        }
                    }
                    // short entries (fewer than 32 edges, dangling returns included), packed
                    const unsigned s_len = len < 32u ? len : 0u;
                    unsigned incl = s_len;
    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += y;
                    }
                    const unsigned excl = incl - s_len;
                    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
                    for (unsigned base = 0; base < total; base += 128u) {
                        int v4[4];
                        double a4[4];
    #pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const unsigned e = base + 32u * q + lane;
                            // owner = the last lane whose exclusive offset is <= e (zero-length lanes never win: see the scan)
                            int own = 0;
    #pragma unroll
                            for (int step = 16; step >= 1; step >>= 1) {
                                const unsigned x = __shfl_sync(0xffffffffu, excl, own + step);
                                if (x <= e) own += step;
                            }
                            const int o_st = __shfl_sync(0xffffffffu, st, own);
                            const unsigned o_ex = __shfl_sync(0xffffffffu, excl, own);
                            a4[q] = __shfl_sync(0xffffffffu, add, own);
                            v4[q] = kEmpty;
                            if (e < total) v4[q] = o_st >= 0 ? __ldcs(P.packed + o_st + (e - o_ex)) : src_key;
                        }
                        push4(v4, a4);
                    }
                }
            }
            __syncthreads();
            // (reset HERE, one barrier before the settle: a direct level enters its scan without another barrier, and the scan
            // already counts into n_sel / the next push list)
            if (tid == 0) { sm.n_push = 0; sm.n_edges = 0; sm.n_sel = 0; sm.next_item = 0; }
            {
                const int n_big = min(sm.n_big, kBigCap);
                for (int bi = 0; bi < n_big; bi++) {
                    const int st = sm.big_st[bi];
                    const unsigned len = sm.big_len[bi];
                    const double add = sm.big_add[bi];
                    for (unsigned base0 = 0; base0 < len; base0 += BB * 4) {   // (uniform trip count: sink is a warp operation)
                        const unsigned base = base0 + tid;
                        int vp[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const unsigned e = base + BB * q;
                            vp[q] = src_key;
                            if (e < len && st >= 0) vp[q] = __ldcs(P.packed + st + e);
                        }
                        if (direct) {
                            int v4[4];
                            double a4[4];
#pragma unroll
                            for (int q = 0; q < 4; q++) { v4[q] = base + BB * q < len ? vp[q] : kEmpty; a4[q] = add; }
                            accum4(v4, a4, P.max_probe);
                            continue;
                        }
                        unsigned bk[4], ps[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) ps[q] = reserve(vp[q], base + BB * q < len, bk[q]);
#pragma unroll
                        for (int q = 0; q < 4; q++) store(vp[q], add, base + BB * q < len, bk[q], ps[q]);
                    }
                }
                __syncthreads();
            }
            if (tid == 0) sm.n_big = 0;   // (only the next level's expansion touches it again)
            GPB_PHASE(1);
            // ---------------------------------------------------------------- settle of level + 1, bucket by bucket
            const int nl = level + 1;
            const bool will_push = nl < P.L - 1;
            const double c = P.coef[nl];
            // Candidate merge (see the merge below): a node can only be among the K largest if ONE of its <= L contributions
            // coef * r is >= (K-th largest reserve) / L.  tau_lb is a lower bound of that reserve, so every contribution
            // >= tau_lb / L notes its node as a candidate (while tau_lb is 0: every node).
            const double cand_thr = tau_lb / (double)P.L * (1.0 - 1e-9);
            // A large level that arrives before any bound exists (a source with fewer than K neighbours whose neighbours have
            // thousands) would note every node: its candidates are picked from its log entries AFTER the level has set tau_lb.
            unsigned level_pairs = level_edges;
            if (!direct) { level_pairs = 0; for (int b = 0; b < P.nb; b++) level_pairs += s_cnt[b]; }
            const bool defer = !P.full_merge && tau_lb == 0.0 && level_pairs > 2048u;
            if (defer) {   // (warp-uniform, rare)
                if (tid < P.nb) s_lmark[tid] = min(s_lcnt[tid], (unsigned)P.capLog);
                // a DIRECT level goes straight into the scan: without this barrier a fast warp appends to the logs before a slow
                // thread has taken its mark, the deferred pick then skips those entries and their nodes never become candidates
                __syncthreads();
            }
            long long lvl_max = 0;   // this thread's largest contribution of the level (a node of its own: a level's nodes are distinct)
            for (int b = 0; b < P.nb;) {
                // one visit = consecutive buckets [b, e) whose pairs together cannot overfill the table (pairs bound the
                // distinct nodes); a bucket above the limit is a visit of its own.  (s_cnt is stable since the barrier that
                // ended the expansion, so every thread forms the same groups.)
                int e = P.nb;   // (direct: the table already holds the whole level, one visit over all buckets)
                if (!direct) {
                    unsigned tot = min(s_cnt[b], (unsigned)P.capPair);
                    e = b + 1;
                    while (e < P.nb && tot + min(s_cnt[e], (unsigned)P.capPair) <= (unsigned)kGroupPairs) { tot += min(s_cnt[e], (unsigned)P.capPair); e++; }
                    if (tot == 0) { b = e; continue; }
                    for (int bb = b; bb < e; bb++)
                        accumulate(pair_id + (long long)bb * P.capPair, pair_val + (long long)bb * P.capPair, min(s_cnt[bb], (unsigned)P.capPair), P.max_probe);
                    __syncthreads();
                }
                const bool multi = e - b > 1;
                // settle: every thread scans its slots and the warp COMPACTS the occupied ones into a queue (a table filled to a
                // third would otherwise walk the body below sixteen times per warp with a third of the lanes); 32 queued slots at
                // a time: reserve += coef * r (graph.h:90 / :106) goes to the log of the node's bucket (one counter update per
                // warp and bucket), the push decision comes from the degree code in the key, the slot is emptied
                unsigned short *wq = sm.wq[tid >> 5];
                int qn = 0;
                auto settle32 = [&](const int off, const int cnt) {   // queue entries [off, off + cnt), cnt <= 32
                    const bool have = lane < cnt;
                    double rr = 0.0;
                    unsigned key = 0u;
                    if (have) {
                        const int slot = wq[off + lane];
                        rr = s_vals[slot]; key = (unsigned)s_keys[slot];
                        s_vals[slot] = 0.0; s_keys[slot] = kEmpty;   // the table is empty again for the next visit
                    }
                    unsigned lb = (unsigned)b, pos;
                    if (!multi) {
                        unsigned base = 0;
                        if (lane == 0) base = atomicAdd(&s_lcnt[b], (unsigned)cnt);
                        pos = __shfl_sync(0xffffffffu, base, 0) + lane;
                    } else {
                        lb = have ? hash_node(key & idmask) >> bshift : 0xFFFFFFFFu;
                        const unsigned peers = __match_any_sync(0xffffffffu, lb);
                        const int leader = __ffs(peers) - 1;
                        unsigned p0 = 0;
                        if (have && lane == leader) p0 = atomicAdd(&s_lcnt[lb], (unsigned)__popc(peers));
                        pos = __shfl_sync(0xffffffffu, p0, leader) + __popc(peers & ((1u << lane) - 1u));
                    }
                    if (have) {
                        src_front++;
                        const double cr = c * rr;
                        if (pos < (unsigned)P.capLog) { const unsigned o = lb * capLog32 + pos; log_id[o] = (int)key; log_val[o] = cr; }
                        else ovf = true;
                        lvl_max = max(lvl_max, __double_as_longlong(cr));
                        if (!defer && cr >= cand_thr) {   // (rare once the first two or three levels have set tau_lb)
                            const int p = atomicAdd(&sm.n_cand, 1);
                            if (p < capC) sup_id[p] = (int)key;
                        }
                        if (will_push) {
                            const unsigned code = has_code ? key >> P.idbits : 0u;   // min(deg, cap): a lower bound of deg
                            if (rr >= P.rmax * (double)code) {                       // necessary for graph.h:94; the exact test follows
                                const int p = atomicAdd(&sm.n_sel, 1);
                                if (p < kCandCap) { sm.cand.key[p] = key; sm.cand.r[p] = rr; }
                                else consider(key, rr);
                            }
                        }
                    }
                };
                const unsigned lt = (1u << lane) - 1u;
#pragma unroll 2
                for (int j = 0; j < SPT / 2; j++) {
                    const int slot = 2 * (j * BB + tid);
                    const double2 r2 = *reinterpret_cast<const double2 *>(s_vals + slot);
                    const bool got0 = r2.x != 0.0, got1 = r2.y != 0.0;
                    const unsigned m0 = __ballot_sync(0xffffffffu, got0), m1 = __ballot_sync(0xffffffffu, got1);
                    if (m0) {   // (warp-uniform; qn < 32 here, the queue holds 64)
                        if (got0) wq[qn + __popc(m0 & lt)] = (unsigned short)slot;
                        qn += __popc(m0);
                        __syncwarp();
                        if (qn >= 32) { qn -= 32; settle32(qn, 32); __syncwarp(); }
                    }
                    if (m1) {
                        if (got1) wq[qn + __popc(m1 & lt)] = (unsigned short)(slot + 1);
                        qn += __popc(m1);
                        __syncwarp();
                        if (qn >= 32) { qn -= 32; settle32(qn, 32); __syncwarp(); }
                    }
                }
                if (qn) settle32(0, qn);
                __syncthreads();
                if (tid < e - b) s_cnt[b + tid] = 0;   // (sm.full stays set until the source ends: it is handed over anyway)
                b = e;
            }
            __syncthreads();
            if (will_push) {
                const int n_sel = min(sm.n_sel, kCandCap);
                for (int i = tid; i < n_sel; i += BB) consider(sm.cand.key[i], sm.cand.r[i]);
            }
            __syncthreads();
            if (!P.full_merge && P.K <= BB) {
                // tau_lb: the K-th largest of the threads' largest contributions of this level -- K different nodes whose
                // reserve is at least that (a level's nodes are distinct, contributions only add up)
                const double lvl_bound = block_kth_lower_bound<BB>(sm, P.K, lvl_max);   // (one histogram pass: within 1/32 of the K-th largest)
                tau_lb = fmax(tau_lb, lvl_bound);   // (the same value in every thread: nothing to publish, no barrier)
                if (defer) {   // this level's log entries (just written: L2) against the bound the level itself gave
                    const double thr = tau_lb / (double)P.L * (1.0 - 1e-9);
                    for (int b = 0; b < P.nb; b++) {
                        const unsigned hi = min(s_lcnt[b], (unsigned)P.capLog);
                        for (unsigned i = s_lmark[b] + tid; i < hi; i += BB) {
                            if (log_val[(long long)b * P.capLog + i] >= thr) {
                                const int p = atomicAdd(&sm.n_cand, 1);
                                if (p < capC) sup_id[p] = log_id[(long long)b * P.capLog + i];
                            }
                        }
                    }
                }
            }
            GPB_PHASE(2);
        }
        if (ovf) sm.ovf = 1;
        __syncthreads();
        const bool redo = sm.ovf != 0;
        // ------------------------------------------------------------------ reserve: merge the logs bucket by bucket
        // Only reserves that can still be among the K largest are written out: a running threshold tau_run with at least K
        // reserves >= it.  Every warp publishes the r-th largest of its lanes' maxima, r = ceil(K / warps) (a lane's maximum is a
        // node of its own, so the warp has seen r reserves >= that value); the minimum over the warps bounds K of them.  The
        // published values only grow, so a reader that sees a mix of old and new ones still holds a valid (lower) bound.
        const int rth = (P.K + BB / 32 - 1) / (BB / 32);
        if (tid == 0) {
            sm.n_sup = 0; sm.n_all = 0;
        }
        if (lane == 0) sm.wtau[tid >> 5] = 0.0;
        __syncthreads();
        long long m1x = 0;   // largest reserve this thread produced (non-negative doubles order like their bit patterns)
        unsigned src_support = 0;
        // Writes the table's nodes with a reserve >= tau_run to the compact arrays and empties the table.
        auto drain = [&](const double tau_run) {
#pragma unroll 2
            for (int j = 0; j < SPT / 2; j++) {
                const int slot = 2 * (j * BB + tid);
                const int2 k2 = *reinterpret_cast<const int2 *>(s_keys + slot);
                const bool got0 = k2.x != kEmpty, got1 = k2.y != kEmpty;
                if (__any_sync(0xffffffffu, got0 | got1)) {
                    double2 x2 = make_double2(0.0, 0.0);
                    if (got0 | got1) {
                        x2 = *reinterpret_cast<const double2 *>(s_vals + slot);
                        *reinterpret_cast<double2 *>(s_vals + slot) = make_double2(0.0, 0.0);
                        *reinterpret_cast<int2 *>(s_keys + slot) = make_int2(kEmpty, kEmpty);
                    }
                    const bool keep0 = got0 && x2.x >= tau_run, keep1 = got1 && x2.y >= tau_run;
                    const unsigned m0 = __ballot_sync(0xffffffffu, keep0), m1 = __ballot_sync(0xffffffffu, keep1);
                    if (m0 | m1) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&sm.n_sup, __popc(m0) + __popc(m1));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        const unsigned lt = (1u << lane) - 1u;
                        const long long p0 = base + __popc(m0 & lt), p1 = base + __popc(m0) + __popc(m1 & lt);
                        if (keep0 && p0 < P.capS) { sup_id[p0] = (int)((unsigned)k2.x & idmask); sup_val[p0] = x2.x; }
                        if (keep1 && p1 < P.capS) { sup_id[p1] = (int)((unsigned)k2.y & idmask); sup_val[p1] = x2.y; }
                    }
                    if (got0) { src_support++; m1x = max(m1x, __double_as_longlong(x2.x)); }
                    if (got1) { src_support++; m1x = max(m1x, __double_as_longlong(x2.y)); }
                }
            }
        };
        // CANDIDATE MERGE (default).  The K largest reserves are wanted, not the whole reserve vector: only the candidates
        // noted during the levels (a few hundred to a few thousand nodes against a support of 10^5) get a table slot, then
        // the logs stream through ONCE with a read-only lookup per entry -- almost always a miss decided by one 16-byte
        // shared-memory read, no claim, no add -- and one table scan writes the candidates' reserves out.  The support itself
        // is never materialised (gp_push_stats.support_total does not count these sources).  Falls back to the full merge
        // when the candidates outgrow half a table, when K > 1024, or when P.full_merge asks for the support count.
        const bool cand_merge = !redo && !P.full_merge && P.K <= BB && sm.n_cand <= capC;
        if (cand_merge) {
            const int n_cand = sm.n_cand;
            for (int i = tid; i < n_cand; i += BB) {
                const int vp = sup_id[i];
                bool claimed;
                find_slot<kBuckets4>(s_keys, (hash_node((unsigned)vp & idmask) >> (bshift - kHashBits)) & (kBuckets4 - 1), vp, kBuckets4, claimed);
            }
            __syncthreads();
            for (int b = 0; b < P.nb; b++) {
                const unsigned n = min(s_lcnt[b], (unsigned)P.capLog);
                const int *lid = log_id + (long long)b * P.capLog;
                const double *lval = log_val + (long long)b * P.capLog;
                for (unsigned i0 = tid; i0 < n; i0 += BB * 4) {
                    int vp[4], at[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) vp[q] = i0 + BB * q < n ? __ldcs(lid + i0 + BB * q) : kEmpty;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        at[q] = -1;
                        if (vp[q] != kEmpty) at[q] = lookup_slot<kBuckets4>(s_keys, (hash_node((unsigned)vp[q] & idmask) >> (bshift - kHashBits)) & (kBuckets4 - 1), vp[q]);
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (at[q] >= 0) atomicAdd(s_vals + at[q], __ldcs(lval + i0 + BB * q));   // (the value is only read for a candidate)
                }
            }
            __syncthreads();
            drain(0.0);
            __syncthreads();
            if (tid < P.nb) s_lcnt[tid] = 0;
            src_support = 0;   // (candidates, not the support)
        } else
        for (int b = 0; b < P.nb;) {
            unsigned tot = min(s_lcnt[b], (unsigned)P.capLog);
            int e = b + 1;
            while (e < P.nb && tot + min(s_lcnt[e], (unsigned)P.capLog) <= (unsigned)kGroupPairs) { tot += min(s_lcnt[e], (unsigned)P.capLog); e++; }
            if (tot != 0 && !redo) {
                // (read before the barrier below: the warps publish their next values right after their drain)
                double tau_run = sm.wtau[0];
#pragma unroll 8
                for (int w = 1; w < BB / 32; w++) tau_run = fmin(tau_run, sm.wtau[w]);
                // (the table held these nodes at every level, or they are fewer than kGroupPairs: it cannot overflow)
                for (int bb = b; bb < e; bb++)
                    accumulate(log_id + (long long)bb * P.capLog, log_val + (long long)bb * P.capLog, min(s_lcnt[bb], (unsigned)P.capLog), kBuckets4);
                __syncthreads();
                // every claimed slot is one node of the support (its reserve may be exactly 0.0 in `single` mode)
                drain(tau_run);
                if (rth <= 32) {   // the warp's rth-largest lane maximum (0.0 while fewer than rth lanes hold a positive reserve)
                    const long long r_val = warp_rth_largest(m1x, rth);
                    if (lane == 0) sm.wtau[tid >> 5] = __longlong_as_double(r_val);
                }
                __syncthreads();
            }
            if (tid < e - b) s_lcnt[b + tid] = 0;
            b = e;
        }
        {
            const unsigned ws = __reduce_add_sync(0xffffffffu, src_support);
            if (lane == 0 && ws) atomicAdd(&sm.n_all, (int)ws);
        }
        __syncthreads();
        if (tid == 0 && sm.n_all) atomicMax(P.max_support, (unsigned long long)sm.n_all);   // (full merge only: sizes the buckets of later calls)
        GPB_PHASE(3);
        // ------------------------------------------------------------------ top-k, graph.h:111-126
        const int n_sup = (int)min((long long)sm.n_sup, P.capS);
        if (!redo) {
            st_frontier += src_front; st_edges += src_edges;
            st_support += src_support;
            // Threshold: the K-th largest of the threads' maxima -- every thread's maximum is a distinct node's reserve, so at
            // least K reserves are >= tau and nothing below tau can be among the K largest (K <= threads); any lower bound of
            // it serves as well.
            const double tau = block_kth_lower_bound<BB>(sm, P.K, m1x);   // (a lower bound within 1/32 of it: a few more survivors)
            if (tid == 0) sm.n_list = 0;
            __syncthreads();
            bool listed = tau > 0.0;
            if (listed) {
                for (int j0 = tid; j0 < n_sup; j0 += 4 * BB) {
                    double x[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) x[q] = j0 + q * BB < n_sup ? sup_val[j0 + q * BB] : 0.0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (x[q] >= tau) {
                            const int pos = atomicAdd(&sm.n_list, 1);
                            if (pos < kListCap) { sm.add[pos] = x[q]; sm.start[pos] = sup_id[j0 + q * BB]; }
                        }
                    }
                }
                __syncthreads();
                listed = sm.n_list <= kListCap;
            }
            const int n_list = listed ? sm.n_list : 0;
            auto each = [&](auto f) {
                if (listed) {
                    for (int i = tid; i < n_list; i += BB) f(sm.add[i], sm.start[i]);
                } else {
                    for (int j = tid; j < n_sup; j += BB) {
                        const double x = sup_val[j];
                        if (x > 0.0) f(x, sup_id[j]);
                    }
                }
            };
            const int n = block_topk<BB>(sm, P.K, listed && n_list <= kBucketCap, each, [&](int pos, int id, double v) {
                const long long o = it * P.K + pos;
                P.out_row[o] = src; P.out_col[o] = id; P.out_val[o] = v;
                if (P.out_val32) P.out_val32[o] = (float)v;
            });
            // unfilled slots read (0, 0, 0.0): what graph.h:117-126 leaves in the caller-zeroed arrays
            for (int i = n + tid; i < P.K; i += BB) {
                const long long o = it * P.K + i;
                P.out_row[o] = 0; P.out_col[o] = 0; P.out_val[o] = 0.0;
                if (P.out_val32) P.out_val32[o] = 0.f;
            }
        } else if (tid == 0) {
            P.redo[atomicAdd(P.redo_count, 1ull)] = (int)it; st_redo++; st_sources--;
        }
        if (tid == 0) { sm.ovf = 0; sm.full = 0; sm.n_push = 0; sm.n_edges = 0; sm.n_sel = 0; }
        GPB_PHASE(4);
    }
    // counters
    for (int o = 16; o >= 1; o >>= 1) {
        st_frontier += __shfl_xor_sync(0xffffffffu, st_frontier, o); st_support += __shfl_xor_sync(0xffffffffu, st_support, o);
        st_edges += __shfl_xor_sync(0xffffffffu, st_edges, o);
    }
    if (lane == 0) {
        if (st_support) { atomicAdd(P.stats + 2, st_support); atomicAdd(P.cum + 2, st_support); }
        if (st_edges) { atomicAdd(P.stats + 0, st_edges); atomicAdd(P.cum + 0, st_edges); }
        if (st_frontier) { atomicAdd(P.stats + 1, st_frontier); atomicAdd(P.cum + 1, st_frontier); }
    }
    if (tid == 0) {
        sm.ph[7] = clock64() - t_begin;
        for (int i = 0; i < 8; i++) atomicAdd(P.phase + i, (unsigned long long)sm.ph[i]);
        atomicAdd(P.cum + 3, st_sources);
        atomicAdd(P.cum + 4, st_sources);   // "cluster_sources": sources finished by the first-pass kernel
        atomicAdd(P.cum + 5, st_redo);
    }
}
#undef GPB_PHASE

}  // namespace

size_t gpb_dynamic_smem(int nb, int block) { return (size_t)gpb_slots(block) * 12 + (size_t)nb * 12; }

namespace {
template <int BB>
int configure_kernel(size_t smem) {
    static size_t configured = 0;
    if (smem > configured) {
        GP_CUDA_TRY(cudaFuncSetAttribute(gfpush_bucket_kernel<BB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GP_CUDA_TRY(cudaFuncSetAttribute(gfpush_bucket_kernel<BB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    return GP_OK;
}
template <int BB>
int resident_ctas(int nb, int *per_sm) {
    const size_t smem = gpb_dynamic_smem(nb, BB);
    int rc = configure_kernel<BB>(smem);
    if (rc != GP_OK) return rc;
    int n = 0;
    GP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gfpush_bucket_kernel<BB>, BB, smem));
    *per_sm = n < 1 ? 1 : n;
    return GP_OK;
}
template <int BB>
int launch(const BucketPushParams &P, int ctas, cudaStream_t stream) {
    const size_t smem = gpb_dynamic_smem(P.nb, BB);
    int rc = configure_kernel<BB>(smem);
    if (rc != GP_OK) return rc;
    gfpush_bucket_kernel<BB><<<(unsigned)ctas, BB, smem, stream>>>(P);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}
}  // namespace

int gpb_ctas_per_sm(int nb, int block, int *per_sm) {
    return block == 256 ? resident_ctas<256>(nb, per_sm) : block == 512 ? resident_ctas<512>(nb, per_sm) : resident_ctas<1024>(nb, per_sm);
}

int gpb_launch(const BucketPushParams &P, int ctas, cudaStream_t stream) {
    return P.block == 256 ? launch<256>(P, ctas, stream) : P.block == 512 ? launch<512>(P, ctas, stream) : launch<1024>(P, ctas, stream);
}

}  // namespace gpp
