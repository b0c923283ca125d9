// gfpush_cluster.cu -- GFPush + top-k with ONE SOURCE PER THREAD-BLOCK CLUSTER (sm_100a).
// Same computation as gfpush.cu (Graph::gfpush_omp, /root/reference/precompute/graph.h:53-131); this kernel is the
// path for supports that outgrow one SM's shared memory (Reddit-, MAG-, Amazon2M-shape graphs):
//
//   * a cluster of G CTAs (1, 2, 4, 8 or 16) owns a source.  Node v belongs to CTA  hash(v) >> (32 - log2 G);  every
//     CTA keeps an open-addressed {key, residue} table of 16 384 slots in its shared memory, PERSISTENT for the source,
//     so the cluster offers G x 16 384 slots and the residues never leave the chip;
//   * the reserve of a slot lives in a REGISTER of the thread that owns the slot (32 slots per thread, 512 threads):
//     "reserve += coef * r" (graph.h:90) is one FMA, nothing is logged, merged or read-modify-written in memory, and
//     the top-k selects straight out of the register file;
//   * a level is   settle  ->  expand  ->  exchange:
//       settle   every thread scans its 32 slots: a non-zero residue is credited to the reserve and zeroed; whether the
//                node pushes (r >= rmax * deg, graph.h:94) is decided from a DEGREE CODE packed into the spare high
//                bits of every CSR entry (and therefore of every table key): the {start, degree} record of a node is
//                only fetched for the ~2 % of nodes that can pass -- no list of touched nodes, no first-touch
//                detection, no per-node global access;
//       expand   edge-balanced over the CTA's push list (block scan of the degrees + owner search, coalesced reads of
//                the packed CSR); with G == 1 every edge is find-or-claim + fp64 add in the local table, with G > 1
//                every edge (node, r/deg) is APPENDED to the stream of the CTA that owns the node (one shared-memory
//                counter per destination, one atomic per warp and destination via match.any);
//       exchange after a cluster barrier every CTA reads the G streams addressed to it (L2-resident, coalesced) and
//                accumulates them into its table.  Entries of degree >= hub_min_deg are put on a cluster-wide list and
//                every CTA expands a 1/G slice, so a 240 K-degree hub does not serialise on its owner;
//   * top-k: every CTA radix-selects its own K largest reserves (gfpush.cu's MSD select, reading registers), writes
//     them to a candidate buffer, and the cluster's first CTA selects the K largest of the G x K candidates;
//   * a source whose table or stream overflows is handed to gfpush.cu's slab kernel through a redo list.
#include "gfpush_cluster.h"
#include "gfpush_shared.cuh"

#include <cooperative_groups.h>

#include <algorithm>

namespace cg = cooperative_groups;

namespace gpp {
namespace {

constexpr int CB = kClusterBlock;
constexpr int SPT = kClusterSlots / CB;       // table slots (and reserve registers) per thread
constexpr int kBuckets = kClusterSlots / 4;   // 4-key buckets: one 16-byte shared-memory read per probe
constexpr int kEmpty = -1;
constexpr int kEdgeUnroll = 4;

constexpr int kBigLen = 2048;    // entries longer than this are expanded by all warps of the CTA together
constexpr int kBigCap = 32;
constexpr int kItemBatch = 2;    // push-list entries a warp expands together
constexpr int kChunk = 128;      // edges per push-list entry (longer entries are cut when they are appended)
constexpr int kCandCap = 1024;   // nodes per level (and CTA) whose degree code allows a push; more are handled inline

struct CSmem {
    unsigned off[CB + 1];      // exclusive scan of the tile's degrees, off[CB] = total
    int start[CB];
    double add[CB];
    unsigned warp_scan[CB / 32 + 1];
    double wtau[CB / 32];      // top-k pre-filter: per-warp lower bounds of the K-th largest reserve
    union {
        struct {               // top-k
            unsigned hist[kHistBins];
            unsigned long long bkey[kBucketCap];
            int bid[kBucketCap];
        } sel;
        struct {               // settle: candidates for a push
            double r[kCandCap];
            unsigned key[kCandCap];
        } cand;
    };
    unsigned cnt[kClusterMaxG];      // entries this CTA appended to the stream of each destination (this level)
    unsigned inbox[2][kClusterMaxG];    // [level parity] entries every sender appended to MY stream (written by the senders before the barrier)
    unsigned pushed[2][kClusterMaxG];   // [level parity] push-list entries every CTA expanded at this level
    unsigned pre[kClusterMaxG + 1];  // exchange / final select: prefix sums over the senders
    long long seg_off[kClusterMaxG]; // exchange: element offset of every sender's stream to me, minus its prefix
    long long it;                    // (first CTA) the cluster's current source
    int n_push;                      // local push-list entries of this level
    int n_sel;                       // settle candidates
    int n_out, n_bucket, n_cand;
    int sel_bin, sel_above, sel_inbin;
    int seed_slot;
    int next_item;                   // expand: next push-list entry to hand to a warp
    int n_big;                       // expand: entries left for the all-warps pass
    int big_st[kBigCap];
    unsigned big_len[kBigCap];
    double big_add[kBigCap];
    int n_list;                      // top-k: reserves above the pre-filter threshold, compacted into add[] / start[]
    int full;                        // a probe sequence ran out: this source goes to the slab kernel, stop probing
    // cluster-wide state, valid in the first CTA's copy
    unsigned c_push[2];              // push-list entries of all CTAs, by level parity
    unsigned c_hub[2];               // hub entries, by level parity
    unsigned c_flags;                // 1 = hand the source to the slab kernel
    long long ph[8], t_prev;
};

__device__ __forceinline__ unsigned hash_node(unsigned id) { return id * 2654435761u; }

template <int G>
struct Log2 { static constexpr int v = 1 + Log2<G / 2>::v; };
template <>
struct Log2<1> { static constexpr int v = 0; };

// Slot of packed node `vp` in the table (claiming one when it is new), or -1 when `max_probe` buckets hold neither it
// nor an empty slot.  Keys are never removed while a source is live and empties are taken in index order, so an
// observed key is final and a node can never end up in two slots.
__device__ __forceinline__ int find_slot(int *keys, unsigned bucket, int vp, int max_probe) {
    unsigned b = bucket;
    for (int probe = 0; probe < max_probe; probe++, b = (b + 1) & (kBuckets - 1)) {
        const int4 k4 = *reinterpret_cast<const int4 *>(keys + 4 * b);
        const int kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int k = kk[i];
            if (k == kEmpty) {
                k = atomicCAS(keys + 4 * b + i, kEmpty, vp);
                if (k == kEmpty) return (int)(4 * b + i);
            }
            if (k == vp) return (int)(4 * b + i);
        }
    }
    return -1;
}

template <int G>
__global__ void __launch_bounds__(CB, 1) gfpush_cluster_kernel(const ClusterPushParams P) {
    constexpr int LOGG = Log2<G>::v;
    constexpr int kOwnerShift = G > 1 ? 32 - LOGG : 31;   // hash >> kOwnerShift = owning CTA (unused when G == 1)
    constexpr int kBucketShift = 32 - LOGG - 12;          // the 12 hash bits below the owner bits pick the bucket
    // Long entries are shared by the whole cluster (a second cluster barrier per level) only when the cluster is large
    // enough for one CTA's hub to matter; small clusters expand their own entries and pay one barrier per level.
    constexpr bool kShareHubs = G >= 4;
    __shared__ CSmem sm;
    extern __shared__ double s_vals[];                      // [kClusterSlots] next-level residues
    int *s_keys = reinterpret_cast<int *>(s_vals + kClusterSlots);   // [kClusterSlots] packed node, kEmpty = free

    const int tid = threadIdx.x;
    const int lane = gp_lane();
    unsigned rank = 0;
    if (G > 1) rank = cg::this_cluster().block_rank();
    const long long cta = blockIdx.x;
    const long long cta0 = cta - rank;          // the cluster's first CTA
    const long long cid = cta / G;
    CSmem *ldr = &sm;                            // the first CTA's shared state
    if (G > 1) ldr = cg::this_cluster().map_shared_rank(&sm, 0);
    auto csync = [&]() { if (G > 1) cg::this_cluster().sync(); else __syncthreads(); };

    const unsigned idmask = P.idbits >= 32 ? 0xFFFFFFFFu : ((1u << P.idbits) - 1u);
    const bool has_code = P.idbits < 32;
    int *push_start = P.push_start + cta * P.capP;
    int *push_len = P.push_len + cta * P.capP;
    double *push_add = P.push_add + cta * P.capP;
    int *hub_start = P.hub_start + cid * P.capHub;
    int *hub_deg = P.hub_deg + cid * P.capHub;
    double *hub_add = P.hub_add + cid * P.capHub;
    // my streams, one per destination, in two generations (level parity): without a barrier between settle and expand
    // a fast CTA already writes level l+1 while a slow one still reads level l
    int *x_id_base = P.x_id + cta * 2 * G * P.capX;
    double *x_val_base = P.x_val + cta * 2 * G * P.capX;
    unsigned long long *err = P.stats + 3;

    for (int i = tid; i < kClusterSlots; i += CB) { s_vals[i] = 0.0; s_keys[i] = kEmpty; }
    double rsv[SPT];                             // reserve of slot j * CB + tid
#pragma unroll
    for (int j = 0; j < SPT; j++) rsv[j] = 0.0;
    if (tid < kClusterMaxG) sm.cnt[tid] = 0;
    if (tid == 0) {
        sm.c_push[0] = sm.c_push[1] = 0; sm.c_hub[0] = sm.c_hub[1] = 0; sm.c_flags = 0; sm.n_push = 0; sm.n_sel = 0; sm.full = 0; sm.next_item = 0; sm.n_big = 0;
        for (int i = 0; i < 8; i++) sm.ph[i] = 0;
    }
    unsigned long long st_sources = 0, st_cluster = 0, st_redo = 0;   // thread 0 only
    unsigned long long st_edges = 0;                                    // lane 0 of every warp
    unsigned st_frontier = 0, st_support = 0;                           // every thread, reduced at the end
    const long long t_begin = clock64();
    if (tid == 0) sm.t_prev = t_begin;
#define GPC_PHASE(i) do { if (tid == 0) { const long long t_now = clock64(); sm.ph[i] += t_now - sm.t_prev; sm.t_prev = t_now; } } while (0)

    for (;;) {
        if (rank == 0 && tid == 0) sm.it = (long long)atomicAdd(P.queue, 1ull);
        csync();   // #B
        const long long it = ldr->it;
        if (it >= P.S) break;
        const int src = P.node_idx[it];
        if (src < 0 || src >= P.n) {   // refuse instead of reading out of bounds
            if (rank == 0) {
                if (tid == 0) atomicOr(err, kErrBadSource);
                for (int i = tid; i < P.K; i += CB) {
                    const long long o = it * P.K + i;
                    P.out_row[o] = 0; P.out_col[o] = 0; P.out_val[o] = 0.0;
                    if (P.out_val32) P.out_val32[o] = 0.f;
                }
            }
            csync();   // the next fetch must not overwrite sm.it before everyone has read it
            continue;
        }
        const int2 src_rec = __ldg(P.node_rec + src);
        unsigned src_front = 0;                 // work counters of this source (dropped when it is handed over)
        unsigned long long src_edges = 0;       // (lane 0 of every warp counts the entries it expanded)
        int src_packed = src;
        if (has_code) {
            const unsigned cap = (1u << (31 - P.idbits)) - 1u;
            src_packed = (int)((unsigned)src | (min((unsigned)src_rec.y, cap) << P.idbits));
        }
        bool ovf = false;
        // A push-list entry {start, len, add}: the first CB of a level stay in the tile arrays of the expansion.
        auto add_entry = [&](int e_start, unsigned e_len, double e_add, int par) {
            if (kShareHubs && e_len >= (unsigned)P.hub_min_deg) {
                const unsigned p = atomicAdd(&ldr->c_hub[par], 1u);
                if (p < (unsigned)P.capHub) { hub_start[p] = e_start; hub_deg[p] = (int)e_len; hub_add[p] = e_add; }
                else ovf = true;
            } else {
                // entries are cut into chunks of kChunk edges -- the unit a warp expands in one go, so the level balances
                // over the warps like an edge-parallel expansion without a search for the owner of an edge; entries
                // beyond kBigLen stay whole and are expanded by all warps together
                const unsigned step = e_len > (unsigned)kBigLen ? e_len : (unsigned)kChunk;
                for (unsigned o = 0; o < e_len; o += step) {
                    const int p = atomicAdd(&sm.n_push, 1);
                    const unsigned l = min(step, e_len - o);
                    if (p < CB) { sm.start[p] = e_start + (int)o; sm.off[p] = l; sm.add[p] = e_add; }
                    else if (p < P.capP) { push_start[p] = e_start + (int)o; push_len[p] = (int)l; push_add[p] = e_add; }
                    else ovf = true;
                }
            }
        };
        // Exact push decision of a node whose degree code allows it (graph.h:91-95); fetches its {start, degree} record.
        auto consider = [&](unsigned key, double r, int par) {
            const int2 rec = __ldg(P.node_rec + (key & idmask));
            const unsigned d = (unsigned)rec.y;
            if (d == 0) add_entry(-1, 1u, r, par);                                     // graph.h:91-93: back to the source
            else if (r >= P.rmax * (double)d) add_entry(rec.x, d, r / (double)d, par);   // graph.h:94-95
        };
        // ------------------------------------------------------------------ level 0 (graph.h:80-82)
        // residue = {src: 1}: the reserve of the source's slot gets coef[0], and the source's adjacency is expanded by
        // the whole cluster (every CTA takes a 1/G slice) without a settle pass or a barrier.
        const unsigned hsrc = hash_node((unsigned)src);
        const bool mine = (G > 1 ? hsrc >> kOwnerShift : 0u) == rank;
        const unsigned deg0 = (unsigned)src_rec.y;
        const bool push0 = P.L > 1 && (deg0 == 0 || 1.0 >= P.rmax * (double)deg0);
        if (tid == 0) {
            if (rank == 0) { st_sources++; st_cluster++; }
            if (mine) sm.seed_slot = find_slot(s_keys, (hsrc >> kBucketShift) & (kBuckets - 1), src_packed, 1);   // empty table
            sm.n_push = 0;
            if (push0) {
                if (deg0 == 0) {
                    if (rank == 0) { sm.start[0] = -1; sm.off[0] = 1u; sm.add[0] = 1.0; sm.n_push = 1; }
                } else {
                    const unsigned lo = (unsigned)((unsigned long long)deg0 * rank / G), hi = (unsigned)((unsigned long long)deg0 * (rank + 1) / G);
                    const unsigned l0 = hi - lo;
                    const unsigned step = l0 > (unsigned)kBigLen ? l0 : (unsigned)kChunk;
                    int n0 = 0;
                    for (unsigned o = 0; o < l0; o += step, n0++) {   // (at most kBigLen / kChunk chunks, all in the tile arrays)
                        sm.start[n0] = src_rec.x + (int)(lo + o); sm.off[n0] = min(step, l0 - o); sm.add[n0] = 1.0 / (double)deg0;
                    }
                    sm.n_push = n0;
                }
            }
        }
        __syncthreads();
        if (mine) {
            const int slot = sm.seed_slot;
            if ((slot & (CB - 1)) == tid) {
                const double c0 = P.coef[0];
#pragma unroll
                for (int j = 0; j < SPT; j++)
                    rsv[j] += j == slot / CB ? c0 : 0.0;   // graph.h:90 at level 0 (written so that rsv[] stays in registers)
                src_front++;
            }
        }
        GPC_PHASE(0);

        for (int level = 0; level < P.L - 1; level++) {   // graph.h:83
            const int par = level & 1;
            int n_local, n_hub = 0;
            if (level == 0) {
                if (!push0) break;
                n_local = sm.n_push;
            } else if (kShareHubs) {
                if (tid == 0 && sm.n_push) atomicAdd(&ldr->c_push[par], (unsigned)sm.n_push);
                csync();   // #1: every CTA's push list and the hub list of this level are complete
                n_local = min((long long)sm.n_push, P.capP);
                n_hub = (int)min(ldr->c_hub[par], (unsigned)P.capHub);
                if (ldr->c_push[par] + ldr->c_hub[par] == 0) break;   // nothing pushes: every later residue is zero
                if (rank == 0 && tid == 0) { sm.c_push[par ^ 1] = 0; sm.c_hub[par ^ 1] = 0; }
            } else {
                n_local = min((long long)sm.n_push, P.capP);
                if (G == 1 && n_local == 0) break;
            }
            // ---------------------------------------------------------------- expand (graph.h:94-100)
            // One warp per push-list entry (entries are handed out through a shared counter): the lanes read consecutive
            // CSR entries, so there is no search for the owner of an edge and no barrier inside the level.  Pushing nodes
            // have tens to hundreds of neighbours on every BASELINE shape (a node only pushes while deg <= r / rmax);
            // the rare long entry (a hub source's slice) is expanded by all warps together in a second pass.
            // One edge per lane: the neighbour goes to the table (G == 1) or to its owner's stream (G > 1).
            int *x_id = x_id_base + (long long)par * G * P.capX;
            double *x_val = x_val_base + (long long)par * G * P.capX;
            auto sink = [&](const int vp, const double av, const bool ok) {
                if (G == 1) {
                    if (ok && !*(volatile int *)&sm.full) {
                        const int slot = find_slot(s_keys, (hash_node((unsigned)vp & idmask) >> kBucketShift) & (kBuckets - 1), vp, P.max_probe);
                        if (slot >= 0) atomicAdd(s_vals + slot, av);   // graph.h:98
                        else { ovf = true; sm.full = 1; }
                    }
                } else {
                    const unsigned act = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const unsigned dst = hash_node((unsigned)vp & idmask) >> kOwnerShift;
                        const unsigned peers = __match_any_sync(act, dst);
                        const int leader = __ffs(peers) - 1;
                        unsigned pos = 0;
                        if (lane == leader) pos = atomicAdd(&sm.cnt[dst], (unsigned)__popc(peers));
                        pos = __shfl_sync(peers, pos, leader) + __popc(peers & ((1u << lane) - 1u));
                        if (pos < (unsigned)P.capX) {
                            x_id[dst * P.capX + pos] = vp;
                            x_val[dst * P.capX + pos] = av;
                        } else ovf = true;
                    }
                }
            };
            const int n_items = n_local + n_hub;
            auto load_item = [&](const int i, int &st, unsigned &len, double &add) {
                st = 0; len = 0; add = 0.0;
                if (i < n_local) {
                    if (i < CB) { st = sm.start[i]; len = sm.off[i]; add = sm.add[i]; }   // written by settle
                    else { st = push_start[i]; len = (unsigned)push_len[i]; add = push_add[i]; }
                } else if (i < n_items) {   // this CTA's slice of a hub entry
                    const int h = i - n_local;
                    const unsigned d = (unsigned)__ldcg(hub_deg + h);
                    const unsigned lo = (unsigned)((unsigned long long)d * rank / G), hi = (unsigned)((unsigned long long)d * (rank + 1) / G);
                    st = __ldcg(hub_start + h) + (int)lo; len = hi - lo; add = __ldcg(hub_add + h);
                }
            };
            for (;;) {
                // a warp takes kItemBatch entries at a time and has the first 64 edges of each in flight together: the
                // entries of a level are short, so the level costs about one memory round trip, not one per entry
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
                    if (lane == 0) src_edges += len[k];
                    if (len[k] > (unsigned)kChunk) {   // an uncut entry or a hub slice: all warps expand it together after this pass
                        int bpos = 0;
                        if (lane == 0) bpos = atomicAdd(&sm.n_big, 1);
                        bpos = __shfl_sync(0xffffffffu, bpos, 0);
                        if (bpos < kBigCap) {
                            if (lane == 0) { sm.big_st[bpos] = st[k]; sm.big_len[bpos] = len[k]; sm.big_add[bpos] = add[k]; }
                        } else {   // (more long entries than the list holds: this warp expands it alone)
                            for (unsigned base = 0; base < len[k]; base += 32) {
                                const bool ok = base + lane < len[k];
                                int v = src_packed;
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
                        vp[k][q] = src_packed;
                        if (st[k] >= 0 && (unsigned)(32 * q + lane) < len[k]) vp[k][q] = __ldcs(P.packed + st[k] + 32 * q + lane);   // graph.h:96-97
                    }
                }
#pragma unroll
                for (int k = 0; k < kItemBatch; k++) {
#pragma unroll
                    for (int q = 0; q < kChunk / 32; q++)
                        if ((unsigned)(32 * q) < len[k]) sink(vp[k][q], add[k], (unsigned)(32 * q + lane) < len[k]);   // (warp-uniform condition)
                }
            }
            __syncthreads();
            {
                const int n_big = min(sm.n_big, kBigCap);
                for (int bi = 0; bi < n_big; bi++) {
                    const int st = sm.big_st[bi];
                    const unsigned len = sm.big_len[bi];
                    const double add = sm.big_add[bi];
                    for (unsigned base = (unsigned)(tid & ~31); base < len; base += CB * kEdgeUnroll) {
                        int vp[kEdgeUnroll];
                        bool ok[kEdgeUnroll];
#pragma unroll
                        for (int q = 0; q < kEdgeUnroll; q++) {
                            const unsigned e = base + CB * q + lane;
                            ok[q] = e < len;
                            vp[q] = src_packed;
                            if (ok[q] && st >= 0) vp[q] = __ldcs(P.packed + st + e);
                        }
#pragma unroll
                        for (int q = 0; q < kEdgeUnroll; q++)
                            if (base + CB * q < len) sink(vp[q], add, ok[q]);
                    }
                }
                __syncthreads();   // (also orders every thread's read of n_big before its reset below)
            }
            if (tid == 0) { sm.n_push = 0; sm.n_sel = 0; sm.next_item = 0; sm.n_big = 0; }
            GPC_PHASE(1);
            if (G > 1) {
                // One table update per lane: find-or-claim + fp64 add (graph.h:98).
                auto accumulate = [&](const int vp, const double av, const bool ok) {
                    if (ok && !*(volatile int *)&sm.full) {
                        const int slot = find_slot(s_keys, (hash_node((unsigned)vp & idmask) >> kBucketShift) & (kBuckets - 1), vp, P.max_probe);
                        if (slot >= 0) atomicAdd(s_vals + slot, av);
                        else { ovf = true; sm.full = 1; }
                    }
                };
                // Tell every receiver how much I sent it and ARRIVE at the cluster barrier (every stream of this level is
                // then complete and visible); while the other CTAs get there, accumulate what I sent to myself.
                const unsigned own = min(sm.cnt[rank], (unsigned)P.capX);
                if (tid < G) {
                    CSmem *peer = cg::this_cluster().map_shared_rank(&sm, tid);
                    peer->inbox[par][rank] = (unsigned)tid == rank ? 0u : min(sm.cnt[tid], (unsigned)P.capX);
                    peer->pushed[par][rank] = (unsigned)n_items;   // lets everyone see when no CTA pushed anything
                }
                cg::this_cluster().barrier_arrive();
                {
                    const int *oid = x_id + rank * P.capX;
                    const double *oval = x_val + rank * P.capX;
                    for (unsigned e0 = tid; e0 < own; e0 += CB * kEdgeUnroll) {
                        int vp[kEdgeUnroll];
                        double av[kEdgeUnroll];
#pragma unroll
                        for (int q = 0; q < kEdgeUnroll; q++) {
                            const unsigned e = e0 + q * CB;
                            vp[q] = 0; av[q] = 0.0;
                            if (e < own) { vp[q] = __ldcg(oid + e); av[q] = __ldcg(oval + e); }
                        }
#pragma unroll
                        for (int q = 0; q < kEdgeUnroll; q++) accumulate(vp[q], av[q], e0 + q * CB < own);
                    }
                }
                cg::this_cluster().barrier_wait();
                if (tid < G) sm.cnt[tid] = 0;   // (barriers separate this from the next level's appends)
                if (!kShareHubs) {
                    unsigned any = 0;
#pragma unroll
                    for (int q = 0; q < G; q++) any |= sm.pushed[par][q];
                    if (any == 0) break;   // nothing was pushed anywhere: every later residue is zero (the reserve is complete)
                }
                // ------------------------------------------------------------ exchange: accumulate what the others sent me
                if (tid < 32) {
                    unsigned v = tid < G ? sm.inbox[par][tid] : 0u, incl = v;
#pragma unroll
                    for (int o = 1; o < kClusterMaxG; o <<= 1) {
                        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += y;
                    }
                    if (tid < G) {
                        sm.pre[tid + 1] = incl;
                        sm.seg_off[tid] = (((cta0 + tid) * 2 + par) * G + rank) * P.capX - (long long)(incl - v);   // stream of sender tid, minus its prefix
                    }
                    if (tid == 0) sm.pre[0] = 0;
                }
                __syncthreads();
                const unsigned total = sm.pre[G];
                for (unsigned e0 = tid; e0 < total; e0 += CB * kEdgeUnroll) {
                    int vp[kEdgeUnroll];
                    double av[kEdgeUnroll];
#pragma unroll
                    for (int q = 0; q < kEdgeUnroll; q++) {
                        const unsigned e = e0 + q * CB;
                        vp[q] = 0; av[q] = 0.0;
                        if (e < total) {
                            int sdr = 0;
#pragma unroll
                            for (int k = 1; k < G; k++) sdr += e >= sm.pre[k];
                            const long long a = sm.seg_off[sdr] + e;
                            vp[q] = __ldcg(P.x_id + a);
                            av[q] = __ldcg(P.x_val + a);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < kEdgeUnroll; q++) accumulate(vp[q], av[q], e0 + q * CB < total);
                }
                GPC_PHASE(3);
            }
            __syncthreads();
            // ---------------------------------------------------------------- settle of level + 1 (graph.h:85-93,102; :104-110 for the last)
            {
                const int nl = level + 1;
                const int npar = nl & 1;
                const bool will_push = nl < P.L - 1;
                const double c = P.coef[nl];
                // credit every slot: reserve += coef * r (graph.h:90 / :106) -- an untouched slot adds an exact zero, so the
                // pass has no branch; the bits of `nz` remember which of my slots held a residue
                unsigned nz = 0;
#pragma unroll
                for (int j = 0; j < SPT; j++) {
                    const double r = s_vals[j * CB + tid];
                    rsv[j] = fma(c, r, rsv[j]);
                    if (__double2hiint(r) | __double2loint(r)) nz |= 1u << j;
                    if (!will_push && r != 0.0) s_vals[j * CB + tid] = 0.0;
                }
                src_front += __popc(nz);
                if (will_push) {
                    // take the residues; a node whose degree code allows it becomes a candidate for the next push list
                    while (nz) {
                        const int slot = (__ffs(nz) - 1) * CB + tid;
                        nz &= nz - 1;
                        const double r = s_vals[slot];
                        s_vals[slot] = 0.0;
                        const unsigned key = (unsigned)s_keys[slot];
                        const unsigned code = has_code ? key >> P.idbits : 0u;   // min(deg, cap): a lower bound of deg
                        if (r >= P.rmax * (double)code) {                        // necessary for graph.h:94; the exact test follows
                            const int p = atomicAdd(&sm.n_sel, 1);
                            if (p < kCandCap) { sm.cand.key[p] = key; sm.cand.r[p] = r; }
                            else consider(key, r, npar);
                        }
                    }
                }
                __syncthreads();
                if (will_push) {
                    // the few nodes that may push fetch their {start, degree} record in parallel
                    const int n_sel = min(sm.n_sel, kCandCap);
                    for (int i = tid; i < n_sel; i += CB) consider(sm.cand.key[i], sm.cand.r[i], npar);
                    __syncthreads();
                }
            }
            GPC_PHASE(2);
        }
        if (ovf) atomicOr(&ldr->c_flags, 1u);
        // ------------------------------------------------------------------ top-k, graph.h:111-126
        csync();   // #C: every CTA's overflow flag has landed
        const bool redo = ldr->c_flags != 0;
        if (!redo) { st_frontier += src_front; st_edges += src_edges; }
        // Pre-filter: v_r = the r-th largest (distinct) of a warp's 32 lane maxima, r = ceil(K / warps): every warp holds
        // at least r reserves >= its v_r, so at least K reserves of this CTA are >= tau = min over the warps, and nothing
        // below tau can be among its K largest.  The bulk of a source's support (thousands of reserves of a few 1e-7)
        // never reaches the select; the few dozen survivors are compacted into shared memory and rank-counted.
        double tau = 0.0;
        const int per_warp = (P.K + CB / 32 - 1) / (CB / 32);
        if (!redo && per_warp <= 8) {
            long long m1 = 0;   // positive doubles order like their bit patterns
#pragma unroll
            for (int j = 0; j < SPT; j++) m1 = max(m1, __double_as_longlong(rsv[j]));
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
            for (int i = 1; i < CB / 32; i++) tau = fmin(tau, sm.wtau[i]);
        }
        if (tid == 0) sm.n_list = 0;
        __syncthreads();
        bool listed = !redo && tau > 0.0;
        if (listed) {
#pragma unroll
            for (int j = 0; j < SPT; j++) {
                if (rsv[j] >= tau) {
                    const int pos = atomicAdd(&sm.n_list, 1);
                    if (pos < CB) { sm.add[pos] = rsv[j]; sm.start[pos] = (int)((unsigned)s_keys[j * CB + tid] & idmask); }
                }
            }
            __syncthreads();
            listed = sm.n_list <= CB;   // (uniform) more than the list holds: select from the registers instead
        }
        const int n_list = listed ? sm.n_list : 0;
        auto each_list = [&](auto f) {
            for (int i = tid; i < n_list; i += CB) f(sm.add[i], sm.start[i]);
        };
        auto each_reg = [&](auto f) {
#pragma unroll
            for (int j = 0; j < SPT; j++)
                if (rsv[j] > 0.0) f(rsv[j], (int)((unsigned)s_keys[j * CB + tid] & idmask));
        };
        if (!redo) {
            auto emit_out = [&](int pos, int id, double v) {
                const long long o = it * P.K + pos;
                P.out_row[o] = src; P.out_col[o] = id; P.out_val[o] = v;
                if (P.out_val32) P.out_val32[o] = (float)v;
            };
            int *cand_id = P.cand_id + cta * P.K;
            double *cand_val = P.cand_val + cta * P.K;
            auto emit_cand = [&](int pos, int id, double v) { cand_id[pos] = id; cand_val[pos] = v; };
            int n;
            if (G == 1) n = listed ? block_topk<CB>(sm, P.K, n_list <= kBucketCap, each_list, emit_out) : block_topk<CB>(sm, P.K, false, each_reg, emit_out);
            else n = listed ? block_topk<CB>(sm, P.K, n_list <= kBucketCap, each_list, emit_cand) : block_topk<CB>(sm, P.K, false, each_reg, emit_cand);
            if (G == 1) {
                // unfilled slots read (0, 0, 0.0): what graph.h:117-126 leaves in the caller-zeroed arrays
                for (int i = n + tid; i < P.K; i += CB) {
                    const long long o = it * P.K + i;
                    P.out_row[o] = 0; P.out_col[o] = 0; P.out_val[o] = 0.0;
                    if (P.out_val32) P.out_val32[o] = 0.f;
                }
            } else if (tid == 0) sm.n_cand = n;
        }
        // the table is empty again for the next source
#pragma unroll
        for (int j = 0; j < SPT; j++) {
            const int slot = j * CB + tid;
            if (s_keys[slot] != kEmpty) { st_support += redo ? 0u : 1u; s_keys[slot] = kEmpty; }
            rsv[j] = 0.0;
        }
        if (G > 1) {
            csync();   // #A: candidates of every CTA are visible
            if (rank == 0) {
                if (redo) {
                    if (tid == 0) { P.redo[atomicAdd(P.redo_count, 1ull)] = (int)it; st_redo++; st_sources--; st_cluster--; }
                } else {
                    if (tid < 32) {
                        unsigned v = tid < G ? (unsigned)cg::this_cluster().map_shared_rank(&sm, tid)->n_cand : 0u, incl = v;
#pragma unroll
                        for (int o = 1; o < kClusterMaxG; o <<= 1) {
                            const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += y;
                        }
                        if (tid < G) sm.pre[tid + 1] = incl;
                        if (tid == 0) sm.pre[0] = 0;
                    }
                    __syncthreads();
                    const unsigned total = sm.pre[G];
                    auto each_cand = [&](auto f) {
                        for (unsigned e = tid; e < total; e += CB) {
                            int s = 0;
#pragma unroll
                            for (int k = 1; k < G; k++) s += e >= sm.pre[k];
                            const long long a = (cta0 + s) * P.K + (e - sm.pre[s]);
                            f(__ldcg(P.cand_val + a), __ldcg(P.cand_id + a));
                        }
                    };
                    const int n = block_topk<CB>(sm, P.K, total <= (unsigned)kBucketCap, each_cand, [&](int pos, int id, double v) {
                        const long long o = it * P.K + pos;
                        P.out_row[o] = src; P.out_col[o] = id; P.out_val[o] = v;
                        if (P.out_val32) P.out_val32[o] = (float)v;
                    });
                    for (int i = n + tid; i < P.K; i += CB) {
                        const long long o = it * P.K + i;
                        P.out_row[o] = 0; P.out_col[o] = 0; P.out_val[o] = 0.0;
                        if (P.out_val32) P.out_val32[o] = 0.f;
                    }
                }
                if (tid == 0) { sm.c_push[0] = sm.c_push[1] = 0; sm.c_hub[0] = sm.c_hub[1] = 0; sm.c_flags = 0; }
            }
        } else {
            if (redo && tid == 0) { P.redo[atomicAdd(P.redo_count, 1ull)] = (int)it; st_redo++; st_sources--; st_cluster--; }
            if (tid == 0) sm.c_flags = 0;
        }
        if (tid == 0) { sm.n_push = 0; sm.n_sel = 0; sm.full = 0; }
        __syncthreads();
        GPC_PHASE(4);
    }
    st_frontier = __reduce_add_sync(0xffffffffu, st_frontier);
    st_support = __reduce_add_sync(0xffffffffu, st_support);
    if (lane == 0) {
        if (st_edges) { atomicAdd(P.stats + 0, st_edges); atomicAdd(P.cum + 0, st_edges); }
        if (st_frontier) { atomicAdd(P.stats + 1, (unsigned long long)st_frontier); atomicAdd(P.cum + 1, (unsigned long long)st_frontier); }
        if (st_support) { atomicAdd(P.stats + 2, (unsigned long long)st_support); atomicAdd(P.cum + 2, (unsigned long long)st_support); }
    }
    if (tid == 0) {
        sm.ph[7] = clock64() - t_begin;
        for (int i = 0; i < 8; i++) atomicAdd(P.phase + i, (unsigned long long)sm.ph[i]);
        atomicAdd(P.cum + 3, st_sources);
        atomicAdd(P.cum + 4, st_cluster);
        atomicAdd(P.cum + 5, st_redo);
    }
    if (G > 1) cg::this_cluster().sync();   // nobody leaves while its shared memory may still be read
}
#undef GPC_PHASE

__global__ void pack_indices_kernel(const int2 *node_rec, const int *indices, long long nnz, int idbits, int *packed) {
    const unsigned cap = (1u << (31 - idbits)) - 1u;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const unsigned v = (unsigned)indices[i];
        const unsigned d = (unsigned)__ldg(&node_rec[v].y);
        packed[i] = (int)(v | (min(d, cap) << idbits));
    }
}

template <int G>
int launch_config(int clusters, cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attr, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        GP_CUDA_TRY(cudaFuncSetAttribute(gfpush_cluster_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)gpc_dynamic_smem()));
        if (G > 8) GP_CUDA_TRY(cudaFuncSetAttribute(gfpush_cluster_kernel<G>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured = true;
    }
    *cfg = cudaLaunchConfig_t{};
    cfg->blockDim = dim3(CB);
    cfg->gridDim = dim3((unsigned)(clusters * G));
    cfg->dynamicSmemBytes = gpc_dynamic_smem();
    cfg->stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = G > 1 ? 1 : 0;
    return GP_OK;
}

template <int G>
int max_clusters_t(int num_sms, int *out) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    int rc = launch_config<G>(std::max(num_sms / G, 1), &cfg, attr, nullptr);
    if (rc != GP_OK) return rc;
    if (G == 1) { *out = num_sms; return GP_OK; }
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, gfpush_cluster_kernel<G>, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *out = n;
    return GP_OK;
}

template <int G>
int launch_t(const ClusterPushParams &P, int clusters, cudaStream_t stream) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    int rc = launch_config<G>(clusters, &cfg, attr, stream);
    if (rc != GP_OK) return rc;
    GP_CUDA_TRY(cudaLaunchKernelEx(&cfg, gfpush_cluster_kernel<G>, P));
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

}  // namespace

size_t gpc_dynamic_smem() { return (size_t)kClusterSlots * 12; }

int gpc_pack_indices(const int2 *node_rec, const int *indices, long long nnz, int idbits, int *packed, int num_sms,
                     cudaStream_t stream) {
    if (nnz == 0) return GP_OK;
    pack_indices_kernel<<<num_sms * 8, 256, 0, stream>>>(node_rec, indices, nnz, idbits, packed);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

int gpc_max_clusters(int G, int num_sms, int *out) {
    switch (G) {
        case 1: return max_clusters_t<1>(num_sms, out);
        case 2: return max_clusters_t<2>(num_sms, out);
        case 4: return max_clusters_t<4>(num_sms, out);
        case 8: return max_clusters_t<8>(num_sms, out);
        case 16: return max_clusters_t<16>(num_sms, out);
    }
    gp_set_error("cluster size must be 1, 2, 4, 8 or 16 (got %d)", G);
    return GP_ERR_INVALID;
}

int gpc_launch(const ClusterPushParams &P, int G, int clusters, cudaStream_t stream) {
    switch (G) {
        case 1: return launch_t<1>(P, clusters, stream);
        case 2: return launch_t<2>(P, clusters, stream);
        case 4: return launch_t<4>(P, clusters, stream);
        case 8: return launch_t<8>(P, clusters, stream);
        case 16: return launch_t<16>(P, clusters, stream);
    }
    gp_set_error("cluster size must be 1, 2, 4, 8 or 16 (got %d)", G);
    return GP_ERR_INVALID;
}

}  // namespace gpp
