/*
 * grandplus_b200.h -- C ABI of libgrandplus_b200.so: the B200 (sm_100a) implementation of
 * GRAND+'s propagation hot path.  Plain pointers and sizes only; no torch / pybind types.
 *
 * Part 1 replaces the reference's pybind11 module `precompute.propagation`
 *   (/root/reference/precompute/propagation.cpp:8-12, class Graph in
 *    /root/reference/precompute/graph.h:17-133; callers /root/reference/model.py:251,268 and
 *    /root/reference/model_mag.py:271,289).
 * Part 2 replaces the torch_scatter calls inside Grand_Plus.random_prop
 *   (/root/reference/model.py:80-87, /root/reference/model_mag.py:80-86), the host-side
 *   gather around it (/root/reference/model.py:310-316) and MLP.emb
 *   (/root/reference/model_mag.py:48-55).
 *
 * Conventions
 *   - every function returns 0 on success or a negative gp_status; gp_last_error() gives the
 *     message of the calling thread's last failure (the reference checks nothing and overruns
 *     silently, graph.h:59-71 -- this ABI validates and refuses instead);
 *   - "host" pointers are ordinary process memory, "device" pointers are CUDA device memory on
 *     the handle's / current device; `stream` is a cudaStream_t passed as void* (NULL = legacy
 *     default stream);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     GP_ERR_CUDA.
 */
#ifndef GRANDPLUS_B200_H
#define GRANDPLUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_ABI_VERSION 1

typedef enum {
    GP_OK = 0,
    GP_ERR_INVALID = -1,  /* bad argument (null pointer, negative size, unsorted CSR, id out of range) */
    GP_ERR_CUDA = -2,     /* CUDA runtime error or no device                                          */
    GP_ERR_NOMEM = -3,    /* device or host allocation failed                                         */
    GP_ERR_OVERFLOW = -4  /* a frontier/support list outgrew its analytic bound (should not happen)   */
} gp_status;

const char *gp_last_error(void);
int gp_abi_version(void);
/* Number of visible CUDA devices (0 when there is none); never fails. */
int gp_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Part 1: GFPush + top-k
 * ------------------------------------------------------------------------------------------ */

/* Replaces Graph::Graph(indptr, indices, seed) (graph.h:32-47).  The CSR is COPIED to `device`
 * (the reference borrows the caller's NumPy buffers, graph.h:34-36); out-degrees are
 * indptr[i+1]-indptr[i] as in graph.h:42-45.  indptr is int32[n_nodes+1], indices int32[nnz],
 * both host.  `seed` is accepted and ignored exactly like the reference's (graph.h:30,40). */
typedef struct gp_graph gp_graph;
int gp_graph_create(const int32_t *indptr, int64_t n_nodes, const int32_t *indices, int64_t nnz,
                    int32_t seed, int device, gp_graph **out);
/* Same, from CSR arrays already resident on `device` (used by the multi-GPU driver and bench). */
int gp_graph_create_device(const int32_t *d_indptr, int64_t n_nodes, const int32_t *d_indices, int64_t nnz,
                           int device, gp_graph **out);
void gp_graph_destroy(gp_graph *g);
int64_t gp_graph_num_nodes(const gp_graph *g);
int64_t gp_graph_num_edges(const gp_graph *g);

/* Scratch policy for the per-source residue/reserve tables. */
typedef enum {
    GP_SCRATCH_AUTO = 0,   /* shared memory when the graph is small enough, else HBM slabs */
    GP_SCRATCH_SMEM = 1,   /* dense next-residue table in shared memory (n_nodes*8 B must fit)  */
    GP_SCRATCH_HBM = 2     /* direct-addressed per-CTA slabs in HBM / L2                        */
} gp_scratch_mode;

typedef struct {
    int32_t scratch_mode;      /* gp_scratch_mode                                      */
    int32_t block_threads;     /* 0 = default                                           */
    int32_t ctas_per_sm;       /* 0 = default                                           */
    int64_t max_scratch_bytes; /* 0 = default (half of free device memory)              */
} gp_push_config;
int gp_graph_configure(gp_graph *g, const gp_push_config *cfg);

/* Replaces Graph::gfpush_omp(node_idx, row_idx, col_idx, value, coef, rmax, K) (graph.h:53-131).
 * All pointers HOST.  node_idx int32[S]; coef float64[L] (L = order+1 levels); outputs int32 /
 * int32 / float64 [S*K].  For source `it` the selected entries land in slots it*K .. it*K+cnt-1
 * (order within a row unspecified, as with nth_element, graph.h:115); every other slot is written
 * as (0, 0, 0.0), which is what the reference leaves in the caller-zeroed arrays
 * (model.py:252-254, graph.h:121).  Synchronous. */
int gp_gfpush(gp_graph *g, const int32_t *node_idx, int64_t S, const double *coef, int32_t L,
              double rmax, int32_t K, int32_t *row_idx, int32_t *col_idx, double *value);

/* Same computation with node_idx and the outputs resident on the graph's device; coef stays a
 * host array.  d_value32 (nullable) additionally receives the values cast to fp32 -- the dtype
 * the aggregation consumes (model.py:315).  Asynchronous on `stream`. */
int gp_gfpush_device(gp_graph *g, const int32_t *d_node_idx, int64_t S, const double *coef, int32_t L,
                     double rmax, int32_t K, int32_t *d_row_idx, int32_t *d_col_idx, double *d_value,
                     float *d_value32, void *stream);

/* Work counters of the most recent gfpush on this handle (synchronises the handle's stream).
 * edges_pushed / frontier_total are SURVEY 8(d)'s E_push / F_tot. */
typedef struct {
    int64_t edges_pushed;
    int64_t frontier_total;
    int64_t support_total;
    int64_t sources;
    int64_t ctas;            /* persistent CTAs launched                              */
    int64_t scratch_bytes;   /* device scratch held by the handle                     */
    int32_t scratch_mode;    /* mode actually used (gp_scratch_mode)                  */
    int32_t kernel_launches; /* kernels launched by the call                          */
    int64_t cluster_sources; /* cumulative stats only: sources finished by the cluster kernel          */
    int64_t redo_sources;    /* cumulative stats only: sources it handed over to the slab kernel       */
    int32_t cluster_size;    /* CTAs per source of the cluster kernel (0 = per-CTA kernels only)       */
    int32_t table_slots;     /* shared-memory residue-table slots per CTA (0 = residues on the HBM slabs)  */
    int32_t bucket_count;    /* node-range buckets of the bucket kernel (0 = not used)                     */
    int32_t reserved;
} gp_push_stats;
int gp_gfpush_last_stats(gp_graph *g, gp_push_stats *out);
/* Counters summed over every gfpush since creation / the last reset (device-wide synchronise);
 * lets a caller time a burst of asynchronous calls and read the work afterwards. */
int gp_gfpush_cumulative_stats(gp_graph *g, gp_push_stats *out, int reset);

/* Profiling hook (the reference only takes an unused gettimeofday pair, graph.h:54-56,129-130): SM cycles the
 * persistent CTAs spent per phase, summed over CTAs since the last reset:
 * [0] fetch + level 0, [1] expand, [2] settle, [3] reserve merge (per-CTA kernels) / exchange (cluster kernel), [4] top-k,
 * [5] expand of each source's widest level, [6] its settle (per-CTA kernels), [7] kernel residency.
 * Device-wide synchronise. */
int gp_gfpush_phase_cycles(gp_graph *g, uint64_t out[8], int reset);

/* ------------------------------------------------------------------------------------------
 * Part 2: fused gather - mask - scale - reduce aggregation (all pointers DEVICE)
 * ------------------------------------------------------------------------------------------
 * out[a, b, :] = sum_j m_{a,j} * table[nbr_j, :] / (sum_j m_{a,j} + eps)     j over row b's entries
 *   m_{a,j} = score_j                        in eval mode (training == 0) or p == 0
 *           = keep_{a,j} * score_j / (1-p)   in training mode  (F.dropout, model.py:82)
 * keep_{a,j} is drawn from Philox4x32-10 keyed by (seed, offset) at counter (entry index, a), or
 * read from mask_in.  One launch produces n_aug augmentations (model.py:321 loops `sample`
 * times over the same batch) and reads each table row at most once.
 */
typedef struct {
    /* table: [n_table_rows, F] fp32, row stride ld_table elements (>= F) */
    const float *table;
    int64_t n_table_rows;
    int32_t F;
    int64_t ld_table;
    /* entries.  Layout A (CSR): row_ptr int32[B+1]; entry j of row b is j in [row_ptr[b], row_ptr[b+1]).
     * Layout B (slots, row_ptr == NULL): entry i of row b is slot_rows[b]*slot_K + i, i < slot_K, and
     * is skipped when score <= 0 (the zero pads of graph.h:117-126).  slot_rows == NULL -> b itself. */
    const int32_t *row_ptr;
    const int32_t *slot_rows;
    int32_t slot_K;
    const int32_t *nbr;      /* int32 per entry: table row; NULL = entry index itself (pre-gathered feats) */
    const float *score;      /* fp32 per entry                                                          */
    int64_t B;               /* output rows                                                             */
    int64_t n_entries;       /* length of nbr/score/mask arrays (CSR: row_ptr[B]; slots: n_slot_rows*K)  */
    /* DropNode */
    double p;                /* drop probability as the caller's Python float (scale = 1/(1-p) in fp32) */
    int32_t training;
    int32_t n_aug;           /* 1..4 */
    uint64_t seed, offset;
    const uint8_t *mask_in;  /* nullable [n_aug, n_entries], 1 = keep */
    uint8_t *mask_out;       /* nullable [n_aug, n_entries]           */
    float eps;               /* 1e-12 for random_prop (model.py:87), 1e-10 for emb (model_mag.py:54) */
    /* outputs */
    float *out;              /* [n_aug, B, F], row stride ld_out */
    int64_t ld_out;
    float *denom_out;        /* nullable [n_aug, B]: sum_j m + eps, saved for the backward pass */
} gp_aggregate_args;
int gp_aggregate_fwd(const gp_aggregate_args *args, void *stream);

/* Backward of the above w.r.t. the table rows (model_mag.py:356 keeps autograd through
 * random_prop and emb):  grad_table[nbr_j, :] += m_{a,j} / denom[a,b] * grad_out[a,b,:].
 * With nbr == NULL every entry owns its row and the result is written, not accumulated.
 * The mask must be supplied (mask_in) when training != 0 && p > 0. */
typedef struct {
    const float *grad_out;   /* [n_aug, B, F], row stride ld_grad_out */
    int64_t ld_grad_out;
    const float *denom;      /* [n_aug, B] from the forward pass      */
    const int32_t *row_ptr;  /* CSR layout only                       */
    const int32_t *nbr;
    const float *score;
    int64_t B, n_entries;
    int32_t F;
    double p;
    int32_t training;
    int32_t n_aug;
    const uint8_t *mask_in;
    float *grad_table;       /* [n_table_rows, F] (accumulated into) or [n_entries, F] when nbr == NULL */
    int64_t ld_grad_table;
    int64_t n_table_rows;
} gp_aggregate_bwd_args;
int gp_aggregate_bwd(const gp_aggregate_bwd_args *args, void *stream);

/* MLP.emb with a non-zero input dropout (model_mag.py:48-55; scripts/run_mag.sh uses 0.0, which is gp_aggregate_fwd with
 * eps 1e-10): out[b,:] = sum_j w_j * drop(table[idx_j,:]) / (sum_j w_j + eps), drop(x) = keep ? x/(1-p) : 0 per ELEMENT,
 * keep drawn from Philox4x32-10 at counter (entry j, column block) keyed by (seed, offset).  row_ptr int32[B+1] segments
 * the entries by output row (node_idx ascending, dim_size = node_idx[-1]+1); denom_out (nullable) receives sum_j w_j + eps.
 * The backward regenerates the same mask and accumulates into COMPACT gradient rows: slot[j] is the position of entry j's
 * table row among the batch's distinct rows, grad_rows is [n_distinct, H] (zeroed by the caller). */
int gp_emb_dropout_fwd(const float *table, int64_t n_table_rows, int64_t ld_table, int32_t H, const int32_t *row_ptr,
                       const int32_t *idx, const float *weight, int64_t B, double p, uint64_t seed, uint64_t offset,
                       float eps, float *out, int64_t ld_out, float *denom_out, void *stream);
int gp_emb_dropout_bwd(const float *grad_out, int64_t ld_grad_out, int32_t H, const int32_t *row_ptr, const int32_t *slot,
                       const float *weight, const float *denom, int64_t B, double p, uint64_t seed, uint64_t offset,
                       float *grad_rows, int64_t ld_grad_rows, void *stream);
/* The element mask the two calls above draw, uint8 [nza, H] (for export / tests). */
int gp_emb_dropout_mask(int64_t nza, int32_t H, double p, uint64_t seed, uint64_t offset, uint8_t *d_mask, void *stream);

/* Optimizer step of the embedding table on the touched rows only, EXACTLY equal to the reference's dense
 * torch.optim.Adam(lr, betas, eps, weight_decay=0) over the whole table (model_mag.py:312-313,369): a row that is not touched
 * still moves under dense Adam while its moments decay, so every row remembers the step it was last brought up to date
 * (last_step int32[n_rows]) and the skipped zero-gradient steps last_step+1 .. upto are replayed when it is next read or
 * updated.  rows int64[R] (distinct; NULL = all rows 0..R-1).  grad_rows == NULL: catch up to step `upto` only (call before
 * a forward pass that reads the rows, and with rows == NULL before inference / saving).  grad_rows != NULL ([R, H]):
 * catch up to `upto`, then apply step upto+1 with the gradient. */
int gp_lazy_adam_rows(float *param, float *exp_avg, float *exp_avg_sq, int32_t *last_step, int64_t ld, int32_t H,
                      const int64_t *rows, int64_t R, const float *grad_rows, int64_t ld_grad, int32_t upto, float lr,
                      float beta1, float beta2, float eps, void *stream);

/* mat_idx (int64, ascending; model.py:84 takes dim_size = mat_idx[-1]+1) -> CSR row_ptr int32[B+1].
 * *d_flags (int32[2], device) receives {1 if unsorted or out of range, 0 otherwise; unused}. */
int gp_segments_from_sorted_index(const int64_t *d_idx, int64_t n, int64_t B, int32_t *d_row_ptr,
                                  int32_t *d_flags, void *stream);
/* int64 -> int32 index narrowing with range check against n_rows (flag set when out of range). */
int gp_narrow_index(const int64_t *d_idx, int64_t n, int64_t n_rows, int32_t *d_out, int32_t *d_flags, void *stream);

/* Philox4x32-10 keep-mask exactly as gp_aggregate_fwd draws it (for export / tests). */
int gp_dropnode_mask(int64_t n_entries, int32_t n_aug, double p, uint64_t seed, uint64_t offset,
                     uint8_t *d_mask, void *stream);

/* Performance knobs for sweeps (profiles/); defaults are the measured best and results never depend on them.
 * Aggregation: "agg_kernel" (0 auto, 1 register-staged LDG kernel, 2 TMA-staged cp.async.bulk kernel), "agg_nbuf",
 * "agg_max_vec", "agg_max_chunk", "agg_smem_kb".
 * GFPush (graphs beyond the dense shared-memory mode): "push_smem_hash" (shared-memory residue table in front of the slabs:
 * 0 off, 1 auto from rmax, 2 always), "push_smem_probe" (4-key buckets tried before a node goes to the slab), "push_max_ctas"
 * (cap on persistent CTAs, for scaling experiments); the opt-in cluster kernel (one source per thread-block cluster):
 * "push_cluster" (0 off, 1 auto, -1 = one CTA, 2, 4, 8, 16 CTAs per source), "push_cluster_probe", "push_hub_deg",
 * "push_max_clusters"; "push_bucket" (the hash-bucket kernel, the default beyond the dense shared-memory mode: 0 off, 1 auto, 2 always),
 * "push_bucket_block" (threads per CTA: 1024 = one CTA per SM, 512 = two, 256 = three; 0 = from the expected support), "push_bucket_fill"
 * (eighths of the table one visit may fill, 3..7), "push_bucket_nb" (buckets per source, 0 = automatic), "push_bucket_merge" (0 = the bucket kernel merges only the top-k candidates of a source and does not count its support, 1 = it merges the whole reserve).  The same keys are read from the GP_TUNING environment variable ("key=value,key=value") by the
 * Python loader. */
int gp_set_tuning(const char *key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* GRANDPLUS_B200_H */
