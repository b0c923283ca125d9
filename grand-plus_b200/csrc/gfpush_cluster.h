// gfpush_cluster.h -- host interface of gfpush_cluster.cu (one GFPush source per thread-block cluster).
#pragma once
#include <cuda_runtime.h>

namespace gpp {

constexpr int kClusterBlock = 512;          // threads per CTA
constexpr int kClusterSlots = 16384;        // residue-table slots per CTA
constexpr int kClusterMaxG = 16;            // CTAs per source

struct ClusterPushParams {
    const int2 *node_rec;      // [n] {indptr[v], degree}
    const int *packed;         // [nnz] neighbour id | min(degree, cap) << idbits  (gpc_pack_indices)
    int n;
    int idbits;                // ids occupy the low idbits bits of a packed entry, the code the bits up to 30; 32 = no degree code
    const int *node_idx;
    long long S;
    const double *coef;        // device [L]
    int L;
    double rmax;
    int K;
    int *out_row;
    int *out_col;
    double *out_val;
    float *out_val32;          // nullable
    // per-CTA scratch
    int *push_start;           // [ctas][capP]  push list of the level: CSR offset (-1 = dangling -> source)
    int *push_len;             // [ctas][capP]
    double *push_add;          // [ctas][capP]  r / deg
    long long capP;
    int *x_id;                 // [ctas][2][G][capX]  exchange streams sender -> owner (two generations): packed node
    double *x_val;             // [ctas][2][G][capX]  ... and the pushed amount
    long long capX;
    int *cand_id;              // [ctas][K]  local top-k candidates
    double *cand_val;          // [ctas][K]
    // per-cluster scratch
    int *hub_start;            // [clusters][capHub]  entries expanded by the whole cluster
    int *hub_deg;
    double *hub_add;
    int capHub;
    int hub_min_deg;
    int max_probe;             // 4-key buckets tried before a source is handed to the slab kernel
    unsigned long long *queue; // [1] next source
    unsigned long long *stats; // [0] edges [1] frontier [2] support [3] error flags
    unsigned long long *cum;   // [0] edges [1] frontier [2] support [3] sources [4] cluster sources [5] redone sources
    unsigned long long *phase; // [8] SM cycles per phase, summed over CTAs
    int *redo;                 // sources handed over (table / stream overflow)
    unsigned long long *redo_count;
};

// packed[e] = indices[e] | min(deg(indices[e]), cap) << idbits, cap = 2^(31-idbits) - 1 (packed entries are non-negative)
int gpc_pack_indices(const int2 *node_rec, const int *indices, long long nnz, int idbits, int *packed, int num_sms,
                     cudaStream_t stream);
// dynamic shared memory one CTA needs
size_t gpc_dynamic_smem();
// clusters of G CTAs the device keeps resident (0 = this cluster size cannot be scheduled)
int gpc_max_clusters(int G, int num_sms, int *out);
int gpc_launch(const ClusterPushParams &P, int G, int clusters, cudaStream_t stream);

}  // namespace gpp
