// gfpush_bucket.h -- host interface of gfpush_bucket.cu (hash buckets, one at a time in a shared-memory table).
#pragma once
#include <cuda_runtime.h>

namespace gpp {

// threads per CTA: 1024 (one CTA per SM), 512 (two per SM) or 256 (three per SM); the shared-memory {key, residue} table has
// 16 slots per thread
constexpr int gpb_slots(int block) { return block * 16; }
constexpr int kBucketMaxBuckets = 256;   // bucket counters live in shared memory

struct BucketPushParams {
    const int2 *node_rec;      // [n] {indptr[v], degree}
    const int *packed;         // [nnz] neighbour id | degree code << idbits (gpc_pack_indices)
    int n;
    int idbits;
    int block;                 // threads per CTA: 1024, 512 or 256 (gpb_slots(block) table slots)
    int nb;                    // buckets = 2^log_nb (>= 2): bucket of a node = hash(node) >> (32 - log_nb)
    int log_nb;
    int max_probe;             // 4-key buckets of the table tried before the source is handed to the slab kernel
    int group_pairs;           // consecutive buckets are one table fill while their pairs stay below this (default 5/8 of the slots)
    int full_merge;            // 1: merge the whole reserve (counts the support); 0: only the top-k candidates (default)
    const int *node_idx;
    long long S;               // sources [it_base, S) of node_idx are processed by this launch
    long long it_base;
    const double *coef;        // device [L]
    int L;
    double rmax;
    int K;
    int *out_row;
    int *out_col;
    double *out_val;
    float *out_val32;          // nullable
    // per-CTA scratch
    int *pair_id;              // [ctas][nb][capPair]  pushed edges of the level, by bucket: packed node
    double *pair_val;          // [ctas][nb][capPair]  ... and the pushed amount r / deg
    long long capPair;
    long long pair_stride;     // nb * capPair: one CTA's streams
    int *log_id;               // [ctas][nb][capLog]   reserve log of the source, by bucket: packed node
    double *log_val;           // [ctas][nb][capLog]   ... and coef * r
    long long capLog;
    long long log_stride;      // nb * capLog
    int *push_start;           // [ctas][capP]  push list of the level beyond the entries kept in shared memory
    int *push_len;
    double *push_add;
    long long capP;
    int *sup_id;               // [ctas][capS]  merged reserve: node ...
    double *sup_val;           // [ctas][capS]  ... and value
    long long capS;
    unsigned long long *queue;
    unsigned long long *max_support;   // largest support of any source (the planner sizes the buckets from it)
    unsigned long long *stats; // [0] edges [1] frontier [2] support [3] error flags
    unsigned long long *cum;
    unsigned long long *phase;
    int *redo;                 // sources handed to the slab kernel (a bucket stream outgrew its capacity)
    unsigned long long *redo_count;
};

size_t gpb_dynamic_smem(int nb, int block);
int gpb_ctas_per_sm(int nb, int block, int *per_sm);   // resident CTAs per SM of that geometry (occupancy query)
int gpb_launch(const BucketPushParams &P, int ctas, cudaStream_t stream);

}  // namespace gpp
