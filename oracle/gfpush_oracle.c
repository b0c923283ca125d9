/*
 * oracle/gfpush_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement, in plain C, of the reference's GFPush + top-k
 * (/root/reference/precompute/graph.h:53-131).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product path
 * (grand-plus_b200/) never does and fails loudly without its CUDA library.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against golden vectors
 * produced by the reference's own propagation.cpp compiled unmodified (oracle/Makefile,
 * target `ref`) on the real Cora / Citeseer / Pubmed graphs at the scripts/run_*.sh
 * parameters (tests/golden/make_golden.py), and -- when oracle/_ref is present --
 * against the reference module live.
 *
 * What is restated, line by line:
 *   graph.h:76-82   residue = {src: 1}, reserve = {src: 0}
 *   graph.h:83      for i in 0 .. L-2                      (level-synchronous)
 *   graph.h:85-90     pop (u, r); reserve[u] += coef[i]*r  (credited BEFORE the test)
 *   graph.h:91-93     deg(u)==0 -> next[src] += r          (dangling mass to the source)
 *   graph.h:94-100    r >= rmax*deg(u) -> next[v] += r/deg(u) for v in N(u); else dropped
 *   graph.h:102       residue = next
 *   graph.h:104-110 last level: reserve[u] += coef[L-1]*r
 *   graph.h:111-115 k = min(K, |reserve|); nth_element descending by value
 *   graph.h:117-126 write the first k with v > 0 to slots it*K+i; other slots untouched
 *
 * Deliberate differences (none observable beyond fp64 summation order, <= 1e-15 rel):
 *   - std::unordered_map is replaced by dense arrays + touched lists, so per-node sums
 *     are taken in frontier order instead of hash-bucket order;
 *   - nth_element (ties arbitrary, slot order arbitrary) is replaced by a full sort,
 *     descending by value, ties broken by ascending node id, so the oracle's own output
 *     is deterministic; parity tests compare sets modulo the tie band;
 *   - the reference's read-after-erase of iter->second (graph.h:86-89) is not copied.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t edges_pushed;   /* sum over levels, over pushed u, of deg(u)          */
    int64_t frontier_total; /* sum over levels of |frontier| (incl. the last one) */
    int64_t support_total;  /* sum over sources of |reserve|                      */
    int64_t max_frontier;   /* largest single-level frontier seen                 */
    int64_t max_support;    /* largest |reserve| seen                             */
} gp_oracle_stats;

typedef struct {
    int n;
    double *cur, *nxt, *rsv; /* dense [n]                       */
    uint8_t *in_nxt, *seen;  /* membership flags [n]            */
    int *cur_l, *nxt_l, *sup_l;
    int n_cur, n_nxt, n_sup;
} scratch_t;

static int scratch_init(scratch_t *s, int n) {
    s->n = n;
    s->cur = (double *)calloc((size_t)n, sizeof(double));
    s->nxt = (double *)calloc((size_t)n, sizeof(double));
    s->rsv = (double *)calloc((size_t)n, sizeof(double));
    s->in_nxt = (uint8_t *)calloc((size_t)n, 1);
    s->seen = (uint8_t *)calloc((size_t)n, 1);
    s->cur_l = (int *)malloc((size_t)n * sizeof(int));
    s->nxt_l = (int *)malloc((size_t)n * sizeof(int));
    s->sup_l = (int *)malloc((size_t)n * sizeof(int));
    s->n_cur = s->n_nxt = s->n_sup = 0;
    return s->cur && s->nxt && s->rsv && s->in_nxt && s->seen && s->cur_l && s->nxt_l && s->sup_l;
}

static void scratch_free(scratch_t *s) {
    free(s->cur); free(s->nxt); free(s->rsv); free(s->in_nxt); free(s->seen);
    free(s->cur_l); free(s->nxt_l); free(s->sup_l);
}

static inline void touch_reserve(scratch_t *s, int u) {
    if (!s->seen[u]) { s->seen[u] = 1; s->sup_l[s->n_sup++] = u; }
}

static inline void add_next(scratch_t *s, int v, double x) {
    if (!s->in_nxt[v]) { s->in_nxt[v] = 1; s->nxt_l[s->n_nxt++] = v; s->nxt[v] = 0.0; }
    s->nxt[v] += x;
}

/* graph.h:76-110 for one source.  Leaves the reserve in s->rsv over s->sup_l[0..n_sup). */
static void push_one(const int32_t *indptr, const int32_t *indices, int src,
                     const double *coef, int L, double rmax, scratch_t *s, gp_oracle_stats *st) {
    s->n_cur = s->n_nxt = s->n_sup = 0;
    s->cur_l[s->n_cur++] = src;            /* graph.h:80 */
    s->cur[src] = 1.0;
    touch_reserve(s, src);                 /* graph.h:81 */
    s->rsv[src] = 0.0;
    for (int i = 0; i < L - 1; i++) {      /* graph.h:83 */
        if (s->n_cur > st->max_frontier) st->max_frontier = s->n_cur;
        st->frontier_total += s->n_cur;
        for (int j = 0; j < s->n_cur; j++) {
            int u = s->cur_l[j];
            double r = s->cur[u];
            touch_reserve(s, u);
            s->rsv[u] += coef[i] * r;      /* graph.h:90 */
            uint32_t deg = (uint32_t)(indptr[u + 1] - indptr[u]);
            if (deg == 0) {
                add_next(s, src, r);       /* graph.h:91-93 */
            } else if (r >= rmax * deg) {  /* graph.h:94: double * unsigned -> double */
                double val = r / deg;      /* graph.h:95 */
                for (int e = indptr[u]; e < indptr[u + 1]; e++) add_next(s, indices[e], val);
                st->edges_pushed += deg;
            }
        }
        /* graph.h:102  residue = next */
        for (int j = 0; j < s->n_cur; j++) s->cur[s->cur_l[j]] = 0.0;
        for (int j = 0; j < s->n_nxt; j++) {
            int v = s->nxt_l[j];
            s->cur[v] = s->nxt[v]; s->nxt[v] = 0.0; s->in_nxt[v] = 0;
            s->cur_l[j] = v;
        }
        s->n_cur = s->n_nxt; s->n_nxt = 0;
    }
    if (s->n_cur > st->max_frontier) st->max_frontier = s->n_cur;
    st->frontier_total += s->n_cur;
    for (int j = 0; j < s->n_cur; j++) {   /* graph.h:104-110 */
        int u = s->cur_l[j];
        touch_reserve(s, u);
        s->rsv[u] += coef[L - 1] * s->cur[u];
        s->cur[u] = 0.0;
    }
    s->n_cur = 0;
    st->support_total += s->n_sup;
    if (s->n_sup > st->max_support) st->max_support = s->n_sup;
}

static void clear_reserve(scratch_t *s) {
    for (int j = 0; j < s->n_sup; j++) { int u = s->sup_l[j]; s->rsv[u] = 0.0; s->seen[u] = 0; }
    s->n_sup = 0;
}

typedef struct { double v; int c; } pair_t;
static int cmp_desc(const void *a, const void *b) {
    const pair_t *x = (const pair_t *)a, *y = (const pair_t *)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return (x->c > y->c) - (x->c < y->c);
}

static void stats_merge(gp_oracle_stats *dst, const gp_oracle_stats *src) {
    dst->edges_pushed += src->edges_pushed;
    dst->frontier_total += src->frontier_total;
    dst->support_total += src->support_total;
    if (src->max_frontier > dst->max_frontier) dst->max_frontier = src->max_frontier;
    if (src->max_support > dst->max_support) dst->max_support = src->max_support;
}

/* Same argument meaning as Graph::gfpush_omp (graph.h:53): outputs are caller-allocated,
 * caller-zeroed [S*K]; slot it*K+i is written only when the i-th selected value is > 0.
 * nthreads <= 0 -> all cores (the reference hard-codes 40, graph.h:41).
 * Returns 0, or -1 on allocation failure. */
int gp_oracle_gfpush(const int32_t *indptr, const int32_t *indices, int32_t n,
                     const int32_t *node_idx, int64_t S, const double *coef, int32_t L,
                     double rmax, int32_t K, int32_t *row_idx, int32_t *col_idx, double *value,
                     gp_oracle_stats *stats_out, int32_t nthreads) {
    gp_oracle_stats total; memset(&total, 0, sizeof total);
    int failed = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_num_procs();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        scratch_t s; gp_oracle_stats st; memset(&st, 0, sizeof st);
        pair_t *res = (pair_t *)malloc((size_t)n * sizeof(pair_t));
        int ok = scratch_init(&s, n) && res;
        if (!ok) {
#pragma omp atomic write
            failed = 1;
        }
#pragma omp for schedule(dynamic)          /* graph.h:73 */
        for (int64_t it = 0; it < S; it++) {
            if (!ok) continue;
            int src = node_idx[it];
            push_one(indptr, indices, src, coef, L, rmax, &s, &st);
            int m = s.n_sup;                /* graph.h:111 */
            for (int j = 0; j < m; j++) { res[j].c = s.sup_l[j]; res[j].v = s.rsv[s.sup_l[j]]; }
            int k = m > K ? K : m;          /* graph.h:113 */
            qsort(res, (size_t)m, sizeof(pair_t), cmp_desc);
            for (int i = 0; i < k; i++) {   /* graph.h:117-126 */
                if (res[i].v > 0.0) {
                    int64_t idx = it * K + i;
                    row_idx[idx] = src; col_idx[idx] = res[i].c; value[idx] = res[i].v;
                }
            }
            clear_reserve(&s);
        }
#pragma omp critical
        stats_merge(&total, &st);
        if (ok) scratch_free(&s);
        free(res);
    }
    if (stats_out) *stats_out = total;
    return failed ? -1 : 0;
}

/* The full (un-truncated) reserve vector of one source, dense [n], for the tie-band
 * checks in tests/: out[v] = reserve value, seen[v] = 1 where v is a key of the
 * reference's reserve map (may be NULL). */
int gp_oracle_reserve_row(const int32_t *indptr, const int32_t *indices, int32_t n, int32_t src,
                          const double *coef, int32_t L, double rmax, double *out, uint8_t *seen,
                          gp_oracle_stats *stats_out) {
    scratch_t s; gp_oracle_stats st; memset(&st, 0, sizeof st);
    if (!scratch_init(&s, n)) return -1;
    push_one(indptr, indices, src, coef, L, rmax, &s, &st);
    memset(out, 0, (size_t)n * sizeof(double));
    if (seen) memset(seen, 0, (size_t)n);
    for (int j = 0; j < s.n_sup; j++) {
        int u = s.sup_l[j];
        out[u] = s.rsv[u];
        if (seen) seen[u] = 1;
    }
    if (stats_out) *stats_out = st;
    scratch_free(&s);
    return 0;
}
