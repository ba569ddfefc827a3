/*
 * oracle/cpu_flat.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the flat-index search that TextReact's retrieval step
 * delegates to the third-party `faiss` wheel:
 *     reference call sites  retrieve/retrieve_faiss.py:65  faiss.IndexFlatL2(d)
 *                           retrieve/retrieve_faiss.py:66  index.add(train_fps)
 *                           retrieve/retrieve_faiss.py:71  index.search(query_fps, k)
 * and, for the 768-d neural retriever whose output enters the repo through
 * retrieve/convert_format.py:7-16, faiss.IndexFlatIP with depth 100.
 *
 * PARITY UNPINNED: `faiss` is neither vendored nor pinned by the reference
 * (absent from environment.yml) and is not importable in the build image, and the
 * reference ships no tests / golden vectors for this path.  What is restated
 * here is FAISS's *published* flat-search algorithm:
 *   - scalar path (FAISS uses it for nq < 20): one dot product / squared
 *     distance per (query, row), a k-element binary heap per query that is
 *     replaced only by strictly better entries, then a final reorder;
 *   - BLAS path (nq >= 20): sgemm over query-block x database-block tiles
 *     (done by the numpy caller in oracle/cpu_flat.py), and the same heap fed
 *     one score block at a time -> trx_oracle_heap_block below.
 * Ties are resolved on (score, id): better score first, then lower id, which is
 * the (val, id) comparison recent FAISS heaps use.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file.  The product (textreact_b200) never does.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TRX_METRIC_IP 0
#define TRX_METRIC_L2 1

/* "a is worse than b" under the total order (score, id).
 * For IP larger score is better, for L2 smaller distance is better; on equal
 * score the larger id is the worse one. */
static inline int worse_ip(float va, int64_t ia, float vb, int64_t ib) {
    return (va < vb) || (va == vb && ia > ib);
}
static inline int worse_l2(float va, int64_t ia, float vb, int64_t ib) {
    return (va > vb) || (va == vb && ia > ib);
}

/* Binary heap whose root is the WORST kept entry (FAISS: CMin heap for IP, CMax
 * heap for L2).  1-based sift like faiss/utils/Heap.h. */
static void heap_replace_top(int metric, int k, float* v, int64_t* id, float nv, int64_t nid) {
    int i = 1;
    v--; id--;                               /* 1-based */
    for (;;) {
        int l = 2 * i, r = l + 1, c;
        if (l > k) break;
        if (r > k) c = l;
        else {
            int l_worse = metric == TRX_METRIC_IP ? worse_ip(v[l], id[l], v[r], id[r])
                                                  : worse_l2(v[l], id[l], v[r], id[r]);
            c = l_worse ? l : r;             /* child that is worse (closer to root) */
        }
        int new_worse = metric == TRX_METRIC_IP ? worse_ip(nv, nid, v[c], id[c])
                                                : worse_l2(nv, nid, v[c], id[c]);
        if (new_worse) break;
        v[i] = v[c]; id[i] = id[c];
        i = c;
    }
    v[i] = nv; id[i] = nid;
}

static inline void heap_offer(int metric, int k, float* v, int64_t* id, float nv, int64_t nid) {
    int top_worse = metric == TRX_METRIC_IP ? worse_ip(v[0], id[0], nv, nid)
                                            : worse_l2(v[0], id[0], nv, nid);
    if (top_worse) heap_replace_top(metric, k, v, id, nv, nid);
}

/* Fill value FAISS leaves in unfilled slots: -FLT_MAX for IP, +FLT_MAX for L2, id -1.
 * The (score,id) order treats id -1 with the neutral score as worst because a
 * real score can only tie it at +-FLT_MAX, which finite inputs never produce. */
void trx_oracle_heap_init(int metric, int64_t nq, int k, float* D, int64_t* I) {
    float fill = metric == TRX_METRIC_IP ? -FLT_MAX : FLT_MAX;
    for (int64_t i = 0; i < nq * (int64_t)k; i++) { D[i] = fill; I[i] = -1; }
}

/* Feed one block of scores[nq][nb] (row-major, ld = nb) for database rows
 * base_id .. base_id+nb-1 into the per-query heaps (D, I hold heaps in place).
 * groups/excl implement the gold-removed mode restated from
 * textreact/dataset.py:74-76 as an index-side mask: row j is ineligible for
 * query i iff excl[i] >= 0 && groups[j] == excl[i]. */
void trx_oracle_heap_block(int metric, int64_t nq, int64_t nb, int64_t base_id, int k,
                           const float* scores, const int32_t* groups /*global, nullable*/,
                           const int32_t* excl /*nullable*/, float* D, int64_t* I) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; i++) {
        float* v = D + i * k;
        int64_t* id = I + i * k;
        const float* s = scores + i * nb;
        int32_t ex = excl ? excl[i] : -1;
        for (int64_t j = 0; j < nb; j++) {
            if (ex >= 0 && groups && groups[base_id + j] == ex) continue;
            /* -1 slots: replace unconditionally while unfilled (root is -1 until k seen) */
            if (id[0] < 0) heap_replace_top(metric, k, v, id, s[j], base_id + j);
            else heap_offer(metric, k, v, id, s[j], base_id + j);
        }
    }
}

static int cmp_metric;
typedef struct { float v; int64_t id; } pair_t;
static int pair_cmp(const void* a, const void* b) {
    const pair_t* x = (const pair_t*)a; const pair_t* y = (const pair_t*)b;
    if (x->id < 0 && y->id < 0) return 0;
    if (x->id < 0) return 1;                 /* unfilled slots last */
    if (y->id < 0) return -1;
    if (x->v != y->v) {
        if (cmp_metric == TRX_METRIC_IP) return x->v > y->v ? -1 : 1;
        return x->v < y->v ? -1 : 1;
    }
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}

/* Final reorder: best first, ties by ascending id, -1 padding last. */
void trx_oracle_heap_reorder(int metric, int64_t nq, int k, float* D, int64_t* I) {
    cmp_metric = metric;
    pair_t* tmp = (pair_t*)malloc(sizeof(pair_t) * (size_t)k);
    for (int64_t i = 0; i < nq; i++) {
        for (int j = 0; j < k; j++) { tmp[j].v = D[i * k + j]; tmp[j].id = I[i * k + j]; }
        qsort(tmp, (size_t)k, sizeof(pair_t), pair_cmp);
        for (int j = 0; j < k; j++) { D[i * k + j] = tmp[j].v; I[i * k + j] = tmp[j].id; }
    }
    free(tmp);
}

/* Scalar path (FAISS nq < 20: exhaustive_inner_product_seq / exhaustive_L2sqr_seq):
 * direct fp32 dot product or sum of squared differences, 8 partial sums like the
 * 8-wide SIMD accumulators of fvec_inner_product / fvec_L2sqr. */
static float dot8(const float* a, const float* b, int64_t d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t i = 0;
    for (; i + 8 <= d; i += 8)
        for (int l = 0; l < 8; l++) acc[l] += a[i + l] * b[i + l];
    float s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
    for (; i < d; i++) s += a[i] * b[i];
    return s;
}
static float l2sqr8(const float* a, const float* b, int64_t d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t i = 0;
    for (; i + 8 <= d; i += 8)
        for (int l = 0; l < 8; l++) { float t = a[i + l] - b[i + l]; acc[l] += t * t; }
    float s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
    for (; i < d; i++) { float t = a[i] - b[i]; s += t * t; }
    return s;
}

void trx_oracle_search_seq(int metric, const float* xb, int64_t nb, const float* xq, int64_t nq,
                           int64_t d, int k, const int32_t* groups, const int32_t* excl,
                           float* D, int64_t* I) {
    trx_oracle_heap_init(metric, nq, k, D, I);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < nq; i++) {
        float* v = D + i * k;
        int64_t* id = I + i * k;
        const float* q = xq + i * d;
        int32_t ex = excl ? excl[i] : -1;
        for (int64_t j = 0; j < nb; j++) {
            if (ex >= 0 && groups && groups[j] == ex) continue;
            float s = metric == TRX_METRIC_IP ? dot8(q, xb + j * d, d) : l2sqr8(q, xb + j * d, d);
            if (id[0] < 0) heap_replace_top(metric, k, v, id, s, j);
            else heap_offer(metric, k, v, id, s, j);
        }
    }
    trx_oracle_heap_reorder(metric, nq, k, D, I);
}

/* fp64 arbiter: scores in double, used only to compute the rank-k gap that decides
 * where ids MUST agree (north_star: relative gap > 1e-5) and to settle disputes. */
void trx_oracle_scores_f64(int metric, const float* xb, int64_t nb, const float* xq, int64_t nq,
                           int64_t d, double* out /*[nq][nb]*/) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; i++) {
        const float* q = xq + i * d;
        for (int64_t j = 0; j < nb; j++) {
            const float* x = xb + j * d;
            double s = 0.0;
            if (metric == TRX_METRIC_IP) for (int64_t t = 0; t < d; t++) s += (double)q[t] * (double)x[t];
            else for (int64_t t = 0; t < d; t++) { double u = (double)q[t] - (double)x[t]; s += u * u; }
            out[i * nb + j] = s;
        }
    }
}
