/* examples/c_abi_example.c -- the drop-in boundary from plain C (no Python, no torch, no C++).
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -Ltextreact_b200 -ltrx -Wl,-rpath,$PWD/textreact_b200 -o /tmp/trx_example
 *
 * What retrieve/retrieve_faiss.py:62-74 does through the faiss object, through libtrx.so:
 * IndexFlatL2(d) -> add(train_fps) -> search(query_fps, k) -> (distance, rank).  Needs a B200 to RUN; compiling and
 * linking it is part of the CPU test suite (tests/test_abi.py) as the check that include/trx.h is valid C. */
#include <stdio.h>
#include <stdlib.h>

#include "trx.h"

int main(void) {
    const int d = 64, k = 5;
    const int64_t n = 20000, nq = 3;
    float* xb = (float*)malloc((size_t)n * d * sizeof(float));
    float* D = (float*)malloc((size_t)nq * k * sizeof(float));
    int64_t* I = (int64_t*)malloc((size_t)nq * k * sizeof(int64_t));
    unsigned s = 1u;
    for (int64_t i = 0; i < n * d; i++) { s = s * 1664525u + 1013904223u; xb[i] = (float)(s >> 8) / 16777216.0f - 0.5f; }

    trx_index* idx = NULL;
    if (trx_create(d, TRX_METRIC_L2, 0, &idx) != TRX_OK) { fprintf(stderr, "create: %s\n", trx_last_error()); return 2; }
    if (trx_add(idx, xb, n) != TRX_OK) { fprintf(stderr, "add: %s\n", trx_last_error()); return 3; }
    /* the first three stored rows as queries: each must find itself at distance 0 */
    if (trx_search(idx, xb, nq, k, NULL, D, I, NULL) != TRX_OK) { fprintf(stderr, "search: %s\n", trx_last_error()); return 4; }
    for (int64_t q = 0; q < nq; q++) {
        printf("query %lld:", (long long)q);
        for (int j = 0; j < k; j++) printf(" (%lld, %.4f)", (long long)I[q * k + j], D[q * k + j]);
        printf("\n");
        if (I[q * k] != q || D[q * k] != 0.0f) { fprintf(stderr, "self match missing\n"); return 5; }
    }
    trx_search_params_t p;
    p.exclude = NULL; p.attr_below = 2147483647; p.dedup_groups = 0; p.self_row0 = 100;   /* rows 100.. as queries */
    if (trx_search_ex(idx, NULL, nq, k, &p, D, I, NULL) != TRX_OK || I[0] != 100) { fprintf(stderr, "search_ex: %s\n", trx_last_error()); return 6; }
    trx_stats_t st;
    trx_stats(idx, &st);
    printf("%s: ntotal=%lld path=%d launches=%lld\n", trx_version(), (long long)trx_ntotal(idx), st.last_path, (long long)st.launches);
    trx_destroy(idx);
    free(xb); free(D); free(I);
    return 0;
}
