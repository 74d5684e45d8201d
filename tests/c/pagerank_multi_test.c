/* pagerank_multi_test.c -- a plain C program (no Python, no C++) that links against
 * libspmv_b200.so through include/spmv_b200.h and runs PageRank on n_gpus devices from one process
 * (spmv_b200_pagerank_multi), then checks the result against the single-GPU drop-in pagerank()
 * entry point (spmv_b200_pagerank, reference include/spmv/pagerank.h:29-32) on the same graph:
 * ranks within an L1 distance of 1e-6, same iteration count (+-1), sum of ranks 1.  The drop-in
 * pagerank() divides by the reference's sequential fp32 sum (src/pagerank.cu:142-150; off by ~1e-4
 * at n = 2^16) while the multi-GPU form divides by an f64 sum, so both vectors
 * are renormalised in f64 here before they are compared.
 *
 *   cc -std=c11 -Iinclude tests/c/pagerank_multi_test.c -Lgpu-spmv_b200/lib -lspmv_b200 -lm
 *   ./a.out <n_gpus> [scale] [exchange: -1 auto, 0 nccl, 1 p2p, 2 multicast]
 *
 * The graph: a deterministic scale-free-ish digraph built here (each node links to 8 hashed targets
 * biased towards low ids), column-normalised as pagerank() expects (rows = destinations). */
#include "spmv_b200.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint32_t mix(uint32_t x) {
    x = (x ^ (x >> 16)) * 0x7FEB352Du;
    x = (x ^ (x >> 15)) * 0x846CA68Bu;
    return x ^ (x >> 16);
}

int main(int argc, char** argv) {
    const int n_gpus = argc > 1 ? atoi(argv[1]) : 2;
    const int scale = argc > 2 ? atoi(argv[2]) : 16;
    const int exchange = argc > 3 ? atoi(argv[3]) : SPMV_B200_EXCHANGE_AUTO;
    const int n = 1 << scale, deg = 8;
    const long long m = (long long)n * deg;

    /* edges src -> dst; a quarter of the nodes have no out-links (dangling) */
    int* src = (int*)malloc(sizeof(int) * m);
    int* dst = (int*)malloc(sizeof(int) * m);
    int* outdeg = (int*)calloc(n, sizeof(int));
    int* indeg = (int*)calloc(n + 1, sizeof(int));
    long long e = 0;
    for (int s = 0; s < n; ++s) {
        if ((mix((uint32_t)s * 7u + 1u) & 3u) == 0u) continue;
        for (int k = 0; k < deg; ++k) {
            uint32_t h = mix((uint32_t)s * 31u + (uint32_t)k + 99u);
            uint32_t span = 1u << (1 + (mix(h) % (uint32_t)scale)); /* biased towards low ids */
            int d = (int)(h % span) % n;
            src[e] = s; dst[e] = d; ++e;
            ++outdeg[s]; ++indeg[d + 1];
        }
    }
    for (int i = 0; i < n; ++i) indeg[i + 1] += indeg[i];
    spmv_b200_csr* A = spmv_b200_csr_create(n, n, (int)e);
    if (!A) { fprintf(stderr, "csr_create failed\n"); return 2; }
    memcpy(A->row_ptrs, indeg, sizeof(int) * (n + 1));
    int* cursor = (int*)malloc(sizeof(int) * n);
    memcpy(cursor, indeg, sizeof(int) * n);
    for (long long j = 0; j < e; ++j) { /* sources are generated in ascending order: columns stay sorted per row */
        const int p = cursor[dst[j]]++;
        A->col_indices[p] = src[j];
        A->values[p] = 1.0f / (float)outdeg[src[j]];
    }
    free(cursor); free(src); free(dst); free(indeg); free(outdeg);

    spmv_b200_pagerank_config cfg;
    cfg.damping_factor = 0.85f; cfg.tolerance = 1e-6f; cfg.max_iterations = 100;

    /* single GPU, drop-in entry point */
    int rc = spmv_b200_csr_to_gpu(A);
    if (rc != 0) { fprintf(stderr, "csr_to_gpu: %s\n", spmv_b200_error_string(rc)); return 2; }
    spmv_b200_pagerank_result one;
    rc = spmv_b200_pagerank(A, &cfg, &one);
    if (rc != 0 || !one.ranks) { fprintf(stderr, "pagerank: %s\n", spmv_b200_error_string(rc)); return 2; }

    /* n_gpus devices, one process */
    float* ranks = (float*)malloc(sizeof(float) * n);
    spmv_b200_pr_dist_result res;
    rc = spmv_b200_pagerank_multi(A, &cfg, n_gpus, NULL, exchange, 4, 0, ranks, &res);
    if (rc != 0) { fprintf(stderr, "pagerank_multi: %s\n", spmv_b200_error_string(rc)); return 3; }

    double l1 = 0.0, sum = 0.0, sum_one = 0.0;
    for (int i = 0; i < n; ++i) { sum += ranks[i]; sum_one += one.ranks[i]; }
    for (int i = 0; i < n; ++i) l1 += fabs((double)ranks[i] / sum - (double)one.ranks[i] / sum_one);
    printf("n = %d, nnz = %lld, %d GPUs, exchange %d (used %d), graph replay %d, %d kernels/iteration\n", n, e, n_gpus,
           exchange, res.exchange, res.graph_replay, res.kernels_per_iteration);
    printf("single GPU: %d iterations, residual %.3e, converged %d\n", one.iterations, one.final_residual, (int)one.converged);
    printf("%d GPUs   : %d iterations, residual %.3e, converged %d, %.3f ms/iteration (device)\n", n_gpus, res.iterations,
           res.final_residual, res.converged, res.device_seconds / (res.iterations_launched > 0 ? res.iterations_launched : 1) * 1e3);
    printf("L1 distance %.3e, sum of ranks %.9f\n", l1, sum);
    const int ok = l1 <= 1e-6 && fabs(sum - 1.0) <= 1e-5 && abs(res.iterations - one.iterations) <= 1 &&
                   res.converged == (int)one.converged;
    spmv_b200_pagerank_free(&one);
    spmv_b200_csr_destroy(A);
    free(ranks);
    printf("%s\n", ok ? "PASS" : "FAIL");
    return ok ? 0 : 1;
}
