/*
 * spmv_oracle.c -- CPU restatement of the LessUp/gpu-spmv hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (gpu-spmv_b200/, include/)
 * may include, link, dlopen or call this file.  It is used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker / timed CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks every function
 * here (a) against the reference's own known answers (README 3x3, design.md
 * 3x4 layout, tests/test_spmv.cu 5*2=10 and [3,0,7], Appendix-D fingerprints
 * of SURVEY.md) stored in tests/golden/, and (b) bit-for-bit against the
 * unmodified reference sources compiled into oracle/_ref/libspmv_ref.so when
 * that library is present.
 *
 * All citations are file:line in /root/reference.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off + baseline x86-64 == what the reference gets from
 * `g++ -O2` (no FMA contraction: mul then add, each rounded to fp32).
 */
#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* CSR conversion: src/csr_matrix.cpp:50-95                            */
/* ------------------------------------------------------------------ */

/* nnz count pass, src/csr_matrix.cpp:56-61 (keeps v != 0.0f: -0.0 dropped,
 * NaN kept). */
ORC_API int orc_dense_count_nnz(const float *dense, int rows, int cols)
{
    int nnz = 0;
    for (int i = 0; i < rows * cols; i++)
        if (dense[i] != 0.0f) nnz++;
    return nnz;
}

/* fill pass, src/csr_matrix.cpp:79-92 */
ORC_API void orc_csr_from_dense(const float *dense, int rows, int cols,
                                float *values, int *col_indices, int *row_ptrs)
{
    int idx = 0;
    for (int i = 0; i < rows; i++) {
        row_ptrs[i] = idx;
        for (int j = 0; j < cols; j++) {
            float v = dense[i * cols + j];
            if (v != 0.0f) {
                values[idx] = v;
                col_indices[idx] = j;
                idx++;
            }
        }
    }
    row_ptrs[rows] = idx;
}

/* src/csr_matrix.cpp:97-114 (memset, scatter, last duplicate wins) */
ORC_API void orc_csr_to_dense(int rows, int cols, const float *values,
                              const int *col_indices, const int *row_ptrs,
                              float *dense)
{
    memset(dense, 0, (size_t)rows * (size_t)cols * sizeof(float));
    for (int i = 0; i < rows; i++)
        for (int j = row_ptrs[i]; j < row_ptrs[i + 1]; j++)
            dense[i * cols + col_indices[j]] = values[j];
}

/* src/csr_matrix.cpp:116-136 (linear scan, early break on sorted columns) */
ORC_API float orc_csr_get_element(int rows, int cols, const float *values,
                                  const int *col_indices, const int *row_ptrs,
                                  int row, int col)
{
    if (row < 0 || row >= rows || col < 0 || col >= cols) return 0.0f;
    for (int i = row_ptrs[row]; i < row_ptrs[row + 1]; i++) {
        if (col_indices[i] == col) return values[i];
        if (col_indices[i] > col) break;
    }
    return 0.0f;
}

/* ------------------------------------------------------------------ */
/* Row statistics + selector                                           */
/* ------------------------------------------------------------------ */

typedef struct {
    float avg_nnz_per_row;
    int max_nnz_per_row;
    int min_nnz_per_row;
    float skewness;
} orc_stats;

/* src/csr_matrix.cpp:281-300 */
ORC_API void orc_csr_stats(int rows, int nnz, const int *row_ptrs, orc_stats *out)
{
    out->avg_nnz_per_row = 0.0f;
    out->max_nnz_per_row = 0;
    out->min_nnz_per_row = INT_MAX;
    out->skewness = 0.0f;
    if (rows == 0) {
        out->min_nnz_per_row = 0;
        return;
    }
    out->avg_nnz_per_row = (float)nnz / rows;
    for (int i = 0; i < rows; i++) {
        int n = row_ptrs[i + 1] - row_ptrs[i];
        if (n > out->max_nnz_per_row) out->max_nnz_per_row = n;
        if (n < out->min_nnz_per_row) out->min_nnz_per_row = n;
    }
    out->skewness = (float)out->max_nnz_per_row / (out->min_nnz_per_row + 1);
}

/* src/spmv_cpu.cpp:34-50.  Returns kernel_type (0 scalar, 1 vector,
 * 2 merge-path); block_size and use_texture through out-params. */
ORC_API int orc_auto_config(int rows, int cols, int nnz, const int *row_ptrs,
                            int *block_size, int *use_texture)
{
    orc_stats s;
    *block_size = 256;
    *use_texture = cols > 10000;
    orc_csr_stats(rows, nnz, row_ptrs, &s);
    if (s.avg_nnz_per_row < 4.0f) return 0;
    if (s.skewness < 10.0f) return 1;
    return 2;
}

/* ------------------------------------------------------------------ */
/* ELL conversion: src/ell_matrix.cpp:53-159                           */
/* ------------------------------------------------------------------ */

/* width pass of ell_from_csr, src/ell_matrix.cpp:117-121 */
ORC_API int orc_ell_width_from_csr(int rows, const int *row_ptrs)
{
    int w = 0;
    for (int i = 0; i < rows; i++) {
        int n = row_ptrs[i + 1] - row_ptrs[i];
        if (n > w) w = n;
    }
    return w;
}

/* fill pass of ell_from_csr, src/ell_matrix.cpp:139-156: padding col=-1,
 * val=0.0f; entry k of row i at k*rows+i in CSR order. */
ORC_API void orc_ell_from_csr(int rows, int width, const float *values,
                              const int *col_indices, const int *row_ptrs,
                              float *ell_values, int *ell_cols)
{
    size_t size = (size_t)rows * (size_t)width;
    for (size_t i = 0; i < size; i++) {
        ell_cols[i] = -1;
        ell_values[i] = 0.0f;
    }
    for (int i = 0; i < rows; i++) {
        int k = 0;
        for (int j = row_ptrs[i]; j < row_ptrs[i + 1]; j++) {
            size_t idx = (size_t)k * rows + i;
            ell_values[idx] = values[j];
            ell_cols[idx] = col_indices[j];
            k++;
        }
    }
}

/* width pass of ell_from_dense, src/ell_matrix.cpp:59-68 */
ORC_API int orc_ell_width_from_dense(const float *dense, int rows, int cols)
{
    int w = 0;
    for (int i = 0; i < rows; i++) {
        int n = 0;
        for (int j = 0; j < cols; j++)
            if (dense[i * cols + j] != 0.0f) n++;
        if (n > w) w = n;
    }
    return w;
}

/* fill pass of ell_from_dense, src/ell_matrix.cpp:86-106 */
ORC_API void orc_ell_from_dense(const float *dense, int rows, int cols, int width,
                                float *ell_values, int *ell_cols)
{
    size_t size = (size_t)rows * (size_t)width;
    for (size_t i = 0; i < size; i++) {
        ell_cols[i] = -1;
        ell_values[i] = 0.0f;
    }
    for (int i = 0; i < rows; i++) {
        int k = 0;
        for (int j = 0; j < cols; j++) {
            float v = dense[i * cols + j];
            if (v != 0.0f) {
                size_t idx = (size_t)k * rows + i;
                ell_values[idx] = v;
                ell_cols[idx] = j;
                k++;
            }
        }
    }
}

/* src/ell_matrix.cpp:161-182 */
ORC_API void orc_ell_to_dense(int rows, int cols, int width, const float *ell_values,
                              const int *ell_cols, float *dense)
{
    memset(dense, 0, (size_t)rows * (size_t)cols * sizeof(float));
    for (int i = 0; i < rows; i++)
        for (int k = 0; k < width; k++) {
            size_t idx = (size_t)k * rows + i;
            int c = ell_cols[idx];
            if (c >= 0) dense[i * cols + c] = ell_values[idx];
        }
}

/* src/ell_matrix.cpp:184-200 (stops at first padding slot) */
ORC_API float orc_ell_get_element(int rows, int cols, int width, const float *ell_values,
                                  const int *ell_cols, int row, int col)
{
    if (row < 0 || row >= rows || col < 0 || col >= cols) return 0.0f;
    for (int k = 0; k < width; k++) {
        size_t idx = (size_t)k * rows + row;
        if (ell_cols[idx] == col) return ell_values[idx];
        if (ell_cols[idx] < 0) break;
    }
    return 0.0f;
}

/* ------------------------------------------------------------------ */
/* CPU SpMV: src/spmv_cpu.cpp:6-32                                      */
/* ------------------------------------------------------------------ */

/* spmv_cpu_csr, src/spmv_cpu.cpp:6-16: sequential fp32, mul then add. */
ORC_API void orc_spmv_csr(int rows, const float *values, const int *col_indices,
                          const int *row_ptrs, const float *x, float *y)
{
    for (int i = 0; i < rows; i++) {
        float sum = 0.0f;
        for (int j = row_ptrs[i]; j < row_ptrs[i + 1]; j++)
            sum += values[j] * x[col_indices[j]];
        y[i] = sum;
    }
}

/* spmv_cpu_ell, src/spmv_cpu.cpp:18-32 */
ORC_API void orc_spmv_ell(int rows, int width, const float *ell_values,
                          const int *ell_cols, const float *x, float *y)
{
    for (int i = 0; i < rows; i++) {
        float sum = 0.0f;
        for (int k = 0; k < width; k++) {
            size_t idx = (size_t)k * rows + i;
            int c = ell_cols[idx];
            if (c >= 0) sum += ell_values[idx] * x[c];
        }
        y[i] = sum;
    }
}

/* Supplementary restatement: same row dot as
 * src/spmv_cpu.cpp:6-16 with an f64 accumulator, plus the per-row scale
 * sum_j |a_ij x_j| that north_star's tolerance is relative to. */
ORC_API void orc_spmv_csr_f64(int rows, const float *values, const int *col_indices,
                              const int *row_ptrs, const float *x, double *y,
                              double *abs_scale)
{
    for (int i = 0; i < rows; i++) {
        double sum = 0.0, asum = 0.0;
        for (int j = row_ptrs[i]; j < row_ptrs[i + 1]; j++) {
            double p = (double)values[j] * (double)x[col_indices[j]];
            sum += p;
            asum += fabs(p);
        }
        y[i] = sum;
        if (abs_scale) abs_scale[i] = asum;
    }
}

ORC_API void orc_spmv_ell_f64(int rows, int width, const float *ell_values,
                              const int *ell_cols, const float *x, double *y,
                              double *abs_scale)
{
    for (int i = 0; i < rows; i++) {
        double sum = 0.0, asum = 0.0;
        for (int k = 0; k < width; k++) {
            size_t idx = (size_t)k * rows + i;
            int c = ell_cols[idx];
            if (c >= 0) {
                double p = (double)ell_values[idx] * (double)x[c];
                sum += p;
                asum += fabs(p);
            }
        }
        y[i] = sum;
        if (abs_scale) abs_scale[i] = asum;
    }
}

/* ------------------------------------------------------------------ */
/* Bandwidth model: src/bandwidth.cpp:22-88                             */
/* ------------------------------------------------------------------ */

/* byte count of compute_bandwidth_csr, src/bandwidth.cpp:34-42 */
ORC_API uint64_t orc_bytes_csr(int rows, int cols, int nnz)
{
    size_t rd = 0;
    rd += nnz * sizeof(float);
    rd += nnz * sizeof(int);
    rd += (rows + 1) * sizeof(int);
    rd += cols * sizeof(float);
    size_t wr = rows * sizeof(float);
    return (uint64_t)(rd + wr);
}

/* byte count of compute_bandwidth_ell, src/bandwidth.cpp:66-75 */
ORC_API uint64_t orc_bytes_ell(int rows, int cols, int width)
{
    size_t ell = (size_t)rows * width;
    size_t rd = ell * sizeof(float) + ell * sizeof(int) + cols * sizeof(float);
    size_t wr = rows * sizeof(float);
    return (uint64_t)(rd + wr);
}

/* achieved GB/s exactly as src/bandwidth.cpp:45-46 (fp32 arithmetic) */
ORC_API float orc_achieved_gbs(uint64_t total_bytes, float elapsed_ms)
{
    if (elapsed_ms <= 0.0f) return 0.0f;
    float elapsed_s = elapsed_ms / 1000.0f;
    return ((size_t)total_bytes / 1e9f) / elapsed_s;
}

/* ------------------------------------------------------------------ */
/* PageRank: src/pagerank.cu:11-153                                     */
/* ------------------------------------------------------------------ */

/* find_dangling_nodes, src/pagerank.cu:20-48: fp32 column sums in row-major
 * order; dangling iff the sum == 0.0f.  flags[col] = 1 for dangling. */
ORC_API int orc_find_dangling(int rows, int cols, const float *values,
                              const int *col_indices, const int *row_ptrs,
                              unsigned char *flags)
{
    float *sums = (float *)calloc((size_t)(cols > 0 ? cols : 1), sizeof(float));
    int count = 0;
    for (int r = 0; r < rows; r++)
        for (int j = row_ptrs[r]; j < row_ptrs[r + 1]; j++) {
            int c = col_indices[j];
            if (c >= 0 && c < cols) sums[c] += values[j];
        }
    for (int c = 0; c < cols; c++) {
        flags[c] = (sums[c] == 0.0f);
        count += flags[c];
    }
    free(sums);
    return count;
}

/*
 * Literal fp32 recurrence of pagerank(), src/pagerank.cu:50-153, with
 * spmv_cpu_csr standing in for the GPU VECTOR_CSR call at :102 (the
 * reference has no CPU PageRank; every other line is the host code it runs).
 * Valid as a literal oracle for n <~ 2^24.
 *   ranks      out [n]
 *   returns    iterations; *residual = final L2; *converged = flag
 */
ORC_API int orc_pagerank_f32(int n, int cols, const float *values, const int *col_indices,
                             const int *row_ptrs, float damping, float tolerance,
                             int max_iterations, float *ranks, float *residual,
                             int *converged)
{
    float *r_old = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    float *r_new = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    unsigned char *dang = (unsigned char *)calloc((size_t)(cols > 0 ? cols : 1), 1);
    float init = 1.0f / n;                            /* :69 */
    for (int i = 0; i < n; i++) r_old[i] = init;
    float teleport = (1.0f - damping) / n;            /* :86 */
    orc_find_dangling(n, cols, values, col_indices, row_ptrs, dang);
    int iters = 0, conv = 0, from_new = 0;
    float res = 0.0f;
    for (int it = 0; it < max_iterations; it++) {
        float dsum = 0.0f;                            /* :94-99 */
        for (int c = 0; c < cols; c++)
            if (dang[c] && c < n) dsum += r_old[c];
        orc_spmv_csr(n, values, col_indices, row_ptrs, r_old, r_new); /* :102 */
        float dcontrib = (n > 0) ? (damping * dsum / n) : 0.0f;       /* :111 */
        for (int i = 0; i < n; i++)
            r_new[i] = damping * r_new[i] + dcontrib + teleport;     /* :113 */
        float s = 0.0f;                               /* :11-18 */
        for (int i = 0; i < n; i++) {
            float d = r_new[i] - r_old[i];
            s += d * d;
        }
        res = sqrtf(s);
        iters = it + 1;
        if (res < tolerance) { conv = 1; from_new = 1; break; }      /* :123 */
        float *t = r_old; r_old = r_new; r_new = t;   /* :130-131 */
    }
    const float *fin = from_new ? r_new : r_old;      /* :135-139 */
    for (int i = 0; i < n; i++) ranks[i] = fin[i];
    float sum = 0.0f;                                 /* :142-150 */
    for (int i = 0; i < n; i++) sum += ranks[i];
    if (sum > 0.0f)
        for (int i = 0; i < n; i++) ranks[i] /= sum;
    *residual = res;
    *converged = conv;
    free(r_old); free(r_new); free(dang);
    return iters;
}

/*
 * Supplementary restatement: the same recurrence with f64
 * accumulators for the three host sums that break in fp32 at n = 2^26
 * (dangling mass, sum d^2, normalisation) and an f64 row dot.  Storage stays
 * fp32, the update expression stays ((d*y) + dc) + tp in fp32, the stop rule
 * stays L2 < tolerance.  Also reports the L1 residual north_star asks for.
 * fixed_iterations > 0 runs exactly that many iterations (ignores tolerance)
 * so two implementations can be compared at equal iteration counts.
 */
ORC_API int orc_pagerank_f64(int n, int cols, const float *values, const int *col_indices,
                             const int *row_ptrs, float damping, float tolerance,
                             int max_iterations, int fixed_iterations, float *ranks,
                             double *residual_l2, double *residual_l1, int *converged)
{
    float *r_old = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    float *r_new = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    unsigned char *dang = (unsigned char *)calloc((size_t)(cols > 0 ? cols : 1), 1);
    float init = 1.0f / n;
    for (int i = 0; i < n; i++) r_old[i] = init;
    float teleport = (1.0f - damping) / n;
    orc_find_dangling(n, cols, values, col_indices, row_ptrs, dang);
    int iters = 0, conv = 0, from_new = 0;
    double l2 = 0.0, l1 = 0.0;
    int limit = fixed_iterations > 0 ? fixed_iterations : max_iterations;
    for (int it = 0; it < limit; it++) {
        double dsum = 0.0;
        for (int c = 0; c < cols; c++)
            if (dang[c] && c < n) dsum += (double)r_old[c];
        float dsum_f = (float)dsum;
        float dcontrib = (n > 0) ? (damping * dsum_f / n) : 0.0f;
        double s2 = 0.0, s1 = 0.0;
        for (int i = 0; i < n; i++) {
            double acc = 0.0;
            for (int j = row_ptrs[i]; j < row_ptrs[i + 1]; j++)
                acc += (double)values[j] * (double)r_old[col_indices[j]];
            float y = (float)acc;
            float v = damping * y + dcontrib + teleport;
            r_new[i] = v;
            double d = (double)v - (double)r_old[i];
            s2 += d * d;
            s1 += fabs(d);
        }
        l2 = sqrt(s2);
        l1 = s1;
        iters = it + 1;
        if (fixed_iterations <= 0 && (float)l2 < tolerance) { conv = 1; from_new = 1; break; }
        if (fixed_iterations > 0 && it + 1 == limit) { from_new = 1; conv = ((float)l2 < tolerance); break; }
        float *t = r_old; r_old = r_new; r_new = t;
    }
    const float *fin = from_new ? r_new : r_old;
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += (double)fin[i];
    float sum_f = (float)sum;
    for (int i = 0; i < n; i++) ranks[i] = (sum_f > 0.0f) ? fin[i] / sum_f : fin[i];
    *residual_l2 = l2;
    *residual_l1 = l1;
    *converged = conv;
    free(r_old); free(r_new); free(dang);
    return iters;
}

/* ------------------------------------------------------------------ */
/* Merge-path coordinates and row partitioning                          */
/* ------------------------------------------------------------------ */

/*
 * Canonical Merrill-Garland diagonal search (the corrected form of
 * merge_path_search, src/spmv_kernels.cu:48-72): list A is
 * the row-END offsets row_ptrs[1..rows], list B the nz indices 0..nnz-1.
 */
ORC_API void orc_merge_path_search(int diagonal, const int *row_ptrs, int rows, int nnz,
                                   int *out_row, int *out_nz)
{
    int lo = diagonal - nnz > 0 ? diagonal - nnz : 0;
    int hi = diagonal < rows ? diagonal : rows;
    while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (row_ptrs[mid + 1] <= diagonal - mid - 1) lo = mid + 1;
        else hi = mid;
    }
    *out_row = lo;
    *out_nz = diagonal - lo;
}

/*
 * Contiguous row split: boundary p is the first row whose prefix
 * work row_ptrs[i] + i*row_weight is >= p*total/parts (lower_bound), with
 * bounds[0]=0 and bounds[parts]=rows.  row_weight 0 = nnz-balanced, 1 =
 * (rows + nnz)-balanced, the coarse level of the merge-path search.
 */
ORC_API void orc_partition_rows_weighted(int rows, const int *row_ptrs, int parts, int row_weight, int *bounds)
{
    int64_t total = (int64_t)row_ptrs[rows] + (int64_t)rows * row_weight;
    bounds[0] = 0;
    for (int p = 1; p < parts; p++) {
        int64_t target = (total * p) / parts;
        int lo = 0, hi = rows;
        while (lo < hi) {
            int mid = lo + (hi - lo) / 2;
            if ((int64_t)row_ptrs[mid] + (int64_t)mid * row_weight < target) lo = mid + 1;
            else hi = mid;
        }
        bounds[p] = lo;
    }
    bounds[parts] = rows;
    for (int p = 1; p <= parts; p++)
        if (bounds[p] < bounds[p - 1]) bounds[p] = bounds[p - 1];
}

ORC_API void orc_partition_rows(int rows, const int *row_ptrs, int parts, int *bounds)
{
    orc_partition_rows_weighted(rows, row_ptrs, parts, 0, bounds);
}

/* ------------------------------------------------------------------ */
/* top-k: src/pagerank.cu:162-185 ("identical up to ties")              */
/* ------------------------------------------------------------------ */

typedef struct { int node_id; float rank; } orc_topk_node;

static int orc_topk_cmp(const void *a, const void *b)
{
    const orc_topk_node *x = (const orc_topk_node *)a, *y = (const orc_topk_node *)b;
    if (x->rank > y->rank) return -1;
    if (x->rank < y->rank) return 1;
    return (x->node_id > y->node_id) - (x->node_id < y->node_id);
}

/* full sort (small n only); ties broken by node id so the checker can
 * compare rank values position by position. */
ORC_API void orc_top_k(const float *ranks, int n, int k, orc_topk_node *out)
{
    orc_topk_node *all = (orc_topk_node *)malloc((size_t)(n > 0 ? n : 1) * sizeof(*all));
    for (int i = 0; i < n; i++) { all[i].node_id = i; all[i].rank = ranks[i]; }
    qsort(all, (size_t)n, sizeof(*all), orc_topk_cmp);
    int kk = k < n ? k : n;
    for (int i = 0; i < kk; i++) out[i] = all[i];
    free(all);
}

/* FNV-1a-64 over raw bytes: the fingerprint used in SURVEY Appendix D. */
ORC_API uint64_t orc_fnv1a64(const void *data, size_t nbytes)
{
    const unsigned char *p = (const unsigned char *)data;
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < nbytes; i++) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return h;
}
