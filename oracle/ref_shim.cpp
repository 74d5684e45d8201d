// ref_shim.cpp -- extern "C" handles over the UNMODIFIED reference library.
//
// TEST INFRASTRUCTURE ONLY (see oracle/spmv_oracle.c header).  This file is
// compiled together with the reference's own sources, read in place from
// /root/reference/src (never copied into this repo), into
// oracle/_ref/libspmv_ref.so by oracle/Makefile.  It adds no arithmetic: every
// function forwards to the reference's public API (include/spmv/*.h) so that
// Python tests can drive it through ctypes, and so that bench.py's
// `--impl reference` / cpu_baseline legs can time spmv_cpu_csr / spmv_cpu_ell.
//
// The input generators replay the reference's own test fixtures
// (include/spmv/test_utils.h:12-58, tests/test_spmv.cu:40-48,
// tests/test_pagerank.cu:18-40) so golden vectors are the inputs the
// reference's tests actually see.
#include "spmv/spmv.h"
#include "spmv/pagerank.h"
#include "spmv/bandwidth.h"
#include "spmv/benchmark.h"
#include "spmv/test_utils.h"

#include <chrono>
#include <cstring>
#include <vector>

using namespace spmv;

#define REF_API extern "C" __attribute__((visibility("default")))

// ---- CSR handles -----------------------------------------------------------

// non-owning view over caller arrays (host side only)
REF_API void* ref_csr_wrap(int rows, int cols, int nnz, float* values, int* col_indices,
                           int* row_ptrs) {
    CSRMatrix* m = new CSRMatrix();
    std::memset(m, 0, sizeof(*m));
    m->num_rows = rows; m->num_cols = cols; m->nnz = nnz;
    m->values = values; m->col_indices = col_indices; m->row_ptrs = row_ptrs;
    m->owns_host_memory = false; m->owns_device_memory = false;
    return m;
}

REF_API void* ref_csr_from_dense(const float* dense, int rows, int cols, int* status) {
    CSRMatrix* m = csr_create(0, 0, 0);
    int rc = csr_from_dense(m, dense, rows, cols);
    if (status) *status = rc;
    return m;
}

REF_API void ref_csr_fields(void* h, int* rows, int* cols, int* nnz, float** values,
                            int** col_indices, int** row_ptrs) {
    CSRMatrix* m = static_cast<CSRMatrix*>(h);
    *rows = m->num_rows; *cols = m->num_cols; *nnz = m->nnz;
    *values = m->values; *col_indices = m->col_indices; *row_ptrs = m->row_ptrs;
}

REF_API void ref_csr_destroy(void* h) { csr_destroy(static_cast<CSRMatrix*>(h)); }
REF_API int ref_csr_to_dense(void* h, float* dense) { return csr_to_dense(static_cast<CSRMatrix*>(h), dense); }
REF_API float ref_csr_get_element(void* h, int r, int c) { return csr_get_element(static_cast<CSRMatrix*>(h), r, c); }
REF_API int ref_csr_to_gpu(void* h) { return csr_to_gpu(static_cast<CSRMatrix*>(h)); }
REF_API int ref_csr_serialize(void* h, const char* f) { return csr_serialize(static_cast<CSRMatrix*>(h), f); }
REF_API int ref_csr_deserialize(void* h, const char* f) { return csr_deserialize(static_cast<CSRMatrix*>(h), f); }

REF_API void ref_csr_stats(void* h, float* avg, int* mx, int* mn, float* skew) {
    CSRStats s = csr_compute_stats(static_cast<CSRMatrix*>(h));
    *avg = s.avg_nnz_per_row; *mx = s.max_nnz_per_row; *mn = s.min_nnz_per_row; *skew = s.skewness;
}

REF_API void ref_auto_config(void* h, int* kernel_type, int* block_size, int* use_texture) {
    SpMVConfig c = spmv_auto_config(static_cast<CSRMatrix*>(h));
    *kernel_type = static_cast<int>(c.kernel_type); *block_size = c.block_size;
    *use_texture = c.use_texture ? 1 : 0;
}

// ---- ELL handles -----------------------------------------------------------

REF_API void* ref_ell_from_dense(const float* dense, int rows, int cols, int* status) {
    ELLMatrix* e = ell_create(0, 0, 0);
    int rc = ell_from_dense(e, dense, rows, cols);
    if (status) *status = rc;
    return e;
}

REF_API void* ref_ell_from_csr(void* csr, int* status) {
    ELLMatrix* e = ell_create(0, 0, 0);
    int rc = ell_from_csr(e, static_cast<CSRMatrix*>(csr));
    if (status) *status = rc;
    return e;
}

REF_API void ref_ell_fields(void* h, int* rows, int* cols, int* width, float** values,
                            int** col_indices) {
    ELLMatrix* e = static_cast<ELLMatrix*>(h);
    *rows = e->num_rows; *cols = e->num_cols; *width = e->max_nnz_per_row;
    *values = e->values; *col_indices = e->col_indices;
}

REF_API void ref_ell_destroy(void* h) { ell_destroy(static_cast<ELLMatrix*>(h)); }
REF_API int ref_ell_to_dense(void* h, float* dense) { return ell_to_dense(static_cast<ELLMatrix*>(h), dense); }
REF_API float ref_ell_get_element(void* h, int r, int c) { return ell_get_element(static_cast<ELLMatrix*>(h), r, c); }
REF_API int ref_ell_to_gpu(void* h) { return ell_to_gpu(static_cast<ELLMatrix*>(h)); }
REF_API int ref_ell_serialize(void* h, const char* f) { return ell_serialize(static_cast<ELLMatrix*>(h), f); }
REF_API int ref_ell_deserialize(void* h, const char* f) { return ell_deserialize(static_cast<ELLMatrix*>(h), f); }

// ---- CPU SpMV (the ground-truth oracle) -------------------------------------

REF_API void ref_spmv_cpu_csr(void* h, const float* x, float* y) { spmv_cpu_csr(static_cast<CSRMatrix*>(h), x, y); }
REF_API void ref_spmv_cpu_ell(void* h, const float* x, float* y) { spmv_cpu_ell(static_cast<ELLMatrix*>(h), x, y); }

// timed repetitions of the reference CPU path (steady_clock, seconds per call)
REF_API double ref_time_spmv_cpu_csr(void* h, const float* x, float* y, int reps) {
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) spmv_cpu_csr(static_cast<CSRMatrix*>(h), x, y);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count() / (reps > 0 ? reps : 1);
}
REF_API double ref_time_spmv_cpu_ell(void* h, const float* x, float* y, int reps) {
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) spmv_cpu_ell(static_cast<ELLMatrix*>(h), x, y);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count() / (reps > 0 ? reps : 1);
}

// ---- reference CUDA path (secondary oracle; needs a GPU) ---------------------

REF_API int ref_spmv_csr_gpu(void* h, const float* d_x, float* d_y, int kernel_type,
                             int block_size, int use_texture, int vec_size,
                             float* elapsed_ms, float* gflops, float* bandwidth) {
    SpMVConfig c;
    c.kernel_type = static_cast<SpMVConfig::KernelType>(kernel_type);
    c.block_size = block_size; c.use_texture = use_texture != 0;
    SpMVResult r = spmv_csr(static_cast<CSRMatrix*>(h), d_x, d_y, &c, vec_size);
    if (elapsed_ms) *elapsed_ms = r.elapsed_ms;
    if (gflops) *gflops = r.gflops;
    if (bandwidth) *bandwidth = r.bandwidth_gb_s;
    return r.error_code;
}

REF_API int ref_spmv_ell_gpu(void* h, const float* d_x, float* d_y, int vec_size,
                             float* elapsed_ms, float* gflops, float* bandwidth) {
    SpMVResult r = spmv_ell(static_cast<ELLMatrix*>(h), d_x, d_y, nullptr, vec_size);
    if (elapsed_ms) *elapsed_ms = r.elapsed_ms;
    if (gflops) *gflops = r.gflops;
    if (bandwidth) *bandwidth = r.bandwidth_gb_s;
    return r.error_code;
}

REF_API int ref_pagerank(void* h, float damping, float tolerance, int max_iterations,
                         float* ranks_out, float* residual, int* converged) {
    PageRankConfig c; c.damping_factor = damping; c.tolerance = tolerance;
    c.max_iterations = max_iterations;
    CSRMatrix* m = static_cast<CSRMatrix*>(h);
    PageRankResult r = pagerank(m, &c);
    if (r.ranks && ranks_out) std::memcpy(ranks_out, r.ranks, sizeof(float) * m->num_rows);
    if (residual) *residual = r.final_residual;
    if (converged) *converged = r.converged ? 1 : 0;
    int it = r.iterations;
    pagerank_free(&r);
    return it;
}

REF_API void ref_pagerank_top_k(const float* ranks, int n, int k, int* ids, float* vals) {
    PageRankResult r; r.ranks = const_cast<float*>(ranks);
    std::vector<TopKNode> out(k > 0 ? k : 1);
    pagerank_top_k(&r, n, k, out.data());
    int kk = k < n ? k : n;
    for (int i = 0; i < kk; i++) { ids[i] = out[i].node_id; vals[i] = out[i].rank; }
}

// ---- bandwidth model ---------------------------------------------------------

REF_API void ref_bandwidth_csr(void* h, float ms, float* theo, float* ach, float* eff) {
    BandwidthMetrics b = compute_bandwidth_csr(static_cast<CSRMatrix*>(h), ms);
    *theo = b.theoretical_bandwidth_gb_s; *ach = b.achieved_bandwidth_gb_s; *eff = b.efficiency;
}
REF_API void ref_bandwidth_ell(void* h, float ms, float* theo, float* ach, float* eff) {
    BandwidthMetrics b = compute_bandwidth_ell(static_cast<ELLMatrix*>(h), ms);
    *theo = b.theoretical_bandwidth_gb_s; *ach = b.achieved_bandwidth_gb_s; *eff = b.efficiency;
}

// ---- benchmark JSON ------------------------------------------------------------

REF_API int ref_benchmark_to_json(const char* name, const float* f7, int num_runs, char* out, int cap) {
    BenchmarkResult r; r.name = name;
    r.execution_time_ms = f7[0]; r.gflops = f7[1]; r.bandwidth_gb_s = f7[2];
    r.avg_time_ms = f7[3]; r.min_time_ms = f7[4]; r.max_time_ms = f7[5]; r.stddev_time_ms = f7[6];
    r.num_runs = num_runs;
    std::string s = benchmark_to_json(r);
    if (static_cast<int>(s.size()) + 1 > cap) return -1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return static_cast<int>(s.size());
}

REF_API void ref_benchmark_from_json(const char* json, float* f7, int* num_runs) {
    BenchmarkResult r = benchmark_from_json(json);
    f7[0] = r.execution_time_ms; f7[1] = r.gflops; f7[2] = r.bandwidth_gb_s;
    f7[3] = r.avg_time_ms; f7[4] = r.min_time_ms; f7[5] = r.max_time_ms; f7[6] = r.stddev_time_ms;
    *num_runs = r.num_runs;
}

// ---- replay of the reference's own test fixtures -------------------------------

// One persistent generator, as the gtest fixture holds one per test
// (tests/test_spmv.cu:14 `RandomGenerator rng{42}`).
static test::RandomGenerator* g_rng = nullptr;

REF_API void ref_rng_seed(unsigned seed) {
    delete g_rng;
    g_rng = new test::RandomGenerator(seed);
}
REF_API int ref_rng_int(int lo, int hi) { return g_rng->randInt(lo, hi); }
REF_API float ref_rng_float(float lo, float hi) { return g_rng->randFloat(lo, hi); }

// generateRandomDenseMatrix, include/spmv/test_utils.h:35-46
REF_API void ref_rng_dense(int rows, int cols, float density, float lo, float hi, float* out) {
    std::vector<float> m = test::generateRandomDenseMatrix(rows, cols, density, *g_rng, lo, hi);
    std::memcpy(out, m.data(), sizeof(float) * m.size());
}
// generateRandomVector, include/spmv/test_utils.h:49-58
REF_API void ref_rng_vector(int n, float lo, float hi, float* out) {
    std::vector<float> v = test::generateRandomVector(n, *g_rng, lo, hi);
    std::memcpy(out, v.data(), sizeof(float) * v.size());
}
