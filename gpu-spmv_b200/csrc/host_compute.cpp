// host_compute.cpp -- the host-only pieces of the compute layer:
//   * spmv_cpu_csr / spmv_cpu_ell: the API's sequential fp32 host SpMV
//     (src/spmv_cpu.cpp:6-32).  These are public API functions; the device
//     path never calls them (there is no CPU fallback anywhere).
//   * the kernel selector (src/spmv_cpu.cpp:34-50) and its B200 policy.
//   * the compulsory-bytes bandwidth model (src/bandwidth.cpp).
//   * pagerank_top_k / pagerank_free (src/pagerank.cu:155-185), which operate
//     on the host rank array by definition of the API.
#include "internal.hpp"

#include <algorithm>
#include <cmath>
#include <mutex>

namespace spmv {

// --------------------------------------------------------------- host SpMV --
// Build note: this translation unit is compiled with -ffp-contract=off so the
// multiply and the add round separately, as in the reference's x86-64 build.

void spmv_cpu_csr(const CSRMatrix* A, const float* x, float* y) {
    if (!A || !x || !y) return;
    const int* rp = A->row_ptrs;
    for (int r = 0; r < A->num_rows; ++r) {
        float acc = 0.0f;
        for (int p = rp[r]; p < rp[r + 1]; ++p) acc += A->values[p] * x[A->col_indices[p]];
        y[r] = acc;
    }
}

void spmv_cpu_ell(const ELLMatrix* A, const float* x, float* y) {
    if (!A || !x || !y) return;
    for (int r = 0; r < A->num_rows; ++r) {
        float acc = 0.0f;
        for (int k = 0; k < A->max_nnz_per_row; ++k) {
            const int slot = ell_index(r, k, A->num_rows);
            const int c = A->col_indices[slot];
            if (c >= 0) acc += A->values[slot] * x[c];
        }
        y[r] = acc;
    }
}

// ---------------------------------------------------------------- selector --

namespace b200 {

// The reference decision (src/spmv_cpu.cpp:41-47) on already-computed stats.
SpMVConfig::KernelType reference_kernel_choice(const CSRStats& s) {
    if (s.avg_nnz_per_row < 4.0f) return SpMVConfig::SCALAR_CSR;
    if (s.skewness < 10.0f) return SpMVConfig::VECTOR_CSR;
    return SpMVConfig::MERGE_PATH;
}

SpMVConfig reference_policy(const CSRMatrix* A) {
    SpMVConfig cfg;
    cfg.block_size = 256;
    cfg.use_texture = A->num_cols > 10000;
    cfg.kernel_type = reference_kernel_choice(csr_compute_stats(A));
    return cfg;
}

}  // namespace b200

// Reference policy plus one documented B200 divergence (selector):
// a short-row matrix (avg < 4) that also holds an outlier row longer than
// kOutlierRowNnz would serialise one CTA of the row-owner kernel on that row,
// so it is routed to merge-path instead (BASELINE config 3).
SpMVConfig spmv_auto_config(const CSRMatrix* A) {
    SpMVConfig cfg;
    cfg.block_size = 256;
    cfg.use_texture = A->num_cols > 10000;
    const CSRStats s = csr_compute_stats(A);
    cfg.kernel_type = b200::reference_kernel_choice(s);
    if (cfg.kernel_type == SpMVConfig::SCALAR_CSR && s.max_nnz_per_row > b200::kOutlierRowNnz)
        cfg.kernel_type = SpMVConfig::MERGE_PATH;
    return cfg;
}

// --------------------------------------------------------- bandwidth model --

namespace b200 {

// src/bandwidth.cpp:34-42
size_t csr_compulsory_bytes(const CSRMatrix* A) {
    size_t rd = 0;
    rd += A->nnz * sizeof(float);
    rd += A->nnz * sizeof(int);
    rd += (A->num_rows + 1) * sizeof(int);
    rd += A->num_cols * sizeof(float);
    return rd + A->num_rows * sizeof(float);
}

// src/bandwidth.cpp:66-75 (padding slots are counted)
size_t ell_compulsory_bytes(const ELLMatrix* A) {
    const size_t slots = static_cast<size_t>(A->num_rows) * A->max_nnz_per_row;
    return slots * (sizeof(float) + sizeof(int)) + A->num_cols * sizeof(float) +
           A->num_rows * sizeof(float);
}

static BandwidthMetrics metrics_for(size_t bytes, float elapsed_ms) {
    BandwidthMetrics m;
    const float seconds = elapsed_ms / 1000.0f;
    m.achieved_bandwidth_gb_s = (bytes / 1e9f) / seconds;   // fp32, as src/bandwidth.cpp:45-46
    m.theoretical_bandwidth_gb_s = get_gpu_peak_bandwidth();
    if (m.theoretical_bandwidth_gb_s > 0.0f)
        m.efficiency = std::min(m.achieved_bandwidth_gb_s / m.theoretical_bandwidth_gb_s, 1.0f);
    return m;
}

}  // namespace b200

// Theoretical peak = memory clock x bus width x 2 (DDR), GB/s, the formula of
// src/bandwidth.cpp:7-20 -- read through attributes of the CURRENT device and
// cached per device (the reference re-queries device 0's full property struct
// on every call).
float get_gpu_peak_bandwidth() {
    static std::mutex mu;
    static float cached[64];
    static bool have[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0.0f;
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev]) {
        int clock_khz = 0, bus_bits = 0;
        if (cudaDeviceGetAttribute(&clock_khz, cudaDevAttrMemoryClockRate, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&bus_bits, cudaDevAttrGlobalMemoryBusWidth, dev) != cudaSuccess)
            return 0.0f;
        const float khz = static_cast<float>(clock_khz);
        const float bits = static_cast<float>(bus_bits);
        cached[dev] = (khz * 1000.0f) * (bits / 8.0f) * 2.0f / 1e9f;
        have[dev] = true;
    }
    return cached[dev];
}

BandwidthMetrics compute_bandwidth_csr(const CSRMatrix* A, float elapsed_ms) {
    if (!A || elapsed_ms <= 0.0f) return BandwidthMetrics();
    return b200::metrics_for(b200::csr_compulsory_bytes(A), elapsed_ms);
}

BandwidthMetrics compute_bandwidth_ell(const ELLMatrix* A, float elapsed_ms) {
    if (!A || elapsed_ms <= 0.0f) return BandwidthMetrics();
    return b200::metrics_for(b200::ell_compulsory_bytes(A), elapsed_ms);
}

// ------------------------------------------------------------ top-k / free --

void pagerank_free(PageRankResult* result) {
    if (!result || !result->ranks) return;
    delete[] result->ranks;
    result->ranks = nullptr;
}

// Highest-rank k nodes, descending; order among equal ranks is unspecified
// (as with the reference's partial_sort).  nth_element + sort of the head keeps
// the cost O(n + k log k) for the 2^26-node case.
void pagerank_top_k(const PageRankResult* result, int num_nodes, int k, TopKNode* top_k) {
    if (!result || !result->ranks || !top_k || k <= 0 || num_nodes <= 0) return;
    const int keep = std::min(k, num_nodes);
    std::vector<TopKNode> all(static_cast<size_t>(num_nodes));
    for (int i = 0; i < num_nodes; ++i) {
        all[i].node_id = i;
        all[i].rank = result->ranks[i];
    }
    auto higher = [](const TopKNode& a, const TopKNode& b) { return a.rank > b.rank; };
    if (keep < num_nodes) std::nth_element(all.begin(), all.begin() + keep, all.end(), higher);
    std::sort(all.begin(), all.begin() + keep, higher);
    std::copy(all.begin(), all.begin() + keep, top_k);
}

}  // namespace spmv
