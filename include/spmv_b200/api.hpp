// spmv_b200/api.hpp -- C++ surface of the B200-native SpMV library.
//
// This single header declares every type and free function of the
// LessUp/gpu-spmv C++ API (namespace spmv) so that code written against the
// reference relinks unchanged: identical names, argument meaning, struct
// layouts (static_asserts at the bottom; x86-64 SysV) and therefore identical
// mangled symbols.  The per-topic headers the reference ships
// (include/spmv/{common,cuda_buffer,csr_matrix,ell_matrix,spmv,bandwidth,
// pagerank,benchmark}.h) exist here as thin forwarders to this file.
//
// Reference interface replaced, by section:
//   errors / CudaBuffer   include/spmv/common.h:13-67, include/spmv/cuda_buffer.h:13-101
//   CSR storage           include/spmv/csr_matrix.h:11-71
//   ELL storage           include/spmv/ell_matrix.h:13-66
//   SpMV + selector       include/spmv/spmv.h:11-54
//   bandwidth model       include/spmv/bandwidth.h:10-27
//   PageRank              include/spmv/pagerank.h:9-43
//   benchmark harness     include/spmv/benchmark.h:13-78
//
// The plain-C twin of this surface (the FFI boundary) is include/spmv_b200.h.
#ifndef SPMV_B200_API_HPP
#define SPMV_B200_API_HPP

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace spmv {

// ---------------------------------------------------------------------------
// Status codes.  Functions returning int give 0 or one of these (negative).
// ---------------------------------------------------------------------------
enum class SpMVError {
    SUCCESS = 0,
    INVALID_DIMENSION = -1,
    CUDA_MALLOC = -2,
    CUDA_MEMCPY = -3,
    KERNEL_LAUNCH = -4,
    INVALID_FORMAT = -5,
    FILE_IO = -6,
    OUT_OF_MEMORY = -7,
    INVALID_ARGUMENT = -8
};

// Text for a status code; the wording is part of the contract
// (reference tests/test_common.cpp:8-18 compares it verbatim).
inline const char* spmv_error_string(SpMVError code) {
    switch (code) {
        case SpMVError::SUCCESS:           return "Success";
        case SpMVError::INVALID_DIMENSION: return "Invalid matrix/vector dimension";
        case SpMVError::CUDA_MALLOC:       return "CUDA memory allocation failed";
        case SpMVError::CUDA_MEMCPY:       return "CUDA memory copy failed";
        case SpMVError::KERNEL_LAUNCH:     return "CUDA kernel launch failed";
        case SpMVError::INVALID_FORMAT:    return "Invalid sparse matrix format";
        case SpMVError::FILE_IO:           return "File I/O error";
        case SpMVError::OUT_OF_MEMORY:     return "Out of memory";
        case SpMVError::INVALID_ARGUMENT:  return "Invalid argument";
    }
    return "Unknown error";
}

// Thrown by CudaBuffer and CUDA_CHECK_THROW.
class CudaException : public std::runtime_error {
public:
    explicit CudaException(cudaError_t e)
        : std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e)), error_(e) {}
    cudaError_t error() const { return error_; }
private:
    cudaError_t error_;
};

}  // namespace spmv

// Return-code flavour: logs and returns CUDA_MALLOC whatever failed (that is
// the reference's behaviour, common.h:53-60, and callers test for != 0).
#define CUDA_CHECK(call)                                                            \
    do {                                                                            \
        cudaError_t spmv_cuda_status_ = (call);                                     \
        if (spmv_cuda_status_ != cudaSuccess) {                                     \
            fprintf(stderr, "CUDA error at %s:%d: %s\n", __FILE__, __LINE__,        \
                    cudaGetErrorString(spmv_cuda_status_));                         \
            return static_cast<int>(spmv::SpMVError::CUDA_MALLOC);                  \
        }                                                                           \
    } while (0)

// Throwing flavour.
#define CUDA_CHECK_THROW(call)                                                      \
    do {                                                                            \
        cudaError_t spmv_cuda_status_ = (call);                                     \
        if (spmv_cuda_status_ != cudaSuccess) throw spmv::CudaException(spmv_cuda_status_); \
    } while (0)

namespace spmv {

// ---------------------------------------------------------------------------
// Move-only owner of a device array.
// ---------------------------------------------------------------------------
template <typename T>
class CudaBuffer {
public:
    CudaBuffer() = default;
    explicit CudaBuffer(size_t count) : count_(count) {
        if (count_ == 0) return;
        cudaError_t e = cudaMalloc(&data_, count_ * sizeof(T));
        if (e != cudaSuccess) throw CudaException(e);
    }
    ~CudaBuffer() { drop(); }

    CudaBuffer(const CudaBuffer&) = delete;
    CudaBuffer& operator=(const CudaBuffer&) = delete;

    CudaBuffer(CudaBuffer&& o) noexcept : data_(o.data_), count_(o.count_) {
        o.data_ = nullptr;
        o.count_ = 0;
    }
    CudaBuffer& operator=(CudaBuffer&& o) noexcept {
        if (this != &o) {
            drop();
            data_ = o.data_;
            count_ = o.count_;
            o.data_ = nullptr;
            o.count_ = 0;
        }
        return *this;
    }

    T* get() { return data_; }
    const T* get() const { return data_; }
    size_t size() const { return count_; }
    bool empty() const { return data_ == nullptr || count_ == 0; }

    void copyFromHost(const T* src, size_t count) {
        if (count > count_) throw std::runtime_error("Copy size exceeds buffer size");
        CUDA_CHECK_THROW(cudaMemcpy(data_, src, count * sizeof(T), cudaMemcpyHostToDevice));
    }
    void copyToHost(T* dst, size_t count) const {
        if (count > count_) throw std::runtime_error("Copy size exceeds buffer size");
        CUDA_CHECK_THROW(cudaMemcpy(dst, data_, count * sizeof(T), cudaMemcpyDeviceToHost));
    }

    // Discards contents; allocates new_count elements (no-op when unchanged).
    void resize(size_t new_count) {
        if (new_count == count_) return;
        drop();
        count_ = new_count;
        if (count_ > 0) CUDA_CHECK_THROW(cudaMalloc(&data_, count_ * sizeof(T)));
    }
    void release() {
        drop();
        count_ = 0;
    }

private:
    void drop() {
        if (data_) cudaFree(data_);
        data_ = nullptr;
    }
    T* data_ = nullptr;
    size_t count_ = 0;
};

// ---------------------------------------------------------------------------
// CSR storage.  Host arrays and their device mirrors live side by side; the
// fields are public and callers read/write them directly.
// ---------------------------------------------------------------------------
struct CSRMatrix {
    int num_rows;
    int num_cols;
    int nnz;
    float* values;        // [nnz]          host
    int* col_indices;     // [nnz]          host
    int* row_ptrs;        // [num_rows + 1] host
    float* d_values;      // device mirrors (nullptr until csr_to_gpu)
    int* d_col_indices;
    int* d_row_ptrs;
    bool owns_host_memory;
    bool owns_device_memory;
};

struct CSRStats {
    float avg_nnz_per_row;
    int max_nnz_per_row;
    int min_nnz_per_row;
    float skewness;  // max / (min + 1)
};

CSRMatrix* csr_create(int rows, int cols, int nnz);
void csr_destroy(CSRMatrix* mat);
int csr_from_dense(CSRMatrix* csr, const float* dense, int rows, int cols);  // dense is row-major
int csr_to_dense(const CSRMatrix* csr, float* dense);
float csr_get_element(const CSRMatrix* mat, int row, int col);
int csr_to_gpu(CSRMatrix* mat);
int csr_from_gpu(CSRMatrix* mat);
void csr_free_gpu(CSRMatrix* mat);
int csr_serialize(const CSRMatrix* mat, const char* filename);
int csr_deserialize(CSRMatrix* mat, const char* filename);
CSRStats csr_compute_stats(const CSRMatrix* mat);

// ---------------------------------------------------------------------------
// ELL storage, column-major: entry k of row i sits at k * num_rows + i;
// unused slots carry col = -1, value = 0.
// ---------------------------------------------------------------------------
struct ELLMatrix {
    int num_rows;
    int num_cols;
    int max_nnz_per_row;
    float* values;        // [num_rows * max_nnz_per_row] host
    int* col_indices;     // same shape, -1 marks padding
    float* d_values;
    int* d_col_indices;
    bool owns_host_memory;
    bool owns_device_memory;
};

ELLMatrix* ell_create(int rows, int cols, int max_nnz_per_row);
void ell_destroy(ELLMatrix* mat);
int ell_from_dense(ELLMatrix* ell, const float* dense, int rows, int cols);
int ell_from_csr(ELLMatrix* ell, const CSRMatrix* csr);
int ell_to_dense(const ELLMatrix* ell, float* dense);
float ell_get_element(const ELLMatrix* mat, int row, int col);
int ell_to_gpu(ELLMatrix* mat);
int ell_from_gpu(ELLMatrix* mat);
void ell_free_gpu(ELLMatrix* mat);
int ell_serialize(const ELLMatrix* mat, const char* filename);
int ell_deserialize(ELLMatrix* mat, const char* filename);

inline int ell_index(int row, int k, int num_rows) { return k * num_rows + row; }

// ---------------------------------------------------------------------------
// SpMV.
// ---------------------------------------------------------------------------
struct SpMVConfig {
    enum KernelType {
        SCALAR_CSR,  // one owner thread per row, rows staged block-cooperatively
        VECTOR_CSR,  // several lanes per row + shuffle reduction
        MERGE_PATH,  // (rows + nnz)-balanced two-level merge path
        ELL_KERNEL   // ELL storage only
    };
    KernelType kernel_type;
    int block_size;    // accepted for compatibility; sm_100a kernels fix their own CTA shape
    bool use_texture;  // accepted for compatibility; x is always read via ld.global.nc

    SpMVConfig() : kernel_type(SCALAR_CSR), block_size(256), use_texture(false) {}
};

struct SpMVResult {
    float* y;              // aliases the caller's d_y on success
    float elapsed_ms;      // device time of the kernel(s)
    float gflops;          // 2 * nnz / time
    float bandwidth_gb_s;  // compulsory bytes / time
    int error_code;        // 0 or a SpMVError

    SpMVResult() : y(nullptr), elapsed_ms(0.0f), gflops(0.0f), bandwidth_gb_s(0.0f), error_code(0) {}
};

// Host reference (sequential fp32); part of the API, never used by the GPU path.
void spmv_cpu_csr(const CSRMatrix* A, const float* x, float* y);
void spmv_cpu_ell(const ELLMatrix* A, const float* x, float* y);

// Device SpMV: y = A x with d_x, d_y device pointers; blocking.
SpMVResult spmv_csr(const CSRMatrix* A, const float* d_x, float* d_y,
                    const SpMVConfig* config, int vec_size = -1);
SpMVResult spmv_ell(const ELLMatrix* A, const float* d_x, float* d_y,
                    const SpMVConfig* config, int vec_size = -1);

// Kernel selector driven by measured row-length statistics.
SpMVConfig spmv_auto_config(const CSRMatrix* A);

inline bool spmv_validate_dimensions(int num_cols, int vec_size) { return num_cols == vec_size; }

// ---------------------------------------------------------------------------
// Bandwidth model (compulsory bytes; the algorithmic-bytes definition).
// ---------------------------------------------------------------------------
struct BandwidthMetrics {
    float theoretical_bandwidth_gb_s;
    float achieved_bandwidth_gb_s;
    float efficiency;  // achieved / theoretical, capped at 1

    BandwidthMetrics()
        : theoretical_bandwidth_gb_s(0.0f), achieved_bandwidth_gb_s(0.0f), efficiency(0.0f) {}
};

BandwidthMetrics compute_bandwidth_csr(const CSRMatrix* A, float elapsed_ms);
BandwidthMetrics compute_bandwidth_ell(const ELLMatrix* A, float elapsed_ms);
float get_gpu_peak_bandwidth();

// ---------------------------------------------------------------------------
// PageRank on a column-normalised adjacency matrix (rows = destinations).
// ---------------------------------------------------------------------------
struct PageRankConfig {
    float damping_factor;
    float tolerance;      // on the L2 norm of the rank delta
    int max_iterations;

    PageRankConfig() : damping_factor(0.85f), tolerance(1e-6f), max_iterations(100) {}
};

struct PageRankResult {
    float* ranks;          // host array [num_nodes], release with pagerank_free
    int iterations;
    float final_residual;
    bool converged;

    PageRankResult() : ranks(nullptr), iterations(0), final_residual(0.0f), converged(false) {}
};

struct TopKNode {
    int node_id;
    float rank;
};

PageRankResult pagerank(const CSRMatrix* adj_matrix, const PageRankConfig* config = nullptr);
void pagerank_free(PageRankResult* result);
void pagerank_top_k(const PageRankResult* result, int num_nodes, int k, TopKNode* top_k);

// ---------------------------------------------------------------------------
// Benchmark harness.
// ---------------------------------------------------------------------------
struct BenchmarkResult {
    std::string name;
    float execution_time_ms;  // == avg_time_ms
    float gflops;
    float bandwidth_gb_s;
    float avg_time_ms;
    float min_time_ms;
    float max_time_ms;
    float stddev_time_ms;     // sample (n-1) standard deviation
    int num_runs;

    BenchmarkResult()
        : execution_time_ms(0.0f), gflops(0.0f), bandwidth_gb_s(0.0f), avg_time_ms(0.0f),
          min_time_ms(0.0f), max_time_ms(0.0f), stddev_time_ms(0.0f), num_runs(0) {}
};

struct BenchmarkConfig {
    int num_warmup_runs;
    int num_runs;
    bool compare_cpu;

    BenchmarkConfig() : num_warmup_runs(5), num_runs(20), compare_cpu(true) {}
};

struct ComparisonResult {
    BenchmarkResult gpu_result;
    BenchmarkResult cpu_result;
    float speedup;  // cpu avg / gpu avg

    ComparisonResult() : speedup(0.0f) {}
};

// A must already be on the device; x is a HOST vector of num_cols floats.
BenchmarkResult benchmark_csr(const CSRMatrix* A, const float* x, const SpMVConfig* config,
                              const BenchmarkConfig* bench_config = nullptr);
BenchmarkResult benchmark_ell(const ELLMatrix* A, const float* x,
                              const BenchmarkConfig* bench_config = nullptr);
ComparisonResult compare_gpu_cpu_csr(const CSRMatrix* A, const float* x, const SpMVConfig* config,
                                     const BenchmarkConfig* bench_config = nullptr);

std::string benchmark_to_json(const BenchmarkResult& result);
std::string comparison_to_json(const ComparisonResult& result);
BenchmarkResult benchmark_from_json(const std::string& json);

// ---------------------------------------------------------------------------
// Layout contract.
// ---------------------------------------------------------------------------
static_assert(sizeof(CSRMatrix) == 72 && offsetof(CSRMatrix, values) == 16 &&
              offsetof(CSRMatrix, d_values) == 40 && offsetof(CSRMatrix, owns_host_memory) == 64,
              "CSRMatrix layout");
static_assert(sizeof(ELLMatrix) == 56 && offsetof(ELLMatrix, values) == 16 &&
              offsetof(ELLMatrix, d_values) == 32 && offsetof(ELLMatrix, owns_host_memory) == 48,
              "ELLMatrix layout");
static_assert(sizeof(SpMVConfig) == 12 && sizeof(SpMVResult) == 24 && sizeof(CSRStats) == 16,
              "SpMV struct layout");
static_assert(sizeof(PageRankConfig) == 12 && sizeof(PageRankResult) == 24 && sizeof(TopKNode) == 8,
              "PageRank struct layout");
static_assert(sizeof(BandwidthMetrics) == 12 && sizeof(BenchmarkConfig) == 12, "metric struct layout");

}  // namespace spmv

#endif  // SPMV_B200_API_HPP
