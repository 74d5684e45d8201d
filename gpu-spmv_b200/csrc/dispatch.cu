// dispatch.cu -- the device entry points of the API: spmv_csr / spmv_ell
// (reference src/spmv_kernels.cu:215-326, :328-420), their stream-ordered
// twins, and the kernel dispatch they share with the benchmark harness and
// PageRank.
//
// Contract kept from the reference launchers: argument checks in the same
// order with the same codes (null -> INVALID_ARGUMENT, vec_size mismatch ->
// INVALID_DIMENSION, missing device arrays -> INVALID_FORMAT), config == NULL
// -> {SCALAR_CSR, 256, false}, unknown kernel_type -> scalar, launch failure
// -> KERNEL_LAUNCH, and on success elapsed_ms (CUDA events around every kernel
// the call launches, merge-path partition and fix-up included), gflops =
// 2*nnz/(ms*1e6), bandwidth from the compulsory-bytes model, y = d_y.
// Differences ("launcher"): events are created once per thread and
// reused; use_texture and block_size are accepted and ignored; a matrix with
// zero rows succeeds with no launch (the reference reports KERNEL_LAUNCH for
// its 0-block grid); spmv_ell counts true non-zeros on the device instead of
// re-walking the host arrays after every call.
#include "internal.hpp"

#include <atomic>
#include <mutex>
#include <unordered_map>

namespace spmv {
namespace b200 {

// ---- launch accounting ----------------------------------------------------------
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(static_cast<unsigned long long>(n), std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ---- scratch ----------------------------------------------------------------------
Scratch::~Scratch() { release(); }
void Scratch::release() {
    if (ptr_) cudaFree(ptr_);
    ptr_ = nullptr;
    cap_ = 0;
}
void* Scratch::reserve(size_t bytes) {
    if (bytes <= cap_) return ptr_;
    if (ptr_) cudaFree(ptr_);
    ptr_ = nullptr;
    cap_ = 0;
    const size_t want = bytes + bytes / 4;  // head-room so slightly larger inputs do not reallocate
    if (cudaMalloc(&ptr_, want) != cudaSuccess) {
        cudaGetLastError();
        ptr_ = nullptr;
        return nullptr;
    }
    cap_ = want;
    return ptr_;
}

cudaError_t dispatch_csr(const CsrView& A, const float* x, float* y, int kernel_type, Scratch& scratch,
                         cudaStream_t stream) {
    if (A.rows <= 0) return cudaSuccess;
    switch (kernel_type) {
        case SpMVConfig::MERGE_PATH: {
            if (A.nnz <= 0) return launch_csr_stream(A, x, y, 1, stream);  // writes zeros
            void* block = scratch.reserve(merge_plan_bytes(A.rows, A.nnz, false));
            if (!block) return cudaErrorMemoryAllocation;
            const MergePlan plan = merge_plan_carve(block, A.rows, A.nnz, false);
            cudaError_t e = launch_merge_partition(A, plan, stream);
            if (e != cudaSuccess) return e;
            return launch_merge_spmv(A, x, y, plan, stream);
        }
        case SpMVConfig::VECTOR_CSR:
            return launch_csr_vector(A, x, y, stream);
        case SpMVConfig::SCALAR_CSR:
        default:
            return launch_csr_stream(A, x, y, 1, stream);
    }
}

namespace {

// Per host thread and device: the two timing events, a merge-path scratch block
// and the 8-byte non-zero counter used by spmv_ell.
struct ThreadCtx {
    cudaEvent_t start = nullptr, stop = nullptr;
    Scratch scratch;
    unsigned long long* d_counter = nullptr;
    bool ok = false;
};

ThreadCtx* thread_ctx() {
    thread_local std::unordered_map<int, ThreadCtx*> per_device;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    auto it = per_device.find(dev);
    if (it != per_device.end()) return it->second->ok ? it->second : nullptr;
    ThreadCtx* c = new ThreadCtx();  // lives for the process (device resources outlive static dtors safely)
    c->ok = cudaEventCreate(&c->start) == cudaSuccess && cudaEventCreate(&c->stop) == cudaSuccess &&
            cudaMalloc(&c->d_counter, sizeof(unsigned long long)) == cudaSuccess;
    if (!c->ok) cudaGetLastError();
    per_device[dev] = c;
    return c->ok ? c : nullptr;
}

SpMVResult failed(SpMVError code) {
    SpMVResult r;
    r.error_code = static_cast<int>(code);
    return r;
}

// async stream scratch: one block per (device, stream), guarded by a mutex
std::mutex g_async_mu;
std::unordered_map<unsigned long long, Scratch*> g_async_scratch;

Scratch* async_scratch(cudaStream_t stream) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long key = (static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(stream)) << 6) ^
                                   static_cast<unsigned long long>(dev);
    std::lock_guard<std::mutex> lock(g_async_mu);
    auto it = g_async_scratch.find(key);
    if (it != g_async_scratch.end()) return it->second;
    Scratch* s = new Scratch();
    g_async_scratch[key] = s;
    return s;
}

}  // namespace

int spmv_csr_async(const CSRMatrix* A, const float* d_x, float* d_y, const SpMVConfig* config,
                   cudaStream_t stream) {
    if (!A || !d_x || !d_y) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    if (!A->d_row_ptrs || !A->d_col_indices || (A->nnz > 0 && !A->d_values))
        return static_cast<int>(SpMVError::INVALID_FORMAT);
    const int kernel = config ? static_cast<int>(config->kernel_type) : static_cast<int>(SpMVConfig::SCALAR_CSR);
    cudaError_t e = dispatch_csr(view_of(A), d_x, d_y, kernel, *async_scratch(stream), stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return static_cast<int>(e == cudaErrorMemoryAllocation ? SpMVError::CUDA_MALLOC : SpMVError::KERNEL_LAUNCH);
    }
    return 0;
}

int spmv_ell_async(const ELLMatrix* A, const float* d_x, float* d_y, cudaStream_t stream) {
    if (!A || !d_x || !d_y) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    if (!A->d_col_indices || !A->d_values) return static_cast<int>(SpMVError::INVALID_FORMAT);
    cudaError_t e = launch_ell(A->num_rows, A->max_nnz_per_row, A->d_col_indices, A->d_values, d_x, d_y, nullptr, stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return static_cast<int>(SpMVError::KERNEL_LAUNCH);
    }
    return 0;
}

}  // namespace b200

// ------------------------------------------------------------------- spmv_csr --
SpMVResult spmv_csr(const CSRMatrix* A, const float* d_x, float* d_y, const SpMVConfig* config, int vec_size) {
    if (!A || !d_x || !d_y) return b200::failed(SpMVError::INVALID_ARGUMENT);
    if (vec_size >= 0 && !spmv_validate_dimensions(A->num_cols, vec_size))
        return b200::failed(SpMVError::INVALID_DIMENSION);
    if (!A->d_row_ptrs || !A->d_col_indices || (A->nnz > 0 && !A->d_values))
        return b200::failed(SpMVError::INVALID_FORMAT);

    const int kernel = config ? static_cast<int>(config->kernel_type) : static_cast<int>(SpMVConfig::SCALAR_CSR);
    b200::ThreadCtx* ctx = b200::thread_ctx();
    if (!ctx) return b200::failed(SpMVError::KERNEL_LAUNCH);  // no usable device: fail, never compute on the host

    cudaStream_t stream = nullptr;  // legacy default stream, as the reference
    cudaEventRecord(ctx->start, stream);
    cudaError_t e = b200::dispatch_csr(b200::view_of(A), d_x, d_y, kernel, ctx->scratch, stream);
    cudaEventRecord(ctx->stop, stream);
    cudaError_t sync = cudaEventSynchronize(ctx->stop);
    cudaError_t last = cudaGetLastError();
    if (e != cudaSuccess || sync != cudaSuccess || last != cudaSuccess)
        return b200::failed(SpMVError::KERNEL_LAUNCH);

    SpMVResult r;
    cudaEventElapsedTime(&r.elapsed_ms, ctx->start, ctx->stop);
    r.gflops = (2.0f * A->nnz) / (r.elapsed_ms * 1e6f);
    r.bandwidth_gb_s = compute_bandwidth_csr(A, r.elapsed_ms).achieved_bandwidth_gb_s;
    r.y = d_y;
    r.error_code = static_cast<int>(SpMVError::SUCCESS);
    return r;
}

// ------------------------------------------------------------------- spmv_ell --
SpMVResult spmv_ell(const ELLMatrix* A, const float* d_x, float* d_y, const SpMVConfig* /*config*/, int vec_size) {
    if (!A || !d_x || !d_y) return b200::failed(SpMVError::INVALID_ARGUMENT);
    if (vec_size >= 0 && !spmv_validate_dimensions(A->num_cols, vec_size))
        return b200::failed(SpMVError::INVALID_DIMENSION);
    if (!A->d_col_indices || !A->d_values) return b200::failed(SpMVError::INVALID_FORMAT);

    b200::ThreadCtx* ctx = b200::thread_ctx();
    if (!ctx) return b200::failed(SpMVError::KERNEL_LAUNCH);

    cudaStream_t stream = nullptr;
    cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), stream);
    cudaEventRecord(ctx->start, stream);
    cudaError_t e = b200::launch_ell(A->num_rows, A->max_nnz_per_row, A->d_col_indices, A->d_values, d_x, d_y,
                                     ctx->d_counter, stream);
    cudaEventRecord(ctx->stop, stream);
    unsigned long long live = 0;
    cudaError_t copy = cudaMemcpyAsync(&live, ctx->d_counter, sizeof(live), cudaMemcpyDeviceToHost, stream);
    cudaError_t sync = cudaStreamSynchronize(stream);
    cudaError_t last = cudaGetLastError();
    if (e != cudaSuccess || copy != cudaSuccess || sync != cudaSuccess || last != cudaSuccess)
        return b200::failed(SpMVError::KERNEL_LAUNCH);

    SpMVResult r;
    cudaEventElapsedTime(&r.elapsed_ms, ctx->start, ctx->stop);
    const int actual_nnz = static_cast<int>(live);  // entries with col >= 0 (reference :399-405)
    r.gflops = (2.0f * actual_nnz) / (r.elapsed_ms * 1e6f);
    r.bandwidth_gb_s = compute_bandwidth_ell(A, r.elapsed_ms).achieved_bandwidth_gb_s;
    r.y = d_y;
    r.error_code = static_cast<int>(SpMVError::SUCCESS);
    return r;
}

}  // namespace spmv
