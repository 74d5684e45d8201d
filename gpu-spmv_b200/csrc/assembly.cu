// assembly.cu -- device-side CSR assembly: the step BEFORE the hot path at scale
// (SURVEY 8f rank 1).  The reference can only build a CSR matrix from a dense
// host array (csr_from_dense, src/csr_matrix.cpp:50-95: O(rows x cols), unusable
// beyond ~50k x 50k) and has no graph input at all; PageRank expects a
// column-normalised adjacency matrix that the caller must produce somehow
// (include/spmv/pagerank.h:28).  Here:
//
//   csr_from_coo_device       (row, col, value) triplets in device memory ->
//                             device CSR sorted by (row, col); duplicates are
//                             kept, in input order (stable), as csr_to_dense /
//                             spmv_cpu_csr accept them.
//   csr_normalize_columns_device   values[j] /= sum of column col[j]  (columns
//                             that sum to 0 are left alone: dangling nodes)
//
// The sort is CUB's radix sort (a library call, set-up only -- not on the SpMV
// path); keys are (row << 32 | col) so only the bits that can be set are sorted.
// The result layout is exactly csr_to_gpu's (include/spmv/csr_matrix.h:11-28):
// row_ptrs i32[rows+1], col_indices i32[nnz], values f32[nnz].
#include "device_utils.cuh"
#include "internal.hpp"

#include <cub/device/device_radix_sort.cuh>

namespace spmv {
namespace b200 {
namespace {

constexpr int kBlock = 256;

inline unsigned grid_for(long long n) {
    long long b = (n + kBlock - 1) / kBlock;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b < 148 * 32 ? b : 148 * 32);
}

__global__ void coo_keys_kernel(long long n, int rows, int cols, const int* __restrict__ r, const int* __restrict__ c,
                                unsigned long long* __restrict__ keys, int* __restrict__ bad) {
    for (long long j = blockIdx.x * static_cast<long long>(kBlock) + threadIdx.x; j < n;
         j += static_cast<long long>(gridDim.x) * kBlock) {
        const int rr = r[j], cc = c[j];
        if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) *bad = 1;
        keys[j] = (static_cast<unsigned long long>(static_cast<unsigned>(rr)) << 32) | static_cast<unsigned>(cc);
    }
}

// row_ptrs[q] = first sorted entry whose row is >= q
__global__ void coo_row_ptrs_kernel(long long n, int rows, const unsigned long long* __restrict__ keys,
                                    int* __restrict__ row_ptrs, int* __restrict__ col_indices) {
    for (long long j = blockIdx.x * static_cast<long long>(kBlock) + threadIdx.x; j <= n;
         j += static_cast<long long>(gridDim.x) * kBlock) {
        const int prev = j == 0 ? -1 : static_cast<int>(keys[j - 1] >> 32);
        const int cur = j == n ? rows : static_cast<int>(keys[j] >> 32);
        for (int q = prev + 1; q <= cur; ++q) row_ptrs[q] = static_cast<int>(j);
        if (j < n) col_indices[j] = static_cast<int>(keys[j] & 0xffffffffull);
    }
}

__global__ void divide_by_colsum_kernel(int nnz, int cols, const int* __restrict__ col_indices,
                                        const double* __restrict__ colsum, float* __restrict__ values) {
    for (long long j = blockIdx.x * static_cast<long long>(kBlock) + threadIdx.x; j < nnz;
         j += static_cast<long long>(gridDim.x) * kBlock) {
        const int c = col_indices[j];
        if (c < 0 || c >= cols) continue;
        const float s = static_cast<float>(__ldg(colsum + c));
        if (s != 0.0f) values[j] = __fdiv_rn(values[j], s);
    }
}

int bits_for(int n) {  // bits needed to represent values in [0, n)
    int b = 1;
    while (b < 32 && (1ll << b) < n) ++b;
    return b;
}

}  // namespace

int csr_from_coo_device(CSRMatrix* out, int rows, int cols, long long n, const int* d_rows, const int* d_cols,
                        const float* d_vals) {
    if (!out || rows < 0 || cols < 0 || n < 0 || n > 0x7fffffffll) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    if (n > 0 && (!d_rows || !d_cols || !d_vals)) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    cudaStream_t stream = nullptr;
    unsigned long long *keys_a = nullptr, *keys_b = nullptr;
    float *vals_b = nullptr, *d_values = nullptr;
    int *d_col = nullptr, *d_rp = nullptr, *d_bad = nullptr;
    void* temp = nullptr;
    auto cleanup = [&]() {
        cudaFree(keys_a); cudaFree(keys_b); cudaFree(vals_b); cudaFree(temp); cudaFree(d_bad);
    };
    auto fail = [&](SpMVError code) {
        cudaGetLastError();
        cleanup();
        cudaFree(d_values); cudaFree(d_col); cudaFree(d_rp);
        return static_cast<int>(code);
    };
    const size_t un = static_cast<size_t>(n);
    if (cudaMalloc(&d_rp, sizeof(int) * (static_cast<size_t>(rows) + 1)) != cudaSuccess) return fail(SpMVError::CUDA_MALLOC);
    if (n > 0) {
        if (cudaMalloc(&keys_a, 8 * un) != cudaSuccess || cudaMalloc(&keys_b, 8 * un) != cudaSuccess ||
            cudaMalloc(&vals_b, 4 * un) != cudaSuccess || cudaMalloc(&d_values, 4 * un) != cudaSuccess ||
            cudaMalloc(&d_col, 4 * un) != cudaSuccess || cudaMalloc(&d_bad, sizeof(int)) != cudaSuccess)
            return fail(SpMVError::CUDA_MALLOC);
        cudaMemsetAsync(d_bad, 0, sizeof(int), stream);
        coo_keys_kernel<<<grid_for(n), kBlock, 0, stream>>>(n, rows, cols, d_rows, d_cols, keys_a, d_bad);
        if (cudaMemcpyAsync(d_values, d_vals, 4 * un, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
            return fail(SpMVError::CUDA_MEMCPY);
        int bad = 0;
        if (cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(SpMVError::CUDA_MEMCPY);
        if (bad) return fail(SpMVError::INVALID_ARGUMENT);  // an index outside the matrix
        cub::DoubleBuffer<unsigned long long> kb(keys_a, keys_b);
        cub::DoubleBuffer<float> vb(d_values, vals_b);
        const int end_bit = 32 + bits_for(rows);
        size_t temp_bytes = 0;
        if (cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, kb, vb, static_cast<int>(n), 0, end_bit, stream) != cudaSuccess)
            return fail(SpMVError::KERNEL_LAUNCH);
        if (cudaMalloc(&temp, temp_bytes ? temp_bytes : 16) != cudaSuccess) return fail(SpMVError::CUDA_MALLOC);
        if (cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kb, vb, static_cast<int>(n), 0, end_bit, stream) != cudaSuccess)
            return fail(SpMVError::KERNEL_LAUNCH);
        if (vb.Current() != d_values) {  // keep the result in the buffer we hand out
            float* t = d_values;
            d_values = vals_b;
            vals_b = t;
        }
        coo_row_ptrs_kernel<<<grid_for(n + 1), kBlock, 0, stream>>>(n, rows, kb.Current(), d_rp, d_col);
        count_launches(2);
    } else {
        cudaMemsetAsync(d_rp, 0, sizeof(int) * (static_cast<size_t>(rows) + 1), stream);
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return fail(SpMVError::KERNEL_LAUNCH);
    cleanup();

    // hand the arrays to the struct with csr_to_gpu's ownership rules (src/csr_matrix.cpp:138-165);
    // host arrays are re-allocated to the new size so that csr_from_gpu can fill them
    csr_free_gpu(out);
    if (out->owns_host_memory) {
        delete[] out->values;
        delete[] out->col_indices;
        delete[] out->row_ptrs;
    }
    out->num_rows = rows;
    out->num_cols = cols;
    out->nnz = static_cast<int>(n);
    out->values = n > 0 ? new float[un] : nullptr;
    out->col_indices = n > 0 ? new int[un] : nullptr;
    out->row_ptrs = new int[static_cast<size_t>(rows) + 1]();
    out->owns_host_memory = true;
    out->d_values = d_values;
    out->d_col_indices = d_col;
    out->d_row_ptrs = d_rp;
    out->owns_device_memory = true;
    note_device_csr(out);
    return 0;
}

int csr_normalize_columns_device(CSRMatrix* A) {
    if (!A) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    if (A->nnz <= 0 || A->num_cols <= 0) return 0;
    if (!A->d_col_indices || !A->d_values || !A->d_row_ptrs) return static_cast<int>(SpMVError::INVALID_FORMAT);
    cudaStream_t stream = nullptr;
    double* d_colsum = nullptr;
    if (cudaMalloc(&d_colsum, sizeof(double) * static_cast<size_t>(A->num_cols)) != cudaSuccess) {
        cudaGetLastError();
        return static_cast<int>(SpMVError::CUDA_MALLOC);
    }
    cudaMemsetAsync(d_colsum, 0, sizeof(double) * static_cast<size_t>(A->num_cols), stream);
    cudaError_t e = launch_colsum(view_of(A), d_colsum, stream);
    divide_by_colsum_kernel<<<grid_for(A->nnz), kBlock, 0, stream>>>(A->nnz, A->num_cols, A->d_col_indices, d_colsum,
                                                                     A->d_values);
    count_launches(1);
    const cudaError_t s = cudaStreamSynchronize(stream);
    cudaFree(d_colsum);
    if (e != cudaSuccess || s != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        cudaGetLastError();
        return static_cast<int>(SpMVError::KERNEL_LAUNCH);
    }
    return 0;
}

}  // namespace b200
}  // namespace spmv
