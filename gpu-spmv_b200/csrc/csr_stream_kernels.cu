// csr_stream_kernels.cu -- row-owner CSR SpMV kernels for sm_100a
// (SCALAR_CSR and VECTOR_CSR of the reference API).
//
// Replaces spmv_csr_scalar_kernel / spmv_csr_vector_kernel (reference
// src/spmv_kernels.cu:168-188, :133-165).  The reference lets every row owner
// fetch its own non-zeros, so lanes stride by the row length (scalar) or most
// lanes idle on short rows (vector, 5-nnz rows use 5 of 32 lanes).
//
// csr_stream_kernel<LPR> separates the two jobs:
//   1. STAGE   a CTA owns a window of consecutive rows.  It stages the
//              window's row_ptrs in shared memory with one 1-D TMA bulk copy
//              (cp.async.bulk + mbarrier), then streams the window's
//              contiguous non-zero range with fully coalesced 128-bit loads of
//              values and col_indices (row boundaries are irrelevant here),
//              gathers x and parks the PRODUCTS in shared memory.
//   2. REDUCE  each row is owned by LPR lanes that sum the row's products out
//              of shared memory.
//      LPR == 1 (SCALAR_CSR): one thread, sequential order, separately rounded
//              multiply and add -> bit-identical to spmv_cpu_csr
//              (reference src/spmv_cpu.cpp:6-16).
//      LPR in {2,4,8,16} (VECTOR_CSR): lanes stride the row, shuffle tree.
//
// csr_warp_row_kernel is VECTOR_CSR for long rows (average >= 64): one warp per
// row, 128-bit loads on the 16-byte aligned interior of the row, shuffle tree.
//
// Roofline: HBM.  Algorithmic bytes per launch = 8*nnz + 4*(rows+1) + 4*cols +
// 4*rows (reference src/bandwidth.cpp:34-42).
#include "device_utils.cuh"
#include "internal.hpp"

#include <algorithm>

namespace spmv {
namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kProductCap = 5632;  // products parked per pass (22 KB)

// CTA geometry for LPR lanes per row: `groups` row owners, each holding at most
// `max_rows_per_group` running sums in registers.
template <int LPR>
struct StreamGeom {
    static constexpr int groups = kThreads / LPR;
    static constexpr int max_rows_per_group = (1024 / groups) < 8 ? (1024 / groups) : 8;
    static constexpr int max_window_rows = groups * max_rows_per_group;  // <= 1024
};

// shared-memory layout: [mbarrier 16 B][row window (R + 4) ints][products kProductCap floats]
__host__ __device__ constexpr size_t stream_smem_bytes(int window_rows) {
    return 16 + static_cast<size_t>(window_rows + 4) * sizeof(int) + kProductCap * sizeof(float);
}

template <int LPR>
__global__ void __launch_bounds__(kThreads)
csr_stream_kernel(int rows, int nnz, const int* __restrict__ row_ptrs,
                  const int* __restrict__ col_indices, const float* __restrict__ values,
                  const float* __restrict__ x, float* __restrict__ y, int window_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* s_rp = reinterpret_cast<int*>(smem_raw + 16);
    float* s_prod = reinterpret_cast<float*>(smem_raw + 16 + static_cast<size_t>(window_rows + 4) * sizeof(int));

    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * window_rows;
    const int nr = min(window_rows, rows - r0);  // rows in this window (>= 1)

    // ---- stage row_ptrs[r0 .. r0 + nr] -------------------------------------
    // Full windows go through one TMA bulk copy of (window_rows + 4) ints (the
    // copy size must be a multiple of 16 bytes); ragged / unaligned windows
    // fall back to plain loads.
    const int* gp = row_ptrs + r0;
    const bool bulk = (nr == window_rows) && (r0 + window_rows + 4 <= rows + 1) && dev::aligned16(gp);
    if (bulk) {
        if (tid == 0) {
            dev::mbar_init(bar, 1);
            dev::mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = static_cast<uint32_t>(window_rows + 4) * sizeof(int);
            dev::mbar_arrive_expect_tx(bar, bytes);
            dev::tma_bulk_g2s(s_rp, gp, bytes, bar);
        }
        dev::mbar_wait(bar, 0);
    } else {
        for (int i = tid; i <= nr; i += kThreads) s_rp[i] = row_ptrs[r0 + i];
        __syncthreads();
    }

    const int n0 = s_rp[0];
    const int n1 = s_rp[nr];

    constexpr int kGroups = StreamGeom<LPR>::groups;  // row owners per CTA
    constexpr int kMaxRowsPerGroup = StreamGeom<LPR>::max_rows_per_group;
    const int group = tid / LPR;
    const int lane = tid % LPR;
    float acc[kMaxRowsPerGroup];
#pragma unroll
    for (int i = 0; i < kMaxRowsPerGroup; ++i) acc[i] = 0.0f;

    // ---- passes over the window's non-zero range ----------------------------
    // `base` is 4-aligned so that 128-bit loads of values/col_indices are
    // aligned (when the arrays themselves are 16-byte aligned; any device
    // pointer is legal in the public struct, so that is checked here); slot
    // (j - base) of s_prod holds the product of non-zero j.
    const bool vec_ok = dev::aligned16(values) && dev::aligned16(col_indices);
    for (int base = n0 & ~3; base < n1; base += kProductCap) {
        const int hi = min(base + kProductCap, n1);  // exclusive end of this pass
        const int lo = max(base, n0);
        if (base != (n0 & ~3)) __syncthreads();      // previous pass fully reduced

        for (int j = base + 4 * tid; j < hi; j += 4 * kThreads) {
            float p0, p1, p2, p3;
            if (vec_ok && j >= lo && j + 4 <= hi) {
                const float4 v = dev::ld_stream_f4(values + j);
                const int4 c = dev::ld_stream_i4(col_indices + j);
                p0 = __fmul_rn(v.x, dev::ld_x(x + c.x));
                p1 = __fmul_rn(v.y, dev::ld_x(x + c.y));
                p2 = __fmul_rn(v.z, dev::ld_x(x + c.z));
                p3 = __fmul_rn(v.w, dev::ld_x(x + c.w));
            } else {  // ragged head / tail of the range
                p0 = (j + 0 >= lo && j + 0 < hi) ? __fmul_rn(values[j + 0], dev::ld_x(x + col_indices[j + 0])) : 0.0f;
                p1 = (j + 1 >= lo && j + 1 < hi) ? __fmul_rn(values[j + 1], dev::ld_x(x + col_indices[j + 1])) : 0.0f;
                p2 = (j + 2 >= lo && j + 2 < hi) ? __fmul_rn(values[j + 2], dev::ld_x(x + col_indices[j + 2])) : 0.0f;
                p3 = (j + 3 >= lo && j + 3 < hi) ? __fmul_rn(values[j + 3], dev::ld_x(x + col_indices[j + 3])) : 0.0f;
            }
            *reinterpret_cast<float4*>(s_prod + (j - base)) = make_float4(p0, p1, p2, p3);
        }
        __syncthreads();

        // ---- reduce: row owners pick up their rows' products -----------------
#pragma unroll
        for (int i = 0; i < kMaxRowsPerGroup; ++i) {
            const int r = group + i * kGroups;
            if (r < nr) {
                const int a = max(s_rp[r], lo);
                const int b = min(s_rp[r + 1], hi);
                float s = acc[i];
                for (int j = a + lane; j < b; j += LPR) s = __fadd_rn(s, s_prod[j - base]);
                acc[i] = s;
            }
        }
    }

    // ---- finish: combine lanes, write y ---------------------------------------
#pragma unroll
    for (int i = 0; i < kMaxRowsPerGroup; ++i) {
        float s = acc[i];
#pragma unroll
        for (int d = LPR / 2; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d, LPR);
        const int r = group + i * kGroups;
        if (lane == 0 && r < nr) y[r0 + r] = s;
    }
}

// One warp per row; rows are expected to be long (>= 64 non-zeros on average).
__global__ void __launch_bounds__(kThreads)
csr_warp_row_kernel(int rows, const int* __restrict__ row_ptrs, const int* __restrict__ col_indices,
                    const float* __restrict__ values, const float* __restrict__ x,
                    float* __restrict__ y) {
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int a = row_ptrs[warp];
    const int b = row_ptrs[warp + 1];
    const bool vec_ok = dev::aligned16(values) && dev::aligned16(col_indices);
    float s = 0.0f;
    // 4-element groups starting at the aligned address at or below `a`
    for (int j = (a & ~3) + 4 * lane; j < b; j += 128) {
        if (vec_ok && j >= a && j + 4 <= b) {
            const float4 v = dev::ld_stream_f4(values + j);
            const int4 c = dev::ld_stream_i4(col_indices + j);
            s = fmaf(v.x, dev::ld_x(x + c.x), s);
            s = fmaf(v.y, dev::ld_x(x + c.y), s);
            s = fmaf(v.z, dev::ld_x(x + c.z), s);
            s = fmaf(v.w, dev::ld_x(x + c.w), s);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (j + e >= a && j + e < b) s = fmaf(values[j + e], dev::ld_x(x + col_indices[j + e]), s);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if (lane == 0) y[warp] = s;
}

__global__ void zero_rows_kernel(int rows, float* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) y[i] = 0.0f;
}

// Window size: as many rows as keep the expected non-zero count of a window
// inside one product pass, limited to what the row owners can hold in
// registers, a multiple of the group count.
template <int LPR>
int pick_window_rows(int rows, int nnz) {
    constexpr int groups = StreamGeom<LPR>::groups;
    constexpr int max_mult = StreamGeom<LPR>::max_rows_per_group;
    const double avg = rows > 0 ? static_cast<double>(nnz) / rows : 0.0;
    int mult = max_mult;
    while (mult > 1 && avg * groups * mult > kProductCap - 8) mult >>= 1;
    return groups * mult;
}

template <int LPR>
cudaError_t launch_stream_lpr(const CsrView& A, const float* x, float* y, cudaStream_t stream) {
    const int window = pick_window_rows<LPR>(A.rows, A.nnz);
    const size_t smem = stream_smem_bytes(window);
    cudaError_t e = cudaFuncSetAttribute(csr_stream_kernel<LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(stream_smem_bytes(StreamGeom<LPR>::max_window_rows)));
    if (e != cudaSuccess) return e;
    const unsigned blocks = static_cast<unsigned>((A.rows + window - 1) / window);
    csr_stream_kernel<LPR><<<blocks, kThreads, smem, stream>>>(A.rows, A.nnz, A.row_ptrs, A.col_indices,
                                                               A.values, x, y, window);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_csr_stream(const CsrView& A, const float* x, float* y, int lanes_per_row,
                              cudaStream_t stream) {
    if (A.rows <= 0) return cudaSuccess;
    if (A.nnz <= 0) {
        zero_rows_kernel<<<(A.rows + 255) / 256, 256, 0, stream>>>(A.rows, y);
        count_launches(1);
        return cudaGetLastError();
    }
    switch (lanes_per_row) {
        case 1:  return launch_stream_lpr<1>(A, x, y, stream);
        case 2:  return launch_stream_lpr<2>(A, x, y, stream);
        case 4:  return launch_stream_lpr<4>(A, x, y, stream);
        case 8:  return launch_stream_lpr<8>(A, x, y, stream);
        default: return launch_stream_lpr<16>(A, x, y, stream);
    }
}

cudaError_t launch_csr_warp_per_row(const CsrView& A, const float* x, float* y, cudaStream_t stream) {
    if (A.rows <= 0) return cudaSuccess;
    const long long threads = static_cast<long long>(A.rows) * 32;
    const unsigned blocks = static_cast<unsigned>((threads + kThreads - 1) / kThreads);
    csr_warp_row_kernel<<<blocks, kThreads, 0, stream>>>(A.rows, A.row_ptrs, A.col_indices, A.values, x, y);
    count_launches(1);
    return cudaGetLastError();
}

// lanes per row for VECTOR_CSR from the average row length:
// <8 -> 2, <16 -> 4, <32 -> 8, <64 -> 16, else a full warp
int vector_lanes_for(int rows, int nnz) {
    const double avg = rows > 0 ? static_cast<double>(nnz) / rows : 0.0;
    if (avg < 8.0) return 2;
    if (avg < 16.0) return 4;
    if (avg < 32.0) return 8;
    if (avg < 64.0) return 16;
    return 32;
}

cudaError_t launch_csr_vector(const CsrView& A, const float* x, float* y, cudaStream_t stream) {
    const int lanes = vector_lanes_for(A.rows, A.nnz);
    if (lanes == 32) return launch_csr_warp_per_row(A, x, y, stream);
    return launch_csr_stream(A, x, y, lanes, stream);
}

}  // namespace b200
}  // namespace spmv
