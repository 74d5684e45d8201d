// ell_kernels.cu -- column-major ELL SpMV for sm_100a.
//
// Replaces spmv_ell_kernel (reference src/spmv_kernels.cu:191-213): there one
// thread owns one row and issues 32-bit loads.  Here one thread owns RPT
// ADJACENT rows so every slice k is read with one 128-bit (RPT=4) or 64-bit
// (RPT=2) load of values and one of col_indices -- a warp covers 128 rows and
// reads 512 contiguous bytes per array per slice.  Matrix data streams through
// ld.global.nc.L1::no_allocate; x is gathered through the read-only path and
// stays L1/L2 resident; y leaves with one vector store.
//
// Per row the products are accumulated in slice order k = 0..W-1 with a
// separately rounded multiply and add (__fmul_rn/__fadd_rn), i.e. exactly the
// operation order of spmv_cpu_ell (reference src/spmv_cpu.cpp:18-32): the
// result is bit-identical to the reference's CPU path.
//
// Roofline: HBM.  Algorithmic bytes per launch = 8*rows*W + 4*cols + 4*rows
// (reference src/bandwidth.cpp:66-75).
#include "device_utils.cuh"
#include "internal.hpp"

namespace spmv {
namespace b200 {
namespace {

constexpr int kEllThreads = 256;
constexpr int kEllBatch = 4;  // slices in flight per thread

template <int RPT> struct VecLoad;
template <> struct VecLoad<4> {
    static __device__ __forceinline__ void load(const float* v, const int* c, float (&fv)[4], int (&ic)[4]) {
        const float4 a = dev::ld_stream_f4(v);
        const int4 b = dev::ld_stream_i4(c);
        fv[0] = a.x; fv[1] = a.y; fv[2] = a.z; fv[3] = a.w;
        ic[0] = b.x; ic[1] = b.y; ic[2] = b.z; ic[3] = b.w;
    }
    static __device__ __forceinline__ void store(float* y, const float (&s)[4]) {
        dev::st_stream_f4(y, make_float4(s[0], s[1], s[2], s[3]));
    }
};
template <> struct VecLoad<2> {
    static __device__ __forceinline__ void load(const float* v, const int* c, float (&fv)[2], int (&ic)[2]) {
        const float2 a = dev::ld_stream_f2(v);
        const int2 b = dev::ld_stream_i2(c);
        fv[0] = a.x; fv[1] = a.y;
        ic[0] = b.x; ic[1] = b.y;
    }
    static __device__ __forceinline__ void store(float* y, const float (&s)[2]) {
        dev::st_stream_f2(y, make_float2(s[0], s[1]));
    }
};
template <> struct VecLoad<1> {
    static __device__ __forceinline__ void load(const float* v, const int* c, float (&fv)[1], int (&ic)[1]) {
        fv[0] = dev::ld_stream_f(v);
        ic[0] = dev::ld_stream_i(c);
    }
    static __device__ __forceinline__ void store(float* y, const float (&s)[1]) { y[0] = s[0]; }
};

// rows % RPT == 0 and all pointers aligned to 4*RPT bytes (checked by the launcher)
template <int RPT>
__global__ void __launch_bounds__(kEllThreads)
ell_slice_kernel(int rows, int width, const int* __restrict__ col_indices,
                 const float* __restrict__ values, const float* __restrict__ x,
                 float* __restrict__ y, unsigned long long* __restrict__ nnz_counter) {
    const long long row0 = (static_cast<long long>(blockIdx.x) * kEllThreads + threadIdx.x) * RPT;
    unsigned live = 0;
    if (row0 < rows) {
        float acc[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) acc[i] = 0.0f;

        const size_t stride = static_cast<size_t>(rows);
        const float* vp = values + row0;
        const int* cp = col_indices + row0;

        for (int k0 = 0; k0 < width; k0 += kEllBatch) {
            float fv[kEllBatch][RPT];
            int ic[kEllBatch][RPT];
            float xv[kEllBatch][RPT];
            // issue every load of the batch before the first use
#pragma unroll
            for (int b = 0; b < kEllBatch; ++b) {
                if (k0 + b < width) VecLoad<RPT>::load(vp + (k0 + b) * stride, cp + (k0 + b) * stride, fv[b], ic[b]);
            }
#pragma unroll
            for (int b = 0; b < kEllBatch; ++b) {
                if (k0 + b < width) {
#pragma unroll
                    for (int i = 0; i < RPT; ++i) xv[b][i] = ic[b][i] >= 0 ? dev::ld_x(x + ic[b][i]) : 0.0f;
                }
            }
#pragma unroll
            for (int b = 0; b < kEllBatch; ++b) {
                if (k0 + b < width) {
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        if (ic[b][i] >= 0) {  // padding slots are skipped, not multiplied
                            acc[i] = __fadd_rn(acc[i], __fmul_rn(fv[b][i], xv[b][i]));
                            ++live;
                        }
                    }
                }
            }
        }
        VecLoad<RPT>::store(y + row0, acc);
    }
    if (nnz_counter != nullptr) {  // true (non-padding) entry count for the GFLOPS figure
        live = dev::warp_sum(live);
        if ((threadIdx.x & 31) == 0 && live) atomicAdd(nnz_counter, static_cast<unsigned long long>(live));
    }
}

template <int RPT>
cudaError_t launch_rpt(int rows, int width, const int* ci, const float* va, const float* x, float* y,
                       unsigned long long* counter, cudaStream_t stream) {
    const long long threads = (static_cast<long long>(rows) + RPT - 1) / RPT;
    const unsigned blocks = static_cast<unsigned>((threads + kEllThreads - 1) / kEllThreads);
    ell_slice_kernel<RPT><<<blocks, kEllThreads, 0, stream>>>(rows, width, ci, va, x, y, counter);
    count_launches(1);
    return cudaGetLastError();
}

bool all_aligned(const void* a, const void* b, const void* c, unsigned bytes) {
    const uintptr_t m = bytes - 1;
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
             reinterpret_cast<uintptr_t>(c)) & m) == 0;
}

__global__ void zero_rows_kernel(int rows, float* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) y[i] = 0.0f;
}

}  // namespace

cudaError_t launch_ell(int rows, int width, const int* col_indices, const float* values,
                       const float* x, float* y, unsigned long long* nnz_counter, cudaStream_t stream) {
    if (rows <= 0) return cudaSuccess;
    if (width <= 0) {  // no stored entries: y = 0
        zero_rows_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(rows, y);
        count_launches(1);
        return cudaGetLastError();
    }
    if (rows % 4 == 0 && all_aligned(col_indices, values, y, 16))
        return launch_rpt<4>(rows, width, col_indices, values, x, y, nnz_counter, stream);
    if (rows % 2 == 0 && all_aligned(col_indices, values, y, 8))
        return launch_rpt<2>(rows, width, col_indices, values, x, y, nnz_counter, stream);
    return launch_rpt<1>(rows, width, col_indices, values, x, y, nnz_counter, stream);
}

}  // namespace b200
}  // namespace spmv
