// gather_bench.cu -- design microbenchmark (not product): how fast can one B200 SM gather 4-byte
// words (a) from global memory through L1/L2, (b) from a shared-memory table, (c) from a table
// spread over the shared memory of a thread-block cluster (DSMEM)?  The answers size the hub-column
// table of the power-law SpMV kernels (csr_hot_kernels.cu).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x = (x ^ (x >> 16)) * 0x7FEB352Du;
    x = (x ^ (x >> 15)) * 0x846CA68Bu;
    return x ^ (x >> 16);
}

// idx[i] uniform in [0, footprint); window > 0: the 32 indices of a warp fall into `window` consecutive
// 128-byte lines chosen at random (models locality inside a warp-wide gather)
__global__ void make_idx(uint32_t* idx, size_t n, uint32_t footprint, int window) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = mix32((uint32_t)i * 2654435761u + 12345u);
        if (window > 0) {
            uint32_t warp = (uint32_t)(i >> 5);
            uint32_t lines = footprint / 32;
            uint32_t base = mix32(warp * 0x9E3779B1u + 7u) % (lines - window);
            idx[i] = (base + h % window) * 32 + (mix32(h) & 31);
        } else {
            idx[i] = h % footprint;
        }
    }
}
__global__ void fill(float* x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = (float)(i & 1023) * 0.001f;
}

constexpr int U = 8;

// (a) global gathers, persistent grid
__global__ void __launch_bounds__(1024, 1) k_global(const uint32_t* __restrict__ idx, size_t n, const float* __restrict__ x, float* out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n; i0 += stride) {
        uint32_t c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = (i0 + u * blockDim.x < n) ? __ldcs(idx + i0 + u * blockDim.x) : 0;
#pragma unroll
        for (int u = 0; u < U; ++u) acc += __ldg(x + c[u]);
    }
    if (acc == 123.456f) out[0] = acc;
}

// (b) shared-memory table (entries floats), filled from x[0..entries)
__global__ void __launch_bounds__(1024, 1) k_smem(const uint32_t* __restrict__ idx, size_t n, const float* __restrict__ x, int entries, float* out) {
    extern __shared__ float tab[];
    for (int i = threadIdx.x; i < entries; i += blockDim.x) tab[i] = x[i];
    __syncthreads();
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n; i0 += stride) {
        uint32_t c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = (i0 + u * blockDim.x < n) ? __ldcs(idx + i0 + u * blockDim.x) : 0;
#pragma unroll
        for (int u = 0; u < U; ++u) acc += tab[c[u]];
    }
    if (acc == 123.456f) out[0] = acc;
}

// (c) cluster-wide table: entry e lives in CTA (e % csize) at slot e / csize
__global__ void __launch_bounds__(1024, 1) k_dsmem(const uint32_t* __restrict__ idx, size_t n, const float* __restrict__ x, int entries_per_cta, float* out) {
    extern __shared__ float tab[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned csize = cluster.num_blocks();
    const unsigned crank = cluster.block_rank();
    for (int i = threadIdx.x; i < entries_per_cta; i += blockDim.x) tab[i] = x[(size_t)i * csize + crank];
    cluster.sync();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n; i0 += stride) {
        uint32_t c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = (i0 + u * blockDim.x < n) ? __ldcs(idx + i0 + u * blockDim.x) : 0;
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = c[u] % csize, slot = c[u] / csize;
            uint32_t ra;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base + slot * 4), "r"(r));
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v[u]) : "r"(ra));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 123.456f) out[0] = acc;
    cluster.sync();
}

// (d) mixed: a fraction of the lanes reads the local table, the rest global (predicated, as gather_enc)
__global__ void __launch_bounds__(1024, 1) k_mixed(const uint32_t* __restrict__ idx, size_t n, const float* __restrict__ x, int entries, uint32_t hot_per_256, float* out) {
    extern __shared__ float tab[];
    for (int i = threadIdx.x; i < entries; i += blockDim.x) tab[i] = x[i];
    __syncthreads();
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n; i0 += stride) {
        uint32_t c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = (i0 + u * blockDim.x < n) ? __ldcs(idx + i0 + u * blockDim.x) : 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool hot = (mix32(c[u]) & 255u) < hot_per_256;
            acc += hot ? tab[c[u] % entries] : __ldg(x + c[u]);
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <class F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / reps;
}

int main() {
    const size_t n = 1ull << 27;  // 128 M gathers
    uint32_t* idx; float* x; float* out;
    const size_t xn = 1ull << 26;  // 256 MB of x
    CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&x, xn * 4)); CK(cudaMalloc(&out, 4));
    fill<<<1184, 256>>>(x, xn);
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, clock %d kHz, n = %zu gathers\n", sms, clk, n);
    auto report = [&](const char* name, float ms, int active_sms) {
        double gps = n / (ms * 1e-3) / 1e9;
        printf("%-58s %8.4f ms  %7.1f G/s  %6.3f elem/clk/SM (at %d SMs, 1.965 GHz)\n", name, ms, gps, gps / active_sms / 1.965, active_sms);
        fflush(stdout);
    };
    // ---- (a) global
    for (uint32_t fp : {1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 18, 1u << 21, 1u << 24, 1u << 26}) {
        for (int window : {0, 16, 8, 4, 2}) {
            if (window && fp != (1u << 24)) continue;
            make_idx<<<1184, 256>>>(idx, n, fp, window);
            char name[128]; snprintf(name, sizeof name, "global  footprint %4u KB window %2d", fp / 256, window);
            report(name, time_ms([&] { k_global<<<sms, 1024>>>(idx, n, x, out); }), sms);
            if (!window) {
                snprintf(name, sizeof name, "global  footprint %4u KB 2 CTAs/SM", fp / 256);
                report(name, time_ms([&] { k_global<<<2 * sms, 1024>>>(idx, n, x, out); }), sms);
            }
        }
    }
    // ---- (b) local smem table
    for (int entries : {8192, 24576, 49152}) {
        make_idx<<<1184, 256>>>(idx, n, entries, 0);
        CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, entries * 4));
        char name[128]; snprintf(name, sizeof name, "smem    table %6d entries", entries);
        report(name, time_ms([&] { k_smem<<<sms, 1024, entries * 4>>>(idx, n, x, entries, out); }), sms);
    }
    // ---- (d) mixed hot/cold at a 96 KB table, x footprint 64 MB
    make_idx<<<1184, 256>>>(idx, n, 1u << 24, 0);
    for (int entries : {24576, 40960}) {
        CK(cudaFuncSetAttribute(k_mixed, cudaFuncAttributeMaxDynamicSharedMemorySize, entries * 4));
        for (uint32_t hot : {0u, 85u, 128u, 179u, 218u, 256u}) {
            char name[128]; snprintf(name, sizeof name, "mixed   table %6d entries, hot share %3u/256", entries, hot);
            report(name, time_ms([&] { k_mixed<<<sms, 1024, entries * 4>>>(idx, n, x, entries, hot, out); }), sms);
        }
    }
    // ---- (c) DSMEM
    for (int csize : {1, 2, 4, 8, 16}) {
        for (int per_cta : {24576, 40960}) {
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = per_cta * 4; cfg.attrs = attr; cfg.numAttrs = 1;
            CK(cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, per_cta * 4));
            if (csize > 8 && cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
            cfg.gridDim = dim3(csize);
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, k_dsmem, &cfg) != cudaSuccess || max_clusters < 1) { cudaGetLastError(); printf("dsmem csize %d: no occupancy\n", csize); continue; }
            cfg.gridDim = dim3(max_clusters * csize);
            const uint32_t entries = (uint32_t)per_cta * csize;
            make_idx<<<1184, 256>>>(idx, n, entries, 0);
            char name[128]; snprintf(name, sizeof name, "dsmem   cluster %2d x %6d entries (%d clusters)", csize, per_cta, max_clusters);
            int ep = per_cta;
            report(name, time_ms([&] { CK(cudaLaunchKernelEx(&cfg, k_dsmem, (const uint32_t*)idx, n, (const float*)x, ep, out)); }), max_clusters * csize);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
