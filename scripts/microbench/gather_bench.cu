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

// (a2) the same gathers while 8 more bytes per gather are STREAMED, as the matrix stream of an SpMV is:
// MODE 1 = by the gathering threads themselves (two more coalesced 4-byte loads per gather: the stream
// shares the LSU / L1 miss path with the gathers), MODE 2 = by 1-D TMA bulk copies into a shared-memory ring
// issued by one thread (async proxy: the stream does not pass through the LSU).  Question: is the
// ~1 sector per clock and SM of random L2 gathers a limit of the SM's LSU miss path (then TMA staging of the
// stream frees it) or of the L2 (then it does not)?
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_global_stream(const uint32_t* __restrict__ idx, size_t n, const float* __restrict__ x,
                                                           const float* __restrict__ s1, const float* __restrict__ s2, float* out) {
    extern __shared__ __align__(128) unsigned char ring[];
    constexpr int kStage = 2 * 1024 * U * 4;  // bytes of s1 + s2 one CTA iteration covers (8 B per gather)
    __shared__ __align__(8) unsigned long long bars[2];
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    if (MODE == 2 && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    int it = 0;
    auto issue = [&](size_t i0_cta, int stage) {  // thread 0: both streams of one CTA iteration -> ring[stage]
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + stage * kStage);
        const uint32_t bar = bar0 + 8 * stage;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(kStage) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(s1 + i0_cta), "r"(kStage / 2), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst + kStage / 2), "l"(s2 + i0_cta), "r"(kStage / 2), "r"(bar) : "memory");
    };
    const size_t first = (size_t)blockIdx.x * blockDim.x * U;
    if (MODE == 2 && threadIdx.x == 0) {
        if (first + 1024 * U <= n) issue(first, 0);
        if (first + stride + 1024 * U <= n) issue(first + stride, 1);
    }
    for (size_t base = first; base + 1024 * U <= n; base += stride, ++it) {
        const size_t i0 = base + threadIdx.x;
        uint32_t c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = __ldcs(idx + i0 + u * blockDim.x);
        float v[U];
        if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldcs(s1 + i0 + u * blockDim.x) + __ldcs(s2 + i0 + u * blockDim.x);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += __ldg(x + c[u]);
        if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u];
        }
        if (MODE == 2) {
            const int stage = it & 1;
            const uint32_t bar = bar0 + 8 * stage, parity = (it >> 1) & 1;
            asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                         :: "r"(bar), "r"(parity) : "memory");
            const float* sv = reinterpret_cast<const float*>(ring + stage * kStage);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += sv[threadIdx.x + u * 1024] + sv[1024 * U + threadIdx.x + u * 1024];
            __syncthreads();
            const size_t nxt = base + 2 * stride;
            if (threadIdx.x == 0 && nxt + 1024 * U <= n) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(nxt, stage);
            }
        }
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
    // ---- (a2) gathers + an 8 B/gather stream: by LDG of the gathering threads vs by TMA bulk copies
    {
        float *s1, *s2;
        CK(cudaMalloc(&s1, n * 4)); CK(cudaMalloc(&s2, n * 4));
        fill<<<1184, 256>>>(s1, n); fill<<<1184, 256>>>(s2, n);
        make_idx<<<1184, 256>>>(idx, n, 1u << 24, 0);
        const int ring_bytes = 2 * 2 * 1024 * U * 4;
        CK(cudaFuncSetAttribute(k_global_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes));
        report("global  footprint 65536 KB, gathers only (again)", time_ms([&] { k_global<<<sms, 1024>>>(idx, n, x, out); }), sms);
        report("global  footprint 65536 KB + 8 B/gather stream by LDG", time_ms([&] { k_global_stream<1><<<sms, 1024>>>(idx, n, x, s1, s2, out); }), sms);
        report("global  footprint 65536 KB + 8 B/gather stream by TMA", time_ms([&] { k_global_stream<2><<<sms, 1024, ring_bytes>>>(idx, n, x, s1, s2, out); }), sms);
        CK(cudaFree(s1)); CK(cudaFree(s2));
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
