// ell_gated_kernel.cu -- the ELL product of the host-buffer call (host_pipeline.cu) as ONE persistent
// kernel that runs WHILE x is still arriving over PCIe and WHILE finished rows are already leaving.
//
// The reference's caller copies x up, calls spmv_ell (src/spmv_kernels.cu:191-213, 369-420) and copies
// y down (README.md:98-118, src/benchmark.cu:36-38,95-102).  The chunked form of host_pipeline.cu cuts
// that into row chunks with one upload, one launch and one download per chunk; measured on B200 every
// stream operation that depends on a copy costs 5-10 us of idle link, so ~12 chunks is its optimum and
// the download trails the upload by two of them (profiles/r2_host_gated.txt).
//
// Here the upload is ONE copy and the product ONE launch, and the data is its own arrival flag:
//   * before the upload the device copy of x holds a SENTINEL bit pattern (a signalling NaN nobody
//     computes with) in every entry; the copy engine overwrites it front to back;
//   * the PRODUCER warp of a CTA holds back the TMA refill of a 256-row window until the LAST x entry
//     that window reads no longer holds the sentinel (one L1-bypassing poll per window);
//   * the row owners (the consumers of ell_tma_pipe_kernel<1>, unchanged arithmetic: slots from shared
//     memory, slice order, separately rounded multiply / add -- bit-identical to spmv_cpu_ell) check
//     every gathered x for the sentinel and re-read it past the L1 until it has arrived, so the
//     result never depends on the ORDER in which the copy engine writes;
//   * a stream-ordered 32-bit write after the copy (cuStreamWriteValue32) publishes "all of x is
//     there": an x entry that REALLY holds the sentinel pattern is accepted then -- correct, only late;
//   * y leaves as it is finished.  y_host page-locked (the default case): the row owners store their rows
//     STRAIGHT into it -- 128-byte posted PCIe writes in step with the upload, no copy call at all (1.42 ms on
//     config 2; two bare DMA copies with no compute take 1.39 ms).  Any other y_host: every consumer warp adds 1
//     to the progress counter of its row chunk; the warp that completes a chunk writes the call's epoch into
//     ready[chunk] in page-locked HOST memory (one posted PCIe write, ~1 us), and the calling thread -- the call
//     is blocking, it has nothing else to do -- polls those words in order and queues the D2H copy of a chunk
//     the moment it is ready (1.64 ms: each copy costs ~19 us of idle link).  (Stream-ordered waits,
//     cuStreamWaitValue32, were measured first: the front end re-polls a failed wait so rarely that the download
//     trailed the product by ~50 us PER CHUNK.)
// A window therefore waits for the 16 KB of x behind it instead of 1/12 of x.
//
// Safety: a producer that sees neither its entry nor the completion flag within `timeout_ns` (a failed
// copy) sets *abort_word; everybody stops waiting, the kernel drains with meaningless rows and the
// host repeats the call in the chunked form.  The kernel can therefore not hang.
#include "device_utils.cuh"
#include "internal.hpp"

#include <vector>

namespace spmv {
namespace b200 {
namespace {

constexpr int kGatedRows = 256;               // rows per window = consumer threads
constexpr int kGatedThreads = kGatedRows + 32;  // + one producer warp
constexpr int kGatedMaxStages = 8;
constexpr unsigned kSentinel = 0x7FA3C0DEu;   // signalling NaN with an arbitrary payload

__device__ __forceinline__ unsigned ld_relaxed_sys(const void* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned atom_acq_rel_add(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// slow path of a gather: the entry still holds the sentinel (or the L1 holds a line fetched before it arrived)
__device__ __noinline__ float wait_for_x(const float* p, const unsigned* done_flag, unsigned epoch, const unsigned* abort_word,
                                         unsigned poll_sleep_ns) {
    for (;;) {
        unsigned v = ld_relaxed_sys(p);
        if (v != kSentinel) return __uint_as_float(v);
        if (ld_relaxed_sys(done_flag) == epoch) {  // the whole copy has landed: the entry really holds this pattern
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            return __uint_as_float(ld_relaxed_sys(p));
        }
        if (ld_relaxed_sys(abort_word) != 0) return __uint_as_float(v);
        __nanosleep(poll_sleep_ns ? poll_sleep_ns : 100);
    }
}

// window_poll[w] = largest column window w reads (-1: none).  progress[window_chunk[w]] counts the consumer
// warps of that row chunk that have stored their rows (zeroed by the host between calls; chunk_warps[c] = how
// many there are).
__global__ void __launch_bounds__(kGatedThreads)
ell_tma_gated_kernel(int rows, int width, int stages, const int* __restrict__ col_indices,
                     const float* __restrict__ values, const float* x, float* __restrict__ y,
                     const int* __restrict__ window_poll, const unsigned* done_flag, unsigned epoch,
                     unsigned* abort_word, unsigned* abort_host, unsigned long long timeout_ns, unsigned poll_sleep_ns,
                     unsigned* progress, const unsigned char* __restrict__ window_chunk, const unsigned* __restrict__ chunk_warps,
                     unsigned* ready_host) {
    extern __shared__ __align__(16) unsigned char gated_smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(gated_smem);   // [stages] producer -> consumers (TMA bytes)
    uint64_t* empty = full + kGatedMaxStages;                   // [stages] consumers -> producer
    unsigned char* buffers = gated_smem + 128;
    const size_t stage_bytes = static_cast<size_t>(width) * kGatedRows * 8;  // values then col_indices

    const int tid = threadIdx.x;
    const size_t stride = static_cast<size_t>(rows);
    const int num_windows = (rows + kGatedRows - 1) / kGatedRows;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            dev::mbar_init(full + s, 1);
            dev::mbar_init(empty + s, kGatedRows / 32);  // one arrival per consumer warp
        }
        dev::mbar_fence_init();
    }
    __syncthreads();

    if (tid >= kGatedRows) {
        // ---- producer warp: one lane gates and issues
        if (tid != kGatedRows) return;
        bool all_there = false;  // the completion flag was seen (or everybody gave up): no more polling
        int it = 0;
        for (int w = blockIdx.x; w < num_windows; w += gridDim.x, ++it) {
            const int stage = it % stages;
            if (it >= stages) dev::mbar_wait(empty + stage, ((it / stages) - 1) & 1u);
            const int last = window_poll[w];
            if (!all_there && last >= 0) {
                unsigned long long t0 = 0;
                unsigned polls = 0;
                while (ld_relaxed_sys(x + last) == kSentinel) {
                    if (ld_relaxed_sys(done_flag) == epoch) { all_there = true; break; }
                    if ((++polls & 15u) == 0) {
                        if (ld_relaxed_sys(abort_word) != 0) { all_there = true; break; }
                        const unsigned long long now = global_timer_ns();
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > timeout_ns) {  // device word: seen by everybody here; host word: by the caller
                            atomicExch(abort_word, 1u);
                            *reinterpret_cast<volatile unsigned*>(abort_host) = 1u;
                            all_there = true;
                            break;
                        }
                    }
                    if (poll_sleep_ns) __nanosleep(poll_sleep_ns);
                }
            }
            const long long r0 = static_cast<long long>(w) * kGatedRows;
            const int nr = static_cast<int>(min(static_cast<long long>(kGatedRows), rows - r0));
            const uint32_t slice_bytes = static_cast<uint32_t>(nr) * sizeof(float);
            float* s_val = reinterpret_cast<float*>(buffers + stage * stage_bytes);
            int* s_col = reinterpret_cast<int*>(s_val + width * kGatedRows);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            dev::mbar_arrive_expect_tx(full + stage, 2u * width * slice_bytes);
            for (int k = 0; k < width; ++k) {
                const size_t off = k * stride + static_cast<size_t>(r0);
                dev::tma_bulk_g2s(s_val + k * kGatedRows, values + off, slice_bytes, full + stage);
                dev::tma_bulk_g2s(s_col + k * kGatedRows, col_indices + off, slice_bytes, full + stage);
            }
        }
        return;
    }

    // ---- consumers: thread r owns row r of the window
    int it = 0;
    for (int w = blockIdx.x; w < num_windows; w += gridDim.x, ++it) {
        const int stage = it % stages;
        const long long r0 = static_cast<long long>(w) * kGatedRows;
        const int nr = static_cast<int>(min(static_cast<long long>(kGatedRows), rows - r0));
        const float* s_val = reinterpret_cast<const float*>(buffers + stage * stage_bytes);
        const int* s_col = reinterpret_cast<const int*>(s_val + width * kGatedRows);
        dev::mbar_wait(full + stage, (it / stages) & 1u);
        if (tid < nr) {
            float a = 0.0f;
#pragma unroll 5
            for (int k = 0; k < width; ++k) {
                const int c = s_col[k * kGatedRows + tid];
                if (c >= 0) {
                    float xv = dev::ld_x(x + c);
                    if (__float_as_uint(xv) == kSentinel) xv = wait_for_x(x + c, done_flag, epoch, abort_word, poll_sleep_ns);
                    a = __fadd_rn(a, __fmul_rn(s_val[k * kGatedRows + tid], xv));
                }
            }
            y[r0 + tid] = a;
        }
        __syncwarp();
        if ((tid & 31) == 0) {
            dev::mbar_arrive(empty + stage);  // this warp has read the stage
            if (progress) {  // ... and stored its rows; the warp that completes a row chunk tells the host
                const int chunk = window_chunk[w];
                const unsigned in_chunk = chunk_warps[chunk];
                if (atom_acq_rel_add(progress + chunk, 1u) + 1u == in_chunk) {
                    asm volatile("fence.acq_rel.sys;" ::: "memory");
                    st_relaxed_sys(ready_host + chunk, epoch);
                }
            }
        }
    }
}

// largest column (padding excluded) of every 256-row window
__global__ void ell_window_max_col_kernel(int rows, int width, const int* __restrict__ col_indices, int* __restrict__ wmax) {
    const int w = blockIdx.x;
    const int r = w * kGatedRows + threadIdx.x;
    int hi = -1;
    if (r < rows)
        for (int k = 0; k < width; ++k) hi = max(hi, col_indices[static_cast<size_t>(k) * rows + r]);
    hi = __reduce_max_sync(0xffffffffu, hi);
    __shared__ int s[kGatedRows / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = hi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kGatedRows / 32; ++i) hi = max(hi, s[i]);
        wmax[w] = hi;
    }
}

__global__ void fill_sentinel_kernel(unsigned* __restrict__ p, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        p[i] = kSentinel;
}

// diagnostic: block i polls entry i * step until it no longer holds the sentinel and records the time
// mode 0: ld.relaxed.sys, 1: ld.relaxed.gpu, 2: ld.volatile
__global__ void probe_arrival_kernel(const unsigned* x, size_t step, int samples, unsigned long long* t_ns, unsigned long long timeout_ns,
                                     int mode, unsigned sleep_ns) {
    const int i = blockIdx.x;
    if (i >= samples || threadIdx.x != 0) return;
    const unsigned* p = x + i * step;
    const unsigned long long t0 = global_timer_ns();
    unsigned long long now = t0;
    for (;;) {
        unsigned v;
        if (mode == 0) v = ld_relaxed_sys(p);
        else if (mode == 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        else asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        now = global_timer_ns();
        if (v != kSentinel || now - t0 > timeout_ns) break;
        if (sleep_ns) __nanosleep(sleep_ns);
    }
    t_ns[i] = now;
}

}  // namespace

// Diagnostic (profiles/r2_host_gated.txt): in which ORDER does one cudaMemcpyAsync of n floats land in device memory,
// and does polling the destination disturb it?  out_ns[i] = arrival time of entry i * (n / samples) relative to the
// earliest arrival; out_ns[samples] = duration of the copy by CUDA events.  samples == 0 in effect: mode < 0 skips the probe.
int probe_h2d_order(const float* x_host, size_t n, int samples, long long* out_ns, int mode, unsigned sleep_ns) {
    if (!x_host || !out_ns || samples < 1 || samples > 4096 || n < static_cast<size_t>(samples)) return -8;
    float* d_x = nullptr;
    unsigned long long* d_t = nullptr;
    cudaStream_t s_probe = nullptr, s_copy = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<unsigned long long> t(samples, 0);
    float ms = 0.0f;
    bool ok = cudaMalloc(&d_x, n * sizeof(float)) == cudaSuccess && cudaMalloc(&d_t, samples * sizeof(unsigned long long)) == cudaSuccess &&
              cudaStreamCreateWithFlags(&s_probe, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess &&
              cudaMemset(d_t, 0, samples * sizeof(unsigned long long)) == cudaSuccess;
    if (ok) {
        fill_sentinel_kernel<<<148 * 4, 512>>>(reinterpret_cast<unsigned*>(d_x), n);
        ok = cudaDeviceSynchronize() == cudaSuccess;
        if (mode >= 0)
            probe_arrival_kernel<<<samples, 32, 0, s_probe>>>(reinterpret_cast<const unsigned*>(d_x), n / samples, samples, d_t,
                                                               500000000ull, mode, sleep_ns);
        count_launches(2);
        ok = ok && cudaEventRecord(e0, s_copy) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(d_x, x_host, n * sizeof(float), cudaMemcpyHostToDevice, s_copy) == cudaSuccess;
        ok = ok && cudaEventRecord(e1, s_copy) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(s_copy) == cudaSuccess && cudaStreamSynchronize(s_probe) == cudaSuccess;
        ok = ok && cudaMemcpy(t.data(), d_t, samples * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess;
        ok = ok && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess;
    }
    if (ok) {
        unsigned long long lo = t[0];
        for (auto v : t) lo = v < lo ? v : lo;
        for (int i = 0; i < samples; ++i) out_ns[i] = static_cast<long long>(t[i] - lo);
        out_ns[samples] = static_cast<long long>(ms * 1e6);
    }
    cudaFree(d_x); cudaFree(d_t);
    if (s_probe) cudaStreamDestroy(s_probe);
    if (s_copy) cudaStreamDestroy(s_copy);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (!ok) { cudaGetLastError(); return -4; }
    return 0;
}

bool ell_gated_applies(int rows, int width, const int* col_indices, const float* values) {
    return rows > 0 && width >= 1 && width <= 8 && rows % 4 == 0 &&
           ((reinterpret_cast<uintptr_t>(col_indices) | reinterpret_cast<uintptr_t>(values)) & 15u) == 0;
}

int ell_gated_windows(int rows) { return (rows + kGatedRows - 1) / kGatedRows; }
int ell_gated_window_rows() { return kGatedRows; }
int ell_gated_warps_per_window() { return kGatedRows / 32; }

cudaError_t launch_ell_window_max_col(int rows, int width, const int* col_indices, int* d_wmax, cudaStream_t stream) {
    const int windows = ell_gated_windows(rows);
    if (windows <= 0) return cudaSuccess;
    ell_window_max_col_kernel<<<windows, kGatedRows, 0, stream>>>(rows, width, col_indices, d_wmax);
    count_launches(1);
    return cudaGetLastError();
}

// x[0 .. n) <- sentinel
cudaError_t launch_fill_sentinel(float* x, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    fill_sentinel_kernel<<<148 * 4, 512, 0, stream>>>(reinterpret_cast<unsigned*>(x), n);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_ell_gated(int rows, int width, const int* col_indices, const float* values, const float* x, float* y,
                             const int* d_window_poll, const unsigned* d_done_flag, unsigned epoch, unsigned* abort_word,
                             unsigned* abort_host, unsigned long long timeout_ns, unsigned poll_sleep_ns, unsigned* d_progress,
                             const unsigned char* d_window_chunk, const unsigned* d_chunk_warps, unsigned* ready_host,
                             cudaStream_t stream) {
    if (!ell_gated_applies(rows, width, col_indices, values)) return cudaErrorInvalidConfiguration;
    const size_t stage_bytes = static_cast<size_t>(width) * kGatedRows * 8;
    const int stages = 2;
    const size_t smem = 128 + stages * stage_bytes;
    cudaError_t e = cudaFuncSetAttribute(ell_tma_gated_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    int fit = 1;  // the grid must be co-resident: a waiting CTA never keeps one that is not yet scheduled from its turn
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, ell_tma_gated_kernel, kGatedThreads, smem);
    if (e != cudaSuccess) return e;
    if (fit < 1) return cudaErrorInvalidConfiguration;
    int sms = 148, dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    int blocks = sms * fit;
    const int windows = ell_gated_windows(rows);
    if (blocks > windows) blocks = windows;
    ell_tma_gated_kernel<<<blocks, kGatedThreads, smem, stream>>>(rows, width, stages, col_indices, values, x, y, d_window_poll,
                                                                   d_done_flag, epoch, abort_word, abort_host, timeout_ns,
                                                                   poll_sleep_ns, d_progress, d_window_chunk, d_chunk_warps, ready_host);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace b200
}  // namespace spmv
