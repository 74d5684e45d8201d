// pagerank.cu -- PageRank on the fused merge-path iteration.
//
// Replaces pagerank() (reference src/pagerank.cu:50-153).  The reference runs
// only the SpMV on the GPU and, EVERY iteration, copies the rank vector to the
// host, applies damping/teleport/dangling mass there with single-threaded fp32
// loops, copies it back and computes the L2 residual on the host.  Here the
// whole iteration is one pass of merge_tile_kernel<PageRankRow> (+ its fix-up
// and a 1-CTA reduction): the rank vectors never leave HBM and the only
// per-iteration host traffic is the 24-byte {sum d^2, sum |d|, dangling mass}.
//
// Recurrence, initial vector, update expression order, stop rule (L2 norm of
// the delta < tolerance, checked after every iteration), result fields and the
// final normalisation follow the reference line by line; the three global
// sums use f64 accumulators because the reference's sequential fp32 host sums
// lose all accuracy at n = 2^26.
//
// The same building blocks work on a row shard of the matrix (row_offset,
// n_global) so that one process per GPU can run a sharded PageRank with an
// all-gather of the rank slices and an all-reduce of the three sums in between
// (gpu-spmv_b200/dist.py).
#include "internal.hpp"

#include <cmath>
#include <cstdlib>
#include <new>

namespace spmv {
namespace b200 {

constexpr int kTmpDoubles = 148 * 16 + 8;

// A plan over one row shard: merge-path coordinates are computed once (the
// matrix is constant across iterations) and the work arrays are owned here.
struct PrPlan {
    CsrView A{};
    int row_offset = 0;
    int n_global = 0;
    Scratch merge_block;
    MergePlan merge;
    Scratch tmp_block;   // kTmpDoubles doubles
    double* tmp = nullptr;
    // A ROW SHARD of a scale-free graph takes the hub-column plan: the tile epilogue of that kernel
    // overlaps the slice exchange with the product (8 GPUs, R-MAT 26: 1.17 vs 1.46 ms per iteration
    // against the segmented-stream kernel, whose exchange is a separate pass).  The WHOLE graph on
    // one GPU has nothing to exchange and takes the segmented stream with its streaming epilogue
    // (R-MAT 26: 5.45 vs 5.89 ms per iteration; R-MAT 24: 1.21 vs 1.29).  Neither: plain tile kernel.
    PlannedCsr planned;
    bool whole_graph = false;
    ~PrPlan() { planned.release(); }
};

static int pr_hot_env() {
    static const int mode = [] {
        const char* v = getenv("SPMV_B200_HOT");
        return v ? atoi(v) : -1;
    }();
    return mode;
}

int pr_plan_create(const CSRMatrix* shard, int row_offset, int n_global, cudaStream_t stream, PrPlan** out) {
    if (!shard || !out || row_offset < 0 || n_global < 0) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    if (static_cast<long long>(row_offset) + shard->num_rows > n_global)
        return static_cast<int>(SpMVError::INVALID_DIMENSION);
    // every kernel gathers r[col] for col < num_cols from a rank vector of n_global entries: a shard (or a whole
    // adjacency matrix) with another column count would read past it.  The reference rejects the same input in
    // spmv_csr(..., vec_size = n) -> INVALID_DIMENSION (src/spmv_kernels.cu:224-226, src/pagerank.cu:102-107).
    if (shard->num_cols != n_global) return static_cast<int>(SpMVError::INVALID_DIMENSION);
    if (!shard->d_row_ptrs || (shard->nnz > 0 && (!shard->d_col_indices || !shard->d_values)))
        return static_cast<int>(SpMVError::INVALID_FORMAT);
    PrPlan* p = new (std::nothrow) PrPlan();
    if (!p) return static_cast<int>(SpMVError::OUT_OF_MEMORY);
    p->A = view_of(shard);
    p->row_offset = row_offset;
    p->n_global = n_global;
    void* block = p->merge_block.reserve(merge_plan_bytes(p->A.rows, p->A.nnz, true));
    p->tmp = static_cast<double*>(p->tmp_block.reserve(kTmpDoubles * sizeof(double)));
    if (!block || !p->tmp) {
        delete p;
        return static_cast<int>(SpMVError::CUDA_MALLOC);
    }
    p->merge = merge_plan_carve(block, p->A.rows, p->A.nnz, true);
    if (p->A.rows > 0 && launch_merge_partition(p->A, p->merge, stream) != cudaSuccess) {
        cudaGetLastError();
        delete p;
        return static_cast<int>(SpMVError::KERNEL_LAUNCH);
    }
    // the matrix is constant over the iterations: the hub-column plan pays for itself after a few
    p->whole_graph = row_offset == 0 && shard->num_rows == n_global;
    if (pr_hot_env() != 0 &&
        planned_build(p->A, &p->planned, 0, pr_hot_env() > 0, p->whole_graph, stream, p->whole_graph) != cudaSuccess) {
        cudaGetLastError();
        p->planned.release();  // not fatal: the plain tile kernel is used
    }
    *out = p;
    return 0;
}

void pr_plan_destroy(PrPlan* p) { delete p; }

// max_hot_columns: 0 drops the plan (plain tile kernel), < 0 device maximum; force skips the
// size / benefit thresholds.  Returns the number of hub columns now in use, or a negative error.
int pr_plan_set_hot(PrPlan* p, int max_hot_columns, bool force, cudaStream_t stream) {
    if (!p) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    cudaStreamSynchronize(stream);
    p->planned.release();
    if (max_hot_columns == 0) return 0;
    if (planned_build(p->A, &p->planned, max_hot_columns < 0 ? 0 : max_hot_columns, force, p->whole_graph, stream,
                      p->whole_graph) != cudaSuccess) {
        cudaGetLastError();
        p->planned.release();
        return static_cast<int>(SpMVError::CUDA_MALLOC);
    }
    return p->planned.seg.valid() ? (p->planned.seg.n_hot > 0 ? p->planned.seg.n_hot : 1) : p->planned.hot.n_hot;
}

int pr_step(PrPlan* p, const float* d_r_old, float* d_r_new, float damping, const float* d_dsum,
            const uint32_t* d_bits, double* d_partial, cudaStream_t stream, float* const* peer_r_new, int n_peers,
            int self_rank, float* mc_r_new) {
    if (!p || !d_r_old || !d_r_new || !d_dsum || !d_bits || !d_partial)
        return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    if (n_peers > kMaxPeers || (n_peers > 1 && ((!peer_r_new && !mc_r_new) || self_rank < 0 || self_rank >= n_peers)))
        return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    PageRankStepArgs a;
    a.n_peers = n_peers > 1 ? n_peers : 0;
    a.self_rank = self_rank;
    for (int i = 0; i < kMaxPeers; ++i) a.peers[i] = (n_peers > 1 && i < n_peers && peer_r_new) ? peer_r_new[i] : nullptr;
    a.mc_r_new = n_peers > 1 ? mc_r_new : nullptr;
    a.r_old = d_r_old;
    a.r_new = d_r_new;
    a.row_offset = p->row_offset;
    a.n_global = p->n_global;
    a.damping = damping;
    a.teleport = (1.0f - damping) / p->n_global;  // reference src/pagerank.cu:86
    a.d_dsum = d_dsum;
    a.bits = d_bits;
    a.out = d_partial;
    const cudaError_t e = p->planned.seg.valid()     ? launch_seg_pagerank(p->A, p->planned.seg, a, stream)
                          : p->planned.hot.n_hot > 0 ? launch_hot_pagerank(p->A, p->planned.hot, p->merge, a, stream)
                                                     : launch_merge_pagerank(p->A, p->merge, a, stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return static_cast<int>(SpMVError::KERNEL_LAUNCH);
    }
    return 0;
}

const CsrView& pr_plan_view(const PrPlan* p) { return p->A; }
int pr_plan_hub_columns(const PrPlan* p) {
    if (!p) return 0;
    return p->planned.seg.valid() ? p->planned.seg.n_hot : p->planned.hot.n_hot;
}
double* pr_plan_tmp(PrPlan* p) { return p->tmp; }

// pagerank_small.cu
bool pagerank_small_applies(int n, int nnz);
cudaError_t pagerank_small_run(const CsrView& A, float damping, float tolerance, int max_iterations, const uint32_t* d_bits,
                               float* buf_a, float* buf_b, const float* d_dsum, float* l2_history, int history_capacity,
                               int* iterations, float* residual, double* l1, bool* converged, int* final_buffer,
                               cudaStream_t stream);

// The loop of a small graph in one persistent kernel (pagerank_small.cu); same set-up, same outputs as the
// multi-kernel loop below.  kNotHandled: take that loop instead.
constexpr int kNotHandled = 1 << 20;
static int pagerank_device_small(const CSRMatrix* adj, const PageRankConfig* config, float* d_ranks, int* iterations,
                                 float* final_residual, bool* converged, double* l1_residual, bool normalize, float* l2_history,
                                 int history_capacity) {
    const int n = adj->num_rows;
    if (!adj->d_row_ptrs || (adj->nnz > 0 && (!adj->d_col_indices || !adj->d_values))) return kNotHandled;  // the loop below reports it
    const CsrView A = view_of(adj);
    cudaStream_t stream = nullptr;
    const size_t words = (static_cast<size_t>(n) + 31) / 32;
    float *d_a = nullptr, *d_b = nullptr, *d_dsum = nullptr;
    double *d_colsum = nullptr, *d_tmp = nullptr;
    uint32_t* d_bits = nullptr;
    const bool ok = cudaMalloc(&d_a, sizeof(float) * n) == cudaSuccess && cudaMalloc(&d_b, sizeof(float) * n) == cudaSuccess &&
                    cudaMalloc(&d_colsum, sizeof(double) * n) == cudaSuccess && cudaMalloc(&d_bits, sizeof(uint32_t) * words) == cudaSuccess &&
                    cudaMalloc(&d_dsum, sizeof(float)) == cudaSuccess && cudaMalloc(&d_tmp, sizeof(double) * kTmpDoubles) == cudaSuccess;
    int rc = 0;
    if (!ok) {
        cudaGetLastError();
        rc = static_cast<int>(SpMVError::CUDA_MALLOC);
    } else {
        // dangling nodes, r = 1/n and its dangling mass: exactly as below
        cudaMemsetAsync(d_colsum, 0, sizeof(double) * n, stream);
        launch_colsum(A, d_colsum, stream);
        launch_dangling_bits(d_colsum, n, n, d_bits, stream);
        launch_pr_init(n, d_bits, d_a, d_dsum, d_tmp, stream);
        int fin = 0;
        const cudaError_t e = pagerank_small_run(A, config->damping_factor, config->tolerance, config->max_iterations, d_bits, d_a, d_b,
                                                 d_dsum, l2_history, history_capacity, iterations, final_residual, l1_residual,
                                                 converged, &fin, stream);
        if (e == cudaErrorNotSupported) {
            rc = kNotHandled;
        } else if (e != cudaSuccess) {
            cudaGetLastError();
            rc = static_cast<int>(SpMVError::KERNEL_LAUNCH);
        } else {
            const float* vec = fin ? d_b : d_a;
            if (normalize) launch_normalize(vec, n, d_ranks, d_tmp, stream);
            else cudaMemcpyAsync(d_ranks, vec, sizeof(float) * n, cudaMemcpyDeviceToDevice, stream);
            if (cudaStreamSynchronize(stream) != cudaSuccess) {
                cudaGetLastError();
                rc = static_cast<int>(SpMVError::KERNEL_LAUNCH);
            }
        }
    }
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_colsum); cudaFree(d_bits); cudaFree(d_dsum); cudaFree(d_tmp);
    return rc;
}

// The whole loop on one device; d_ranks receives the ranks, normalised on the device with an
// f64 sum when `normalize` is set, else the raw final iterate.
int pagerank_device(const CSRMatrix* adj, const PageRankConfig* config, float* d_ranks, int* iterations,
                    float* final_residual, bool* converged, double* l1_residual, bool normalize, float* l2_history,
                    int history_capacity) {
    NvtxRange nvtx_range("spmv_b200:pagerank_device");
    if (!adj || !d_ranks) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    PageRankConfig defaults;
    if (!config) config = &defaults;
    const int n = adj->num_rows;
    if (iterations) *iterations = 0;
    if (final_residual) *final_residual = 0.0f;
    if (converged) *converged = false;
    if (l1_residual) *l1_residual = 0.0;
    if (n <= 0) return 0;
    // non-square adjacency: the reference's first spmv_csr(adj, ..., vec_size = n) fails with INVALID_DIMENSION and
    // pagerank() returns the uniform vector with iterations = 0 (src/pagerank.cu:102-107); pagerank() below does the same
    if (adj->num_cols != n) return static_cast<int>(SpMVError::INVALID_DIMENSION);

    if (pagerank_small_applies(n, adj->nnz)) {  // launch-bound sizes: the whole loop in one persistent kernel
        const int r = pagerank_device_small(adj, config, d_ranks, iterations, final_residual, converged, l1_residual, normalize,
                                            l2_history, history_capacity);
        if (r != kNotHandled) return r;
    }

    cudaStream_t stream = nullptr;
    PrPlan* plan = nullptr;
    int rc = pr_plan_create(adj, 0, n, stream, &plan);
    if (rc != 0) return rc;

    const int cols = adj->num_cols;
    const size_t colsum_n = static_cast<size_t>(cols > n ? cols : n);
    const size_t words = (static_cast<size_t>(n) + 31) / 32;
    float *d_a = nullptr, *d_b = nullptr, *d_dsum = nullptr;
    double* d_colsum = nullptr;
    uint32_t* d_bits = nullptr;
    double* d_partial = nullptr;
    double* h_partial = nullptr;
    bool ok = cudaMalloc(&d_a, sizeof(float) * n) == cudaSuccess &&
              cudaMalloc(&d_b, sizeof(float) * n) == cudaSuccess &&
              cudaMalloc(&d_colsum, sizeof(double) * colsum_n) == cudaSuccess &&
              cudaMalloc(&d_bits, sizeof(uint32_t) * words) == cudaSuccess &&
              cudaMalloc(&d_dsum, sizeof(float)) == cudaSuccess &&
              cudaMalloc(&d_partial, 3 * sizeof(double)) == cudaSuccess &&
              cudaMallocHost(&h_partial, 6 * sizeof(double)) == cudaSuccess;
    auto cleanup = [&]() {
        cudaFree(d_a); cudaFree(d_b); cudaFree(d_colsum); cudaFree(d_bits); cudaFree(d_dsum); cudaFree(d_partial);
        if (h_partial) cudaFreeHost(h_partial);
        pr_plan_destroy(plan);
    };
    if (!ok) {
        cudaGetLastError();
        cleanup();
        return static_cast<int>(SpMVError::CUDA_MALLOC);
    }

    // dangling nodes: columns whose stored values sum to 0 (reference :20-48, :87)
    cudaMemsetAsync(d_colsum, 0, sizeof(double) * colsum_n, stream);
    launch_colsum(plan->A, d_colsum, stream);
    launch_dangling_bits(d_colsum, n, cols, d_bits, stream);
    // r = 1/n and its dangling mass (reference :69-72, :94-99 for iteration 0)
    launch_pr_init(n, d_bits, d_a, d_dsum, plan->tmp, stream);

    // The stop rule is the reference's (L2 norm of the delta < tolerance, every iteration,
    // :118-127) but it is read one iteration late: the 24-byte sums of iteration i go to pinned
    // host memory asynchronously and are inspected while iteration i+1 is already queued, so the
    // device never idles on the host.  If iteration i had converged, its output is the INPUT of
    // the speculative iteration i+1 (which writes the other buffer), so vector, iteration count
    // and residual are exactly those of the eager rule.
    float* r_old = d_a;
    float* r_new = d_b;
    const float* fin = r_old;  // what the reference returns when the loop never runs (:135-139)
    int iters = 0;
    float residual = 0.0f;
    double l1 = 0.0;
    bool conv = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
    struct Pending { int iter; int slot; const float* vec; bool valid; } pending = {0, 0, nullptr, false};
    auto settle = [&](const Pending& p) -> bool {  // true when that iteration met the tolerance
        if (cudaEventSynchronize(ev[p.slot]) != cudaSuccess) {
            cudaGetLastError();
            rc = static_cast<int>(SpMVError::KERNEL_LAUNCH);
            return true;
        }
        residual = std::sqrt(static_cast<float>(h_partial[3 * p.slot + 0]));  // L2 norm of the delta (:118)
        l1 = h_partial[3 * p.slot + 1];
        iters = p.iter;
        if (l2_history && p.iter >= 1 && p.iter <= history_capacity) l2_history[p.iter - 1] = residual;
        fin = p.vec;
        return residual < config->tolerance;  // :123-127
    };
    // One iteration = the fused step (3-4 kernels), the dangling-mass hand-over and the 24-byte download.
    auto enqueue_iteration = [&](float* from, float* to, int slot, cudaStream_t s) -> int {
        const int r = pr_step(plan, from, to, config->damping_factor, d_dsum, d_bits, d_partial, s);
        if (r != 0) return r;
        launch_next_dsum(d_partial, d_dsum, s);
        cudaMemcpyAsync(h_partial + 3 * slot, d_partial, 3 * sizeof(double), cudaMemcpyDeviceToHost, s);
        return 0;
    };
    // CUDA-graph replay: the two ping-pong iterations (a -> b with host slot 0,
    // b -> a with slot 1) are captured once and replayed, one graph launch per iteration instead of
    // 5-6 API calls -- small graphs are launch-bound.  The stop rule stays per iteration (above).
    // Capture needs a real stream; a blocking one keeps the legacy default stream's ordering.  Any
    // failure falls back to the eager loop.  SPMV_B200_PR_GRAPH=0 disables it.
    cudaStream_t gs = nullptr;
    cudaGraphExec_t exec[2] = {nullptr, nullptr};
    unsigned long long launches_per_iteration = 0;
    static const int use_graph = [] {
        const char* v = getenv("SPMV_B200_PR_GRAPH");
        return v ? atoi(v) : 1;
    }();
    if (use_graph && config->max_iterations >= 4 && cudaStreamCreate(&gs) == cudaSuccess) {
        for (int p = 0; p < 2; ++p) {
            cudaGraph_t g = nullptr;
            const unsigned long long before = launch_count();
            bool good = cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (good) {
                const int r = enqueue_iteration(p == 0 ? d_a : d_b, p == 0 ? d_b : d_a, p, gs);
                good = cudaStreamEndCapture(gs, &g) == cudaSuccess && r == 0 && g != nullptr;
            }
            launches_per_iteration = launch_count() - before;
            count_launches(-static_cast<int>(launches_per_iteration));  // captured, not launched
            if (good) good = cudaGraphInstantiate(&exec[p], g, 0) == cudaSuccess;
            if (g) cudaGraphDestroy(g);
            if (!good) {
                cudaGetLastError();
                for (int q = 0; q < 2; ++q) {
                    if (exec[q]) cudaGraphExecDestroy(exec[q]);
                    exec[q] = nullptr;
                }
                break;
            }
        }
    }
    const bool replay = exec[0] && exec[1];
    cudaStream_t loop_stream = replay ? gs : stream;
    for (int it = 0; it < config->max_iterations; ++it) {
        const int slot = it & 1;
        if (replay) {
            if (cudaGraphLaunch(exec[slot], gs) != cudaSuccess) {
                cudaGetLastError();
                rc = static_cast<int>(SpMVError::KERNEL_LAUNCH);
                break;
            }
            count_launches(static_cast<int>(launches_per_iteration));
        } else {
            rc = enqueue_iteration(r_old, r_new, slot, stream);
            if (rc != 0) break;  // the reference also leaves the loop on a failed SpMV (:105-107)
        }
        cudaEventRecord(ev[slot], loop_stream);
        if (pending.valid) {
            const bool done = settle(pending);
            pending.valid = false;
            if (done) {
                conv = rc == 0;
                break;
            }
        }
        pending = {it + 1, slot, r_new, true};
        float* t = r_old; r_old = r_new; r_new = t;  // :130-131
    }
    if (pending.valid && rc == 0) conv = settle(pending) && rc == 0;
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (gs) cudaStreamSynchronize(gs);  // a speculative iteration may still be running
    for (int q = 0; q < 2; ++q)
        if (exec[q]) cudaGraphExecDestroy(exec[q]);
    if (gs) cudaStreamDestroy(gs);
    // final vector (:135-139) and normalisation (:142-150)
    if (normalize) launch_normalize(fin, n, d_ranks, plan->tmp, stream);
    else cudaMemcpyAsync(d_ranks, fin, sizeof(float) * n, cudaMemcpyDeviceToDevice, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (rc == 0) rc = static_cast<int>(SpMVError::KERNEL_LAUNCH);
    }
    if (iterations) *iterations = iters;
    if (final_residual) *final_residual = residual;
    if (converged) *converged = conv;
    if (l1_residual) *l1_residual = l1;
    cleanup();
    return rc;
}

}  // namespace b200

// -------------------------------------------------------------------- pagerank --
// API-compatible wrapper: result.ranks is a host new[] array (pagerank_free
// deletes it).  adj must already be on the device; unlike the reference the
// host arrays are not needed (the dangling scan runs on the device).
PageRankResult pagerank(const CSRMatrix* adj_matrix, const PageRankConfig* config) {
    PageRankResult result;
    if (!adj_matrix) return result;
    const int n = adj_matrix->num_rows;
    result.ranks = new float[n > 0 ? n : 0];
    if (n <= 0) return result;

    const float init = 1.0f / n;
    auto fill_initial = [&]() {  // what the reference returns when its first SpMV fails (:105-107, :135-150)
        for (int i = 0; i < n; ++i) result.ranks[i] = init;
        float sum = 0.0f;
        for (int i = 0; i < n; ++i) sum += result.ranks[i];
        if (sum > 0.0f)
            for (int i = 0; i < n; ++i) result.ranks[i] /= sum;
    };

    float* d_ranks = nullptr;
    if (cudaMalloc(&d_ranks, sizeof(float) * n) != cudaSuccess) throw CudaException(cudaGetLastError());
    int iters = 0;
    float residual = 0.0f;
    bool conv = false;
    // Up to 2^24 nodes the final normalisation is the reference's own host code (sequential fp32
    // sum, :142-150) on the un-normalised vector, so the returned ranks carry the same rounding as
    // the reference's; beyond that its fp32 sum is meaningless (0.25 at n = 2^26) and the
    // f64 device normalisation is used.
    const bool literal_normalisation = n <= (1 << 24);
    const int rc = b200::pagerank_device(adj_matrix, config, d_ranks, &iters, &residual, &conv, nullptr,
                                         !literal_normalisation);
    if (rc != 0 && iters == 0) {
        cudaFree(d_ranks);
        fill_initial();
        return result;
    }
    cudaError_t e = cudaMemcpy(result.ranks, d_ranks, sizeof(float) * n, cudaMemcpyDeviceToHost);
    cudaFree(d_ranks);
    if (e != cudaSuccess) throw CudaException(e);
    if (literal_normalisation) {
        float sum = 0.0f;
        for (int i = 0; i < n; ++i) sum += result.ranks[i];
        if (sum > 0.0f)
            for (int i = 0; i < n; ++i) result.ranks[i] /= sum;
    }
    result.iterations = iters;
    result.final_residual = residual;
    result.converged = conv;
    return result;
}

}  // namespace spmv
