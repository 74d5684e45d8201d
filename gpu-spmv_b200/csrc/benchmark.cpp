// benchmark.cpp -- benchmark harness and its JSON form.
//
// Same protocol and result fields as the reference (src/benchmark.cu:21-237):
// copy the host x to the device once, num_warmup_runs untimed calls, num_runs
// timed calls of the blocking API, min / max / mean / sample (n-1) standard
// deviation of SpMVResult.elapsed_ms; gflops and bandwidth are those of the
// last successful run; JSON is fixed-point with 6 decimals and the key order
// name, execution_time_ms, gflops, bandwidth_gb_s, avg, min, max, stddev,
// num_runs; the reader is a "key": + stof scan.
#include "internal.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <iomanip>
#include <sstream>

namespace spmv {
namespace {

void fill_statistics(const std::vector<float>& samples, BenchmarkResult& out) {
    if (samples.empty()) return;
    out.num_runs = static_cast<int>(samples.size());
    out.min_time_ms = *std::min_element(samples.begin(), samples.end());
    out.max_time_ms = *std::max_element(samples.begin(), samples.end());
    float total = 0.0f;
    for (float t : samples) total += t;
    out.avg_time_ms = total / samples.size();
    out.execution_time_ms = out.avg_time_ms;
    float spread = 0.0f;
    if (samples.size() > 1) {
        float ss = 0.0f;
        for (float t : samples) {
            const float d = t - out.avg_time_ms;
            ss += d * d;
        }
        spread = std::sqrt(ss / (samples.size() - 1));
    }
    out.stddev_time_ms = spread;
}

// Runs `call` warm-up + timed times and folds the SpMVResults into `out`.
template <typename Call>
void run_protocol(const BenchmarkConfig& cfg, Call call, BenchmarkResult& out) {
    for (int i = 0; i < cfg.num_warmup_runs; ++i) call();
    std::vector<float> samples;
    samples.reserve(cfg.num_runs > 0 ? cfg.num_runs : 0);
    for (int i = 0; i < cfg.num_runs; ++i) {
        const SpMVResult r = call();
        if (r.error_code != static_cast<int>(SpMVError::SUCCESS)) continue;
        samples.push_back(r.elapsed_ms);
        out.gflops = r.gflops;
        out.bandwidth_gb_s = r.bandwidth_gb_s;
    }
    fill_statistics(samples, out);
}

}  // namespace

BenchmarkResult benchmark_csr(const CSRMatrix* A, const float* x, const SpMVConfig* config,
                              const BenchmarkConfig* bench_config) {
    BenchmarkResult out;
    out.name = "CSR SpMV";
    const BenchmarkConfig defaults;
    const BenchmarkConfig& cfg = bench_config ? *bench_config : defaults;
    CudaBuffer<float> d_x(A->num_cols), d_y(A->num_rows);
    d_x.copyFromHost(x, A->num_cols);
    run_protocol(cfg, [&] { return spmv_csr(A, d_x.get(), d_y.get(), config, A->num_cols); }, out);
    return out;
}

BenchmarkResult benchmark_ell(const ELLMatrix* A, const float* x, const BenchmarkConfig* bench_config) {
    BenchmarkResult out;
    out.name = "ELL SpMV";
    const BenchmarkConfig defaults;
    const BenchmarkConfig& cfg = bench_config ? *bench_config : defaults;
    CudaBuffer<float> d_x(A->num_cols), d_y(A->num_rows);
    d_x.copyFromHost(x, A->num_cols);
    run_protocol(cfg, [&] { return spmv_ell(A, d_x.get(), d_y.get(), nullptr, A->num_cols); }, out);
    return out;
}

// GPU numbers as above; the CPU side times the API's own single-threaded
// spmv_cpu_csr num_runs times (the reference brackets it with CUDA events,
// src/benchmark.cu:151-167; a steady clock measures the same interval).
ComparisonResult compare_gpu_cpu_csr(const CSRMatrix* A, const float* x, const SpMVConfig* config,
                                     const BenchmarkConfig* bench_config) {
    ComparisonResult cmp;
    cmp.gpu_result = benchmark_csr(A, x, config, bench_config);
    const BenchmarkConfig defaults;
    const BenchmarkConfig& cfg = bench_config ? *bench_config : defaults;

    cmp.cpu_result.name = "CPU CSR SpMV";
    std::vector<float> y(A->num_rows > 0 ? A->num_rows : 0);
    std::vector<float> samples;
    for (int i = 0; i < cfg.num_runs; ++i) {
        const auto t0 = std::chrono::steady_clock::now();
        spmv_cpu_csr(A, x, y.data());
        const auto t1 = std::chrono::steady_clock::now();
        samples.push_back(std::chrono::duration<float, std::milli>(t1 - t0).count());
    }
    fill_statistics(samples, cmp.cpu_result);
    if (cmp.gpu_result.avg_time_ms > 0.0f) cmp.speedup = cmp.cpu_result.avg_time_ms / cmp.gpu_result.avg_time_ms;
    return cmp;
}

std::string benchmark_to_json(const BenchmarkResult& r) {
    std::ostringstream js;
    js << std::fixed << std::setprecision(6);
    js << "{\n"
       << "  \"name\": \"" << r.name << "\",\n"
       << "  \"execution_time_ms\": " << r.execution_time_ms << ",\n"
       << "  \"gflops\": " << r.gflops << ",\n"
       << "  \"bandwidth_gb_s\": " << r.bandwidth_gb_s << ",\n"
       << "  \"avg_time_ms\": " << r.avg_time_ms << ",\n"
       << "  \"min_time_ms\": " << r.min_time_ms << ",\n"
       << "  \"max_time_ms\": " << r.max_time_ms << ",\n"
       << "  \"stddev_time_ms\": " << r.stddev_time_ms << ",\n"
       << "  \"num_runs\": " << r.num_runs << "\n"
       << "}";
    return js.str();
}

std::string comparison_to_json(const ComparisonResult& c) {
    std::ostringstream js;
    js << std::fixed << std::setprecision(6);
    js << "{\n"
       << "  \"gpu\": " << benchmark_to_json(c.gpu_result) << ",\n"
       << "  \"cpu\": " << benchmark_to_json(c.cpu_result) << ",\n"
       << "  \"speedup\": " << c.speedup << "\n"
       << "}";
    return js.str();
}

// Minimal reader for the writer above: value after the first `"key":`.
BenchmarkResult benchmark_from_json(const std::string& json) {
    auto number_after = [&json](const char* key) -> float {
        const std::string tag = std::string("\"") + key + "\":";
        const size_t at = json.find(tag);
        if (at == std::string::npos) return 0.0f;
        return std::stof(json.substr(at + tag.size()));
    };
    BenchmarkResult r;
    r.execution_time_ms = number_after("execution_time_ms");
    r.gflops = number_after("gflops");
    r.bandwidth_gb_s = number_after("bandwidth_gb_s");
    r.avg_time_ms = number_after("avg_time_ms");
    r.min_time_ms = number_after("min_time_ms");
    r.max_time_ms = number_after("max_time_ms");
    r.stddev_time_ms = number_after("stddev_time_ms");
    r.num_runs = static_cast<int>(number_after("num_runs"));
    return r;
}

}  // namespace spmv
