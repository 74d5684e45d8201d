// spmv/benchmark.h -- drop-in include path of the LessUp/gpu-spmv API.
// Everything is declared in spmv_b200/api.hpp; this file only forwards.
#ifndef SPMV_B200_FWD_BENCHMARK_H
#define SPMV_B200_FWD_BENCHMARK_H
#include "../spmv_b200/api.hpp"
#endif
