// Shared host/device helpers for libdram_b200.so (internal; the public ABI is include/dram_b200.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "dram_b200.h"

namespace dram {

// Per-thread error text behind dram_last_error().
void set_error(const char *fmt, ...);

inline int check_cuda(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return DRAM_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return DRAM_E_LAUNCH;
}

#define DRAM_REQUIRE(cond, ...)       \
  do {                                \
    if (!(cond)) {                    \
      ::dram::set_error(__VA_ARGS__); \
      return DRAM_E_ARG;              \
    }                                 \
  } while (0)

#define DRAM_CHECK_LAUNCH(what)                                       \
  do {                                                                \
    int _rc = ::dram::check_cuda(cudaGetLastError(), what);           \
    if (_rc != DRAM_OK) return _rc;                                   \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grid for grid-stride memory-bound kernels: a multiple of the SM count.
int sm_count();
inline int stream_grid(int64_t work_items, int threads, int ctas_per_sm = 8) {
  int64_t want = ceil_div64(work_items, threads);
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace dram
