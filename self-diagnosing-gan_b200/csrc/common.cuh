// common.cuh -- error plumbing, launch accounting and small device helpers shared by all kernels.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/sdg.h"

namespace sdg {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

#define SDG_CUDA(expr)                                   \
  do {                                                   \
    int _rc = ::sdg::check((expr), #expr);               \
    if (_rc) return _rc;                                 \
  } while (0)

#define SDG_REQUIRE(cond, code, ...)                     \
  do {                                                   \
    if (!(cond)) {                                       \
      ::sdg::set_error(__VA_ARGS__);                     \
      return (code);                                     \
    }                                                    \
  } while (0)

// every kernel launch of the library goes through here so bench.py can report gpu_launches
#define SDG_LAUNCH(kernel, grid, block, smem, stream, ...)                    \
  do {                                                                        \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); \
    ::sdg::g_launches.fetch_add(1, std::memory_order_relaxed);                \
    SDG_CUDA(cudaGetLastError());                                             \
  } while (0)

// fp16 range guard (include/sdg.h, sdg_ctx_set_range_flag): the caller-owned device flag of the ctx an entry point is
// working for, visible to the kernel launchers of this thread for the duration of the call
extern thread_local int* t_range_flag;
struct RangeScope {
  explicit RangeScope(int* flag) { t_range_flag = flag; }
  ~RangeScope() { t_range_flag = nullptr; }
};
constexpr float kF16Max = 65504.0f;

constexpr int kNumSMs = 148;   // B200

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid for a grid-stride streaming kernel: whole waves of the 148 SMs, capped by the work
inline int stream_grid(int64_t work_items, int per_block, int blocks_per_sm = 8) {
  int64_t need = cdiv(work_items, per_block);
  int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace sdg
