// Shared helpers for the cer_mvs_b200 kernels (sm_100a only, no torch headers).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cer_mvs_b200.h"

namespace cer {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define CER_REQUIRE(cond, ...)                     \
  do {                                             \
    if (!(cond)) {                                 \
      cer::set_error(__VA_ARGS__);                 \
      return CER_ERR_INVALID;                      \
    }                                              \
  } while (0)

#define CER_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      cer::set_error("%s failed: %s", #call, cudaGetErrorString(e_));               \
      return (int)e_;                                                               \
    }                                                                               \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kFeatC = 64;        // dim_fmap (core/raft.py:18)
constexpr int kNumSMs = 148;      // B200

// launch counter (cer_plan_last_launch_count): every kernel launch in the library goes through LAUNCH
extern thread_local long long g_launches;

}  // namespace cer

#define CER_LAUNCH(kernel, grid, block, smem, stream, ...)            \
  do {                                                                \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); \
    ++cer::g_launches;                                                \
  } while (0)
