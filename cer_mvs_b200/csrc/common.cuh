// Shared helpers for the cer_mvs_b200 kernels (sm_100a only, no torch headers).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

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
      cudaGetLastError(); /* do not leave a sticky error for the host framework */  \
      cer::set_error("%s failed: %s", #call, cudaGetErrorString(e_));               \
      return (int)e_;                                                               \
    }                                                                               \
  } while (0)

// One-time-per-DEVICE guard for cudaFuncSetAttribute (the attribute lives in the device's context, so a process-wide
// flag would leave the kernels of a second GPU unconfigured): true exactly once per (guard, current device).
static inline bool first_time_on_device(std::atomic<unsigned long long>& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  return !(mask.fetch_or(bit) & bit);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kFeatC = 64;        // dim_fmap (core/raft.py:18)
constexpr int kNumSMs = 148;      // B200

// launch counter (cer_plan_last_launch_count): every kernel launch in the library goes through CER_LAUNCH
extern thread_local long long g_launches;

// Kernel classes, for the per-kernel CUDA-event timing of cer_plan_set_kernel_timing().
enum KernelKind {
  KK_LAYOUT = 0, KK_PROJ, KK_BUILD, KK_POOL, KK_LOOKUP, KK_CORR_DROPIN, KK_DISP_ENC, KK_CORR_ENC1, KK_CONV_E,
  KK_CONV_GATES, KK_CONV_Q, KK_CONV_DELTA, KK_DISP_UPDATE, KK_FINISH, KK_COUNT
};
struct KernelTimer;                       // defined in plan.cu
extern thread_local KernelTimer* g_timer; // non-null while a plan runs eagerly with timing enabled
void timer_begin(int kind, cudaStream_t s);
void timer_end(cudaStream_t s);

// Programmatic dependent launch (PDL): a kernel launched with CER_LAUNCH_PDL may start while its predecessor in
// the stream is still draining; it must call pdl_wait() before touching anything the predecessor wrote (or
// writing anything the predecessor reads).  Launch latency, barrier/TMEM set-up and constant-weight prefetch
// then overlap the predecessor's tail.  g_pdl = 0 (CER_PDL=0) falls back to plain stream order.
extern int g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace cer

#define CER_LAUNCH_PDL(kind, kernel, grid, block, smem, stream, ...)                         \
  do {                                                                                       \
    if (cer::g_timer) cer::timer_begin((kind), (cudaStream_t)(stream));                      \
    cudaLaunchConfig_t cfg_ = {};                                                            \
    cfg_.gridDim = dim3(grid);                                                               \
    cfg_.blockDim = dim3(block);                                                             \
    cfg_.dynamicSmemBytes = (smem);                                                          \
    cfg_.stream = (cudaStream_t)(stream);                                                    \
    cudaLaunchAttribute attr_[1];                                                            \
    attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                        \
    attr_[0].val.programmaticStreamSerializationAllowed = 1;                                 \
    cfg_.attrs = attr_;                                                                      \
    cfg_.numAttrs = cer::g_pdl ? 1 : 0;                                                      \
    cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                                          \
    if (cer::g_timer) cer::timer_end((cudaStream_t)(stream));                                \
    ++cer::g_launches;                                                                       \
  } while (0)

#define CER_LAUNCH(kind, kernel, grid, block, smem, stream, ...)                \
  do {                                                                          \
    if (cer::g_timer) cer::timer_begin((kind), (cudaStream_t)(stream));         \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);   \
    if (cer::g_timer) cer::timer_end((cudaStream_t)(stream));                   \
    ++cer::g_launches;                                                          \
  } while (0)
