// Micro-benchmark: back-to-back tcgen05.mma (M128, N, K16, f16) on static smem operands with the
// conv kernel's descriptor strides -> cycles per MMA (is the no-swizzle K-major operand fetch full rate?)
#include <cstdio>
#include <cuda_runtime.h>
#include "../cer_mvs_b200/csrc/tc_common.cuh"
using namespace cer;
namespace cer { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } thread_local long long g_launches = 0; thread_local KernelTimer* g_timer = nullptr; void timer_begin(int, cudaStream_t) {} void timer_end(cudaStream_t) {} int g_pdl = 0; }

template <int N>
__global__ void __launch_bounds__(128, 1) k(int iters, int a_lbo, int a_sbo, int b_lbo, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t s0 = smem_addr_u32(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_addr_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_u32(&tslot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t a0 = s0 + (i % 9) * 16;
      const uint32_t b0 = s0 + 100 * 1024 + (i % 4) * 24576 % (90 * 1024);
      const uint64_t ad = umma_desc(a0 + 2 * (i & 3) * a_lbo, a_lbo, a_sbo);
      const uint64_t bd = umma_desc(b0 + 2 * (i & 3) * b_lbo, b_lbo, 128);
      tc_mma_f16(tm, ad, bd, idesc, 1u);
    }
    tc_commit(smem_addr_u32(&bar));
    mbar_wait(smem_addr_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u));
}

template <int N> void run(const char* name, int a_lbo, int a_sbo, int b_lbo, int grid) {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 20000;
  k<N><<<grid, 128, 200 * 1024>>>(iters, a_lbo, a_sbo, b_lbo, d);
  cudaDeviceSynchronize();
  k<N><<<grid, 128, 200 * 1024>>>(iters, a_lbo, a_sbo, b_lbo, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-44s N=%3d grid=%3d: %7.1f cycles/MMA (ideal %d)  %s\n", name, N, grid, (double)h / iters, 128 * N / 256, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<192>("conv strides (A lbo 2880 sbo 160, B lbo 3072)", 2880, 160, 192 * 16, 148);
  run<192>("conv strides, 1 CTA only", 2880, 160, 192 * 16, 1);
  run<192>("dense A (lbo 2048 sbo 128)", 2048, 128, 192 * 16, 148);
  run<256>("conv strides N=256", 2880, 160, 256 * 16, 148);
  run<64>("conv strides N=64", 2880, 160, 64 * 16, 148);
  run<128>("conv strides N=128", 2880, 160, 128 * 16, 148);
  return 0;
}
