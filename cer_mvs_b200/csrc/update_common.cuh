// Shared between the HMMA (v1) and tcgen05 (v2) UpdateBlock kernels.
#pragma once
#include "common.cuh"
#include "update_blob.h"

namespace cer {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float h_round(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }

enum Epilogue { EPI_RELU = 0, EPI_GATES = 1, EPI_GRUOUT = 2, EPI_DELTA = 3 };

struct ConvArgs {
  const __half* src[4];
  int n_src;
  const __half* wpk;   // HMMA weights [n_src][9][64][N]
  const __half* wtc;   // tcgen05 weights: [n_src*9] UMMA K-major tiles [8][N][8]
  const __half* wtc2;  // CTA-pair tcgen05 weights: [n_src*9] tiles [2][8][N/2][8]
  const float* bias;   // [N] or null
  int h, w;
  const float* disp;   // tcgen05 gate conv: disparity map for the in-kernel disparity encoder
  int dn_chunk;        // index of the chunk generated from disp (-1: every chunk is read from src[])
  __half* out_h;       // EPI_RELU: [px][64]
  __half* net;         // EPI_GATES: read; EPI_GRUOUT: read + written in place
  __half* z;           // EPI_GATES: write; EPI_GRUOUT: read
  __half* rnet;        // EPI_GATES: write
  float* qx;           // EPI_GATES: write; EPI_GRUOUT: read
  const float* w2;     // EPI_DELTA: [9][256]
  float* s9;           // EPI_DELTA: [px][2][9] partial dots of the second delta conv
  unsigned long long* prof;   // optional (tools/conv_roles.py): per-role wait/work cycle counters of CTA 0, else null
};

struct UpdateWs {
  __half *dn, *e1, *e, *z, *rnet;
  float *qx, *s9;
  size_t total;
};

inline UpdateWs carve_ws(void* base, long long px) {
  UpdateWs w{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return (char*)base + r; };
  w.dn = (__half*)take(px * 64 * 2);
  w.e1 = (__half*)take(px * 64 * 2);
  w.e = (__half*)take(px * 64 * 2);
  w.z = (__half*)take(px * 64 * 2);
  w.rnet = (__half*)take(px * 64 * 2);
  w.qx = (float*)take(px * 64 * 4);
  w.s9 = (float*)take(px * 18 * 4);   // [px][2 column halves][9 taps] (the mma.sync path fills half 0 only)
  w.total = o;
  return w;
}

// tcgen05 variant (update_tc.cu)
int tc_configure();
template <int N, int EPI>
int launch_conv_tc(const ConvArgs& a, cudaStream_t stream);
int launch_conv_tc_dispatch(int n, int epi, const ConvArgs& a, cudaStream_t stream);

// 0 = mma.sync (v1), 1 = tcgen05 (v2); process-wide, set by cer_set_conv_variant / CER_CONV env
int conv_variant();

}  // namespace cer
