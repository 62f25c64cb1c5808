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
  float* s9;           // EPI_DELTA: [2][9][px] partial dots of the second delta conv (s9_index)
  unsigned long long* prof;   // optional (tools/conv_roles.py): per-role wait/work cycle counters of CTA 0, else null
  // Tile-level dependencies between consecutive tcgen05 convs of one iteration (same 16 x 8 tiling).  flags_out: this
  // launch sets flags_out[tile] = 1 once every store of the tile is visible.  flags_in: instead of waiting for the whole
  // upstream grid (griddepcontrol.wait), the roles that read its output wait for the 3 x 3 tile neighbourhood, so this
  // grid's CTAs start on the SMs the upstream grid's tail leaves idle.  Both null: plain grid-level dependency.
  int* flags_out;
  const int* flags_in;
};

// Partial dots of the second delta conv: plane-major [part 2][tap 9][px], so that the consumer's nine neighbour reads
// per pixel are nine coalesced row reads per warp (the pixel-major layout of round 1 cost 18 sector-scattered loads).
__host__ __device__ __forceinline__ long long s9_index(long long px, int part, int t, long long p) {
  return (long long)(part * 9 + t) * px + p;
}

constexpr int kFlagIters = 64;      // iterations per stage with tile flags (later ones fall back to grid dependencies)
constexpr int kFlagKernels = 3;     // corr-encoder 3x3, gates, q/GRU publish; gates, q/GRU, delta consume
inline int flag_tiles(int h, int w) { return ((h + 15) / 16) * ((w + 7) / 8); }   // the tcgen05 convs' 16 x 8 tiling

struct UpdateWs {
  __half *dn, *e1, *e, *z, *rnet;
  float *qx, *s9;
  int* flags;                         // [kFlagIters][kFlagKernels][flag_tiles(h, w)]
  size_t flags_bytes;
  size_t total;
};

inline UpdateWs carve_ws(void* base, int h, int wd) {
  const long long px = (long long)h * wd;
  UpdateWs w{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return (char*)base + r; };
  w.dn = (__half*)take(px * 64 * 2);
  w.e1 = (__half*)take(px * 64 * 2);
  w.e = (__half*)take(px * 64 * 2);
  w.z = (__half*)take(px * 64 * 2);
  w.rnet = (__half*)take(px * 64 * 2);
  w.qx = (float*)take(px * 64 * 4);
  w.s9 = (float*)take(px * 18 * 4);   // [2 column halves][9 taps][px] (the mma.sync path fills half 0 only)
  w.flags_bytes = (size_t)kFlagIters * kFlagKernels * flag_tiles(h, wd) * sizeof(int);
  w.flags = (int*)take(w.flags_bytes);
  w.total = o;
  return w;
}

// tcgen05 variant (update_tc.cu)
int tc_configure();
int tc_num_tiles(int h, int w);
template <int N, int EPI>
int launch_conv_tc(const ConvArgs& a, cudaStream_t stream);
int launch_conv_tc_dispatch(int n, int epi, const ConvArgs& a, cudaStream_t stream);

// 0 = mma.sync (v1), 1 = tcgen05 (v2); process-wide, set by cer_set_conv_variant / CER_CONV env
int conv_variant();

}  // namespace cer
