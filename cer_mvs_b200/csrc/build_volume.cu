// Fused epipolar cost-volume build: CorrBlock.__init__ of the reference (core/corr.py:46-97) in
// one kernel -- hypotheses, projection, clamp, bilinear gather, 64-channel dot, view sum, D-minor
// layout.  No coords tensor, no per-view transposes, no global read-modify-write.
//
// Mapping: an 8-lane group owns one reference pixel.  Lane L of the group computes the epipolar
// sample position of hypothesis d0+L (so the projective arithmetic is done once per sample, not
// once per lane), the group then walks the 8 samples: every lane reads its 8-channel slice
// (16 B fp16 / 32 B fp32) of the four corner pixels -- a warp instruction touches 4 source pixels
// = 4 full 128-byte lines (fp16) -- and accumulates slice dots weighted by the bilinear weights.
// A 7-shuffle butterfly leaves the full dot of sample d0+L in lane L, which adds it to its
// view-sum accumulator and finally stores 8 consecutive hypotheses (32 B) per group.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cer {

constexpr int kMaxPairs = 64;
constexpr int kChunksPerBlock = 2;  // hypothesis chunks (of 8) handled by one block in sequence

// 8 channels of one pixel and their dot product with another pixel's 8 channels.
template <typename T>
struct FeatSlice;

// fp16 features: Blackwell's mixed-precision FMA (PTX fma.rn.f32.f16 -> SASS FHFMA) multiplies two fp16
// values exactly and accumulates in fp32 -- the same result as fmaf(float(a), float(b), c) without the
// 64 conversions per corner row.
__device__ __forceinline__ float fhfma_lo(uint32_t a, uint32_t b, float c) {
  float r;
  asm("{\n .reg .b16 al, ah, bl, bh;\n mov.b32 {al, ah}, %1;\n mov.b32 {bl, bh}, %2;\n fma.rn.f32.f16 %0, al, bl, %3;\n}"
      : "=f"(r) : "r"(a), "r"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fhfma_hi(uint32_t a, uint32_t b, float c) {
  float r;
  asm("{\n .reg .b16 al, ah, bl, bh;\n mov.b32 {al, ah}, %1;\n mov.b32 {bl, bh}, %2;\n fma.rn.f32.f16 %0, ah, bh, %3;\n}"
      : "=f"(r) : "r"(a), "r"(b), "f"(c));
  return r;
}

// base + 32-bit byte offset as ONE instruction (IMAD.WIDE.U32) instead of a 64-bit add pair
__device__ __forceinline__ const void* ptr_add_u32(const void* base, uint32_t byte_off) {
  uint64_t r;
  asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(byte_off), "l"(reinterpret_cast<uint64_t>(base)));
  return reinterpret_cast<const void*>(r);
}

template <>
struct FeatSlice<__half> {
  uint4 v;
  __device__ __forceinline__ void load(const __half* p) { v = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void load_off(const __half* base, uint32_t byte_off) {
    v = __ldg(reinterpret_cast<const uint4*>(ptr_add_u32(base, byte_off)));
  }
  __device__ __forceinline__ float dot(const FeatSlice& o) const {
    float d = fhfma_lo(v.x, o.v.x, 0.f);
    d = fhfma_hi(v.x, o.v.x, d);
    d = fhfma_lo(v.y, o.v.y, d);
    d = fhfma_hi(v.y, o.v.y, d);
    d = fhfma_lo(v.z, o.v.z, d);
    d = fhfma_hi(v.z, o.v.z, d);
    d = fhfma_lo(v.w, o.v.w, d);
    d = fhfma_hi(v.w, o.v.w, d);
    return d;
  }
};

template <>
struct FeatSlice<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void load_off(const float* base, uint32_t byte_off) {
    load(reinterpret_cast<const float*>(ptr_add_u32(base, byte_off)));
  }
  __device__ __forceinline__ float dot(const FeatSlice& o) const {
    float d = a.x * o.a.x;
    d = fmaf(a.y, o.a.y, d);
    d = fmaf(a.z, o.a.z, d);
    d = fmaf(a.w, o.a.w, d);
    d = fmaf(b.x, o.b.x, d);
    d = fmaf(b.y, o.b.y, d);
    d = fmaf(b.z, o.b.z, d);
    d = fmaf(b.w, o.b.w, d);
    return d;
  }
};

template <typename T>
__global__ void __launch_bounds__(256) build_volume_kernel(
    const T* __restrict__ feats, const float* __restrict__ Pij, const int* __restrict__ ii,
    const int* __restrict__ jj, int n_pairs, const float* __restrict__ disp_in, int shift, int D, float incre,
    float lo_origin, float* __restrict__ origin_out, float* __restrict__ volume, float out_scale, int per_view,
    int h, int w) {
  __shared__ float sP[kMaxPairs][12];
  __shared__ int sI[kMaxPairs], sJ[kMaxPairs];
  __shared__ __align__(16) int sS[32 * 8 * 8];   // per 8-lane group: 8 samples x {4 offsets, 4 weights}
  for (int t = threadIdx.x; t < n_pairs * 12; t += blockDim.x) sP[t / 12][t % 12] = Pij[(t / 12) * 16 + (t % 12)];
  for (int t = threadIdx.x; t < n_pairs; t += blockDim.x) {
    sI[t] = ii[t];
    sJ[t] = jj[t];
  }
  __syncthreads();

  const int lane = threadIdx.x & 7;
  const long long px = (long long)h * w;
  const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool valid = pix < px;
  const long long p = valid ? pix : px - 1;
  const int x = (int)(p % w), y = (int)(p / w);
  const float xf = (float)x, yf = (float)y;

  // hypothesis origin (core/corr.py:59-63)
  const float din = __ldg(disp_in + p);
  const float org = shift ? (din < lo_origin ? lo_origin : din) : din;
  if (valid && lane == 0 && blockIdx.y == 0) origin_out[p] = org;

  const long long img_stride = px * kFeatC;
  int cached_ref = -1;
  FeatSlice<T> f1;

  const int chunk0 = blockIdx.y * kChunksPerBlock;
  for (int ch = chunk0; ch < chunk0 + kChunksPerBlock; ++ch) {
    const int d0 = ch * 8;
    if (d0 >= D) break;
    const int d = d0 + lane;
    // d_k = origin + (k - D//2) * incre, rounded like the reference's two tensor ops (corr.py:56,66)
    const float dval = __fadd_rn(__fmul_rn((float)(min(d, D - 1) - D / 2), incre), org);
    float acc = 0.f;
    for (int k = 0; k < n_pairs; ++k) {
      const int ri = sI[k];
      if (ri != cached_ref) {  // block-uniform
        f1.load(feats + ri * img_stride + p * kFeatC + lane * 8);
        cached_ref = ri;
      }
      const T* img2 = feats + sJ[k] * img_stride + lane * 8;
      const float* P = sP[k];
      // X = Pij . (x, y, 1, d)   (utils/projective_ops.py:25-27), then /X2 and clamp (corr.py:88)
      const float X0 = fmaf(P[3], dval, fmaf(P[1], yf, P[0] * xf) + P[2]);
      const float X1 = fmaf(P[7], dval, fmaf(P[5], yf, P[4] * xf) + P[6]);
      const float X2 = fmaf(P[11], dval, fmaf(P[9], yf, P[8] * xf) + P[10]);
      float u = __fdiv_rn(X0, X2), v = __fdiv_rn(X1, X2);
      u = u < -1e4f ? -1e4f : (u > 1e4f ? 1e4f : u);  // NaN-preserving clamp
      v = v < -1e4f ? -1e4f : (v > 1e4f ? 1e4f : v);
      const float fu = floorf(u), fv = floorf(v);
      const float dx = u - fu, dy = v - fv;
      const int ix = (int)fu, iy = (int)fv;
      // The owner lane resolves the sample once and publishes it through shared memory (two 16-byte broadcast
      // reads per lane instead of shuffles + per-lane bounds logic):
      //   * element offsets of the four corner pixels, clamped into fmap2 so every load is legal and unpredicated;
      //   * row / column weights with out-of-range rows / columns zeroed: (dot * wy) * wx is then exactly 0 for a
      //     corner outside fmap2, which is what the reference adds for it (zero features, kernel.cu:81-84).
      //     NaN coordinates keep NaN weights, so the output is NaN like the reference's.
      {
        const int y0c = min(max(iy, 0), h - 1), y1c = min(max(iy + 1, 0), h - 1);
        const int x0c = min(max(ix, 0), w - 1), x1c = min(max(ix + 1, 0), w - 1);
        const float wy0 = (iy >= 0 && iy < h) ? 1.f - dy : ((dy != dy) ? dy : 0.f);
        const float wy1 = (iy + 1 >= 0 && iy + 1 < h) ? dy : ((dy != dy) ? dy : 0.f);
        const float wx0 = (ix >= 0 && ix < w) ? 1.f - dx : ((dx != dx) ? dx : 0.f);
        const float wx1 = (ix + 1 >= 0 && ix + 1 < w) ? dx : ((dx != dx) ? dx : 0.f);
        constexpr int kPixBytes = kFeatC * (int)sizeof(T);     // byte offsets (an image is < 4 GB)
        int4 o4 = make_int4((y0c * w + x0c) * kPixBytes, (y0c * w + x1c) * kPixBytes, (y1c * w + x0c) * kPixBytes,
                            (y1c * w + x1c) * kPixBytes);
        int4* slot = reinterpret_cast<int4*>(sS + (threadIdx.x >> 3) * 8 * 8 + lane * 8);
        slot[0] = o4;
        reinterpret_cast<float4*>(slot)[1] = make_float4(wy0, wy1, wx0, wx1);
      }
      __syncwarp();

      float part[8];
      const int* grp = sS + (threadIdx.x >> 3) * 8 * 8;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int4 o4 = *reinterpret_cast<const int4*>(grp + s * 8);
        const float4 wt = *reinterpret_cast<const float4*>(grp + s * 8 + 4);
        FeatSlice<T> f2;
        f2.load_off(img2, (uint32_t)o4.x);
        const float d00 = f1.dot(f2);
        f2.load_off(img2, (uint32_t)o4.y);
        const float d01 = f1.dot(f2);
        f2.load_off(img2, (uint32_t)o4.z);
        const float d10 = f1.dot(f2);
        f2.load_off(img2, (uint32_t)o4.w);
        const float d11 = f1.dot(f2);
        // (dot * wy) * wx per corner (correlation_kernel.cu:97-100)
        part[s] = ((d00 * wt.x) * wt.z + (d01 * wt.x) * wt.w) + ((d10 * wt.y) * wt.z + (d11 * wt.y) * wt.w);
      }
      __syncwarp();   // the slots are rewritten for the next view
      // butterfly: lane L ends with the group-wide total of sample L
      float k4[4], k2[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = (lane & 4) ? part[i] : part[i + 4];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
        k4[i] = ((lane & 4) ? part[i + 4] : part[i]) + recv;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = (lane & 2) ? k4[i] : k4[i + 2];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        k2[i] = ((lane & 2) ? k4[i + 2] : k4[i]) + recv;
      }
      const float send = (lane & 1) ? k2[0] : k2[1];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
      const float total = ((lane & 1) ? k2[1] : k2[0]) + recv;
      if (per_view) {
        if (valid && d < D) volume[((long long)k * px + p) * D + d] = total * out_scale;
      } else {
        acc += total;
      }
    }
    if (!per_view && valid && d < D) volume[p * D + d] = acc * out_scale;
  }
}

// ------------------------------------------------------------------------------------------
// fp16 features, 4 lanes per pixel: every lane owns 16 channels and fetches them with ONE 256-bit load
// (LDG.E.256, sm_100), so a warp instruction still covers 8 full 128-byte rows (one L1 wavefront per row) but the
// per-sample bookkeeping is shared by 4 lanes instead of 8 and there are half as many load instructions.
// ------------------------------------------------------------------------------------------
struct Slice16 {
  uint32_t v[8];
  __device__ __forceinline__ void load(const void* p) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
  }
  __device__ __forceinline__ float dot(const Slice16& o) const {
    float d = fhfma_lo(v[0], o.v[0], 0.f);
    d = fhfma_hi(v[0], o.v[0], d);
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      d = fhfma_lo(v[i], o.v[i], d);
      d = fhfma_hi(v[i], o.v[i], d);
    }
    return d;
  }
};

constexpr int kH16Chunks = 4;   // chunks of 4 hypotheses per block in sequence

// The L1-gather formulation (cer_set_build_variant(1) / CER_BUILD=gather; the A/B partner of the shared-memory-staged
// tcgen05 build in build_volume_tc.cu and the path used for row bands of odd geometry): a 4-lane group owns a pixel, the
// owner lane projects hypothesis d0 + lane and publishes clamped corner offsets + weights through shared memory, every
// lane fetches 16 channels of a corner row with one 256-bit load.
__global__ void __launch_bounds__(256, 4) build_volume_h16_kernel(
    const __half* __restrict__ feats, const float* __restrict__ Pij, const int* __restrict__ ii,
    const int* __restrict__ jj, int n_pairs, const float* __restrict__ disp_in, int shift, int D, float incre,
    float lo_origin, float* __restrict__ origin_out, float* __restrict__ volume, float out_scale, int per_view,
    int h, int w, int y_begin, int d_begin, int d_end, int accumulate, const int* __restrict__ tile_only) {
  // second pass behind the staged kernel: only the 16 x 8 tiles it flagged (this block's 8 x 8 tile is half of one)
  if (tile_only != nullptr) {
    const int tx8 = (w + 7) >> 3, tx16 = (w + 15) >> 4;
    if (tile_only[(blockIdx.y / tx8) * tx16 + ((blockIdx.y % tx8) >> 1)] == 0) return;
  }
  __shared__ float sP[kMaxPairs][12];
  __shared__ int sI[kMaxPairs], sJ[kMaxPairs];
  // per 4-lane group: 4 samples x {4 byte offsets, 4 weights} = 32 words, padded to 36: a quarter-warp (two groups)
  // then reads / writes its two 16-byte slots in different banks
  constexpr int kGrpWords = 36;
  __shared__ __align__(16) int sS[64 * kGrpWords];
  for (int t = threadIdx.x; t < n_pairs * 12; t += blockDim.x) sP[t / 12][t % 12] = Pij[(t / 12) * 16 + (t % 12)];
  for (int t = threadIdx.x; t < n_pairs; t += blockDim.x) {
    sI[t] = ii[t];
    sJ[t] = jj[t];
  }
  __syncthreads();

  const int lane = threadIdx.x & 3;
  const long long px = (long long)h * w;
  // a block = an 8 x 8 pixel tile (a warp = one tile row); grid = (hypothesis-chunk groups, pixel tiles)
  const int tiles_x = (w + 7) >> 3;
  const int pl = threadIdx.x >> 2;
  const int x_ = (blockIdx.y % tiles_x) * 8 + (pl & 7), y_ = y_begin + (blockIdx.y / tiles_x) * 8 + (pl >> 3);
  const bool valid = x_ < w && y_ < h;
  const int x = valid ? x_ : w - 1, y = valid ? y_ : h - 1;
  const long long p = (long long)y * w + x;
  const float xf = (float)x, yf = (float)y;
  const float din = __ldg(disp_in + p);
  const float org = shift ? (din < lo_origin ? lo_origin : din) : din;      // core/corr.py:59-63
  if (valid && lane == 0 && blockIdx.x == 0) origin_out[p] = org;

  const long long img_stride = px * kFeatC;
  int cached_ref = -1;
  Slice16 f1;
  int* grp = sS + (threadIdx.x >> 2) * kGrpWords;

  const int chunk0 = blockIdx.x * kH16Chunks;
  for (int ch = chunk0; ch < chunk0 + kH16Chunks; ++ch) {
    const int d0 = d_begin + ch * 4;
    if (d0 >= d_end) break;
    const int d = d0 + lane;
    const float dval = __fadd_rn(__fmul_rn((float)(min(d, D - 1) - D / 2), incre), org);   // corr.py:56,66
    float acc = 0.f;
    for (int k = 0; k < n_pairs; ++k) {
      const int ri = sI[k];
      if (ri != cached_ref) {
        f1.load(feats + ri * img_stride + p * kFeatC + lane * 16);
        cached_ref = ri;
      }
      const __half* img2 = feats + sJ[k] * img_stride + lane * 16;
      const float* P = sP[k];
      // X = Pij . (x, y, 1, d), /X2, clamp (corr.py:88), floor; clamped corner offsets (every load legal) and row / column
      // weights with out-of-range rows / columns zeroed (NaN coordinates keep NaN weights like the reference)
      {
        const float bx = fmaf(P[1], yf, P[0] * xf) + P[2], by = fmaf(P[5], yf, P[4] * xf) + P[6];
        const float bz = fmaf(P[9], yf, P[8] * xf) + P[10];
        const float X0 = fmaf(P[3], dval, bx), X1 = fmaf(P[7], dval, by), X2 = fmaf(P[11], dval, bz);
        float u = __fdiv_rn(X0, X2), v = __fdiv_rn(X1, X2);
        u = u < -1e4f ? -1e4f : (u > 1e4f ? 1e4f : u);
        v = v < -1e4f ? -1e4f : (v > 1e4f ? 1e4f : v);
        const float fu = floorf(u), fv = floorf(v);
        const float dx = u - fu, dy = v - fv;
        const int ix = (int)fu, iy = (int)fv;
        const int y0c = min(max(iy, 0), h - 1), y1c = min(max(iy + 1, 0), h - 1);
        const int x0c = min(max(ix, 0), w - 1), x1c = min(max(ix + 1, 0), w - 1);
        float4 wt;
        wt.x = (iy >= 0 && iy < h) ? 1.f - dy : ((dy != dy) ? dy : 0.f);
        wt.y = (iy + 1 >= 0 && iy + 1 < h) ? dy : ((dy != dy) ? dy : 0.f);
        wt.z = (ix >= 0 && ix < w) ? 1.f - dx : ((dx != dx) ? dx : 0.f);
        wt.w = (ix + 1 >= 0 && ix + 1 < w) ? dx : ((dx != dx) ? dx : 0.f);
        int4* slot = reinterpret_cast<int4*>(grp + lane * 8);
        slot[0] = make_int4((y0c * w + x0c) * 128, (y0c * w + x1c) * 128, (y1c * w + x0c) * 128, (y1c * w + x1c) * 128);
        reinterpret_cast<float4*>(slot)[1] = wt;
        __syncwarp();
      }
      float part[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int4 o4 = *reinterpret_cast<const int4*>(grp + s * 8);
        const float4 wt = *reinterpret_cast<const float4*>(grp + s * 8 + 4);
        Slice16 f2;
        f2.load(ptr_add_u32(img2, (uint32_t)o4.x));
        const float d00 = f1.dot(f2);
        f2.load(ptr_add_u32(img2, (uint32_t)o4.y));
        const float d01 = f1.dot(f2);
        f2.load(ptr_add_u32(img2, (uint32_t)o4.z));
        const float d10 = f1.dot(f2);
        f2.load(ptr_add_u32(img2, (uint32_t)o4.w));
        const float d11 = f1.dot(f2);
        part[s] = ((d00 * wt.x) * wt.z + (d01 * wt.x) * wt.w) + ((d10 * wt.y) * wt.z + (d11 * wt.y) * wt.w);
      }
      __syncwarp();      // the slots are rewritten for the next view
      // butterfly over the 4 lanes: lane L ends with the total of sample L
      float k2[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = (lane & 2) ? part[i] : part[i + 2];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        k2[i] = ((lane & 2) ? part[i + 2] : part[i]) + recv;
      }
      const float send = (lane & 1) ? k2[0] : k2[1];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
      const float total = ((lane & 1) ? k2[1] : k2[0]) + recv;
      if (per_view) {
        if (valid && d < d_end) {
          float* o = volume + ((long long)k * px + p) * D + d;
          *o = accumulate ? *o + total * out_scale : total * out_scale;
        }
      } else {
        acc += total;
      }
    }
    if (!per_view && valid && d < d_end) {
      float* o = volume + p * D + d;
      *o = accumulate ? *o + acc * out_scale : acc * out_scale;
    }
  }
}

}  // namespace cer

namespace cer {
int build_volume_tc(const void* feats, const float* Pij, const int* ii, const int* jj, int n_pairs,
                    const float* disp_in, int shift, int D, float incre, float lo_origin, float* origin,
                    float* volume, float out_scale, int per_view, int h, int w, int y_begin, int y_end, int d_begin,
                    int d_end, int accumulate, int** tile_skip_out, cudaStream_t stream);
// 0 = shared-memory-staged source boxes + tcgen05 (build_volume_tc.cu, default for fp16 features);
// 1 = L1 row gather with FHFMA (this file).  fp32 features always take the generic gather kernel.
static int g_build_variant = -1;
static int build_variant() {
  if (g_build_variant < 0) {
    const char* e = getenv("CER_BUILD");
    g_build_variant = (e && !strcmp(e, "gather")) ? 1 : 0;
  }
  return g_build_variant;
}
}  // namespace cer

using namespace cer;

namespace cer {
void build_set_profile(unsigned long long* dev);
}
// Debug: per-role counters of the staged build kernel, 16 x u64 on the device (nullptr switches them off).
extern "C" int cer_debug_set_build_profile(unsigned long long* dev_counters) {
  cer::build_set_profile(dev_counters);
  return CER_OK;
}

extern "C" int cer_set_build_variant(int variant) {
  CER_REQUIRE(variant >= 0 && variant <= 1,
              "cer_set_build_variant: 0 TMA-staged source boxes + tcgen05 (default); 1 L1 row gather (FHFMA)");
  g_build_variant = variant;
  return CER_OK;
}

extern "C" int cer_build_volume_part(const void* feats, int feats_f16, const float* Pij, const int* ii, const int* jj,
                                     int n_pairs, const float* disp_in, int shift, int D, float incre, float lo_origin,
                                     float* origin, float* volume, float out_scale, int per_view, int h, int w,
                                     int d_begin, int d_end, int accumulate, cer_stream_t stream);

extern "C" int cer_build_volume(const void* feats, int feats_f16, const float* Pij, const int* ii, const int* jj,
                                int n_pairs, const float* disp_in, int shift, int D, float incre, float lo_origin,
                                float* origin, float* volume, float out_scale, int per_view, int h, int w,
                                cer_stream_t stream) {
  return cer_build_volume_part(feats, feats_f16, Pij, ii, jj, n_pairs, disp_in, shift, D, incre, lo_origin, origin,
                               volume, out_scale, per_view, h, w, 0, D, 0, stream);
}

// Hypotheses [d_begin, d_end) of the given views only, optionally ADDED to the volume: the unit of work of the sharded
// build (a rank owns a contiguous run of (view, hypothesis) units; partial volumes are summed by one all-reduce).
extern "C" int cer_build_volume_part(const void* feats, int feats_f16, const float* Pij, const int* ii, const int* jj,
                                     int n_pairs, const float* disp_in, int shift, int D, float incre, float lo_origin,
                                     float* origin, float* volume, float out_scale, int per_view, int h, int w,
                                     int d_begin, int d_end, int accumulate, cer_stream_t stream) {
  CER_REQUIRE(feats && Pij && ii && jj && disp_in && origin && volume, "cer_build_volume: null pointer");
  CER_REQUIRE(n_pairs > 0 && n_pairs <= kMaxPairs, "cer_build_volume: n_pairs must be 1..%d", kMaxPairs);
  CER_REQUIRE(D > 0 && h > 0 && w > 0, "cer_build_volume: bad sizes");
  CER_REQUIRE(d_begin >= 0 && d_begin < d_end && d_end <= D, "cer_build_volume_part: hypotheses [%d, %d) outside 0..%d",
              d_begin, d_end, D);
  CER_REQUIRE(aligned16(feats), "cer_build_volume: feats must be 16-byte aligned");
  const bool whole = d_begin == 0 && d_end == D && !accumulate;
  CER_REQUIRE(whole || feats_f16, "cer_build_volume_part: partial / accumulating builds need fp16 features");
  const long long px = (long long)h * w;
  // fp16 features: the staged tcgen05 kernel first; it flags the tiles whose hypothesis origins are too incoherent for a
  // common source box, and the gather kernel then computes exactly those (tile_only); CER_BUILD=gather: gather only
  int* tile_only = nullptr;
  if (feats_f16 && build_variant() == 0 && (reinterpret_cast<uintptr_t>(feats) & 127) == 0) {
    int rc = build_volume_tc(feats, Pij, ii, jj, n_pairs, disp_in, shift, D, incre, lo_origin, origin, volume, out_scale,
                             per_view, h, w, 0, h, d_begin, d_end, accumulate, &tile_only, (cudaStream_t)stream);
    if (rc) return rc;
  }
  if (feats_f16) {
    dim3 g16(ceil_div(ceil_div(d_end - d_begin, 4), kH16Chunks), ((w + 7) / 8) * ((h + 7) / 8));
    CER_REQUIRE(g16.y <= 65535u, "cer_build_volume: image too large for the tile grid (%u tiles)", g16.y);
    CER_LAUNCH(KK_BUILD, build_volume_h16_kernel, g16, 256, 0, stream, (const __half*)feats, Pij, ii, jj, n_pairs,
               disp_in, shift, D, incre, lo_origin, origin, volume, out_scale, per_view, h, w, 0, d_begin, d_end,
               accumulate, (const int*)tile_only);
  } else {
    const int chunks = ceil_div(D, 8);
    dim3 grid(ceil_div(px * 8, 256), ceil_div(chunks, kChunksPerBlock));
    CER_LAUNCH(KK_BUILD, build_volume_kernel<float>, grid, 256, 0, stream, (const float*)feats, Pij, ii, jj, n_pairs,
               disp_in, shift, D, incre, lo_origin, origin, volume, out_scale, per_view, h, w);
  }
  return check_launch("cer_build_volume");
}
