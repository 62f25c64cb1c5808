// UpdateBlock.forward (core/update.py:87-120), v1: implicit-GEMM 3x3 convolutions on mma.sync
// (m16n8k16, fp16 operands, fp32 accumulate) with the element-wise GRU algebra fused into the
// epilogues.  NHWC fp16 activations, 64 channels per tensor, one CTA = 8 x 16 pixels.
//
// Kernels per iteration:
//   K0 disp_encode     dn = fp16(100 * (unfold7x7(disp) - disp))                      update.py:80-85,97
//   K1 corr_enc1       e1 = relu(conv1x1(mean_v corr))                                update.py:103,62-63
//   K2 conv3x3<64>     e  = relu(conv3x3(e1))                                         update.py:64-65
//   K3 conv3x3<192>    z = sigma(.), r*net, qx = convq over [inp|dn|e] + bq           update.py:20-23
//   K4 conv3x3<64>     q = tanh(convq(r*net) + qx); net = (1-z) net + z q             update.py:23-24
//   K5 conv3x3<256>    d = relu(conv3x3(net)); s9[p][t] = w2[t] . d[p]                update.py:69-70
//   K6 disp_update     delta = 0.01 * (b + sum_t s9[p+off_t][t]); disp += delta       update.py:71,114 raft.py:101
#include <string.h>

#include <stdlib.h>

#include "lookup_common.cuh"
#include "tc_common.cuh"
#include "update_common.cuh"

namespace cer {

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}


// ------------------------------------------------------------------------------------------
// K0: disparity-neighbourhood encoder -> [px][64] fp16 (channels 49..63 zero)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) disp_encode_kernel(const float* __restrict__ disp, __half* __restrict__ dn,
                                                         int h, int w) {
  const long long px = (long long)h * w;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // pixel*8 + group of 8 channels
  if (idx >= px * 8) return;
  const long long p = idx >> 3;
  const int g = (int)(idx & 7);
  const int x = (int)(p % w), y = (int)(p / w);
  const float c = __ldg(disp + p);
  __align__(16) __half v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = g * 8 + e;
    float val = 0.f;
    if (k < kDispEnc) {
      const int yy = y + k / 7 - 3, xx = x + k % 7 - 3;
      const float nb = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(disp + (long long)yy * w + xx) : 0.f;
      val = __fmul_rn(100.f, __fsub_rn(nb, c));
    }
    v[e] = __float2half_rn(val);
  }
  *reinterpret_cast<uint4*>(dn + p * 64 + g * 8) = *reinterpret_cast<const uint4*>(v);
}

// ------------------------------------------------------------------------------------------
// K1: e1 = relu(conv1x1(mean over slots of corr [slots][33][px])) -> [px][64] fp16
// ------------------------------------------------------------------------------------------
constexpr int kA1Pitch = kCorrK + 8;   // halfs
constexpr int kW1Pitch = kHid + 8;

__global__ void __launch_bounds__(256) corr_enc1_kernel(const float* __restrict__ corr, int slots,
                                                       const __half* __restrict__ w1, const float* __restrict__ b1,
                                                       __half* __restrict__ e1, long long px) {
  __shared__ __align__(16) __half sA[128 * kA1Pitch];
  __shared__ __align__(16) __half sW[kCorrK * kW1Pitch];
  const long long p0 = (long long)blockIdx.x * 128;
  const int tid = threadIdx.x;
  const float inv = 1.f / (float)slots;
  for (int i = tid; i < kCorrK * 128; i += 256) {
    const int k = i >> 7, pl = i & 127;
    float v = 0.f;
    if (k < kCorrPlanes && p0 + pl < px) {
      for (int s = 0; s < slots; ++s) v += __ldg(corr + ((long long)s * kCorrPlanes + k) * px + p0 + pl);
      v *= inv;
    }
    sA[pl * kA1Pitch + k] = __float2half_rn(v);
  }
  for (int i = tid; i < kCorrK * kHid / 8; i += 256) {
    const int k = i / (kHid / 8), c = i % (kHid / 8);
    *reinterpret_cast<uint4*>(sW + k * kW1Pitch + c * 8) = __ldg(reinterpret_cast<const uint4*>(w1 + k * kHid) + c);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  const uint32_t aBase = smem_u32(sA) + ((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kA1Pitch + 8 * (lane >> 4)) * 2;
  const uint32_t bBase = smem_u32(sW) + (((lane & 7) + 8 * ((lane >> 3) & 1)) * kW1Pitch + 8 * (lane >> 4)) * 2;
#pragma unroll
  for (int k16 = 0; k16 < kCorrK / 16; ++k16) {
    uint32_t a[4];
    ldmatrix_x4(a, aBase + k16 * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, bBase + (k16 * 16 * kW1Pitch + j * 16) * 2);
      mma16816(acc[2 * j], a, b[0], b[1]);
      mma16816(acc[2 * j + 1], a, b[2], b[3]);
    }
  }
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const long long p = p0 + warp * 16 + g + 8 * half;
    if (p >= px) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = j * 8 + q * 2;
      const float v0 = fmaxf(h_round(acc[j][2 * half] + __ldg(b1 + n)), 0.f);
      const float v1 = fmaxf(h_round(acc[j][2 * half + 1] + __ldg(b1 + n + 1)), 0.f);
      *reinterpret_cast<__half2*>(e1 + p * 64 + n) = __floats2half2_rn(v0, v1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// KA (plan only): [disp += delta of the previous iteration] + pyramid lookup + e1 = relu(conv1x1(corr)).
// Fuses K6 of iteration i-1, CorrBlock.__call__ and K1 of iteration i: the 33 correlation planes never
// touch HBM.  Block = 128 pixels; the level-0 volume rows are staged in shared memory (coalesced), every
// thread pair produces the 33 taps of one pixel as the fp16 A row of the 1x1-conv MMA.
// ------------------------------------------------------------------------------------------
constexpr int kLE_PIX = 64;     // pixels per block (4 threads per pixel for the lookup, 4 MMA warps)

__global__ void __launch_bounds__(256) lookup_enc1_kernel(
    const float* __restrict__ volume, const float* __restrict__ origin, float* __restrict__ disp,
    const float* __restrict__ s9, int parts, const float* __restrict__ bd1, int apply_prev, int D, float incre,
    const __half* __restrict__ w1, const float* __restrict__ b1, __half* __restrict__ e1, int h, int w) {
  extern __shared__ __align__(16) unsigned char fsm[];
  __half* sA = reinterpret_cast<__half*>(fsm);                          // [kLE_PIX][kA1Pitch]
  __half* sW = sA + kLE_PIX * kA1Pitch;                                 // [48][kW1Pitch]
  float* sC = reinterpret_cast<float*>(sW + kCorrK * kW1Pitch);         // [kLE_PIX] lookup coordinate
  float* rows = sC + kLE_PIX;                                           // [kLE_PIX][D + 1]
  const long long px = (long long)h * w;
  const long long p0 = (long long)blockIdx.x * kLE_PIX;
  const int npix = (int)min((long long)kLE_PIX, px - p0);
  const int tid = threadIdx.x;
  const int pitch = D + 1;
  pdl_trigger();
  for (int i = tid; i < kCorrK * kHid / 8; i += 256) {   // 1x1 weights: constants, fetched before the dependency wait
    const int k = i / (kHid / 8), c = i % (kHid / 8);
    *reinterpret_cast<uint4*>(sW + k * kW1Pitch + c * 8) = __ldg(reinterpret_cast<const uint4*>(w1 + k * kHid) + c);
  }
  pdl_wait();     // volume (build kernel), disp / s9 (previous iteration) are produced upstream; e1 is read upstream
  // stage volume rows
  const float* vsrc = volume + p0 * D;
  if ((D & 3) == 0) {
    const int nvec = npix * D / 4;
    for (int i = tid; i < nvec; i += 256) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(vsrc) + i);
      const int e = i * 4;
      float* d = rows + (e / D) * pitch + (e % D);
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  } else {
    for (int i = tid; i < npix * D; i += 256) rows[(i / D) * pitch + (i % D)] = __ldg(vsrc + i);
  }
  // disparity of this block's pixels (after applying the pending delta) -> lookup coordinate
  if (tid < kLE_PIX) {
    float c = 0.f;
    if (tid < npix) {
      const long long p = p0 + tid;
      float dsp = disp[p];
      if (apply_prev) {     // K6 of the previous iteration: delta = fp16(0.01 * fp16(b + sum_t s9[p + off_t][t]))
        const int x = (int)(p % w), y = (int)(p / w);
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
          if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
            const float* q = s9 + s9_index(px, 0, t, (long long)yy * w + xx);
            s += (parts == 2) ? __ldg(q) + __ldg(q + 9 * (long long)px) : __ldg(q);
          }
        }
        dsp += h_round(0.01f * h_round(s + __ldg(bd1)));
        disp[p] = dsp;
      }
      c = lookup_coord(dsp, __ldg(origin + p), incre, D);
    }
    sC[tid] = c;
  }
  __syncthreads();
  {  // lookup: 4 threads per pixel, 12 of the 48 A columns each (33 real taps, the rest zero); the tap loops are
     // fully unrolled so level and offset are compile-time and the independent taps overlap
    const int pl = tid & (kLE_PIX - 1), quarter = tid / kLE_PIX;
    const float* row = rows + pl * pitch;
    const float c = sC[pl];
    const bool live = pl < npix;
    __half* arow = sA + pl * kA1Pitch;
#define CER_TAPS(K0)                                                                                        \
  _Pragma("unroll") for (int k = (K0); k < (K0) + 12; ++k)                                                  \
      arow[k] = __float2half_rn((live && k < kCorrPlanes) ? lookup_tap(row, D, k / 11, k % 11 - 5, c) : 0.f);
    if (quarter == 0) { CER_TAPS(0) } else if (quarter == 1) { CER_TAPS(12) } else if (quarter == 2) { CER_TAPS(24) } else { CER_TAPS(36) }
#undef CER_TAPS
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp >= kLE_PIX / 16) return;          // 4 MMA warps x 16 pixels
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  const uint32_t aBase = smem_u32(sA) + ((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kA1Pitch + 8 * (lane >> 4)) * 2;
  const uint32_t bBase = smem_u32(sW) + (((lane & 7) + 8 * ((lane >> 3) & 1)) * kW1Pitch + 8 * (lane >> 4)) * 2;
#pragma unroll
  for (int k16 = 0; k16 < kCorrK / 16; ++k16) {
    uint32_t a[4];
    ldmatrix_x4(a, aBase + k16 * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, bBase + (k16 * 16 * kW1Pitch + j * 16) * 2);
      mma16816(acc[2 * j], a, b[0], b[1]);
      mma16816(acc[2 * j + 1], a, b[2], b[3]);
    }
  }
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const long long p = p0 + warp * 16 + g + 8 * half;
    if (p >= px) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = j * 8 + q * 2;
      const float v0 = fmaxf(h_round(acc[j][2 * half] + __ldg(b1 + n)), 0.f);
      const float v1 = fmaxf(h_round(acc[j][2 * half + 1] + __ldg(b1 + n + 1)), 0.f);
      *reinterpret_cast<__half2*>(e1 + p * 64 + n) = __floats2half2_rn(v0, v1);
    }
  }
}

static size_t lookup_enc1_smem(int D) {
  return (size_t)kLE_PIX * kA1Pitch * 2 + (size_t)kCorrK * kW1Pitch * 2 + kLE_PIX * 4 + (size_t)kLE_PIX * (D + 1) * 4;
}

// ------------------------------------------------------------------------------------------
// KA v4 (plan, D = 64 / 44): the same fused step as one wave of warp-autonomous chunks.
//   * a warp owns 32 consecutive pixels end to end and the kernel has NO block barrier: 5 warps per CTA, 5 CTAs per SM
//     = 25 resident warps per SM = 3 700 chunks at cfg 2 = exactly one wave on 148 SMs (round 1: 1.25 waves of 4-warp CTAs).
//   * the level-0 volume rows go global -> shared memory asynchronously (D = 44: one bulk copy of the copy engine per
//     chunk; D = 64: cp.async into a padded pitch): no row registers, no store instructions.  Row pitch = 4 (mod 32)
//     words or 12, so a 16-byte load per lane at equal offsets is conflict-free.
//   * ONE shared window per pixel: the 12 aligned 16-byte pieces [f2-5, f2+6] around the level-2 centre hold every value
//     the 33 taps touch (the level-1 and level-0 windows are nested in it).  Pieces outside the row come from one zero
//     piece (= grid_sample's zero padding, exact because D % 4 == 0).  Pair sums S and quad sums Q of the pieces are the
//     pooled levels (same expressions as pyr_value up to the exact power-of-two scalings, folded into the weights).
//   * per level the centre coordinate runs through the reference's normalise / unnormalise round trip exactly
//     (bilinear_sampler.py:12 + grid_sample); the ten other taps of the level reuse its floor and weights.  The reference
//     recomputes the round trip per tap, which moves a tap position by <= 1 ulp of the coordinate (4e-6 at 64): the taps
//     differ from the general kernel by <= 5e-6 * |v1 - v0| before they are rounded to fp16 (autocast) -- 12 loads and
//     ~220 instructions per pixel instead of 154 loads and ~1 000.
//   * the 1x1 weights are read as ready-made mma.sync B fragments (BlobLayout::w1f, 6 KB, L1-resident: every other
//     global read of the kernel bypasses L1): no weight tile in shared memory.  The bias rides in the K padding (plane
//     33 = 1, fragment row 33 = bias), so the epilogue is round + max.  The 18 delta partials are 18 coalesced row reads (plane-major s9).
// ------------------------------------------------------------------------------------------
constexpr int kL4_WARPS = 5;
constexpr int kL4_OUT_PITCH = 144;                            // bytes per pixel of the e1 staging tile
constexpr int kL4_TILE_BYTES = 32 * kA1Pitch * 2 + 32 * kL4_OUT_PITCH;   // A tile + e1 staging tile (alias the rows)
template <int D>
struct L4 {
  static constexpr int NV = D / 4;                             // 16-byte pieces per row
  static constexpr int PITCH = (D % 32 == 0) ? D + 4 : D;      // floats
  static constexpr int ROW_BYTES = 32 * PITCH * 4;
  static constexpr int ZERO_OFF = ROW_BYTES > kL4_TILE_BYTES ? ROW_BYTES : kL4_TILE_BYTES;
  static constexpr int BAR_OFF = ZERO_OFF + 16;               // one mbarrier per warp: arrival of the rows
  static constexpr int WARP_BYTES = BAR_OFF + 16;
  static_assert(D % 4 == 0 && D >= 8 && D <= 64, "16-byte row pieces; one window covers a level-2 row of <= 16 values");
  static_assert(PITCH % 4 == 0 && (PITCH % 32) % 8 == 4, "pitch = 4, 12, 20 or 28 (mod 32) words");
};

__device__ __forceinline__ void lds128(float (&r)[48], int i, uint32_t addr) {
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(r[i]), "=f"(r[i + 1]), "=f"(r[i + 2]), "=f"(r[i + 3]) : "r"(addr));
}

// streaming read that leaves L1 to the 6 KB of weight fragments every warp of the SM re-reads
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// centre tap (offset 0) of one level: x0 = cl -> 2*x0/(W-1) - 1 -> ((xn + 1) / 2) * (W-1), rounded like the reference
__device__ __forceinline__ float lookup_centre(const LevelConst k, float cl) {
  const float q0 = __fmul_rn(cl, k.rcp);
  const float q = __fmaf_rn(__fmaf_rn(-q0, k.wm1, cl), k.rcp, q0);
  const float xn = __fmaf_rn(2.f, q, -1.f);
  return __fmul_rn(__fadd_rn(xn, 1.f), k.hw);
}

template <int D>
__global__ void __launch_bounds__(kL4_WARPS * 32, 5) lookup_enc1_v4_kernel(
    const float* __restrict__ volume, const float* __restrict__ origin, float* __restrict__ disp,
    const float* __restrict__ s9, int parts, const float* __restrict__ bd1, int apply_prev, float incre,
    const uint4* __restrict__ w1f, __half* __restrict__ e1, int h, int w) {
  using C = L4<D>;
  constexpr int NV = C::NV, P = C::PITCH;
  extern __shared__ __align__(16) unsigned char fsm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* wbase = fsm + warp * C::WARP_BYTES;
  const int px = h * w;
  const int p0 = (blockIdx.x * kL4_WARPS + warp) * 32;
  pdl_trigger();
  if (p0 >= px) return;                 // no block barrier in this kernel: a surplus warp of the last CTA just leaves
  const int npix = min(32, px - p0);
  const bool live = lane < npix;
  const int p = p0 + (live ? lane : 0);
  const uint32_t rows_s = smem_u32(wbase), zero_s = rows_s + C::ZERO_OFF;
  const uint32_t bar = rows_s + C::BAR_OFF;
  if (lane == 0) {
    *reinterpret_cast<float4*>(wbase + C::ZERO_OFF) = make_float4(0.f, 0.f, 0.f, 0.f);
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  pdl_wait();     // volume (build kernel), disp / s9 (previous iteration) are produced upstream; e1 is read upstream

  // ---- the chunk's rows: 32 * D contiguous floats ----
  // pitch == row (D = 44): ONE bulk copy of the copy engine per chunk.  Padded pitch (D = 64): cp.async, piece
  // f = it * 32 + lane (coalesced) -> row f / NV.  (One bulk copy per row was measured: the copy instruction takes
  // uniform registers, so 32 rows become a 32-trip lane loop -- 28 % of the kernel's stall samples.)
  if (P == D) {
    if (lane == 0) {
      mbar_expect_tx(bar, (uint32_t)npix * D * 4);
      bulk_g2s(rows_s, volume + (long long)p0 * D, (uint32_t)npix * D * 4, bar);
    }
  } else {
    const float4* vsrc = reinterpret_cast<const float4*>(volume + (long long)p0 * D);
    const int nvec = npix * NV;
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int f = it * 32 + lane;
      if (f < nvec) cp_async16(rows_s + ((f / NV) * P + (f % NV) * 4) * 4, vsrc + f, true);
    }
    cp_async_commit();
  }
  float dsp = ldg_stream(disp + p);     // (written below by this thread only)
  const float org = ldg_stream(origin + p);
  if (apply_prev) {     // K6 of the previous iteration: delta = fp16(0.01 * fp16(b + sum_t s9[p + off_t][t]))
    // all 18 partials are requested before the first use: the addresses of neighbours outside the image are clamped
    // to the pixel itself (a legal address) and the value is dropped, so no load sits behind a branch
    const int x = p % w, y = p / w;
    const bool oky[3] = {y > 0, true, y < h - 1}, okx[3] = {x > 0, true, x < w - 1};
    float sa[9], sb[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const bool ok = oky[t / 3] && okx[t % 3];
      const float* q = s9 + (long long)t * px + (ok ? p + (t / 3 - 1) * w + (t % 3 - 1) : p);     // s9_index(px, 0, t, .)
      sa[t] = ldg_stream(q);
      sb[t] = (parts == 2) ? ldg_stream(q + 9 * (long long)px) : 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float sv = (parts == 2) ? sa[t] + sb[t] : sa[t];
      if (oky[t / 3] && okx[t % 3]) s += sv;
    }
    dsp += h_round(0.01f * h_round(s + __ldg(bd1)));
    if (live) disp[p] = dsp;
  }
  // coordinates past the window are all "every tap outside": cap them so that the integer arithmetic below is safe.
  // A lane without a pixel (last chunk) takes the cap: all of its pieces are the zero piece, its A row is zero.
  const float c = live ? fminf(lookup_coord(dsp, org, incre, D), 4096.f) : 4096.f;

  constexpr LevelConst k0 = level_const(D, 0), k1 = level_const(D, 1), k2 = level_const(D, 2);
  const float xp0 = lookup_centre(k0, c), xp1 = lookup_centre(k1, c * 0.5f), xp2 = lookup_centre(k2, c * 0.25f);
  const int f2 = (int)floorf(xp2);
  // nest the level-1 / level-0 floors in the level-2 window; a floor that rounding put one step outside moves back and
  // the weights (computed from the moved floor) extrapolate by that ulp, which is the same value by continuity
  const int f1 = min(max((int)floorf(xp1), 2 * f2), 2 * f2 + 1);
  const int f0 = min(max((int)floorf(xp0), 4 * f2), 4 * f2 + 3);
  const float w21 = (xp2 - (float)f2) * 0.25f, w20 = (((float)f2 + 1.f) - xp2) * 0.25f;     // * 1/4: level-2 pooling
  const float w11 = (xp1 - (float)f1) * 0.5f, w10 = (((float)f1 + 1.f) - xp1) * 0.5f;       // * 1/2: level-1 pooling
  const float w01 = xp0 - (float)f0, w00 = ((float)f0 + 1.f) - xp0;
  const int o1 = f1 - 2 * f2, o0 = f0 - 4 * f2;

  if (P == D) {
    mbar_wait(bar, 0);
  } else {
    cp_async_wait<0>();
    __syncwarp();
  }

  // ---- 33 taps of this lane's pixel -> fp16 A row (registers) ----
  uint32_t arow[kCorrK / 2];
  {
    float R[48];
    const uint32_t row_s = rows_s + lane * (P * 4);
#pragma unroll
    for (int n = 0; n < 12; ++n) {
      const int pi = f2 - 5 + n;
      lds128(R, 4 * n, (unsigned)pi < (unsigned)NV ? row_s + pi * 16 : zero_s);
    }
    auto pack = [&](float lo, float hi) {
      const __half2 hh = __floats2half2_rn(lo, hi);
      return *reinterpret_cast<const uint32_t*>(&hh);
    };
#pragma unroll
    for (int k = 17; k < kCorrK / 2; ++k) arow[k] = 0u;        // planes 34..47: K padding
    float S[24];
#pragma unroll
    for (int m = 0; m < 24; ++m) S[m] = __fadd_rn(R[2 * m], R[2 * m + 1]);
    float t11;                                                  // plane 11 pairs with plane 10 of level 0
    {   // level 2 (planes 22..32): value k2 = f2 - 5 + n is quad sum n
      float Q[12], t[11];
#pragma unroll
      for (int n = 0; n < 12; ++n) Q[n] = __fadd_rn(S[2 * n], S[2 * n + 1]);
#pragma unroll
      for (int j = 0; j < 11; ++j) t[j] = __fmaf_rn(Q[j + 1], w21, __fmul_rn(Q[j], w20));
#pragma unroll
      for (int k = 0; k < 5; ++k) arow[11 + k] = pack(t[2 * k], t[2 * k + 1]);
      arow[16] = pack(t[10], 1.f);                              // plane 33 = 1: row 33 of the fragments is the bias
    }
    {   // level 1 (planes 11..21): value k1 = f1 - 5 + i is pair sum 5 + o1 + i
      float V[12], t[11];
#pragma unroll
      for (int i = 0; i < 12; ++i) V[i] = o1 ? S[6 + i] : S[5 + i];
#pragma unroll
      for (int j = 0; j < 11; ++j) t[j] = __fmaf_rn(V[j + 1], w11, __fmul_rn(V[j], w10));
      t11 = t[0];
#pragma unroll
      for (int k = 0; k < 5; ++k) arow[6 + k] = pack(t[1 + 2 * k], t[2 + 2 * k]);
    }
    {   // level 0 (planes 0..10): value k0 = f0 - 5 + i is window element 15 + o0 + i
      float T[14], t[11];
#pragma unroll
      for (int i = 0; i < 14; ++i) T[i] = (o0 & 1) ? R[16 + i] : R[15 + i];
#pragma unroll
      for (int j = 0; j < 11; ++j) {
        const float v0 = (o0 & 2) ? T[j + 2] : T[j], v1 = (o0 & 2) ? T[j + 3] : T[j + 1];
        t[j] = __fmaf_rn(v1, w01, __fmul_rn(v0, w00));
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) arow[k] = pack(t[2 * k], t[2 * k + 1]);
      arow[5] = pack(t[10], t11);
    }
  }
  __syncwarp();                       // every lane is done with its row: the A tile may overwrite the rows
  __half* sA = reinterpret_cast<__half*>(wbase);
#pragma unroll
  for (int k = 0; k < kCorrK / 8; ++k)
    *reinterpret_cast<uint4*>(sA + lane * kA1Pitch + k * 8) =
        make_uint4(arow[4 * k], arow[4 * k + 1], arow[4 * k + 2], arow[4 * k + 3]);
  __syncwarp();

  // ---- 1x1 conv on mma.sync (two 16-pixel tiles), bias, fp16 rounding, ReLU -> staging tile ----
  unsigned char* sO = wbase + 32 * kA1Pitch * 2;     // behind the A tile
  const int g = lane >> 2, q = lane & 3;
  const __half2 hzero = __floats2half2_rn(0.f, 0.f);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    const uint32_t aBase = smem_u32(sA) + ((mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kA1Pitch + 8 * (lane >> 4)) * 2;
#pragma unroll
    for (int kb = 0; kb < kCorrK / 16; ++kb) {
      uint32_t a[4];
      ldmatrix_x4(a, aBase + kb * 32);
      const uint4* wf = w1f + (kb * 32 + lane) * 4;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint4 b = __ldg(wf + jj);          // n-blocks 2jj, 2jj+1: (b0, b1) each
        mma16816(acc[2 * jj], a, b.x, b.y);
        mma16816(acc[2 * jj + 1], a, b.z, b.w);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)         // relu(fp16(acc)) == fp16 max(., 0) of the rounded sum; the bias came in through K
#pragma unroll
      for (int half = 0; half < 2; ++half)
        *reinterpret_cast<__half2*>(sO + (mt * 16 + g + 8 * half) * kL4_OUT_PITCH + (j * 8 + q * 2) * 2) =
            __hmax2(__floats2half2_rn(acc[j][2 * half], acc[j][2 * half + 1]), hzero);
  }
  __syncwarp();
  // ---- e1: 4 pixels (512 contiguous bytes) per warp store ----
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + (lane >> 3), cchunk = lane & 7;
    if (row < npix)
      *reinterpret_cast<uint4*>(e1 + (long long)(p0 + row) * 64 + cchunk * 8) =
          *reinterpret_cast<const uint4*>(sO + row * kL4_OUT_PITCH + cchunk * 16);
  }
}

template <int D>
static size_t lookup_enc1_v4_smem() { return (size_t)kL4_WARPS * L4<D>::WARP_BYTES; }


// ------------------------------------------------------------------------------------------
// 3x3 implicit-GEMM convolution
// ------------------------------------------------------------------------------------------
constexpr int TH = 8, TW = 16;                 // output tile (128 pixels)
constexpr int HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int HALO_PX = HALO_W * HALO_H;       // 180
constexpr int A_PITCH = 64 + 8;                // halfs per halo pixel (144 B, conflict-free ldmatrix)
constexpr int A_BYTES = HALO_PX * A_PITCH * 2; // 25920
constexpr int NSTAGE = 3;

template <int N_TILE>
struct ConvSmem {
  static constexpr int B_PITCH = N_TILE + 8;             // halfs
  static constexpr int B_BYTES = 64 * B_PITCH * 2;
  static constexpr int TOTAL = 2 * A_BYTES + NSTAGE * B_BYTES + (N_TILE == 256 ? (9 * 256 + 4 * 128 * 9) * 4 : 0);
};

template <int N_TILE, int EPI>
__global__ void __launch_bounds__(256, 1) conv3x3_hmma_kernel(const ConvArgs a) {
  using S = ConvSmem<N_TILE>;
  constexpr int NW = N_TILE / 4;     // columns per warp
  constexpr int NJ = NW / 8;         // n8 tiles per warp
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + 2 * A_BYTES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3;
  const int tiles_x = (a.w + TW - 1) / TW;
  const int x0 = (blockIdx.x % tiles_x) * TW, y0 = (blockIdx.x / tiles_x) * TH;
  const int n_steps = a.n_src * 9;

  auto load_A = [&](int chunk) {
    const __half* src = a.src[chunk];
    const uint32_t dst0 = sA + (chunk & 1) * A_BYTES;
    for (int i = tid; i < HALO_PX * 8; i += 256) {
      const int hp = i >> 3, c = i & 7;
      const int yy = y0 - 1 + hp / HALO_W, xx = x0 - 1 + hp % HALO_W;
      const bool ok = yy >= 0 && yy < a.h && xx >= 0 && xx < a.w;
      const __half* g = src + ((long long)(ok ? yy : 0) * a.w + (ok ? xx : 0)) * 64 + c * 8;
      cp_async16(dst0 + hp * (A_PITCH * 2) + c * 16, g, ok);
    }
  };
  auto load_B = [&](int step) {
    const __half* src = a.wpk + (long long)step * 64 * N_TILE;
    const uint32_t dst0 = sB + (step % NSTAGE) * S::B_BYTES;
    constexpr int CPR = N_TILE / 8;  // 16-byte chunks per row
    for (int i = tid; i < 64 * CPR; i += 256) {
      const int k = i / CPR, c = i % CPR;
      cp_async16(dst0 + k * (S::B_PITCH * 2) + c * 16, src + k * N_TILE + c * 8, true);
    }
  };

  float acc[4][NJ][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  // prologue: steps 0 and 1
  load_A(0);
  load_B(0);
  cp_async_commit();
  if (n_steps > 1) load_B(1);
  cp_async_commit();

  const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1);   // ldmatrix row supplied by this lane
  const int lcol = 8 * (lane >> 4);                      // ldmatrix column offset (halfs)

  for (int step = 0; step < n_steps; ++step) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    {  // prefetch step + 2 (its B buffer was consumed in step - 1)
      const int nxt = step + NSTAGE - 1;
      if (nxt < n_steps) {
        if (nxt % 9 == 0) load_A(nxt / 9);
        load_B(nxt);
      }
      cp_async_commit();
    }
    const int chunk = step / 9, tap = step % 9;
    const int ky = tap / 3, kx = tap % 3;
    const uint32_t aBuf = sA + (chunk & 1) * A_BYTES;
    const uint32_t bBuf = sB + (step % NSTAGE) * S::B_BYTES;
#pragma unroll
    for (int k16 = 0; k16 < 4; ++k16) {
      uint32_t af[4][4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int hp = (wm * 4 + mi + ky) * HALO_W + kx + lrow;
        ldmatrix_x4(af[mi], aBuf + hp * (A_PITCH * 2) + (k16 * 16 + lcol) * 2);
      }
#pragma unroll
      for (int jp = 0; jp < NJ / 2; ++jp) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, bBuf + (k16 * 16 + lrow) * (S::B_PITCH * 2) + (wn * NW + jp * 16 + lcol) * 2);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          mma16816(acc[mi][2 * jp], af[mi], bf[0], bf[1]);
          mma16816(acc[mi][2 * jp + 1], af[mi], bf[2], bf[3]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---------------- epilogues ----------------
  const int g = lane >> 2, q = lane & 3;
  if (EPI == EPI_DELTA) {
    // d = relu(fp16(acc + b)); s9[p][t] = sum_n w2[t][n] * d[n]  (second delta conv as 9 per-pixel dots)
    float* sW2 = reinterpret_cast<float*>(smem + 2 * A_BYTES + NSTAGE * S::B_BYTES);
    float* sS9 = sW2 + 9 * 256;          // [column quarter wn][pixel 128][tap 9]: summed in a fixed order below
    __syncthreads();
    for (int i = tid; i < 9 * 256; i += 256) sW2[i] = __ldg(a.w2 + i);
    __syncthreads();
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      float t0[9], t1[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) t0[t] = t1[t] = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int n = wn * NW + j * 8 + q * 2;
        const float bn0 = __ldg(a.bias + n), bn1 = __ldg(a.bias + n + 1);
        const float d00 = fmaxf(h_round(acc[mi][j][0] + bn0), 0.f), d01 = fmaxf(h_round(acc[mi][j][1] + bn1), 0.f);
        const float d10 = fmaxf(h_round(acc[mi][j][2] + bn0), 0.f), d11 = fmaxf(h_round(acc[mi][j][3] + bn1), 0.f);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float2 wv = *reinterpret_cast<const float2*>(sW2 + t * 256 + n);
          t0[t] = fmaf(d00, wv.x, fmaf(d01, wv.y, t0[t]));
          t1[t] = fmaf(d10, wv.x, fmaf(d11, wv.y, t1[t]));
        }
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        t0[t] += __shfl_xor_sync(0xffffffffu, t0[t], 1);
        t0[t] += __shfl_xor_sync(0xffffffffu, t0[t], 2);
        t1[t] += __shfl_xor_sync(0xffffffffu, t1[t], 1);
        t1[t] += __shfl_xor_sync(0xffffffffu, t1[t], 2);
      }
      if (q == 0) {
        const int pl = (wm * 4 + mi) * TW + g;
        // every (pixel, tap) of a column quarter is written by exactly one lane: no atomics, so the sum over the four
        // quarters below has a fixed order (shared-memory float atomics made the result depend on warp timing)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          sS9[(wn * 128 + pl) * 9 + t] = t0[t];
          sS9[(wn * 128 + pl + 8) * 9 + t] = t1[t];
        }
      }
    }
    __syncthreads();
    for (int k = tid; k < 128 * 9; k += 256) {
      const int t = k / 128, pl = k % 128, i = pl * 9 + t;     // plane-major: consecutive threads = consecutive pixels of a tap
      const int yy = y0 + pl / TW, xx = x0 + pl % TW;
      if (yy < a.h && xx < a.w)
        a.s9[s9_index((long long)a.h * a.w, 0, t, (long long)yy * a.w + xx)] =
            ((sS9[i] + sS9[128 * 9 + i]) + sS9[2 * 128 * 9 + i]) + sS9[3 * 128 * 9 + i];
    }
    return;
  }

#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int yy = y0 + wm * 4 + mi;
    if (yy >= a.h) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int xx = x0 + g + 8 * half;
      if (xx >= a.w) continue;
      const long long p = (long long)yy * a.w + xx;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int n = wn * NW + j * 8 + q * 2;
        float v0 = acc[mi][j][2 * half], v1 = acc[mi][j][2 * half + 1];
        if (EPI == EPI_RELU) {
          v0 = fmaxf(h_round(v0 + __ldg(a.bias + n)), 0.f);
          v1 = fmaxf(h_round(v1 + __ldg(a.bias + n + 1)), 0.f);
          *reinterpret_cast<__half2*>(a.out_h + p * 64 + n) = __floats2half2_rn(v0, v1);
        } else if (EPI == EPI_GATES) {
          v0 += __ldg(a.bias + n);
          v1 += __ldg(a.bias + n + 1);
          if (n < 64) {            // z = sigmoid(convz)                                update.py:20-21
            const float z0 = sigmoid_f(h_round(v0)), z1 = sigmoid_f(h_round(v1));
            *reinterpret_cast<__half2*>(a.z + p * 64 + n) = __floats2half2_rn(z0, z1);
          } else if (n < 128) {    // r = sigmoid(convr); r * net                         update.py:22-23
            const float r0 = h_round(sigmoid_f(h_round(v0))), r1 = h_round(sigmoid_f(h_round(v1)));
            const float2 nt = __half22float2(*reinterpret_cast<const __half2*>(a.net + p * 64 + (n - 64)));
            *reinterpret_cast<__half2*>(a.rnet + p * 64 + (n - 64)) = __floats2half2_rn(r0 * nt.x, r1 * nt.y);
          } else {                 // x-part of convq, kept in fp32 until K4 adds the r*net part
            *reinterpret_cast<float2*>(a.qx + p * 64 + (n - 128)) = make_float2(v0, v1);
          }
        } else if (EPI == EPI_GRUOUT) {  // q = tanh(convq); net = (1-z)*net + z*q          update.py:23-24
          const float2 qx = *reinterpret_cast<const float2*>(a.qx + p * 64 + n);
          const float2 zz = __half22float2(*reinterpret_cast<const __half2*>(a.z + p * 64 + n));
          const float2 nt = __half22float2(*reinterpret_cast<const __half2*>(a.net + p * 64 + n));
          const float q0 = h_round(tanhf(h_round(v0 + qx.x))), q1 = h_round(tanhf(h_round(v1 + qx.y)));
          const float n0 = h_round(h_round(h_round(1.f - zz.x) * nt.x) + h_round(zz.x * q0));
          const float n1 = h_round(h_round(h_round(1.f - zz.y) * nt.y) + h_round(zz.y * q1));
          *reinterpret_cast<__half2*>(a.net + p * 64 + n) = __floats2half2_rn(n0, n1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K6: delta = fp16(0.01 * fp16(b + sum_t s9[p + off_t][t])); disp += delta
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) disp_update_kernel(const float* __restrict__ s9, int parts,
                                                         const float* __restrict__ bd1, float* __restrict__ disp,
                                                         float* __restrict__ delta, int apply, int h, int w) {
  const long long px = (long long)h * w;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (p >= px) return;
  const int x = (int)(p % w), y = (int)(p / w);
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
      const float* q = s9 + s9_index(px, 0, t, (long long)yy * w + xx);
      s += (parts == 2) ? __ldg(q) + __ldg(q + 9 * px) : __ldg(q);
    }
  }
  const float d = h_round(0.01f * h_round(s + __ldg(bd1)));
  if (delta) delta[p] = d;
  if (apply) disp[p] += d;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int N_TILE, int EPI>
static int configure_conv() {
  CER_CUDA(cudaFuncSetAttribute(conv3x3_hmma_kernel<N_TILE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ConvSmem<N_TILE>::TOTAL));
  return CER_OK;
}

static int g_variant = -1;
int conv_variant() {
  if (g_variant < 0) {
    const char* e = getenv("CER_CONV");
    g_variant = (e && !strcmp(e, "hmma")) ? 0 : 1;
  }
  return g_variant;
}

// fused lookup kernel of the plan: 2 = warp-autonomous v2 (default), 1 = v1 (CER_LOOKUP=v1 / cer_set_lookup_variant)
static int g_lookup_variant = -1;
int lookup_variant() {
  if (g_lookup_variant < 0) {
    const char* e = getenv("CER_LOOKUP");
    g_lookup_variant = (e && !strcmp(e, "general")) ? 1 : 2;
  }
  return g_lookup_variant;
}
void set_lookup_variant(int v) { g_lookup_variant = v; }

// Opt in to >48 KB dynamic shared memory once per process (not capturable, so done up front).
int update_configure() {
  static std::atomic<unsigned long long> done{0};
  if (!first_time_on_device(done)) return CER_OK;
  int rc;
  if ((rc = tc_configure())) return rc;
  if ((rc = configure_conv<64, EPI_RELU>())) return rc;
  if ((rc = configure_conv<192, EPI_GATES>())) return rc;
  if ((rc = configure_conv<64, EPI_GRUOUT>())) return rc;
  if ((rc = configure_conv<256, EPI_DELTA>())) return rc;
  CER_CUDA(cudaFuncSetAttribute(lookup_enc1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)lookup_enc1_smem(256)));
  CER_CUDA(cudaFuncSetAttribute(lookup_enc1_v4_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)lookup_enc1_v4_smem<64>()));
  CER_CUDA(cudaFuncSetAttribute(lookup_enc1_v4_kernel<44>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)lookup_enc1_v4_smem<44>()));
  CER_CUDA(cudaFuncSetAttribute(lookup_enc1_v4_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CER_CUDA(cudaFuncSetAttribute(lookup_enc1_v4_kernel<44>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  return CER_OK;
}

static unsigned long long* g_conv_prof = nullptr;   // device buffer [4 kernels][32] set by cer_debug_set_conv_profile

template <int N_TILE, int EPI>
static int launch_conv(const ConvArgs& a_in, cudaStream_t stream) {
  if (conv_variant() == 1) {
    ConvArgs a2 = a_in;
    a2.prof = g_conv_prof ? g_conv_prof + EPI * 32 : nullptr;
    return launch_conv_tc_dispatch(N_TILE, EPI, a2, stream);
  }
  const ConvArgs& a = a_in;
  constexpr int smem = ConvSmem<N_TILE>::TOTAL;
  const int tiles = ((a.w + TW - 1) / TW) * ((a.h + TH - 1) / TH);
  constexpr int kind = EPI == EPI_RELU ? KK_CONV_E : EPI == EPI_GATES ? KK_CONV_GATES : EPI == EPI_GRUOUT ? KK_CONV_Q : KK_CONV_DELTA;
  CER_LAUNCH(kind, (conv3x3_hmma_kernel<N_TILE, EPI>), tiles, 256, smem, stream, a);
  return check_launch("conv3x3_hmma");
}

int update_step_hmma(const void* blob, void* workspace, void* net, const void* inp, float* disp, const float* corr,
                     int slots, float* delta, int apply_delta, int stage, int h, int w, cudaStream_t stream) {
  const BlobLayout L = blob_layout();
  const char* B = (const char*)blob;
  const long long px = (long long)h * w;
  UpdateWs ws = carve_ws(workspace, h, w);
  int rc;
  if ((rc = update_configure())) return rc;
  const bool tc = conv_variant() == 1;
  if (!tc) CER_LAUNCH(KK_DISP_ENC, disp_encode_kernel, ceil_div(px * 8, 256), 256, 0, stream, disp, ws.dn, h, w);
  CER_LAUNCH(KK_CORR_ENC1, corr_enc1_kernel, ceil_div(px, 128), 256, 0, stream, corr, slots, (const __half*)(B + L.w1),
             (const float*)(B + L.b1), ws.e1, px);
  if ((rc = check_launch("update prologue"))) return rc;
  ConvArgs a{};
  a.h = h;
  a.w = w;
  a.dn_chunk = -1;
  // K2
  a.src[0] = ws.e1; a.n_src = 1; a.wpk = (const __half*)(B + L.w2); a.wtc = (const __half*)(B + L.t_w2); a.bias = (const float*)(B + L.b2); a.out_h = ws.e;
  if ((rc = launch_conv<64, EPI_RELU>(a, stream))) return rc;
  // K3
  a.src[0] = (const __half*)net; a.src[1] = (const __half*)inp; a.src[2] = ws.dn; a.src[3] = ws.e; a.n_src = 4;
  a.wpk = (const __half*)(B + L.wg); a.wtc = (const __half*)(B + L.t_wg); a.wtc2 = (const __half*)(B + L.p_wg); a.bias = (const float*)(B + L.bg);
  a.net = (__half*)net; a.z = ws.z; a.rnet = ws.rnet; a.qx = ws.qx;
  if (tc) { a.disp = disp; a.dn_chunk = 2; }      // disparity encoder generated inside the gate conv
  if ((rc = launch_conv<192, EPI_GATES>(a, stream))) return rc;
  a.dn_chunk = -1;
  // K4
  a.src[0] = ws.rnet; a.n_src = 1; a.wpk = (const __half*)(B + L.wq); a.wtc = (const __half*)(B + L.t_wq); a.bias = nullptr;
  if ((rc = launch_conv<64, EPI_GRUOUT>(a, stream))) return rc;
  // K5
  a.src[0] = (const __half*)net; a.n_src = 1; a.wpk = (const __half*)(B + L.wd0[stage]); a.wtc = (const __half*)(B + L.t_wd0[stage]); a.wtc2 = (const __half*)(B + L.p_wd0[stage]);
  a.bias = (const float*)(B + L.bd0[stage]); a.w2 = (const float*)(B + L.wd1[stage]); a.s9 = ws.s9;
  if ((rc = launch_conv<256, EPI_DELTA>(a, stream))) return rc;
  // K6
  CER_LAUNCH(KK_DISP_UPDATE, disp_update_kernel, ceil_div(px, 256), 256, 0, stream, ws.s9, tc ? 2 : 1,
             (const float*)(B + L.bd1[stage]), disp, delta, apply_delta, h, w);
  return check_launch("disp_update");
}

// One GRU iteration of the plan (core/raft.py:96-101): KA (apply pending delta, lookup, 1x1) + K2..K5.
// The delta of THIS iteration stays pending in ws.s9 (applied by the next KA or by update_apply_delta).
// Tile-level dependencies between the tcgen05 convs of an iteration (ConvArgs::flags_in / flags_out) were measured in
// round 1 (100.8 depth-maps/s with flags against 101.3 without) and are switched off: every kernel waits for its whole
// predecessor through programmatic dependent launch.
static int tile_flags() { return 0; }

// Zero the tile flags of the next `iters` iterations (start of a stage; a memset node when captured).
int update_reset_flags(void* workspace, int iters, int h, int w, cudaStream_t stream) {
  if (conv_variant() != 1 || !tile_flags()) return CER_OK;
  UpdateWs ws = carve_ws(workspace, h, w);
  const size_t n = (size_t)(iters < kFlagIters ? iters : kFlagIters) * kFlagKernels * flag_tiles(h, w) * sizeof(int);
  CER_CUDA(cudaMemsetAsync(ws.flags, 0, n, stream));
  return CER_OK;
}

// KA: [pending delta] + pyramid lookup + 1x1 corr encoder.  The two cascade widths of the reference (core/raft.py:77-81)
// take the warp-autonomous kernel; any other D (or CER_LOOKUP=general) the general one.
static void launch_lookup_enc1(const void* blob, const float* volume, const float* origin, float* disp, const float* s9,
                               int parts, const float* bd1, int apply_prev, int D, float incre, __half* e1, int h, int w,
                               cudaStream_t stream) {
  const BlobLayout L = blob_layout();
  const char* B = (const char*)blob;
  const long long px = (long long)h * w;
  if (lookup_variant() == 2 && (D == 64 || D == 44)) {
    const int grid = ceil_div(px, kL4_WARPS * 32);
    if (D == 64)
      CER_LAUNCH_PDL(KK_LOOKUP, lookup_enc1_v4_kernel<64>, grid, kL4_WARPS * 32, lookup_enc1_v4_smem<64>(), stream, volume, origin,
                     disp, s9, parts, bd1, apply_prev, incre, (const uint4*)(B + L.w1f), e1, h, w);
    else
      CER_LAUNCH_PDL(KK_LOOKUP, lookup_enc1_v4_kernel<44>, grid, kL4_WARPS * 32, lookup_enc1_v4_smem<44>(), stream, volume, origin,
                     disp, s9, parts, bd1, apply_prev, incre, (const uint4*)(B + L.w1f), e1, h, w);
  } else {
    CER_LAUNCH_PDL(KK_LOOKUP, lookup_enc1_kernel, ceil_div(px, kLE_PIX), 256, lookup_enc1_smem(D), stream, volume, origin, disp,
                   s9, parts, bd1, apply_prev, D, incre, (const __half*)(B + L.w1), (const float*)(B + L.b1), e1, h, w);
  }
}

int update_iteration_fused(const void* blob, void* workspace, void* net, const void* inp, float* disp,
                           const float* volume, const float* origin, int D, float incre, int apply_prev, int iter,
                           int stage, int h, int w, cudaStream_t stream) {
  const BlobLayout L = blob_layout();
  const char* B = (const char*)blob;
  const long long px = (long long)h * w;
  UpdateWs ws = carve_ws(workspace, h, w);
  int rc;
  if ((rc = update_configure())) return rc;
  if (D > 256) {
    set_error("update_iteration_fused: D > 256 unsupported");
    return CER_ERR_INVALID;
  }
  const bool tc = conv_variant() == 1;
  // tile flags of this iteration: [0] corr-encoder 3x3 -> gates, [1] gates -> q/GRU, [2] q/GRU -> delta
  const bool flags_on = tc && tile_flags() && iter >= 0 && iter < kFlagIters;
  const int n_flag_tiles = flag_tiles(h, w);
  auto F = [&](int k) { return flags_on ? ws.flags + ((size_t)iter * kFlagKernels + k) * n_flag_tiles : (int*)nullptr; };
  launch_lookup_enc1(blob, volume, origin, disp, ws.s9, tc ? 2 : 1, (const float*)(B + L.bd1[stage]), apply_prev, D, incre,
                     ws.e1, h, w, stream);
  if (!tc) CER_LAUNCH(KK_DISP_ENC, disp_encode_kernel, ceil_div(px * 8, 256), 256, 0, stream, disp, ws.dn, h, w);
  if ((rc = check_launch("lookup_enc1"))) return rc;
  ConvArgs a{};
  a.h = h;
  a.w = w;
  a.dn_chunk = -1;
  a.src[0] = ws.e1; a.n_src = 1; a.wpk = (const __half*)(B + L.w2); a.wtc = (const __half*)(B + L.t_w2);
  a.bias = (const float*)(B + L.b2); a.out_h = ws.e;
  a.flags_in = nullptr; a.flags_out = F(0);
  if ((rc = launch_conv<64, EPI_RELU>(a, stream))) return rc;
  a.src[0] = (const __half*)net; a.src[1] = (const __half*)inp; a.src[2] = ws.dn; a.src[3] = ws.e; a.n_src = 4;
  a.wpk = (const __half*)(B + L.wg); a.wtc = (const __half*)(B + L.t_wg); a.wtc2 = (const __half*)(B + L.p_wg); a.bias = (const float*)(B + L.bg);
  a.net = (__half*)net; a.z = ws.z; a.rnet = ws.rnet; a.qx = ws.qx;
  if (tc) { a.disp = disp; a.dn_chunk = 2; }
  a.flags_in = F(0); a.flags_out = F(1);
  if ((rc = launch_conv<192, EPI_GATES>(a, stream))) return rc;
  a.dn_chunk = -1;
  a.src[0] = ws.rnet; a.n_src = 1; a.wpk = (const __half*)(B + L.wq); a.wtc = (const __half*)(B + L.t_wq); a.bias = nullptr;
  a.flags_in = F(1); a.flags_out = F(2);
  if ((rc = launch_conv<64, EPI_GRUOUT>(a, stream))) return rc;
  a.src[0] = (const __half*)net; a.n_src = 1; a.wpk = (const __half*)(B + L.wd0[stage]);
  a.wtc = (const __half*)(B + L.t_wd0[stage]); a.wtc2 = (const __half*)(B + L.p_wd0[stage]); a.bias = (const float*)(B + L.bd0[stage]);
  a.w2 = (const float*)(B + L.wd1[stage]); a.s9 = ws.s9;
  a.flags_in = F(2); a.flags_out = nullptr;
  return launch_conv<256, EPI_DELTA>(a, stream);
}

// K6 on its own: apply the pending delta (after the last iteration of a stage).
int update_apply_delta(const void* blob, void* workspace, float* disp, int stage, int h, int w, cudaStream_t stream) {
  const BlobLayout L = blob_layout();
  const char* B = (const char*)blob;
  const long long px = (long long)h * w;
  UpdateWs ws = carve_ws(workspace, h, w);
  CER_LAUNCH_PDL(KK_DISP_UPDATE, disp_update_kernel, ceil_div(px, 256), 256, 0, stream, ws.s9, conv_variant() == 1 ? 2 : 1,
             (const float*)(B + L.bd1[stage]), disp, (float*)nullptr, 1, h, w);
  return check_launch("disp_update");
}

}  // namespace cer

using namespace cer;

extern "C" {

size_t cer_update_blob_bytes(void) { return blob_layout().total; }

size_t cer_update_workspace_bytes(int h, int w) { return carve_ws(nullptr, h, w).total; }

int cer_pack_update_weights(const float* const* w, void* blob_host) {
  CER_REQUIRE(w && blob_host, "cer_pack_update_weights: null pointer");
  for (int i = 0; i < 18; ++i) CER_REQUIRE(w[i], "cer_pack_update_weights: tensor %d is null", i);
  const BlobLayout L = blob_layout();
  char* B = (char*)blob_host;
  memset(B, 0, L.total);
  auto H = [](float v) { return __float2half_rn(v); };
  // corr_encoder.0: [64][33][1][1] -> w1[k][n]
  {
    __half* d = (__half*)(B + L.w1);
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < kCorrPlanes; ++k) d[k * 64 + n] = H(w[0][n * kCorrPlanes + k]);
    memcpy(B + L.b1, w[1], 64 * 4);
    // the same matrix as mma.sync m16n8k16 B fragments: [k16 block][lane][n8 block][2] words, word r of lane (g, q) =
    // (w1[kb*16 + 8r + 2q][nb*8 + g], w1[kb*16 + 8r + 2q + 1][nb*8 + g]), with the fp16 bias as row 33
    uint32_t* f = (uint32_t*)(B + L.w1f);
    for (int kb = 0; kb < kCorrK / 16; ++kb)
      for (int lane = 0; lane < 32; ++lane)
        for (int nb = 0; nb < 8; ++nb)
          for (int r = 0; r < 2; ++r) {
            const int k = kb * 16 + 8 * r + 2 * (lane & 3), n = nb * 8 + (lane >> 2);
            const __half hb = H(w[1][n]);                     // row 33 = bias (the kernel feeds a constant 1 as plane 33)
            const __half hlo = k == kCorrPlanes ? hb : d[k * 64 + n], hhi = k + 1 == kCorrPlanes ? hb : d[(k + 1) * 64 + n];
            unsigned short lo, hi;
            memcpy(&lo, &hlo, 2);
            memcpy(&hi, &hhi, 2);
            f[((kb * 32 + lane) * 8 + nb) * 2 + r] = (uint32_t)lo | ((uint32_t)hi << 16);
          }
  }
  // generic 3x3 OIHW [cout][cin][3][3] slice -> [tap][k][n_total] at column offset n0
  auto pack3x3 = [&](const float* src, int cout, int cin_total, int cin0, int cin_n, __half* dst, int n_total, int n0) {
    for (int o = 0; o < cout; ++o)
      for (int k = 0; k < cin_n; ++k)
        for (int t = 0; t < 9; ++t)
          dst[((long long)t * 64 + k) * n_total + n0 + o] = H(src[((long long)o * cin_total + cin0 + k) * 9 + t]);
  };
  pack3x3(w[2], 64, 64, 0, 64, (__half*)(B + L.w2), 64, 0);
  memcpy(B + L.b2, w[3], 64 * 4);
  // gates: chunk order net | inp | dn | e = input channels 0 | 64 | 128 (49) | 177
  const int cin0[4] = {0, 64, 128, 177};
  const int cinn[4] = {64, 64, kDispEnc, 64};
  for (int c = 0; c < 4; ++c) {
    __half* dst = (__half*)(B + L.wg) + (long long)c * 9 * 64 * kGateN;
    pack3x3(w[4], 64, kGruIn, cin0[c], cinn[c], dst, kGateN, 0);     // convz
    pack3x3(w[6], 64, kGruIn, cin0[c], cinn[c], dst, kGateN, 64);    // convr
    if (c > 0) pack3x3(w[8], 64, kGruIn, cin0[c], cinn[c], dst, kGateN, 128);  // convq, x part
  }
  {
    float* bg = (float*)(B + L.bg);
    memcpy(bg, w[5], 64 * 4);
    memcpy(bg + 64, w[7], 64 * 4);
    memcpy(bg + 128, w[9], 64 * 4);
  }
  pack3x3(w[8], 64, kGruIn, 0, 64, (__half*)(B + L.wq), 64, 0);      // convq, r*net part
  for (int s = 0; s < 2; ++s) {
    const float* w0 = w[10 + 4 * s];
    const float* b0 = w[11 + 4 * s];
    const float* w1 = w[12 + 4 * s];
    const float* b1 = w[13 + 4 * s];
    pack3x3(w0, 256, 64, 0, 64, (__half*)(B + L.wd0[s]), 256, 0);
    memcpy(B + L.bd0[s], b0, 256 * 4);
    float* d = (float*)(B + L.wd1[s]);
    for (int c = 0; c < 256; ++c)
      for (int t = 0; t < 9; ++t) d[t * 256 + c] = __half2float(H(w1[c * 9 + t]));
    *(float*)(B + L.bd1[s]) = b1[0];
  }
  // tcgen05 layout: re-tile every [tap][k][n] matrix into [tap][k/8][n][k%8]
  auto retile = [&](size_t src_off, size_t dst_off, int n_chunks, int N) {
    const __half* src = (const __half*)(B + src_off);
    __half* dst = (__half*)(B + dst_off);
    for (long long st = 0; st < (long long)n_chunks * 9; ++st)
      for (int k = 0; k < 64; ++k)
        for (int n = 0; n < N; ++n)
          dst[st * 64 * N + ((long long)(k / 8) * N + n) * 8 + (k % 8)] = src[st * 64 * N + (long long)k * N + n];
  };
  retile(L.w2, L.t_w2, 1, 64);
  retile(L.wg, L.t_wg, 4, kGateN);
  retile(L.wq, L.t_wq, 1, 64);
  for (int s = 0; s < 2; ++s) retile(L.wd0[s], L.t_wd0[s], 1, kDelta0);
  // CTA-pair layout: [tap][half][k/8][n % (N/2)][k%8]
  auto retile_pair = [&](size_t src_off, size_t dst_off, int n_chunks, int N) {
    const __half* src = (const __half*)(B + src_off);
    __half* dst = (__half*)(B + dst_off);
    const int NL = N / 2;
    for (long long st = 0; st < (long long)n_chunks * 9; ++st)
      for (int k = 0; k < 64; ++k)
        for (int n = 0; n < N; ++n)
          dst[st * 64 * N + (((long long)(n / NL) * 8 + k / 8) * NL + n % NL) * 8 + (k % 8)] =
              src[st * 64 * N + (long long)k * N + n];
  };
  retile_pair(L.wg, L.p_wg, 4, kGateN);
  for (int s = 0; s < 2; ++s) retile_pair(L.wd0[s], L.p_wd0[s], 1, kDelta0);
  // biases are rounded to fp16 by autocast as well (conv2d casts every floating argument)
  auto round_bias = [&](size_t off, int n) {
    float* b = (float*)(B + off);
    for (int i = 0; i < n; ++i) b[i] = __half2float(H(b[i]));
  };
  round_bias(L.b1, 64);
  round_bias(L.b2, 64);
  round_bias(L.bg, 192);
  for (int s = 0; s < 2; ++s) {
    round_bias(L.bd0[s], 256);
    round_bias(L.bd1[s], 1);
  }
  return CER_OK;
}

// Debug hook: device buffer of 4 x 32 uint64 receiving per-role wait/work cycle counters of CTA 0 of each tcgen05 conv
// kernel (epilogue kind index: 0 corr-encoder, 1 gates, 2 q/GRU, 3 delta); pass NULL to switch it off.
int cer_debug_set_conv_profile(void* dev_buf) {
  g_conv_prof = (unsigned long long*)dev_buf;
  return CER_OK;
}

int cer_set_conv_variant(int variant) {
  CER_REQUIRE(variant == 0 || variant == 1,
              "cer_set_conv_variant: 1 tcgen05.mma + TMEM (default), 0 mma.sync (the A/B twin every GPU test also runs on)");
  g_variant = variant;
  return CER_OK;
}

int cer_set_lookup_variant(int variant) {
  CER_REQUIRE(variant == 1 || variant == 2,
              "cer_set_lookup_variant: 2 warp-autonomous kernels for the reference configuration (default), 1 general kernels");
  set_lookup_variant(variant);
  return CER_OK;
}

int cer_lookup_encode(const void* blob, const float* volume, const float* origin, float* disp, int D, float incre, int h,
                      int w, void* e1, cer_stream_t stream) {
  CER_REQUIRE(blob && volume && origin && disp && e1, "cer_lookup_encode: null pointer");
  CER_REQUIRE(h > 0 && w > 0 && D >= 8 && D <= 256, "cer_lookup_encode: bad arguments (8 <= D <= 256)");
  CER_REQUIRE(aligned16(blob) && aligned16(volume) && aligned16(e1), "cer_lookup_encode: pointers must be 16-byte aligned");
  int rc;
  if ((rc = update_configure())) return rc;
  launch_lookup_enc1(blob, volume, origin, disp, nullptr, 1, nullptr, 0, D, incre, (__half*)e1, h, w, (cudaStream_t)stream);
  return check_launch("cer_lookup_encode");
}

int cer_gru_step(const void* blob, void* workspace, void* net, const void* inp, const void* dn, const void* e, int h,
                 int w, cer_stream_t stream) {
  CER_REQUIRE(blob && workspace && net && inp && dn && e && h > 0 && w > 0, "cer_gru_step: bad arguments");
  int rc;
  if ((rc = update_configure())) return rc;
  const BlobLayout L = blob_layout();
  const char* B = (const char*)blob;
  UpdateWs ws = carve_ws(workspace, h, w);
  ConvArgs a{};
  a.h = h; a.w = w; a.dn_chunk = -1;
  a.src[0] = (const __half*)net; a.src[1] = (const __half*)inp; a.src[2] = (const __half*)dn; a.src[3] = (const __half*)e;
  a.n_src = 4; a.wpk = (const __half*)(B + L.wg); a.wtc = (const __half*)(B + L.t_wg); a.wtc2 = (const __half*)(B + L.p_wg); a.bias = (const float*)(B + L.bg);
  a.net = (__half*)net; a.z = ws.z; a.rnet = ws.rnet; a.qx = ws.qx;
  if ((rc = launch_conv<192, EPI_GATES>(a, (cudaStream_t)stream))) return rc;
  a.src[0] = ws.rnet; a.n_src = 1; a.wpk = (const __half*)(B + L.wq); a.wtc = (const __half*)(B + L.t_wq); a.bias = nullptr;
  return launch_conv<64, EPI_GRUOUT>(a, (cudaStream_t)stream);
}

int cer_update_step(const void* blob, void* workspace, void* net, const void* inp, float* disp, const float* corr,
                    int slots, float* delta, int apply_delta, int stage, int h, int w, cer_stream_t stream) {
  CER_REQUIRE(blob && workspace && net && inp && disp && corr, "cer_update_step: null pointer");
  CER_REQUIRE(slots > 0 && h > 0 && w > 0 && (stage == 0 || stage == 1), "cer_update_step: bad arguments");
  CER_REQUIRE(aligned16(blob) && aligned16(workspace) && aligned16(net) && aligned16(inp),
              "cer_update_step: pointers must be 16-byte aligned");
  return update_step_hmma(blob, workspace, net, inp, disp, corr, slots, delta, apply_delta, stage, h, w,
                          (cudaStream_t)stream);
}

}  // extern "C"
