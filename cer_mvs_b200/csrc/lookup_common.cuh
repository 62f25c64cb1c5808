// Pyramid lookup arithmetic shared by the stand-alone lookup kernel (corr_ops.cu) and the fused
// lookup + corr-encoder kernel (update_hmma.cu).  Restates core/corr.py:102-143 +
// utils/bilinear_sampler.py:6-25 (grid_sample bilinear / zeros / align_corners=True on a 1 x W_l row).
#pragma once
#include "common.cuh"

namespace cer {

int lookup_variant();   // 2 = warp-autonomous kernels for the reference configuration (default), 1 = general kernels only

__device__ __forceinline__ float pyr_value(const float* row, int lvl, int i, int D) {
  // value i of pyramid level lvl (floor pooling); caller guarantees 0 <= i < (D >> lvl)
  if (lvl == 0) return row[i];
  if (lvl == 1) return (row[2 * i] + row[2 * i + 1]) * 0.5f;
  const float a = (row[4 * i] + row[4 * i + 1]) * 0.5f;
  const float b = (row[4 * i + 2] + row[4 * i + 3]) * 0.5f;
  return (a + b) * 0.5f;
}

// coords = max((zinv - origin) / incre + D//2, 0)   (core/corr.py:107)
__device__ __forceinline__ float lookup_coord(float z, float o, float incre, int D) {
  return fmaxf(__fadd_rn(__fdiv_rn(__fsub_rn(z, o), incre), (float)(D / 2)), 0.f);
}

// One tap: level lvl (width D >> lvl), offset j, coordinate c (level-0 units); row = level-0 volume row.
__device__ __forceinline__ float lookup_tap(const float* row, int D, int lvl, int j, float c) {
  const int Wl = D >> lvl;
  const float cl = c * (1.f / (float)(1 << lvl));                              // exact (power of two)
  const float wm1 = (float)(Wl - 1);
  const float x0 = __fadd_rn((float)j, cl);                                   // corr.py:129
  // 2*x0/wm1 - 1 (bilinear_sampler.py:12) with the correctly rounded quotient from three FMAs (see lookup_tap_padded)
  const float rcp = __frcp_rn(wm1);
  const float q0 = __fmul_rn(x0, rcp);
  const float q = __fmaf_rn(__fmaf_rn(-q0, wm1, x0), rcp, q0);
  const float xn = __fmaf_rn(2.f, q, -1.f);
  const float xp = __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.f), 0.5f), wm1);       // grid_sample unnormalize (/2 is exact)
  const float fl = floorf(xp);
  const float w1 = xp - fl;
  const float w0 = (fl + 1.f) - xp;
  float v = 0.f;
  // zero padding: taps outside [0, Wl) contribute nothing; huge |xp| is out on both sides
  if (fl >= -1.f && fl < (float)Wl) {
    const int i0 = (int)fl;
    const float v0 = (i0 >= 0) ? pyr_value(row, lvl, i0, D) : 0.f;
    const float v1 = (i0 + 1 < Wl) ? pyr_value(row, lvl, i0 + 1, D) : 0.f;
    v = v0 * w0 + v1 * w1;
  }
  return v;
}

// ---- the same tap on zero-padded per-level rows (lookup_enc1_v2_kernel, lookup_v2_kernel) ----------------------
// lrow[-1] = lrow[Wl] = 0 (the grid_sample zero padding) and lrow[0..Wl) holds pyramid level l, so a tap is two
// unconditional loads.  The division by (Wl - 1) is done as q0 = x0 * rcp, r = fma(-q0, wm1, x0), q = fma(r, rcp, q0)
// with rcp = RN(1 / wm1): q is the correctly rounded quotient (Markstein), i.e. bit-identical to __fdiv_rn -- the
// exact multiplications by 2 and 0.5 of the reference formula commute with the roundings and are folded into
// fma(2, q, -1) and hw = 0.5 * wm1.  (tests/test_cpu_abi_and_host.py checks the identity on the host for every
// divisor the kernels can see.)
struct LevelConst {
  float fW, wm1, rcp, hw;
};
__host__ __device__ constexpr LevelConst level_const(int D, int lvl) {
  const int Wl = D >> lvl;
  return LevelConst{(float)Wl, (float)(Wl - 1), 1.f / (float)(Wl - 1), 0.5f * (float)(Wl - 1)};
}
__device__ __forceinline__ float lookup_tap_padded(const float* lrow, const LevelConst k, float cl, int j) {
  const float x0 = __fadd_rn((float)j, cl);                                   // corr.py:129
  const float q0 = __fmul_rn(x0, k.rcp);
  const float q = __fmaf_rn(__fmaf_rn(-q0, k.wm1, x0), k.rcp, q0);            // == x0 / wm1, correctly rounded
  const float xn = __fmaf_rn(2.f, q, -1.f);                                   // == 2*x0/wm1 - 1 (bilinear_sampler.py:12)
  const float xp = __fmul_rn(__fadd_rn(xn, 1.f), k.hw);                       // == ((xn + 1) / 2) * wm1
  const float fl = floorf(xp);
  const float w1 = xp - fl;
  const float w0 = (fl + 1.f) - xp;
  const bool inr = fl >= -1.f && fl < k.fW;                                   // false for NaN
  const int i0 = inr ? (int)fl : -1;
  const float v = lrow[i0] * w0 + lrow[i0 + 1] * w1;
  return inr ? v : 0.f;
}

// ---- warp-autonomous pyramid staging (lookup_enc1_v2_kernel, lookup_v2_kernel) -------------------------------------
// A warp owns 32 consecutive pixels = 32 * D contiguous floats of the D-minor volume.  Per pixel and level one padded
// row in shared memory: [0 | level values | 0], odd pitches so that "lane = pixel" tap reads hit 32 banks.
constexpr int kPyrP0 = 67, kPyrP1 = 35, kPyrP2 = 19;        // floats per pixel row incl. the two pads (D <= 64)
constexpr int kPyrWarpFloats = 32 * (kPyrP0 + kPyrP1 + kPyrP2);

// every lane's D/4 16-byte pieces of the chunk, all in flight at once (piece f = it * 32 + lane: coalesced)
template <int D>
__device__ __forceinline__ void pyr_load_chunk(const float* __restrict__ chunk, int npix, int lane, float4 (&rv)[D / 4]) {
  const float4* vsrc = reinterpret_cast<const float4*>(chunk);
  const int nvec = npix * (D / 4);
#pragma unroll
  for (int it = 0; it < D / 4; ++it) {
    const int f = it * 32 + lane;
    rv[it] = f < nvec ? __ldg(vsrc + f) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// registers -> padded level rows; levels 1 and 2 (floor pooling, same expressions as pyr_value) come straight from the
// piece.  Consecutive lanes are 4 words apart, so the four scalar stores are rotated by lane / 8 (level 1: by lane / 16)
// to spread over the banks.  Caller: __syncwarp() before the first tap.
template <int D>
__device__ __forceinline__ void pyr_store_chunk(const float4 (&rv)[D / 4], float* L0, float* L1, float* L2, int lane) {
  static_assert(D % 4 == 0 && D + 2 <= kPyrP0 && D / 2 + 2 <= kPyrP1 && D / 4 + 2 <= kPyrP2, "pitches are sized for D <= 64");
  constexpr int NV = D / 4;
  {
    float* r0 = L0 + lane * kPyrP0;
    float* r1 = L1 + lane * kPyrP1;
    float* r2 = L2 + lane * kPyrP2;
    r0[0] = 0.f; r0[D + 1] = 0.f;
    r1[0] = 0.f; r1[D / 2 + 1] = 0.f;
    r2[0] = 0.f; r2[D / 4 + 1] = 0.f;
  }
  const int rot = lane >> 3, sw = (lane >> 4) & 1;
#pragma unroll
  for (int it = 0; it < NV; ++it) {
    const int f = it * 32 + lane;
    const int pp = f / NV, i = (f % NV) * 4;               // pixel within the chunk, first element of the piece
    const float4 v = rv[it];
    float* d0 = L0 + pp * kPyrP0 + 1 + i;
    // (c0..c3) = (x, y, z, w) rotated left by rot: component k is element (k + rot) & 3
    const float a0 = (rot & 1) ? v.y : v.x, a1 = (rot & 1) ? v.z : v.y, a2 = (rot & 1) ? v.w : v.z, a3 = (rot & 1) ? v.x : v.w;
    const float c0 = (rot & 2) ? a2 : a0, c1 = (rot & 2) ? a3 : a1, c2 = (rot & 2) ? a0 : a2, c3 = (rot & 2) ? a1 : a3;
    d0[(0 + rot) & 3] = c0;
    d0[(1 + rot) & 3] = c1;
    d0[(2 + rot) & 3] = c2;
    d0[(3 + rot) & 3] = c3;
    const float l1a = (v.x + v.y) * 0.5f, l1b = (v.z + v.w) * 0.5f;     // pyr_value(lvl 1)
    float* d1 = L1 + pp * kPyrP1 + 1 + i / 2;
    d1[sw] = sw ? l1b : l1a;
    d1[sw ^ 1] = sw ? l1a : l1b;
    L2[pp * kPyrP2 + 1 + i / 4] = (l1a + l1b) * 0.5f;                   // pyr_value(lvl 2)
  }
}

// the 3 x 11 taps of this lane's pixel (core/corr.py:126-142 order: level-major, then offset -5..5)
template <int D>
__device__ __forceinline__ void pyr_taps33(const float* L0, const float* L1, const float* L2, int lane, float c,
                                           float (&tp)[33]) {
  const float* r0 = L0 + lane * kPyrP0 + 1;
  const float* r1 = L1 + lane * kPyrP1 + 1;
  const float* r2 = L2 + lane * kPyrP2 + 1;
  constexpr LevelConst k0 = level_const(D, 0), k1 = level_const(D, 1), k2 = level_const(D, 2);
  const float c1 = c * 0.5f, c2 = c * 0.25f;
#pragma unroll
  for (int j = 0; j < 11; ++j) {
    tp[j] = lookup_tap_padded(r0, k0, c, j - 5);
    tp[11 + j] = lookup_tap_padded(r1, k1, c1, j - 5);
    tp[22 + j] = lookup_tap_padded(r2, k2, c2, j - 5);
  }
}

}  // namespace cer
