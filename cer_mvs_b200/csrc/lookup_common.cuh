// Pyramid lookup arithmetic shared by the stand-alone lookup kernel (corr_ops.cu) and the fused
// lookup + corr-encoder kernel (update_hmma.cu).  Restates core/corr.py:102-143 +
// utils/bilinear_sampler.py:6-25 (grid_sample bilinear / zeros / align_corners=True on a 1 x W_l row).
#pragma once
#include "common.cuh"

namespace cer {

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

}  // namespace cer
