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
  const float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, x0), wm1), 1.f);        // bilinear_sampler.py:12
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

}  // namespace cer
