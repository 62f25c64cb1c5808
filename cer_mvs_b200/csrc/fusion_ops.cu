// the geometric-consistency filter of fusion.py -- reproject_with_depth (fusion.py:39-85) +
// check_geometric_consistency (:88-106) + the per-reference-view aggregation of fusion() (:239-249) -- as ONE kernel
// per reference view instead of ~40 tensor ops per (reference, threshold) pair with [S, h, w] temporaries.
//
// Per reference pixel and source view: back-project with the reference depth, transform into the source camera, project,
// sample the source depth map bilinearly (grid_sample, zeros padding, align_corners=True), back-project with the sampled
// depth, transform back, project: reprojection distance and relative depth difference -> nine (i = 2..10) threshold
// masks.  The same projective-gather pattern as the cost-volume build, with one 4-byte gather per tap: HBM/L2 bound,
// one thread per reference pixel looping over the source views, every per-view quantity kept in registers.
#include "common.cuh"

namespace cer {

struct GeoMats {       // per source view, fp32 (computed in fp64 by geo_matrices_kernel)
  float Kri[9];        // inverse(K_ref)
  float T1[12];        // (E_src * inverse(E_ref))[:3]
  float Ks[9];         // K_src
  float Ksi[9];        // inverse(K_src)
  float T2[12];        // (E_ref * inverse(E_src))[:3]
  float Kr[9];         // K_ref
};

__device__ inline void m4_mul(const double* a, const double* b, double* c) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      c[i * 4 + j] = s;
    }
}
// Gauss-Jordan with partial pivoting, n = 3 or 4 (row-major, stride n)
__device__ inline void m_inv(const double* m, double* inv, int n) {
  double a[4][8];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      a[i][j] = m[i * n + j];
      a[i][n + j] = i == j ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    for (int j = 0; j < 2 * n; ++j) {
      const double t = a[c][j];
      a[c][j] = a[piv][j];
      a[piv][j] = t;
    }
    const double d = 1.0 / a[c][c];
    for (int j = 0; j < 2 * n; ++j) a[c][j] *= d;
    for (int r = 0; r < n; ++r)
      if (r != c) {
        const double f = a[r][c];
        for (int j = 0; j < 2 * n; ++j) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = a[i][n + j];
}

__global__ void geo_matrices_kernel(const float* __restrict__ K_ref, const float* __restrict__ E_ref,
                                    const float* __restrict__ K_src, const float* __restrict__ E_src, int S,
                                    GeoMats* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double kr[9], er[16], ks[9], es[16], t[16], u[16];
  for (int i = 0; i < 9; ++i) {
    kr[i] = K_ref[i];
    ks[i] = K_src[s * 9 + i];
  }
  for (int i = 0; i < 16; ++i) {
    er[i] = E_ref[i];
    es[i] = E_src[s * 16 + i];
  }
  GeoMats g;
  m_inv(kr, t, 3);
  for (int i = 0; i < 9; ++i) g.Kri[i] = (float)t[i];
  m_inv(er, t, 4);
  m4_mul(es, t, u);
  for (int i = 0; i < 12; ++i) g.T1[i] = (float)u[i];
  for (int i = 0; i < 9; ++i) g.Ks[i] = (float)ks[i];
  m_inv(ks, t, 3);
  for (int i = 0; i < 9; ++i) g.Ksi[i] = (float)t[i];
  m_inv(es, t, 4);
  m4_mul(er, t, u);
  for (int i = 0; i < 12; ++i) g.T2[i] = (float)u[i];
  for (int i = 0; i < 9; ++i) g.Kr[i] = (float)kr[i];
  out[s] = g;
}

__device__ __forceinline__ void mul3(const float* M, float a, float b, float c, float& x, float& y, float& z) {
  x = fmaf(M[2], c, fmaf(M[1], b, M[0] * a));
  y = fmaf(M[5], c, fmaf(M[4], b, M[3] * a));
  z = fmaf(M[8], c, fmaf(M[7], b, M[6] * a));
}
__device__ __forceinline__ void mul34(const float* M, float a, float b, float c, float& x, float& y, float& z) {
  x = fmaf(M[2], c, fmaf(M[1], b, M[0] * a)) + M[3];
  y = fmaf(M[6], c, fmaf(M[5], b, M[4] * a)) + M[7];
  z = fmaf(M[10], c, fmaf(M[9], b, M[8] * a)) + M[11];
}

// F.grid_sample(bilinear, zeros, align_corners=True) of one [h, w] map at pixel coordinates (x, y), through the
// normalise / un-normalise round trip of utils/bilinear_sampler.py:33-39
__device__ __forceinline__ float sample_depth(const float* __restrict__ img, int h, int w, float x, float y) {
  const float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, x), (float)(w - 1)), 1.f);
  const float yn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, y), (float)(h - 1)), 1.f);
  const float ix = __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.f), 0.5f), (float)(w - 1));
  const float iy = __fmul_rn(__fmul_rn(__fadd_rn(yn, 1.f), 0.5f), (float)(h - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  if (!(fx >= -1.f && fx <= (float)w && fy >= -1.f && fy <= (float)h)) return 0.f;     // every tap outside (or NaN)
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  auto tap = [&](int yy, int xx) { return (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(img + (long long)yy * w + xx) : 0.f; };
  float v = 0.f;
  v = fmaf(tap(y0, x0), wx0 * wy0, v);              // nw
  v = fmaf(tap(y0, x0 + 1), wx1 * wy0, v);          // ne
  v = fmaf(tap(y0 + 1, x0), wx0 * wy1, v);          // sw
  v = fmaf(tap(y0 + 1, x0 + 1), wx1 * wy1, v);      // se
  return v;
}

struct GeoThresholds {
  float dist[9], rel[9];      // float32(i / thre1), float32(i / thre2), i = 2..10
};

// FULL: also write the per-source tensors check_geometric_consistency returns (masks, masked reprojected depth, source
// coordinates, relative depth difference); otherwise only the aggregated mask / averaged depth of fusion():239-249.
template <bool FULL>
__global__ void __launch_bounds__(256) geo_filter_kernel(
    const float* __restrict__ depth_ref, const float* __restrict__ depth_src, const GeoMats* __restrict__ mats, int S,
    int h, int w, GeoThresholds th, unsigned char* __restrict__ masks, float* __restrict__ depth_rep_out,
    float* __restrict__ xsrc_out, float* __restrict__ ysrc_out, float* __restrict__ rel_out,
    unsigned char* __restrict__ geo_mask, float* __restrict__ depth_est, int* __restrict__ n_valid) {
  extern __shared__ GeoMats sM[];
  for (int i = threadIdx.x; i < S * (int)(sizeof(GeoMats) / 4); i += blockDim.x)
    reinterpret_cast<float*>(sM)[i] = reinterpret_cast<const float*>(mats)[i];
  __syncthreads();
  const long long px = (long long)h * w;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int valid = 0;
  if (p < px) {
    const int x = (int)(p % w), y = (int)(p / w);
    const float xf = (float)x, yf = (float)y;
    const float d = __ldg(depth_ref + p);
    int cnt[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) cnt[i] = 0;
    float dsum = 0.f;
    for (int s = 0; s < S; ++s) {
      const GeoMats& M = sM[s];
      float ax, ay, az, bx, by, bz, cx, cy, cz;
      mul3(M.Kri, xf * d, yf * d, d, ax, ay, az);           // reference 3-D space (fusion.py:50-52)
      mul34(M.T1, ax, ay, az, bx, by, bz);                  // source 3-D space (:55-56)
      mul3(M.Ks, bx, by, bz, cx, cy, cz);                   // source pixel (:58-59)
      const float xs = __fdiv_rn(cx, cz), ys = __fdiv_rn(cy, cz);
      const float sd = sample_depth(depth_src + (long long)s * px, h, w, xs, ys);      // :68
      mul3(M.Ksi, xs * sd, ys * sd, sd, ax, ay, az);        // source 3-D space from the sampled depth (:72-73)
      mul34(M.T2, ax, ay, az, bx, by, bz);                  // back in the reference camera (:75-76)
      mul3(M.Kr, bx, by, bz, cx, cy, cz);
      const float xr = __fdiv_rn(cx, cz), yr = __fdiv_rn(cy, cz);
      const float ddx = xr - xf, ddy = yr - yf;
      const float dist = sqrtf(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));   // :96
      const float rel = __fdiv_rn(fabsf(bz - d), d);                                   // :99-100
      bool last = false;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const bool m = dist < th.dist[i] && rel < th.rel[i];                           // :103-105
        cnt[i] += m ? 1 : 0;
        if (FULL) masks[((long long)i * S + s) * px + p] = m ? 1 : 0;
        if (i == 8) last = m;
      }
      const float dr = last ? bz : 0.f;                                                 // :106
      dsum += dr;
      if (FULL) {
        depth_rep_out[(long long)s * px + p] = dr;
        xsrc_out[(long long)s * px + p] = xs;
        ysrc_out[(long long)s * px + p] = ys;
        rel_out[(long long)s * px + p] = rel;
      }
    }
    // fusion():239-249 with n = S + 1: pixel kept if at least i source views agree at threshold level i (i = 2..S)
    bool keep = cnt[8] >= S + 1;
    for (int i = 2; i < S + 1 && i <= 10; ++i) keep = keep || cnt[i - 2] >= i;
    if (geo_mask) geo_mask[p] = keep ? 1 : 0;
    if (depth_est) depth_est[p] = __fdiv_rn(dsum + d, (float)(cnt[8] + 1));
    valid = keep ? 1 : 0;
  }
  if (n_valid) {
    const int total = __syncthreads_count(valid);
    if (threadIdx.x == 0 && total) atomicAdd(n_valid, total);
  }
}

}  // namespace cer

using namespace cer;

extern "C" {

size_t cer_geo_mats_bytes(int n_src) { return (size_t)n_src * sizeof(GeoMats); }

int cer_geometric_filter(const float* depth_ref, const float* K_ref, const float* E_ref, const float* depth_src,
                         const float* K_src, const float* E_src, int n_src, int h, int w, double thre1, double thre2,
                         void* mats_ws, unsigned char* masks, float* depth_reprojected, float* x_src, float* y_src,
                         float* rel_diff, unsigned char* geo_mask, float* depth_est, int* n_valid, cer_stream_t stream) {
  CER_REQUIRE(depth_ref && K_ref && E_ref && depth_src && K_src && E_src && mats_ws, "cer_geometric_filter: null pointer");
  CER_REQUIRE(n_src >= 1 && n_src <= 10 && h > 1 && w > 1, "cer_geometric_filter: 1..10 source views, h, w > 1");
  const bool full = masks || depth_reprojected || x_src || y_src || rel_diff;
  CER_REQUIRE(!full || (masks && depth_reprojected && x_src && y_src && rel_diff),
              "cer_geometric_filter: the per-source outputs come all together or not at all");
  GeoThresholds th;
  for (int i = 2; i <= 10; ++i) {
    th.dist[i - 2] = (float)((double)i / thre1);      // `dist < i/thre1`: a Python double compared in fp32
    th.rel[i - 2] = (float)((double)i / thre2);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n_valid) CER_CUDA(cudaMemsetAsync(n_valid, 0, sizeof(int), st));
  CER_LAUNCH(KK_PROJ, geo_matrices_kernel, 1, 32, 0, st, K_ref, E_ref, K_src, E_src, n_src, (GeoMats*)mats_ws);
  const long long px = (long long)h * w;
  const size_t smem = (size_t)n_src * sizeof(GeoMats);
  if (full)
    CER_LAUNCH(KK_FINISH, geo_filter_kernel<true>, ceil_div(px, 256), 256, smem, st, depth_ref, depth_src,
               (const GeoMats*)mats_ws, n_src, h, w, th, masks, depth_reprojected, x_src, y_src, rel_diff, geo_mask,
               depth_est, n_valid);
  else
    CER_LAUNCH(KK_FINISH, geo_filter_kernel<false>, ceil_div(px, 256), 256, smem, st, depth_ref, depth_src,
               (const GeoMats*)mats_ws, n_src, h, w, th, masks, depth_reprojected, x_src, y_src, rel_diff, geo_mask,
               depth_est, n_valid);
  return check_launch("cer_geometric_filter");
}

}  // extern "C"
