// the small element-wise kernels either side of the hot path, so that the per-image
// critical path has no host round trip:
//   * image preparation: bilinear rescale with align_corners=True (utils/data_utils.py:58-66, the rescale=2 pass)
//     and the [0,255] -> [-1,1] normalisation (core/raft.py:40-41);
//   * output: depth = where(disp == 0, 0, 1 / disp) (inference.py:57-58), optionally written bottom-up like the PFM
//     writer stores it (utils/frame_utils.py:145);
//   * multi-resolution merge (multires.py:26-28): bilinear (cv2.INTER_LINEAR, half-pixel centres) upsampling of the
//     scale-1 depth map to the scale-2 grid, then the relative-consistency select.
// All of them are HBM-bound streaming kernels: 16-byte accesses where the layout allows, grid = a multiple of the SM
// count, grid-stride loops.
#include "common.cuh"

namespace cer {

constexpr int kIoThreads = 256;
static inline int io_grid(long long work_items) {
  const long long blocks = (work_items + kIoThreads - 1) / kIoThreads;
  const long long cap = (long long)kNumSMs * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// images *= 2 / 255.; images -= 1   (two roundings, like the two in-place tensor ops)
__global__ void __launch_bounds__(kIoThreads) normalize_images_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                     long long n, float mul) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = __ldcs(reinterpret_cast<const float4*>(src) + i);
    v.x = __fsub_rn(__fmul_rn(v.x, mul), 1.f);
    v.y = __fsub_rn(__fmul_rn(v.y, mul), 1.f);
    v.z = __fsub_rn(__fmul_rn(v.z, mul), 1.f);
    v.w = __fsub_rn(__fmul_rn(v.w, mul), 1.f);
    reinterpret_cast<float4*>(dst)[i] = v;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __fsub_rn(__fmul_rn(src[i], mul), 1.f);
}

// F.interpolate(mode='bilinear', align_corners=True): src index = dst index * (in - 1) / (out - 1), i0 = floor,
// lambda = frac, second tap clamped to the last pixel; out = wy0 (wx0 a + wx1 b) + wy1 (wx0 c + wx1 d).
// One thread = VEC consecutive output pixels of one output row (row taps computed once, 16-byte streaming store);
// blockIdx.x = (plane, output row), so there is no integer division per pixel.
template <int VEC>
__global__ void __launch_bounds__(kIoThreads) resize_bilinear_ac_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                       int h, int w, int h2, int w2, float sy, float sx) {
  const int y2 = blockIdx.x % h2;
  const long long pl = blockIdx.x / h2;
  const int xq = (blockIdx.y * blockDim.x + threadIdx.x) * VEC;
  if (xq >= w2) return;
  const float fy = __fmul_rn(sy, (float)y2);
  const int y0 = min((int)floorf(fy), h - 1);
  const float ly = fminf(fmaxf(fy - (float)y0, 0.f), 1.f);
  const int y1 = min(y0 + 1, h - 1);
  const float wy0 = 1.f - ly;
  const float* r0 = src + (pl * h + y0) * (long long)w;
  const float* r1 = src + (pl * h + y1) * (long long)w;
  float o[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const int x2 = min(xq + k, w2 - 1);
    const float fx = __fmul_rn(sx, (float)x2);
    const int x0 = min((int)floorf(fx), w - 1);
    const float lx = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
    const int x1 = min(x0 + 1, w - 1);
    const float wx0 = 1.f - lx;
    const float a = __ldg(r0 + x0), b = __ldg(r0 + x1), c = __ldg(r1 + x0), d = __ldg(r1 + x1);
    o[k] = __fadd_rn(__fmul_rn(wy0, __fadd_rn(__fmul_rn(wx0, a), __fmul_rn(lx, b))),
                     __fmul_rn(ly, __fadd_rn(__fmul_rn(wx0, c), __fmul_rn(lx, d))));
  }
  float* out = dst + (pl * h2 + y2) * (long long)w2 + xq;
  if (VEC == 4 && xq + 4 <= w2) {
    __stcs(reinterpret_cast<float4*>(out), make_float4(o[0], o[1], o[2], o[3]));
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      if (xq + k < w2) out[k] = o[k];
  }
}

// depth = where(disp == 0, 0, 1 / disp); flip != 0 writes row y to row h-1-y (np.flipud of the PFM writer)
__global__ void __launch_bounds__(kIoThreads) disp_to_depth_kernel(const float* __restrict__ disp, float* __restrict__ depth,
                                                                  int h, int w, int flip) {
  const long long total = (long long)h * w;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const float d = __ldg(disp + i);
    const float v = d == 0.f ? 0.f : __fdiv_rn(1.f, d);
    long long o = i;
    if (flip) {
      const int y = (int)(i / w), x = (int)(i % w);
      o = (long long)(h - 1 - y) * w + x;
    }
    depth[o] = v;
  }
}

// cv2.resize(im1, (w2, h2)) with INTER_LINEAR on float32 (source coordinate (x + 0.5) * (w1 / w2) - 0.5, taps clamped
// to the image with zero weight on the clamped side; horizontal pass first, then vertical), followed by
// mask = |im1r - im2| < th * im1r; out = mask ? im2 : im1r.
__global__ void __launch_bounds__(kIoThreads) multires_merge_kernel(const float* __restrict__ im1, int h1, int w1,
                                                                   const float* __restrict__ im2, int h2, int w2,
                                                                   float th, double sx, double sy, float* __restrict__ out) {
  const long long total = (long long)h2 * w2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int x = (int)(i % w2), y = (int)(i / w2);
    float fx = (float)((x + 0.5) * sx - 0.5), fy = (float)((y + 0.5) * sy - 0.5);
    int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    fx -= (float)x0;
    fy -= (float)y0;
    if (x0 < 0) { x0 = 0; fx = 0.f; }
    if (x0 >= w1 - 1) { x0 = w1 - 1; fx = 0.f; }
    if (y0 < 0) { y0 = 0; fy = 0.f; }
    if (y0 >= h1 - 1) { y0 = h1 - 1; fy = 0.f; }
    const int x1 = min(x0 + 1, w1 - 1), y1 = min(y0 + 1, h1 - 1);
    const float a = __ldg(im1 + (long long)y0 * w1 + x0), b = __ldg(im1 + (long long)y0 * w1 + x1);
    const float c = __ldg(im1 + (long long)y1 * w1 + x0), d = __ldg(im1 + (long long)y1 * w1 + x1);
    const float r0 = __fadd_rn(__fmul_rn(a, 1.f - fx), __fmul_rn(b, fx));
    const float r1 = __fadd_rn(__fmul_rn(c, 1.f - fx), __fmul_rn(d, fx));
    const float v1 = __fadd_rn(__fmul_rn(r0, 1.f - fy), __fmul_rn(r1, fy));
    const float v2 = __ldcs(im2 + i);
    const bool keep2 = fabsf(__fsub_rn(v1, v2)) < __fmul_rn(th, v1);
    __stcs(out + i, keep2 ? v2 : v1);
  }
}

}  // namespace cer

using namespace cer;

extern "C" {

int cer_normalize_images(const float* src, float* dst, long long n, cer_stream_t stream) {
  CER_REQUIRE(src && dst && n > 0, "cer_normalize_images: bad arguments");
  CER_REQUIRE(aligned16(src) && aligned16(dst), "cer_normalize_images: buffers must be 16-byte aligned");
  CER_LAUNCH(KK_LAYOUT, normalize_images_kernel, io_grid(n / 4 + 1), kIoThreads, 0, stream, src, dst, n, (float)(2 / 255.));
  return check_launch("cer_normalize_images");
}

int cer_resize_bilinear_ac(const float* src, float* dst, int planes, int h, int w, int h2, int w2, cer_stream_t stream) {
  CER_REQUIRE(src && dst && planes > 0 && h > 0 && w > 0 && h2 > 0 && w2 > 0, "cer_resize_bilinear_ac: bad arguments");
  const float sy = h2 > 1 ? (float)(h - 1) / (float)(h2 - 1) : 0.f;
  const float sx = w2 > 1 ? (float)(w - 1) / (float)(w2 - 1) : 0.f;
  CER_REQUIRE((long long)planes * h2 <= 0x7fffffffLL, "cer_resize_bilinear_ac: too many output rows");
  if ((w2 & 3) == 0 && aligned16(dst)) {
    dim3 grid((unsigned)((long long)planes * h2), ceil_div(w2 / 4, kIoThreads));
    CER_LAUNCH(KK_LAYOUT, resize_bilinear_ac_kernel<4>, grid, kIoThreads, 0, stream, src, dst, h, w, h2, w2, sy, sx);
  } else {
    dim3 grid((unsigned)((long long)planes * h2), ceil_div(w2, kIoThreads));
    CER_LAUNCH(KK_LAYOUT, resize_bilinear_ac_kernel<1>, grid, kIoThreads, 0, stream, src, dst, h, w, h2, w2, sy, sx);
  }
  return check_launch("cer_resize_bilinear_ac");
}

int cer_disp_to_depth(const float* disp, float* depth, int h, int w, int flip_rows, cer_stream_t stream) {
  CER_REQUIRE(disp && depth && h > 0 && w > 0, "cer_disp_to_depth: bad arguments");
  CER_REQUIRE(disp != depth || !flip_rows, "cer_disp_to_depth: in-place conversion cannot flip rows");
  CER_LAUNCH(KK_FINISH, disp_to_depth_kernel, io_grid((long long)h * w), kIoThreads, 0, stream, disp, depth, h, w, flip_rows);
  return check_launch("cer_disp_to_depth");
}

int cer_multires_merge(const float* im1, int h1, int w1, const float* im2, int h2, int w2, float th, float* out,
                       cer_stream_t stream) {
  CER_REQUIRE(im1 && im2 && out && h1 > 0 && w1 > 0 && h2 > 0 && w2 > 0, "cer_multires_merge: bad arguments");
  CER_LAUNCH(KK_FINISH, multires_merge_kernel, io_grid((long long)h2 * w2), kIoThreads, 0, stream, im1, h1, w1, im2, h2, w2,
             th, (double)w1 / w2, (double)h1 / h2, out);
  return check_launch("cer_multires_merge");
}

}  // extern "C"
