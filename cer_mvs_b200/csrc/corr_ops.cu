// Correlation-side small kernels: error plumbing, the alt_cuda_corr.forward drop-in, layout
// changes, projection matrices, pyramid pooling and the fused pyramid lookup.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "lookup_common.cuh"

namespace cer {

static thread_local char g_err[512] = "";
thread_local long long g_launches = 0;
int g_pdl = []() { const char* e = getenv("CER_PDL"); return (e && e[0] == '0') ? 0 : 1; }();

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return CER_OK;
}

// ------------------------------------------------------------------------------------------
// alt_cuda_corr.forward drop-in.  One 8-lane group per (b, n, y, x) sample; lanes stride over the
// channel dimension (float4 when C % 4 == 0), 3 xor-shuffles reduce the dot products.
// Restates correlation_kernel.cu:59-116: output (oy, ox) of the (2r+1)^2 window is the bilinear
// blend of the dots at integer pixels (fy-r+oy+{0,1}, fx-r+ox+{0,1}); index oy + rd*ox.
// ------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256) corr_forward_dropin_kernel(
    const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ coords,
    float* __restrict__ corr, int B, int H1, int W1, int H2, int W2, int C, int N, int r) {
  const long long total = (long long)B * N * H1 * W1;
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int lane = threadIdx.x & 7;
  const bool valid = gid < total;
  const long long g = valid ? gid : total - 1;
  const int x = (int)(g % W1);
  const int y = (int)((g / W1) % H1);
  const int n = (int)((g / ((long long)W1 * H1)) % N);
  const int b = (int)(g / ((long long)W1 * H1 * N));
  const float cx = __ldg(coords + 2 * g);
  const float cy = __ldg(coords + 2 * g + 1);
  const float fxf = floorf(cx), fyf = floorf(cy);
  const float dx = cx - fxf, dy = cy - fyf;
  const int fx = (int)fxf, fy = (int)fyf;  // NaN -> 0 like the reference's cvt
  const float* p1 = f1 + (((long long)b * H1 + y) * W1 + x) * C;
  const float* img2 = f2 + (long long)b * H2 * W2 * C;
  const int rd = 2 * r + 1;
  const long long plane = (long long)H1 * W1;
  float* out = corr + (((long long)b * N + n) * rd * rd) * plane + (long long)y * W1 + x;
  for (int ox = 0; ox < rd; ++ox) {
    for (int oy = 0; oy < rd; ++oy) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int h2 = fy - r + oy + (k >> 1);
        const int w2 = fx - r + ox + (k & 1);
        float s = 0.f;
        if (h2 >= 0 && h2 < H2 && w2 >= 0 && w2 < W2) {
          const float* p2 = img2 + ((long long)h2 * W2 + w2) * C;
          if (VEC4) {
            for (int c = lane * 4; c < C; c += 32) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(p1 + c));
              const float4 q = __ldg(reinterpret_cast<const float4*>(p2 + c));
              s = fmaf(a.x, q.x, s);
              s = fmaf(a.y, q.y, s);
              s = fmaf(a.z, q.z, s);
              s = fmaf(a.w, q.w, s);
            }
          } else {
            for (int c = lane; c < C; c += 8) s = fmaf(__ldg(p1 + c), __ldg(p2 + c), s);
          }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        acc += (s * ((k >> 1) ? dy : 1.f - dy)) * ((k & 1) ? dx : 1.f - dx);   // s*(1-dy)*(1-dx) etc., kernel.cu:97-100
      }
      if (valid && lane == 0) out[(long long)(oy + rd * ox) * plane] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------
// NCHW <-> NHWC with dtype change and scale.  Tile = 64 channels x 32 pixels through smem.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const TS* __restrict__ src, TD* __restrict__ dst,
                                                          int C, int dstC, long long px, float scale) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const TS* s = src + (long long)n * C * px;
  TD* d = dst + (long long)n * dstC * px;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    const long long p = p0 + tx;
    tile[ty + 8 * k][tx] = (c < C && p < px) ? to_f<TS>(s[(long long)c * px + p]) * scale : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long p = p0 + ty + 8 * k;
    const int c = c0 + tx;
    if (c < dstC && p < px) d[p * dstC + c] = from_f<TD>(tile[tx][ty + 8 * k]);   // channels C..dstC-1 are zero
  }
}


// fp16 -> fp16, 64 channels (the feature / context maps of the plan): 32 pixels x 64 channels per block, 16-byte loads
// along the pixel axis of every channel plane, 16-byte stores that cover whole 128-byte NHWC pixels (512 contiguous
// bytes per warp store).  `scale` is applied exactly like the general kernel (float multiply, round to fp16).
__global__ void __launch_bounds__(256) nchw_to_nhwc_c64_h16_kernel(const __half* __restrict__ src, __half* __restrict__ dst,
                                                                  long long px, float scale) {
  constexpr int kPitch = 72;                       // halfs per pixel row of the tile (144 B: 16-byte aligned rows)
  __shared__ __align__(16) __half tile[32 * kPitch];
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * 32;
  const __half* s = src + (long long)n * 64 * px;
  __half* d = dst + (long long)n * 64 * px;
  const int t = threadIdx.x;
  {
    const int c = t >> 2, q = t & 3;               // channel, group of 8 pixels
    const long long p = p0 + 8 * q;
    __align__(16) __half v[8];
    if (p + 8 <= px) {
      *reinterpret_cast<uint4*>(v) = __ldcs(reinterpret_cast<const uint4*>(s + (long long)c * px + p));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (p + i < px) ? s[(long long)c * px + p + i] : __float2half_rn(0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[(8 * q + i) * kPitch + c] = __float2half_rn(__half2float(v[i]) * scale);
  }
  __syncthreads();
  {
    const int pl = t >> 3, ch = t & 7;             // pixel of the tile, group of 8 channels
    if (p0 + pl < px)
      *reinterpret_cast<uint4*>(d + (p0 + pl) * 64 + ch * 8) = *reinterpret_cast<const uint4*>(tile + pl * kPitch + ch * 8);
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const TS* __restrict__ src, TD* __restrict__ dst,
                                                          int C, long long px) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const TS* s = src + (long long)n * C * px;
  TD* d = dst + (long long)n * C * px;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long p = p0 + ty + 8 * k;
    const int c = c0 + tx;
    tile[ty + 8 * k][tx] = (c < C && p < px) ? to_f<TS>(s[p * C + c]) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    const long long p = p0 + tx;
    if (c < C && p < px) d[(long long)c * px + p] = from_f<TD>(tile[tx][ty + 8 * k]);
  }
}

// ------------------------------------------------------------------------------------------
// Pij = K4_j P_j P_i^-1 K4_i^-1 in fp64 (one thread per pair), utils/projective_ops.py:16-23
// ------------------------------------------------------------------------------------------
__device__ void mat4_mul(const double* a, const double* b, double* c) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      c[i * 4 + j] = s;
    }
}

__device__ void mat4_inv(const double* m, double* inv) {  // Gauss-Jordan, partial pivoting
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      a[i][j] = m[i * 4 + j];
      a[i][4 + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (piv != c)
      for (int j = 0; j < 8; ++j) {
        double t = a[c][j];
        a[c][j] = a[piv][j];
        a[piv][j] = t;
      }
    const double d = 1.0 / a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] *= d;
    for (int r = 0; r < 4; ++r)
      if (r != c) {
        const double f = a[r][c];
        for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) inv[i * 4 + j] = a[i][4 + j];
}

__global__ void projection_matrices_kernel(const float* __restrict__ poses, const float* __restrict__ K,
                                           const int* __restrict__ ii, const int* __restrict__ jj,
                                           int n_pairs, float* __restrict__ Pij) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pairs) return;
  const int i = ii[k], j = jj[k];
  double Ki[16], Kj[16], Pi[16], Pj[16], t0[16], t1[16], t2[16];
  for (int a = 0; a < 16; ++a) {
    Ki[a] = Kj[a] = 0.0;
    Pi[a] = poses[i * 16 + a];
    Pj[a] = poses[j * 16 + a];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      Ki[r * 4 + c] = K[i * 9 + r * 3 + c];
      Kj[r * 4 + c] = K[j * 9 + r * 3 + c];
    }
  Ki[15] = Kj[15] = 1.0;
  mat4_mul(Kj, Pj, t0);
  mat4_inv(Pi, t1);
  mat4_mul(t0, t1, t2);
  mat4_inv(Ki, t1);
  mat4_mul(t2, t1, t0);
  for (int a = 0; a < 16; ++a) Pij[k * 16 + a] = (float)t0[a];
}

// ------------------------------------------------------------------------------------------
// avg_pool2d([1,2]) (core/corr.py:96)
// ------------------------------------------------------------------------------------------
__global__ void pool_pairs_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int W) {
  const int Wo = W / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Wo) return;
  const long long r = idx / Wo;
  const int c = (int)(idx % Wo);
  dst[idx] = (src[r * W + 2 * c] + src[r * W + 2 * c + 1]) * 0.5f;
}

// ------------------------------------------------------------------------------------------
// Fused pyramid lookup (core/corr.py:102-143).  Block = 128 pixels of one slot: the level-0 rows are
// staged in shared memory with coalesced 16-byte loads (4*D bytes per pixel is the only volume read),
// levels 1..2 are rebuilt from it on the fly, each thread produces the L*(2r+1) taps of its pixel.
// ------------------------------------------------------------------------------------------
constexpr int kLookupPix = 128;

__global__ void __launch_bounds__(kLookupPix) lookup_kernel(
    const float* __restrict__ volume, const float* __restrict__ origin, const float* __restrict__ zinv,
    long long zinv_stride, int D, float incre, int radius, int num_levels, float* __restrict__ out,
    long long px) {
  extern __shared__ float rows[];  // [kLookupPix][D + 1]
  const int slot = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * kLookupPix;
  const int npix = (int)min((long long)kLookupPix, px - p0);
  const float* vsrc = volume + ((long long)slot * px + p0) * D;
  const int pitch = D + 1;
  if ((D & 3) == 0) {
    const int nvec = npix * D / 4;
    for (int i = threadIdx.x; i < nvec; i += kLookupPix) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(vsrc) + i);
      const int e = i * 4;
      const int r = e / D, c = e % D;
      float* d = rows + r * pitch + c;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  } else {
    for (int i = threadIdx.x; i < npix * D; i += kLookupPix) rows[(i / D) * pitch + (i % D)] = __ldg(vsrc + i);
  }
  __syncthreads();
  if ((int)threadIdx.x >= npix) return;
  const long long p = p0 + threadIdx.x;
  const float* row = rows + threadIdx.x * pitch;
  const float z = __ldg(zinv + slot * zinv_stride + p);
  const float o = __ldg(origin + p);
  const float c = lookup_coord(z, o, incre, D);
  const int taps = 2 * radius + 1;
  float* o_ptr = out + (long long)slot * num_levels * taps * px + p;
  for (int lvl = 0; lvl < num_levels; ++lvl)
    for (int j = -radius; j <= radius; ++j)
      o_ptr[(long long)(lvl * taps + (j + radius)) * px] = lookup_tap(row, D, lvl, j, c);
}

// ------------------------------------------------------------------------------------------
// The reference configuration (radius 5, 3 levels, D = 64 / 44: core/raft.py:77-81) takes a warp-autonomous kernel:
// a warp owns 32 consecutive pixels of one slot, brings their 32 * D contiguous floats in with 16-byte loads that
// are all in flight at once, materialises the zero-padded three-level pyramid in its own shared-memory slice
// (no block barrier anywhere), then lane = pixel evaluates the 33 taps and every tap leaves as one coalesced
// 128-byte store.  HBM traffic per (pixel, slot): 4 * D + 8 in, 132 out (SURVEY 8d counts 284 B algorithmic).
// ------------------------------------------------------------------------------------------
constexpr int kLookupV2Warps = 4;

template <int D>
__global__ void __launch_bounds__(kLookupV2Warps * 32, 3) lookup_v2_kernel(
    const float* __restrict__ volume, const float* __restrict__ origin, const float* __restrict__ zinv,
    long long zinv_stride, float incre, float* __restrict__ out, int px) {
  extern __shared__ __align__(16) float pyr[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* L0 = pyr + warp * kPyrWarpFloats;
  float* L1 = L0 + 32 * kPyrP0;
  float* L2 = L1 + 32 * kPyrP1;
  const int slot = blockIdx.y;
  const int p0 = (blockIdx.x * kLookupV2Warps + warp) * 32;
  if (p0 >= px) return;
  const int npix = min(32, px - p0);
  const bool live = lane < npix;
  const int p = p0 + (live ? lane : 0);
  float4 rv[D / 4];
  pyr_load_chunk<D>(volume + ((long long)slot * px + p0) * D, npix, lane, rv);
  const float z = __ldg(zinv + slot * zinv_stride + p);
  const float o = __ldg(origin + p);
  pyr_store_chunk<D>(rv, L0, L1, L2, lane);
  const float c = lookup_coord(z, o, incre, D);
  __syncwarp();
  float tp[33];
  pyr_taps33<D>(L0, L1, L2, lane, c, tp);
  if (live) {
    float* o_ptr = out + (long long)slot * 33 * px + p;
#pragma unroll
    for (int k = 0; k < 33; ++k) __stcs(o_ptr + (long long)k * px, tp[k]);      // streamed: written once, read by the encoder
  }
}

}  // namespace cer

using namespace cer;

extern "C" {

int cer_abi_version(void) { return 1; }
const char* cer_last_error(void) { return cer::g_err; }

int cer_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device: %s (cer_mvs_b200 has no CPU fallback)", cudaGetErrorString(e));
    return CER_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  CER_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    return CER_ERR_NO_DEVICE;
  }
  return CER_OK;
}

int cer_corr_forward_f32(const float* fmap1, const float* fmap2, const float* coords, float* corr, int B,
                         int H1, int W1, int H2, int W2, int C, int N, int radius, cer_stream_t stream) {
  CER_REQUIRE(B >= 0 && H1 >= 0 && W1 >= 0 && H2 > 0 && W2 > 0 && C > 0 && N >= 0 && radius >= 0,
              "cer_corr_forward_f32: bad sizes");
  const long long total = (long long)B * N * H1 * W1;
  if (total == 0) return CER_OK;   // empty output (pointers of empty tensors may be null)
  CER_REQUIRE(fmap1 && fmap2 && coords && corr, "cer_corr_forward_f32: null pointer");
  const long long threads = total * 8;
  CER_REQUIRE(threads / 256 < 0x7fffffffLL, "cer_corr_forward_f32: problem too large");
  const bool vec = (C % 4 == 0) && aligned16(fmap1) && aligned16(fmap2);
  const int grid = ceil_div(threads, 256);
  if (vec)
    CER_LAUNCH(KK_CORR_DROPIN, corr_forward_dropin_kernel<true>, grid, 256, 0, stream, fmap1, fmap2, coords, corr, B, H1, W1, H2,
               W2, C, N, radius);
  else
    CER_LAUNCH(KK_CORR_DROPIN, corr_forward_dropin_kernel<false>, grid, 256, 0, stream, fmap1, fmap2, coords, corr, B, H1, W1, H2,
               W2, C, N, radius);
  return check_launch("cer_corr_forward_f32");
}

int cer_nchw_to_nhwc_pad(const void* src, int src_f16, void* dst, int dst_f16, int n, int C, int dstC, int h,
                         int w, float scale, cer_stream_t stream) {
  CER_REQUIRE(src && dst && n > 0 && C > 0 && dstC >= C && h > 0 && w > 0, "cer_nchw_to_nhwc: bad arguments");
  const long long px = (long long)h * w;
  dim3 grid(ceil_div(px, 32), ceil_div(dstC, 32), n);
  if (src_f16 && dst_f16 && C == 64 && dstC == 64 && (px % 8) == 0 && aligned16(src) && aligned16(dst) && n <= 65535) {
    CER_LAUNCH(KK_LAYOUT, nchw_to_nhwc_c64_h16_kernel, dim3(ceil_div(px, 32), n), 256, 0, stream, (const __half*)src,
               (__half*)dst, px, scale);
    return check_launch("cer_nchw_to_nhwc");
  }
  if (src_f16 && dst_f16)
    CER_LAUNCH(KK_LAYOUT, (nchw_to_nhwc_kernel<__half, __half>), grid, 256, 0, stream, (const __half*)src, (__half*)dst, C, dstC, px, scale);
  else if (src_f16 && !dst_f16)
    CER_LAUNCH(KK_LAYOUT, (nchw_to_nhwc_kernel<__half, float>), grid, 256, 0, stream, (const __half*)src, (float*)dst, C, dstC, px, scale);
  else if (!src_f16 && dst_f16)
    CER_LAUNCH(KK_LAYOUT, (nchw_to_nhwc_kernel<float, __half>), grid, 256, 0, stream, (const float*)src, (__half*)dst, C, dstC, px, scale);
  else
    CER_LAUNCH(KK_LAYOUT, (nchw_to_nhwc_kernel<float, float>), grid, 256, 0, stream, (const float*)src, (float*)dst, C, dstC, px, scale);
  return check_launch("cer_nchw_to_nhwc");
}

int cer_nchw_to_nhwc(const void* src, int src_f16, void* dst, int dst_f16, int n, int C, int h, int w,
                     float scale, cer_stream_t stream) {
  return cer_nchw_to_nhwc_pad(src, src_f16, dst, dst_f16, n, C, C, h, w, scale, stream);
}

int cer_nhwc_to_nchw(const void* src, int src_f16, void* dst, int dst_f16, int n, int C, int h, int w,
                     cer_stream_t stream) {
  CER_REQUIRE(src && dst && n > 0 && C > 0 && h > 0 && w > 0, "cer_nhwc_to_nchw: bad arguments");
  const long long px = (long long)h * w;
  dim3 grid(ceil_div(px, 32), ceil_div(C, 32), n);
  if (src_f16 && dst_f16)
    CER_LAUNCH(KK_LAYOUT, (nhwc_to_nchw_kernel<__half, __half>), grid, 256, 0, stream, (const __half*)src, (__half*)dst, C, px);
  else if (src_f16 && !dst_f16)
    CER_LAUNCH(KK_LAYOUT, (nhwc_to_nchw_kernel<__half, float>), grid, 256, 0, stream, (const __half*)src, (float*)dst, C, px);
  else if (!src_f16 && dst_f16)
    CER_LAUNCH(KK_LAYOUT, (nhwc_to_nchw_kernel<float, __half>), grid, 256, 0, stream, (const float*)src, (__half*)dst, C, px);
  else
    CER_LAUNCH(KK_LAYOUT, (nhwc_to_nchw_kernel<float, float>), grid, 256, 0, stream, (const float*)src, (float*)dst, C, px);
  return check_launch("cer_nhwc_to_nchw");
}

int cer_projection_matrices(const float* poses, const float* intrinsics, const int* ii, const int* jj,
                            int n_pairs, float* Pij, cer_stream_t stream) {
  CER_REQUIRE(poses && intrinsics && ii && jj && Pij && n_pairs > 0, "cer_projection_matrices: bad arguments");
  CER_LAUNCH(KK_PROJ, projection_matrices_kernel, ceil_div(n_pairs, 32), 32, 0, stream, poses, intrinsics, ii, jj, n_pairs, Pij);
  return check_launch("cer_projection_matrices");
}

int cer_pool_pairs(const float* src, float* dst, long long rows, int W, cer_stream_t stream) {
  CER_REQUIRE(src && dst && rows >= 0 && W >= 2, "cer_pool_pairs: bad arguments");
  const long long total = rows * (W / 2);
  if (total == 0) return CER_OK;
  CER_LAUNCH(KK_POOL, pool_pairs_kernel, ceil_div(total, 256), 256, 0, stream, src, dst, rows, W);
  return check_launch("cer_pool_pairs");
}

int cer_lookup_strided(const float* volume, int slots, const float* origin, const float* zinv,
                       long long zinv_stride, int D, float incre, int radius, int num_levels, float* out, int h,
                       int w, cer_stream_t stream) {
  CER_REQUIRE(volume && origin && zinv && out, "cer_lookup: null pointer");
  CER_REQUIRE(slots > 0 && h > 0 && w > 0 && radius >= 0, "cer_lookup: bad sizes");
  CER_REQUIRE(num_levels >= 1 && num_levels <= 3, "cer_lookup: num_levels must be 1..3");
  CER_REQUIRE((D >> (num_levels - 1)) >= 2, "cer_lookup: D too small for %d levels", num_levels);
  CER_REQUIRE(D <= 1024, "cer_lookup: D > 1024 unsupported");
  const long long px = (long long)h * w;
  if (lookup_variant() >= 2 && radius == 5 && num_levels == 3 && (D == 64 || D == 44) && px < (1ll << 31) - 64) {
    static std::atomic<unsigned long long> configured{0};
    const size_t smem2 = (size_t)kLookupV2Warps * kPyrWarpFloats * sizeof(float);
    if (first_time_on_device(configured)) {
      CER_CUDA(cudaFuncSetAttribute(lookup_v2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      CER_CUDA(cudaFuncSetAttribute(lookup_v2_kernel<44>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    }
    dim3 grid2(ceil_div(px, kLookupV2Warps * 32), slots);
    if (D == 64)
      CER_LAUNCH(KK_LOOKUP, lookup_v2_kernel<64>, grid2, kLookupV2Warps * 32, smem2, stream, volume, origin, zinv, zinv_stride,
                 incre, out, (int)px);
    else
      CER_LAUNCH(KK_LOOKUP, lookup_v2_kernel<44>, grid2, kLookupV2Warps * 32, smem2, stream, volume, origin, zinv, zinv_stride,
                 incre, out, (int)px);
    return check_launch("cer_lookup");
  }
  const size_t smem = (size_t)kLookupPix * (D + 1) * sizeof(float);
  if (smem > 48 * 1024)
    CER_CUDA(cudaFuncSetAttribute(lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(px, kLookupPix), slots);
  CER_LAUNCH(KK_LOOKUP, lookup_kernel, grid, kLookupPix, smem, stream, volume, origin, zinv, zinv_stride, D, incre, radius,
             num_levels, out, px);
  return check_launch("cer_lookup");
}

int cer_lookup(const float* volume, int slots, const float* origin, const float* zinv, int D, float incre,
               int radius, int num_levels, float* out, int h, int w, cer_stream_t stream) {
  return cer_lookup_strided(volume, slots, origin, zinv, 0, D, incre, radius, num_levels, out, h, w, stream);
}

}  // extern "C"
