// UpdateBlock 3x3 convolutions, v2: implicit GEMM on tcgen05.mma (5th-gen tensor cores) with the
// accumulators in TMEM.  Same math, inputs, outputs and fused epilogues as update_hmma.cu.
//
// One CTA = one 16 x 16 pixel tile = two M=128 MMA tiles (16 rows x 8 columns each) that share every
// weight tile, N = 64 / 192 / 256 output channels, K = n_src x 9 taps x 64 channels.
//
//   warps 0-3  A producers: per 64-channel chunk the 18 x 18 halo tile is brought in with 16-byte
//              cp.async (zero fill outside the image = the conv padding) in the UMMA K-major
//              no-swizzle layout [k-group 8][halo pixel 324][8 halfs]; a 3x3 tap is then just a shifted
//              start address of the same tile (LBO = 324*16 B between k-groups, SBO = 18*16 B between
//              image rows = 8-row core-matrix groups).  Afterwards the same warps run the epilogue:
//              tcgen05.ld gives every thread all N channels of one pixel.
//   warp 4     B producer: one elected lane streams the pre-tiled weights of each (chunk, tap) with
//              cp.async.bulk (TMA 1-D) into an NB-stage ring; also owns the TMEM allocation.
//   warp 5     MMA issuer: one elected lane, 4 (k16) x 2 (M tiles) tcgen05.mma per (chunk, tap),
//              tcgen05.commit hands smem stages back to the producers and the accumulators to the epilogue.
#include "update_common.cuh"

namespace cer {

constexpr int TC_TH = 16, TC_TW = 16;
constexpr int TC_HW = TC_TW + 2, TC_HH = TC_TH + 2;
constexpr int TC_HPX = TC_HW * TC_HH;          // 324 halo pixels
constexpr int TC_A_LBO = TC_HPX * 16;          // bytes between the two 8-channel groups of one K=16 slice
constexpr int TC_A_SBO = TC_HW * 16;           // bytes between 8-row groups (= image rows of the M tile)
constexpr int TC_A_BYTES = 8 * TC_A_LBO;       // one 64-channel chunk: 41 472 B
constexpr int TC_THREADS = 192;

template <int N>
struct TcCfg {
  static constexpr int NB = (N == 256) ? 3 : 4;
  static constexpr int B_BYTES = 64 * N * 2;
  static constexpr int TMEM_COLS = (2 * N <= 128) ? 128 : (2 * N <= 256 ? 256 : 512);
  static constexpr int OFF_B = 2 * TC_A_BYTES;
  static constexpr int OFF_EXTRA = OFF_B + NB * B_BYTES;                  // DELTA: w2 [9][256] f32 + bias [256] f32
  static constexpr int EXTRA_BYTES = (N == 256) ? (9 * 256 + 256) * 4 : 0;
  static constexpr int OFF_BAR = OFF_EXTRA + EXTRA_BYTES;                 // 8-byte aligned
  static constexpr int NUM_BAR = 5 + 2 * NB;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout type 0 [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=B=f16 (0), K-major both,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void st_half32(__half* dst, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 pk;
    __half2 h0 = __floats2half2_rn(v[8 * q + 0], v[8 * q + 1]);
    __half2 h1 = __floats2half2_rn(v[8 * q + 2], v[8 * q + 3]);
    __half2 h2 = __floats2half2_rn(v[8 * q + 4], v[8 * q + 5]);
    __half2 h3 = __floats2half2_rn(v[8 * q + 6], v[8 * q + 7]);
    pk.x = *reinterpret_cast<uint32_t*>(&h0);
    pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2);
    pk.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(dst + 8 * q) = pk;
  }
}
__device__ __forceinline__ void ld_half32(const __half* src, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 pk = *reinterpret_cast<const uint4*>(src + 8 * q);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&pk.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&pk.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&pk.z));
    const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&pk.w));
    v[8 * q + 0] = a.x; v[8 * q + 1] = a.y; v[8 * q + 2] = b.x; v[8 * q + 3] = b.y;
    v[8 * q + 4] = c.x; v[8 * q + 5] = c.y; v[8 * q + 6] = d.x; v[8 * q + 7] = d.y;
  }
}

// ---- the kernel -----------------------------------------------------------------------------------
template <int N, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_tc_kernel(const ConvArgs a) {
  using C = TcCfg<N>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sA = s0, sB = s0 + C::OFF_B, sBar = s0 + C::OFF_BAR;
  // barrier slots
  auto bar_a_full = [&](int i) { return sBar + 8 * i; };
  auto bar_a_empty = [&](int i) { return sBar + 8 * (2 + i); };
  auto bar_b_full = [&](int i) { return sBar + 8 * (4 + i); };
  auto bar_b_empty = [&](int i) { return sBar + 8 * (4 + C::NB + i); };
  const uint32_t bar_acc = sBar + 8 * (4 + 2 * C::NB);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_x = (a.w + TC_TW - 1) / TC_TW;
  const int x0 = (blockIdx.x % tiles_x) * TC_TW, y0 = (blockIdx.x / tiles_x) * TC_TH;
  const int n_steps = a.n_src * 9;

  if (tid == 0) {
    mbar_init(bar_a_full(0), 128);
    mbar_init(bar_a_full(1), 128);
    mbar_init(bar_a_empty(0), 1);
    mbar_init(bar_a_empty(1), 1);
    for (int i = 0; i < C::NB; ++i) {
      mbar_init(bar_b_full(i), 1);
      mbar_init(bar_b_empty(i), 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(s0 + C::OFF_TMEM), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (EPI == EPI_DELTA && warp < 4) {
    float* ex = reinterpret_cast<float*>(smem + C::OFF_EXTRA);
    for (int i = tid; i < 9 * 256; i += 128) ex[i] = __ldg(a.w2 + i);
    for (int i = tid; i < 256; i += 128) ex[9 * 256 + i] = __ldg(a.bias + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ================= A producers =================
    for (int c = 0; c < a.n_src; ++c) {
      const int buf = c & 1;
      mbar_wait(bar_a_empty(buf), ((c >> 1) & 1) ^ 1);
      const __half* src = a.src[c];
      const uint32_t dst0 = sA + buf * TC_A_BYTES;
      for (int i = tid; i < TC_HPX * 8; i += 128) {
        const int hp = i >> 3, g = i & 7;
        const int yy = y0 - 1 + hp / TC_HW, xx = x0 - 1 + hp % TC_HW;
        const bool ok = yy >= 0 && yy < a.h && xx >= 0 && xx < a.w;
        const __half* gp = src + ((long long)(ok ? yy : 0) * a.w + (ok ? xx : 0)) * 64 + g * 8;
        cp_async16_zfill(dst0 + g * TC_A_LBO + hp * 16, gp, ok);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy (MMA) reads
      mbar_arrive(bar_a_full(buf));
    }
  } else if (warp == 4) {
    // ================= B producer =================
    if (lane == 0) {
      const char* wsrc = reinterpret_cast<const char*>(a.wtc);
      for (int s = 0; s < n_steps; ++s) {
        const int st = s % C::NB;
        mbar_wait(bar_b_empty(st), ((s / C::NB) & 1) ^ 1);
        mbar_expect_tx(bar_b_full(st), C::B_BYTES);
        bulk_g2s(sB + st * C::B_BYTES, wsrc + (size_t)s * C::B_BYTES, C::B_BYTES, bar_b_full(st));
      }
    }
  } else {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(128, N);
      for (int s = 0; s < n_steps; ++s) {
        const int chunk = s / 9, tap = s % 9, buf = chunk & 1, st = s % C::NB;
        if (tap == 0) mbar_wait(bar_a_full(buf), (chunk >> 1) & 1);
        mbar_wait(bar_b_full(st), (s / C::NB) & 1);
        tc_fence_after();
        const int ky = tap / 3, kx = tap % 3;
        const uint32_t a0 = sA + buf * TC_A_BYTES + (ky * TC_HW + kx) * 16;
        const uint32_t b0 = sB + st * C::B_BYTES;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint64_t bd = umma_desc(b0 + 2 * k16 * (N * 16), N * 16, 128);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t ad = umma_desc(a0 + 2 * k16 * TC_A_LBO + j * 8 * 16, TC_A_LBO, TC_A_SBO);
            tc_mma_f16(tmem_base + j * N, ad, bd, idesc, (s > 0 || k16 > 0) ? 1u : 0u);
          }
        }
        tc_commit(bar_b_empty(st));                 // smem stage free once these MMAs have read it
        if (tap == 8) tc_commit(bar_a_empty(buf));
      }
      tc_commit(bar_acc);                           // accumulators complete
    }
  }

  if (warp < 4) {
    // ================= epilogue: thread = one pixel of each M tile, all N channels =================
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int m = warp * 32 + lane;                 // row of the M tile = TMEM lane
    const int r = m >> 3, cc = m & 7;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      const int yy = y0 + r, xx = x0 + j * 8 + cc;
      const bool ok = yy < a.h && xx < a.w;
      const long long p = ok ? (long long)yy * a.w + xx : 0;
      float t9[9];
      if (EPI == EPI_DELTA) {
#pragma unroll
        for (int t = 0; t < 9; ++t) t9[t] = 0.f;
      }
#pragma unroll 1
      for (int cb = 0; cb < N / 32; ++cb) {
        uint32_t raw[32];
        tc_ld32(lane_addr + j * N + cb * 32, raw);     // warp-collective: executed by every lane
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(raw[e]);
        const int n0 = cb * 32;
        if (EPI == EPI_RELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(h_round(v[e] + __ldg(a.bias + n0 + e)), 0.f);
          if (ok) st_half32(a.out_h + p * 64 + n0, v);
        } else if (EPI == EPI_GATES) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] += __ldg(a.bias + n0 + e);
          if (n0 < 64) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = sigmoid_f(h_round(v[e]));
            if (ok) st_half32(a.z + p * 64 + n0, v);
          } else if (n0 < 128) {
            if (ok) {
              float nt[32];
              ld_half32(a.net + p * 64 + (n0 - 64), nt);
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = h_round(sigmoid_f(h_round(v[e]))) * nt[e];
              st_half32(a.rnet + p * 64 + (n0 - 64), v);
            }
          } else if (ok) {
            float4* q = reinterpret_cast<float4*>(a.qx + p * 64 + (n0 - 128));
#pragma unroll
            for (int e = 0; e < 8; ++e) q[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
          }
        } else if (EPI == EPI_GRUOUT) {
          if (ok) {
            float zz[32], nt[32];
            ld_half32(a.z + p * 64 + n0, zz);
            ld_half32(a.net + p * 64 + n0, nt);
            const float4* q = reinterpret_cast<const float4*>(a.qx + p * 64 + n0);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 qq = q[e];
              v[4 * e] += qq.x; v[4 * e + 1] += qq.y; v[4 * e + 2] += qq.z; v[4 * e + 3] += qq.w;
            }
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float qv = h_round(tanhf(h_round(v[e])));
              v[e] = h_round(h_round(h_round(1.f - zz[e]) * nt[e]) + h_round(zz[e] * qv));
            }
            st_half32(a.net + p * 64 + n0, v);
          }
        } else {  // EPI_DELTA
          const float* ex = reinterpret_cast<const float*>(smem + C::OFF_EXTRA);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(h_round(v[e] + ex[9 * 256 + n0 + e]), 0.f);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            float acc = t9[t];
#pragma unroll
            for (int e = 0; e < 32; ++e) acc = fmaf(v[e], ex[t * 256 + n0 + e], acc);
            t9[t] = acc;
          }
        }
      }
      if (EPI == EPI_DELTA && ok) {
#pragma unroll
        for (int t = 0; t < 9; ++t) a.s9[p * 9 + t] = t9[t];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ---- host ----------------------------------------------------------------------------------------
template <int N, int EPI>
static int tc_configure_one() {
  CER_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                TcCfg<N>::TOTAL));
  return CER_OK;
}

int tc_configure() {
  int rc;
  if ((rc = tc_configure_one<64, EPI_RELU>())) return rc;
  if ((rc = tc_configure_one<192, EPI_GATES>())) return rc;
  if ((rc = tc_configure_one<64, EPI_GRUOUT>())) return rc;
  if ((rc = tc_configure_one<256, EPI_DELTA>())) return rc;
  return CER_OK;
}

template <int N, int EPI>
int launch_conv_tc(const ConvArgs& a, cudaStream_t stream) {
  const int tiles = ((a.w + TC_TW - 1) / TC_TW) * ((a.h + TC_TH - 1) / TC_TH);
  constexpr int kind = EPI == EPI_RELU ? KK_CONV_E : EPI == EPI_GATES ? KK_CONV_GATES : EPI == EPI_GRUOUT ? KK_CONV_Q : KK_CONV_DELTA;
  CER_LAUNCH(kind, (conv3x3_tc_kernel<N, EPI>), tiles, TC_THREADS, TcCfg<N>::TOTAL, stream, a);
  return check_launch("conv3x3_tc");
}

int launch_conv_tc_dispatch(int n, int epi, const ConvArgs& a, cudaStream_t stream) {
  if (n == 64 && epi == EPI_RELU) return launch_conv_tc<64, EPI_RELU>(a, stream);
  if (n == 192 && epi == EPI_GATES) return launch_conv_tc<192, EPI_GATES>(a, stream);
  if (n == 64 && epi == EPI_GRUOUT) return launch_conv_tc<64, EPI_GRUOUT>(a, stream);
  if (n == 256 && epi == EPI_DELTA) return launch_conv_tc<256, EPI_DELTA>(a, stream);
  set_error("conv3x3_tc: unsupported configuration N=%d epilogue=%d", n, epi);
  return CER_ERR_INVALID;
}

}  // namespace cer
