// UpdateBlock 3x3 convolutions, v3: implicit GEMM on tcgen05.mma (5th-gen tensor cores), accumulators
// in TMEM.  Same math, inputs, outputs and fused epilogues as update_hmma.cu.
//
// Persistent CTAs (one per SM; the two wide convs as cta_group::2 CTA pairs by default) walk over 16 x 8 pixel tiles
// (M = 128 per CTA), N = 64 / 192 / 256 output channels, K = n_src x 9 taps x 64 channels; the TMEM accumulator is
// double-buffered so the epilogue of tile t overlaps the MMAs of tile t+1.  Warp roles (512 threads for the gate conv,
// 416 for the others):
//
//   warps 0-7    epilogue: tcgen05.ld gives every thread 32 channels of one pixel (TMEM lane = pixel; warps w and w+4
//                split the columns); bias / sigmoid / tanh / GRU blend / delta dots, 256-bit NHWC loads and stores;
//                then releases the accumulator.
//   warp 8       MMA issuer: one elected lane, 4 (k16) tcgen05.mma per (chunk, tap); tcgen05.commit hands smem stages
//                back to the producers and accumulators to the epilogue.  Owns the TMEM allocation.
//   warp 9       B producer: lane 0 streams the pre-tiled weights with cp.async.bulk / TMA into an mbarrier ring (N = 64
//                and the paired delta conv keep all 9 taps resident); lane 1 publishes tile flags when they are enabled.
//   warps 10-12  A producers: per 64-channel chunk ONE 4-D TMA tensor load brings the 18 x 10 halo tile in (zero fill
//                outside the image = the conv padding) in the UMMA K-major no-swizzle layout
//                [k-group 8][halo pixel 180][8 halfs]; a 3x3 tap is then just a shifted start address of the same tile
//                (LBO = 180*16 B between k-groups, SBO = 10*16 B between image rows = 8-row core-matrix groups).
//                (cer_set_conv_a_tma(0): the same tile by 16-byte cp.async from these three warps.)
//   warps 13-15  gate conv only: the disparity-encoder chunk (core/update.py:80-85,97) is computed straight from disp
//                into the A ring instead of being read from HBM.
#include <string.h>

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "tc_common.cuh"
#include "update_common.cuh"

namespace cer {

constexpr int TC_TH = 16, TC_TW = 8;            // one M = 128 tile
constexpr int TC_HW = TC_TW + 2, TC_HH = TC_TH + 2;
constexpr int TC_HPX = TC_HW * TC_HH;          // 180 halo pixels
constexpr int TC_A_LBO = TC_HPX * 16;          // bytes between the two 8-channel groups of one K=16 slice
constexpr int TC_A_SBO = TC_HW * 16;           // bytes between 8-row groups (= image rows of the M tile)
constexpr int TC_A_BYTES = 8 * TC_A_LBO;       // one 64-channel chunk: 23 040 B
constexpr int TC_PROD = 96;                    // A producer threads (80 active: one per (halo column, k-group))
constexpr int TC_W_MMA = 8, TC_W_BPROD = 9, TC_W_APROD = 10;   // warp roles; warps 0-7 are the epilogue
constexpr int TC_W_MMA2 = TC_W_APROD + TC_PROD / 32;            // warp 13: dn producer (or second MMA issuer, see kTwoIssuers)
constexpr int TC_W_DN = TC_W_MMA2;                               // warps 13-15: dn producers
constexpr int TC_DN = 96;                                        // dn producer threads (== TC_PROD: same arrival count)
constexpr int TC_THREADS = TC_W_DN * 32 + TC_DN;                 // 512
// only the gate conv has a dn chunk: the other convs are launched without the three dn warps (416 threads), which frees
// three idle warps' worth of registers and scheduler slots
__host__ __device__ constexpr int tc_threads(int epi) { return epi == EPI_GATES ? TC_THREADS : TC_W_DN * 32; }
// Two issuer warps taking alternate (chunk, tap) steps buy ~5 % on the gate conv but make the fp32 accumulation order
// (and therefore the last bit of some outputs) depend on how the two warps interleave: off, results are bit-reproducible.
constexpr bool kTwoIssuers = false;      // (warp 13 generates the dn chunk instead)
constexpr int TC_DT_H = TC_HH + 6, TC_DT_W = TC_HW + 6;   // disparity tile for the 7x7 encoder: 24 x 16

// CG2: the CTA pair of a 2-CTA cluster runs one M=256 tcgen05.mma.cta_group::2 per step: each CTA supplies its own
// 128 pixel rows of A and HALF of the weight tile (N/2 output channels), which halves the weight traffic from L2
// and the shared-memory operand reads per SM -- the limiter of the 1-CTA form.
// MC2: a 2-CTA cluster of ordinary 1-CTA MMAs in which each CTA fetches HALF of every weight tile and multicasts it
// into both CTAs' shared memory (cp.async.bulk ... .multicast::cluster): the weight traffic out of L2 -- 885 KB per
// 128-pixel tile for the gate conv, the measured limiter -- is halved with no cross-CTA software signalling at all
// (stage release = tcgen05.commit multicast to both CTAs' barriers).
// MT2: one CTA owns TWO M = 128 tiles and issues both MMAs against every weight stage, so the weight bytes an SM
// has to pull through its ~65 GB/s L2 port per output pixel are halved (that port, not the tensor pipe, bounds the
// streamed-weight convs); the accumulators then fill TMEM (2 x N columns), so the epilogue is not double-buffered.
// S3: the 1-CTA form with a whole kernel row (3 taps, 72 KB) per weight stage, two stages: the issuing thread pays one
// barrier wait + one commit per 12 MMAs (1 150 tensor cycles) instead of per 4 -- its per-step cost (~500 cycles of waits,
// fences and commits around 384 cycles of MMA work) is what the role profile shows on the gate conv.  N = 192 only
// (two 3-tap stages of the N = 256 conv do not fit in shared memory).
// HALF (delta conv): CTAs come in (even, odd) pairs that are NOT a cluster: both walk the same tiles, each computes half
// of the 256 output channels (template N = 128) with ITS half of the weights resident in shared memory (9 x 16 KB), so a
// tile costs one run of 36 M128 x N128 MMAs and nothing is streamed (the 1-CTA N = 256 form pulls 295 KB of weights per
// tile through L2 and pays a barrier wait + commit per 4 MMAs: 790 cycles of issue for 512 cycles of tensor work).
enum TcMode { TC_SINGLE = 0, TC_CG2 = 1, TC_MC2 = 2, TC_MT2 = 3, TC_S3 = 4, TC_HALF = 5 };

template <int N, int MODE = TC_SINGLE>
struct TcCfg {
  static constexpr bool CG2 = MODE == TC_CG2;
  static constexpr bool MC2 = MODE == TC_MC2;
  static constexpr int MT = MODE == TC_MT2 ? 2 : 1;             // M tiles per CTA work unit
  // all 9 weight tiles stay in smem: the N = 64 convs, and the delta conv (N = 256) as a CTA pair -- each CTA's half of
  // its 288 KB weight set is 144 KB, exactly the three 3-tap stages the pair form has room for, so the pair never
  // re-streams weights (the 1-CTA form pulls 295 KB per 128-pixel tile through its L2 port for only 36 MMAs)
  static constexpr bool HALF = MODE == TC_HALF;
  static_assert(!HALF || N == 128, "the half form computes 128 of the delta conv's 256 channels per CTA");
  static constexpr bool RESIDENT = (N == 64) || (MODE == TC_CG2 && N == 256) || HALF;
  // taps per weight stage: the CTA-pair form moves a whole kernel row (3 taps) per stage so that the issuing thread
  // pays one barrier wait + one commit per 12 MMAs instead of per 4 (its per-step cost, not the tensor pipe, bounds
  // the streamed-weight convs); the pair's halved weight footprint is what makes room for it
  static constexpr bool S3 = MODE == TC_S3;
  static_assert(!S3 || N == 192, "3-tap stages of the 1-CTA form are sized for the gate conv");
  // resident weights (N = 64): nothing to wait for between taps, so all nine go out in one run of 36 MMAs
  static constexpr int TPS = ((N == 64 && !CG2) || HALF) ? 9 : (CG2 || S3) ? 3 : 1;
  static constexpr int NG = 9 / TPS;                          // stages per 64-channel chunk
  static constexpr int NB = RESIDENT ? NG : S3 ? 2 : (CG2 ? (N == 256 ? 3 : 4) : (MT == 2 ? (N == 256 ? 3 : 5) : ((N == 256) ? 4 : 6)));
  static constexpr int NLOC = CG2 ? N / 2 : N;                // weight rows held by this CTA
  // A ring depth (64-channel chunks).  Deeper rings / more accumulator stages for the N = 64 convs were measured
  // slower (q/GRU 36 -> 44 us): those kernels are bound by their per-tile latency chain, not by ring capacity.
  static constexpr int NA = MT == 2 ? 4 : 3;
  static constexpr int NACC = MT == 2 ? 1 : 2;                    // TMEM accumulator stages (of MT x N columns)
  static constexpr int B_BYTES = 64 * NLOC * 2;               // one tap
  static constexpr int STAGE_BYTES = TPS * B_BYTES;
  static constexpr int TMEM_COLS = (NACC * MT * N <= 128) ? 128 : (NACC * MT * N <= 256 ? 256 : 512);
  static constexpr int OFF_B = NA * TC_A_BYTES;
  static constexpr int OFF_EXTRA = OFF_B + NB * STAGE_BYTES;                  // DELTA: bias [256] f32 + w2 [9][256] f16; GATES: disparity tile
  // HALF: + the 9 partial dots of 128 pixels handed from the warps of column half 1 to those of column half 0
  static constexpr int EXTRA_BYTES = (N == 256 || HALF) ? 256 * 4 + 9 * (256 + 8) * 2 + (HALF ? 128 * 9 * 4 : 0)
                                                       : (N == 192 ? TC_DT_H * TC_DT_W * 4 : 0);
  // N = 64: staging tile of the q/GRU epilogue's element-wise operands (z, net: 16 x 8 pixels x 128 B; q's x-part: x 256 B),
  // written by TMA one tile ahead with the 128-byte swizzle (1024-byte aligned), read conflict-free by thread = pixel
  static constexpr int OFF_GRU = (OFF_EXTRA + EXTRA_BYTES + 1023) / 1024 * 1024;
  static constexpr int GRU_BYTES = (N == 64 && MODE == TC_SINGLE) ? 128 * (128 + 128 + 256) : 0;
  static constexpr int OFF_BAR = GRU_BYTES ? OFF_GRU + GRU_BYTES : OFF_EXTRA + EXTRA_BYTES;                 // 8-byte aligned
  static constexpr int NUM_BAR = 2 * NA + 3 * NB + 2 * NACC + 2;  // a_full/empty, b_full/empty/peer_full, acc_full/empty, gru_full/empty
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
};

// Blackwell mixed-precision FMA (PTX fma.rn.f32.f16 -> SASS FHFMA): exact fp16 x fp16 product, fp32 accumulate
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

// ---- tile-level dependencies between consecutive convs (ConvArgs::flags_in / flags_out) ----
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_cta_shared_inc(uint32_t addr) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// Whole warp: wait for one tile's flag (lane 0 polls).
__device__ __forceinline__ void wait_tile_flag(const int* flag, int lane) {
  if (lane == 0) {
    unsigned spins = 0;
    while (ld_acquire_gpu(flag) == 0) {
      __nanosleep(200);
      if (++spins > (1u << 23)) __trap();
    }
  }
  __syncwarp();
}
// Whole warp: lanes 0..8 each watch one tile of the 3 x 3 neighbourhood of `tile` (tiles outside the grid count as done).
// The upstream grid never waits for this one, so the wait always ends; the spin limit only turns a programming error
// (a flag that is never published) into a trap instead of a hung GPU.
__device__ __forceinline__ void wait_tiles3x3(const int* flags, int tile, int tiles_x, int n_tiles, int lane) {
  const int tx = tile % tiles_x, ty = tile / tiles_x, tiles_y = n_tiles / tiles_x;
  const int nx = tx + lane % 3 - 1, ny = ty + lane / 3 - 1;
  const bool watch = lane < 9 && nx >= 0 && nx < tiles_x && ny >= 0 && ny < tiles_y;
  const int* f = flags + (watch ? ny * tiles_x + nx : 0);
  unsigned spins = 0;
  while (true) {
    const int v = watch ? ld_acquire_gpu(f) : 1;
    if (__all_sync(0xffffffffu, v != 0)) break;
    __nanosleep(200);
    if (++spins > (1u << 23)) __trap();
  }
}

// tcgen05.ld 16x256b.x2: 16 TMEM lanes (rows) x 16 fp32 columns, delivered per 8-column block in the mma.sync m16n8
// accumulator layout (lane = 4 g + q: r0,r1 = row g, cols 2q,2q+1; r2,r3 = row g+8; r4..r7 = the same for cols +8), which
// after packing to fp16 pairs IS the m16n8k16 A fragment (a0 = r0:r1, a1 = r2:r3, a2 = r4:r5, a3 = r6:r7).
__device__ __forceinline__ void tc_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void mma16816_f32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Delta head (second delta conv as 9 per-pixel dots over 256 channels): 1 = on mma.sync straight from the TMEM
// accumulator fragments (32 HMMA per warp and tile), 0 = 288 FHFMA per thread and 32-channel block (the round's earlier
// form, kept for A/B: the delta conv was bound by this epilogue's issue slots, not by its MMAs).
constexpr int kDeltaHeadMma = 1;
constexpr int kW2Pitch = 256 + 8;      // halfs per tap row of the delta.2 weights in shared memory (conflict-free B fragments)

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  // 1 - 2/(e^{2x}+1): abs error ~1e-7, far below the fp16 rounding that follows; saturates cleanly
  return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f);
}

// 256-bit stores: one full 32-byte sector per request (a thread owns 64 contiguous bytes of fp16 or 128 of fp32)
__device__ __forceinline__ void st256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void st_half32(__half* dst, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t pk[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const __half2 h = __floats2half2_rn(v[16 * q + 2 * e], v[16 * q + 2 * e + 1]);
      pk[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    st256(dst + 16 * q, pk);
  }
}
__device__ __forceinline__ void st_float32(float* dst, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t pk[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) pk[e] = __float_as_uint(v[8 * q + e]);
    st256(dst + 8 * q, pk);
  }
}
// Activations are re-written every iteration, and with tile flags a consumer CTA no longer passes a grid-level
// dependency wait (which is what invalidates L1): read them with ld.global.cg (L2, the coherence point) so a line this
// SM cached an iteration ago can never be served.  256-bit loads: one request per 32-byte sector, which is what the
// L1-allocating 128-bit loads amounted to after their sector fill (128-bit .cg loads were measured 10 % slower).
__device__ __forceinline__ void ldcg256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void ldg_half32_raw(const __half* src, uint4 (&pk)[4]) {     // 32 halfs, 64-byte aligned
  uint32_t a[8], b[8];
  ldcg256(src, a);
  ldcg256(src + 16, b);
  pk[0] = make_uint4(a[0], a[1], a[2], a[3]);
  pk[1] = make_uint4(a[4], a[5], a[6], a[7]);
  pk[2] = make_uint4(b[0], b[1], b[2], b[3]);
  pk[3] = make_uint4(b[4], b[5], b[6], b[7]);
}
__device__ __forceinline__ void unpack_half32(const uint4 (&pk)[4], float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&pk[q].x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&pk[q].y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&pk[q].z));
    const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&pk[q].w));
    v[8 * q + 0] = a.x; v[8 * q + 1] = a.y; v[8 * q + 2] = b.x; v[8 * q + 3] = b.y;
    v[8 * q + 4] = c.x; v[8 * q + 5] = c.y; v[8 * q + 6] = d.x; v[8 * q + 7] = d.y;
  }
}
__device__ __forceinline__ void ld_half32(const __half* src, float (&v)[32]) {
  uint4 pk[4];
  ldg_half32_raw(src, pk);
  unpack_half32(pk, v);
}
__device__ __forceinline__ void ld_float32_add(const float* src, float (&v)[32]) {          // v += 32 floats, 128-byte aligned
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r[8];
    ldcg256(src + 8 * q, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[8 * q + e] += __uint_as_float(r[e]);
  }
}

// TMA descriptors of the (up to four) NHWC fp16 source tensors, viewed as 4-D (8 channels, x, y, 8 channel groups) so that
// ONE tensor load of the box (8, 10, 18, 8) writes the whole 18 x 10-pixel halo chunk in the UMMA K-major layout
// [k-group][halo pixel][8 halfs]; pixels outside the image come back as zeros, which is the conv padding.
struct alignas(64) TcMaps {
  CUtensorMap m[4];
  int use_tma;
};

// ---- the kernel -----------------------------------------------------------------------------------
template <int N, int EPI, int MODE>
__global__ void __launch_bounds__(tc_threads(EPI), 1) conv3x3_tc_kernel(const ConvArgs a,
                                                                  const __grid_constant__ CUtensorMap wmap,
                                                                  const __grid_constant__ TcMaps amaps) {
  using C = TcCfg<N, MODE>;
  constexpr bool CG2 = C::CG2, MC2 = C::MC2;
  constexpr int MT = C::MT;
  static_assert(MT == 1 || (!C::RESIDENT && N * 2 <= 512), "MT2 is for the streamed-weight convs");
  static_assert(!(MC2 && C::RESIDENT), "resident weights are only used by the 1-CTA form");
  static_assert(!CG2 || C::NB >= C::NA, "the CTA pair reuses bar_b_peer(0..NA) as peer-A-full barriers");
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sA = s0, sB = s0 + C::OFF_B, sBar = s0 + C::OFF_BAR;
  auto bar_a_full = [&](int i) { return sBar + 8 * i; };
  auto bar_a_empty = [&](int i) { return sBar + 8 * (C::NA + i); };
  auto bar_b_full = [&](int i) { return sBar + 8 * (2 * C::NA + i); };
  auto bar_b_empty = [&](int i) { return sBar + 8 * (2 * C::NA + C::NB + i); };
  auto bar_b_peer = [&](int i) { return sBar + 8 * (2 * C::NA + 2 * C::NB + i); };      // CG2, leader only
  auto bar_acc_full = [&](int i) { return sBar + 8 * (2 * C::NA + 3 * C::NB + i); };
  auto bar_acc_empty = [&](int i) { return sBar + 8 * (2 * C::NA + 3 * C::NB + C::NACC + i); };
  const uint32_t bar_gru_full = sBar + 8 * (2 * C::NA + 3 * C::NB + 2 * C::NACC), bar_gru_empty = bar_gru_full + 8;
  constexpr bool GRU_TMA = EPI == EPI_GRUOUT && C::GRU_BYTES > 0;
  const uint32_t rank = (CG2 || MC2) ? cluster_ctarank() : 0u;      // CG2: 0 = leader (issues the MMAs)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();      // let the next kernel of the chain get scheduled as SMs drain (it waits on pdl_wait itself)
  const int tiles_x = (a.w + TC_TW - 1) / TC_TW;
  const int n_tiles = tiles_x * ((a.h + TC_TH - 1) / TC_TH);
  // every CTA runs the same number of tile iterations (a CTA pair must stay in lock step); surplus tile indices
  // lie below the image: all loads zero-fill, nothing is stored
  const int n_units = (n_tiles + MT - 1) / MT;                   // a work unit = MT consecutive tiles
  constexpr bool HALF = C::HALF;
  const int nhalf = HALF ? (int)(blockIdx.x & 1) : 0;            // which 128 of the 256 output channels
  const int cta_i = HALF ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, cta_n = HALF ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_iter = (n_units + cta_n - 1) / cta_n;
  const int n_src = a.n_src;
  // Every CTA (pair) walks the (chunk, tap) sum in its own rotated order so that the 148 SMs do not all pull the
  // same weight tile out of L2 at the same moment (the accumulation order is free).
  const int rot_id = (CG2 || MC2) ? blockIdx.x / 2 : blockIdx.x;      // a pair consumes weight stages in lock step
  const int rot_g = rot_id % C::NG, rot_chunk = (rot_id / C::NG) % n_src;

  if (tid == 0) {
    for (int i = 0; i < C::NA; ++i) {
      mbar_init(bar_a_full(i), TC_PROD);   // local producers; CG2: the peer relays its completions to bar_b_peer(i)
      mbar_init(bar_a_empty(i), kTwoIssuers ? 2 : 1);      // one tcgen05.commit per MMA issuer warp
    }
    for (int i = 0; i < C::NB; ++i) {
      mbar_init(bar_b_full(i), 1);
      mbar_init(bar_b_empty(i), MC2 ? 2 : 1);      // MC2: both CTAs must have consumed a stage before it is refilled
      mbar_init(bar_b_peer(i), 1);
    }
    for (int i = 0; i < C::NACC; ++i) {
      mbar_init(bar_acc_full(i), kTwoIssuers ? 2 : 1);
      mbar_init(bar_acc_empty(i), CG2 ? 512 : 256);
    }
    mbar_init(bar_gru_full, 1);
    mbar_init(bar_gru_empty, 256);
    *reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM + 8) = 0u;     // epilogue warps that finished a tile (flags_out)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_W_MMA) {
    if (CG2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(s0 + C::OFF_TMEM), "r"((uint32_t)C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(s0 + C::OFF_TMEM), "r"((uint32_t)C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (EPI == EPI_DELTA && warp < 8) {
    // delta head constants: bias f32 [256], then the 3x3 -> 1 weights as fp16 [9][256] (the packed values are
    // fp16-representable, so the conversion is exact)
    float* exb = reinterpret_cast<float*>(smem + C::OFF_EXTRA);
    __half* exw = reinterpret_cast<__half*>(smem + C::OFF_EXTRA + 256 * 4);
    for (int i = tid; i < 256; i += 256) exb[i] = __ldg(a.bias + i);
    for (int i = tid; i < 9 * 256; i += 256) exw[(i >> 8) * kW2Pitch + (i & 255)] = __float2half_rn(__ldg(a.w2 + i));
  }
  tc_fence_before();
  __syncthreads();
  if (CG2 || MC2) cluster_sync_all();       // barriers of both CTAs are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // optional role profiling (a.prof != null): cycles CTA 0 spends in each barrier wait, per role
  long long prof_acc[4] = {0, 0, 0, 0};
  const bool prof_on = a.prof != nullptr && blockIdx.x == 0 && lane == 0;
  const long long prof_t0 = clock64();
  auto pwait = [&](uint32_t bar, uint32_t parity, int slot) {
    if (prof_on) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      prof_acc[slot] += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };
  // arrive on a barrier that lives in the leader CTA
  auto arrive_leader = [&](uint32_t bar) {
    if (!CG2 || rank == 0) mbar_arrive(bar);
    else mbar_arrive_cluster(bar, 0);
  };

  // Everything above (barriers, TMEM, delta-head constants) and the weight stream below touch only constants; the
  // activations belong to the previous kernel of the chain: wait for it here (PDL), except in the weight producer.
  // With tile flags the roles below wait per tile for the 3 x 3 neighbourhood of the upstream conv instead; everything
  // else they read was complete before the upstream grid passed its own dependency wait.
  const bool dep_flags = a.flags_in != nullptr;
  if (warp != TC_W_BPROD && !dep_flags) pdl_wait();

  if (warp >= TC_W_APROD && warp < TC_W_MMA2) {
    // ================= A producers =================
    const int pt = tid - TC_W_APROD * 32;
    const int p_hx = pt >> 3, p_g = pt & 7;          // this thread's halo column and k-group (pt < 80)
    int seq = 0;           // chunk sequence number over all tiles of this CTA
    // A stage is published by the hardware: every producer thread attaches a cp.async.mbarrier.arrive.noinc to its
    // copies, so the producers never wait for (or fence) their own copies and run ahead of the MMA by the ring depth.
    // (A producer-side wait_group + fence.proxy.async serialises on the copies of the NEXT chunk too: measured, the
    // producers then became the critical path of the CTA-pair form.)  The MMA warp issues the generic->async proxy
    // fence after it has observed the barrier.
    for (int it = 0; it < n_iter; ++it) {
      const int tile0 = (it * cta_n + cta_i) * MT;
      for (int cj = 0; cj < n_src * MT; ++cj, ++seq) {
        const int ci = cj / MT, tile = tile0 + cj % MT;          // chunk-major: (chunk 0: tile 0, tile 1), (chunk 1: ...)
        const int x0 = (tile % tiles_x) * TC_TW, y0 = (tile / tiles_x) * TC_TH;
        const int c = (ci + rot_chunk) % n_src;
        const int st = seq % C::NA;
        if (dep_flags && ci == 0 && tile < n_tiles) wait_tiles3x3(a.flags_in, tile, tiles_x, n_tiles, lane);
        pwait(bar_a_empty(st), ((seq / C::NA) & 1) ^ 1, 0);
        const uint32_t dst0 = sA + st * TC_A_BYTES;
        const long long t_fill0 = prof_on ? clock64() : 0;
        if (c == a.dn_chunk) {
          continue;       // generated by the dn warp below; the wait above keeps this warp in step with the ring phases
        } else if (amaps.use_tma) {
          // one thread hands the whole chunk to the TMA unit (23 KB, zero-filled outside the image); the other producer
          // threads only keep the barrier's arrival count
          if (pt == 0) {
            mbar_expect_tx(bar_a_full(st), TC_A_BYTES);
            tma4d(dst0, &amaps.m[c], 0, x0 - 1, y0 - 1, 0, bar_a_full(st));
          } else {
            mbar_arrive(bar_a_full(st));
          }
          if (prof_on) prof_acc[3] += clock64() - t_fill0;
          continue;
        } else if (p_hx < TC_HW) {
          // one halo column x one k-group per thread, walking down the 18 halo rows: 2 adds per copy
          const int xx = x0 - 1 + p_hx;
          const bool xok = xx >= 0 && xx < a.w;
          const __half* gp = a.src[c] + ((long long)(y0 - 1) * a.w + (xok ? xx : 0)) * 64 + p_g * 8;
          uint32_t dst = dst0 + p_g * TC_A_LBO + p_hx * 16;
#pragma unroll 6
          for (int hy = 0; hy < TC_HH; ++hy) {
            const int yy = y0 - 1 + hy;
            const bool ok = xok && yy >= 0 && yy < a.h;
            cp_async16_zfill(dst, ok ? gp : a.src[c], ok);
            gp += (long long)a.w * 64;
            dst += TC_A_SBO;
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_a_full(st)) : "memory");
        if (prof_on) prof_acc[3] += clock64() - t_fill0;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp >= TC_W_DN) {
    // ================= dn producers (gate conv only) =================
    // disparity-neighbourhood encoder (core/update.py:41-49), generated straight into the A ring: channel
    // k = 100 * (disp(y + k/7 - 3, x + k%7 - 3) - disp(y, x)), zero padded.  Three warps of their own, off the copy
    // warps' path: they stage the 24 x 16 disparity tile, then each thread writes the 8 k-groups of its halo pixels.
    if (a.dn_chunk >= 0) {
      static_assert(TC_DN == TC_PROD, "the dn threads stand in for the copy threads' arrivals on a_full");
      float* sD = reinterpret_cast<float*>(smem + C::OFF_EXTRA);
      constexpr int DT = TC_DT_H * TC_DT_W, DT_PER = DT / TC_DN;   // 384 / 96 = 4 values per thread
      static_assert(DT % TC_DN == 0, "disparity tile is spread evenly over the dn threads");
      const int dt = tid - TC_W_DN * 32;
      int seq = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int tile0 = (it * cta_n + cta_i) * MT;
        for (int cj = 0; cj < n_src * MT; ++cj, ++seq) {
          const int ci = cj / MT, tile = tile0 + cj % MT;
          const int st = seq % C::NA;
          if ((ci + rot_chunk) % n_src != a.dn_chunk) {
            // Not mine, but observe the release of EVERY ring use in order: a parity wait issued two phases ahead of
            // the barrier would alias with the phase before and succeed at once.  Costs nothing: my next stage is
            // released after this one anyway.
            mbar_wait(bar_a_empty(st), ((seq / C::NA) & 1) ^ 1);
            continue;
          }
          const int x0 = (tile % tiles_x) * TC_TW, y0 = (tile / tiles_x) * TC_TH;
          if (dep_flags && tile < n_tiles) {      // disp was complete before the upstream conv published anything
            unsigned spins = 0;
            while (ld_acquire_gpu(a.flags_in + tile) == 0) {
              __nanosleep(200);
              if (++spins > (1u << 23)) __trap();
            }
          }
          float dv[DT_PER];                                        // in flight while we wait for the stage
#pragma unroll
          for (int q = 0; q < DT_PER; ++q) {
            const int i = q * TC_DN + dt;
            const int yy = y0 - 4 + i / TC_DT_W, xx = x0 - 4 + i % TC_DT_W;
            dv[q] = (yy >= 0 && yy < a.h && xx >= 0 && xx < a.w) ? __ldcg(a.disp + (long long)yy * a.w + xx) : 0.f;
          }
          pwait(bar_a_empty(st), ((seq / C::NA) & 1) ^ 1, 0);
          const long long t_fill0 = prof_on ? clock64() : 0;
          asm volatile("bar.sync 1, 96;" ::: "memory");            // previous tile's readers are done
#pragma unroll
          for (int q = 0; q < DT_PER; ++q) sD[q * TC_DN + dt] = dv[q];
          asm volatile("bar.sync 1, 96;" ::: "memory");
          unsigned char* dstp = smem + st * TC_A_BYTES;
          for (int hp = dt; hp < TC_HPX; hp += TC_DN) {
            const int hy = hp / TC_HW, hx = hp % TC_HW;
            const int yy = y0 - 1 + hy, xx = x0 - 1 + hx;
            const bool ok = yy >= 0 && yy < a.h && xx >= 0 && xx < a.w;
            const float* win = sD + hy * TC_DT_W + hx;             // window origin = (hy+3-3, hx+3-3)
            const float ctr = win[3 * TC_DT_W + 3];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              __align__(16) __half v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int k = g * 8 + e;                           // compile-time after unrolling
                float val = 0.f;
                if (k < kDispEnc) val = __fmul_rn(100.f, __fsub_rn(win[(k / 7) * TC_DT_W + (k % 7)], ctr));
                v[e] = __float2half_rn(ok ? val : 0.f);
              }
              *reinterpret_cast<uint4*>(dstp + g * TC_A_LBO + hp * 16) = *reinterpret_cast<const uint4*>(v);
            }
          }
          mbar_arrive(bar_a_full(st));       // plain stores, ordinary (release) arrival; the MMA warp fences the proxy
          if (prof_on) prof_acc[2] += clock64() - t_fill0;
        }
      }
    }
  } else if (warp == TC_W_BPROD) {
    // ================= tile publisher (lane 1): flags_out[tile] = 1 once all 8 epilogue warps have stored the tile ======
    // (st.release.gpu after an acquire of the warps' release-increments: cumulativity makes every epilogue thread's
    // stores visible before the flag; the epilogue itself never blocks on the publication)
    if (lane == 1 && a.flags_out != nullptr) {
      uint32_t done = 0;
      for (int t = 0; t < n_iter; ++t) {
        for (int j = 0; j < MT; ++j) {
          done += 8;
          unsigned spins = 0;
          while (ld_acquire_cta_shared(s0 + C::OFF_TMEM + 8) < done) {
            __nanosleep(64);
            if (++spins > (1u << 24)) __trap();
          }
          const int tile = (t * cta_n + cta_i) * MT + j;
          if (tile < n_tiles) st_release_gpu(a.flags_out + tile, 1);
        }
      }
    }
    // ================= B producer =================
    if (lane == 0) {
      const char* wsrc = reinterpret_cast<const char*>(CG2 ? a.wtc2 : a.wtc);
      if (C::RESIDENT) {
        for (int s = 0; s < C::NG; ++s) {      // loaded once, never released
          if (CG2) {
            if (rank == 0) mbar_expect_tx(bar_b_full(s), 2 * C::STAGE_BYTES);
#pragma unroll
            for (int kx = 0; kx < C::TPS; ++kx)
              tma2d_cg2(sB + s * C::STAGE_BYTES + kx * C::B_BYTES, &wmap, 0,
                        ((s * C::TPS + kx) * 2 + (int)rank) * (C::B_BYTES / 256), bar_b_full(s));
          } else if (HALF) {     // pair layout [tap][half][k-group][128][8]: this CTA's half of every tap is one 16 KB run
            mbar_expect_tx(bar_b_full(s), C::STAGE_BYTES);
            const char* w2src = reinterpret_cast<const char*>(a.wtc2);
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap)
              bulk_g2s(sB + tap * C::B_BYTES, w2src + (size_t)(tap * 2 + nhalf) * C::B_BYTES, C::B_BYTES, bar_b_full(s));
          } else {
            mbar_expect_tx(bar_b_full(s), C::STAGE_BYTES);
            bulk_g2s(sB + s * C::STAGE_BYTES, wsrc + (size_t)s * C::STAGE_BYTES, C::STAGE_BYTES, bar_b_full(s));
          }
        }
      } else {
        int seq = 0;
        const int n_steps = n_src * C::NG;
        for (int it = 0; it < n_iter; ++it) {
          for (int si = 0; si < n_steps; ++si, ++seq) {
            const int s = ((si / C::NG + rot_chunk) % n_src) * 9 + ((si % C::NG + rot_g) % C::NG) * C::TPS;   // first tap
            const int st = seq % C::NB;
            pwait(bar_b_empty(st), ((seq / C::NB) & 1) ^ 1, 0);
            if (CG2) {
              // each CTA fetches its half of the output channels with a TMA tensor load that credits the LEADER's
              // barrier (cta_group::2): the leader arms it for both halves, no software relay between the CTAs
              if (rank == 0) mbar_expect_tx(bar_b_full(st), 2 * C::STAGE_BYTES);
#pragma unroll
              for (int kx = 0; kx < C::TPS; ++kx)
                tma2d_cg2(sB + st * C::STAGE_BYTES + kx * C::B_BYTES, &wmap, 0,
                          ((s + kx) * 2 + (int)rank) * (C::B_BYTES / 256), bar_b_full(st));
              continue;
            }
            mbar_expect_tx(bar_b_full(st), MC2 ? C::B_BYTES : C::STAGE_BYTES);
            if (MC2) {
              // my half of the tile goes to both CTAs (same offset), the peer sends the other half
              constexpr uint32_t HALF = C::B_BYTES / 2;
              bulk_g2s_mc(sB + st * C::STAGE_BYTES + rank * HALF, wsrc + (size_t)s * C::B_BYTES + rank * HALF, HALF,
                          bar_b_full(st), (uint16_t)3);
            } else {
              // one copy per stage: the TPS taps of a kernel row are adjacent tiles of the packed weights
              bulk_g2s(sB + st * C::STAGE_BYTES, wsrc + (size_t)s * C::B_BYTES, C::STAGE_BYTES, bar_b_full(st));
            }
          }
        }
      }
      if (GRU_TMA) {
        // ---- q/GRU conv: this lane is idle once the resident weights are requested; it moves the element-wise operands
        // of the epilogue (z, net, q's x-part of one 16 x 8 tile = 64 KB) global -> shared memory one tile ahead, so the
        // epilogue never exposes their latency (round 1: eight epilogue warps in lock step each waited for eight 32-byte
        // loads per thread, once per tile; prefetching into registers spills, per-thread cp.async costs more than it hides)
        pdl_wait();                                   // z / qx / net come from the previous kernels of the chain
        const uint32_t sG = s0 + C::OFF_GRU;
        for (int t = 0; t < n_iter; ++t) {
          const int tile = t * cta_n + cta_i;
          const int x0 = (tile % tiles_x) * TC_TW, y0 = (tile / tiles_x) * TC_TH;
          pwait(bar_gru_empty, (t & 1) ^ 1, 0);       // every epilogue thread has copied tile t-1's operands to registers
          mbar_expect_tx(bar_gru_full, (uint32_t)C::GRU_BYTES);
          tma3d(sG, &amaps.m[1], 0, x0, y0, bar_gru_full);                         // z    [16][8][128 B]
          tma3d(sG + 128 * 128, &amaps.m[2], 0, x0, y0, bar_gru_full);             // net  [16][8][128 B]
          tma4d(sG + 2 * 128 * 128, &amaps.m[3], 0, 0, x0, y0, bar_gru_full);      // qx   [16][8][2][128 B]
        }
      }
    }
  } else if (warp == TC_W_MMA || (kTwoIssuers && warp == TC_W_MMA2)) {
    // ================= MMA issuers =================
    // Two warps take alternate (chunk, tap) steps: waiting on the stage barrier, building descriptors and issuing four
    // UTCHMMA + commit costs one thread ~790 cycles per step, the four MMAs only 384 -- a single issuer leaves the tensor
    // pipe half idle.  MMAs of both warps accumulate into the same TMEM tile (the sum is order-free); only the
    // step that starts a tile (accumulate = 0) must be issued first: its owner signals the other warp (named barrier 3).
    // CG2: leader CTA only.
    if (!CG2 || rank == 0) {     // each warp runs its loop converged (uniform control flow); one elected lane issues
      constexpr uint32_t idesc = umma_idesc(CG2 ? 256 : 128, N);
      const int mw = warp == TC_W_MMA ? 0 : 1;
      int aseq = 0, bseq = 0;
      for (int t = 0; t < n_iter; ++t) {
        const int as = t % C::NACC;
        pwait(bar_acc_empty(as), ((t / C::NACC) & 1) ^ 1, 0);     // epilogue(s) have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (MT * N);
        const bool own_first = !kTwoIssuers || (bseq & 1) == mw;   // do I issue the tile's first step?
        bool ordered = !kTwoIssuers;
        for (int c = 0; c < n_src; ++c, aseq += MT) {
          int ast[MT];
#pragma unroll
          for (int j = 0; j < MT; ++j) {
            ast[j] = (aseq + j) % C::NA;
            pwait(bar_a_full(ast[j]), ((aseq + j) / C::NA) & 1, 1);
            if (CG2) pwait(bar_b_peer(ast[j]), ((aseq + j) / C::NA) & 1, 1);   // the peer CTA's half of the pixel rows
          }
          // generic-proxy writes of the A stage (the dn chunk's st.shared) -> UMMA reads; stages filled by TMA alone need none
          if (EPI == EPI_GATES) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          for (int ti = 0; ti < C::NG; ++ti, ++bseq) {
            if (kTwoIssuers && (bseq & 1) != mw) continue;         // the other issuer's step
            const int g = C::RESIDENT ? ti : (ti + rot_g) % C::NG;
            int bst;
            if (C::RESIDENT) {
              bst = g;
              if (t == 0) mbar_wait(bar_b_full(bst), 0);
            } else {
              bst = bseq % C::NB;
              pwait(bar_b_full(bst), (bseq / C::NB) & 1, 2);         // CG2: armed for both CTAs' halves
            }
            if (!own_first && !ordered) {                          // the accumulate = 0 step has been issued
              asm volatile("bar.sync %0, 64;" ::"r"(3 + (t & 1)) : "memory");     // ids alternate: issuers are <= 1 tile apart
              ordered = true;
            }
            tc_fence_after();
            const long long t_issue0 = prof_on ? clock64() : 0;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < C::TPS; ++kk) {
                const int tap = g * C::TPS + kk;
                const int ky = tap / 3, kx = tap % 3;
                // descriptors of the 4 K=16 slices differ only in the start-address field: one 64-bit add each
                const uint64_t bd0 = umma_desc(sB + bst * C::STAGE_BYTES + kk * C::B_BYTES, C::NLOC * 16, 128);
#pragma unroll
                for (int j = 0; j < MT; ++j) {           // every M tile of the unit consumes this weight stage
                  const uint64_t ad0 = umma_desc(sA + ast[j] * TC_A_BYTES + (ky * TC_HW + kx) * 16, TC_A_LBO, TC_A_SBO);
#pragma unroll
                  for (int k16 = 0; k16 < 4; ++k16) {
                    const uint64_t ad = ad0 + (uint64_t)((2 * k16 * TC_A_LBO) >> 4);
                    const uint64_t bd = bd0 + (uint64_t)((2 * k16 * (C::NLOC * 16)) >> 4);
                    const uint32_t accum = (c > 0 || ti > 0 || kk > 0 || k16 > 0) ? 1u : 0u;
                    if (CG2) tc_mma2_f16(d_tmem + j * N, ad, bd, idesc, accum);
                    else tc_mma_f16(d_tmem + j * N, ad, bd, idesc, accum);
                  }
                }
              }
              if (!C::RESIDENT) {            // stage free (in both CTAs) once these MMAs have read it
                if (CG2) tc_commit2_mc(bar_b_empty(bst));
                else if (MC2) tc_commit1_mc(bar_b_empty(bst), (uint16_t)3);
                else tc_commit(bar_b_empty(bst));
              }
            }
            __syncwarp();
            if (own_first && !ordered) {                           // let the other issuer start on this tile
              asm volatile("bar.arrive %0, 64;" ::"r"(3 + (t & 1)) : "memory");
              ordered = true;
            }
            if (prof_on) prof_acc[3] += clock64() - t_issue0;
          }
          if (elect_one()) {                                       // my MMAs on this chunk's A stage(s)
#pragma unroll
            for (int j = 0; j < MT; ++j) {
              if (CG2) tc_commit2_mc(bar_a_empty(ast[j]));
              else tc_commit(bar_a_empty(ast[j]));
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (CG2) tc_commit2_mc(bar_acc_full(as));           // my share of the accumulation is complete
          else tc_commit(bar_acc_full(as));
        }
        __syncwarp();
      }
    } else if (warp == TC_W_MMA && lane == 0) {
      // CG2 peer: relay the completion of this CTA's A stages to the leader (its bar_b_peer slots double as "peer A full")
      int aseq = 0;
      for (int t = 0; t < n_iter; ++t) {
        for (int c = 0; c < n_src; ++c, ++aseq) {
          const int st = aseq % C::NA;
          mbar_wait(bar_a_full(st), (aseq / C::NA) & 1);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive_cluster(bar_b_peer(st), 0);
        }
      }
    }
  } else if (warp < 8) {
    // ================= epilogue: thread = one pixel, half of the N channels =================
    // warps w and w+4 share TMEM lane quadrant w (a warp may only touch lanes 32*(warp%4)..+31) and split the columns
    const int quad = warp & 3, chalf = warp >> 2;
    const int m = quad * 32 + lane;                 // row of the M tile = TMEM lane
    const int r = m >> 3, cc = m & 7;
    for (int t = 0; t < n_iter; ++t) {
      const int as = t % C::NACC;
#pragma unroll 1
      for (int j = 0; j < MT; ++j) {
      const int tile = (t * cta_n + cta_i) * MT + j;
      const int x0 = (tile % tiles_x) * TC_TW, y0 = (tile / tiles_x) * TC_TH;
      const int yy = y0 + r, xx = x0 + cc;
      const bool ok = yy < a.h && xx < a.w;
      const long long p = ok ? (long long)yy * a.w + xx : 0;
      // operands of the element-wise GRU algebra do not depend on the accumulator: fetch them before waiting for it
      // With tile flags this grid may have been launched long before its predecessors finished: nothing they (or the
      // grids before them) wrote may be read until a flag of the upstream conv has been seen -- a published tile
      // implies that conv passed its own grid-level wait, i.e. every earlier grid of the stream is complete.
      if (dep_flags && tile < n_tiles) wait_tile_flag(a.flags_in + tile, lane);
      uint4 pre_a[4];
      if (EPI == EPI_GATES && ok) ldg_half32_raw(a.net + p * 64 + chalf * 32, pre_a);   // net slice of this thread's r chunk
      // q/GRU conv: this thread's operands (pixel m, channels chalf*32..+31) out of the swizzled staging tile, then the
      // tile goes back to the producer -- before the accumulator wait, so the next tile's loads overlap this tile's math
      uint4 gz[4], gn[4];
      float4 gq[8];
      if (GRU_TMA) {
        mbar_wait(bar_gru_full, t & 1);
        const unsigned char* gt = smem + C::OFF_GRU;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int ch = ((chalf * 4 + k) ^ (m & 7)) * 16;                    // 128-byte swizzle: chunk ^ (row & 7)
          gz[k] = *reinterpret_cast<const uint4*>(gt + m * 128 + ch);
          gn[k] = *reinterpret_cast<const uint4*>(gt + 128 * 128 + m * 128 + ch);
        }
        const int qrow = 2 * m + chalf;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          gq[k] = *reinterpret_cast<const float4*>(gt + 2 * 128 * 128 + qrow * 128 + ((k ^ (qrow & 7)) * 16));
        // The loads must have DELIVERED before the tile is handed back: the arrive does not depend on their registers, the
        // loads queue behind the tensor core's operand reads (the MMAs of the next tiles saturate the shared-memory port),
        // and the copy engine overwrites the tile as soon as the last arrival is in.  Fold every loaded word into one value
        // the arrive is ordered behind.
        uint32_t sink = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) sink ^= gz[k].x ^ gz[k].y ^ gz[k].z ^ gz[k].w ^ gn[k].x ^ gn[k].y ^ gn[k].z ^ gn[k].w;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sink ^= __float_as_uint(gq[k].x) ^ __float_as_uint(gq[k].y) ^ __float_as_uint(gq[k].z) ^ __float_as_uint(gq[k].w);
        // (a conditional store that practically never fires: an empty asm leaves no trace in the PTX and ptxas drops the
        // whole chain -- measured: the arrive then overtakes loads that queue behind the MMAs' shared-memory reads and
        // the next tile's copy overwrites what they were about to read)
        if (sink == 0x5bd1e995u) *reinterpret_cast<volatile uint32_t*>(smem + C::OFF_TMEM + 12) = sink;
        mbar_arrive(bar_gru_empty);
      }
      if (j == 0) pwait(bar_acc_full(as), (t / C::NACC) & 1, 0);
      tc_fence_after();

      const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (MT * N) + j * N;
      float t9[9];
      if (EPI == EPI_DELTA) {
#pragma unroll
        for (int q = 0; q < 9; ++q) t9[q] = 0.f;
      }
      if (EPI == EPI_DELTA && kDeltaHeadMma) {
        // relu(fp16(acc + bias)) -> fp16 A fragments -> mma.sync against the delta.2 weights (taps 0..7 = n-tile 0, tap 8
        // = column 0 of n-tile 1): this warp's 32 pixels (two m16 tiles) x its 128 channels (8 k16 steps)
        const float* exb = reinterpret_cast<const float*>(smem + C::OFF_EXTRA);
        const __half* exw = reinterpret_cast<const __half*>(smem + C::OFF_EXTRA + 256 * 4);
        const int g = lane >> 2, q4 = lane & 3;
        const __half2 hzero = __floats2half2_rn(0.f, 0.f);
        float d0[2][4], d1[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int e = 0; e < 4; ++e) d0[mt][e] = d1[mt][e] = 0.f;
        // All of this warp's accumulator columns come out of TMEM first (N/32 x 2 loads in flight, one wait), and the
        // accumulator stage goes back to the MMA warp BEFORE the bias / ReLU / delta-head arithmetic: the conv is bound
        // by this epilogue (role profile: the MMA warp waits 17 % of its time for acc_empty), not by its MMAs.
        constexpr int NKS = N / 32;
        uint32_t r0[NKS][8], r1[NKS][8];
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
          const int ct = chalf * (N / 2) + ks * 16;        // TMEM column of this CTA's accumulator
          tc_ld_16x256b_x2(lane_addr + ct, r0[ks]);
          tc_ld_16x256b_x2(lane_addr + (16u << 16) + ct, r1[ks]);
        }
        tc_ld_wait();
        if (j == MT - 1) {
          tc_fence_before();
          arrive_leader(bar_acc_empty(as));           // accumulator stage may be overwritten
        }
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
          const int c0 = nhalf * 128 + chalf * (N / 2) + ks * 16;      // output channel of the delta.0 conv
          const uint32_t b00 = *reinterpret_cast<const uint32_t*>(exw + g * kW2Pitch + c0 + 2 * q4);
          const uint32_t b01 = *reinterpret_cast<const uint32_t*>(exw + g * kW2Pitch + c0 + 8 + 2 * q4);
          const uint32_t b10 = g == 0 ? *reinterpret_cast<const uint32_t*>(exw + 8 * kW2Pitch + c0 + 2 * q4) : 0u;
          const uint32_t b11 = g == 0 ? *reinterpret_cast<const uint32_t*>(exw + 8 * kW2Pitch + c0 + 8 + 2 * q4) : 0u;
          const float2 bb0 = *reinterpret_cast<const float2*>(exb + c0 + 2 * q4);
          const float2 bb1 = *reinterpret_cast<const float2*>(exb + c0 + 8 + 2 * q4);
          auto frag = [&](const uint32_t (&r)[8], uint32_t (&af)[4]) {
            const __half2 h0 = __hmax2(__floats2half2_rn(__uint_as_float(r[0]) + bb0.x, __uint_as_float(r[1]) + bb0.y), hzero);
            const __half2 h1 = __hmax2(__floats2half2_rn(__uint_as_float(r[2]) + bb0.x, __uint_as_float(r[3]) + bb0.y), hzero);
            const __half2 h2 = __hmax2(__floats2half2_rn(__uint_as_float(r[4]) + bb1.x, __uint_as_float(r[5]) + bb1.y), hzero);
            const __half2 h3 = __hmax2(__floats2half2_rn(__uint_as_float(r[6]) + bb1.x, __uint_as_float(r[7]) + bb1.y), hzero);
            af[0] = *reinterpret_cast<const uint32_t*>(&h0);
            af[1] = *reinterpret_cast<const uint32_t*>(&h1);
            af[2] = *reinterpret_cast<const uint32_t*>(&h2);
            af[3] = *reinterpret_cast<const uint32_t*>(&h3);
          };
          uint32_t a0[4], a1[4];
          frag(r0[ks], a0);
          frag(r1[ks], a1);
          mma16816_f32(d0[0], a0, b00, b01);
          mma16816_f32(d1[0], a0, b10, b11);
          mma16816_f32(d0[1], a1, b00, b01);
          mma16816_f32(d1[1], a1, b10, b11);
        }
        // accumulator fragment (row g / g+8 of m-tile mt, taps 2q, 2q+1; tap 8 in column 0 of the second n-tile)
        // HALF: this CTA owns part `nhalf` of the delta partials; its two column halves are added here in a fixed order
        // (warps 4..7 park theirs in shared memory, warps 0..3 add and store), so the consumers still sum two parts.
        float* sX = reinterpret_cast<float*>(smem + C::OFF_EXTRA + 256 * 4 + 9 * kW2Pitch * 2);      // [128 px][9]
        if (HALF && chalf == 1) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              float* d = sX + (quad * 32 + mt * 16 + g + 8 * hf) * 9;
              d[2 * q4] = d0[mt][2 * hf];
              d[2 * q4 + 1] = d0[mt][2 * hf + 1];
              if (q4 == 0) d[8] = d1[mt][2 * hf];
            }
        }
        if (HALF) asm volatile("bar.sync %0, 64;" ::"r"(8 + quad) : "memory");       // the two warps of this lane quadrant
        if (!HALF || chalf == 0) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int mm = quad * 32 + mt * 16 + g + 8 * hf;
              const int y2 = y0 + (mm >> 3), x2 = x0 + (mm & 7);
              float v0 = d0[mt][2 * hf], v1 = d0[mt][2 * hf + 1], v8 = d1[mt][2 * hf];
              if (HALF) {
                const float* d = sX + mm * 9;
                v0 += d[2 * q4];
                v1 += d[2 * q4 + 1];
                if (q4 == 0) v8 += d[8];
              }
              if (y2 < a.h && x2 < a.w) {
                const long long npx = (long long)a.h * a.w, pp = (long long)y2 * a.w + x2;   // 8 lanes (g) = one 32-byte sector per tap plane
                const int part = HALF ? nhalf : chalf;
                a.s9[s9_index(npx, part, 2 * q4, pp)] = v0;
                a.s9[s9_index(npx, part, 2 * q4 + 1, pp)] = v1;
                if (q4 == 0) a.s9[s9_index(npx, part, 8, pp)] = v8;
              }
            }
        }
        if (HALF) asm volatile("bar.sync %0, 64;" ::"r"(8 + quad) : "memory");       // the parked values have been read
      } else {
#pragma unroll 1
      for (int cb = chalf * (N / 64); cb < (chalf + 1) * (N / 64); ++cb) {
        uint32_t raw[32];
        tc_ld32(lane_addr + cb * 32, raw);          // warp-collective: executed by every lane
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(raw[e]);
        const int n0 = cb * 32;
        if (EPI == EPI_RELU || EPI == EPI_GATES) {      // bias: 8 uniform 16-byte loads
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n0) + q4);
            v[4 * q4] += bb.x; v[4 * q4 + 1] += bb.y; v[4 * q4 + 2] += bb.z; v[4 * q4 + 3] += bb.w;
          }
        }
        if (EPI == EPI_RELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);     // relu commutes with the (monotone) fp16 rounding of st_half32
          if (ok) st_half32(a.out_h + p * 64 + n0, v);
        } else if (EPI == EPI_GATES) {
          if (n0 < 64) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = fast_sigmoid(h_round(v[e]));
            if (ok) st_half32(a.z + p * 64 + n0, v);
          } else if (n0 < 128) {
            if (ok) {      // fp16(fp16(sigmoid(fp16(convr))) * net): the product of two fp16 values rounded once = HMUL2
              const __half2* n2 = reinterpret_cast<const __half2*>(pre_a);          // n0 - 64 == chalf * 32
              uint32_t outp[16];
#pragma unroll
              for (int e2 = 0; e2 < 16; ++e2) {
                const float2 af = __half22float2(__floats2half2_rn(v[2 * e2], v[2 * e2 + 1]));
                const __half2 o = __hmul2(__floats2half2_rn(fast_sigmoid(af.x), fast_sigmoid(af.y)), n2[e2]);
                outp[e2] = *reinterpret_cast<const uint32_t*>(&o);
              }
              const uint32_t(&lo)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&outp[0]);
              const uint32_t(&hi)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&outp[8]);
              st256(a.rnet + p * 64 + (n0 - 64), lo);
              st256(a.rnet + p * 64 + (n0 - 64) + 16, hi);
            }
          } else if (ok) {
            st_float32(a.qx + p * 64 + (n0 - 128), v);
          }
        } else if (EPI == EPI_GRUOUT) {
          if (ok) {
            // (prefetching z / net / qx before the accumulator wait was measured slower: 39 vs 36 us -- ptxas holds this
            // kernel at 128 registers, the extra 64 live registers spill, and the loads are L2 hits that the eight
            // epilogue warps already overlap.  Round 2: a cp.async prefetch one tile ahead into a per-thread shared-memory
            // slot, no extra registers: 33.9 vs 28.8 us -- sixteen 16-byte copies per thread cost more than they hide.)
            // The element-wise GRU algebra on packed fp16 pairs: autocast rounds every op to fp16, and for fp16 operands
            // HSUB2(1, z) and HMUL2 are exactly fp16(float op) (1 - z and the products are exact in fp32); the final sum
            // is done in fp32 and rounded once, like torch's opmath path.  ~35 % fewer epilogue instructions than the
            // unpack-to-float version (the q conv is bound by its epilogue's issue slots, not by its MMAs).
            uint4 zk[4], nk[4];
            if (GRU_TMA) {
#pragma unroll
              for (int k = 0; k < 4; ++k) { zk[k] = gz[k]; nk[k] = gn[k]; }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                v[4 * k] += gq[k].x; v[4 * k + 1] += gq[k].y; v[4 * k + 2] += gq[k].z; v[4 * k + 3] += gq[k].w;
              }
            } else {
              ldg_half32_raw(a.z + p * 64 + n0, zk);
              ldg_half32_raw(a.net + p * 64 + n0, nk);
              ld_float32_add(a.qx + p * 64 + n0, v);
            }
            const __half2* z2 = reinterpret_cast<const __half2*>(zk);
            const __half2* n2 = reinterpret_cast<const __half2*>(nk);
            const __half2 one2 = __floats2half2_rn(1.f, 1.f);
            uint32_t outp[16];
#pragma unroll
            for (int e2 = 0; e2 < 16; ++e2) {
              const float2 af = __half22float2(__floats2half2_rn(v[2 * e2], v[2 * e2 + 1]));        // fp16(convq)
              const __half2 q = __floats2half2_rn(fast_tanh(af.x), fast_tanh(af.y));                  // fp16(tanh)
              const float2 f1 = __half22float2(__hmul2(__hsub2(one2, z2[e2]), n2[e2]));              // fp16(fp16(1-z)*net)
              const float2 f2 = __half22float2(__hmul2(z2[e2], q));                                   // fp16(z*q)
              const __half2 o = __floats2half2_rn(f1.x + f2.x, f1.y + f2.y);
              outp[e2] = *reinterpret_cast<const uint32_t*>(&o);
            }
            {
              const uint32_t(&lo)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&outp[0]);
              const uint32_t(&hi)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&outp[8]);
              st256(a.net + p * 64 + n0, lo);
              st256(a.net + p * 64 + n0 + 16, hi);
            }
          }
        } else {  // EPI_DELTA
          // relu(fp16(acc + bias)) stays packed in fp16; the 9-tap dot with the (fp16) delta.2 weights runs on the
          // mixed-precision FMA (FHFMA: exact fp16 x fp16 product, fp32 accumulate, same channel order as before)
          const float* exb = reinterpret_cast<const float*>(smem + C::OFF_EXTRA);
          const __half* exw = reinterpret_cast<const __half*>(smem + C::OFF_EXTRA + 256 * 4);
          uint32_t hv[16];
          const __half2 hzero = __floats2half2_rn(0.f, 0.f);
#pragma unroll
          for (int e2 = 0; e2 < 16; ++e2) {
            const float2 bb = *reinterpret_cast<const float2*>(exb + n0 + 2 * e2);
            const __half2 h = __hmax2(__floats2half2_rn(v[2 * e2] + bb.x, v[2 * e2 + 1] + bb.y), hzero);
            hv[e2] = *reinterpret_cast<const uint32_t*>(&h);
          }
#pragma unroll
          for (int q = 0; q < 9; ++q) {
            float acc = t9[q];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 w = *reinterpret_cast<const uint4*>(exw + q * kW2Pitch + n0 + 8 * j4);
              acc = fhfma_lo(hv[4 * j4 + 0], w.x, acc); acc = fhfma_hi(hv[4 * j4 + 0], w.x, acc);
              acc = fhfma_lo(hv[4 * j4 + 1], w.y, acc); acc = fhfma_hi(hv[4 * j4 + 1], w.y, acc);
              acc = fhfma_lo(hv[4 * j4 + 2], w.z, acc); acc = fhfma_hi(hv[4 * j4 + 2], w.z, acc);
              acc = fhfma_lo(hv[4 * j4 + 3], w.w, acc); acc = fhfma_hi(hv[4 * j4 + 3], w.w, acc);
            }
            t9[q] = acc;
          }
        }
      }
      if (j == MT - 1) {
        tc_fence_before();
        arrive_leader(bar_acc_empty(as));           // accumulator stage may be overwritten
      }
      if (EPI == EPI_DELTA && ok) {   // two partial sums per pixel (one per column half), added by the consumer
#pragma unroll
        for (int q = 0; q < 9; ++q) a.s9[s9_index((long long)a.h * a.w, chalf, q, p)] = t9[q];
      }
      }   // generic (per-pixel) epilogue
      if (a.flags_out != nullptr) {   // this warp's stores of the tile are issued: count it (the publisher thread releases the flag)
        __syncwarp();
        if (lane == 0) red_release_cta_shared_inc(s0 + C::OFF_TMEM + 8);
      }
      }   // j
    }
  }

  if (prof_on && (warp == 0 || warp == TC_W_MMA || warp == TC_W_BPROD || warp == TC_W_APROD)) {   // (first MMA issuer only)
    const int role = warp == 0 ? 0 : warp == TC_W_MMA ? 1 : warp == TC_W_BPROD ? 2 : 3;   // epilogue, mma, b, a
    for (int i = 0; i < 4; ++i) a.prof[role * 8 + i] = (unsigned long long)prof_acc[i];
    a.prof[role * 8 + 7] = (unsigned long long)(clock64() - prof_t0);
  }
  if (prof_on && warp == TC_W_DN) {      // dn warp: fill cycles / wait for a free stage, in the A row's spare slots
    a.prof[3 * 8 + 4] = (unsigned long long)prof_acc[2];
    a.prof[3 * 8 + 5] = (unsigned long long)prof_acc[0];
  }
  tc_fence_before();
  __syncthreads();
  if (CG2 || MC2) cluster_sync_all();     // no CTA leaves while its partner can still arrive on / write to its shared memory
  if (warp == TC_W_MMA) {
    if (CG2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"((uint32_t)C::TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// ---- host ----------------------------------------------------------------------------------------
// Pair modes apply to the N = 192 / 256 convs (weights streamed per tile); the N = 64 convs keep their weights resident.
// Mode of the streamed-weight convs (A/B switch), per kernel: [0] gate conv (N = 192), [1] delta conv (N = 256).
// Measured default: CTA pairs for the gate conv (79 vs 84 us); the delta conv as single CTAs (43 us; as a pair with
// resident half weight sets its M256 x N256 MMAs run at ~200 instead of 128 cycles: 48 us).
template <int N, int EPI>
static int tc_configure_one() {
  if constexpr (N == 256) {        // the delta conv only exists as channel halves (TC_HALF)
    CER_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<128, EPI_DELTA, TC_HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TcCfg<128, TC_HALF>::TOTAL));
  } else {
    CER_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<N, EPI, TC_SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TcCfg<N, TC_SINGLE>::TOTAL));
    if constexpr (N == 192)
      CER_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<N, EPI, TC_CG2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    TcCfg<N, TC_CG2>::TOTAL));
  }
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

int tc_num_tiles(int h, int w) { return ((w + TC_TW - 1) / TC_TW) * ((h + TC_TH - 1) / TC_TH); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D byte view of the CTA-pair weight layout: rows of 256 B, one box = one CTA's half of a (chunk, tap) tile.
static int make_weight_map(const void* base, size_t total_bytes, int half_bytes, CUtensorMap* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CER_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return CER_ERR_INVALID;
    }
    fn = (EncodeTiledFn)p;
  }
  const cuuint64_t gdim[2] = {256, (cuuint64_t)(total_bytes / 256)};
  const cuuint64_t gstride[1] = {256};
  const cuuint32_t box[2] = {256, (cuuint32_t)(half_bytes / 256)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CER_ERR_INVALID;
  }
  return CER_OK;
}

// A operand by TMA tensor loads (the cp.async path of round 1 remains in the kernel as dead `use_tma == 0` code)
static int a_tma() { return 1; }

static int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CER_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return CER_ERR_INVALID;
    }
    fn = (EncodeTiledFn)p;
  }
  *out = fn;
  return CER_OK;
}

static int make_act_maps(const ConvArgs& a, TcMaps* out) {
  memset(out, 0, sizeof(*out));
  out->use_tma = a_tma();
  if (!out->use_tma) return CER_OK;
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc) return rc;
  for (int c = 0; c < a.n_src; ++c) {
    if (c == a.dn_chunk) continue;
    if (!aligned16(a.src[c])) {
      set_error("conv3x3_tc: source tensor %d is not 16-byte aligned", c);
      return CER_ERR_INVALID;
    }
    // dims, innermost first: 8 channels of a group (16 B) | x | y | the 8 channel groups of a pixel
    const cuuint64_t gdim[4] = {8, (cuuint64_t)a.w, (cuuint64_t)a.h, 8};
    const cuuint64_t gstride[3] = {128, (cuuint64_t)a.w * 128, 16};
    const cuuint32_t box[4] = {8, (cuuint32_t)TC_HW, (cuuint32_t)TC_HH, 8};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&out->m[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(a.src[c]), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (activation halo map) failed (%d)", (int)r);
      return CER_ERR_INVALID;
    }
  }
  return CER_OK;
}

// q/GRU conv: maps of the epilogue's element-wise operands, one 16 x 8-pixel tile per load, 128-byte swizzle
// (m[1] = z, m[2] = net: fp16 [h][w][64] as (64, w, h); m[3] = q's x-part: fp32 [h][w][64] as (32, 2, w, h))
static int make_gru_maps(const ConvArgs& a, TcMaps* out) {
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc) return rc;
  const void* src16[2] = {a.z, a.net};
  for (int i = 0; i < 2; ++i) {
    const cuuint64_t gdim[3] = {64, (cuuint64_t)a.w, (cuuint64_t)a.h};
    const cuuint64_t gstride[2] = {128, (cuuint64_t)a.w * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)TC_TW, (cuuint32_t)TC_TH};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&out->m[1 + i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(src16[i]), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (GRU operand map %d) failed (%d)", i, (int)r);
      return CER_ERR_INVALID;
    }
  }
  {
    const cuuint64_t gdim[4] = {32, 2, (cuuint64_t)a.w, (cuuint64_t)a.h};
    const cuuint64_t gstride[3] = {128, 256, (cuuint64_t)a.w * 256};
    const cuuint32_t box[4] = {32, 2, (cuuint32_t)TC_TW, (cuuint32_t)TC_TH};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&out->m[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.qx), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (GRU q map) failed (%d)", (int)r);
      return CER_ERR_INVALID;
    }
  }
  return CER_OK;
}

template <int N, int EPI, int MODE>
static int launch_pair(const ConvArgs& a, int tiles, int kind, cudaStream_t stream) {
  CUtensorMap wmap;
  memset(&wmap, 0, sizeof(wmap));
  if (MODE == TC_CG2) {
    int rc = make_weight_map(a.wtc2, (size_t)a.n_src * 9 * 64 * N * 2, TcCfg<N, MODE>::B_BYTES, &wmap);
    if (rc) return rc;
  }
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  grid &= ~1;                                             // whole CTA pairs
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc_threads(EPI));
  cfg.dynamicSmemBytes = TcCfg<N, MODE>::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cer::g_pdl ? 2 : 1;
  if (cer::g_timer) cer::timer_begin(kind, stream);
  TcMaps amaps;
  {
    int rc = make_act_maps(a, &amaps);
    if (rc) return rc;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<N, EPI, MODE>, a, wmap, amaps);
  if (cer::g_timer) cer::timer_end(stream);
  ++cer::g_launches;
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("conv3x3_tc (cta pair) launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return check_launch("conv3x3_tc");
}

template <int N, int EPI>
int launch_conv_tc(const ConvArgs& a, cudaStream_t stream) {
  const int tiles = ((a.w + TC_TW - 1) / TC_TW) * ((a.h + TC_TH - 1) / TC_TH);
  constexpr int kind = EPI == EPI_RELU ? KK_CONV_E : EPI == EPI_GATES ? KK_CONV_GATES : EPI == EPI_GRUOUT ? KK_CONV_Q : KK_CONV_DELTA;
  TcMaps amaps;
  {
    int rc = make_act_maps(a, &amaps);
    if (rc) return rc;
  }
  if constexpr (EPI == EPI_GRUOUT) {
    int rc = make_gru_maps(a, &amaps);
    if (rc) return rc;
  }
  CUtensorMap nomap;
  memset(&nomap, 0, sizeof(nomap));
  if constexpr (N == 256) {
    // the delta conv as channel halves: CTA 2i / 2i+1 walk the same tiles with one half of the weights resident each
    // (43.5 -> 40.8 us event-timed, 9.10 -> 8.86 ms per step against one N = 256 tile per CTA with streamed weights)
    int grid2 = 2 * tiles < kNumSMs ? 2 * tiles : kNumSMs;
    grid2 &= ~1;
    CER_LAUNCH_PDL(kind, (conv3x3_tc_kernel<128, EPI_DELTA, TC_HALF>), grid2, tc_threads(EPI), (TcCfg<128, TC_HALF>::TOTAL),
                   stream, a, nomap, amaps);
    return check_launch("conv3x3_tc");
  } else {
    // the gate conv (N = 192) runs as cta_group::2 CTA pairs (79 vs 84 us); the N = 64 convs one 128-pixel tile per CTA
    if constexpr (N == 192) {
      if (tiles >= 2) return launch_pair<N, EPI, TC_CG2>(a, tiles, kind, stream);
    }
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;     // persistent: one CTA per SM
    CER_LAUNCH_PDL(kind, (conv3x3_tc_kernel<N, EPI, TC_SINGLE>), grid, tc_threads(EPI), (TcCfg<N, TC_SINGLE>::TOTAL), stream, a,
                   nomap, amaps);
    return check_launch("conv3x3_tc");
  }
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
