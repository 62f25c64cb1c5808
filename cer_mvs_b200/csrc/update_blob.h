// Packed UpdateBlock weights (core/update.py:58-78) as the kernels read them.
// One blob, every section 256-byte aligned.  fp16 weights are [chunk][tap][k=64][n] with
// chunk = 64-channel slice of the conv input, tap = ky*3+kx, n = output channel.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace cer {

constexpr int kCorrPlanes = 33;   // num_levels * (2*radius+1)
constexpr int kCorrK = 48;        // 33 padded to 3 x k16
constexpr int kHid = 64;          // dim_net = dim_inp = dim0_corr = dim1_corr
constexpr int kDispEnc = 49;      // size_disp_enc^2
constexpr int kGruIn = 241;       // 64 + 64 + 49 + 64
constexpr int kDelta0 = 256;      // dim0_delta
constexpr int kGateN = 192;       // z | r | q(x-part)

struct BlobLayout {
  size_t w1;      // fp16 [48][64]            corr_encoder.0
  size_t b1;      // f32  [64]
  size_t w2;      // fp16 [1][9][64][64]      corr_encoder.2
  size_t b2;      // f32  [64]
  size_t wg;      // fp16 [4][9][64][192]     gates: chunks net | inp | dn | e ; n = z | r | q
  size_t bg;      // f32  [192]               bz | br | bq
  size_t wq;      // fp16 [1][9][64][64]      convq, input channels 0..63 (r*net)
  size_t wd0[2];  // fp16 [1][9][64][256]     delta{s}.0
  size_t bd0[2];  // f32  [256]
  size_t wd1[2];  // f32  [9][256]            delta{s}.2 (fp16-rounded values)
  size_t bd1[2];  // f32  [1]
  // tcgen05 copies of the 3x3 weights: per (chunk, tap) one UMMA K-major no-swizzle B tile,
  // [kgroup = 8][n = N][8 halfs]  (element (k, n) at ((k/8 * N + n) * 8 + k%8) * 2 bytes)
  size_t t_w2;     // [1*9] tiles, N = 64
  size_t t_wg;     // [4*9] tiles, N = 192
  size_t t_wq;     // [1*9] tiles, N = 64
  size_t t_wd0[2]; // [1*9] tiles, N = 256
  // CTA-pair (cta_group::2) copies: per (chunk, tap) [half 2][kgroup 8][n N/2][8 halfs] so that each CTA's half of
  // the output channels is one contiguous bulk copy
  size_t p_wg;     // [4*9] tiles, N = 192
  size_t p_wd0[2]; // [1*9] tiles, N = 256
  size_t w1f;      // u32 [3][32][8][2]  corr_encoder.0 as mma.sync m16n8k16 B fragments (lookup_enc1_v4_kernel)
  size_t total;
};

inline constexpr size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

inline BlobLayout blob_layout() {
  BlobLayout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return r; };
  L.w1 = take(kCorrK * kHid * 2);
  L.b1 = take(kHid * 4);
  L.w2 = take(9 * 64 * kHid * 2);
  L.b2 = take(kHid * 4);
  L.wg = take(4 * 9 * 64 * kGateN * 2);
  L.bg = take(kGateN * 4);
  L.wq = take(9 * 64 * kHid * 2);
  for (int s = 0; s < 2; ++s) {
    L.wd0[s] = take(9 * 64 * kDelta0 * 2);
    L.bd0[s] = take(kDelta0 * 4);
    L.wd1[s] = take(9 * kDelta0 * 4);
    L.bd1[s] = take(4);
  }
  L.t_w2 = take(9 * 64 * kHid * 2);
  L.t_wg = take(4 * 9 * 64 * kGateN * 2);
  L.t_wq = take(9 * 64 * kHid * 2);
  for (int s = 0; s < 2; ++s) L.t_wd0[s] = take(9 * 64 * kDelta0 * 2);
  L.p_wg = take(4 * 9 * 64 * kGateN * 2);
  for (int s = 0; s < 2; ++s) L.p_wd0[s] = take(9 * 64 * kDelta0 * 2);
  L.w1f = take((kCorrK / 16) * 32 * 8 * 2 * 4);
  L.total = o;
  return L;
}

}  // namespace cer
