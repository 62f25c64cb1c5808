/* cer_mvs_b200 -- C ABI of the B200-native CER-MVS inference hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch types) and a CUDA
 * stream handle, never allocates device memory behind the caller's back (except the opaque
 * `cer_plan`, which owns its workspace), never synchronises unless it says so, and returns 0 on
 * success or a non-zero cudaError_t-style code (`cer_last_error()` gives the text).
 * All device pointers must be 16-byte aligned.  Launches go on the given stream and are CUDA-graph
 * capturable.
 *
 * Each function cites the interface of the reference (princeton-vl/CER-MVS @ 8062ddf) it replaces.
 */
#ifndef CER_MVS_B200_H
#define CER_MVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cer_stream_t; /* cudaStream_t */

#define CER_OK 0
#define CER_ERR_INVALID 1001   /* bad argument (shape / alignment / unsupported size) */
#define CER_ERR_NO_DEVICE 1002 /* no sm_100 device: the library has no CPU fallback */

int cer_abi_version(void);
const char* cer_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x). */
int cer_device_check(void);

/* ---- alt_cuda_corr.forward ---------------------------------------------------------------
 * Replaces corr_forward / corr_cuda_forward / corr_forward_kernel
 * (alt_cuda_corr/correlation.cpp:23-33,52; correlation_kernel.cu:18-119,260-286).
 * fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2] (x,y in fmap2 pixels), all fp32
 * contiguous; corr [B,N,(2r+1)^2,H1,W1] fp32 is fully overwritten (the reference zero-fills and
 * accumulates).  Any radius >= 0, any C >= 1. */
int cer_corr_forward_f32(const float* fmap1, const float* fmap2, const float* coords, float* corr,
                         int B, int H1, int W1, int H2, int W2, int C, int N, int radius,
                         cer_stream_t stream);

/* ---- feature / context layout ------------------------------------------------------------
 * Replaces the permute / "/8.0" / contiguous / float chain of direct_corr (core/corr.py:29-35).
 * src [n,C,h,w] (fp16 or fp32) -> dst [n,h,w,C] (fp16 or fp32), multiplied by `scale`. */
int cer_nchw_to_nhwc(const void* src, int src_f16, void* dst, int dst_f16, int n, int C, int h, int w,
                     float scale, cer_stream_t stream);
/* Same with a destination channel pitch dstC >= C (channels C..dstC-1 are written as zero). */
int cer_nchw_to_nhwc_pad(const void* src, int src_f16, void* dst, int dst_f16, int n, int C, int dstC, int h,
                         int w, float scale, cer_stream_t stream);
/* The inverse layout change (used to hand `net` back as [1,1,64,h,w], core/update.py:116). */
int cer_nhwc_to_nchw(const void* src, int src_f16, void* dst, int dst_f16, int n, int C, int h, int w,
                     cer_stream_t stream);

/* ---- projective_transform, matrix part (utils/projective_ops.py:16-23) ---------------------
 * Pij[k] = K4(jj[k]) . P(jj[k]) . P(ii[k])^-1 . K4(ii[k])^-1, computed in fp64, stored fp32 row-major.
 * poses [n,4,4] world->camera (already scaled, core/raft.py:35), intrinsics [n,3,3] already divided
 * by the encoder stride (core/raft.py:39); ii, jj device int32 [n_pairs]. */
int cer_projection_matrices(const float* poses, const float* intrinsics, const int* ii, const int* jj,
                            int n_pairs, float* Pij, cer_stream_t stream);

/* ---- CorrBlock.__init__ (core/corr.py:46-97), fused ----------------------------------------
 * hypotheses (core/corr.py:56-66) + projection of (x,y,1,d) (utils/projective_ops.py:5-13,25-27) +
 * clamp (corr.py:88) + alt_cuda_corr.forward with radius 0 (correlation_kernel.cu) + D-minor layout
 * (corr.py:41-43,89-91), for all (ii[k], jj[k]) pairs in one launch.
 *   feats   [n_img,h,w,64] NHWC, fp16 or fp32, ALREADY multiplied by 1/8 (cer_nchw_to_nhwc)
 *   Pij     [n_pairs,16] (cer_projection_matrices), ii/jj device int32 [n_pairs]
 *   disp_in [h,w] fp32; shift != 0 for stage 0 (corr.py:59-62)
 *   origin  [h,w] fp32 out (= CorrBlock.disps_origin)
 *   volume  out, fp32, D minor.  per_view == 0: [h*w, D] = out_scale * sum over pairs (use 1/V for the
 *           view mean of core/update.py:103; a rank that owns a subset of views passes 1/V_total and the
 *           partial volumes are summed across ranks).  per_view != 0: [n_pairs, h*w, D], the reference's
 *           corr_pyramid[0] row order ((v*h+y)*w+x), each multiplied by out_scale. */
int cer_build_volume(const void* feats, int feats_f16, const float* Pij, const int* ii, const int* jj,
                     int n_pairs, const float* disp_in, int shift, int D, float incre, float lo_origin,
                     float* origin, float* volume, float out_scale, int per_view, int h, int w,
                     cer_stream_t stream);

/* Hypotheses [d_begin, d_end) of the given views only (fp16 features); accumulate != 0 ADDS the result to `volume`
 * instead of overwriting it.  This is the unit of work of the sharded build: a rank owns a contiguous run of
 * (view, hypothesis) units, builds it into a zeroed partial volume, and the partial volumes are summed by one
 * all-reduce (core/corr.py:84-91 is a sum over views of independent per-hypothesis terms). */
int cer_build_volume_part(const void* feats, int feats_f16, const float* Pij, const int* ii, const int* jj,
                          int n_pairs, const float* disp_in, int shift, int D, float incre, float lo_origin,
                          float* origin, float* volume, float out_scale, int per_view, int h, int w, int d_begin,
                          int d_end, int accumulate, cer_stream_t stream);

/* Kernel used for fp16 features (cer_build_volume / cer_build_volume_part and the plan):
 *   0 = shared-memory-staged source boxes: TMA tensor loads of the epipolar bounding box of a 16 x 8 pixel tile and a
 *       chunk of hypotheses, tile-vs-box dots on tcgen05.mma, bilinear blend of correlation scalars (default;
 *       csrc/build_volume_tc.cu);
 *   1 = L1 row gather, 4 lanes per pixel, 256-bit loads + FHFMA (CER_BUILD=gather; csrc/build_volume.cu).
 * Both restate core/corr.py:46-97 + alt_cuda_corr.forward (radius 0); they differ in the summation order of the
 * 64-channel dot only.  fp32 features always use the generic gather kernel with plain FFMA. */
int cer_set_build_variant(int variant);
/* Debug / profiling: 16 x uint64 device counters filled by the staged build kernel (chunks, plan attempts, MMA / ZERO /
 * DIRECT chunks, sum of chunk lengths, sum of MMA N, cycles per role; layout in csrc/build_volume_tc.cu); NULL = off. */
int cer_debug_set_build_profile(unsigned long long* dev_counters);

/* avg_pool2d([1,2]) pyramid level (core/corr.py:95-97): src [rows, W] -> dst [rows, W/2] (floor). */
int cer_pool_pairs(const float* src, float* dst, long long rows, int W, cer_stream_t stream);

/* ---- BasicEncoder, type "HR" (core/extractor.py:62-155; fnet / cnet of core/raft.py:28-29,57,66-69) --------------
 * the producer of the hot path's inputs.  mma.sync implicit-GEMM convolutions on NHWC fp16
 * activations, instance-norm statistics in fp32, autocast rounding points (csrc/encoder.cu).
 * Weights: 22 arrays in state-dict order (conv1, layer1.{0,1}.conv{1,2}, layer2.0.{conv1,conv2,downsample.0},
 * layer2.1.{conv1,conv2}, conv2; weight then bias each), OIHW fp32. */
size_t cer_encoder_blob_bytes(int out_dim);
size_t cer_encoder_workspace_bytes(int H, int W);
int cer_pack_encoder_weights(const float* const* w, int out_dim, void* blob_host);
/* One image [3][H][W] fp32 (normalize != 0: 0..255 input, x*2/255-1 applied first, core/raft.py:40-41).
 *   fnet (out_dim 64, instance_norm 1): out_nhwc [H/4*W/4][64] fp16 scaled by nhwc_scale (the build's layout; nullable),
 *        out_nchw [64][H/4*W/4] fp16 (what the reference module returns; nullable)
 *   cnet (out_dim 128, instance_norm 0), context_split 1: net = tanh(ch 0..63), inp = relu(ch 64..127)
 *        (core/raft.py:58-60): out_nhwc = net, out_nhwc2 = inp ([H/4*W/4][64] fp16 each; nullable), out_nchw
 *        [2][64][H/4*W/4] (nullable); context_split 0: the raw 128-channel map (out_nchw [128][..] / out_nhwc [..][128]) */
int cer_encoder_forward(const void* blob, void* workspace, const float* image, int H, int W, int normalize, int out_dim,
                        int instance_norm, int context_split, void* out_nhwc, void* out_nhwc2, void* out_nchw,
                        float nhwc_scale, cer_stream_t stream);

/* ---- CorrBlock.__call__ (core/corr.py:102-143 + utils/bilinear_sampler.py:6-25) -------------
 * volume [slots, h*w, D] level 0 (levels 1..L-1 are rebuilt on the fly), origin [h*w], zinv [h*w]
 * -> out [slots, L*(2r+1), h, w] fp32 (channel = level*(2r+1) + tap).  L <= 3 needs D >= 4. */
int cer_lookup(const float* volume, int slots, const float* origin, const float* zinv, int D,
               float incre, int radius, int num_levels, float* out, int h, int w, cer_stream_t stream);

/* Same with one zinv map per slot: zinv + slot*zinv_stride (0 = shared, what core/raft.py:99 passes). */
int cer_lookup_strided(const float* volume, int slots, const float* origin, const float* zinv,
                       long long zinv_stride, int D, float incre, int radius, int num_levels, float* out,
                       int h, int w, cer_stream_t stream);

/* ---- UpdateBlock (core/update.py:29-120) ---------------------------------------------------
 * Weights are handed over as one packed blob.  cer_update_blob_bytes() gives its size;
 * cer_pack_update_weights() fills a HOST blob from the reference's state-dict tensors (fp32, OIHW,
 * host pointers, order below); copy it to the device once.
 *   w[0..1]  corr_encoder.0.weight [64,33,1,1], .bias [64]
 *   w[2..3]  corr_encoder.2.weight [64,64,3,3], .bias
 *   w[4..5]  gru.convz.weight [64,241,3,3], .bias      w[6..7] gru.convr     w[8..9] gru.convq
 *   w[10..11] delta0.0.weight [256,64,3,3], .bias      w[12..13] delta0.2.weight [1,256,3,3], .bias
 *   w[14..17] delta1.* likewise
 * Only the reference's default architecture is supported (3 levels, radius 5, 7x7 disparity
 * encoder, aggregation ["mean"], share_corr, share_gru, per-stage delta). */
size_t cer_update_blob_bytes(void);
int cer_pack_update_weights(const float* const* w, void* blob_host);

/* Workspace (device) for one UpdateBlock.forward at h x w. */
size_t cer_update_workspace_bytes(int h, int w);

/* One UpdateBlock.forward (core/update.py:87-120), autocast semantics (fp16 operands, fp32 accumulate,
 * fp16 rounding where torch.cuda.amp.autocast rounds).
 *   net   [h*w,64] fp16 NHWC, updated IN PLACE (the GRU state)      inp [h*w,64] fp16 NHWC
 *   disp  [h*w] fp32, updated IN PLACE when `apply_delta` (core/raft.py:101)
 *   corr  [slots,33,h,w] fp32 (CorrBlock.__call__ output); the mean over slots is taken (update.py:103)
 *   delta [h*w] fp32 out (may be NULL)                               stage 0/1 selects delta{stage} */
int cer_update_step(const void* blob, void* workspace, void* net, const void* inp, float* disp,
                    const float* corr, int slots, float* delta, int apply_delta, int stage, int h, int w,
                    cer_stream_t stream);

/* Which tensor-core path the 3x3 convolutions use (A/B switch; every GPU test runs on both):
 *   1 = default (CER_CONV unset): tcgen05.mma + TMEM, persistent CTAs, TMA operand loads; the gate conv (N = 192) as CTA
 *       pairs issuing cta_group::2 MMAs (M = 256, each CTA holds half of every weight tile), the delta conv (N = 256) as
 *       couples of CTAs that each keep one 128-channel half of the weights resident, the N = 64 convs one 128-pixel
 *       tile per CTA with resident weights (csrc/update_tc.cu)
 *   0 = mma.sync (the v1 kernels, csrc/update_hmma.cu; CER_CONV=hmma).
 * Takes effect for launches and graph captures issued afterwards (a plan that has already captured its graphs keeps
 * the variant it captured). */
int cer_set_conv_variant(int variant);

/* Which lookup kernels are used (A/B switch):
 *   2 = default: warp-autonomous kernels for the reference configuration (radius 5, 3 levels, D = 64 / 44).  The
 *       drop-in lookup (cer_lookup) is bit-identical to the general kernel.  The plan's fused lookup + 1x1-encoder kernel
 *       (cer_lookup_encode) evaluates the reference's normalise / unnormalise round trip once per pyramid level and shares
 *       floor and weights between the eleven taps of the level: <= 1 ulp of the coordinate per tap position, i.e. rare
 *       one-ulp flips of its fp16 output against the general kernel (tests/test_gpu_lookup_encode.py)
 *   1 = the general kernels only: the reference's per-tap arithmetic, any D / radius (CER_LOOKUP=general).  The drop-in
 *       classes always compute this arithmetic; a plan on variant 1 is bit-identical to them. */
int cer_set_lookup_variant(int variant);

/* Debug: per-role wait-cycle counters of the tcgen05 convolutions (tools/conv_roles.py). dev_buf: 4 x 32 uint64. */
int cer_debug_set_conv_profile(void* dev_buf);

/* ---- image preparation, depth output, multi-resolution merge (csrc/io_ops.cu) ---------------- */

/* images *= 2 / 255.; images -= 1 (core/raft.py:40-41), n floats, src may equal dst. */
int cer_normalize_images(const float* src, float* dst, long long n, cer_stream_t stream);

/* F.interpolate(images, [h2, w2], mode='bilinear', align_corners=True) of `planes` h x w planes
 * (scale_operation, utils/data_utils.py:58-66: the rescale=2 pass of the cascaded configuration). */
int cer_resize_bilinear_ac(const float* src, float* dst, int planes, int h, int w, int h2, int w2, cer_stream_t stream);

/* depth = where(disp == 0, 0, 1 / disp) (inference.py:57-58); flip_rows != 0 stores row y at row h-1-y, the order
 * write_pfm puts on disk (utils/frame_utils.py:145).  In place (depth == disp) only without the flip. */
int cer_disp_to_depth(const float* disp, float* depth, int h, int w, int flip_rows, cer_stream_t stream);

/* multires.py:26-28: im1r = cv2.resize(im1, (w2, h2)) (INTER_LINEAR); out = |im1r - im2| < th * im1r ? im2 : im1r. */
int cer_multires_merge(const float* im1, int h1, int w1, const float* im2, int h2, int w2, float th, float* out,
                       cer_stream_t stream);

/* ---- geometric-consistency filter of fusion.py (csrc/fusion_ops.cu) ----------------------- */

/* reproject_with_depth + check_geometric_consistency (fusion.py:39-106) for one reference view against n_src (<= 10)
 * source views, and the per-view aggregation of fusion() (fusion.py:239-249), in one kernel.
 *   depth_ref [h,w], K_ref [3,3], E_ref [4,4] (world -> camera); depth_src [n_src,h,w], K_src [n_src,3,3],
 *   E_src [n_src,4,4]; thre1 / thre2 as passed to check_geometric_consistency (masks use i/thre1, i/thre2, i = 2..10);
 *   mats_ws: device scratch of cer_geo_mats_bytes(n_src) bytes.
 * Per-source outputs (all five or none, may be NULL): masks [9,n_src,h,w] uint8, depth_reprojected [n_src,h,w]
 * (zero where the i = 10 mask fails), x_src / y_src / rel_diff [n_src,h,w].
 * Aggregated outputs (each may be NULL): geo_mask [h,w] uint8, depth_est [h,w], n_valid (device int = geo_mask.sum()). */
size_t cer_geo_mats_bytes(int n_src);
int cer_geometric_filter(const float* depth_ref, const float* K_ref, const float* E_ref, const float* depth_src,
                         const float* K_src, const float* E_src, int n_src, int h, int w, double thre1, double thre2,
                         void* mats_ws, unsigned char* masks, float* depth_reprojected, float* x_src, float* y_src,
                         float* rel_diff, unsigned char* geo_mask, float* depth_est, int* n_valid, cer_stream_t stream);

/* CorrBlock.__call__ fused with the first corr-encoder layer (core/corr.py:102-143 + core/update.py:62 `corr_encoder[0:2]`
 * under autocast): e1[p][0..63] = fp16 relu(conv1x1(lookup(volume, disp)[p])) from the view-mean volume [h*w][D] fp32 of
 * cer_build_volume, origin [h*w], disp [h*w] fp32; e1 [h*w][64] fp16 NHWC.  The 33 correlation planes never reach HBM.
 * D = 64 and D = 44 (the reference's cascade widths) take the warp-autonomous kernel, any other D the general one. */
int cer_lookup_encode(const void* blob, const float* volume, const float* origin, float* disp, int D, float incre, int h,
                      int w, void* e1, cer_stream_t stream);

/* ConvGRU.forward alone (core/update.py:17-25): net [h*w,64] fp16 NHWC updated in place from
 * inputs inp [h*w,64], dn [h*w,64] (49 disparity-encoder channels + 15 zero), e [h*w,64], all fp16 NHWC. */
int cer_gru_step(const void* blob, void* workspace, void* net, const void* inp, const void* dn, const void* e,
                 int h, int w, cer_stream_t stream);

/* ---- the whole hot path (core/raft.py:75-108) behind one handle ----------------------------- */
typedef struct cer_plan cer_plan;

typedef struct cer_plan_config {
  int h, w;            /* quarter-resolution grid (H/4, W/4) */
  int max_views;       /* largest number of source views */
  int n_stages;        /* <= 4 */
  int D[4];            /* hypotheses per stage (core/raft.py:77-80) */
  double incre[4];     /* 0.0025 / N per stage (core/raft.py:81), the Python double */
  int iters[4];        /* GRU iterations per stage */
  int feats_f16;       /* 1: keep features in fp16 (lossless for autocast fnet output), 0: fp32 */
  int use_graph;       /* 1: capture each stage's iteration loop in a CUDA graph */
} cer_plan_config;

/* Allocates the plan's device workspace (cudaMalloc) on the current device. */
int cer_plan_create(const cer_plan_config* cfg, cer_plan** out);
void cer_plan_destroy(cer_plan* plan);
size_t cer_plan_workspace_bytes(const cer_plan* plan);
/* Copies a HOST blob (cer_pack_update_weights) to the device; synchronous. */
int cer_plan_set_weights(cer_plan* plan, const void* blob_host);

/* Device-resident inputs.  fmaps [n_views+1,64,h,w] NCHW (fp16 if fmaps_f16 else fp32), view 0 is the
 * reference image; net, inp [64,h,w] NCHW (same dtype flag ctx_f16); poses [n_views+1,4,4] fp32 already
 * multiplied by scale (core/raft.py:35); intrinsics [n_views+1,3,3] fp32 already divided by 4
 * (core/raft.py:39).  view_begin/view_end select the source views this rank builds (0..n_views for
 * one GPU); when the range is a strict subset the partial volume of every stage is left in
 * cer_plan_partial_volume() after cer_plan_build_stage() for the caller to all-reduce.
 * disp_out [h,w] fp32 = disparity after the last iteration times out_scale (core/raft.py:108). */
int cer_plan_run_device(cer_plan* plan, const void* fmaps, int fmaps_f16, const void* net, const void* inp,
                        int ctx_f16, const float* poses, const float* intrinsics, int n_views,
                        float out_scale, float* disp_out, cer_stream_t stream);

/* Same, with HOST buffers: pinned-or-pageable host memory in, host memory out; the call copies
 * host->device, runs, copies the disparity back and synchronises the stream (inference.py:43-57). */
int cer_plan_run_host(cer_plan* plan, const void* fmaps, int fmaps_f16, const void* net, const void* inp,
                      int ctx_f16, const float* poses, const float* intrinsics, int n_views,
                      float out_scale, float* disp_out, cer_stream_t stream);

/* Pipelined form of cer_plan_run_host for throughput: submit returns as soon as the work is enqueued (the host->device
 * copy of this job runs on a private copy stream and overlaps the kernels of the previous job; the disparity is copied
 * back on the same copy stream).  At most two jobs are in flight (submit waits for the oldest one otherwise).  Host
 * buffers should be pinned and must stay valid until the matching cer_plan_wait_host() returns; jobs complete in
 * submission order. */
int cer_plan_submit_host(cer_plan* plan, const void* fmaps, int fmaps_f16, const void* net, const void* inp,
                         int ctx_f16, const float* poses, const float* intrinsics, int n_views, float out_scale,
                         float* disp_out_host, cer_stream_t stream);
int cer_plan_wait_host(cer_plan* plan);

/* Stage-wise API for sharded multi-GPU runs: prepare (the views this rank touches; an empty
 * range converts the reference image and the context only) -> for each stage { build_stage_units (this rank's run of
 * (view, hypothesis) units, scaled 1/total_views) ; [caller all-reduces partial_volume] ; iterate_stage } -> finish. */
int cer_plan_prepare(cer_plan* plan, const void* fmaps, int fmaps_f16, const void* net, const void* inp,
                     int ctx_f16, const float* poses, const float* intrinsics, int n_views,
                     int view_begin, int view_end, cer_stream_t stream);
/* The plan's own input buffers, for producers that write the kernels' layouts directly (cer_encoder_forward): feature
 * image i (0 = reference image, 1.. = source views) as [h*w][64] fp16 ALREADY scaled by 1/8 (core/corr.py:30-31), net /
 * inp as [h*w][64] fp16.  Fill them, then call cer_plan_prepare_inplace (projection matrices, disp = 0) in place of
 * cer_plan_prepare: no layout kernels run between the encoders and the hot path. */
void* cer_plan_feature_buffer(cer_plan* plan, int image);
void* cer_plan_net_buffer(cer_plan* plan);
void* cer_plan_inp_buffer(cer_plan* plan);
int cer_plan_prepare_inplace(cer_plan* plan, const float* poses, const float* intrinsics, int n_views, int view_begin,
                             int view_end, cer_stream_t stream);
int cer_plan_build_stage(cer_plan* plan, int stage, cer_stream_t stream);
/* Sharded build: units [unit_begin, unit_end) of the stage's n_views * D (view, hypothesis) units, view-major, into a
 * zeroed partial volume (cer_plan_partial_volume); the views they touch must have been prepared.  An empty range only
 * zeroes the volume and writes the hypothesis origin (every rank runs the lookups). */
int cer_plan_build_stage_units(cer_plan* plan, int stage, long long unit_begin, long long unit_end, cer_stream_t stream);
float* cer_plan_partial_volume(cer_plan* plan, int stage, size_t* n_floats);
int cer_plan_iterate_stage(cer_plan* plan, int stage, cer_stream_t stream);
int cer_plan_finish(cer_plan* plan, float out_scale, float* disp_out, cer_stream_t stream);
/* Per-kernel timing for roofline accounting: when enabled the plan runs eagerly (no graph) and brackets
 * every kernel launch with CUDA events on the launching stream.  cer_plan_kernel_times() synchronises
 * the stream, returns accumulated milliseconds and launch counts per kernel class and resets them.
 * Classes (index): 0 layout, 1 projection, 2 volume build, 3 pool, 4 lookup, 5 corr drop-in,
 * 6 disp encoder, 7 corr encoder 1x1, 8 conv corr-encoder 3x3, 9 conv gates, 10 conv q + GRU,
 * 11 conv delta, 12 disp update, 13 finish. */
#define CER_KERNEL_KINDS 14
int cer_plan_set_kernel_timing(cer_plan* plan, int enable);
int cer_plan_kernel_times(cer_plan* plan, double* ms_by_kind, long long* launches_by_kind, int n_kinds,
                          cer_stream_t stream);
/* Number of kernel launches issued by the last run (graph nodes count as launches). */
long long cer_plan_last_launch_count(const cer_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* CER_MVS_B200_H */
