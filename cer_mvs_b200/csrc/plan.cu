// cer_plan: the stage / iteration loop of RAFT.forward (core/raft.py:75-108) as a native object
// that owns its device workspace, keeps every intermediate resident in HBM in the kernels' own
// layouts (NHWC fp16 activations, D-minor fp32 volume) and replays each stage's iteration loop
// from a CUDA graph.
#include <vector>

#include "common.cuh"
#include "update_blob.h"

namespace cer {

// ---- per-kernel CUDA-event timing (eager mode only; used by bench.py for the roofline numbers) ----
struct KernelTimer {
  std::vector<cudaEvent_t> ev;   // pairs
  std::vector<int> kinds;
  size_t used = 0;
  double ms[KK_COUNT] = {0};
  long long count[KK_COUNT] = {0};
};
thread_local KernelTimer* g_timer = nullptr;

void timer_begin(int kind, cudaStream_t s) {
  KernelTimer* t = g_timer;
  if (t->used + 2 > t->ev.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      t->ev.push_back(e);
    }
  }
  t->kinds.push_back(kind);
  cudaEventRecord(t->ev[t->used], s);
}
void timer_end(cudaStream_t s) {
  KernelTimer* t = g_timer;
  cudaEventRecord(t->ev[t->used + 1], s);
  t->used += 2;
}
static void timer_collect(KernelTimer* t) {   // caller has synchronised the stream
  for (size_t i = 0; i < t->used; i += 2) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t->ev[i], t->ev[i + 1]);
    t->ms[t->kinds[i / 2]] += ms;
    t->count[t->kinds[i / 2]] += 1;
  }
  t->used = 0;
  t->kinds.clear();
}

int update_step_hmma(const void* blob, void* workspace, void* net, const void* inp, float* disp, const float* corr,
                     int slots, float* delta, int apply_delta, int stage, int h, int w, cudaStream_t stream);
int update_configure();
int update_iteration_fused(const void* blob, void* workspace, void* net, const void* inp, float* disp,
                           const float* volume, const float* origin, int D, float incre, int apply_prev, int iter,
                           int stage, int h, int w, cudaStream_t stream);
int update_reset_flags(void* workspace, int iters, int h, int w, cudaStream_t stream);
int update_apply_delta(const void* blob, void* workspace, float* disp, int stage, int h, int w, cudaStream_t stream);

__global__ void scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, float s, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * s;
}

// hypothesis origin of a stage (core/corr.py:59-63), for ranks of a sharded build that own no unit of it
__global__ void origin_kernel(const float* __restrict__ disp, int shift, float lo, float* __restrict__ origin, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float d = disp[i];
    origin[i] = shift ? (d < lo ? lo : d) : d;
  }
}

__global__ void iota_pairs_kernel(int* ii, int* jj, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    ii[i] = 0;
    jj[i] = i + 1;
  }
}
}  // namespace cer

using namespace cer;

struct cer_plan {
  cer_plan_config cfg;
  long long px = 0;
  int Dmax = 0;
  // device buffers
  void* feats = nullptr;      // [(max_views+1)][px][64] fp16|fp32, pre-scaled by 1/8
  __half* net = nullptr;      // [px][64]
  __half* inp = nullptr;      // [px][64]
  float* disp = nullptr;      // [px]
  float* origin = nullptr;    // [px]
  float* volume = nullptr;    // [px][Dmax]
  float* corr = nullptr;      // [33][px]
  float* Pij = nullptr;       // [max_views][16]
  float* poses = nullptr;     // staging for run_host
  float* intr = nullptr;
  int *ii = nullptr, *jj = nullptr;
  void* ws = nullptr;         // update workspace
  void* blob = nullptr;
  void* stage_in = nullptr;   // host-path staging of NCHW inputs
  size_t stage_in_bytes = 0;
  // pipelined host path (cer_plan_submit_host / cer_plan_wait_host): two staging sets, copies on a private stream
  void* pipe_in[2] = {nullptr, nullptr};
  size_t pipe_in_bytes[2] = {0, 0};
  float* pipe_cam[2] = {nullptr, nullptr};     // poses + intrinsics
  float* pipe_out[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr, out_stream = nullptr;   // H2D and D2H on separate streams (both overlap compute)
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_compute[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  long long submitted = 0, waited = 0;
  size_t total_bytes = 0;
  // run state
  int n_views = 0, vb = 0, ve = 0;
  bool have_weights = false;
  cudaStream_t capture_stream = nullptr;   // graphs are captured here (the caller's stream may be the legacy one)
  cudaGraphExec_t graph[4] = {nullptr, nullptr, nullptr, nullptr};
  long long graph_nodes[4] = {0, 0, 0, 0};
  long long launches = 0;
  KernelTimer* timer = nullptr;   // non-null: run eagerly and time every kernel with CUDA events
};

static int plan_alloc(cer_plan* p, void** ptr, size_t bytes) {
  CER_CUDA(cudaMalloc(ptr, bytes));
  p->total_bytes += bytes;
  return CER_OK;
}

extern "C" {

int cer_plan_create(const cer_plan_config* cfg, cer_plan** out) {
  CER_REQUIRE(cfg && out, "cer_plan_create: null pointer");
  CER_REQUIRE(cfg->h > 0 && cfg->w > 0 && cfg->max_views > 0 && cfg->max_views <= 64, "cer_plan_create: bad sizes");
  CER_REQUIRE(cfg->n_stages >= 1 && cfg->n_stages <= 2,
              "cer_plan_create: 1 or 2 cascade stages supported (per-stage delta weights, core/update.py:67)");
  int rc = cer_device_check();
  if (rc) return rc;
  cer_plan* p = new cer_plan();
  p->cfg = *cfg;
  p->px = (long long)cfg->h * cfg->w;
  for (int s = 0; s < cfg->n_stages; ++s) {
    if (cfg->D[s] < 4 || cfg->D[s] > 1024 || cfg->iters[s] < 0 || !(cfg->incre[s] > 0)) {
      delete p;
      set_error("cer_plan_create: bad stage %d parameters", s);
      return CER_ERR_INVALID;
    }
    p->Dmax = cfg->D[s] > p->Dmax ? cfg->D[s] : p->Dmax;
  }
  const size_t fsz = cfg->feats_f16 ? 2 : 4;
  const long long px = p->px;
  bool ok = true;
  ok = ok && !plan_alloc(p, &p->feats, (size_t)(cfg->max_views + 1) * px * 64 * fsz);
  ok = ok && !plan_alloc(p, (void**)&p->net, px * 64 * 2);
  ok = ok && !plan_alloc(p, (void**)&p->inp, px * 64 * 2);
  ok = ok && !plan_alloc(p, (void**)&p->disp, px * 4);
  ok = ok && !plan_alloc(p, (void**)&p->origin, px * 4);
  ok = ok && !plan_alloc(p, (void**)&p->volume, px * p->Dmax * 4);
  ok = ok && !plan_alloc(p, (void**)&p->corr, px * kCorrPlanes * 4);
  ok = ok && !plan_alloc(p, (void**)&p->Pij, cfg->max_views * 16 * 4);
  ok = ok && !plan_alloc(p, (void**)&p->poses, (cfg->max_views + 1) * 16 * 4);
  ok = ok && !plan_alloc(p, (void**)&p->intr, (cfg->max_views + 1) * 9 * 4);
  ok = ok && !plan_alloc(p, (void**)&p->ii, cfg->max_views * 4);
  ok = ok && !plan_alloc(p, (void**)&p->jj, cfg->max_views * 4);
  ok = ok && !plan_alloc(p, &p->ws, cer_update_workspace_bytes(cfg->h, cfg->w));
  ok = ok && !plan_alloc(p, &p->blob, cer_update_blob_bytes());
  if (!ok) {
    cer_plan_destroy(p);
    return CER_ERR_INVALID;
  }
  iota_pairs_kernel<<<1, 64>>>(p->ii, p->jj, cfg->max_views);
  rc = update_configure();
  if (!rc) rc = (int)cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking);
  if (!rc) rc = (int)cudaDeviceSynchronize();
  if (rc) {
    cer_plan_destroy(p);
    return rc;
  }
  *out = p;
  return CER_OK;
}

void cer_plan_destroy(cer_plan* p) {
  if (!p) return;
  for (int s = 0; s < 4; ++s)
    if (p->graph[s]) cudaGraphExecDestroy(p->graph[s]);
  if (p->capture_stream) cudaStreamDestroy(p->capture_stream);
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  if (p->out_stream) cudaStreamDestroy(p->out_stream);
  for (int b = 0; b < 2; ++b) {
    if (p->pipe_in[b]) cudaFree(p->pipe_in[b]);
    if (p->pipe_cam[b]) cudaFree(p->pipe_cam[b]);
    if (p->pipe_out[b]) cudaFree(p->pipe_out[b]);
    if (p->ev_h2d[b]) cudaEventDestroy(p->ev_h2d[b]);
    if (p->ev_compute[b]) cudaEventDestroy(p->ev_compute[b]);
    if (p->ev_out[b]) cudaEventDestroy(p->ev_out[b]);
  }
  if (p->timer) {
    for (cudaEvent_t e : p->timer->ev) cudaEventDestroy(e);
    delete p->timer;
  }
  void* ptrs[] = {p->feats, p->net, p->inp, p->disp, p->origin, p->volume, p->corr, p->Pij, p->poses,
                  p->intr,  p->ii,  p->jj,  p->ws,   p->blob,   p->stage_in};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  delete p;
}

size_t cer_plan_workspace_bytes(const cer_plan* p) { return p ? p->total_bytes + p->stage_in_bytes : 0; }

int cer_plan_set_weights(cer_plan* p, const void* blob_host) {
  CER_REQUIRE(p && blob_host, "cer_plan_set_weights: null pointer");
  CER_CUDA(cudaMemcpy(p->blob, blob_host, cer_update_blob_bytes(), cudaMemcpyHostToDevice));
  p->have_weights = true;
  return CER_OK;
}

int cer_plan_prepare(cer_plan* p, const void* fmaps, int fmaps_f16, const void* net, const void* inp, int ctx_f16,
                     const float* poses, const float* intrinsics, int n_views, int view_begin, int view_end,
                     cer_stream_t stream) {
  CER_REQUIRE(p && fmaps && net && inp && poses && intrinsics, "cer_plan_prepare: null pointer");
  CER_REQUIRE(p->have_weights, "cer_plan_prepare: call cer_plan_set_weights first");
  CER_REQUIRE(n_views >= 1 && n_views <= p->cfg.max_views, "cer_plan_prepare: n_views %d outside 1..%d", n_views,
              p->cfg.max_views);
  CER_REQUIRE(view_begin >= 0 && view_begin <= view_end && view_end <= n_views, "cer_plan_prepare: bad view range");
  const int h = p->cfg.h, w = p->cfg.w;
  const long long px = p->px;
  const size_t in_sz = fmaps_f16 ? 2 : 4, f_sz = p->cfg.feats_f16 ? 2 : 4;
  cer::g_launches = 0;
  int rc;
  // reference image + the owned source views, scaled by 1/8 (core/corr.py:30-31)
  if ((rc = cer_nchw_to_nhwc(fmaps, fmaps_f16, p->feats, p->cfg.feats_f16, 1, 64, h, w, 0.125f, stream))) return rc;
  const size_t img_in = (size_t)px * 64 * in_sz, img_f = (size_t)px * 64 * f_sz;
  if (view_end > view_begin &&
      (rc = cer_nchw_to_nhwc((const char*)fmaps + (1 + view_begin) * img_in, fmaps_f16,
                             (char*)p->feats + (1 + view_begin) * img_f, p->cfg.feats_f16, view_end - view_begin, 64,
                             h, w, 0.125f, stream)))
    return rc;
  if ((rc = cer_nchw_to_nhwc(net, ctx_f16, p->net, 1, 1, 64, h, w, 1.f, stream))) return rc;
  if ((rc = cer_nchw_to_nhwc(inp, ctx_f16, p->inp, 1, 1, 64, h, w, 1.f, stream))) return rc;
  CER_CUDA(cudaMemsetAsync(p->disp, 0, px * 4, (cudaStream_t)stream));   // core/raft.py:52
  if (view_end > view_begin &&
      (rc = cer_projection_matrices(poses, intrinsics, p->ii + view_begin, p->jj + view_begin,
                                    view_end - view_begin, p->Pij + view_begin * 16, stream)))
    return rc;
  p->n_views = n_views;
  p->vb = view_begin;
  p->ve = view_end;
  p->launches = cer::g_launches + 1;  // + memset
  return CER_OK;
}

// The plan's own input buffers, for producers that write the kernels' layouts directly (the encoders of
// csrc/encoder.cu): feature image i (0 = reference image, 1.. = source views) as [px][64] fp16 ALREADY scaled by 1/8,
// net / inp as [px][64] fp16.  After filling them: cer_plan_prepare_inplace instead of cer_plan_prepare.
void* cer_plan_feature_buffer(cer_plan* p, int image) {
  if (!p || !p->cfg.feats_f16 || image < 0 || image > p->cfg.max_views) return nullptr;
  return (char*)p->feats + (size_t)image * p->px * 64 * 2;
}
void* cer_plan_net_buffer(cer_plan* p) { return p ? p->net : nullptr; }
void* cer_plan_inp_buffer(cer_plan* p) { return p ? p->inp : nullptr; }

int cer_plan_prepare_inplace(cer_plan* p, const float* poses, const float* intrinsics, int n_views, int view_begin,
                             int view_end, cer_stream_t stream) {
  CER_REQUIRE(p && poses && intrinsics, "cer_plan_prepare_inplace: null pointer");
  CER_REQUIRE(p->have_weights, "cer_plan_prepare_inplace: call cer_plan_set_weights first");
  CER_REQUIRE(p->cfg.feats_f16, "cer_plan_prepare_inplace: needs a plan with fp16 features");
  CER_REQUIRE(n_views >= 1 && n_views <= p->cfg.max_views, "cer_plan_prepare_inplace: n_views %d outside 1..%d", n_views,
              p->cfg.max_views);
  CER_REQUIRE(view_begin >= 0 && view_begin <= view_end && view_end <= n_views, "cer_plan_prepare_inplace: bad view range");
  cer::g_launches = 0;
  CER_CUDA(cudaMemsetAsync(p->disp, 0, p->px * 4, (cudaStream_t)stream));   // core/raft.py:52
  int rc;
  if (view_end > view_begin &&
      (rc = cer_projection_matrices(poses, intrinsics, p->ii + view_begin, p->jj + view_begin, view_end - view_begin,
                                    p->Pij + view_begin * 16, stream)))
    return rc;
  p->n_views = n_views;
  p->vb = view_begin;
  p->ve = view_end;
  p->launches = cer::g_launches + 1;
  return CER_OK;
}

int cer_plan_build_stage(cer_plan* p, int s, cer_stream_t stream) {
  CER_REQUIRE(p && s >= 0 && s < p->cfg.n_stages, "cer_plan_build_stage: bad stage");
  const int D = p->cfg.D[s];
  const double incre = (double)p->cfg.incre[s];
  const float lo = (float)((D / 2) * incre);   // torch.tensor(nIncre // 2 * incre).float(), core/corr.py:60
  cer::g_launches = 0;
  int rc = cer_build_volume(p->feats, p->cfg.feats_f16, p->Pij + p->vb * 16, p->ii + p->vb, p->jj + p->vb,
                            p->ve - p->vb, p->disp, s == 0, D, (float)incre, lo, p->origin, p->volume,
                            1.f / (float)p->n_views, 0, p->cfg.h, p->cfg.w, stream);
  p->launches += cer::g_launches;
  return rc;
}

// Sharded build: the work of a stage is n_views * D units (view, hypothesis), view-major; a rank
// builds the contiguous run [unit_begin, unit_end) into a zeroed partial volume -- at most three kernel calls (tail of
// the first view, whole views, head of the last view), each adding to the volume -- and the partial volumes of all
// ranks are summed by one all-reduce.  Works for any number of ranks (more ranks than views: cfg 5 of BASELINE.json).
int cer_plan_build_stage_units(cer_plan* p, int s, long long unit_begin, long long unit_end, cer_stream_t stream) {
  CER_REQUIRE(p && s >= 0 && s < p->cfg.n_stages, "cer_plan_build_stage_units: bad stage");
  CER_REQUIRE(p->cfg.feats_f16, "cer_plan_build_stage_units: needs a plan with fp16 features");
  const int D = p->cfg.D[s];
  const long long U = (long long)p->n_views * D;
  CER_REQUIRE(unit_begin >= 0 && unit_begin <= unit_end && unit_end <= U, "cer_plan_build_stage_units: units [%lld, %lld) "
              "outside 0..%lld", unit_begin, unit_end, U);
  const double incre = (double)p->cfg.incre[s];
  const float lo = (float)((D / 2) * incre);
  cer::g_launches = 0;
  CER_CUDA(cudaMemsetAsync(p->volume, 0, (size_t)p->px * D * 4, (cudaStream_t)stream));
  p->launches += 1;
  int rc = CER_OK;
  if (unit_end > unit_begin) {
    const int v0 = (int)(unit_begin / D), a = (int)(unit_begin % D);
    const int v1 = (int)((unit_end - 1) / D), b = (int)((unit_end - 1) % D) + 1;
    CER_REQUIRE(v0 >= p->vb && v1 < p->ve, "cer_plan_build_stage_units: views %d..%d were not prepared (prepared %d..%d)",
                v0, v1, p->vb, p->ve - 1);
    auto part = [&](int vb, int ve, int d0, int d1) {
      return cer_build_volume_part(p->feats, 1, p->Pij + vb * 16, p->ii + vb, p->jj + vb, ve - vb, p->disp, s == 0, D,
                                   (float)incre, lo, p->origin, p->volume, 1.f / (float)p->n_views, 0, p->cfg.h,
                                   p->cfg.w, d0, d1, 1, stream);
    };
    if (v0 == v1) {
      rc = part(v0, v0 + 1, a, b);
    } else {
      int first_whole = v0, last_whole = v1;       // whole views [first_whole, last_whole)
      if (a > 0) {
        rc = part(v0, v0 + 1, a, D);
        first_whole = v0 + 1;
      }
      if (b < D) {
        if (!rc) rc = part(v1, v1 + 1, 0, b);
      } else {
        last_whole = v1 + 1;
      }
      if (!rc && last_whole > first_whole) rc = part(first_whole, last_whole, 0, D);
    }
  } else {
    // a rank without units still needs the hypothesis origin (every rank runs the lookups)
    CER_LAUNCH(KK_BUILD, origin_kernel, ceil_div(p->px, 256), 256, 0, stream, p->disp, s == 0, lo, p->origin, p->px);
    rc = check_launch("cer_plan_build_stage_units (origin)");
  }
  p->launches += cer::g_launches;
  return rc;
}

float* cer_plan_partial_volume(cer_plan* p, int s, size_t* n_floats) {
  if (!p || s < 0 || s >= p->cfg.n_stages) return nullptr;
  if (n_floats) *n_floats = (size_t)p->px * p->cfg.D[s];
  return p->volume;
}

static int issue_iterations(cer_plan* p, int s, cudaStream_t stream) {
  const int D = p->cfg.D[s];
  const float incre = (float)p->cfg.incre[s];
  int rc0 = update_reset_flags(p->ws, p->cfg.iters[s], p->cfg.h, p->cfg.w, stream);
  if (rc0) return rc0;
  for (int it = 0; it < p->cfg.iters[s]; ++it) {
    // lookup of the current disparity fused with the corr encoder; the delta of the previous iteration is
    // applied inside the same kernel (core/raft.py:99-101)
    int rc = update_iteration_fused(p->blob, p->ws, p->net, p->inp, p->disp, p->volume, p->origin, D, incre,
                                    it > 0, it, s, p->cfg.h, p->cfg.w, stream);
    if (rc) return rc;
  }
  return update_apply_delta(p->blob, p->ws, p->disp, s, p->cfg.h, p->cfg.w, stream);
}

int cer_plan_iterate_stage(cer_plan* p, int s, cer_stream_t stream_) {
  CER_REQUIRE(p && s >= 0 && s < p->cfg.n_stages, "cer_plan_iterate_stage: bad stage");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (p->cfg.iters[s] == 0) return CER_OK;
  if (!p->cfg.use_graph || p->timer) {
    cer::g_launches = 0;
    int rc = issue_iterations(p, s, stream);
    p->launches += cer::g_launches;
    return rc;
  }
  if (!p->graph[s]) {
    cudaGraph_t g = nullptr;
    // capture on the plan's own stream (capturing records, it does not execute); replay on the caller's
    CER_CUDA(cudaStreamBeginCapture(p->capture_stream, cudaStreamCaptureModeThreadLocal));
    cer::g_launches = 0;
    int rc = issue_iterations(p, s, p->capture_stream);
    cudaError_t e = cudaStreamEndCapture(p->capture_stream, &g);
    if (rc || e != cudaSuccess) {
      if (g) cudaGraphDestroy(g);
      if (!rc) {
        set_error("graph capture failed: %s", cudaGetErrorString(e));
        rc = (int)e;
      }
      return rc;
    }
    p->graph_nodes[s] = cer::g_launches;
    CER_CUDA(cudaGraphInstantiate(&p->graph[s], g, 0));
    cudaGraphDestroy(g);
  }
  CER_CUDA(cudaGraphLaunch(p->graph[s], stream));
  p->launches += p->graph_nodes[s];
  return CER_OK;
}

int cer_plan_finish(cer_plan* p, float out_scale, float* disp_out, cer_stream_t stream) {
  CER_REQUIRE(p && disp_out, "cer_plan_finish: null pointer");
  CER_LAUNCH(KK_FINISH, scale_copy_kernel, ceil_div(p->px, 256), 256, 0, stream, p->disp, disp_out, out_scale, p->px);
  p->launches += 1;
  return check_launch("cer_plan_finish");
}

int cer_plan_run_device(cer_plan* p, const void* fmaps, int fmaps_f16, const void* net, const void* inp,
                        int ctx_f16, const float* poses, const float* intrinsics, int n_views, float out_scale,
                        float* disp_out, cer_stream_t stream) {
  CER_REQUIRE(p, "cer_plan_run_device: null plan");
  cer::g_timer = p->timer;
  int rc = cer_plan_prepare(p, fmaps, fmaps_f16, net, inp, ctx_f16, poses, intrinsics, n_views, 0, n_views, stream);
  for (int s = 0; !rc && s < p->cfg.n_stages; ++s) {
    rc = cer_plan_build_stage(p, s, stream);
    if (!rc) rc = cer_plan_iterate_stage(p, s, stream);
  }
  if (!rc) rc = cer_plan_finish(p, out_scale, disp_out, stream);
  cer::g_timer = nullptr;
  return rc;
}

int cer_plan_set_kernel_timing(cer_plan* p, int enable) {
  CER_REQUIRE(p, "cer_plan_set_kernel_timing: null plan");
  if (enable && !p->timer) p->timer = new KernelTimer();
  if (!enable && p->timer) {
    for (cudaEvent_t e : p->timer->ev) cudaEventDestroy(e);
    delete p->timer;
    p->timer = nullptr;
  }
  return CER_OK;
}

int cer_plan_kernel_times(cer_plan* p, double* ms_by_kind, long long* launches_by_kind, int n_kinds,
                          cer_stream_t stream) {
  CER_REQUIRE(p && p->timer && ms_by_kind && launches_by_kind, "cer_plan_kernel_times: timing not enabled");
  CER_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  timer_collect(p->timer);
  for (int k = 0; k < n_kinds && k < KK_COUNT; ++k) {
    ms_by_kind[k] = p->timer->ms[k];
    launches_by_kind[k] = p->timer->count[k];
    p->timer->ms[k] = 0;
    p->timer->count[k] = 0;
  }
  return CER_OK;
}

int cer_plan_run_host(cer_plan* p, const void* fmaps, int fmaps_f16, const void* net, const void* inp, int ctx_f16,
                      const float* poses, const float* intrinsics, int n_views, float out_scale, float* disp_out,
                      cer_stream_t stream_) {
  CER_REQUIRE(p && fmaps && net && inp && poses && intrinsics && disp_out, "cer_plan_run_host: null pointer");
  CER_REQUIRE(n_views >= 1 && n_views <= p->cfg.max_views, "cer_plan_run_host: bad n_views");
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long px = p->px;
  const size_t fm_bytes = (size_t)(n_views + 1) * 64 * px * (fmaps_f16 ? 2 : 4);
  const size_t ctx_bytes = (size_t)64 * px * (ctx_f16 ? 2 : 4);
  const size_t need = align256(fm_bytes) + 2 * align256(ctx_bytes) + align256(px * 4);
  if (need > p->stage_in_bytes) {
    if (p->stage_in) cudaFree(p->stage_in);
    p->stage_in = nullptr;
    p->stage_in_bytes = 0;
    CER_CUDA(cudaMalloc(&p->stage_in, need));
    p->stage_in_bytes = need;
  }
  char* d_fm = (char*)p->stage_in;
  char* d_net = d_fm + align256(fm_bytes);
  char* d_inp = d_net + align256(ctx_bytes);
  float* d_out = (float*)(d_inp + align256(ctx_bytes));
  CER_CUDA(cudaMemcpyAsync(d_fm, fmaps, fm_bytes, cudaMemcpyHostToDevice, stream));
  CER_CUDA(cudaMemcpyAsync(d_net, net, ctx_bytes, cudaMemcpyHostToDevice, stream));
  CER_CUDA(cudaMemcpyAsync(d_inp, inp, ctx_bytes, cudaMemcpyHostToDevice, stream));
  CER_CUDA(cudaMemcpyAsync(p->poses, poses, (n_views + 1) * 16 * 4, cudaMemcpyHostToDevice, stream));
  CER_CUDA(cudaMemcpyAsync(p->intr, intrinsics, (n_views + 1) * 9 * 4, cudaMemcpyHostToDevice, stream));
  int rc = cer_plan_run_device(p, d_fm, fmaps_f16, d_net, d_inp, ctx_f16, p->poses, p->intr, n_views, out_scale,
                               d_out, stream);
  if (rc) return rc;
  CER_CUDA(cudaMemcpyAsync(disp_out, d_out, px * 4, cudaMemcpyDeviceToHost, stream));
  CER_CUDA(cudaStreamSynchronize(stream));
  return CER_OK;
}

// ---- pipelined host path: the H2D copy of job i+1 and the D2H copy of job i overlap the kernels of job i ----
int cer_plan_submit_host(cer_plan* p, const void* fmaps, int fmaps_f16, const void* net, const void* inp, int ctx_f16,
                         const float* poses, const float* intrinsics, int n_views, float out_scale,
                         float* disp_out_host, cer_stream_t stream_) {
  CER_REQUIRE(p && fmaps && net && inp && poses && intrinsics && disp_out_host, "cer_plan_submit_host: null pointer");
  CER_REQUIRE(n_views >= 1 && n_views <= p->cfg.max_views, "cer_plan_submit_host: bad n_views");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (p->submitted - p->waited >= 2) {
    int rc = cer_plan_wait_host(p);
    if (rc) return rc;
  }
  const int b = (int)(p->submitted & 1);
  const long long px = p->px;
  const size_t fm_bytes = (size_t)(n_views + 1) * 64 * px * (fmaps_f16 ? 2 : 4);
  const size_t ctx_bytes = (size_t)64 * px * (ctx_f16 ? 2 : 4);
  const size_t need = align256(fm_bytes) + 2 * align256(ctx_bytes);
  if (!p->copy_stream) {
    CER_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    CER_CUDA(cudaStreamCreateWithFlags(&p->out_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CER_CUDA(cudaEventCreateWithFlags(&p->ev_h2d[i], cudaEventDisableTiming));
      CER_CUDA(cudaEventCreateWithFlags(&p->ev_compute[i], cudaEventDisableTiming));
      CER_CUDA(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
      CER_CUDA(cudaMalloc((void**)&p->pipe_cam[i], (p->cfg.max_views + 1) * 25 * 4));
      CER_CUDA(cudaMalloc((void**)&p->pipe_out[i], px * 4));
    }
  }
  if (need > p->pipe_in_bytes[b]) {
    CER_CUDA(cudaStreamSynchronize(p->copy_stream));
    CER_CUDA(cudaStreamSynchronize(p->out_stream));
    CER_CUDA(cudaStreamSynchronize(stream));
    if (p->pipe_in[b]) cudaFree(p->pipe_in[b]);
    p->pipe_in[b] = nullptr;
    p->pipe_in_bytes[b] = 0;
    CER_CUDA(cudaMalloc(&p->pipe_in[b], need));
    p->pipe_in_bytes[b] = need;
  }
  char* d_fm = (char*)p->pipe_in[b];
  char* d_net = d_fm + align256(fm_bytes);
  char* d_inp = d_net + align256(ctx_bytes);
  float* d_pose = p->pipe_cam[b];
  float* d_intr = d_pose + (p->cfg.max_views + 1) * 16;
  // copy stream: staging set b is free once the job that used it two submits ago has been computed
  if (p->submitted >= 2) CER_CUDA(cudaStreamWaitEvent(p->copy_stream, p->ev_compute[b], 0));
  CER_CUDA(cudaMemcpyAsync(d_fm, fmaps, fm_bytes, cudaMemcpyHostToDevice, p->copy_stream));
  CER_CUDA(cudaMemcpyAsync(d_net, net, ctx_bytes, cudaMemcpyHostToDevice, p->copy_stream));
  CER_CUDA(cudaMemcpyAsync(d_inp, inp, ctx_bytes, cudaMemcpyHostToDevice, p->copy_stream));
  CER_CUDA(cudaMemcpyAsync(d_pose, poses, (n_views + 1) * 16 * 4, cudaMemcpyHostToDevice, p->copy_stream));
  CER_CUDA(cudaMemcpyAsync(d_intr, intrinsics, (n_views + 1) * 9 * 4, cudaMemcpyHostToDevice, p->copy_stream));
  CER_CUDA(cudaEventRecord(p->ev_h2d[b], p->copy_stream));
  // compute stream
  CER_CUDA(cudaStreamWaitEvent(stream, p->ev_h2d[b], 0));
  if (p->submitted >= 2) CER_CUDA(cudaStreamWaitEvent(stream, p->ev_out[b], 0));   // pipe_out[b] has been read back
  int rc = cer_plan_run_device(p, d_fm, fmaps_f16, d_net, d_inp, ctx_f16, d_pose, d_intr, n_views, out_scale,
                               p->pipe_out[b], stream);
  if (rc) return rc;
  CER_CUDA(cudaEventRecord(p->ev_compute[b], stream));
  // result back on its own stream: holds up neither the next job's kernels nor its H2D copy
  CER_CUDA(cudaStreamWaitEvent(p->out_stream, p->ev_compute[b], 0));
  CER_CUDA(cudaMemcpyAsync(disp_out_host, p->pipe_out[b], px * 4, cudaMemcpyDeviceToHost, p->out_stream));
  CER_CUDA(cudaEventRecord(p->ev_out[b], p->out_stream));
  p->submitted += 1;
  return CER_OK;
}

int cer_plan_wait_host(cer_plan* p) {
  CER_REQUIRE(p, "cer_plan_wait_host: null plan");
  if (p->waited >= p->submitted) return CER_OK;
  const int b = (int)(p->waited & 1);
  CER_CUDA(cudaEventSynchronize(p->ev_out[b]));
  p->waited += 1;
  return CER_OK;
}

long long cer_plan_last_launch_count(const cer_plan* p) { return p ? p->launches : 0; }

}  // extern "C"
