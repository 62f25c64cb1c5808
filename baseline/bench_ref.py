"""Timed legs of the UNMODIFIED reference for bench.py (BASELINE / TEST INFRASTRUCTURE, not product code).

* ``reference_gpu``   -- SURVEY.md section 8(d)(i): the reference's Python + its own alt_cuda_corr kernel + real autocast
  on the same B200, hot path only (stub encoders), CUDA-event timed.  Denominator of the north star's ">= 4x reference
  single-GPU".
* ``corr_kernel_legs`` -- ``alt_cuda_corr_ref.forward`` per source view ("bar to beat on B200 = this kernel
  compiled for sm_100") next to the drop-in kernel and the fused build.
* ``reference_cpu``   -- the reference's Python on the host cores (``bench.py --impl reference`` and ``cpu_baseline``).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import refrun  # noqa: E402
from cer_mvs_b200 import synth  # noqa: E402

t = torch.from_numpy


def reference_gpu(H, W, V, sc, sd, cascades, device, steps=2, warmup=1):
    """{name: ms per depth map} of the reference's RAFT.forward hot path on `device`, one entry per cascade."""
    ref = refrun.import_reference("gpu")
    refrun.restore_reference_classes()
    fm16 = t(sc["fmaps"]).to(device).half()
    pre16 = t(synth.make_context_pre(H // 4, W // 4, seed=0)).to(device).half()
    images = torch.zeros(1, V + 1, 3, H, W, device=device)
    poses, K = t(sc["poses"]).to(device), t(sc["intrinsics"]).to(device)
    out = {}
    for name, cascade in cascades.items():
        model = refrun.make_model(ref, sd, cascade, fm16, pre16, device)
        ms, _ = refrun.time_forward(model, images, poses, K, 1.0, steps=steps, warmup=warmup)
        out[name] = ms
        del model
    torch.cuda.empty_cache()
    return out


def corr_kernel_legs(H, W, V, sc, device, n=5):
    """Per-view correlation: the reference's kernel vs the drop-in kernel (same fp32 inputs, same coords), and the
    fused build (all V views, fp16 features) -- microseconds, CUDA events."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    import cer_mvs_b200.alt_cuda_corr as ours
    from cer_mvs_b200.corr import CorrBlock
    ext = build_ref.load()
    h1, w1 = H // 4, W // 4
    fm = t(sc["fmaps"]).to(device)
    f1 = (fm[0, 0].permute(1, 2, 0) / 8.0).contiguous()[None]
    f2 = (fm[0, 1].permute(1, 2, 0) / 8.0).contiguous()[None]
    poses, K = t(sc["poses"]).to(device), t(sc["intrinsics"]).to(device).clone()
    K[:, :, :2] /= 4
    ii, jj = torch.zeros(V, dtype=torch.long, device=device), torch.arange(1, V + 1, device=device)
    res = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=n):
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / reps

    for stage, (D, incre, shift) in enumerate([(64, 0.0025 / 64, True), (44, 0.0025 / 320, False)]):
        disp = torch.zeros(1, 1, h1, w1, device=device) if stage == 0 else t(sc["true_disp"]).to(device)[None, None]
        cb = CorrBlock(fm.half(), poses, K, ii, jj, nIncre=D, incre=incre, disps_input=disp, shift=shift, num_levels=3,
                       radius=5, test_mode=True, do_report=False)
        # coords of view 1 exactly as core/corr.py:86-88 feeds them to the extension
        d = (torch.arange(D, device=device) - D // 2).float() * incre
        dd = d.view(1, D, 1, 1) + cb.disps_origin.view(1, 1, h1, w1)
        ys, xs = torch.meshgrid(torch.arange(h1, device=device).float(), torch.arange(w1, device=device).float(),
                                indexing="ij")
        X = torch.stack([xs.expand_as(dd), ys.expand_as(dd), torch.ones_like(dd), dd], -1)
        P = cb.Pij[0].view(4, 4)
        x1 = X @ P.T
        coords = (x1[..., :2] / x1[..., 2:3]).clamp(-1e4, 1e4).contiguous()
        res[f"stage{stage}_reference_kernel_per_view_us"] = timed(lambda: ext.forward(f1, f2, coords, 0))
        res[f"stage{stage}_dropin_kernel_per_view_us"] = timed(lambda: ours.forward(f1, f2, coords, 0))
        res[f"stage{stage}_fused_build_all_views_us"] = timed(lambda: CorrBlock(
            fm.half(), poses, K, ii, jj, nIncre=D, incre=incre, disps_input=disp, shift=shift, num_levels=3, radius=5,
            test_mode=True, do_report=False))
        res[f"stage{stage}_fused_build_per_view_us"] = res[f"stage{stage}_fused_build_all_views_us"] / V
        del cb, coords, x1, X
    res["note"] = ("reference kernel = correlation_kernel.cu compiled for sm_100 (oracle/_ref), coords precomputed; "
                   "the reference additionally spends ~10 torch ops per view on projective_transform / permutes "
                   "(core/corr.py:84-91) that the fused build contains")
    return res


def reference_cpu(H, W, V, cascade, rows, iters=(1, 1), threads=None, seed=0):
    """The reference's RAFT.forward on the host cores on a bounded sample: a band of ``rows`` feature rows (all
    columns, all views), ``iters`` iterations per stage; the cost-volume build and the iterations are timed
    separately (a timing subclass of the reference's CorrBlock) and extrapolated linearly in pixels and iterations to
    one full depth map of ``cascade``.  Returns (seconds per full depth map, description, threads)."""
    if threads:
        torch.set_num_threads(int(threads))
    ref = refrun.import_reference("cpu")
    h1 = H // 4
    rows = min(rows, h1)
    sc = synth.make_scene(4 * rows, W, V, seed=seed)
    sd = synth.make_update_weights(seed=0, delta_scale=0.1, delta_bias=0.005)
    pre = t(synth.make_context_pre(rows, W // 4, seed=seed))
    small = [(cascade[0][0], cascade[0][1], iters[0]), (cascade[1][0], cascade[1][1], iters[1])]
    model = refrun.make_model(ref, sd, small, t(sc["fmaps"]), pre, "cpu")
    acc = {"build": 0.0}
    Base = ref.corr.CorrBlock

    class TimedCorrBlock(Base):
        def __init__(self, *a, **k):
            t0 = time.perf_counter()
            super().__init__(*a, **k)
            acc["build"] += time.perf_counter() - t0

    ref.raft.CorrBlock = TimedCorrBlock
    try:
        images = torch.zeros(1, V + 1, 3, 4 * rows, W)
        t0 = time.perf_counter()
        refrun.run_forward(model, images, t(sc["poses"]), t(sc["intrinsics"]), 1.0)
        total = time.perf_counter() - t0
    finally:
        ref.raft.CorrBlock = Base
    t_build, t_iter = acc["build"], (total - acc["build"]) / max(sum(iters), 1)
    spx = h1 / rows
    n_it = sum(c[2] for c in cascade)
    full = (t_build + t_iter * n_it) * spx
    desc = (f"reference Python (core/raft.py, corr.py, update.py; alt_cuda_corr.forward served by the oracle's CPU "
            f"restatement), fp32, rows 0..{rows - 1} of {h1} x {W // 4} cols x {V} views, both volume builds + "
            f"{iters[0]}+{iters[1]} of {cascade[0][2]}+{cascade[1][2]} iterations, extrapolated linearly in pixels and "
            f"iterations ({total:.1f}s of CPU work)")
    return full, desc, torch.get_num_threads()


def reference_cpu_cfg1(threads=None):
    """One UN-EXTRAPOLATED depth map of BASELINE configs[0] (448x576, 2 views, 2+2 iterations, fp32) on the host."""
    if threads:
        torch.set_num_threads(int(threads))
    ref = refrun.import_reference("cpu")
    H, W, V = synth.CONFIGS["cfg1_dtu_448x576_v2"]
    cascade = [(64, 64, 2), (-1, 320, 2)]
    sc = synth.make_scene(H, W, V, seed=21)
    sd = synth.make_update_weights(seed=21, delta_scale=0.1, delta_bias=0.02)
    pre = t(synth.make_context_pre(H // 4, W // 4, seed=21))
    model = refrun.make_model(ref, sd, cascade, t(sc["fmaps"]), pre, "cpu")
    images = torch.zeros(1, V + 1, 3, H, W)
    t0 = time.perf_counter()
    refrun.run_forward(model, images, t(sc["poses"]), t(sc["intrinsics"]), 1.0)
    return time.perf_counter() - t0
