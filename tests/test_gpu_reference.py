"""GPU: the CUDA path against the UNMODIFIED reference at BASELINE.json sizes.

* configs[0] at full size (448x576, 2 source views, 2+2 iterations, fp32): golden of the reference's own CPU run
  (oracle/gen_golden_cfg1.py) -- needs nothing but the committed fixture;
* configs[0] and configs[1] (1184x1600, 10 views, 16+16 iterations, fp16 autocast) against the reference RUN ON
  THIS GPU: its Python from baseline/_ref, its own alt_cuda_corr kernel from oracle/_ref, real
  torch.cuda.amp.autocast (baseline/refrun.py).  Skipped when those git-ignored artefacts did not travel;
* the reference's own RAFT.forward after cer_mvs_b200.install.install() == DepthHotPath (general lookup kernel), bit for bit.

North-star bar: relative L1 on disparity <= 1e-3; every test prints the number it achieved.
"""
import os
import sys

import numpy as np
import pytest
import torch

from cer_mvs_b200 import synth
from util import ROOT, rel_l1, t

sys.path.insert(0, os.path.join(ROOT, "baseline"))
import refrun  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-3

needs_ref = pytest.mark.skipif(not refrun.available("gpu"),
                               reason="baseline/_ref or oracle/_ref not present (built in the build container only)")


def _scene(cfg, seed, dscale, dbias):
    H, W, V = synth.CONFIGS[cfg]
    sc = synth.make_scene(H, W, V, seed=seed)
    sd = synth.make_update_weights(seed=seed, delta_scale=dscale, delta_bias=dbias)
    pre = synth.make_context_pre(H // 4, W // 4, seed=seed)
    return H, W, V, sc, sd, pre


def _ours(H, W, V, sc, sd, net, inp, cascade, dtype, scale=1.0, feats_f16=True):
    from cer_mvs_b200.hotpath import DepthHotPath
    hp = DepthHotPath(H // 4, W // 4, max_views=V, cascade=cascade, feats_f16=feats_f16)
    hp.load_update_block(sd)
    out = hp(t(sc["fmaps"]).cuda().to(dtype), net.to(dtype), inp.to(dtype), t(sc["poses"]).cuda(),
             t(sc["intrinsics"]).cuda(), scale=scale)
    return out.cpu().numpy().copy()


def test_cfg1_full_size_vs_reference_fp32_golden(golden):
    """BASELINE configs[0]: 448x576, 2 views, 2+2 iterations, fp32 -- the reference's own CPU run, un-extrapolated."""
    g = golden("e2e_fp32_cfg1")
    cascade = [tuple(int(v) for v in row) for row in g["cascade"]]
    H, W, V, sc, sd, pre = _scene("cfg1_dtu_448x576_v2", int(g["seed"]), float(g["delta_scale"]), float(g["delta_bias"]))
    pre = t(pre).cuda()
    net, inp = torch.tanh(pre[:, :, :64]), torch.relu(pre[:, :, 64:])
    out = _ours(H, W, V, sc, sd, net, inp, cascade, torch.float32, feats_f16=False)
    err = rel_l1(out, g["disp"])
    print(f"cfg1 448x576 V=2 2+2: rel L1 vs reference fp32 (CPU) = {err:.3e}")
    assert out.shape == g["disp"].shape
    assert err < TOL, err


def _reference_gpu(cfg, cascade, seed, dscale, dbias, scale=1.0):
    ref = refrun.import_reference("gpu")
    refrun.restore_reference_classes()
    H, W, V, sc, sd, pre = _scene(cfg, seed, dscale, dbias)
    dev = torch.device("cuda")
    fm16 = t(sc["fmaps"]).to(dev).half()            # what fnet emits under autocast (core/raft.py:55,66-69)
    pre16 = t(pre).to(dev).half()
    model = refrun.make_model(ref, sd, cascade, fm16, pre16, dev)
    images = torch.zeros(1, V + 1, 3, H, W, device=dev)
    want = refrun.run_forward(model, images, t(sc["poses"]).to(dev), t(sc["intrinsics"]).to(dev), scale)
    net, inp = torch.tanh(pre16[:, :, :64]), torch.relu(pre16[:, :, 64:])     # core/raft.py:58-60, fp16 under autocast
    return ref, (H, W, V, sc, sd), (fm16, pre16, net, inp, images), want


@needs_ref
@pytest.mark.parametrize("cfg,cascade,dbias", [
    ("cfg1_dtu_448x576_v2", [(64, 64, 2), (-1, 320, 2)], 0.02),
    ("cfg1_dtu_448x576_v2", [(64, 64, 8), (-1, 320, 8)], 0.01),
    ("cfg2_dtu_1184x1600_v10", [(64, 64, 16), (-1, 320, 16)], 0.005),
])
def test_vs_reference_run_on_this_gpu(cfg, cascade, dbias):
    """Reference Python + reference kernel + real autocast on this GPU vs DepthHotPath, same inputs and weights."""
    ref, (H, W, V, sc, sd), (fm16, pre16, net, inp, images), want = _reference_gpu(cfg, cascade, 31, 0.1, dbias)
    assert want.dtype == torch.float64                      # disp * scale with a float64 scale (core/raft.py:108)
    out = _ours(H, W, V, sc, sd, net, inp, cascade, torch.float16)
    err = rel_l1(out, want.cpu().numpy())
    print(f"{cfg} {cascade[0][2]}+{cascade[1][2]}: rel L1 vs the reference on this GPU (autocast) = {err:.3e}; "
          f"mean disp {float(want.mean()):.3e}")
    assert err < TOL, err


@needs_ref
def test_reference_raft_forward_after_install_is_depth_hot_path():
    """core/raft.py + inference-style call, unmodified, with cer_mvs_b200.install.install(): same bits as the plan."""
    import cer_mvs_b200.install as I
    cascade = [(64, 64, 3), (-1, 320, 3)]
    ref, (H, W, V, sc, sd), (fm16, pre16, net, inp, images), want = _reference_gpu("cfg1_dtu_448x576_v2", cascade, 32,
                                                                                    0.1, 0.02)
    try:
        I.install()
        assert ref.raft.CorrBlock.__module__.startswith("cer_mvs_b200")
        model = refrun.make_model(ref, sd, cascade, fm16, pre16, torch.device("cuda"))
        assert type(model.update_block).__module__.startswith("cer_mvs_b200")
        got = refrun.run_forward(model, images, t(sc["poses"]).cuda(), t(sc["intrinsics"]).cuda(), 1.0)
    finally:
        refrun.restore_reference_classes()
    assert got.dtype == torch.float64 and got.shape == want.shape
    # the drop-in classes run the general lookup kernel: the plan on that kernel is the same arithmetic, bit for bit; the
    # plan's default kernel shares floor and weights between the taps of a level (tests/test_gpu_lookup_encode.py)
    from cer_mvs_b200 import _lib
    try:
        _lib.check(_lib.lib().cer_set_lookup_variant(1))
        hot_general = _ours(H, W, V, sc, sd, net, inp, cascade, torch.float16)
    finally:
        _lib.lib().cer_set_lookup_variant(2)
    hot = _ours(H, W, V, sc, sd, net, inp, cascade, torch.float16)
    err = rel_l1(got.cpu().numpy(), want.cpu().numpy())
    gap = rel_l1(hot, hot_general)
    print(f"reference RAFT.forward with the drop-ins installed vs the stock reference: rel L1 = {err:.3e}; "
          f"plan default vs general lookup kernel {gap:.2e}")
    assert np.array_equal(got.cpu().numpy().astype(np.float32), hot_general)
    assert gap < 1e-4
    assert err < TOL, err


@needs_ref
def test_build_and_lookup_vs_reference_corrblock_full_size():
    """CorrBlock of the reference (its kernel, its pyramid, its 528 grid_samples) vs the drop-in at cfg-2 size."""
    from cer_mvs_b200.corr import CorrBlock
    ref = refrun.import_reference("gpu")
    refrun.restore_reference_classes()
    H, W, V = synth.CONFIGS["cfg2_dtu_1184x1600_v10"]
    V = 4                                             # 4 views keep the reference's per-view volumes small
    sc = synth.make_scene(H, W, V, seed=33)
    h1, w1 = H // 4, W // 4
    fmaps = t(sc["fmaps"]).cuda().half()
    poses, K = t(sc["poses"]).cuda(), t(sc["intrinsics"]).cuda().clone()
    K[:, :, :2] /= 4
    ii, jj = torch.zeros(V, dtype=torch.long).cuda(), torch.arange(1, V + 1).cuda()
    disp = t(sc["true_disp"]).cuda()[None, None].contiguous()
    for stage, (D, incre, shift, din) in enumerate([(64, 0.0025 / 64, True, torch.zeros_like(disp)),
                                                    (44, 0.0025 / 320, False, disp)]):
        kw = dict(nIncre=D, incre=incre, disps_input=din, shift=shift, num_levels=3, radius=5, test_mode=True,
                  do_report=False)
        with torch.no_grad():
            want_cb = ref.corr.CorrBlock(fmaps, poses, K, ii, jj, **kw)
            got_cb = CorrBlock(fmaps, poses, K, ii, jj, **kw)
            z = disp + 3.3 * incre
            want = want_cb(z[:, ii]).mean(dim=1)
            got = got_cb(z[:, ii]).mean(dim=1)
        err = rel_l1(got.cpu().numpy(), want.cpu().numpy())
        print(f"stage {stage}: lookup of the reference's CorrBlock vs drop-in at 296x400, {V} views: rel L1 = {err:.3e}")
        assert err < 2e-4, err
