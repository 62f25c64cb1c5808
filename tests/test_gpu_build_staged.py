"""GPU: the shared-memory-staged tcgen05 cost-volume build (csrc/build_volume_tc.cu, default) against the L1-gather
kernel (csrc/build_volume.cu) -- the same restatement of core/corr.py:46-97 + alt_cuda_corr.forward(radius 0), pinned
against the reference's goldens in test_gpu_corrblock.py -- on geometry the golden scene does not have: ragged image
sizes, out-of-image epipolar lines, projective poles (points behind the source camera, NaN coordinates), strong zoom
(the box of one hypothesis does not fit: sample-by-sample mode), per-view output, row bands, BASELINE sizes.
The two kernels sum the 64 channels in a different order: tolerance 2e-5 absolute on correlations of magnitude ~1."""
import numpy as np
import pytest
import torch

from cer_mvs_b200 import _lib, synth
from util import t

pytestmark = pytest.mark.gpu

STAGES = [(64, 0.0025 / 64, True), (44, 0.0025 / 320, False)]


def _build(variant, fm, poses, K, disp_in, stage, per_view=False, parts=None):
    """cer_build_volume(_rows) on NHWC fp16 features prepared by cer_nchw_to_nhwc; returns (volume, origin)."""
    L = _lib.lib()
    st = _lib.stream_ptr()
    _, n, C, h, w = fm.shape
    V = n - 1
    D, incre, shift = STAGES[stage]
    feats = torch.empty(n, h, w, C, device="cuda", dtype=torch.float16)
    _lib.check(L.cer_nchw_to_nhwc(fm.data_ptr(), 1, feats.data_ptr(), 1, n, C, h, w, 0.125, st))
    ii = torch.zeros(V, dtype=torch.int32, device="cuda")
    jj = torch.arange(1, V + 1, dtype=torch.int32, device="cuda")
    Pij = torch.empty(V, 16, device="cuda")
    _lib.check(L.cer_projection_matrices(poses.data_ptr(), K.data_ptr(), ii.data_ptr(), jj.data_ptr(), V, Pij.data_ptr(), st))
    origin = torch.zeros(h, w, device="cuda")
    vol = torch.full((V if per_view else 1, h * w, D), 7.0, device="cuda")
    lo = float(torch.tensor(D // 2 * incre).float())
    _lib.check(L.cer_set_build_variant(variant))
    try:
        if parts is None:
            _lib.check(L.cer_build_volume(feats.data_ptr(), 1, Pij.data_ptr(), ii.data_ptr(), jj.data_ptr(), V,
                                          disp_in.data_ptr(), int(shift), D, incre, lo, origin.data_ptr(),
                                          vol.data_ptr(), 1.0 if per_view else 1.0 / V, int(per_view), h, w, st))
        else:
            # parts: [(view_begin, view_end, d_begin, d_end)] accumulated into a zeroed volume (the sharded build)
            vol.zero_()
            for vb, ve, d0, d1 in parts:
                _lib.check(L.cer_build_volume_part(feats.data_ptr(), 1, Pij[vb:].data_ptr(), ii[vb:].data_ptr(),
                                                   jj[vb:].data_ptr(), ve - vb, disp_in.data_ptr(), int(shift), D, incre,
                                                   lo, origin.data_ptr(), vol.data_ptr(), 1.0 / V, 0, h, w, d0, d1, 1, st))
        torch.cuda.synchronize()
    finally:
        L.cer_set_build_variant(0)
    return vol, origin


def _scene(H, W, V, seed, stage):
    sc = synth.make_scene(H, W, V, seed=seed)
    fm = t(sc["fmaps"]).cuda().half()
    poses = t(sc["poses"])[0].cuda().contiguous()
    K = t(sc["intrinsics"])[0].clone()
    K[:, :2] /= 4
    K = K.cuda().contiguous()
    disp = torch.zeros(H // 4, W // 4, device="cuda") if stage == 0 else t(sc["true_disp"]).cuda().contiguous()
    return fm, poses, K, disp


def _compare(a, b, tol=2e-5):
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(b)
    err = float(np.abs(a[m] - b[m]).max()) if m.any() else 0.0
    assert err <= tol, err
    return err


@pytest.mark.parametrize("stage", [0, 1])
@pytest.mark.parametrize("hw,V", [((80, 112), 3), ((148, 212), 2), ((64, 68), 4), ((448, 576), 2)])
def test_staged_equals_gather(hw, V, stage):
    fm, poses, K, disp = _scene(hw[0], hw[1], V, seed=40 + stage, stage=stage)
    got, o1 = _build(0, fm, poses, K, disp, stage)
    want, o2 = _build(1, fm, poses, K, disp, stage)
    assert torch.equal(o1, o2)
    err = _compare(got, want)
    assert float(want.abs().mean()) > 1e-3
    print(f"{hw} V={V} stage {stage}: max |staged - gather| = {err:.2e}")


@pytest.mark.parametrize("stage", [0, 1])
def test_per_view_output(stage):
    fm, poses, K, disp = _scene(80, 112, 3, seed=44, stage=stage)
    got, _ = _build(0, fm, poses, K, disp, stage, per_view=True)
    want, _ = _build(1, fm, poses, K, disp, stage, per_view=True)
    _compare(got, want)
    mean, _ = _build(0, fm, poses, K, disp, stage)
    _compare(mean[0], want.mean(0), tol=3e-5)


@pytest.mark.parametrize("case", ["far_baseline", "behind_camera", "zoom", "rotated90", "tiny_source_focal", "nan_pose"])
@pytest.mark.parametrize("stage", [0, 1])
def test_degenerate_geometry(case, stage):
    """Epipolar lines that leave the image, cross a projective pole, or cover many source pixels per hypothesis."""
    fm, poses, K, disp = _scene(96, 128, 3, seed=50, stage=stage)
    poses, K = poses.clone(), K.clone()
    if case == "far_baseline":
        poses[1:, :3, 3] *= 8.0                 # samples run out of the source image, +-1e4 clamp region
    elif case == "behind_camera":
        poses[1, 2, 3] = -700.0                 # the plane (depth ~600) ends up behind source camera 1: X2 changes sign
        poses[2, 2, 3] = -590.0
    elif case == "zoom":
        K[1, :2] *= 6.0                         # one hypothesis of a 16x8 tile covers > 64 source columns: DIRECT mode
        K[2, :2] *= 2.5
    elif case == "rotated90":
        R = torch.tensor([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]], device="cuda")
        poses[1, :3, :3] = R @ poses[1, :3, :3]
        poses[1, :3, 3] = R @ poses[1, :3, 3]
    elif case == "tiny_source_focal":
        K[2, :2] *= 0.05                        # the whole source image is a few pixels wide
    elif case == "nan_pose":
        poses[3, 0, 3] = float("nan")
    got, _ = _build(0, fm, poses, K, disp, stage, per_view=True)
    want, _ = _build(1, fm, poses, K, disp, stage, per_view=True)
    err = _compare(got, want)
    print(f"{case} stage {stage}: max |staged - gather| = {err:.2e}, NaN fraction {float(torch.isnan(want).float().mean()):.3f}, "
          f"nonzero fraction {float((want != 0).float().mean()):.3f}")


@pytest.mark.parametrize("noise_steps", [0.5, 3.0, 30.0, 300.0])
def test_noisy_disparity_input(noise_steps):
    """Second cascade stage on a disparity map that is noisy from pixel to pixel (what an untrained GRU produces): the
    samples of neighbouring pixels are far apart, tiles whose hypothesis origins differ by more than a few steps are
    flagged by the staged kernel and computed by the gather kernel -- the result is the same either way."""
    fm, poses, K, disp = _scene(160, 224, 3, seed=48, stage=1)
    g = torch.Generator(device="cuda").manual_seed(1)
    incre = STAGES[1][1]
    noisy = disp + (torch.rand(disp.shape, device="cuda", generator=g) - 0.5) * 2 * noise_steps * incre
    noisy[:8, :] = disp[:8, :]                     # a coherent band stays on the staged path
    got, _ = _build(0, fm, poses, K, noisy.contiguous(), 1)
    want, _ = _build(1, fm, poses, K, noisy.contiguous(), 1)
    err = _compare(got, want)
    frac_equal = float((got == want).float().mean())
    print(f"noise {noise_steps} steps: max |staged - gather| = {err:.2e}, bit-equal fraction {frac_equal:.3f}")
    if noise_steps >= 30:
        assert frac_equal > 0.8                    # most tiles were handed to the gather kernel


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("stage", [0, 1])
def test_unit_parts_add_up_to_whole(variant, stage):
    """(view, hypothesis) parts accumulated into a zeroed volume == the whole build (up to the order of the view sum)."""
    fm, poses, K, disp = _scene(160, 112, 3, seed=46, stage=stage)
    D = STAGES[stage][0]
    whole, _ = _build(variant, fm, poses, K, disp, stage)
    parts = [(0, 1, 0, D), (1, 2, 0, 17), (1, 2, 17, D), (2, 3, 0, 5), (2, 3, 5, 6), (2, 3, 6, D)]
    got, _ = _build(variant, fm, poses, K, disp, stage, parts=parts)
    _compare(got, whole, tol=1e-6)
    two, _ = _build(variant, fm, poses, K, disp, stage, parts=[(0, 2, 0, D), (2, 3, 0, D)])
    _compare(two, whole, tol=1e-6)


@pytest.mark.parametrize("cfg", ["cfg2_dtu_1184x1600_v10", "cfg5_blended_1536x2048_v7"])
def test_baseline_sizes(cfg):
    H, W, V = synth.CONFIGS[cfg]
    for stage in (0, 1):
        fm, poses, K, disp = _scene(H, W, V, seed=47, stage=stage)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = {}
        outs = {}
        for variant in (0, 1):
            _build(variant, fm, poses, K, disp, stage)
            e0.record()
            outs[variant], _ = _build(variant, fm, poses, K, disp, stage)
            e1.record()
            torch.cuda.synchronize()
            ms[variant] = e0.elapsed_time(e1)
        err = _compare(outs[0], outs[1])
        print(f"{cfg} stage {stage}: staged {ms[0]:.2f} ms, gather {ms[1]:.2f} ms (incl. layout + projection), "
              f"max diff {err:.2e}")
