"""GPU: CorrBlock drop-in (fused build + lookup kernels) against golden outputs of the reference's
core/corr.py (tests/golden/ops_corrblock.npz) and the oracle."""
import numpy as np
import pytest
import torch

import cer_oracle as O
from cer_mvs_b200 import synth
from util import H, V, W, h1, t, w1

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["staged_tcgen05", "fhfma"], autouse=True)
def build_variant(request):
    """fp16-feature builds run on both kernels: TMA-staged boxes + tcgen05 (default) and the FHFMA L1 gather."""
    from cer_mvs_b200 import _lib
    _lib.check(_lib.lib().cer_set_build_variant({"staged_tcgen05": 0, "fhfma": 1}[request.param]))
    yield request.param
    _lib.lib().cer_set_build_variant(0)
STAGES = [(64, 0.0025 / 64, True), (44, 0.0025 / 320, False)]


def _block(golden, stage, per_view, dtype=torch.float32, seed=1):
    from cer_mvs_b200.corr import CorrBlock
    g = golden("ops_corrblock")
    sc = synth.make_scene(H, W, V, seed=seed)
    K = t(sc["intrinsics"]).clone()
    K[:, :, :2] /= 4
    D, incre, shift = STAGES[stage]
    ii = torch.zeros(V, dtype=torch.long).cuda()
    jj = torch.arange(1, V + 1).cuda()
    cb = CorrBlock(t(sc["fmaps"]).cuda().to(dtype), t(sc["poses"]).cuda(), K.cuda(), ii, jj, nIncre=D, incre=incre,
                   disps_input=t(g[f"s{stage}_disp_in"]).cuda(), shift=shift, num_levels=3, radius=5, test_mode=True,
                   do_report=False, per_view=per_view)
    return cb, g


@pytest.mark.parametrize("stage", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_volume_and_pyramid_per_view(golden, stage, dtype):
    """fp16 features are lossless here: the synthetic fmaps are fp16-representable (like autocast fnet output)."""
    cb, g = _block(golden, stage, per_view=True, dtype=dtype)
    assert np.array_equal(cb.disps_origin.cpu().numpy(), g[f"s{stage}_origin"])
    pyr = cb.corr_pyramid
    for l in range(3):
        want = g[f"s{stage}_pyr{l}"]
        got = pyr[l].reshape(want.shape).cpu().numpy()
        # Pij is built in fp64 here and in fp32 by the reference: sample positions differ by ~1e-4 px
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("stage", [0, 1])
def test_volume_view_mean(golden, stage):
    cb, g = _block(golden, stage, per_view=False)
    want = g[f"s{stage}_pyr0"].reshape(V, h1 * w1, -1).mean(0)
    np.testing.assert_allclose(cb.volume[0].cpu().numpy(), want, rtol=2e-4, atol=2e-4)


def test_build_with_reference_matrices_is_tight(golden):
    """Feed the oracle's fp32 Pij to the kernel: everything else must agree to fp32 rounding."""
    from cer_mvs_b200 import _lib
    g = golden("ops_corrblock")
    sc = synth.make_scene(H, W, V, seed=1)
    K = t(sc["intrinsics"]).clone()
    K[:, :, :2] /= 4
    Pij = O.projection_matrices(t(sc["poses"]), K, [0] * V, [1, 2, 3])[0].reshape(V, 16).contiguous().cuda()
    feats = torch.empty(V + 1, h1, w1, 64, device="cuda")
    fm = t(sc["fmaps"]).cuda()
    L = _lib.lib()
    st = _lib.stream_ptr()
    _lib.check(L.cer_nchw_to_nhwc(fm.data_ptr(), 0, feats.data_ptr(), 0, V + 1, 64, h1, w1, 0.125, st))
    ii = torch.zeros(V, dtype=torch.int32, device="cuda")
    jj = torch.arange(1, V + 1, dtype=torch.int32, device="cuda")
    for stage, (D, incre, shift) in enumerate(STAGES):
        disp = t(g[f"s{stage}_disp_in"]).cuda().reshape(h1, w1).contiguous()
        origin = torch.empty(h1, w1, device="cuda")
        vol = torch.empty(V, h1 * w1, D, device="cuda")
        lo = float(torch.tensor(D // 2 * incre).float())
        _lib.check(L.cer_build_volume(feats.data_ptr(), 0, Pij.data_ptr(), ii.data_ptr(), jj.data_ptr(), V,
                                      disp.data_ptr(), int(shift), D, incre, lo, origin.data_ptr(), vol.data_ptr(),
                                      1.0, 1, h1, w1, st))
        want = g[f"s{stage}_pyr0"].reshape(V, h1 * w1, D)
        np.testing.assert_allclose(vol.cpu().numpy(), want, rtol=1e-5, atol=2e-5)


def test_projection_matrices(golden):
    from cer_mvs_b200 import _lib
    sc = synth.make_scene(H, W, V, seed=1)
    K = t(sc["intrinsics"]).clone()
    K[:, :, :2] /= 4
    want = O.projection_matrices(t(sc["poses"]), K, [0] * V, [1, 2, 3])[0].reshape(V, 16)
    P = t(sc["poses"])[0].cuda().contiguous()
    Kc = K[0].cuda().contiguous()
    ii = torch.zeros(V, dtype=torch.int32, device="cuda")
    jj = torch.arange(1, V + 1, dtype=torch.int32, device="cuda")
    out = torch.empty(V, 16, device="cuda")
    _lib.check(_lib.lib().cer_projection_matrices(P.data_ptr(), Kc.data_ptr(), ii.data_ptr(), jj.data_ptr(), V,
                                                  out.data_ptr(), _lib.stream_ptr()))
    torch.testing.assert_close(out.cpu(), want, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("stage", [0, 1])
@pytest.mark.parametrize("per_view", [True, False])
def test_lookup(golden, stage, per_view):
    """Lookup on the golden volume itself (uploaded), so only the lookup kernel is under test."""
    cb, g = _block(golden, stage, per_view=per_view)
    vol = t(g[f"s{stage}_pyr0"]).reshape(V, h1 * w1, -1)
    cb.volume = (vol if per_view else vol.mean(0, keepdim=True)).contiguous().cuda()
    for name in ("true", "zero", "far", "rand"):
        z = t(g[f"s{stage}_z_{name}"]).cuda()
        out = cb(z[:, [0] * V]).cpu().numpy()
        want = g[f"s{stage}_lookup_{name}"]
        if not per_view:
            assert out.shape == (1, 1, 33, h1, w1)
            want = want.mean(1, keepdims=True)
        np.testing.assert_allclose(out, want, rtol=1e-5, atol=3e-6)


@pytest.mark.parametrize("stage", [0, 1])
def test_lookup_variants_bit_identical(golden, stage, build_variant):
    """Warp-autonomous lookup kernel (reference configuration, default) == general kernel, bit for bit, on a ragged
    pixel count (560 = 17.5 warps) and on lookups that leave the volume on both sides."""
    if build_variant != "fhfma":
        pytest.skip("independent of the build kernel")
    from cer_mvs_b200 import _lib
    cb, g = _block(golden, stage, per_view=True)
    outs = {}
    try:
        for v in (1, 2):
            _lib.check(_lib.lib().cer_set_lookup_variant(v))
            outs[v] = [cb(t(g[f"s{stage}_z_{name}"]).cuda()[:, [0] * V]).cpu().numpy() for name in ("true", "zero", "far", "rand")]
    finally:
        _lib.lib().cer_set_lookup_variant(2)
    for a, b in zip(outs[1], outs[2]):
        assert np.array_equal(a, b)


def test_identity_view_known_answer():
    """Source view == reference view with the same pose: every hypothesis reprojects onto the pixel
    itself, so the volume is |f|^2/64 for every d (size-independent known answer)."""
    from cer_mvs_b200.corr import CorrBlock
    hh, ww = 37, 53
    g = torch.Generator().manual_seed(0)
    f = torch.randn(1, 1, 64, hh, ww, generator=g).half().float()
    fm = torch.cat([f, f], 1).cuda()
    poses = torch.eye(4).repeat(1, 2, 1, 1).cuda()
    K = torch.tensor([[50.0, 0, 26], [0, 50, 18], [0, 0, 1]]).repeat(1, 2, 1, 1).cuda()
    cb = CorrBlock(fm, poses, K, torch.zeros(1, dtype=torch.long).cuda(), torch.ones(1, dtype=torch.long).cuda(),
                   nIncre=64, incre=0.0025 / 64, disps_input=torch.zeros(1, 1, hh, ww).cuda(), shift=True,
                   num_levels=3, radius=5, test_mode=True, do_report=False)
    want = (f[0, 0] ** 2).sum(0).reshape(-1, 1) / 64
    got = cb.volume[0].cpu()
    torch.testing.assert_close(got, want.expand_as(got), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("hw", [(20, 28), (37, 53), (296, 400)])
def test_layout_kernels_match_permute(hw, build_variant):
    """NCHW -> NHWC (x 1/8) in fp16: the 64-channel fast path (px % 8 == 0) and the general kernel against torch."""
    if build_variant != "fhfma":
        pytest.skip("independent of the build kernel")
    from cer_mvs_b200 import _lib
    h, w = hw
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(3, 64, h, w, generator=g, device="cuda").half()
    out = torch.empty(3, h, w, 64, device="cuda", dtype=torch.float16)
    _lib.check(_lib.lib().cer_nchw_to_nhwc(x.data_ptr(), 1, out.data_ptr(), 1, 3, 64, h, w, 0.125, _lib.stream_ptr()))
    want = (x.float() * 0.125).half().permute(0, 2, 3, 1).contiguous()
    assert torch.equal(out, want)
