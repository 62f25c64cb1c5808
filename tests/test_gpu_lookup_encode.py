"""GPU: the fused pyramid lookup + first corr-encoder layer (cer_lookup_encode; core/corr.py:102-143 + core/update.py:62
under autocast) -- the warp-autonomous kernel for the reference's cascade widths (D = 64 / 44) and the general kernel,
against the oracle's lookup + autocast conv and against each other.

The general kernel recomputes the reference's normalise / unnormalise round trip per tap (bit-exact tap positions); the
warp-autonomous kernel does it once per level and shares floor and weights between the eleven taps, which moves a tap
position by <= 1 ulp of the coordinate (4e-6): both are then rounded to fp16 (autocast), so they differ by rare one-ulp
fp16 flips.  Tolerances below: <= 2 fp16 ulps at the value scale (4e-3 at |e1| ~ 2), mean 1e-4."""
import numpy as np
import pytest
import torch

import cer_oracle as O
from cer_mvs_b200 import _lib, synth
from cer_mvs_b200.update import pack_update_weights

pytestmark = pytest.mark.gpu


def _case(D, h, w, seed, mode):
    rs = np.random.RandomState(seed)
    px = h * w
    volume = (rs.standard_normal((px, D)) * 2.0).astype(np.float32)
    origin = rs.uniform(0.2, 0.4, px).astype(np.float32)
    incre = np.float32(1.0 / 64)
    if mode == "wide":            # coordinates from far below 0 (clamped by corr.py:107) to far past the volume
        c = rs.uniform(-12.0, D + 12.0, px)
    elif mode == "smooth":        # neighbouring pixels look up neighbouring columns (what a trained model does)
        yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        c = (D / 2 + 0.37 * D * np.sin(xx / 9.0) * np.cos(yy / 7.0)).reshape(-1) + rs.uniform(-0.5, 0.5, px)
    else:                         # exact integers and half-integers, both borders: floor boundaries of every level
        c = rs.randint(-2, 2 * D + 6, px) * 0.5
    disp = (origin + (c - D // 2) * incre).astype(np.float32)
    return volume, origin, disp, float(incre)


def _run(variant, blob, volume, origin, disp, D, incre, h, w):
    L = _lib.lib()
    e1 = torch.full((h * w, 64), float("nan"), device="cuda", dtype=torch.float16)
    d = disp.clone()
    try:
        _lib.check(L.cer_set_lookup_variant(variant))
        _lib.check(L.cer_lookup_encode(blob.data_ptr(), volume.data_ptr(), origin.data_ptr(), d.data_ptr(), D, incre, h, w,
                                       e1.data_ptr(), _lib.stream_ptr()), "cer_lookup_encode")
        torch.cuda.synchronize()
    finally:
        L.cer_set_lookup_variant(2)
    assert torch.equal(d, disp)          # no pending delta: the disparity map is read only
    return e1.float().cpu().numpy()


def _oracle(sd, volume, origin, disp, D, incre, h, w):
    corr = torch.from_numpy(volume).reshape(1, h, w, 1, D)
    pyr = [corr]
    for _ in range(2):
        W = corr.shape[-1]
        corr = 0.5 * (corr[..., 0:2 * (W // 2):2] + corr[..., 1:2 * (W // 2):2])
        pyr.append(corr)
    taps = O.lookup(pyr, torch.from_numpy(origin).reshape(1, h, w), D, incre, torch.from_numpy(disp).reshape(1, 1, h, w))
    e = torch.relu(O._conv(taps[:, 0], O.to_torch_sd(sd), "corr_encoder.0", True))     # [1,64,h,w]
    return e[0].permute(1, 2, 0).reshape(h * w, 64).numpy()


@pytest.mark.parametrize("mode", ["wide", "smooth", "borders"])
@pytest.mark.parametrize("D,grid", [(64, (37, 53)), (44, (37, 53)), (64, (16, 32)), (44, (5, 7)), (48, (21, 40))])
def test_lookup_encode(D, grid, mode):
    h, w = grid
    volume, origin, disp, incre = _case(D, h, w, 11 + D + h, mode)
    sd = synth.make_update_weights(seed=5)
    blob = torch.from_numpy(pack_update_weights(sd)).cuda()
    tv, to, td = (torch.from_numpy(a).cuda() for a in (volume, origin, disp))
    want = _oracle(sd, volume, origin, disp, D, incre, h, w)
    general = _run(1, blob, tv, to, td, D, incre, h, w)
    fast = _run(2, blob, tv, to, td, D, incre, h, w)          # D = 48 has no warp-autonomous kernel: same kernel twice
    assert np.isfinite(fast).all() and np.isfinite(general).all()
    for name, got in (("general", general), ("warp-autonomous", fast)):
        d = np.abs(got - want)
        print(f"D={D} {h}x{w} {mode} {name}: max {d.max():.2e} mean {d.mean():.2e} differing {np.mean(d > 0):.4f}")
        assert d.max() <= 4e-3 and d.mean() < 1e-4, name
    d = np.abs(fast - general)
    assert d.max() <= 4e-3 and np.mean(d > 0) < 0.03
    if D == 48:
        assert np.array_equal(fast, general)
