"""GPU: the WHOLE RAFT.forward (core/raft.py:34-108: normalisation, cnet, fnet per image, cascade stages) of
cer_mvs_b200.raft.RAFT against the unmodified reference model run on this GPU under torch.cuda.amp.autocast
(baseline/_ref + the reference's own alt_cuda_corr kernel): real encoders, images in, disparity out."""
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


def _state_dict(seed, dscale, dbias):
    sd = {}
    for k, v in synth.make_encoder_weights(seed=seed, out_dim=64).items():
        sd["fnet." + k] = t(v)
    for k, v in synth.make_encoder_weights(seed=seed + 1, out_dim=128).items():
        sd["cnet." + k] = t(v)
    for k, v in synth.make_update_weights(seed=seed, delta_scale=dscale, delta_bias=dbias).items():
        sd["update_block." + k] = t(v)
    return sd


def _inputs(H, W, V, seed):
    sc = synth.make_scene(H, W, V, seed=seed)
    # views of one textured plane are not available as images; a shared low-passed noise image shifted per view gives the
    # cost volume some structure (the parity statement does not depend on it)
    base = synth.make_image(H, W + 64, n=1, seed=seed)[0]
    images = np.stack([base[:, :, 4 * v:4 * v + W] for v in range(V + 1)])[None]
    return t(np.ascontiguousarray(images)).cuda(), t(sc["poses"]).cuda(), t(sc["intrinsics"]).cuda()


def test_state_dict_keys_match_reference_model():
    from cer_mvs_b200.raft import RAFT
    m = RAFT(cascade=[(64, 64, 8), (-1, 320, 8)], test_mode=True)
    want = set(_state_dict(0, 0.1, 0.0).keys())
    assert set(m.state_dict().keys()) == want
    m.load_state_dict(_state_dict(0, 0.1, 0.0), strict=True)


@pytest.mark.skipif(not refrun.available("gpu"), reason="baseline/_ref or oracle/_ref not present")
@pytest.mark.parametrize("iters,dbias", [(2, 0.02), (8, 0.01)])
def test_whole_forward_vs_reference_model_on_this_gpu(iters, dbias):
    from cer_mvs_b200.raft import RAFT
    ref = refrun.import_reference("gpu")
    refrun.restore_reference_classes()
    H, W, V = synth.CONFIGS["cfg1_dtu_448x576_v2"]
    cascade = [(64, 64, iters), (-1, 320, iters)]
    sd = _state_dict(61, 0.1, dbias)
    images, poses, K = _inputs(H, W, V, 61)
    ours = RAFT(cascade=cascade, test_mode=True)
    ours.load_state_dict(sd, strict=True)
    ours = ours.cuda().eval()
    model = ref.raft.RAFT(cascade=cascade, test_mode=True)
    model.load_state_dict(sd, strict=True)               # the same checkpoint, both ways
    model = model.cuda().eval()
    scale = torch.tensor([1.0], dtype=torch.float64, device="cuda")
    with torch.no_grad():
        want = model(images.clone(), poses.clone(), K.clone(), scale=scale)
        got = ours(images, poses, K, scale=scale)
        again = ours(images, poses, K, scale=scale)
    assert got.dtype == want.dtype == torch.float64 and got.shape == want.shape
    assert torch.equal(got, again)
    err = rel_l1(got.cpu().numpy(), want.cpu().numpy())
    print(f"whole RAFT.forward 448x576 V=2 {iters}+{iters}: rel L1 vs the reference model on this GPU = {err:.3e}; "
          f"mean disp {float(want.mean()):.3e}")
    assert err < 1e-3, err
