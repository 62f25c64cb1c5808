"""GPU: the BasicEncoder drop-in (csrc/encoder.cu) against (a) the golden output of the reference's own
module (fp32 CPU run, oracle/gen_golden_encoder.py), (b) the oracle's autocast emulation, (c) the reference module run on
this GPU under torch.cuda.amp.autocast (baseline/_ref, when it travelled).  The kernels compute like autocast: fp16
operands / activations, fp32 accumulation; the bar against an fp32 run is therefore fp16-level (relative L1 < 2e-2 after
ten layers and nine instance norms), against autocast runs a few fp16 ulps on the output scale."""
import os
import sys

import numpy as np
import pytest
import torch

import cer_oracle as O
from cer_mvs_b200 import synth
from util import ROOT, rel_l1, t

sys.path.insert(0, os.path.join(ROOT, "baseline"))
import refrun  # noqa: E402

pytestmark = pytest.mark.gpu
H, W = 72, 104


def _enc(kind, seed):
    from cer_mvs_b200.extractor import BasicEncoder
    dim, norm = (64, "instance") if kind == "fnet" else (128, "none")
    enc = BasicEncoder(output_dim=dim, norm_fn=norm, type="HR")
    sd = synth.make_encoder_weights(seed=seed, out_dim=dim)
    enc.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
    return enc.cuda().eval(), sd


@pytest.mark.parametrize("kind", ["fnet", "cnet"])
def test_vs_reference_golden_and_autocast_oracle(golden, kind):
    g = golden("ops_encoder")
    img = synth.make_image(H, W, n=1, seed=int(g["image_seed"]))
    x = t(img) * (2 / 255.) - 1
    enc, sd = _enc(kind, int(g[f"{kind}_seed"]))
    with torch.no_grad():
        got = enc(x.cuda()).float().cpu().numpy()
    want32 = g[f"{kind}_out"]
    assert got.shape == want32.shape
    auto = O.basic_encoder(O.to_torch_sd(sd), x, kind == "fnet", autocast=True).numpy()
    e32, eauto, gap = rel_l1(got, want32), rel_l1(got, auto), rel_l1(auto, want32)
    print(f"{kind}: rel L1 vs reference fp32 = {e32:.3e} (autocast-vs-fp32 gap of the reference numerics {gap:.3e}); "
          f"vs autocast oracle = {eauto:.3e}")
    assert e32 < 2e-2 and eauto < 1e-2


def test_layouts_and_context_split(golden):
    g = golden("ops_encoder")
    img = t(synth.make_image(H, W, n=2, seed=3)).cuda()
    fnet, _ = _enc("fnet", 100)
    cnet, _ = _enc("cnet", 101)
    x = img * (2 / 255.) - 1
    with torch.no_grad():
        f = fnet(x)                                             # [2,64,h,w]
        fh = fnet.forward_features(img, scale=0.125, normalize=True)          # raw images in, build layout out
        raw = cnet(x[:1])
        net, inp = cnet.forward_context(x[:1])
        net_c, inp_c = cnet.forward_context(x[:1], nchw=True)
    assert f.dtype == torch.float16 and f.shape == (2, 64, H // 4, W // 4)
    assert torch.equal(fh, (f.float() * 0.125).half().permute(0, 2, 3, 1).contiguous())
    want_net = torch.tanh(raw[:, :64].float()).half()
    want_inp = torch.relu(raw[:, 64:])
    assert torch.equal(inp.permute(0, 3, 1, 2), want_inp) and torch.equal(inp_c, want_inp)
    torch.testing.assert_close(net.permute(0, 3, 1, 2).float(), want_net.float(), rtol=0, atol=1e-3)   # tanhf vs torch.tanh
    assert torch.equal(net.permute(0, 3, 1, 2), net_c)


@pytest.mark.parametrize("hw", [(64, 96), (448, 576), (1184, 1600)])
def test_sizes_and_determinism(hw):
    fnet, _ = _enc("fnet", 7)
    img = t(synth.make_image(hw[0], hw[1], n=1, seed=5)).cuda()
    with torch.no_grad():
        a = fnet.forward_features(img, normalize=True)
        b = fnet.forward_features(img, normalize=True)
    assert torch.isfinite(a.float()).all() and torch.equal(a, b)
    # instance norm: per-channel statistics of the last normalised tensor are not observable here, but the output must
    # not depend on a constant brightness offset much less than on the image itself
    assert float(a.float().abs().mean()) > 1e-3


@pytest.mark.skipif(not refrun.available("gpu"), reason="baseline/_ref not present")
@pytest.mark.parametrize("kind", ["fnet", "cnet"])
def test_vs_reference_module_on_this_gpu(kind):
    sys.path.insert(0, refrun.ref_dir())
    refrun.import_reference("gpu")
    from core.extractor import BasicEncoder as RefEncoder
    dim, norm = (64, "instance") if kind == "fnet" else (128, "none")
    enc, sd = _enc(kind, 55)
    ref = RefEncoder(output_dim=dim, norm_fn=norm, type="HR")
    ref.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
    ref = ref.cuda().eval()
    x = t(synth.make_image(448, 576, n=1, seed=9)).cuda() * (2 / 255.) - 1
    with torch.no_grad():
        with torch.autocast("cuda", dtype=torch.float16):
            want = ref(x)
        got = enc(x)
    assert want.dtype == torch.float16
    err = rel_l1(got.float().cpu().numpy(), want.float().cpu().numpy())
    print(f"{kind} 448x576: rel L1 vs the reference module on this GPU (autocast) = {err:.3e}")
    assert err < 1e-2
