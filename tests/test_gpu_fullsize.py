"""GPU: size-independent properties at the BASELINE.json configurations' full sizes (no oracle run
needed): known answers that hold for any image size exercise every tile / edge path of the kernels."""
import numpy as np
import pytest
import torch

from cer_mvs_b200 import synth

pytestmark = pytest.mark.gpu

# (h1, w1, V) of BASELINE configs 2, 4, 5; config 3 (592 x 800) is covered by the build test
GRIDS = {"cfg2": (296, 400, 10), "cfg4": (264, 480, 15), "cfg5": (384, 512, 7)}


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_build_identity_view_known_answer(cfg, dtype):
    """Source == reference image and pose: every hypothesis reprojects onto its own pixel, so volume[p, d] = |f_p|^2 / 64."""
    from cer_mvs_b200.corr import CorrBlock
    h1, w1 = (296, 400) if cfg == "cfg2" else (592, 800)
    g = torch.Generator(device="cuda").manual_seed(0)
    f = torch.randn(1, 1, 64, h1, w1, generator=g, device="cuda").half()
    fm = torch.cat([f, f], 1).to(dtype)
    poses = torch.eye(4, device="cuda").repeat(1, 2, 1, 1)
    K = torch.tensor([[700.0, 0, w1 / 2], [0, 700.0, h1 / 2], [0, 0, 1]], device="cuda").repeat(1, 2, 1, 1)
    want = (f.float()[0, 0] ** 2).sum(0).reshape(-1, 1) / 64
    for D, incre, shift in [(64, 0.0025 / 64, True), (44, 0.0025 / 320, False)]:
        cb = CorrBlock(fm, poses, K, torch.zeros(1, dtype=torch.long, device="cuda"),
                       torch.ones(1, dtype=torch.long, device="cuda"), nIncre=D, incre=incre,
                       disps_input=torch.full((1, 1, h1, w1), 0.001, device="cuda"), shift=shift, num_levels=3,
                       radius=5, test_mode=True, do_report=False)
        torch.testing.assert_close(cb.volume[0], want.expand(-1, D), rtol=1e-5, atol=1e-5)


def test_build_is_linear_in_views():
    """mean over {a, b} == (volume{a} + volume{b}) / 2 at config-4 size (the identity view sharding relies on)."""
    from cer_mvs_b200.corr import CorrBlock
    h1, w1, _ = GRIDS["cfg4"]
    poses, K = synth.make_cameras(2, 4 * h1, 4 * w1, seed=3)
    Kq = torch.from_numpy(K).clone()
    Kq[:, :2] /= 4
    g = torch.Generator(device="cuda").manual_seed(1)
    fm = torch.randn(1, 3, 64, h1, w1, generator=g, device="cuda").half()
    P = torch.from_numpy(poses)[None].cuda()
    Kc = Kq[None].cuda()
    disp = torch.full((1, 1, h1, w1), 0.0015, device="cuda")

    def vol(jj):
        ii = torch.zeros(len(jj), dtype=torch.long, device="cuda")
        cb = CorrBlock(fm, P, Kc, ii, torch.tensor(jj, device="cuda"), nIncre=44, incre=0.0025 / 320, disps_input=disp,
                       shift=False, num_levels=3, radius=5, test_mode=True, do_report=False)
        return cb.volume[0]
    both, a, b = vol([1, 2]), vol([1]), vol([2])
    torch.testing.assert_close(both, (a + b) / 2, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("D", [64, 44])
def test_lookup_of_a_ramp_known_answer(D):
    """volume[p, d] = d: bilinear interpolation and floor-pooling of a ramp are exact, so tap j of level l at coordinate x
    reads x + 2^l j + (2^l - 1)/2 while in range (config-5 size)."""
    from cer_mvs_b200 import _lib
    h1, w1, _ = GRIDS["cfg5"]
    px = h1 * w1
    vol = torch.arange(D, dtype=torch.float32, device="cuda").repeat(px, 1).contiguous()
    origin = torch.full((px,), 0.001, device="cuda")
    incre = 0.0025 / 64
    xs = torch.linspace(8.25, D - 8.5, px, device="cuda")             # lookup coordinate per pixel
    zinv = (origin + (xs - D // 2) * incre).contiguous()
    out = torch.empty(33, px, device="cuda")
    _lib.check(_lib.lib().cer_lookup(vol.data_ptr(), 1, origin.data_ptr(), zinv.data_ptr(), D, incre, 5, 3,
                                     out.data_ptr(), h1, w1, _lib.stream_ptr()))
    x = torch.clamp((zinv - origin) / incre + D // 2, min=0)
    for lvl in range(3):
        for j in range(-5, 6):
            pos = x / 2 ** lvl + j
            inside = (pos >= 0) & (pos <= (D >> lvl) - 1)
            want = x + (2 ** lvl) * j + (2 ** lvl - 1) / 2
            got = out[lvl * 11 + j + 5]
            if inside.any():
                torch.testing.assert_close(got[inside], want[inside], rtol=2e-5, atol=2e-4)
            outside = (pos <= -1) | (pos >= (D >> lvl))             # both taps in the zero padding
            assert (got[outside] == 0).all()


@pytest.mark.parametrize("cfg", ["cfg2", "cfg4", "cfg5"])
def test_update_block_zero_weights_known_answer(cfg):
    """All conv weights zero, biases zero except the last delta bias b: z = r = 0.5, q = 0 -> net' = fp16(0.5 net),
    delta = fp16(0.01 * fp16(b)) at every pixel, image borders and partial tiles included."""
    from cer_mvs_b200.update import UpdateBlock
    h1, w1, V = GRIDS[cfg]
    ub = UpdateBlock(cascade=[(64, 64, 8), (-1, 320, 8)], dim_net=64, dim_inp=64)
    with torch.no_grad():
        for p in ub.parameters():
            p.zero_()
        ub.delta1[2].bias.fill_(0.625)
    ub = ub.cuda()
    g = torch.Generator(device="cuda").manual_seed(2)
    net = torch.tanh(torch.randn(1, 1, 64, h1, w1, generator=g, device="cuda")).half()
    inp = torch.relu(torch.randn(1, 1, 64, h1, w1, generator=g, device="cuda")).half()
    disp = torch.rand(1, 1, h1, w1, generator=g, device="cuda") * 0.002
    corr = torch.randn(1, 1, 33, h1, w1, generator=g, device="cuda")
    with torch.no_grad():
        n2, d2 = ub(net, inp, disp, corr, 1)
    assert torch.equal(n2, (net.float() * 0.5).half())
    want = torch.tensor(0.01 * 0.625).half().float()        # fp16(0.01 * fp16(0.625))
    assert torch.equal(d2, torch.full_like(d2, float(want)))


def test_hot_path_full_size_deterministic_and_graph_equals_eager():
    from cer_mvs_b200.hotpath import DepthHotPath
    h1, w1, V = GRIDS["cfg2"]
    poses, K = synth.make_cameras(V, 4 * h1, 4 * w1, seed=0)
    g = torch.Generator(device="cuda").manual_seed(3)
    fm = torch.randn(1, V + 1, 64, h1, w1, generator=g, device="cuda").half()
    net = torch.tanh(torch.randn(1, 1, 64, h1, w1, generator=g, device="cuda")).half()
    inp = torch.relu(torch.randn(1, 1, 64, h1, w1, generator=g, device="cuda")).half()
    sd = synth.make_update_weights(seed=0, delta_scale=0.1, delta_bias=0.005)
    outs = []
    for use_graph in (True, False):
        hp = DepthHotPath(h1, w1, max_views=V, cascade=[(64, 64, 2), (-1, 320, 2)], use_graph=use_graph)
        hp.load_update_block(sd)
        a = hp(fm, net, inp, torch.from_numpy(poses)[None].cuda(), torch.from_numpy(K)[None].cuda(), 1.0).clone()
        b = hp(fm, net, inp, torch.from_numpy(poses)[None].cuda(), torch.from_numpy(K)[None].cuda(), 1.0).clone()
        assert torch.equal(a, b)
        assert torch.isfinite(a).all()
        outs.append(a)
    assert torch.equal(outs[0], outs[1])
