"""GPU: UpdateBlock / ConvGRU drop-ins (tensor-core conv kernels with fused GRU epilogues) against
the oracle's autocast emulation (the reference's GPU numerics) and the reference's fp32 golden."""
import numpy as np
import pytest
import torch

import cer_oracle as O
from cer_mvs_b200 import synth
from util import V, h1, rel_l1, t, w1

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["default", "hmma"], autouse=True)
def conv_variant(request):
    """Every test runs on both tensor-core paths: tcgen05.mma + TMEM (default) and mma.sync (v1)."""
    from cer_mvs_b200 import _lib
    _lib.check(_lib.lib().cer_set_conv_variant({"default": 1, "hmma": 0}[request.param]))
    yield request.param
    _lib.lib().cer_set_conv_variant(1)


def _ub(sd_np):
    from cer_mvs_b200.update import UpdateBlock
    ub = UpdateBlock(cascade=[(64, 64, 8), (-1, 320, 8)], dim_net=64, dim_inp=64)
    ub.load_state_dict({k: t(v) for k, v in sd_np.items()}, strict=True)
    return ub.cuda().eval()


@pytest.mark.parametrize("stage", [0, 1])
def test_update_block_vs_autocast_oracle(golden, stage):
    g = golden("ops_update")
    sd_np = synth.make_update_weights(seed=2, delta_scale=1.0)
    ub = _ub(sd_np)
    net, inp, disp, corr = (t(g[k]) for k in ("net", "inp", "disp", "corr_frames"))
    with torch.no_grad():
        n2, d2 = ub(net.cuda().half(), inp.cuda().half(), disp.cuda(), corr.cuda(), stage)
    assert n2.shape == (1, 1, 64, h1, w1) and n2.dtype == torch.float16 and d2.shape == (1, 1, h1, w1)
    wn, wd = O.update_block(O.to_torch_sd(sd_np), net, inp, disp, corr, stage, autocast=True)
    # identical rounding points; only fp32 summation order differs -> rare 1-ulp fp16 flips
    dn = (n2.float().cpu() - wn).abs()
    assert float(dn.max()) < 4e-3 and float((dn > 1e-3).float().mean()) < 2e-3
    assert rel_l1(d2.cpu().numpy(), wd.numpy()) < 2e-3
    # and against the reference's own fp32 result (golden): fp16-operand error level
    assert rel_l1(n2.float().cpu().numpy(), g[f"net_out{stage}"]) < 2e-3
    assert rel_l1(d2.cpu().numpy(), g[f"delta{stage}"]) < 1e-2


def test_update_block_state_reuse_and_fp32_io(golden):
    """Feeding back the returned net (core/raft.py:100) must equal feeding an equal copy; fp32 in -> fp32 out."""
    g = golden("ops_update")
    sd_np = synth.make_update_weights(seed=2, delta_scale=1.0)
    ub = _ub(sd_np)
    net, inp, disp, corr = (t(g[k]).cuda() for k in ("net", "inp", "disp", "corr_frames"))
    with torch.no_grad():
        n1, d1 = ub(net, inp, disp, corr, 0)
        assert n1.dtype == torch.float32
        n2a, d2a = ub(n1, inp, disp + d1, corr, 1)
        n2b, d2b = ub(n1.clone(), inp.clone(), disp + d1, corr, 1)
    assert torch.equal(n2a, n2b) and torch.equal(d2a, d2b)


def test_conv_gru_standalone(golden):
    from cer_mvs_b200.update import ConvGRU
    g = golden("ops_update")
    sd_np = synth.make_update_weights(seed=2, delta_scale=1.0)
    gru = ConvGRU(h_planes=64, i_planes=64 + 64 + 49)
    gru.load_state_dict({k[4:]: t(v) for k, v in sd_np.items() if k.startswith("gru.")}, strict=True)
    gru = gru.cuda()
    rs = np.random.RandomState(5)
    net, inp = t(g["net"])[0], t(g["inp"])[0]
    dn = t(rs.standard_normal((1, 49, h1, w1)).astype(np.float32) * 0.1)
    e = t(np.maximum(rs.standard_normal((1, 64, h1, w1)), 0).astype(np.float32)).half().float()
    with torch.no_grad():
        out = gru(net.cuda().half(), inp.cuda().half(), dn.cuda(), e.cuda().half())
    sd = O.to_torch_sd(sd_np)
    x = torch.cat([net, inp, dn, e], 1)
    z = O._h(torch.sigmoid(O._conv(x, sd, "gru.convz", True)))
    r = O._h(torch.sigmoid(O._conv(x, sd, "gru.convr", True)))
    q = O._h(torch.tanh(O._conv(torch.cat([O._h(r * net), inp, dn, e], 1), sd, "gru.convq", True)))
    want = O._h(O._h(O._h(1 - z) * net) + O._h(z * q))
    d = (out.float().cpu() - want).abs()
    assert float(d.max()) < 4e-3 and float((d > 1e-3).float().mean()) < 2e-3


def test_unsupported_architecture_fails_loudly():
    from cer_mvs_b200.update import UpdateBlock
    with pytest.raises(NotImplementedError):
        UpdateBlock(cascade=[(64, 64, 8), (-1, 320, 8)], dim_net=64, dim_inp=64, aggregation=["mean", "max"])
    ub = _ub(synth.make_update_weights(seed=2))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ub(torch.zeros(1, 1, 64, 8, 8), torch.zeros(1, 1, 64, 8, 8), torch.zeros(1, 1, 8, 8),
           torch.zeros(1, 3, 33, 8, 8), 0)


@pytest.mark.parametrize("hw", [(8, 8), (16, 8), (17, 9), (5, 40), (33, 130)])
@pytest.mark.parametrize("stage", [0, 1])
def test_update_block_ragged_and_single_tile_grids(hw, stage):
    """Grids of one 16 x 8 tile, of partial tiles on both borders and of odd tile counts (the delta conv runs as couples of
    CTAs over the same tiles, the gate conv as CTA pairs): UpdateBlock against the autocast oracle on seeded inputs."""
    h, w = hw
    rs = np.random.RandomState(100 + h * w + stage)
    sd_np = synth.make_update_weights(seed=2, delta_scale=1.0)
    ub = _ub(sd_np)
    net = t(np.tanh(rs.standard_normal((1, 1, 64, h, w))).astype(np.float16).astype(np.float32))
    inp = t(np.maximum(rs.standard_normal((1, 1, 64, h, w)), 0).astype(np.float16).astype(np.float32))
    disp = t((rs.uniform(0.2, 0.4, (1, 1, h, w))).astype(np.float32))
    corr = t((rs.standard_normal((1, 2, 33, h, w)) * 0.5).astype(np.float32))
    with torch.no_grad():
        n2, d2 = ub(net.cuda().half(), inp.cuda().half(), disp.cuda(), corr.cuda(), stage)
    wn, wd = O.update_block(O.to_torch_sd(sd_np), net, inp, disp, corr, stage, autocast=True)
    dn = (n2.float().cpu() - wn).abs()
    assert float(dn.max()) < 4e-3 and float((dn > 1e-3).float().mean()) < 2e-3
    assert rel_l1(d2.cpu().numpy(), wd.numpy()) < 2e-3
