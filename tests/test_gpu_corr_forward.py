"""GPU: the alt_cuda_corr.forward drop-in (csrc/corr_ops.cu) against (a) the CPU oracle restatement and
(b) the reference's own kernel compiled from its sources (oracle/_ref), which also pins the oracle."""
import numpy as np
import pytest
import torch

import cer_oracle as O
from util import ref_ext

pytestmark = pytest.mark.gpu


def _case(B, H1, W1, H2, W2, C, N, seed, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn(B, H1, W1, C, generator=g)
    f2 = torch.randn(B, H2, W2, C, generator=g)
    cx = torch.rand(B, N, H1, W1, generator=g) * (W2 + 6) * spread - 3
    cy = torch.rand(B, N, H1, W1, generator=g) * (H2 + 6) * spread - 3
    return f1, f2, torch.stack([cx, cy], -1).contiguous()


@pytest.mark.parametrize("r", [0, 1, 2])
@pytest.mark.parametrize("C", [64, 7, 96])
def test_vs_oracle_and_reference_kernel(r, C):
    import cer_mvs_b200.alt_cuda_corr as acc
    f1, f2, co = _case(2, 9, 13, 11, 10, C, 5, seed=10 * r + C)
    want, = O.corr_forward(f1, f2, co, r)
    got, = acc.forward(f1.cuda(), f2.cuda(), co.cuda(), r)
    assert got.shape == want.shape and got.dtype == torch.float32
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-5)
    ext = ref_ext()
    # the reference kernel reads channels in unguarded chunks of 32 (correlation_kernel.cu:43,52): C % 32 only
    if ext is not None and C % 32 == 0:
        ref, = ext.forward(f1.cuda(), f2.cuda(), co.cuda(), r)
        torch.cuda.synchronize()
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(want, ref.cpu(), rtol=1e-5, atol=1e-5)   # pins the oracle itself


def test_edge_coordinates():
    """Integer coords, far out-of-bounds (+-1e4 clamp values, core/corr.py:88), NaN."""
    import cer_mvs_b200.alt_cuda_corr as acc
    f1, f2, co = _case(1, 4, 8, 6, 7, 64, 6, seed=3)
    co[0, 0] = torch.floor(co[0, 0])
    co[0, 1] = 1e4
    co[0, 2] = -1e4
    co[0, 3, 0, 0, 0] = float("nan")
    want, = O.corr_forward(f1, f2, co, 0)
    got, = acc.forward(f1.cuda(), f2.cuda(), co.cuda(), 0)
    got = got.cpu()
    assert torch.isnan(got[0, 3, 0, 0, 0]) and torch.isnan(want[0, 3, 0, 0, 0])
    assert (got[0, 1] == 0).all() and (got[0, 2] == 0).all()
    m = ~torch.isnan(want)
    torch.testing.assert_close(got[m], want[m], rtol=1e-5, atol=1e-5)
    ext = ref_ext()
    if ext is not None:
        ref, = ext.forward(f1.cuda(), f2.cuda(), co.cuda(), 0)
        ref = ref.cpu()
        assert torch.equal(torch.isnan(ref), torch.isnan(got))
        torch.testing.assert_close(got[m], ref[m], rtol=1e-5, atol=1e-5)


def test_empty_and_errors():
    import cer_mvs_b200.alt_cuda_corr as acc
    f1, f2, co = _case(1, 4, 8, 6, 7, 64, 0, seed=1)
    got, = acc.forward(f1.cuda(), f2.cuda(), co.cuda(), 0)
    assert got.shape == (1, 0, 1, 4, 8)
    f1, f2, co = _case(1, 4, 8, 6, 7, 64, 2, seed=1)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        acc.forward(f1, f2.cuda(), co.cuda(), 0)
    with pytest.raises(RuntimeError, match="must be contiguous"):
        acc.forward(f1.cuda().permute(0, 2, 1, 3), f2.cuda(), co.cuda(), 0)
    with pytest.raises(NotImplementedError):
        acc.backward(f1.cuda(), f2.cuda(), co.cuda(), got, 0)


def test_cer_mvs_shape_cfg1():
    """The shape CER-MVS itself calls with at BASELINE configs[0]: B=1, 112x144, C=64, N=64, r=0."""
    import cer_mvs_b200.alt_cuda_corr as acc
    ext = ref_ext()
    if ext is None:
        pytest.skip("oracle/_ref not built")
    f1, f2, co = _case(1, 112, 144, 112, 144, 64, 64, seed=5)
    got, = acc.forward(f1.cuda(), f2.cuda(), co.cuda(), 0)
    ref, = ext.forward(f1.cuda(), f2.cuda(), co.cuda(), 0)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=2e-5)
