"""GPU: image preparation, depth output and multi-resolution merge kernels (csrc/io_ops.cu)
against golden outputs of the reference's own functions (tests/golden/ops_io.npz) and the numpy oracle."""
import numpy as np
import pytest
import torch

import io_oracle as IO
from cer_mvs_b200 import prep
from util import t

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,s", [("s2", 2), ("s15", 1.5)])
def test_scale_operation(golden, name, s):
    g = golden("ops_io")
    K = t(g["scale_K_in"]).clone()
    out, K2 = prep.scale_operation(t(g["scale_in"]).cuda(), K, s)
    np.testing.assert_allclose(out.cpu().numpy(), g[f"scale_{name}_out"], rtol=2e-6, atol=2e-5)   # fp32, values 0..255
    np.testing.assert_array_equal(K2.numpy(), g[f"scale_{name}_K"])


def test_crop_operation_matches_reference_semantics():
    x = torch.arange(2 * 3 * 10 * 12, dtype=torch.float32).reshape(2, 3, 10, 12).cuda()
    K = torch.tensor([[[7.0, 0, 6], [0, 7, 5], [0, 0, 1]]] * 2)
    out, K2 = prep.crop_operation(x, K, 6, 8)
    assert out.shape == (2, 3, 6, 8) and torch.equal(out, x[:, :, 2:8, 2:10])
    assert K2[0, 0, 2] == 4 and K2[0, 1, 2] == 3


def test_normalize_images_bit_exact(golden):
    g = golden("ops_io")
    out = prep.normalize_images(t(g["scale_in"]).cuda())
    np.testing.assert_array_equal(out.cpu().numpy(), g["norm_out"])
    x = torch.rand(3, 1001, device="cuda") * 255            # ragged tail (n % 4 != 0)
    np.testing.assert_array_equal(prep.normalize_images(x).cpu().numpy(), IO.normalize_images(x.cpu().numpy()))


def test_disp_to_depth_and_pfm_file(golden, tmp_path):
    g = golden("ops_io")
    disp = t(g["disp"]).cuda()
    np.testing.assert_array_equal(prep.disp_to_depth(disp).cpu().numpy(), g["depth"])
    flipped = prep.disp_to_depth(disp, flip_rows=True)
    np.testing.assert_array_equal(flipped.cpu().numpy(), np.flipud(g["depth"]))
    p = tmp_path / "d.pfm"
    prep.write_pfm(p, flipped, flipped=True)
    assert open(p, "rb").read() == g["pfm_bytes"].tobytes()           # byte-identical to the reference's file
    np.testing.assert_array_equal(prep.readPFM(p), g["depth"])


def test_multires_merge(golden):
    g = golden("ops_io")
    out = prep.multires_merge(t(g["multires_im1"]).cuda(), t(g["multires_im2"]).cuda(), 0.02).cpu().numpy()
    want, im1r = IO.multires_merge(g["multires_im1"], g["multires_im2"], 0.02)
    margin = np.abs(np.abs(im1r - g["multires_im2"]) - np.float32(0.02) * im1r) > 1e-3
    np.testing.assert_allclose(out[margin], g["multires_out"][margin], rtol=1e-6)
    np.testing.assert_allclose(out[margin], want[margin], rtol=1e-6)


def test_full_size_round_trips():
    """BASELINE cfg 3 sizes: identity rescale is exact, a 2x rescale keeps the corner pixels, depth of depth is disp."""
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.rand(1, 3, 1184, 1600, device="cuda", generator=g) * 255
    K = torch.eye(3)[None].clone()
    same, _ = prep.scale_operation(img, K.clone(), 1)
    assert torch.equal(same, img)
    big, K2 = prep.scale_operation(img, K.clone(), 2)
    assert big.shape == (1, 3, 2368, 3200) and K2[0, 0, 0] == 2
    for yy, xx in ((0, 0), (0, -1), (-1, 0), (-1, -1)):
        assert torch.equal(big[..., yy, xx], img[..., yy, xx])      # align_corners=True
    assert float(big.min()) >= float(img.min()) and float(big.max()) <= float(img.max())
    disp = torch.rand(592, 800, device="cuda", generator=g) * 2.5e-3 + 1e-4
    depth = prep.disp_to_depth(disp)
    back = prep.disp_to_depth(depth)
    assert torch.allclose(back, disp, rtol=3e-7, atol=0)
    merged = prep.multires_merge(depth[::2, ::2].contiguous(), depth, th=1e9)     # everything consistent -> im2
    assert torch.equal(merged, depth)


def test_cpu_tensors_raise():
    with pytest.raises(RuntimeError):
        prep.normalize_images(torch.zeros(4))
    with pytest.raises(RuntimeError):
        prep.disp_to_depth(torch.zeros(2, 2))
