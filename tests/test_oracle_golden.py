"""CPU: the oracle restatement (oracle/cer_oracle.py) against golden outputs of the reference's own
Python (tests/golden/*.npz, made by oracle/gen_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

import cer_oracle as O
from cer_mvs_b200 import synth
from conftest import rel_l1

H, W, V = 80, 112, 3
h1, w1 = H // 4, W // 4


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def _scene(seed):
    sc = synth.make_scene(H, W, V, seed=seed)
    K = t(sc["intrinsics"]).clone()
    K[:, :, :2] /= 4
    return t(sc["fmaps"]), t(sc["poses"]), K, sc


def test_projective_transform(golden):
    g = golden("ops_corrblock")
    fm, poses, K, _ = _scene(1)
    Pij = O.projection_matrices(poses, K, [0], [2])
    x = O.project_hypotheses(Pij, t(g["proj_disps"]))              # [1,1,h,w,D,2]
    ref = t(g["proj_x1"])[..., [0, 1]].permute(0, 1, 3, 4, 2, 5)
    assert torch.allclose(x, ref, rtol=1e-5, atol=1e-4)


def test_build_volume_and_pyramid(golden):
    g = golden("ops_corrblock")
    fm, poses, K, _ = _scene(1)
    for stage, (D, incre, shift) in enumerate([(64, 0.0025 / 64, True), (44, 0.0025 / 320, False)]):
        pyr, origin = O.build_volume(fm, poses, K, [0] * V, [1, 2, 3], D, incre,
                                     t(g[f"s{stage}_disp_in"]), shift)
        assert np.array_equal(origin.numpy(), g[f"s{stage}_origin"])
        for l in range(3):
            got = pyr[l].reshape(V * h1 * w1, -1).numpy()
            assert got.shape == g[f"s{stage}_pyr{l}"].shape
            np.testing.assert_allclose(got, g[f"s{stage}_pyr{l}"], rtol=1e-5, atol=1e-6)


def test_lookup(golden):
    g = golden("ops_corrblock")
    for stage, (D, incre) in enumerate([(64, 0.0025 / 64), (44, 0.0025 / 320)]):
        pyr = [t(g[f"s{stage}_pyr{l}"]).reshape(-1, 1, 1, g[f"s{stage}_pyr{l}"].shape[-1]) for l in range(3)]
        origin = t(g[f"s{stage}_origin"])
        for name in ("true", "zero", "far", "rand"):
            z = t(g[f"s{stage}_z_{name}"])
            out = O.lookup(pyr, origin, D, incre, z[:, [0] * V], radius=5)
            np.testing.assert_allclose(out.numpy(), g[f"s{stage}_lookup_{name}"], rtol=1e-5, atol=2e-6)


def test_update_block(golden):
    g = golden("ops_update")
    sd = O.to_torch_sd(synth.make_update_weights(seed=2, delta_scale=1.0))
    np.testing.assert_array_equal(O.disp_encoder(t(g["disp"])).numpy(), g["disp_enc"])
    for stage in (0, 1):
        net, delta = O.update_block(sd, t(g["net"]), t(g["inp"]), t(g["disp"]), t(g["corr_frames"]), stage)
        np.testing.assert_allclose(net.numpy(), g[f"net_out{stage}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(delta.numpy(), g[f"delta{stage}"], rtol=1e-4, atol=1e-7)


def _e2e(golden, name, autocast):
    g = golden(name)
    seed = int(g["seed"])
    sc = synth.make_scene(H, W, V, seed=seed)
    sd = O.to_torch_sd(synth.make_update_weights(seed=seed, delta_scale=float(g["delta_scale"]),
                                                 delta_bias=float(g["delta_bias"]) if "delta_bias" in g else 0.0))
    cascade = [tuple(int(v) for v in row) for row in g["cascade"]]
    out, trace = O.hot_path(sd, t(sc["fmaps"]), t(g["net"]), t(g["inp"]), t(sc["poses"]), t(sc["intrinsics"]),
                            cascade=cascade, scale=float(g["scale"]), autocast=autocast, return_all=True)
    return out.numpy(), g["disp"], trace, g["deltas"]


def test_e2e_fp32(golden):
    for name in ("trained_like", "unscaled_oob", "drift", "scaled_pose"):
        out, ref, trace, deltas = _e2e(golden, "e2e_fp32_" + name, autocast=False)
        assert ref.dtype == np.float64 and out.shape == ref.shape      # core/raft.py:108 -> float64
        err = rel_l1(out, ref)
        assert err < 1e-4, (name, err)


def test_e2e_cfg1_full_size(golden):
    """BASELINE configs[0] at its full size (448x576, 2 views, 2+2 iterations, fp32): the oracle against the
    reference's own un-extrapolated CPU run (oracle/gen_golden_cfg1.py)."""
    g = golden("e2e_fp32_cfg1")
    Hc, Wc, Vc = synth.CONFIGS["cfg1_dtu_448x576_v2"]
    seed = int(g["seed"])
    sc = synth.make_scene(Hc, Wc, Vc, seed=seed)
    sd = O.to_torch_sd(synth.make_update_weights(seed=seed, delta_scale=float(g["delta_scale"]),
                                                 delta_bias=float(g["delta_bias"])))
    pre = t(synth.make_context_pre(Hc // 4, Wc // 4, seed=seed))
    cascade = [tuple(int(v) for v in row) for row in g["cascade"]]
    out = O.hot_path(sd, t(sc["fmaps"]), torch.tanh(pre[:, :, :64]), torch.relu(pre[:, :, 64:]), t(sc["poses"]),
                     t(sc["intrinsics"]), cascade=cascade, scale=1.0, autocast=False).numpy()
    assert out.shape == g["disp"].shape == (1, 1, 112, 144)
    err = rel_l1(out, g["disp"])
    assert err < 1e-4, err


def test_e2e_autocast_emulation(golden):
    """The fp16-rounding emulation against the reference run under torch.autocast('cpu', float16)
    (proxy for the GPU autocast path, core/raft.py:55)."""
    out, ref, _, _ = _e2e(golden, "e2e_autocast_drift", autocast=True)
    err = rel_l1(out, ref)
    assert err < 1e-3, err


def test_io_oracle_matches_reference_functions(golden):
    """oracle/io_oracle.py against outputs of the reference's scale_operation / write_pfm / multires()."""
    import io_oracle as IO
    g = golden("ops_io")
    for name, s in (("s2", 2), ("s15", 1.5)):
        out, K = IO.scale_operation(g["scale_in"], g["scale_K_in"], s)
        np.testing.assert_allclose(out, g[f"scale_{name}_out"], rtol=2e-6, atol=2e-5)
        np.testing.assert_array_equal(K, g[f"scale_{name}_K"])
    np.testing.assert_array_equal(IO.normalize_images(g["scale_in"]), g["norm_out"])
    np.testing.assert_array_equal(IO.disp_to_depth(g["disp"]), g["depth"])
    assert IO.pfm_bytes(g["depth"]) == g["pfm_bytes"].tobytes()
    merged, im1r = IO.multires_merge(g["multires_im1"], g["multires_im2"], 0.02)
    # the select is discontinuous: compare away from the decision boundary |im1r - im2| == th * im1r
    margin = np.abs(np.abs(im1r - g["multires_im2"]) - np.float32(0.02) * im1r) > 1e-3
    assert margin.mean() > 0.99
    np.testing.assert_allclose(merged[margin], g["multires_out"][margin], rtol=1e-6)
    picked2 = g["multires_out"] == g["multires_im2"]
    assert 0.2 < picked2.mean() < 0.8                 # both branches of the select are exercised


@pytest.mark.parametrize("case", ["a", "b"])
def test_fusion_oracle_matches_reference_functions(golden, case):
    """oracle/fusion_oracle.py against the reference's check_geometric_consistency + fusion():239-249."""
    import fusion_oracle as FO
    g = golden("ops_fusion")
    depths, K, E, (t1, t2) = g[f"{case}_depths"], g[f"{case}_K"], g[f"{case}_E"], g[f"{case}_thre"]
    masks, drep, xs, ys, rel = FO.check_geometric_consistency(depths[0], K[0], E[0], depths[1:], K[1:], E[1:], t1, t2)
    np.testing.assert_allclose(xs, g[f"{case}_x_src"], rtol=0, atol=2e-3)         # pixels, image up to 64 wide
    np.testing.assert_allclose(ys, g[f"{case}_y_src"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(rel, g[f"{case}_rel"], rtol=0, atol=2e-6)
    assert (masks != g[f"{case}_masks"]).mean() < 2e-3                            # threshold comparisons are discontinuous
    same = (masks == g[f"{case}_masks"]).all(axis=(0, 1))
    np.testing.assert_allclose(drep[:, same], g[f"{case}_depth_reprojected"][:, same], rtol=2e-6)
    keep, depth_est = FO.aggregate(masks, drep, depths[0])
    assert (keep != g[f"{case}_geo_mask"]).mean() < 3e-3
    np.testing.assert_allclose(depth_est[same], g[f"{case}_depth_est"][same], rtol=2e-6)
    assert 0.3 < g[f"{case}_geo_mask"].mean() < 0.95


@pytest.mark.parametrize("kind", ["fnet", "cnet"])
def test_encoder_oracle_matches_reference_module(golden, kind):
    """oracle/cer_oracle.basic_encoder against the reference's own BasicEncoder (core/extractor.py, fp32 CPU run)."""
    g = golden("ops_encoder")
    dim = 64 if kind == "fnet" else 128
    sd = O.to_torch_sd(synth.make_encoder_weights(seed=int(g[f"{kind}_seed"]), out_dim=dim))
    x = t(synth.make_image(72, 104, n=1, seed=int(g["image_seed"]))) * (2 / 255.) - 1
    out = O.basic_encoder(sd, x, kind == "fnet", autocast=False).numpy()
    np.testing.assert_allclose(out, g[f"{kind}_out"], rtol=1e-4, atol=1e-4)
    net, inp = O.context_split(t(g["cnet_out"]))
    assert float(net.abs().max()) <= 1.0 and float(inp.min()) >= 0.0
