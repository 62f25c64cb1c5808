"""GPU: the fused geometric-consistency filter (csrc/fusion_ops.cu) against golden outputs of the
reference's own check_geometric_consistency / fusion():239-249 (tests/golden/ops_fusion.npz) and the numpy oracle."""
import numpy as np
import pytest
import torch

import fusion_oracle as FO
from cer_mvs_b200 import fusion_ops
from util import t

pytestmark = pytest.mark.gpu


def _inputs(g, case):
    depths, K, E = g[f"{case}_depths"], g[f"{case}_K"], g[f"{case}_E"]
    t1, t2 = (float(v) for v in g[f"{case}_thre"])
    return depths, K, E, t1, t2


@pytest.mark.parametrize("case", ["a", "b"])
def test_check_geometric_consistency_dropin(golden, case):
    g = golden("ops_fusion")
    depths, K, E, t1, t2 = _inputs(g, case)
    S = depths.shape[0] - 1
    rep = lambda x: t(x).cuda().unsqueeze(0).repeat(S, *([1] * x.ndim))     # noqa: E731  (the reference repeats the ref view)
    masks, mask, drep, xs, ys, rel = fusion_ops.check_geometric_consistency(
        rep(depths[0]), rep(K[0]), rep(E[0]), t(depths[1:]).cuda(), t(K[1:]).cuda(), t(E[1:]).cuda(), t1, t2)
    m = torch.stack(masks).cpu().numpy()
    assert len(masks) == 9 and torch.equal(mask, masks[-1]) and m.dtype == bool
    np.testing.assert_allclose(xs.cpu().numpy(), g[f"{case}_x_src"], rtol=0, atol=2e-3)      # pixels
    np.testing.assert_allclose(ys.cpu().numpy(), g[f"{case}_y_src"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(rel.cpu().numpy(), g[f"{case}_rel"], rtol=0, atol=2e-6)
    assert (m != g[f"{case}_masks"]).mean() < 2e-3        # thresholds are discontinuous: a few boundary pixels may flip
    same = (m == g[f"{case}_masks"]).all(axis=(0, 1))
    np.testing.assert_allclose(drep.cpu().numpy()[:, same], g[f"{case}_depth_reprojected"][:, same], rtol=2e-6)


@pytest.mark.parametrize("case", ["a", "b"])
def test_geometric_filter_fused(golden, case):
    g = golden("ops_fusion")
    depths, K, E, t1, t2 = _inputs(g, case)
    keep, depth_est, ratio = fusion_ops.geometric_filter(t(depths[0]).cuda(), t(K[0]).cuda(), t(E[0]).cuda(),
                                                         t(depths[1:]).cuda(), t(K[1:]).cuda(), t(E[1:]).cuda(), t1, t2)
    keep, depth_est = keep.cpu().numpy(), depth_est.cpu().numpy()
    assert (keep != g[f"{case}_geo_mask"]).mean() < 3e-3
    assert abs(ratio - keep.mean()) < 1e-9 and abs(ratio - g[f"{case}_geo_mask"].mean()) < 3e-3
    # the averaged depth where every mask bit of the oracle agrees with the reference
    masks, drep, _, _, _ = FO.check_geometric_consistency(depths[0], K[0], E[0], depths[1:], K[1:], E[1:], t1, t2)
    same = (masks == g[f"{case}_masks"]).all(axis=(0, 1))
    close = np.isclose(depth_est, g[f"{case}_depth_est"], rtol=3e-6, atol=0)
    assert close[same].mean() > 0.998


def test_identity_views_known_answer():
    """Every source view == the reference view: every mask passes and depth_est == depth up to the bilinear resampling of a
    NOISY depth map at coordinates that are only equal to the pixel centres up to fp32 rounding (~1e-5 pixel x neighbour
    differences of up to 400 -> a few 1e-5 relative)."""
    h, w, S = 296, 400, 10
    g = torch.Generator(device="cuda").manual_seed(0)
    d = torch.rand(h, w, device="cuda", generator=g) * 400 + 400
    K = torch.tensor([[700.0, 0, w / 2], [0, 700.0, h / 2], [0, 0, 1]], device="cuda")
    E = torch.eye(4, device="cuda")
    keep, depth_est, ratio = fusion_ops.geometric_filter(d, K, E, d[None].repeat(S, 1, 1), K[None].repeat(S, 1, 1),
                                                         E[None].repeat(S, 1, 1), 4.4, 1430.0)
    assert bool(keep.all()) and ratio == 1.0
    torch.testing.assert_close(depth_est, d, rtol=2e-4, atol=0)


def test_cpu_tensors_raise():
    with pytest.raises(RuntimeError):
        fusion_ops.geometric_filter(torch.zeros(4, 4), torch.eye(3), torch.eye(4), torch.zeros(1, 4, 4), torch.eye(3)[None],
                                    torch.eye(4)[None], 4.4, 1430.0)


def test_fuse_depth_maps_bisection_loop(golden):
    """fusion():201-299 -- threshold bisection over all reference views -- against the numpy oracle of the same loop.
    The bisection is discontinuous (a kept fraction that lands next to ``glb`` can send the two implementations down
    different branches), so the final threshold is compared up to the last two bisection steps and the masks up to the
    pixels such a threshold difference moves."""
    g = golden("ops_fusion")
    depths, K, E = g["a_depths"], g["a_K"], g["a_E"]
    n = depths.shape[0]
    pairs = [(i, [j for j in range(n) if j != i]) for i in range(n)]
    out = fusion_ops.fuse_depth_maps(t(depths).cuda(), t(K).cuda(), t(E).cuda(), pairs, glb=0.6)
    thre, ratios, masks, depth_est = FO.fuse(depths, K, E, pairs, glb=0.6)
    assert abs(out["thre"] - thre) <= 4.0 / 2 ** 8
    assert abs(np.mean(out["ratios"]) - np.mean(ratios)) < 0.02 and abs(np.mean(ratios) - 0.6) < 0.05
    m = out["masks"].cpu().numpy()
    assert (m != masks).mean() < 0.03
    both = m & masks
    np.testing.assert_allclose(out["depth_est"].cpu().numpy()[both], depth_est[both], rtol=2e-3)
    # world points of view 0 (identity pose): x = (u - cx) z / fx
    p0 = out["points"][0].cpu().numpy()
    assert p0.shape == (int(m[0].sum()), 3)
    vs, us = np.nonzero(m[0])
    z = out["depth_est"].cpu().numpy()[0][m[0]]
    np.testing.assert_allclose(p0[:, 2], z, rtol=1e-5)
    np.testing.assert_allclose(p0[:, 0], (us - K[0, 0, 2]) * z / K[0, 0, 0], rtol=1e-4, atol=1e-3)


def test_whole_fusion_against_reference_run(golden, tmp_path):
    """fusion() end to end -- PFM depth maps and images in, mask PNGs and result.ply out -- against a run of the
    reference's own fusion() on the same files (oracle/gen_golden_fusion_loop.py).  The threshold bisection is
    discontinuous: a kept fraction next to ``glb`` can send the two runs down different branches of the last bisection
    steps, so the masks may differ on the pixels such a threshold difference moves (< 2 %)."""
    import cv2
    from cer_mvs_b200 import prep
    g = golden("fusion_loop")
    depths, K, E, images = g["depths"], g["K"], g["E"], g["images"]
    n = depths.shape[0]
    (tmp_path / "depths").mkdir()
    for i in range(n):
        prep.write_pfm(tmp_path / "depths" / f"{i}_s.pfm", depths[i])
    loader = []
    for row in g["pairs"]:
        ids = [int(v) for v in row]
        loader.append((t(images[ids])[None], t(E[ids])[None], t(K[ids])[None], [(str(j),) for j in ids], 1.0))
    out = fusion_ops.fusion(loader, tmp_path, suffix="_s", glb=float(g["glb"]), rescale=1)
    masks = np.stack([cv2.imread(str(tmp_path / "mask" / f"{i}_s.png"), cv2.IMREAD_GRAYSCALE) > 0 for i in range(n)])
    assert np.array_equal(masks, out["masks"].cpu().numpy())
    diff = (masks != g["masks"]).mean()
    print(f"fusion(): mask pixels that differ from the reference run: {diff:.4f}; kept {masks.mean():.3f} vs {g['masks'].mean():.3f}")
    assert diff < 0.02
    # result.ply: same header layout as the reference's file, the vertices of the commonly kept pixels agree
    raw = open(tmp_path / "result.ply", "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    ref_header = g["ply_header"].tobytes().decode()
    got_header = raw[:end].decode()
    strip = lambda s: "\n".join(l for l in s.split("\n") if not l.startswith("element vertex"))  # noqa: E731
    assert strip(got_header) == strip(ref_header)
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    verts = np.frombuffer(raw[end:], dtype=dt)
    assert len(verts) == int(masks.sum())
    # align the two vertex lists through the masks (both are written view by view, row-major inside a view)
    both = masks & g["masks"]
    sel_got, sel_ref = both[masks], both[g["masks"]]
    xyz = np.stack([verts["x"], verts["y"], verts["z"]], 1)[sel_got]
    np.testing.assert_allclose(xyz, g["ply_xyz"][sel_ref], rtol=2e-3, atol=0.2)      # depth ~600: 0.2 = 3e-4 relative
    rgb = np.stack([verts["red"], verts["green"], verts["blue"]], 1)[sel_got]
    assert (np.abs(rgb.astype(int) - g["ply_rgb"][sel_ref].astype(int)) <= 1).mean() > 0.999
