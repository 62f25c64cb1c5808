"""Generate tests/golden/ops_fusion.npz with the REFERENCE's own ``reproject_with_depth`` / ``check_geometric_consistency``
(fusion.py:39-106).  fusion.py imports plyfile / datasets (absent here), so the two function definitions are compiled
from the source text where it lies under /root/reference -- nothing is copied into the repo.  The per-view aggregation
of fusion() (fusion.py:239-249) is inline in a 200-line function and is re-executed here statement by statement.
TEST INFRASTRUCTURE ONLY; build container only.

    python oracle/gen_golden_fusion.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CER_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from utils.bilinear_sampler import bilinear_sampler  # noqa: E402
from cer_mvs_b200 import synth  # noqa: E402


def reference_functions():
    src = open(os.path.join(REF, "fusion.py")).read()
    a = src.index("def reproject_with_depth(")
    b = src.index("@gin.configurable()\ndef fusion(")
    ns = {"torch": torch, "bilinear_sampler": bilinear_sampler}
    exec(compile(src[a:b], os.path.join(REF, "fusion.py"), "exec"), ns)
    return ns["check_geometric_consistency"]


def make_case(seed, h, w, S):
    """Depth maps of a slanted plane seen from S + 1 cameras, with noise / outliers so that every mask level is mixed."""
    rs = np.random.RandomState(seed)
    poses, K = synth.make_cameras(S, h, w, seed=seed)
    nrm = np.array([0.05, -0.03, 1.0]); nrm /= np.linalg.norm(nrm)
    d_plane = 600.0 * nrm[2]
    depths = []
    for v in range(S + 1):
        R, t = poses[v, :3, :3].astype(np.float64), poses[v, :3, 3].astype(np.float64)
        ys, xs = np.mgrid[0:h, 0:w]
        rays = np.linalg.inv(K[v].astype(np.float64)) @ np.stack([xs.ravel(), ys.ravel(), np.ones(h * w)])
        # camera-space point z * ray; world X = R^T (z ray - t); plane n.X = d  ->  z = (d + n.R^T t) / (n.R^T ray)
        nr = nrm @ R.T
        z = (d_plane + nr @ t) / (nr @ rays)
        z = z.reshape(h, w) * (1 + rs.normal(0, 2e-4, (h, w)))
        bad = rs.rand(h, w) < 0.15
        z[bad] *= 1 + rs.uniform(-0.02, 0.02, bad.sum())
        depths.append(z.astype(np.float32))
    return np.stack(depths), K, poses


def main():
    check = reference_functions()
    out = {}
    for name, (seed, h, w, S, t1, t2) in {"a": (1, 48, 64, 4, 4.4, 1430.0), "b": (2, 37, 53, 3, 10 ** 0.5 * 4, 10 ** 0.5 * 1300)}.items():
        depths, K, E = make_case(seed, h, w, S)
        T = torch.from_numpy
        ref_d = T(depths[0]).unsqueeze(0).repeat(S, 1, 1)
        ref_K = T(K[0]).unsqueeze(0).repeat(S, 1, 1)
        ref_E = T(E[0]).unsqueeze(0).repeat(S, 1, 1)
        masks, geo_mask, drep, xs, ys, rel = check(ref_d, ref_K, ref_E, T(depths[1:]), T(K[1:]), T(E[1:]), t1, t2)
        # fusion.py:239-249 (n = 1 + len(src_views))
        n = 1 + S
        geo_mask_sums = []
        for i in range(2, n):
            geo_mask_sums.append(masks[i - 2].sum(dim=0).int())
        geo_mask_sum = geo_mask.sum(dim=0)
        keep = geo_mask_sum >= n
        for i in range(2, n):
            keep = torch.logical_or(keep, geo_mask_sums[i - 2] >= i)
        depth_est = (drep.sum(dim=0) + ref_d[0]) / (geo_mask_sum + 1)
        out.update({f"{name}_depths": depths, f"{name}_K": K, f"{name}_E": E, f"{name}_thre": np.array([t1, t2]),
                    f"{name}_masks": torch.stack(masks).numpy(), f"{name}_depth_reprojected": drep.numpy(),
                    f"{name}_x_src": xs.numpy(), f"{name}_y_src": ys.numpy(), f"{name}_rel": rel.numpy(),
                    f"{name}_geo_mask": keep.numpy(), f"{name}_depth_est": depth_est.numpy()})
        print(name, "mask fractions", [round(float(m.float().mean()), 3) for m in masks], "kept", float(keep.float().mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ops_fusion.npz"), **out)


if __name__ == "__main__":
    main()
