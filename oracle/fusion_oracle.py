"""CPU restatement (numpy, float32) of the reference's geometric-consistency check, fusion.py:39-106, and of the
per-view aggregation of fusion(), fusion.py:239-249.  TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else; the
product path is csrc/fusion_ops.cu.  Pinned by tests/golden/ops_fusion.npz = outputs of the reference's own functions
(oracle/gen_golden_fusion.py), see tests/test_oracle_golden.py.  Third-party arithmetic restated: torch
``F.grid_sample(bilinear, zeros, align_corners=True)`` (weights (x1-x)(y1-y) etc., taps outside the image are zero)."""
import numpy as np

f32 = np.float32


def _grid_sample(img, x, y):
    h, w = img.shape
    xn = (f32(2) * x / f32(w - 1) - f32(1)).astype(f32)                # utils/bilinear_sampler.py:36-37
    yn = (f32(2) * y / f32(h - 1) - f32(1)).astype(f32)
    ix = ((xn + f32(1)) / f32(2) * f32(w - 1)).astype(f32)
    iy = ((yn + f32(1)) / f32(2) * f32(h - 1)).astype(f32)
    fx, fy = np.floor(ix), np.floor(iy)
    out = np.zeros_like(ix)
    for dy, wy in ((0, (fy + 1) - iy), (1, iy - fy)):
        for dx, wx in ((0, (fx + 1) - ix), (1, ix - fx)):
            xx, yy = fx + dx, fy + dy
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            v = img[np.clip(yy, 0, h - 1).astype(np.int64), np.clip(xx, 0, w - 1).astype(np.int64)]
            out = out + np.where(ok, v, f32(0)) * (wx * wy).astype(f32)
    return out.astype(f32)


def check_geometric_consistency(depth_ref, K_ref, E_ref, depth_src, K_src, E_src, thre1=4.4, thre2=1430.0):
    """depth_ref [h,w]; depth_src [S,h,w]; returns (masks [9,S,h,w] bool, depth_reprojected, x_src, y_src, rel) [S,h,w]."""
    S, h, w = depth_src.shape
    ys, xs = np.mgrid[0:h, 0:w]
    xr, yr = xs.reshape(-1).astype(f32), ys.reshape(-1).astype(f32)
    d = depth_ref.reshape(-1).astype(f32)
    masks = np.zeros((9, S, h, w), bool)
    drep, xo, yo, rel = (np.zeros((S, h, w), f32) for _ in range(4))
    for s in range(S):
        A = np.linalg.inv(K_ref.astype(f32)).astype(f32)                                   # :49
        xyz_ref = A @ (np.stack([xr, yr, np.ones_like(xr)]) * d)                           # :50-52
        T1 = (E_src[s].astype(f32) @ np.linalg.inv(E_ref.astype(f32)).astype(f32)).astype(f32)
        xyz_src = (T1 @ np.concatenate([xyz_ref, np.ones((1, h * w), f32)]))[:3]           # :55-56
        kx = K_src[s].astype(f32) @ xyz_src                                                # :58
        x_src, y_src = (kx[0] / kx[2]).astype(f32), (kx[1] / kx[2]).astype(f32)            # :59
        sd = _grid_sample(depth_src[s].astype(f32), x_src, y_src)                          # :68
        xyz2 = np.linalg.inv(K_src[s].astype(f32)).astype(f32) @ (np.stack([x_src, y_src, np.ones_like(x_src)]) * sd)
        T2 = (E_ref.astype(f32) @ np.linalg.inv(E_src[s].astype(f32)).astype(f32)).astype(f32)
        xyz_rep = (T2 @ np.concatenate([xyz2, np.ones((1, h * w), f32)]))[:3]              # :75-76
        depth_rep = xyz_rep[2].astype(f32)
        kr = K_ref.astype(f32) @ xyz_rep
        xrep, yrep = (kr[0] / kr[2]).astype(f32), (kr[1] / kr[2]).astype(f32)
        dist = np.sqrt((xrep - xr) ** 2 + (yrep - yr) ** 2).astype(f32)                    # :96
        r = (np.abs(depth_rep - d) / d).astype(f32)                                        # :99-100
        for i in range(2, 11):
            masks[i - 2, s] = ((dist < f32(i / thre1)) & (r < f32(i / thre2))).reshape(h, w)   # :103-105
        drep[s] = np.where(masks[8, s], depth_rep.reshape(h, w), f32(0))                   # :106
        xo[s], yo[s], rel[s] = x_src.reshape(h, w), y_src.reshape(h, w), r.reshape(h, w)
    return masks, drep, xo, yo, rel


def aggregate(masks, drep, depth_ref):
    """fusion.py:239-249."""
    S = masks.shape[1]
    n = S + 1
    sums = masks.sum(1)                                  # [9,h,w]
    keep = sums[8] >= n
    for i in range(2, n):
        keep = keep | (sums[i - 2] >= i)
    depth_est = ((drep.sum(0) + depth_ref) / (sums[8] + 1)).astype(f32)
    return keep, depth_est


def fuse(all_depths, all_K, all_E, pair_data, glb=0.25, tot_iter=10):
    """The view / bisection loop of fusion(), fusion.py:201-299 (without the image and PLY handling)."""
    n, h, w = all_depths.shape
    left, right = -2.0, 2.0
    masks = np.zeros((n, h, w), bool)
    depth_est = np.zeros((n, h, w), f32)
    thre, ratios = 0.0, []
    for it in range(tot_iter):
        thre = (left + right) / 2
        ratios = []
        for ref, srcs in pair_data:
            srcs = list(srcs)
            m, drep, _, _, _ = check_geometric_consistency(all_depths[ref], all_K[ref], all_E[ref], all_depths[srcs],
                                                           all_K[srcs], all_E[srcs], 10 ** thre * 4, 10 ** thre * 1300)
            masks[ref], depth_est[ref] = aggregate(m, drep, all_depths[ref])
            ratios.append(float(masks[ref].astype(f32).mean()))
        if it < tot_iter - 1:
            if np.mean(ratios) >= glb:
                left = thre
            else:
                right = thre
    return thre, ratios, masks, depth_est
