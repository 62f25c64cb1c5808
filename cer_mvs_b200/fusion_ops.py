"""Geometric-consistency filter of the reference's fusion stage (fusion.py:39-106, 239-249) on one
fused CUDA kernel (csrc/fusion_ops.cu), under the reference's own function name plus the aggregated form its main loop
needs.  CUDA tensors only -- no CPU fallback.

    masks, mask, depth_reprojected, x2d_src, y2d_src, rel = check_geometric_consistency(
        depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src, thre1, thre2)
    geo_mask, depth_est, ratio = geometric_filter(ref_depth, ref_K, ref_E, src_depths, src_Ks, src_Es, thre1, thre2)
"""
import torch

from . import _lib


def _f32(x, name):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    return x.to(torch.float32).contiguous()


def _run(ref_depth, ref_K, ref_E, src_depths, src_Ks, src_Es, thre1, thre2, full):
    ref_depth, src_depths = _f32(ref_depth, "depth_ref"), _f32(src_depths, "depth_src")
    ref_K, ref_E = _f32(ref_K, "intrinsics_ref"), _f32(ref_E, "extrinsics_ref")
    src_Ks, src_Es = _f32(src_Ks, "intrinsics_src"), _f32(src_Es, "extrinsics_src")
    S, h, w = src_depths.shape
    if ref_depth.shape != (h, w) or ref_K.shape != (3, 3) or ref_E.shape != (4, 4) or src_Ks.shape != (S, 3, 3) or \
            src_Es.shape != (S, 4, 4):
        raise RuntimeError("geometric filter: inconsistent shapes")
    dev = ref_depth.device
    L = _lib.lib()
    ws = torch.empty(L.cer_geo_mats_bytes(S), dtype=torch.uint8, device=dev)
    geo_mask = torch.empty(h, w, dtype=torch.uint8, device=dev)
    depth_est = torch.empty(h, w, dtype=torch.float32, device=dev)
    n_valid = torch.zeros(1, dtype=torch.int32, device=dev)
    if full:
        masks = torch.empty(9, S, h, w, dtype=torch.uint8, device=dev)
        drep, xs, ys, rel = (torch.empty(S, h, w, dtype=torch.float32, device=dev) for _ in range(4))
        extra = [t.data_ptr() for t in (masks, drep, xs, ys, rel)]
    else:
        masks = drep = xs = ys = rel = None
        extra = [None] * 5
    with torch.cuda.device(dev):
        _lib.check(L.cer_geometric_filter(ref_depth.data_ptr(), ref_K.data_ptr(), ref_E.data_ptr(), src_depths.data_ptr(),
                                          src_Ks.data_ptr(), src_Es.data_ptr(), S, h, w, float(thre1), float(thre2),
                                          ws.data_ptr(), *extra, geo_mask.data_ptr(), depth_est.data_ptr(),
                                          n_valid.data_ptr(), _lib.stream_ptr()), "cer_geometric_filter")
    return masks, drep, xs, ys, rel, geo_mask, depth_est, n_valid


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                thre1=4.4, thre2=1430.):
    """fusion.py:88-106.  The reference passes the reference view repeated once per source view ([S,h,w], [S,3,3],
    [S,4,4]); the first copy is used.  Returns (masks: list of 9 bool [S,h,w], mask = masks[-1], depth_reprojected
    (zero where mask is False), x2d_src, y2d_src, relative_depth_diff)."""
    masks, drep, xs, ys, rel, _, _, _ = _run(depth_ref[0], intrinsics_ref[0], extrinsics_ref[0], depth_src,
                                             intrinsics_src, extrinsics_src, thre1, thre2, full=True)
    mlist = [masks[i].bool() for i in range(9)]
    return mlist, mlist[-1], drep, xs, ys, rel


def geometric_filter(ref_depth, ref_K, ref_E, src_depths, src_Ks, src_Es, thre1, thre2):
    """One reference view of fusion()'s inner loop (fusion.py:228-251): (geo_mask bool [h,w], depth_est [h,w],
    geo_mask.float().mean()) without materialising any per-source tensor."""
    _, _, _, _, _, geo_mask, depth_est, n_valid = _run(ref_depth, ref_K, ref_E, src_depths, src_Ks, src_Es, thre1, thre2,
                                                       full=False)
    h, w = geo_mask.shape
    return geo_mask.bool(), depth_est, float(n_valid.item()) / float(h * w)


def fuse_depth_maps(all_depths, all_intrinsics, all_extrinsics, pair_data, glb=0.25, tot_iter=10, images=None):
    """The view loop of ``fusion()`` (fusion.py:201-299) on the fused filter: bisection of the consistency threshold
    (``thre`` in log10 units, between -2 and 2) until the mean kept fraction over all reference views meets ``glb``,
    then the final masks, averaged depths and world-space points.

    all_depths [N,h,w], all_intrinsics [N,3,3], all_extrinsics [N,4,4] (world -> camera), CUDA tensors;
    pair_data: list of (ref_index, [source indices]); images (optional) [N,h,w,3] in 0..1 for point colours.
    Returns dict(thre, ratios, masks [N,h,w] bool, depth_est [N,h,w], points: list of [M,3] world xyz per reference
    view, colors: list of [M,3] uint8 or None)."""
    all_depths = _f32(all_depths, "all_depths")
    all_intrinsics, all_extrinsics = _f32(all_intrinsics, "all_intrinsics"), _f32(all_extrinsics, "all_extrinsics")
    n_images, h, w = all_depths.shape
    thre_left, thre_right = -2.0, 2.0                                    # fusion.py:201-202
    masks = torch.zeros(n_images, h, w, dtype=torch.bool, device=all_depths.device)
    depth_est = torch.zeros(n_images, h, w, device=all_depths.device)
    thre, ratios = 0.0, []
    for it in range(tot_iter):
        thre = (thre_left + thre_right) / 2                              # :206
        ratios = []
        for ref_view, src_views in pair_data:
            src = torch.as_tensor(list(src_views), device=all_depths.device, dtype=torch.long)
            m, d, r = geometric_filter(all_depths[ref_view], all_intrinsics[ref_view], all_extrinsics[ref_view],
                                       all_depths[src], all_intrinsics[src], all_extrinsics[src],
                                       10 ** thre * 4, 10 ** thre * 1300)            # :233-237
            depth_est[ref_view] = d                                                   # :249
            masks[ref_view] = m
            ratios.append(r)                                                          # :253
        if it < tot_iter - 1:
            if sum(ratios) / len(ratios) >= glb:                                      # :296-299
                thre_left = thre
            else:
                thre_right = thre
    points, colors = [], ([] if images is not None else None)
    ys, xs = torch.meshgrid(torch.arange(h, device=all_depths.device), torch.arange(w, device=all_depths.device),
                            indexing="ij")
    for ref_view, _ in pair_data:                                                     # :274-291 (last iteration)
        valid = masks[ref_view]
        x, y, dep = xs[valid].double(), ys[valid].double(), depth_est[ref_view][valid].double()
        xyz_ref = torch.linalg.inv(all_intrinsics[ref_view].double()) @ (torch.stack([x, y, torch.ones_like(x)]) * dep)
        xyz_world = (torch.linalg.inv(all_extrinsics[ref_view].double()) @
                     torch.cat([xyz_ref, torch.ones_like(x)[None]]))[:3]
        points.append(xyz_world.T.float())
        if images is not None:
            colors.append((images[ref_view][valid] * 255).to(torch.uint8))
    return {"thre": thre, "ratios": ratios, "masks": masks, "depth_est": depth_est, "points": points, "colors": colors}


# ------------------------------------------------------------------------------------------------------------
# File layer of the reference's fusion() (fusion.py:110-200, 256-318): depth maps from PFM files, images and cameras
# from the data loader, masks to PNG, the fused point cloud to result.ply.  Host I/O by nature (cv2 / numpy for the
# files and the bilinear image resize, like the reference); the geometry runs on the GPU through fuse_depth_maps.
# ------------------------------------------------------------------------------------------------------------
def write_ply(path, xyz, rgb):
    """Point cloud -> binary little-endian PLY with the element / property layout plyfile writes for the reference's
    structured array (fusion.py:304-316): vertex {float x, y, z; uchar red, green, blue}."""
    import numpy as np
    xyz = np.ascontiguousarray(np.asarray(xyz, dtype="<f4").reshape(-1, 3))
    rgb = np.ascontiguousarray(np.asarray(rgb, dtype=np.uint8).reshape(-1, 3))
    if len(xyz) != len(rgb):
        raise RuntimeError("write_ply: one colour per point expected")
    rec = np.empty(len(xyz), dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    for i, n in enumerate(("x", "y", "z")):
        rec[n] = xyz[:, i]
    for i, n in enumerate(("red", "green", "blue")):
        rec[n] = rgb[:, i]
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
              "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % len(rec))
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rec.tobytes())


def _fit_image_to_depth(img, depth_hw, K):
    """fusion.py:150-171: bring the colour image to the depth map's size -- uniform scale so that one side matches, then
    a centred crop of the other -- and move the intrinsics with it.  img: H x W x 3 in 0..1; returns (image, K)."""
    import math
    import cv2
    dh, dw = depth_hw
    ih, iw = img.shape[:2]
    scale = float(dh) / ih                    # match the heights ...
    crop_rows = False
    if dw / iw > scale:                       # ... unless the depth map is relatively wider: match the widths
        scale = float(dw) / iw
        crop_rows = True
    img = cv2.resize(img, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
    K = K.copy()
    K[:2, :] *= scale
    if not crop_rows:
        off = int(math.ceil((img.shape[1] - dw) / 2))
        img = img[:, off:dw + off, :]
        K[0, 2] -= off
    else:
        off = int(math.ceil((img.shape[0] - dh) / 2))
        img = img[off:img.shape[0] - off, :, :]
        K[1, 2] -= off
    return img, K


def fusion(data_loader, output_folder, suffix="", glb=0.25, rescale=1, device=None):
    """Drop-in for the reference's ``fusion()`` (fusion.py:110-318), same arguments and the same files written:
    ``mask/{view}{suffix}.png`` per reference view and ``result.ply``.  ``data_loader`` yields
    (images [1,n,3,H,W] 0..255, extrinsics [1,n,4,4], intrinsics [1,n,3,3], image_names, _) with the reference view
    first; the depth maps are read from ``output_folder/depths/{name}{suffix}.pfm``.
    Returns the dict of fuse_depth_maps (threshold, masks, averaged depths, points, colours)."""
    import os
    from pathlib import Path

    import cv2
    import numpy as np

    from .prep import readPFM
    output_folder = Path(output_folder)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    name_to_index, pairs = {}, []
    imgs, depths, Ks, Es = [], [], [], []
    for i, (images, extrinsics, intrinsics, image_names, _) in enumerate(data_loader):
        ref_name = image_names[0][0]
        name_to_index[ref_name] = i
        pairs.append((ref_name, [x[0] for x in image_names[1:]]))
        img = images.squeeze(0)[0].permute(1, 2, 0).numpy() / 255.
        depth = np.ascontiguousarray(readPFM(output_folder / "depths" / f"{ref_name}{suffix}.pfm"))
        dh, dw = depth.shape
        depth = cv2.resize(depth, (int(dw * rescale), int(dh * rescale)))            # fusion.py:148
        img, K = _fit_image_to_depth(img, depth.shape, np.array(intrinsics[0][0], dtype=np.float64))
        if i > 0 and (img.shape != imgs[0].shape or depth.shape != depths[0].shape):
            # fusion.py:184-196: views of another size are copied into the first view's frame (top-left aligned)
            fi, fd = np.zeros_like(imgs[0]), np.zeros_like(depths[0])
            sh, sw = min(img.shape[0], fi.shape[0]), min(img.shape[1], fi.shape[1])
            fi[:sh, :sw] = img[:sh, :sw]
            sh, sw = min(depth.shape[0], fd.shape[0]), min(depth.shape[1], fd.shape[1])
            fd[:sh, :sw] = depth[:sh, :sw]
            img, depth = fi, fd
        imgs.append(img)
        depths.append(depth)
        Ks.append(K)
        Es.append(np.array(extrinsics[0][0], dtype=np.float64))
    pair_idx = [(name_to_index[r], [name_to_index[s] for s in srcs]) for r, srcs in pairs]
    t = torch.from_numpy
    images_t = t(np.stack(imgs)).to(dev)
    out = fuse_depth_maps(t(np.stack(depths)).float().to(dev), t(np.stack(Ks)).float().to(dev),
                          t(np.stack(Es)).float().to(dev), pair_idx, glb=glb, images=images_t)
    os.makedirs(output_folder / "mask", exist_ok=True)
    for ref_view, _ in pair_idx:
        cv2.imwrite(str(output_folder / "mask" / f"{ref_view}{suffix}.png"),
                    out["masks"][ref_view].cpu().numpy().astype(np.uint8) * 255)
    xyz = torch.cat(out["points"], 0).cpu().numpy()
    rgb = torch.cat(out["colors"], 0).cpu().numpy()
    write_ply(output_folder / "result.ply", xyz, rgb)
    return out
