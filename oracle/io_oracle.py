"""CPU restatement (numpy) of the reference's image-preparation / depth-output / multi-resolution-merge
arithmetic.  TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else; the product path
is csrc/io_ops.cu.  Pinned by tests/golden/ops_io.npz = outputs of the reference's own functions
(oracle/gen_golden_io.py), see tests/test_oracle_golden.py.

Third-party arithmetic restated here: torch ``F.interpolate(bilinear, align_corners=True)`` (torch 2.11 CPU kernel,
aten/src/ATen/native/cpu/UpSampleKernel.cpp: source index = dst * (in-1)/(out-1), weights 1-l / l, second tap
clamped) and OpenCV 4.13 ``cv2.resize(INTER_LINEAR)`` on float32 (source coordinate (x+0.5)*scale-0.5, taps clamped
with zero weight, horizontal pass then vertical pass in float32).
"""
import numpy as np

f32 = np.float32


def scale_operation(images, intrinsics, s):
    """utils/data_utils.py:58-66.  images [N,C,H,W] float32, intrinsics [N,3,3] -> (resized, scaled intrinsics)."""
    n, c, h, w = images.shape
    h2, w2 = int(s * h), int(s * w)
    K = intrinsics.copy()
    K[:, 0] *= s
    K[:, 1] *= s
    sy = f32(h - 1) / f32(h2 - 1) if h2 > 1 else f32(0)
    sx = f32(w - 1) / f32(w2 - 1) if w2 > 1 else f32(0)
    fy = (sy * np.arange(h2, dtype=f32)).astype(f32)
    fx = (sx * np.arange(w2, dtype=f32)).astype(f32)
    y0 = np.minimum(np.floor(fy).astype(np.int64), h - 1)
    x0 = np.minimum(np.floor(fx).astype(np.int64), w - 1)
    ly = np.clip(fy - y0.astype(f32), 0, 1).astype(f32)[:, None]
    lx = np.clip(fx - x0.astype(f32), 0, 1).astype(f32)[None, :]
    y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
    a, b = images[:, :, y0][:, :, :, x0], images[:, :, y0][:, :, :, x1]
    cc, d = images[:, :, y1][:, :, :, x0], images[:, :, y1][:, :, :, x1]
    one = f32(1)
    top = ((one - lx) * a + lx * b).astype(f32)
    bot = ((one - lx) * cc + lx * d).astype(f32)
    return ((one - ly) * top + ly * bot).astype(f32), K


def normalize_images(images):
    """core/raft.py:40-41: two in-place float32 tensor ops."""
    return ((images.astype(f32) * f32(2 / 255.)).astype(f32) - f32(1)).astype(f32)


def disp_to_depth(res):
    """inference.py:57-58."""
    with np.errstate(divide="ignore"):
        return np.where(res == 0, 0, 1 / res).astype(f32)


def pfm_bytes(image, scale=1):
    """utils/frame_utils.py:138-164 for a float32 H x W image on a little-endian host."""
    assert image.dtype == np.float32 and image.ndim == 2
    flipped = np.flipud(image)
    return b"Pf\n" + b"%d %d\n" % (image.shape[1], image.shape[0]) + b"%f\n" % (-scale) + flipped.tobytes()


def cv2_resize_linear(im, w2, h2):
    h1, w1 = im.shape
    sx, sy = w1 / w2, h1 / h2

    def taps(n_dst, n_src, scale):
        f = ((np.arange(n_dst) + 0.5) * scale - 0.5).astype(f32)
        i0 = np.floor(f).astype(np.int64)
        f = (f - i0.astype(f32)).astype(f32)
        lo, hi = i0 < 0, i0 >= n_src - 1
        i0 = np.where(lo, 0, np.where(hi, n_src - 1, i0))
        f = np.where(lo | hi, f32(0), f).astype(f32)
        return i0, np.minimum(i0 + 1, n_src - 1), f
    x0, x1, fx = taps(w2, w1, sx)
    y0, y1, fy = taps(h2, h1, sy)
    one = f32(1)
    rows = (im[:, x0] * (one - fx)[None, :] + im[:, x1] * fx[None, :]).astype(f32)          # horizontal pass
    return (rows[y0] * (one - fy)[:, None] + rows[y1] * fy[:, None]).astype(f32)            # vertical pass


def multires_merge(im1, im2, th=0.02):
    """multires.py:24-28."""
    im1r = cv2_resize_linear(im1, im2.shape[1], im2.shape[0])
    mask = np.abs(im1r - im2) < f32(th) * im1r
    return np.where(mask, im2, im1r), im1r
