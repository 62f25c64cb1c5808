"""Image preparation, depth output and the multi-resolution merge around the hot path,
under the reference's own function names and argument meaning.  These are DEVICE-side equivalents: the reference
calls ``scale_operation`` / ``crop_operation`` on CPU tensors before ``.cuda()`` (inference.py:45-50) and
``multires`` works on numpy arrays, so a caller moves the images to the GPU first (``images.cuda()``), then calls
these; ``install.install()`` does not substitute them.  CPU tensors raise -- there is no CPU path.

* ``scale_operation`` / ``crop_operation``  (utils/data_utils.py:58-79)
* ``normalize_images``                       (core/raft.py:40-41, ``images *= 2 / 255.; images -= 1``)
* ``disp_to_depth``                          (inference.py:57-58, ``np.where(res == 0, 0, 1 / res)``)
* ``write_pfm`` / ``readPFM``                (utils/frame_utils.py:138-164, 10-40)
* ``multires_merge``                         (multires.py:24-28)

The array work runs in csrc/io_ops.cu on CUDA tensors; there is no CPU fallback (CPU tensors raise).  PFM
reading/writing is host file I/O by nature; the writer takes the depth map already flipped by the kernel.
"""
import re
import sys

import numpy as np
import torch

from . import _lib


def _need_cuda_f32(x, name):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if x.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    return x.contiguous()


def scale_operation(images, intrinsics, s):
    """images [N,3,H,W] float32 cuda, intrinsics [N,3,3] (modified in place like the reference's) -> rescaled pair."""
    images = _need_cuda_f32(images, "images")
    n, c, ht1, wd1 = images.shape
    ht2, wd2 = int(s * ht1), int(s * wd1)
    intrinsics[:, 0] *= s
    intrinsics[:, 1] *= s
    out = torch.empty(n, c, ht2, wd2, device=images.device, dtype=torch.float32)
    with torch.cuda.device(images.device):
        _lib.check(_lib.lib().cer_resize_bilinear_ac(images.data_ptr(), out.data_ptr(), n * c, ht1, wd1, ht2, wd2,
                                                     _lib.stream_ptr()), "cer_resize_bilinear_ac")
    return out, intrinsics


def crop_operation(images, intrinsics, crop_h, crop_w):
    """Centre crop; a view plus an intrinsics shift (no kernel: the next stage reads the strided view once)."""
    ht1, wd1 = images.shape[2], images.shape[3]
    x0, y0 = (wd1 - crop_w) // 2, (ht1 - crop_h) // 2
    images = images[:, :, y0:y0 + crop_h, x0:x0 + crop_w]
    intrinsics[:, 0, 2] -= x0
    intrinsics[:, 1, 2] -= y0
    return images, intrinsics


def normalize_images(images, out=None):
    images = _need_cuda_f32(images, "images")
    out = torch.empty_like(images) if out is None else out
    with torch.cuda.device(images.device):
        _lib.check(_lib.lib().cer_normalize_images(images.data_ptr(), out.data_ptr(), images.numel(),
                                                   _lib.stream_ptr()), "cer_normalize_images")
    return out


def disp_to_depth(disp, flip_rows=False):
    """disp [..., h, w] float32 cuda -> depth of the same shape (rows bottom-up if flip_rows, as PFM stores them)."""
    disp = _need_cuda_f32(disp, "disp")
    h, w = disp.shape[-2:]
    if disp.numel() != h * w:
        raise RuntimeError("disp_to_depth: one map at a time")
    out = torch.empty_like(disp)
    with torch.cuda.device(disp.device):
        _lib.check(_lib.lib().cer_disp_to_depth(disp.data_ptr(), out.data_ptr(), h, w, int(flip_rows),
                                                _lib.stream_ptr()), "cer_disp_to_depth")
    return out


def multires_merge(im1, im2, th=0.02):
    """im1 [h1,w1] (scale-1 depth), im2 [h2,w2] (scale-2 depth), float32 cuda -> merged [h2,w2]."""
    im1, im2 = _need_cuda_f32(im1, "im1"), _need_cuda_f32(im2, "im2")
    out = torch.empty_like(im2)
    with torch.cuda.device(im2.device):
        _lib.check(_lib.lib().cer_multires_merge(im1.data_ptr(), im1.shape[0], im1.shape[1], im2.data_ptr(),
                                                 im2.shape[0], im2.shape[1], float(th), out.data_ptr(),
                                                 _lib.stream_ptr()), "cer_multires_merge")
    return out


_PFM_MAGIC = {1: b"Pf", 3: b"PF"}          # channels -> magic (grey / colour)


def write_pfm(file, image, scale=1, flipped=False):
    """PFM writer with the file layout of utils/frame_utils.py:138-164 (byte-identical files: tests/test_gpu_io.py):
    magic line, "width height", signed scale whose sign says little-endian, then the rows bottom-up as raw float32.
    image: float32 [H, W], [H, W, 1] or [H, W, 3] (numpy or tensor).  ``flipped=True``: the rows already are
    bottom-up (``disp_to_depth(..., flip_rows=True)`` did it on the device)."""
    arr = image.detach().cpu().numpy() if isinstance(image, torch.Tensor) else np.asarray(image)
    if arr.dtype != np.float32:
        raise Exception("Image dtype must be float32.")
    channels = 1 if arr.ndim == 2 else (arr.shape[2] if arr.ndim == 3 else 0)
    if channels not in _PFM_MAGIC:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    rows = arr if flipped else arr[::-1]
    little = arr.dtype.byteorder == "<" or (arr.dtype.byteorder in "=|" and sys.byteorder == "little")
    header = b"%s\n%d %d\n%f\n" % (_PFM_MAGIC[channels], arr.shape[1], arr.shape[0], -scale if little else scale)
    with open(file, "wb") as f:
        f.write(header)
        f.write(np.ascontiguousarray(rows).tobytes())


def readPFM(file):
    """PFM reader (the inverse of write_pfm; utils/frame_utils.py:10-40): float32 [H, W] or [H, W, 3], rows top-down."""
    with open(file, "rb") as f:
        magic = f.readline().strip()
        channels = {v: k for k, v in _PFM_MAGIC.items()}.get(magic)
        if channels is None:
            raise Exception("Not a PFM file.")
        dims = re.fullmatch(rb"(\d+)\s(\d+)\s", f.readline())
        if dims is None:
            raise Exception("Malformed PFM header.")
        width, height = int(dims.group(1)), int(dims.group(2))
        little = float(f.readline().strip()) < 0
        data = np.frombuffer(f.read(), dtype="<f4" if little else ">f4")
    shape = (height, width) if channels == 1 else (height, width, 3)
    return np.ascontiguousarray(data.reshape(shape)[::-1])
