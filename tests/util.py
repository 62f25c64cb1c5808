"""Shared helpers for the GPU parity tests."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

H, W, V = 80, 112, 3
h1, w1 = H // 4, W // 4


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def cuda(x):
    return t(x).cuda() if isinstance(x, np.ndarray) else x.cuda()


def rel_l1(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-30))


def ref_ext():
    """The reference's own alt_cuda_corr compiled from /root/reference sources (oracle/_ref), or None."""
    try:
        import build_ref
        return build_ref.load()
    except Exception:  # noqa: BLE001
        return None
