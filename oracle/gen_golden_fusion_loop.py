"""Golden for the WHOLE ``fusion()`` of the reference (fusion.py:110-318): PFM depth maps + images on a temporary
folder, a list standing in for the DataLoader, the reference's own loading / resizing / threshold bisection / masks /
PLY export, run on the CPU (``.cuda()`` patched to the identity).  Writes tests/golden/fusion_loop.npz: the inputs, the
final masks (read back from the PNGs the reference saves) and the vertices of result.ply.
TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference, cv2).

    python oracle/gen_golden_fusion_loop.py
"""
import os
import sys
import tempfile
import types
from pathlib import Path

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CER_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

from gen_golden_fusion import make_case  # noqa: E402

if not hasattr(np, "bool"):          # fusion.py:33 / datasets use the numpy < 1.24 aliases
    np.bool = bool
    np.float = float
    np.int = int
torch.Tensor.cuda = lambda self, *a, **k: self
_zeros = torch.zeros
sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))

import fusion as ref_fusion  # noqa: E402
from utils.frame_utils import write_pfm  # noqa: E402


def read_ply_vertices(path):
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    n = int([l for l in raw[:end].decode().split("\n") if l.startswith("element vertex")][0].split()[-1])
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    return np.frombuffer(raw[end:end + n * dt.itemsize], dtype=dt), raw[:end]


def main():
    seed, h, w, S = 3, 48, 64, 4
    depths, K, E = make_case(seed, h, w, S)
    n = S + 1
    rs = np.random.RandomState(seed)
    images = rs.uniform(0, 255, (n, 3, h, w)).astype(np.float32)          # image size == depth size: scale 1, no crop
    pairs = [(i, [j for j in range(n) if j != i]) for i in range(n)]
    glb = 0.6
    with tempfile.TemporaryDirectory() as td:
        out = Path(td)
        (out / "depths").mkdir()
        for i in range(n):
            write_pfm(out / "depths" / f"{i}_s.pfm", depths[i])
        loader = []
        for ref, srcs in pairs:
            ids = [ref] + srcs
            loader.append((torch.from_numpy(images[ids])[None], torch.from_numpy(E[ids])[None],
                           torch.from_numpy(K[ids])[None], [(str(j),) for j in ids], 1.0))
        ref_fusion.fusion(loader, out, suffix="_s", glb=glb, rescale=1)
        masks = np.stack([cv2.imread(str(out / "mask" / f"{i}_s.png"), cv2.IMREAD_GRAYSCALE) > 0 for i in range(n)])
        verts, header = read_ply_vertices(out / "result.ply")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fusion_loop.npz"), depths=depths, K=K, E=E, images=images,
                        glb=glb, masks=masks, ply_xyz=np.stack([verts["x"], verts["y"], verts["z"]], 1),
                        ply_rgb=np.stack([verts["red"], verts["green"], verts["blue"]], 1),
                        ply_header=np.frombuffer(header, dtype=np.uint8),
                        pairs=np.array([[r] + s for r, s in pairs]))
    print("masks kept", masks.mean(axis=(1, 2)), "vertices", len(verts))


if __name__ == "__main__":
    main()
