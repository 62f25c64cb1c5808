"""Generate tests/golden/ops_io.npz with the REFERENCE's own code for SURVEY 8f rows 2-3: ``scale_operation``
(utils/data_utils.py:58-66), ``write_pfm`` / ``readPFM`` (utils/frame_utils.py), the whole ``multires()`` function
(multires.py:16-40, run on a temporary folder of PFM files) and the two one-liners of core/raft.py:40-41 and
inference.py:57-58 (executed verbatim).  TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference, cv2).

    python oracle/gen_golden_io.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CER_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, REF)

from utils.data_utils import scale_operation  # noqa: E402
from utils.frame_utils import readPFM, write_pfm  # noqa: E402
import types  # noqa: E402

# matplotlib is absent here and only used by multires(visualize=True): an empty stand-in lets the module import
sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
import multires as ref_multires  # noqa: E402


def main():
    rs = np.random.RandomState(11)
    out = {}
    # scale_operation, rescale = 2 and a non-integer factor
    images = torch.from_numpy(rs.uniform(0, 255, (2, 3, 24, 40)).astype(np.float32))
    K = torch.tensor([[[700.0, 0, 20], [0, 690, 12], [0, 0, 1]]] * 2)
    for name, s in (("s2", 2), ("s15", 1.5)):
        im2, K2 = scale_operation(images.clone(), K.clone(), s)
        out[f"scale_{name}_out"] = im2.numpy()
        out[f"scale_{name}_K"] = K2.numpy()
    out["scale_in"] = images.numpy()
    out["scale_K_in"] = K.numpy()
    # core/raft.py:40-41
    n = images.clone()
    n *= 2 / 255.
    n -= 1
    out["norm_out"] = n.numpy()
    # inference.py:57-58 + write_pfm
    res = rs.uniform(-1e-4, 2.5e-3, (30, 44)).astype(np.float32)
    res[rs.rand(30, 44) < 0.1] = 0.0
    with np.errstate(divide="ignore"):
        im = np.where(res == 0, 0, 1 / res).astype(np.float32)
    out["disp"] = res
    out["depth"] = im
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "d.pfm")
        write_pfm(p, im)
        out["pfm_bytes"] = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)
        assert np.array_equal(readPFM(p), im)
        # multires(): scale-1 map 30 x 44, scale-2 map 60 x 88 that agrees with it on about half of the pixels
        d1 = rs.uniform(400, 900, (30, 44)).astype(np.float32)
        import cv2
        d2 = cv2.resize(d1, (88, 60)) * (1 + rs.uniform(-0.04, 0.04, (60, 88))).astype(np.float32)
        d2 = d2.astype(np.float32)
        os.makedirs(os.path.join(td, "depths"))
        write_pfm(os.path.join(td, "depths", "scan_scale1.pfm"), d1)
        write_pfm(os.path.join(td, "depths", "scan_scale2.pfm"), d2)
        ref_multires.multires(td, th=0.02)
        out["multires_im1"] = d1
        out["multires_im2"] = d2
        out["multires_out"] = np.ascontiguousarray(readPFM(os.path.join(td, "depths", "scan_th0.02.pfm")))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ops_io.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
