"""Golden for SURVEY 8f row 1: the reference's own BasicEncoder (core/extractor.py) -- fnet (instance norm, 64
channels) and cnet (no norm, 128 channels) -- on a small seeded image, fp32 on the CPU.  Writes
tests/golden/ops_encoder.npz (inputs are re-generated from the seed; weights = the module's own seeded init).
TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference).

    python oracle/gen_golden_encoder.py
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

from core.extractor import BasicEncoder  # noqa: E402

from cer_mvs_b200 import synth  # noqa: E402

H, W = 72, 104


def main():
    out = {}
    img = synth.make_image(H, W, n=1, seed=17)
    x = torch.from_numpy(img) * (2 / 255.) - 1
    for name, dim, norm, seed in (("fnet", 64, "instance", 100), ("cnet", 128, "none", 101)):
        enc = BasicEncoder(output_dim=dim, norm_fn=norm, type="HR").eval()
        sd = synth.make_encoder_weights(seed=seed, out_dim=dim)
        enc.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)     # exactly the reference's keys
        with torch.no_grad():
            y = enc(x)
        out[f"{name}_out"] = y.numpy()
        out[f"{name}_seed"] = seed
    out["image_seed"] = 17
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ops_encoder.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_out")}, os.path.getsize(os.path.join(ROOT, "tests", "golden", "ops_encoder.npz")))


if __name__ == "__main__":
    main()
