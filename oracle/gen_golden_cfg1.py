"""Golden for BASELINE.json configs[0] at its FULL size: the reference's own RAFT.forward (fp32, CPU) on the
448x576, 2-source-view, 2+2-iteration case.  Runs only in the build container (needs /root/reference);
writes tests/golden/e2e_fp32_cfg1.npz (the disparity, 112x144 floats; inputs are re-generated from the seed).

    python oracle/gen_golden_cfg1.py

TEST INFRASTRUCTURE ONLY.  The reference is imported unmodified (baseline/refrun.py, mode "cpu": alt_cuda_corr.forward
is CUDA-only and is served by cer_oracle.corr_forward, which the GPU tests pin against the compiled reference kernel).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))

import refrun  # noqa: E402
from cer_mvs_b200 import synth  # noqa: E402

H, W, V = synth.CONFIGS["cfg1_dtu_448x576_v2"]
CASCADE = [(64, 64, 2), (-1, 320, 2)]
SEED, DSCALE, DBIAS = 21, 0.1, 0.02

if __name__ == "__main__":
    torch.manual_seed(0)
    ref = refrun.import_reference("cpu")
    sc = synth.make_scene(H, W, V, seed=SEED)
    sd = synth.make_update_weights(seed=SEED, delta_scale=DSCALE, delta_bias=DBIAS)
    pre = torch.from_numpy(synth.make_context_pre(H // 4, W // 4, seed=SEED))
    t = torch.from_numpy
    model = refrun.make_model(ref, sd, CASCADE, t(sc["fmaps"]), pre, "cpu")
    images = torch.zeros(1, V + 1, 3, H, W)
    t0 = time.perf_counter()
    disp = refrun.run_forward(model, images, t(sc["poses"]), t(sc["intrinsics"]), 1.0)
    dt = time.perf_counter() - t0
    out = os.path.join(ROOT, "tests", "golden", "e2e_fp32_cfg1.npz")
    np.savez_compressed(out, disp=disp.numpy().astype(np.float32), seed=SEED, delta_scale=DSCALE, delta_bias=DBIAS,
                        scale=1.0, cascade=np.array(CASCADE), cpu_seconds=dt, threads=torch.get_num_threads())
    print("cfg1 golden:", disp.shape, disp.dtype, "mean", float(disp.mean()), f"{dt:.2f}s on {torch.get_num_threads()} threads",
          os.path.getsize(out), "bytes")
