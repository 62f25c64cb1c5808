"""Per-role counters of the staged cost-volume build at a BASELINE size (debug tool; run on the GPU box).
    python tools/build_profile.py [cfg] """
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cer_mvs_b200 import _lib, synth  # noqa: E402
from test_gpu_build_staged import STAGES, _build, _scene  # noqa: E402

NAMES = ["chunks", "views", "mma", "zero", "direct", "sum_c", "sum_n16", "plan_cyc", "prod_wait_b", "cons_wait",
         "cons_copy", "cons_blend", "mma_wait_b", "mma_wait_acc", "cons_cyc", "copied_cols"]

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2_dtu_1184x1600_v10"
H, W, V = synth.CONFIGS[cfg]
L = _lib.lib()
for stage in (0, 1):
    fm, poses, K, disp = _scene(H, W, V, seed=47, stage=stage)
    _build(0, fm, poses, K, disp, stage)
    prof = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.cer_debug_set_build_profile(prof.data_ptr())
    _build(0, fm, poses, K, disp, stage)
    L.cer_debug_set_build_profile(None)
    p = prof.cpu().numpy().astype(np.float64)
    d = dict(zip(NAMES, p))
    n_cta = 148
    print(f"--- {cfg} stage {stage} (D={STAGES[stage][0]})")
    print(f"chunks {d['chunks']:.0f}  (item, view) plans {d['views']:.0f}  mma {d['mma']:.0f} zero {d['zero']:.0f} "
          f"direct {d['direct']:.0f}  mean c {d['sum_c'] / d['chunks']:.2f}  mean n16 {d['sum_n16'] / max(d['mma'], 1):.0f}  "
          f"copied cols/chunk (thread 0's warp) {d['copied_cols'] / max(d['mma'], 1):.0f}")
    for k in ("plan_cyc", "prod_wait_b", "cons_cyc", "cons_wait", "cons_copy", "cons_blend", "mma_wait_b", "mma_wait_acc"):
        print(f"  {k:14s} {d[k] / n_cta / 1e3:10.1f} kcycles per CTA   ({d[k] / d['chunks']:8.0f} cycles per chunk)")
