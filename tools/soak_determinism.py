#!/usr/bin/env python
"""Soak test: the hot path at BASELINE cfg 2 (16+16 iterations) N times on the same inputs, every output compared bit for
bit with the first (catches rare intra-kernel races that a two-run determinism test can miss)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cer_mvs_b200 import synth  # noqa: E402
from cer_mvs_b200.hotpath import DepthHotPath  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
H, W, V = synth.CONFIGS["cfg2_dtu_1184x1600_v10"]
sc = synth.make_scene(H, W, V, seed=0)
sd = synth.make_update_weights(seed=0)
t = torch.from_numpy
args = (t(sc["fmaps"]).cuda().half(), t(sc["net"]).cuda().half(), t(sc["inp"]).cuda().half(),
        t(sc["poses"]).cuda(), t(sc["intrinsics"]).cuda(), 1.0)
for use_graph in (True, False):
    hp = DepthHotPath(H // 4, W // 4, max_views=V, cascade=[(64, 64, 16), (-1, 320, 16)], use_graph=use_graph)
    hp.load_update_block(sd)
    first = hp(*args).clone()
    bad = 0
    for i in range(n if use_graph else n // 10):
        out = hp(*args)
        if not torch.equal(out, first):
            bad += 1
    torch.cuda.synchronize()
    print(f"graph={use_graph}: {bad} of {n if use_graph else n // 10} runs differ from the first; finite: {bool(torch.isfinite(first).all())}")
    assert bad == 0
print("soak ok")
