#!/usr/bin/env python
"""Per-kernel time vs image size (tiles per SM) to separate fixed launch/prologue cost from per-tile cost."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cer_mvs_b200 import synth  # noqa: E402
from cer_mvs_b200.hotpath import DepthHotPath  # noqa: E402

V = int(os.environ.get("V", "2"))
out = {}
for h1, w1 in [(64, 296), (128, 296), (256, 296), (296, 400)]:
    sc = synth.make_scene(4 * h1, 4 * w1, V, seed=0)
    sd = synth.make_update_weights(seed=0, delta_scale=0.1, delta_bias=0.005)
    t = torch.from_numpy
    hp = DepthHotPath(h1, w1, max_views=V, cascade=[(64, 64, 8), (-1, 320, 8)], use_graph=False)
    hp.load_update_block(sd)
    args = (t(sc["fmaps"]).cuda().half(), t(sc["net"]).cuda().half(), t(sc["inp"]).cuda().half(),
            t(sc["poses"]).cuda(), t(sc["intrinsics"]).cuda(), 1.0)
    hp.set_kernel_timing(True)
    hp(*args)
    hp.kernel_times()
    for _ in range(3):
        hp(*args)
    kt = hp.kernel_times()
    tiles = ((h1 + 15) // 16) * ((w1 + 7) // 8)
    out[f"{h1}x{w1}"] = {"tiles_per_sm": round(tiles / 148, 2), **{k: round(1e3 * ms / n, 1) for k, (ms, n) in kt.items() if n}}
    print(f"{h1}x{w1}", out[f"{h1}x{w1}"], flush=True)
    del hp
json.dump(out, open("gpurun_out/kernel_scaling.json", "w"), indent=1)
