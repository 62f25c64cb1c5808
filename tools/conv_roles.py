#!/usr/bin/env python
"""Where do the roles of the tcgen05 conv kernels wait?  (cycles of CTA 0, last launch of each kernel)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cer_mvs_b200 import _lib, synth  # noqa: E402
from cer_mvs_b200.hotpath import DepthHotPath  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
_lib.check(_lib.lib().cer_set_conv_variant(variant))
H, W, V = 1184, 1600, 2
sc = synth.make_scene(H, W, V, seed=0)
sd = synth.make_update_weights(seed=0, delta_scale=0.1, delta_bias=0.005)
t = torch.from_numpy
hp = DepthHotPath(H // 4, W // 4, max_views=V, cascade=[(64, 64, 2), (-1, 320, 2)], use_graph=False)
hp.load_update_block(sd)
buf = torch.zeros(4 * 32, dtype=torch.int64, device="cuda")
if not os.environ.get("NO_PROF"):
    _lib.lib().cer_debug_set_conv_profile(buf.data_ptr())
args = (t(sc["fmaps"]).cuda().half(), t(sc["net"]).cuda().half(), t(sc["inp"]).cuda().half(),
        t(sc["poses"]).cuda(), t(sc["intrinsics"]).cuda(), 1.0)
for _ in range(2):
    hp(*args)
torch.cuda.synchronize()
b = buf.cpu().view(4, 4, 8)
names = ["corr-enc 3x3 (N=64)", "gates (N=192)", "q/GRU (N=64)", "delta (2 x N=128, resident halves)"]
roles = ["epilogue warp0: wait acc_full", "mma: wait acc_empty / a_full / b_full / issue", "B producer: wait b_empty", "A prod: a_empty / - / - / issue / dn fill / dn wait"]
print(f"variant {variant}; cycles of CTA 0 (1 cycle ~ 0.52 ns at 1.92 GHz)")
for k in range(4):
    print(names[k])
    for r in range(4):
        tot = int(b[k, r, 7])
        w = [int(x) for x in b[k, r, :6]]
        print(f"   {roles[r]:42s} total {tot:8d}  waits {w}  ({100 * sum(w[:3]) / max(tot, 1):.0f}% waiting)")
