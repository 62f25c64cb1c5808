import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from cer_mvs_b200 import _lib, synth
from cer_mvs_b200.hotpath import DepthHotPath
from util import H, V, W, h1, w1, t, rel_l1
g = dict(np.load("tests/golden/e2e_fp32_trained_like.npz"))
seed = int(g["seed"])
sc = synth.make_scene(H, W, V, seed=seed)
sd = synth.make_update_weights(seed=seed, delta_scale=float(g["delta_scale"]), delta_bias=float(g["delta_bias"]) if "delta_bias" in g else 0.0)
cascade = [tuple(int(v) for v in row) for row in g["cascade"]]
L = _lib.lib()
ref = None
for variant, flags, graph, pdl in itertools.product((1, 2), (0, 1), (False, True), (1,)):
    L.cer_set_conv_variant(variant); L.cer_set_tile_flags(flags)
    hp = DepthHotPath(h1, w1, max_views=V, cascade=cascade, feats_f16=False, use_graph=graph)
    hp.load_update_block(sd)
    outs = []
    for rep in range(3):
        out = hp(t(sc["fmaps"]).cuda(), t(g["net"]).cuda(), t(g["inp"]).cuda(), t(sc["poses"]).cuda(), t(sc["intrinsics"]).cuda(), scale=float(g["scale"]))
        outs.append(out.cpu().numpy().copy())
    if ref is None: ref = outs[0]
    print(f"variant {variant} flags {flags} graph {graph}: err vs golden {[round(rel_l1(o, g['disp']),5) for o in outs]} vs first {[round(rel_l1(o, ref),6) for o in outs]}", flush=True)
