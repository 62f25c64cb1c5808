"""Encoder timing at a BASELINE size (run on the GPU box): fnet per image, cnet, whole RAFT.forward."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cer_mvs_b200 import synth  # noqa: E402
from cer_mvs_b200.extractor import BasicEncoder  # noqa: E402
from cer_mvs_b200.raft import RAFT  # noqa: E402
from test_gpu_raft import _inputs, _state_dict  # noqa: E402

H, W, V = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2_dtu_1184x1600_v10"]
t = torch.from_numpy
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


img = t(synth.make_image(H, W, n=1, seed=1)).cuda()
fnet = BasicEncoder(output_dim=64, norm_fn="instance").cuda().eval()
cnet = BasicEncoder(output_dim=128, norm_fn="none").cuda().eval()
with torch.no_grad():
    print(f"fnet one image {H}x{W}: {timed(lambda: fnet.forward_features(img, normalize=True)):.3f} ms  (71 GFLOP)")
    print(f"cnet one image: {timed(lambda: cnet.forward_context(img, normalize=True)):.3f} ms")
    images, poses, K = _inputs(H, W, V, 1)
    m = RAFT(cascade=[(64, 64, 16), (-1, 320, 16)], test_mode=True)
    m.load_state_dict(_state_dict(1, 0.1, 0.005))
    m = m.cuda().eval()
    ms = timed(lambda: m(images, poses, K, scale=1.0), 5)
    print(f"whole RAFT.forward {H}x{W}, {V} views, 16+16: {ms:.2f} ms = {1e3 / ms:.1f} depth-maps/s")
