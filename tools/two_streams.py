#!/usr/bin/env python
"""Throughput with N depth maps in flight on N streams (one plan each): the CTAs of one depth map's kernel fill the
SMs that the other's kernel leaves idle in its last tile round / pipeline fill."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cer_mvs_b200 import synth  # noqa: E402
from cer_mvs_b200.hotpath import DepthHotPath  # noqa: E402

H, W, V = 1184, 1600, 10
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
sc = synth.make_scene(H, W, V, seed=0)
sd = synth.make_update_weights(seed=0, delta_scale=0.1, delta_bias=0.005)
t = torch.from_numpy
args = (t(sc["fmaps"]).cuda().half(), t(sc["net"]).cuda().half(), t(sc["inp"]).cuda().half(), t(sc["poses"]).cuda(),
        t(sc["intrinsics"]).cuda(), 1.0)
for n in (1, 2, 3):
    hps = [DepthHotPath(H // 4, W // 4, max_views=V, cascade=[(64, 64, 16), (-1, 320, 16)]) for _ in range(n)]
    for hp in hps:
        hp.load_update_block(sd)
    streams = [torch.cuda.Stream() for _ in range(n)]
    for i in range(3 * n):
        with torch.cuda.stream(streams[i % n]):
            hps[i % n](*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for i in range(steps):
        with torch.cuda.stream(streams[i % n]):
            hps[i % n](*args)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    outs = [hp._out.clone() for hp in hps]
    same = all(torch.equal(outs[0], o) for o in outs)
    print(f"{n} depth maps in flight: {ms:.3f} ms per depth map = {1e3 / ms:.1f} depth-maps/s (outputs identical: {same})", flush=True)
    del hps
