#!/usr/bin/env python
"""HBM throughput of the SURVEY 8f kernels (image prep, depth output, multires merge, geometric filter) at the BASELINE
sizes: algorithmic bytes / CUDA-event time against the measured HBM peak.  Writes gpurun_out/aux_kernels.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cer_mvs_b200 import fusion_ops, prep  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


g = torch.Generator(device="cuda").manual_seed(0)
out = {}
V, H, W = 10, 1184, 1600
img = torch.rand(V + 1, 3, H, W, device="cuda", generator=g) * 255                      # 11 views, full resolution
K = torch.eye(3)[None].repeat(V + 1, 1, 1)
buf = torch.empty_like(img)
rows = [("normalize_images (11 x 3 x 1184 x 1600)", lambda: prep.normalize_images(img, out=buf), 2 * img.numel() * 4),
        ("scale_operation x2 (11 x 3 x 1184 x 1600 -> 2368 x 3200)", lambda: prep.scale_operation(img, K.clone(), 2), 5 * img.numel() * 4)]
disp = torch.rand(592, 800, device="cuda", generator=g) * 2e-3 + 1e-4
rows.append(("disp_to_depth + flip (592 x 800)", lambda: prep.disp_to_depth(disp, flip_rows=True), 2 * disp.numel() * 4))
d1 = torch.rand(296, 400, device="cuda", generator=g) * 400 + 400
d2 = torch.rand(592, 800, device="cuda", generator=g) * 400 + 400
rows.append(("multires_merge (296 x 400 + 592 x 800)", lambda: prep.multires_merge(d1, d2, 0.02), (d1.numel() + 2 * d2.numel()) * 4))
S, h, w = 10, 1184, 1600                                                                # fusion works on full-resolution maps
depths = torch.rand(S + 1, h, w, device="cuda", generator=g) * 400 + 400
Kf = torch.tensor([[2892.0, 0, w / 2], [0, 2883.0, h / 2], [0, 0, 1]], device="cuda")
E = torch.eye(4, device="cuda")
Es = E[None].repeat(S, 1, 1).clone()
Es[:, 0, 3] = torch.linspace(-100, 100, S, device="cuda")
rows.append((f"geometric_filter ({S} source views, {h} x {w})",
             lambda: fusion_ops.geometric_filter(depths[0], Kf, E, depths[1:], Kf[None].repeat(S, 1, 1), Es, 4.4, 1430.0),
             ((S + 1) * h * w * 4 + h * w * 5)))
for name, fn, nbytes in rows:
    t = timed(fn)
    out[name] = {"us": t * 1e6, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / peak}
    print(f"{name}: {t * 1e6:.1f} us, {nbytes / t / 1e9:.0f} GB/s = {100 * nbytes / t / 1e9 / peak:.0f} % of {peak:.0f}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "aux_kernels.json"), "w"), indent=1)
