"""Deterministic synthetic inputs for the CER-MVS hot path (no dataset, no checkpoint).

Everything here is generated from ``numpy.random.RandomState`` so the same arrays come out
in the build container (golden generation), in the CPU tests and on the GPU box.

What is synthesised (shapes follow SURVEY.md section 8, reference file:line in brackets):

* cameras: DTU-like intrinsics at full resolution and world->camera poses on a ring around the
  reference camera  [datasets/dtu.py:255-276 yields poses/intrinsics in this form];
* feature maps ``fmaps [1, V+1, C, h1, w1]``: what ``fnet`` would emit for a textured slanted
  plane -- every view samples the same analytic texture (random sinusoids per channel) at the
  3-D point its pixel sees, so the epipolar cost volume has a true peak  [core/raft.py:66-69];
* context ``net, inp [1, 1, 64, h1, w1]``: tanh / relu of smooth random fields  [core/raft.py:57-60];
* ``UpdateBlock`` weights with the reference's state-dict keys and shapes  [core/update.py:58-78].
"""
from __future__ import annotations

import numpy as np

DTU_FX, DTU_FY = 2892.33, 2883.18  # typical DTU rectified intrinsics at 1600x1200


def make_cameras(num_src: int, height: int, width: int, seed: int = 0,
                 depth_centre: float = 600.0, baseline: float = 120.0):
    """Returns (poses [V+1,4,4], intrinsics [V+1,3,3]) float32, full-resolution intrinsics.

    Pose 0 is the reference camera (identity); source cameras sit on a ring of radius
    ``baseline`` in the reference camera's x/y plane and are rotated to look at the point
    ``(0, 0, depth_centre)``.
    """
    rs = np.random.RandomState(seed)
    n = num_src + 1
    K = np.zeros((n, 3, 3), np.float64)
    K[:, 0, 0] = DTU_FX * width / 1600.0
    K[:, 1, 1] = DTU_FY * height / 1200.0
    K[:, 0, 2] = width / 2.0
    K[:, 1, 2] = height / 2.0
    K[:, 2, 2] = 1.0
    poses = np.zeros((n, 4, 4), np.float64)
    poses[0] = np.eye(4)
    target = np.array([0.0, 0.0, depth_centre])
    for v in range(1, n):
        ang = 2.0 * np.pi * (v - 1) / max(num_src, 1) + rs.uniform(-0.2, 0.2)
        rad = baseline * rs.uniform(0.8, 1.2)
        centre = np.array([rad * np.cos(ang), rad * np.sin(ang), rs.uniform(-10.0, 10.0)])
        z = target - centre
        z /= np.linalg.norm(z)
        up = np.array([0.0, 1.0, 0.0])
        x = np.cross(up, z)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z], 0)          # world -> camera rotation
        poses[v, :3, :3] = R
        poses[v, :3, 3] = -R @ centre
        poses[v, 3, 3] = 1.0
    return poses.astype(np.float32), K.astype(np.float32)


def _plane_points(pose, Kq, h1, w1, plane):
    """3-D world points where the quarter-res pixels of one camera hit the plane n.X = d."""
    nrm, d = plane
    R = pose[:3, :3].astype(np.float64)
    t = pose[:3, 3].astype(np.float64)
    centre = -R.T @ t
    ys, xs = np.meshgrid(np.arange(h1, dtype=np.float64), np.arange(w1, dtype=np.float64), indexing="ij")
    rays_c = np.stack([(xs - Kq[0, 2]) / Kq[0, 0], (ys - Kq[1, 2]) / Kq[1, 1], np.ones_like(xs)], -1)
    rays_w = rays_c @ R                      # R^T applied to each ray
    s = (d - centre @ nrm) / (rays_w @ nrm)
    return centre + rays_w * s[..., None]


def make_fmaps(poses, intrinsics, height, width, channels: int = 64, seed: int = 0,
               depth_centre: float = 600.0, fp16_exact: bool = True):
    """Feature maps [1, V+1, C, h1, w1] float32 (values exactly representable in fp16 when
    ``fp16_exact``) for a slanted textured plane, plus the true inverse depth of the reference
    view [h1, w1]."""
    rs = np.random.RandomState(seed + 1000)
    h1, w1 = height // 4, width // 4
    n = poses.shape[0]
    nrm = np.array([0.15, -0.1, 1.0])
    nrm /= np.linalg.norm(nrm)
    plane = (nrm, nrm[2] * depth_centre)
    # texture: per channel a sum of 3 plane waves over world (x, y); wavelengths 6..60 units
    freq = rs.uniform(2 * np.pi / 60.0, 2 * np.pi / 6.0, size=(channels, 3))
    ang = rs.uniform(0, 2 * np.pi, size=(channels, 3))
    phase = rs.uniform(0, 2 * np.pi, size=(channels, 3))
    kx, ky = freq * np.cos(ang), freq * np.sin(ang)
    fmaps = np.zeros((n, channels, h1, w1), np.float32)
    Kq = intrinsics.astype(np.float64).copy()
    Kq[:, :2] /= 4.0                          # core/raft.py:39
    for v in range(n):
        X = _plane_points(poses[v], Kq[v], h1, w1, plane)
        arg = X[..., 0, None, None] * kx + X[..., 1, None, None] * ky + phase
        fmaps[v] = np.sin(arg).sum(-1).transpose(2, 0, 1) * (1.0 / np.sqrt(1.5))
    if fp16_exact:
        fmaps = fmaps.astype(np.float16).astype(np.float32)
    X0 = _plane_points(poses[0], Kq[0], h1, w1, plane)
    true_disp = (1.0 / X0[..., 2]).astype(np.float32)
    return fmaps[None], true_disp


def _smooth_field(rs, channels, h1, w1, cell=6):
    """Band-limited random field: coarse gaussian noise, bilinearly upsampled."""
    gh, gw = h1 // cell + 2, w1 // cell + 2
    g = rs.standard_normal((channels, gh, gw))
    ys = np.linspace(0, gh - 1.001, h1)
    xs = np.linspace(0, gw - 1.001, w1)
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[None, :, None], (xs - x0)[None, None, :]
    a = g[:, y0][:, :, x0]
    b = g[:, y0][:, :, x0 + 1]
    c = g[:, y0 + 1][:, :, x0]
    d = g[:, y0 + 1][:, :, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx)


def make_context_pre(h1, w1, dim_net: int = 64, dim_inp: int = 64, seed: int = 0):
    """What ``cnet`` would emit before the tanh / relu split: [1, 1, dim_net+dim_inp, h1, w1]."""
    rs = np.random.RandomState(seed + 2000)
    pre = np.concatenate([_smooth_field(rs, dim_net, h1, w1), _smooth_field(rs, dim_inp, h1, w1)], 0)
    return np.ascontiguousarray(pre, dtype=np.float32)[None, None]


def make_context(h1, w1, dim_net: int = 64, dim_inp: int = 64, seed: int = 0, fp16_exact: bool = True):
    """(net, inp), each [1, 1, C, h1, w1] float32: tanh / relu of smooth fields (core/raft.py:57-60)."""
    pre = make_context_pre(h1, w1, dim_net, dim_inp, seed)[0, 0]
    net = np.ascontiguousarray(np.tanh(pre[:dim_net]), dtype=np.float32)
    inp = np.ascontiguousarray(np.maximum(pre[dim_net:], 0), dtype=np.float32)
    if fp16_exact:
        net = net.astype(np.float16).astype(np.float32)
        inp = inp.astype(np.float16).astype(np.float32)
    return net[None, None], inp[None, None]


def make_update_weights(seed: int = 0, delta_scale: float = 0.02, delta_bias: float = 0.0, num_levels: int = 3, radius: int = 5,
                        dim_net: int = 64, dim_inp: int = 64, size_disp_enc: int = 7, n_cascade: int = 2,
                        fp16_exact: bool = True):
    """State dict (numpy, OIHW float32) for the reference ``UpdateBlock`` (core/update.py:58-78),
    default sharing flags: corr_encoder and gru shared, delta per stage.

    Scale is PyTorch's default conv init (uniform +-1/sqrt(fan_in)); the last delta conv is
    multiplied by ``delta_scale`` so per-iteration updates stay inside the cost volume
    (synthetic checkpoint); ``delta_bias`` is added to the last delta bias so
    the disparity drifts upwards by about ``0.01*delta_bias`` per iteration (walks through the volume)."""
    rs = np.random.RandomState(seed + 3000)
    sd = {}

    def conv(name, cout, cin, k, scale=1.0):
        bound = 1.0 / np.sqrt(cin * k * k)
        sd[name + ".weight"] = (rs.uniform(-bound, bound, (cout, cin, k, k)) * scale).astype(np.float32)
        sd[name + ".bias"] = (rs.uniform(-bound, bound, (cout,)) * scale).astype(np.float32)

    cor_planes = num_levels * (2 * radius + 1)
    conv("corr_encoder.0", 64, cor_planes, 1)
    conv("corr_encoder.2", 64, 64, 3)
    for i in range(n_cascade):
        conv(f"delta{i}.0", 256, dim_net, 3)
        conv(f"delta{i}.2", 1, 256, 3, scale=delta_scale)
        sd[f"delta{i}.2.bias"] = sd[f"delta{i}.2.bias"] + np.float32(delta_bias)
    cin = dim_net + dim_inp + 64 + size_disp_enc ** 2
    for g in ("convz", "convr", "convq"):
        conv(f"gru.{g}", dim_net, cin, 3)
    if fp16_exact:
        sd = {k: v.astype(np.float16).astype(np.float32) for k, v in sd.items()}
    return sd


# BASELINE.json configs: name -> (H, W, V)
CONFIGS = {
    "cfg1_dtu_448x576_v2": (448, 576, 2),
    "cfg2_dtu_1184x1600_v10": (1184, 1600, 10),
    "cfg3_dtu_2368x3200_v10": (2368, 3200, 10),
    "cfg4_tnt_1056x1920_v15": (1056, 1920, 15),
    "cfg5_blended_1536x2048_v7": (1536, 2048, 7),
}


def make_scene(height, width, num_src, seed=0, fp16_exact=True):
    """All hot-path inputs for one reference image, as numpy arrays."""
    poses, K = make_cameras(num_src, height, width, seed)
    fmaps, true_disp = make_fmaps(poses, K, height, width, seed=seed, fp16_exact=fp16_exact)
    net, inp = make_context(height // 4, width // 4, seed=seed, fp16_exact=fp16_exact)
    return dict(fmaps=fmaps, net=net, inp=inp, poses=poses[None], intrinsics=K[None],
                true_disp=true_disp)


def make_encoder_weights(seed: int = 0, out_dim: int = 64, fp16_exact: bool = True):
    """State dict (numpy, OIHW float32) for the reference ``BasicEncoder`` of type "HR" (core/extractor.py:62-118):
    Kaiming-normal weights (fan_out, relu) like its own init, small random biases."""
    rs = np.random.RandomState(seed + 4000)
    sd = {}

    def conv(name, cout, cin, k):
        std = np.sqrt(2.0 / (cout * k * k))
        sd[name + ".weight"] = (rs.standard_normal((cout, cin, k, k)) * std).astype(np.float32)
        sd[name + ".bias"] = (rs.standard_normal((cout,)) * 0.05).astype(np.float32)

    conv("conv1", 32, 3, 7)
    for b in (0, 1):
        conv(f"layer1.{b}.conv1", 32, 32, 3)
        conv(f"layer1.{b}.conv2", 32, 32, 3)
    conv("layer2.0.conv1", 64, 32, 3)
    conv("layer2.0.conv2", 64, 64, 3)
    conv("layer2.0.downsample.0", 64, 32, 1)
    conv("layer2.1.conv1", 64, 64, 3)
    conv("layer2.1.conv2", 64, 64, 3)
    conv("conv2", out_dim, 64, 1)
    if fp16_exact:
        sd = {k: v.astype(np.float16).astype(np.float32) for k, v in sd.items()}
    return sd


def make_image(height, width, n=1, seed=0):
    """[n,3,H,W] float32 in 0..255: low-passed uniform noise (neighbouring pixels correlate like a photograph)."""
    rs = np.random.RandomState(seed + 5000)
    img = rs.uniform(0, 255, (n, 3, height, width))
    img = (img + np.roll(img, 1, 2) + np.roll(img, 1, 3) + np.roll(img, (1, 1), (2, 3))) / 4
    return np.ascontiguousarray(img, dtype=np.float32)
