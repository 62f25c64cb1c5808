"""ORACLE -- CPU restatement of the CER-MVS inference hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this file; the product (``cer_mvs_b200``) never does and has no CPU
fallback.

Each function restates one reference function in plain fp32 torch-on-CPU and cites the
reference ``file:line`` (relative to princeton-vl/CER-MVS @ 8062ddf) it follows:

    corr_forward           alt_cuda_corr/correlation_kernel.cu:18-119, 260-286
    projection_matrices    utils/projective_ops.py:16-23
    project_hypotheses     utils/projective_ops.py:5-13, 25-27 ; core/corr.py:87-88
    build_volume           core/corr.py:28-43, 46-97
    lookup                 core/corr.py:102-143 ; utils/bilinear_sampler.py:6-25
    disp_encoder           core/update.py:80-85
    update_block           core/update.py:17-25, 87-120
    hot_path               core/raft.py:75-108

How it is pinned (parity is NOT unpinned, but the pin is generated, the reference ships no tests):
  * ``tests/golden/*.npz`` hold outputs of the reference's own Python (``core/corr.py``,
    ``core/update.py``, ``utils/*`` imported from /root/reference through three import shims,
    script ``oracle/gen_golden.py``); ``tests/test_oracle_golden.py`` checks this file against them.
  * ``corr_forward`` restates a CUDA-only kernel; it is checked on the GPU box against
    ``oracle/_ref/alt_cuda_corr_ref*.so`` = the reference's two source files compiled where they lie
    (``oracle/build_ref.py``), see ``tests/test_gpu_corr_forward.py``.

Third-party arithmetic the reference relies on (torch, pinned 1.7.1 in environment.yml:151; run
here with torch 2.11): ``F.grid_sample`` (bilinear, zeros, align_corners=True), ``F.avg_pool2d``
(floor), ``F.unfold`` (zero pad), ``nn.Conv2d``.  The first three are restated explicitly below;
convolutions call ``F.conv2d`` in fp32.

``autocast=True`` emulates what ``torch.cuda.amp.autocast`` does to the reference on a GPU
(core/raft.py:55, 97-100): conv inputs/weights/bias rounded to fp16, fp32 accumulation, conv outputs
rounded to fp16, element-wise ops on fp16 tensors rounded to fp16 after every op.  ``autocast=False``
is the reference's CPU behaviour (autocast inert): everything fp32 (BASELINE.json configs[0]).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CHANNEL_STRIDE = 32  # correlation_kernel.cu:10


def _h(x: torch.Tensor) -> torch.Tensor:
    """Round to fp16 and come back to fp32 (one autocast rounding point)."""
    return x.half().float()


# ----------------------------------------------------------------------------------------------
# alt_cuda_corr.forward
# ----------------------------------------------------------------------------------------------
def corr_forward(fmap1: torch.Tensor, fmap2: torch.Tensor, coords: torch.Tensor, radius: int):
    """correlation_kernel.cu:18-119.  fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2]
    (x, y) in fmap2 pixels  ->  [corr [B,N,(2r+1)^2,H1,W1]] fp32, zero-initialised (:273).

    For every window offset (iy, ix) in [0, 2r+1]^2 the kernel dots fmap1 with fmap2 at integer
    pixel (floor(y)-r+iy, floor(x)-r+ix) (zero outside fmap2, :81-84) and scatters the result
    with the four bilinear weights into up to four outputs (:92-114); output index is
    ``iy + rd*ix`` (x-major).  Channels go in chunks of 32 (:43), accumulated with ``+=``."""
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    N = coords.shape[1]
    r = int(radius)
    rd = 2 * r + 1
    corr = torch.zeros(B, N, rd * rd, H1, W1, dtype=torch.float32)
    x = coords[..., 0].float()
    y = coords[..., 1].float()
    fx, fy = torch.floor(x), torch.floor(y)
    dx, dy = x - fx, y - fy                                   # :67-68
    # NaN/inf coords: the kernel's int cast is undefined; weights are NaN so the output is NaN.
    fxi = torch.nan_to_num(fx, nan=0.0, posinf=1e9, neginf=-1e9).clamp(-2**30, 2**30).long()
    fyi = torch.nan_to_num(fy, nan=0.0, posinf=1e9, neginf=-1e9).clamp(-2**30, 2**30).long()
    f2flat = fmap2.float().reshape(B * H2 * W2, C)
    f1 = fmap1.float().reshape(B, 1, H1, W1, C)
    boff = (torch.arange(B) * (H2 * W2)).view(B, 1, 1, 1)
    for c0 in range(0, C, CHANNEL_STRIDE):                    # :43
        f1c = f1[..., c0:c0 + CHANNEL_STRIDE].contiguous()
        f2c = f2flat[:, c0:c0 + CHANNEL_STRIDE].contiguous()
        for iy in range(rd + 1):                              # :71
            for ix in range(rd + 1):                          # :72
                h2 = fyi - r + iy
                w2 = fxi - r + ix
                inb = (h2 >= 0) & (h2 < H2) & (w2 >= 0) & (w2 < W2)
                lin = (h2.clamp(0, H2 - 1) * W2 + w2.clamp(0, W2 - 1) + boff).reshape(-1)
                g = f2c.index_select(0, lin).reshape(B, N, H1, W1, -1) * inb[..., None]
                s = (f1c * g).sum(-1)                         # :88-90
                if iy > 0 and ix > 0:                         # nw, :97,104
                    corr[:, :, (iy - 1) + rd * (ix - 1)] += s * dy * dx
                if iy > 0 and ix < rd:                        # ne, :98,107
                    corr[:, :, (iy - 1) + rd * ix] += s * dy * (1 - dx)
                if iy < rd and ix > 0:                        # sw, :99,110
                    corr[:, :, iy + rd * (ix - 1)] += s * (1 - dy) * dx
                if iy < rd and ix < rd:                       # se, :100,113
                    corr[:, :, iy + rd * ix] += s * (1 - dy) * (1 - dx)
    return [corr]


# ----------------------------------------------------------------------------------------------
# projective geometry
# ----------------------------------------------------------------------------------------------
def projection_matrices(poses: torch.Tensor, intrinsics: torch.Tensor, ii, jj) -> torch.Tensor:
    """utils/projective_ops.py:16-23.  poses [B,n,4,4] world->camera, intrinsics [B,n,3,3] already
    divided by the encoder stride (core/raft.py:39).  Returns Pij [B,len(ii),4,4] fp32:
    ``K4_j . P_j . P_i^-1 . K4_i^-1``."""
    Ks = torch.zeros_like(poses)
    Ks[..., :3, :3] = intrinsics
    Ks[..., 3, 3] = 1.0
    return Ks[:, jj] @ poses[:, jj] @ poses[:, ii].inverse() @ Ks[:, ii].inverse()


def project_hypotheses(Pij: torch.Tensor, disps: torch.Tensor) -> torch.Tensor:
    """utils/projective_ops.py:5-13,25-27 + core/corr.py:87-88.  Pij [B,1,4,4], disps [B,1,D,h,w]
    -> coords [B,1,h,w,D,2] = clamp((X0/X2, X1/X2), +-1e4) with X = Pij.(x, y, 1, d)."""
    B, _, D, h, w = disps.shape
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    x = xs.float() + torch.zeros_like(disps)
    y = ys.float() + torch.zeros_like(disps)
    x0 = torch.stack([x, y, torch.ones_like(disps), disps], -1)          # [B,1,D,h,w,4]
    x1 = torch.einsum("ijkh,ij...h->ij...k", Pij, x0)
    x1 = x1 / x1[..., [2]]
    x1 = x1[..., [0, 1]].permute(0, 1, 3, 4, 2, 5).contiguous()
    return x1.clamp(min=-1e4, max=1e4)


# ----------------------------------------------------------------------------------------------
# CorrBlock
# ----------------------------------------------------------------------------------------------
def hypothesis_origin(disps_input: torch.Tensor, nIncre: int, incre: float, shift: bool) -> torch.Tensor:
    """core/corr.py:56-63.  disps_input [B,1,h,w] -> origin [B,1,1,h,w]."""
    B, _, h, w = disps_input.shape
    d = disps_input.reshape(B, 1, 1, h, w).float()
    if shift:
        lo = torch.tensor(nIncre // 2 * incre).float()
        return torch.where(d < nIncre // 2 * incre, lo, d)
    return d.clone()


def build_volume(fmaps, poses, intrinsics, ii, jj, nIncre, incre, disps_input, shift, num_levels=3):
    """core/corr.py:46-97 (test_mode branch :79-91).  fmaps [B,n,C,h,w]; returns
    (pyramid, origin): pyramid[l] is [B*V*h*w, 1, 1, D>>l] fp32 (row index ((b*V+v)*h+y)*w+x,
    hypothesis minor), origin [B,1,1,h,w]."""
    fmaps = fmaps.float()
    B, n, C, h, w = fmaps.shape
    origin = hypothesis_origin(disps_input, nIncre, incre, shift)
    disps = ((torch.arange(nIncre) - nIncre // 2) * incre).view(1, 1, nIncre, 1, 1) + origin   # :56,66
    parts = []
    for k in range(len(ii)):
        i, j = int(ii[k]), int(jj[k])
        Pij = projection_matrices(poses, intrinsics, [i], [j])
        x1 = project_hypotheses(Pij, disps)                                  # [B,1,h,w,D,2]
        f = fmaps.permute(0, 1, 3, 4, 2)                                     # :29
        f1 = (f[:, [i]] / 8.0).reshape(B, h, w, C).contiguous()              # :30,34
        f2 = (f[:, [j]] / 8.0).reshape(B, h, w, C).contiguous()              # :31,35
        xc = x1.reshape(B, h, w, -1, 2).permute(0, 3, 1, 2, 4).contiguous()  # :37-38
        corr, = corr_forward(f1, f2, xc, 0)                                  # [B,D,1,h,w]
        corr = corr.permute(0, 2, 3, 4, 1)                                   # :41
        parts.append(corr.reshape(B, 1, h * w, 1, 1, nIncre))
    corr = torch.cat(parts, dim=1).reshape(-1, 1, 1, nIncre)                 # :90-91
    pyramid = [corr]
    for _ in range(num_levels - 1):                                          # :95-97
        W = corr.shape[-1]
        corr = 0.5 * (corr[..., 0:2 * (W // 2):2] + corr[..., 1:2 * (W // 2):2])   # avg_pool2d([1,2]), floor
        pyramid.append(corr)
    return pyramid, origin


def lookup(pyramid, origin, nIncre, incre, zinv, radius=5):
    """core/corr.py:102-143 (test_mode branch :123-139) with bilinear_sampler1
    (utils/bilinear_sampler.py:6-25) written out.  zinv [B,V,h,w] -> [B,V,L*(2r+1),h,w] fp32.

    grid_sample(bilinear, zeros, align_corners=True) on a 1 x W_l row: the reference normalises
    ``xn = 2x/(W_l-1) - 1`` and grid_sample un-normalises ``x' = (xn+1)/2*(W_l-1)``; both steps
    are kept so the fp32 rounding of x' matches."""
    r = radius
    B, V, h, w = zinv.shape
    z = zinv.reshape(B * V, h, w, 1).float()
    o = origin.reshape(B, h, w, 1).repeat(V, 1, 1, 1)
    coords = torch.maximum((z - o) / incre + nIncre // 2, torch.zeros(1))    # :107
    out = []
    for lvl, corr in enumerate(pyramid):
        Wl = corr.shape[-1]
        row = corr.reshape(B * V * h * w, Wl)
        cols = []
        for j in range(-r, r + 1):
            x0 = float(j) + coords.reshape(-1) / 2 ** lvl                    # :127-129
            xn = 2 * x0 / (Wl - 1) - 1                                       # bilinear_sampler.py:12
            xp = ((xn + 1) / 2) * (Wl - 1)                                   # grid_sampler unnormalize
            x_lo = torch.floor(xp)
            wgt = xp - x_lo
            i0 = x_lo.long()
            i1 = i0 + 1
            v0 = torch.gather(row, 1, i0.clamp(0, Wl - 1)[:, None])[:, 0] * ((i0 >= 0) & (i0 < Wl))
            v1 = torch.gather(row, 1, i1.clamp(0, Wl - 1)[:, None])[:, 0] * ((i1 >= 0) & (i1 < Wl))
            cols.append(v0 * (1 - wgt) + v1 * wgt)
        out.append(torch.stack(cols, -1).reshape(B * V, h, w, 2 * r + 1))
    out = torch.cat(out, dim=-1).permute(0, 3, 1, 2)                         # :142
    return out.reshape(B, V, -1, h, w).contiguous()


# ----------------------------------------------------------------------------------------------
# UpdateBlock
# ----------------------------------------------------------------------------------------------
def disp_encoder(disp: torch.Tensor, k: int = 7) -> torch.Tensor:
    """core/update.py:80-85.  disp [B,1,h,w] -> [B,k*k,h,w]: zero-padded k x k neighbourhood
    (row-major, ky outer) minus the centre value."""
    B, _, h, w = disp.shape
    p = k // 2
    pad = F.pad(disp, (p, p, p, p))
    chans = [pad[:, 0, ky:ky + h, kx:kx + w] for ky in range(k) for kx in range(k)]
    return torch.stack(chans, 1) - disp


def _conv(x, sd, name, autocast):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    pad = w.shape[-1] // 2
    if autocast:
        return _h(F.conv2d(_h(x), _h(w), _h(b), padding=pad))
    return F.conv2d(x, w, b, padding=pad)


def update_block(sd, net, inp, disp, corr_frames, stage, autocast=False):
    """core/update.py:87-120 with ConvGRU (:17-25), aggregation ["mean"], share_corr/share_gru
    True, share_delta False.  net, inp [B,1,64,h,w]; disp [B,1,h,w]; corr_frames [B,V,33,h,w].
    Returns (net [B,1,64,h,w], delta [B,1,h,w])."""
    B, _, ch, h, w = net.shape
    net = net.reshape(B, -1, h, w).float()
    inp = inp.reshape(B, -1, h, w).float()
    dn = 100 * disp_encoder(disp.reshape(B, 1, h, w).float())                # :97
    corr = corr_frames.float().mean(dim=1)                                   # :103
    e = torch.relu(_conv(corr, sd, "corr_encoder.0", autocast))              # :62-63
    e = torch.relu(_conv(e, sd, "corr_encoder.2", autocast))                 # :64-65
    x = torch.cat([net, inp, dn, e], dim=1)                                  # update.py:18-19
    if autocast:
        z = _h(torch.sigmoid(_conv(x, sd, "gru.convz", True)))
        r = _h(torch.sigmoid(_conv(x, sd, "gru.convr", True)))
        rn = _h(r * net)
        q = _h(torch.tanh(_conv(torch.cat([rn, inp, dn, e], dim=1), sd, "gru.convq", True)))
        net = _h(_h(_h(1 - z) * net) + _h(z * q))                            # :24, one rounding per op
        d = torch.relu(_conv(net, sd, f"delta{stage}.0", True))
        delta = _h(0.01 * _conv(d, sd, f"delta{stage}.2", True))             # :114
    else:
        z = torch.sigmoid(_conv(x, sd, "gru.convz", False))
        r = torch.sigmoid(_conv(x, sd, "gru.convr", False))
        q = torch.tanh(_conv(torch.cat([r * net, inp, dn, e], dim=1), sd, "gru.convq", False))
        net = (1 - z) * net + z * q
        d = torch.relu(_conv(net, sd, f"delta{stage}.0", False))
        delta = 0.01 * _conv(d, sd, f"delta{stage}.2", False)
    return net.reshape(B, 1, ch, h, w), delta.reshape(B, 1, h, w)


# ----------------------------------------------------------------------------------------------
# the stage / iteration loop
# ----------------------------------------------------------------------------------------------
def stage_params(cascade, num_levels=3, radius=5):
    """core/raft.py:76-81: (D, incre) per stage."""
    out = []
    for nIncre, incre, nIters in cascade:
        if nIncre == -1:
            nIncre = (2 * radius + 1) * 2 ** (num_levels - 1)
        out.append((nIncre, 0.0025 / incre, nIters))
    return out


def hot_path(sd, fmaps, net, inp, poses, intrinsics, cascade=((64, 64, 8), (-1, 320, 8)),
             scale=None, autocast=False, return_all=False, num_levels=3, radius=5):
    """core/raft.py:35-39 (pose/intrinsics prep), :75-108 (stage + iteration loop), test_mode.

    fmaps [B,V+1,C,h1,w1], net/inp [B,1,64,h1,w1], poses [B,V+1,4,4], intrinsics [B,V+1,3,3] at
    full image resolution.  Returns disp*scale [B,1,h1,w1] (and the per-iteration disparities)."""
    poses = poses.clone().float()
    if scale is not None:
        poses[..., :3, 3] *= float(scale)                                    # :35
    intrinsics = intrinsics.clone().float()
    intrinsics[:, :, :2] /= 4                                                # :39 (HR encoder)
    B, n, C, h, w = fmaps.shape
    ii = [0] * (n - 1)
    jj = list(range(1, n))
    disp = torch.zeros(B, 1, h, w)
    if autocast:
        fmaps, net, inp = _h(fmaps), _h(net), _h(inp)
    trace = []
    for stage, (nIncre, incre, nIters) in enumerate(stage_params(cascade, num_levels, radius)):
        pyramid, origin = build_volume(fmaps, poses, intrinsics, ii, jj, nIncre, incre,
                                       disp, stage == 0, num_levels)
        for _ in range(nIters):
            corr_frames = lookup(pyramid, origin, nIncre, incre, disp[:, ii], radius)   # :99
            net, delta = update_block(sd, net, inp, disp, corr_frames, stage, autocast)
            disp = disp + delta.float()                                      # :101
            trace.append(disp.clone())
    out = disp * (1.0 if scale is None else float(scale))                    # :108
    return (out, trace) if return_all else out


def to_torch_sd(sd_np):
    return {k: torch.from_numpy(v).float() for k, v in sd_np.items()}


# ----------------------------------------------------------------------------------------------
# BasicEncoder, type "HR" (core/extractor.py:62-155), SURVEY.md section 8f row 1
# ----------------------------------------------------------------------------------------------
def _enc_conv(x, sd, name, stride, autocast):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    pad = w.shape[-1] // 2
    if autocast:
        return _h(F.conv2d(_h(x), _h(w), _h(b), stride=stride, padding=pad))
    return F.conv2d(x, w, b, stride=stride, padding=pad)


def _enc_norm(x, instance, autocast):
    """nn.InstanceNorm2d defaults (extractor.py:29,74): no affine, no running stats, eps 1e-5, biased variance."""
    if not instance:
        return x
    y = F.instance_norm(x, eps=1e-5)
    return _h(y) if autocast else y


def _enc_block(x, sd, prefix, stride, instance, autocast):
    """ResidualBlock.forward (extractor.py:50-58)."""
    rnd = _h if autocast else (lambda v: v)
    y = torch.relu(_enc_norm(_enc_conv(x, sd, prefix + ".conv1", stride, autocast), instance, autocast))
    y = torch.relu(_enc_norm(_enc_conv(y, sd, prefix + ".conv2", 1, autocast), instance, autocast))
    if stride != 1:
        x = _enc_norm(_enc_conv(x, sd, prefix + ".downsample.0", stride, autocast), instance, autocast)
    return torch.relu(rnd(x + y))


def basic_encoder(sd, x, instance_norm, autocast=False):
    """x [N,3,H,W] (already normalised, core/raft.py:40-41) -> [N,out_dim,H/4,W/4]  (extractor.py:143-155)."""
    x = torch.relu(_enc_norm(_enc_conv(x, sd, "conv1", 2, autocast), instance_norm, autocast))
    x = _enc_block(x, sd, "layer1.0", 1, instance_norm, autocast)
    x = _enc_block(x, sd, "layer1.1", 1, instance_norm, autocast)
    x = _enc_block(x, sd, "layer2.0", 2, instance_norm, autocast)
    x = _enc_block(x, sd, "layer2.1", 1, instance_norm, autocast)
    return _enc_conv(x, sd, "conv2", 1, autocast)


def context_split(net_inp, autocast=False):
    """core/raft.py:58-60."""
    rnd = _h if autocast else (lambda v: v)
    return rnd(torch.tanh(net_inp[:, :64])), torch.relu(net_inp[:, 64:])
