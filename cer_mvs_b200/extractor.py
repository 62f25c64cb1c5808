"""Drop-in for ``core/extractor.py``'s ``BasicEncoder`` (reference :62-155, type "HR"): same constructor, same parameter
names / shapes / state-dict keys (a reference ``.pth`` loads with ``strict=True``), same ``forward`` contract
(``[..., 3, H, W]`` normalised images in, ``[..., output_dim, H/4, W/4]`` out).  ``forward`` runs the mma.sync
implicit-GEMM kernels of csrc/encoder.cu with the reference's autocast numerics (fp16 operands and activations, fp32
accumulation and instance-norm statistics); the ``nn.Conv2d`` children only hold the parameters.

Beyond the reference surface, ``forward_features`` hands the result over in the cost-volume build's own layout (NHWC
fp16, pre-scaled by 1/8, core/corr.py:30-31) and ``forward_context`` returns ``net = tanh``, ``inp = relu``
(core/raft.py:58-60) in the update block's layout, so no separate layout kernels run between the encoders and the hot path.

Supported: ``norm_fn`` in {"instance", "none"} (fnet / cnet of core/raft.py:28-29), ``type="HR"``, no dropout, no multidim.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib

_ORDER = ["conv1", "layer1.0.conv1", "layer1.0.conv2", "layer1.1.conv1", "layer1.1.conv2", "layer2.0.conv1",
          "layer2.0.conv2", "layer2.0.downsample.0", "layer2.1.conv1", "layer2.1.conv2", "conv2"]


def pack_encoder_weights(sd, out_dim) -> np.ndarray:
    """Host blob (uint8) from a BasicEncoder state dict; packing itself is native (cer_pack_encoder_weights)."""
    L = _lib.lib()
    arrs = []
    for name in _ORDER:
        for suffix in (".weight", ".bias"):
            v = sd[name + suffix]
            if isinstance(v, torch.Tensor):
                v = v.detach().float().cpu().numpy()
            arrs.append(np.ascontiguousarray(v, dtype=np.float32))
    want = [(32, 3, 7, 7)] + [(32, 32, 3, 3)] * 4 + [(64, 32, 3, 3), (64, 64, 3, 3), (64, 32, 1, 1), (64, 64, 3, 3),
                                                   (64, 64, 3, 3), (out_dim, 64, 1, 1)]
    for name, a, shp in zip(_ORDER, arrs[0::2], want):
        if a.shape != shp:
            raise RuntimeError(f"{name}.weight: expected shape {shp}, got {a.shape} (only the 'HR' encoder is supported)")
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    blob = np.zeros(L.cer_encoder_blob_bytes(out_dim), np.uint8)
    _lib.check(L.cer_pack_encoder_weights(ptrs, out_dim, blob.ctypes.data), "encoder weight packing")
    return blob


class _Block(nn.Module):
    """Parameter holder with the attribute names of the reference's ResidualBlock (extractor.py:9-47)."""

    def __init__(self, in_planes, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1)
        if stride != 1:
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride))


class BasicEncoder(nn.Module):
    def __init__(self, output_dim=128, norm_fn="batch", dropout=0.0, multidim=False, type="HR"):
        super().__init__()
        if norm_fn not in ("instance", "none"):
            raise NotImplementedError("cer_mvs_b200.BasicEncoder: norm_fn must be 'instance' (fnet) or 'none' (cnet)")
        if type != "HR" or multidim or dropout > 0:
            raise NotImplementedError("cer_mvs_b200.BasicEncoder: only the reference's default HR encoder is implemented")
        if (norm_fn == "instance") != (output_dim == 64) or output_dim not in (64, 128):
            raise NotImplementedError("cer_mvs_b200.BasicEncoder: fnet (64, instance) or cnet (128, none) (core/raft.py:28-29)")
        self.norm_fn, self.multidim, self.type, self.output_dim = norm_fn, multidim, type, output_dim
        DIM = 32
        self.conv1 = nn.Conv2d(3, DIM, kernel_size=7, stride=2, padding=3)
        self.layer1 = nn.Sequential(_Block(DIM, DIM, 1), _Block(DIM, DIM, 1))
        self.layer2 = nn.Sequential(_Block(DIM, 2 * DIM, 2), _Block(2 * DIM, 2 * DIM, 1))
        self.conv2 = nn.Conv2d(2 * DIM, output_dim, kernel_size=1)
        for m in self.modules():                              # extractor.py:111-118
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self._key, self._blob, self._ws, self._ws_key = None, None, None, None

    # ---- packed weights / workspace, rebuilt when a parameter or the image size changes ----
    def _packed(self, device):
        params = list(self.state_dict(keep_vars=True).items())
        key = (str(device),) + tuple((k, p.data_ptr(), p._version) for k, p in params)
        if key != self._key:
            host = pack_encoder_weights({k: p for k, p in params}, self.output_dim)
            self._blob = torch.from_numpy(host).to(device)
            self._key = key
        return self._blob

    def _workspace(self, H, W, device):
        key = (H, W, str(device))
        if key != self._ws_key:
            self._ws = torch.empty(_lib.lib().cer_encoder_workspace_bytes(H, W), dtype=torch.uint8, device=device)
            self._ws_key = key
        return self._ws

    def _run(self, x, nhwc, nhwc2, nchw, scale, normalize=False, split=False):
        if not x.is_cuda:
            raise RuntimeError("BasicEncoder: CUDA tensor expected (cer_mvs_b200 has no CPU path)")
        if x.shape[-3] != 3 or x.shape[-1] % 4 or x.shape[-2] % 4:
            raise RuntimeError("BasicEncoder: [..., 3, H, W] with H, W multiples of 4 expected (core/raft.py:49-50)")
        H, W = x.shape[-2:]
        imgs = x.reshape(-1, 3, H, W).float().contiguous()
        L = _lib.lib()
        with torch.cuda.device(x.device):
            blob, ws = self._packed(x.device), self._workspace(H, W, x.device)
            for i in range(imgs.shape[0]):
                _lib.check(L.cer_encoder_forward(
                    blob.data_ptr(), ws.data_ptr(), imgs[i].data_ptr(), H, W, int(normalize), self.output_dim,
                    int(self.norm_fn == "instance"), int(split), nhwc[i].data_ptr() if nhwc is not None else None,
                    nhwc2[i].data_ptr() if nhwc2 is not None else None, nchw[i].data_ptr() if nchw is not None else None,
                    float(scale), _lib.stream_ptr()), "cer_encoder_forward")

    def forward(self, x):
        """[..., 3, H, W] -> [..., output_dim, H/4, W/4] fp16 (what the reference module returns under autocast)."""
        lead, (H, W) = x.shape[:-3], x.shape[-2:]
        n = int(np.prod(lead)) if len(lead) else 1
        out = torch.empty(n, self.output_dim, H // 4, W // 4, device=x.device, dtype=torch.float16)
        self._run(x, None, None, out, 1.0)          # cnet: the raw 128-channel map, split by the caller (core/raft.py:58-60)
        return out.reshape(*lead, self.output_dim, H // 4, W // 4)

    def forward_features(self, x, scale=0.125, normalize=False, out=None):
        """fnet straight into the cost-volume build's layout: [n, H/4, W/4, 64] fp16, multiplied by ``scale``."""
        if self.norm_fn != "instance":
            raise RuntimeError("forward_features is the fnet path")
        H, W = x.shape[-2:]
        n = x.numel() // (3 * H * W)
        out = torch.empty(n, H // 4, W // 4, 64, device=x.device, dtype=torch.float16) if out is None else out
        self._run(x, out, None, None, scale, normalize)
        return out

    def forward_context(self, x, normalize=False, nchw=False, out=None):
        """cnet + the split of core/raft.py:58-60: (net = tanh(.), inp = relu(.)), NHWC [n, H/4, W/4, 64] fp16 each
        (or NCHW [n, 64, H/4, W/4] when ``nchw``); ``out`` = (net, inp) NHWC tensors to fill."""
        if self.norm_fn != "none":
            raise RuntimeError("forward_context is the cnet path")
        H, W = x.shape[-2:]
        n = x.numel() // (3 * H * W)
        if out is not None:
            self._run(x, out[0], out[1], None, 1.0, normalize, split=True)
            return out
        if nchw:
            both = torch.empty(n, 2, 64, H // 4, W // 4, device=x.device, dtype=torch.float16)
            self._run(x, None, None, both, 1.0, normalize, split=True)
            return both[:, 0], both[:, 1]
        net = torch.empty(n, H // 4, W // 4, 64, device=x.device, dtype=torch.float16)
        inp = torch.empty_like(net)
        self._run(x, net, inp, None, 1.0, normalize, split=True)
        return net, inp
