"""Drop-ins for ``core/update.py``'s ``ConvGRU`` (:9-25) and ``UpdateBlock`` (:29-120): same
constructors, same parameter names / shapes / state-dict keys (so a reference ``.pth`` loads with
``strict=True``), same ``forward`` signatures and return shapes.  ``forward`` runs the fused
tensor-core kernels of libcer_mvs_b200 with autocast numerics (fp16 operands, fp32 accumulate);
the ``nn.Conv2d`` children only hold the parameters and are never called.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib

_ORDER = ["corr_encoder.0", "corr_encoder.2", "gru.convz", "gru.convr", "gru.convq",
          "delta0.0", "delta0.2", "delta1.0", "delta1.2"]
_SHAPES = {"corr_encoder.0": (64, 33, 1, 1), "corr_encoder.2": (64, 64, 3, 3), "gru.convz": (64, 241, 3, 3),
           "gru.convr": (64, 241, 3, 3), "gru.convq": (64, 241, 3, 3), "delta0.0": (256, 64, 3, 3),
           "delta0.2": (1, 256, 3, 3), "delta1.0": (256, 64, 3, 3), "delta1.2": (1, 256, 3, 3)}


def pack_update_weights(sd) -> np.ndarray:
    """Host blob (uint8) from a state dict {name.weight/.bias: array-like}; missing entries are zero
    (used by the stand-alone ConvGRU).  Packing itself is native: cer_pack_update_weights."""
    L = _lib.lib()
    arrs = []
    for name in _ORDER:
        shape = _SHAPES[name]
        for suffix, shp in ((".weight", shape), (".bias", (shape[0],))):
            v = sd.get(name + suffix)
            if v is None:
                a = np.zeros(shp, np.float32)
            else:
                if isinstance(v, torch.Tensor):
                    v = v.detach().float().cpu().numpy()
                a = np.ascontiguousarray(v, dtype=np.float32)
                if a.shape != shp:
                    raise RuntimeError(f"{name}{suffix}: expected shape {shp}, got {a.shape} "
                                       "(only the reference's default UpdateBlock architecture is supported)")
            arrs.append(a)
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    blob = np.zeros(L.cer_update_blob_bytes(), np.uint8)
    _lib.check(L.cer_pack_update_weights(ptrs, blob.ctypes.data), "weight packing")
    return blob


class _PackedWeights:
    """Device blob cache, rebuilt when any parameter changes (checked through tensor versions)."""

    def __init__(self):
        self.key = None
        self.blob = None

    def get(self, module: nn.Module, prefix_map):
        params = [(k, p) for k, p in module.state_dict(keep_vars=True).items()]
        key = tuple((k, p.data_ptr(), p._version) for k, p in params)
        if key != self.key:
            sd = {prefix_map(k): p for k, p in params}
            host = pack_update_weights(sd)
            dev = params[0][1].device
            self.blob = torch.from_numpy(host).to(dev)
            self.key = key
        return self.blob


def _to_nhwc_f16(x, C_dst=64):
    """[1,C,h,w] (fp16|fp32) -> [h*w, C_dst] fp16."""
    _, Cc, h, w = x.shape
    x = x.contiguous()
    out = torch.empty(h * w, C_dst, device=x.device, dtype=torch.float16)
    _lib.check(_lib.lib().cer_nchw_to_nhwc_pad(x.data_ptr(), int(x.dtype == torch.float16), out.data_ptr(), 1, 1, Cc,
                                               C_dst, h, w, 1.0, _lib.stream_ptr()), "NCHW->NHWC")
    return out


def _to_nchw(x_nhwc, h, w, dtype):
    out = torch.empty(1, 64, h, w, device=x_nhwc.device, dtype=dtype)
    _lib.check(_lib.lib().cer_nhwc_to_nchw(x_nhwc.data_ptr(), 1, out.data_ptr(), int(dtype == torch.float16), 1, 64,
                                           h, w, _lib.stream_ptr()), "NHWC->NCHW")
    return out


def _check_act(name, t):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (cer_mvs_b200 has no CPU path)")
    if t.dtype not in (torch.float16, torch.float32):
        raise RuntimeError(f"{name}: expected float16 or float32, got {t.dtype}")


class ConvGRU(nn.Module):
    """core/update.py:9-25.  ``forward(net, inp, disp_enc, corr_enc)`` with NCHW tensors of
    64 / 64 / 49 / 64 channels (the only configuration the reference instantiates, update.py:73-78)."""

    def __init__(self, kernel_z=3, kernel_r=3, kernel_q=3, h_planes=None, i_planes=None):
        super().__init__()
        self.do_checkpoint = False
        self.convz = nn.Conv2d(h_planes + i_planes, h_planes, kernel_z, padding=kernel_z // 2)
        self.convr = nn.Conv2d(h_planes + i_planes, h_planes, kernel_r, padding=kernel_r // 2)
        self.convq = nn.Conv2d(h_planes + i_planes, h_planes, kernel_q, padding=kernel_q // 2)
        self._packed = _PackedWeights()
        self._ws = None

    def _workspace(self, h, w, device):
        need = _lib.lib().cer_update_workspace_bytes(h, w)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, device=device, dtype=torch.uint8)
        return self._ws

    def forward(self, net, *inputs):
        if len(inputs) != 3 or net.shape[1] != 64 or [t.shape[1] for t in inputs] != [64, 49, 64]:
            raise NotImplementedError("cer_mvs_b200.ConvGRU: inputs must be (inp[64], disp_enc[49], corr_enc[64]) "
                                      "with a 64-channel state (core/update.py:73-78)")
        for n_, t in (("net", net), ("inp", inputs[0]), ("disp", inputs[1]), ("corr", inputs[2])):
            _check_act(n_, t)
        if net.shape[0] != 1:
            raise NotImplementedError("cer_mvs_b200.ConvGRU: batch size 1 only")
        _, _, h, w = net.shape
        with torch.cuda.device(net.device):
            blob = self._packed.get(self, lambda k: "gru." + k)
            ws = self._workspace(h, w, net.device)
            n_ = _to_nhwc_f16(net)
            i_ = _to_nhwc_f16(inputs[0])
            d_ = _to_nhwc_f16(inputs[1], 64)
            e_ = _to_nhwc_f16(inputs[2])
            _lib.check(_lib.lib().cer_gru_step(blob.data_ptr(), ws.data_ptr(), n_.data_ptr(), i_.data_ptr(),
                                               d_.data_ptr(), e_.data_ptr(), h, w, _lib.stream_ptr()), "ConvGRU")
            return _to_nchw(n_, h, w, net.dtype)


class UpdateBlock(nn.Module):
    """core/update.py:29-120 with the reference's default architecture."""

    def __init__(self, kernel_corr=3, dim0_corr=64, dim1_corr=64, dim_net=None, dim_inp=None, dim0_delta=256,
                 kernel0_delta=3, kernel1_delta=3, num_levels=3, radius=5, size_disp_enc=7, kernel0_vis=3,
                 kernel1_vis=3, share_corr=True, share_gru=True, share_delta=False, aggregation=("mean",),
                 cascade=None):
        super().__init__()
        for k, v in dict(kernel_corr=kernel_corr, dim0_corr=dim0_corr, dim1_corr=dim1_corr, dim_net=dim_net,
                         dim_inp=dim_inp, dim0_delta=dim0_delta, kernel0_delta=kernel0_delta,
                         kernel1_delta=kernel1_delta, num_levels=num_levels, radius=radius,
                         size_disp_enc=size_disp_enc, kernel0_vis=kernel0_vis, kernel1_vis=kernel1_vis,
                         share_corr=share_corr, share_gru=share_gru, share_delta=share_delta,
                         aggregation=list(aggregation), cascade=cascade).items():
            setattr(self, k, v)                                           # store_attr(), update.py:55
        supported = (kernel_corr == 3 and dim0_corr == 64 and dim1_corr == 64 and dim_net == 64 and dim_inp == 64
                     and dim0_delta == 256 and kernel0_delta == 3 and kernel1_delta == 3 and num_levels == 3
                     and radius == 5 and size_disp_enc == 7 and share_corr and share_gru and not share_delta
                     and list(aggregation) == ["mean"] and cascade is not None and len(cascade) == 2)
        if not supported:
            raise NotImplementedError("cer_mvs_b200.UpdateBlock implements the reference's default architecture only "
                                      "(core/update.py:30-53 defaults, 2 cascade stages)")
        cor_planes = len(self.aggregation) * num_levels * (2 * radius + 1)
        self.corr_encoder = nn.Sequential(nn.Conv2d(cor_planes, dim0_corr, 1, padding=0), nn.ReLU(inplace=True),
                                          nn.Conv2d(dim0_corr, dim1_corr, kernel_corr, padding=kernel_corr // 2),
                                          nn.ReLU(inplace=True))
        for i in range(len(cascade)):
            setattr(self, f"delta{i}", nn.Sequential(
                nn.Conv2d(dim_net, dim0_delta, kernel0_delta, padding=kernel0_delta // 2), nn.ReLU(inplace=True),
                nn.Conv2d(dim0_delta, 1, kernel1_delta, padding=kernel1_delta // 2)))
        self.gru = ConvGRU(h_planes=dim_net, i_planes=dim_inp + dim1_corr + size_disp_enc ** 2)
        self._packed = _PackedWeights()
        self._ws = None
        self._inp_cache = None     # (key, source tensor, NHWC fp16)
        self._net_state = None     # (returned NCHW tensor, its version, NHWC fp16 state)

    def disp_encoder(self, disp):
        raise NotImplementedError("disp_encoder is fused into UpdateBlock.forward (csrc/update_hmma.cu, K0)")

    def forward(self, net, inp, disp, corr_frames, stage):
        for n_, t in (("net", net), ("inp", inp), ("disp", disp), ("corr_frames", corr_frames)):
            _check_act(n_, t)
        batch, num, ch, ht, wd = net.shape
        if batch != 1 or num != 1 or ch != 64:
            raise NotImplementedError("cer_mvs_b200.UpdateBlock: net must be [1,1,64,h,w]")
        if corr_frames.dtype != torch.float32 or corr_frames.shape[2] != 33 or corr_frames.shape[0] != 1:
            raise RuntimeError("corr_frames must be float32 [1,V,33,h,w] (CorrBlock.__call__ output)")
        stage = int(stage)
        L = _lib.lib()
        dev = net.device
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            blob = self._packed.get(self, lambda k: k)
            need = L.cer_update_workspace_bytes(ht, wd)
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(need, device=dev, dtype=torch.uint8)
            # inp is iteration-invariant: convert once per image
            ikey = (inp.data_ptr(), inp._version, tuple(inp.shape), inp.dtype)
            if self._inp_cache is None or self._inp_cache[0] != ikey:
                self._inp_cache = (ikey, inp, _to_nhwc_f16(inp.reshape(1, 64, ht, wd)))
            inp_nhwc = self._inp_cache[2]
            # GRU state: reuse our own NHWC copy when the caller hands back the tensor we returned
            st_ = self._net_state
            if st_ is not None and st_[0] is net and st_[1] == net._version:
                net_nhwc = st_[2]
            else:
                net_nhwc = _to_nhwc_f16(net.reshape(1, 64, ht, wd))
            d = disp.reshape(ht, wd).to(torch.float32).contiguous()
            corr = corr_frames.contiguous()
            delta = torch.empty(1, 1, ht, wd, device=dev, dtype=torch.float32)
            _lib.check(L.cer_update_step(blob.data_ptr(), self._ws.data_ptr(), net_nhwc.data_ptr(),
                                         inp_nhwc.data_ptr(), d.data_ptr(), corr.data_ptr(), corr.shape[1],
                                         delta.data_ptr(), 0, stage, ht, wd, st), "UpdateBlock.forward")
            net_out = _to_nchw(net_nhwc, ht, wd, net.dtype).view(1, 1, 64, ht, wd)
            self._net_state = (net_out, net_out._version, net_nhwc)
        return net_out, delta
