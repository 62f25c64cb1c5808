"""``RAFT``: the reference's model class (core/raft.py:13-109) with the same constructor, the same parameter tree
(``fnet.*``, ``cnet.*``, ``update_block.*``: a reference checkpoint loads with ``strict=True``) and the same
``forward(images, poses, intrinsics, scale, do_report)`` contract in test mode -- running entirely on the kernels of
libcer_mvs_b200:

    fnet per image  (csrc/encoder.cu)  -> straight into the plan's feature buffers, NHWC fp16 pre-scaled by 1/8
    cnet + tanh / relu split            -> straight into the plan's net / inp buffers
    cascade stages                      -> DepthHotPath (cost-volume builds + CUDA-graph GRU loops, csrc/plan.cu)

No layout kernels run between the encoders and the hot path.  Inference only (``test_mode=True``), batch size 1
(inference.py:49), encoder type "HR".
"""
import torch
import torch.nn as nn

from . import _lib
from .extractor import BasicEncoder
from .hotpath import DepthHotPath
from .update import UpdateBlock


class _DevArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class RAFT(nn.Module):
    def __init__(self, cascade=((64, 64, 8), (-1, 320, 8)), encoder_type="HR", dim_fmap=64, dim_net=64, dim_inp=64,
                 test_mode=False):
        super().__init__()
        if encoder_type != "HR" or dim_fmap != 64 or dim_net != 64 or dim_inp != 64:
            raise NotImplementedError("cer_mvs_b200.RAFT: the reference's default dimensions / HR encoder only")
        self.cascade = [tuple(c) for c in cascade]
        self.encoder_type, self.dim_fmap, self.dim_net, self.dim_inp, self.test_mode = encoder_type, dim_fmap, dim_net, dim_inp, test_mode
        self.fnet = BasicEncoder(output_dim=dim_fmap, norm_fn="instance", type=encoder_type)
        self.cnet = BasicEncoder(output_dim=dim_net + dim_inp, norm_fn="none", type=encoder_type)
        self.update_block = UpdateBlock(cascade=self.cascade, dim_net=dim_net, dim_inp=dim_inp)
        self._plans = {}

    def _plan(self, h1, w1, n_views, device):
        key = (h1, w1, n_views, str(device))
        ub_key = tuple((k, p.data_ptr(), p._version) for k, p in self.update_block.state_dict(keep_vars=True).items())
        ent = self._plans.get(key)
        if ent is None:
            hp = DepthHotPath(h1, w1, max_views=n_views, cascade=self.cascade, feats_f16=True, device=device)
            ent = self._plans[key] = [hp, None]
        if ent[1] != ub_key:
            ent[0].load_update_block(self.update_block.state_dict())
            ent[1] = ub_key
        return ent[0]

    def forward(self, images, poses, intrinsics, scale=None, do_report=False):
        """images [1,V+1,3,H,W] float32 in 0..255 (left untouched: the normalisation of core/raft.py:40-41 is fused into
        the first convolution), poses [1,V+1,4,4], intrinsics [1,V+1,3,3] -> disp * scale [1,1,H/4,W/4] (float64 when
        ``scale`` is a float64 tensor, like core/raft.py:108)."""
        if not self.test_mode:
            raise NotImplementedError("cer_mvs_b200.RAFT: inference (test_mode=True) only")
        if images.dim() != 5 or images.shape[0] != 1 or images.shape[2] != 3:
            raise RuntimeError("images must be [1, V+1, 3, H, W] (inference.py:49)")
        if not images.is_cuda:
            raise RuntimeError("cer_mvs_b200.RAFT: CUDA tensors expected (no CPU path)")
        _, num, _, H, W = images.shape
        n_views, h1, w1, dev = num - 1, H // 4, W // 4, images.device
        hp = self._plan(h1, w1, n_views, dev)
        L = _lib.lib()
        s = 1.0 if scale is None else float(scale.reshape(-1)[0]) if isinstance(scale, torch.Tensor) else float(scale)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            px = h1 * w1
            feats = torch.as_tensor(_DevArray(L.cer_plan_feature_buffer(hp._plan, 0), (num, h1, w1, 64), "<f2"), device=dev)
            net = torch.as_tensor(_DevArray(L.cer_plan_net_buffer(hp._plan), (1, h1, w1, 64), "<f2"), device=dev)
            inp = torch.as_tensor(_DevArray(L.cer_plan_inp_buffer(hp._plan), (1, h1, w1, 64), "<f2"), device=dev)
            self.fnet.forward_features(images[0], scale=0.125, normalize=True, out=feats)          # core/raft.py:66-69
            self.cnet.forward_context(images[0, :1], normalize=True, out=(net, inp))              # core/raft.py:57-60
            P, K = hp._prep_cameras(poses, intrinsics, s if scale is not None else None)
            _lib.check(L.cer_plan_prepare_inplace(hp._plan, P.data_ptr(), K.data_ptr(), n_views, 0, n_views, st),
                       "cer_plan_prepare_inplace")
            for stg in range(len(hp.stages)):
                _lib.check(L.cer_plan_build_stage(hp._plan, stg, st), "cer_plan_build_stage")
                _lib.check(L.cer_plan_iterate_stage(hp._plan, stg, st), "cer_plan_iterate_stage")
            out = torch.empty(1, 1, h1, w1, device=dev, dtype=torch.float32)
            _lib.check(L.cer_plan_finish(hp._plan, s, out.data_ptr(), st), "cer_plan_finish")
        if isinstance(scale, torch.Tensor) and scale.dtype == torch.float64:
            return out.double()
        return out
