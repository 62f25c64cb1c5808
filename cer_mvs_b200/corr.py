"""Drop-in for ``core/corr.py``'s ``CorrBlock`` (reference :45-143): same constructor keywords,
same ``__call__``, same attributes -- computed by the fused CUDA kernels of libcer_mvs_b200.

Differences a caller can see (both allowed by the reference's own consumer, core/update.py:99-103):

* by default the cost volume is stored already averaged over the source views (``mean_v`` commutes
  with the pyramid and the lookup), so ``__call__`` returns ``[B, 1, 33, h, w]``
  instead of ``[B, V, 33, h, w]``; ``UpdateBlock`` takes ``mean(dim=1)`` either way.
  Pass ``per_view=True`` (or set ``CorrBlock.per_view_default``) for the reference's exact layout.
* ``corr_pyramid`` is materialised lazily (the lookup kernel rebuilds levels 1..L-1 on the fly).
"""
import weakref

import torch

from . import _lib

_feat_cache = {}       # id(fmaps) -> (weakref(fmaps), key, NHWC copy); an entry dies with the caller's tensor


def clear_feature_cache():
    """Drop every cached NHWC feature copy (call before ``torch.cuda.empty_cache()`` to return the memory)."""
    _feat_cache.clear()


def _prepare_features(fmaps: torch.Tensor):
    """[1,n,64,h,w] (fp16|fp32) -> NHWC [n,h,w,64] scaled by 1/8 (core/corr.py:29-35), cached for the
    second cascade stage (the reference redoes this per view per stage).  The cache holds only a weak
    reference to ``fmaps``: when the caller drops the feature maps (end of ``RAFT.forward``) the NHWC copy
    goes with them, so peak memory follows the reference's (inference.py empties the allocator cache per image)."""
    key = (fmaps.data_ptr(), fmaps._version, tuple(fmaps.shape), fmaps.dtype, fmaps.device)
    ident = id(fmaps)
    hit = _feat_cache.get(ident)
    if hit is not None and hit[0]() is fmaps and hit[1] == key:
        return hit[2]
    B, n, C, h, w = fmaps.shape
    f16 = fmaps.dtype == torch.float16
    src = fmaps if fmaps.is_contiguous() else fmaps.contiguous()
    dst = torch.empty(n, h, w, C, device=fmaps.device, dtype=torch.float16 if f16 else torch.float32)
    _lib.check(_lib.lib().cer_nchw_to_nhwc(src.data_ptr(), int(f16), dst.data_ptr(), int(f16), n, C, h, w, 0.125,
                                           _lib.stream_ptr()), "feature layout")
    _feat_cache[ident] = (weakref.ref(fmaps, lambda _r, ident=ident: _feat_cache.pop(ident, None)), key, dst)
    return dst


class CorrBlock:
    per_view_default = False

    def __init__(self, fmaps, poses, intrinsics, ii, jj, nIncre, incre, disps_input, shift, num_levels, radius,
                 test_mode, do_report, per_view=None):
        self.num_levels = num_levels
        self.radius = radius
        self.test_mode = test_mode
        self.nIncre = nIncre
        self.incre = incre
        self.per_view = CorrBlock.per_view_default if per_view is None else per_view
        if not fmaps.is_cuda:
            raise RuntimeError("fmaps must be a CUDA tensor (cer_mvs_b200 has no CPU path)")
        if fmaps.dtype not in (torch.float16, torch.float32):
            fmaps = fmaps.float()                                          # core/corr.py:53
        B, n, C, h, w = fmaps.shape
        if B != 1:
            raise NotImplementedError("cer_mvs_b200.CorrBlock: batch size 1 only (inference.py:49)")
        if C != 64:
            raise NotImplementedError("cer_mvs_b200.CorrBlock: dim_fmap must be 64 (core/raft.py:18)")
        if num_levels < 1 or num_levels > 3:
            raise NotImplementedError("cer_mvs_b200.CorrBlock: num_levels must be 1..3")
        L = _lib.lib()
        dev = fmaps.device
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            feats = _prepare_features(fmaps)
            V = int(ii.shape[0])
            self._V, self._h, self._w = V, h, w
            ii32 = ii.to(device=dev, dtype=torch.int32).contiguous()
            jj32 = jj.to(device=dev, dtype=torch.int32).contiguous()
            P = poses.reshape(-1, 4, 4).to(device=dev, dtype=torch.float32).contiguous()
            K = intrinsics.reshape(-1, 3, 3).to(device=dev, dtype=torch.float32).contiguous()
            Pij = torch.empty(V, 16, device=dev, dtype=torch.float32)
            _lib.check(L.cer_projection_matrices(P.data_ptr(), K.data_ptr(), ii32.data_ptr(), jj32.data_ptr(), V,
                                                 Pij.data_ptr(), st), "projection matrices")
            disp_in = disps_input.reshape(h, w).to(torch.float32).contiguous()
            origin = torch.empty(h, w, device=dev, dtype=torch.float32)
            slots = V if self.per_view else 1
            volume = torch.empty(slots, h * w, nIncre, device=dev, dtype=torch.float32)
            lo = nIncre // 2 * incre                                       # core/corr.py:60 (Python double)
            _lib.check(L.cer_build_volume(feats.data_ptr(), int(feats.dtype == torch.float16), Pij.data_ptr(),
                                          ii32.data_ptr(), jj32.data_ptr(), V, disp_in.data_ptr(), int(bool(shift)),
                                          int(nIncre), float(incre), float(torch.tensor(lo).float()),
                                          origin.data_ptr(), volume.data_ptr(),
                                          1.0 if self.per_view else 1.0 / V, int(self.per_view), h, w, st),
                       "cost-volume build")
        self.volume = volume
        self.Pij = Pij
        self.disps_origin = origin.view(1, 1, 1, h, w)
        self._pyramid = None

    @property
    def corr_pyramid(self):
        """List of [rows,1,1,W_l] like core/corr.py:94-97 (rows = V*h*w per_view, else h*w view means)."""
        if self._pyramid is None:
            L = _lib.lib()
            lv = [self.volume.reshape(-1, self.nIncre)]
            with torch.cuda.device(self.volume.device):
                for _ in range(self.num_levels - 1):
                    src = lv[-1]
                    dst = torch.empty(src.shape[0], src.shape[1] // 2, device=src.device, dtype=torch.float32)
                    _lib.check(L.cer_pool_pairs(src.data_ptr(), dst.data_ptr(), src.shape[0], src.shape[1],
                                                _lib.stream_ptr()), "pyramid")
                    lv.append(dst)
            self._pyramid = [x.view(x.shape[0], 1, 1, x.shape[1]) for x in lv]
        return self._pyramid

    def __call__(self, zinv):
        """zinv [B,num,h,w] -> [B,slots,L*(2r+1),h,w] fp32 (core/corr.py:102-143)."""
        batch, num, h1, w1 = zinv.shape
        if batch != 1 or h1 != self._h or w1 != self._w:
            raise RuntimeError("CorrBlock.__call__: zinv shape does not match the volume")
        if not zinv.is_cuda:
            raise RuntimeError("zinv must be a CUDA tensor")
        z = zinv.to(torch.float32)
        slots = self._V if self.per_view else 1
        if self.per_view:
            if num != self._V:
                raise RuntimeError("CorrBlock.__call__: per_view volume needs one zinv map per view")
            z = z.contiguous()
            stride = h1 * w1
        else:
            z = z[:, 0].contiguous()          # core/raft.py:99 passes V copies of the same disparity
            stride = 0
        planes = self.num_levels * (2 * self.radius + 1)
        out = torch.empty(1, slots, planes, h1, w1, device=z.device, dtype=torch.float32)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().cer_lookup_strided(self.volume.data_ptr(), slots, self.disps_origin.data_ptr(),
                                                     z.data_ptr(), stride, int(self.nIncre), float(self.incre),
                                                     int(self.radius), int(self.num_levels), out.data_ptr(), h1, w1,
                                                     _lib.stream_ptr()), "pyramid lookup")
        return out
