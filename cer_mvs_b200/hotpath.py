"""``DepthHotPath``: the stage / iteration loop of ``RAFT.forward`` (core/raft.py:75-108) behind one
native plan (``cer_plan`` in csrc/plan.cu): feature layout, projection matrices, fused cost-volume
build, and per stage a CUDA-graph replay of {lookup, UpdateBlock} x iterations, everything resident
on the device in the kernels' layouts.

    hp = DepthHotPath(h1, w1, max_views=10, cascade=[(64, 64, 8), (-1, 320, 8)])
    hp.load_update_block(state_dict)                   # reference UpdateBlock keys (core/update.py)
    disp = hp(fmaps, net, inp, poses, intrinsics, scale)      # device tensors -> [1,1,h1,w1]
    disp = hp.run_host(fmaps_np, net_np, inp_np, poses_np, intrinsics_np, scale)   # host buffers

Multi-GPU: ``forward_sharded`` lets every rank build the partial volume of its own run of
(view, hypothesis) units, sums the partials with one all-reduce per stage, then iterates replicated.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .update import pack_update_weights


def stage_params(cascade, num_levels=3, radius=5):
    """core/raft.py:76-81 -> [(D, incre, iters)]."""
    out = []
    for nIncre, incre, nIters in cascade:
        if nIncre == -1:
            nIncre = (2 * radius + 1) * 2 ** (num_levels - 1)
        out.append((int(nIncre), 0.0025 / incre, int(nIters)))
    return out


class _DevView:
    """Expose plan-owned device memory to torch (zero-copy) through __cuda_array_interface__."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class DepthHotPath:
    def __init__(self, h1, w1, max_views, cascade=((64, 64, 8), (-1, 320, 8)), feats_f16=True, use_graph=True,
                 device=None):
        self.h1, self.w1, self.max_views = int(h1), int(w1), int(max_views)
        self.stages = stage_params(cascade)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DepthHotPath: a CUDA device is required (cer_mvs_b200 has no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        cfg = _lib.PlanConfig()
        cfg.h, cfg.w, cfg.max_views, cfg.n_stages = self.h1, self.w1, self.max_views, len(self.stages)
        for s, (D, incre, iters) in enumerate(self.stages):
            cfg.D[s], cfg.incre[s], cfg.iters[s] = D, incre, iters
        cfg.feats_f16, cfg.use_graph = int(feats_f16), int(use_graph)
        self._plan = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cer_plan_create(C.byref(cfg), C.byref(self._plan)), "cer_plan_create")
        self._out = torch.empty(1, 1, self.h1, self.w1, device=self.device, dtype=torch.float32)

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                _lib.lib().cer_plan_destroy(self._plan)
                self._plan = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    # ---- weights ----
    def load_update_block(self, sd):
        """sd: UpdateBlock state dict (reference keys; tensors or arrays; a 'module.update_block.' or
        'update_block.' prefix from a RAFT / DataParallel checkpoint is stripped, inference.py:31-35)."""
        clean = {}
        for k, v in sd.items():
            for pre in ("module.", "update_block."):
                if k.startswith(pre):
                    k = k[len(pre):]
            clean[k] = v
        blob = pack_update_weights(clean)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cer_plan_set_weights(self._plan, blob.ctypes.data), "cer_plan_set_weights")

    # ---- input prep shared by the device paths (core/raft.py:35-39) ----
    def _prep_cameras(self, poses, intrinsics, scale):
        P = poses.reshape(-1, 4, 4).to(self.device, torch.float32).clone()
        if scale is not None:
            P[:, :3, 3] *= float(scale)
        K = intrinsics.reshape(-1, 3, 3).to(self.device, torch.float32).clone()
        K[:, :2] /= 4
        return P.contiguous(), K.contiguous()

    def _check_maps(self, fmaps, net, inp, poses=None, intrinsics=None, out=None):
        """Shapes, dtypes, devices of everything whose raw pointer goes to the C ABI (device entry points)."""
        if fmaps.dim() != 5 or fmaps.shape[0] != 1 or fmaps.shape[2] != 64 or tuple(fmaps.shape[3:]) != (self.h1, self.w1):
            raise RuntimeError(f"fmaps must be [1,V+1,64,{self.h1},{self.w1}]")
        n_views = fmaps.shape[1] - 1
        if not 1 <= n_views <= self.max_views:
            raise RuntimeError(f"number of source views {n_views} outside 1..{self.max_views}")
        for name, t in (("fmaps", fmaps), ("net", net), ("inp", inp)):
            if not t.is_cuda:
                raise RuntimeError("DepthHotPath: device tensors expected (use run_host for host buffers)")
            if t.device != self.device:
                raise RuntimeError(f"DepthHotPath: {name} lives on {t.device}, the plan on {self.device}")
            if t.dtype not in (torch.float16, torch.float32):
                raise RuntimeError("DepthHotPath: float16 or float32 maps expected")
        for name, t in (("net", net), ("inp", inp)):
            if t.numel() != 64 * self.h1 * self.w1 or tuple(t.shape[-3:]) != (64, self.h1, self.w1):
                raise RuntimeError(f"{name} must be [1,1,64,{self.h1},{self.w1}]")
        if net.dtype != inp.dtype:
            raise RuntimeError("net and inp must share a dtype")
        self._check_cameras(poses, intrinsics, n_views)
        if out is not None:
            if not (out.is_cuda and out.device == self.device and out.dtype == torch.float32 and out.is_contiguous()
                    and out.numel() == self.h1 * self.w1):
                raise RuntimeError(f"out must be a contiguous float32 tensor of {self.h1}x{self.w1} elements on {self.device}")
        return n_views

    @staticmethod
    def _check_cameras(poses, intrinsics, n_views):
        if poses is not None and int(np.prod(tuple(poses.shape))) != (n_views + 1) * 16:
            raise RuntimeError(f"poses must hold {n_views + 1} 4x4 matrices (reference + source views)")
        if intrinsics is not None and int(np.prod(tuple(intrinsics.shape))) != (n_views + 1) * 9:
            raise RuntimeError(f"intrinsics must hold {n_views + 1} 3x3 matrices (reference + source views)")

    def _check_host(self, fm, nt, ip, poses, intrinsics, o, who):
        """The C side copies (n_views+1)*64*px elements from `fm`, 64*px from `nt` / `ip` and writes px floats into `o`:
        every size is checked here, a mismatch would be an out-of-bounds host access."""
        if fm.ndim != 5 or fm.shape[0] != 1 or fm.shape[2] != 64 or tuple(fm.shape[3:]) != (self.h1, self.w1):
            raise RuntimeError(f"{who}: fmaps must be [1,V+1,64,{self.h1},{self.w1}]")
        n_views = fm.shape[1] - 1
        if not 1 <= n_views <= self.max_views:
            raise RuntimeError(f"{who}: number of source views {n_views} outside 1..{self.max_views}")
        for name, a in (("fmaps", fm), ("net", nt), ("inp", ip)):
            if a.dtype not in (np.float16, np.float32) or not a.flags["C_CONTIGUOUS"]:
                raise RuntimeError(f"{who}: {name}: contiguous float16/float32 array expected")
        for name, a in (("net", nt), ("inp", ip)):
            if a.size != 64 * self.h1 * self.w1 or tuple(a.shape[-3:]) != (64, self.h1, self.w1):
                raise RuntimeError(f"{who}: {name} must be [1,1,64,{self.h1},{self.w1}]")
        if nt.dtype != ip.dtype:
            raise RuntimeError(f"{who}: net and inp must share a dtype")
        self._check_cameras(poses, intrinsics, n_views)
        if o.dtype != np.float32 or not o.flags["C_CONTIGUOUS"] or o.size != self.h1 * self.w1 or not o.flags["WRITEABLE"]:
            raise RuntimeError(f"{who}: out must be a writable contiguous float32 array of {self.h1}x{self.w1} elements")
        return n_views

    def __call__(self, fmaps, net, inp, poses, intrinsics, scale=1.0, out=None):
        n_views = self._check_maps(fmaps, net, inp, poses, intrinsics, out)
        with torch.cuda.device(self.device):
            P, K = self._prep_cameras(poses, intrinsics, scale)
            fmaps, net, inp = fmaps.contiguous(), net.contiguous(), inp.contiguous()
            out = self._out if out is None else out
            _lib.check(_lib.lib().cer_plan_run_device(
                self._plan, fmaps.data_ptr(), int(fmaps.dtype == torch.float16), net.data_ptr(), inp.data_ptr(),
                int(net.dtype == torch.float16), P.data_ptr(), K.data_ptr(), n_views,
                1.0 if scale is None else float(scale), out.data_ptr(), _lib.stream_ptr()), "cer_plan_run_device")
        return out

    def run_host(self, fmaps, net, inp, poses, intrinsics, scale=1.0, out=None):
        """Host buffers in (numpy arrays or CPU tensors, ideally pinned), host disparity out.
        poses / intrinsics as RAFT.forward receives them (scaling and the /4 are done here)."""
        def as_np(x):
            return x.numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        fm, nt, ip = as_np(fmaps), as_np(net), as_np(inp)
        if out is None:
            out = np.empty((1, 1, self.h1, self.w1), np.float32)
        o = as_np(out)
        n_views = self._check_host(fm, nt, ip, as_np(poses), as_np(intrinsics), o, "run_host")
        P = np.array(as_np(poses), dtype=np.float32).reshape(-1, 4, 4)
        if scale is not None:
            P[:, :3, 3] *= np.float32(scale)
        K = np.array(as_np(intrinsics), dtype=np.float32).reshape(-1, 3, 3)
        K[:, :2] /= 4
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cer_plan_run_host(
                self._plan, fm.ctypes.data, int(fm.dtype == np.float16), nt.ctypes.data, ip.ctypes.data,
                int(nt.dtype == np.float16), P.ctypes.data, K.ctypes.data, n_views,
                1.0 if scale is None else float(scale), o.ctypes.data, _lib.stream_ptr()), "cer_plan_run_host")
        return out

    def submit_host(self, fmaps, net, inp, poses, intrinsics, scale=1.0, out=None):
        """Pipelined run_host: enqueue one depth map (pinned host buffers) and return; its H2D copy overlaps the kernels
        of the previous job.  Call wait_host() to wait for the oldest job; `out` (pinned) is valid after that."""
        def as_np(x):
            return x.numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        fm, nt, ip = as_np(fmaps), as_np(net), as_np(inp)
        if out is None:
            out = torch.empty(1, 1, self.h1, self.w1).pin_memory()
        o = as_np(out)
        n_views = self._check_host(fm, nt, ip, as_np(poses), as_np(intrinsics), o, "submit_host")
        P = np.array(as_np(poses), dtype=np.float32).reshape(-1, 4, 4)
        if scale is not None:
            P[:, :3, 3] *= np.float32(scale)
        K = np.array(as_np(intrinsics), dtype=np.float32).reshape(-1, 3, 3)
        K[:, :2] /= 4
        self._keep = getattr(self, "_keep", [])[-3:] + [(fm, nt, ip, P, K, o)]     # keep host buffers alive while in flight
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cer_plan_submit_host(
                self._plan, fm.ctypes.data, int(fm.dtype == np.float16), nt.ctypes.data, ip.ctypes.data,
                int(nt.dtype == np.float16), P.ctypes.data, K.ctypes.data, n_views,
                1.0 if scale is None else float(scale), o.ctypes.data, _lib.stream_ptr()), "cer_plan_submit_host")
        return out

    def wait_host(self):
        _lib.check(_lib.lib().cer_plan_wait_host(self._plan), "cer_plan_wait_host")

    def forward_sharded(self, fmaps, net, inp, poses, intrinsics, scale=1.0, group=None, out=None):
        """One depth map over all ranks of ``group``.  The cost-volume build of a stage is
        n_views * D independent (view, hypothesis) units (core/corr.py:84-91 sums independent per-view, per-hypothesis
        terms); rank g builds the contiguous run ``unit_range(n_views * D, g, G)`` into a zeroed partial volume, ONE
        all-reduce(sum) per stage adds the partial volumes, the GRU loop then runs replicated.  Any number of ranks
        works, also more ranks than source views (BASELINE configs[4]: 7 views on 8 GPUs).  Every rank passes the
        same inputs; only the views a rank touches are converted to the kernels' layout."""
        import torch.distributed as dist
        from .dist import unit_range, views_of_units
        n_views = self._check_maps(fmaps, net, inp, poses, intrinsics, out)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        units = [unit_range(n_views * D, rank, world) for (D, _, _) in self.stages]
        touched = [views_of_units(ub, ue, D) for (ub, ue), (D, _, _) in zip(units, self.stages)]
        vb = min((t[0] for t in touched if t[1] > t[0]), default=0)
        ve = max((t[1] for t in touched if t[1] > t[0]), default=0)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr()
            P, K = self._prep_cameras(poses, intrinsics, scale)
            fmaps, net, inp = fmaps.contiguous(), net.contiguous(), inp.contiguous()
            out = self._out if out is None else out
            _lib.check(L.cer_plan_prepare(self._plan, fmaps.data_ptr(), int(fmaps.dtype == torch.float16),
                                          net.data_ptr(), inp.data_ptr(), int(net.dtype == torch.float16),
                                          P.data_ptr(), K.data_ptr(), n_views, vb, ve, st), "cer_plan_prepare")
            for s in range(len(self.stages)):
                n = C.c_size_t()
                ptr = L.cer_plan_partial_volume(self._plan, s, C.byref(n))
                vol = torch.as_tensor(_DevView(ptr, n.value), device=self.device)
                ub, ue = units[s]
                _lib.check(L.cer_plan_build_stage_units(self._plan, s, ub, ue, st), "cer_plan_build_stage_units")
                dist.all_reduce(vol, op=dist.ReduceOp.SUM, group=group)
                _lib.check(L.cer_plan_iterate_stage(self._plan, s, st), "cer_plan_iterate_stage")
            _lib.check(L.cer_plan_finish(self._plan, 1.0 if scale is None else float(scale), out.data_ptr(), st),
                       "cer_plan_finish")
        return out

    forward_view_sharded = forward_sharded      # round-1 name

    KERNEL_KINDS = ["layout", "projection", "volume_build", "pool", "lookup", "corr_dropin", "disp_encoder",
                    "corr_enc_1x1", "conv_corr_enc_3x3", "conv_gates", "conv_q_gru", "conv_delta", "disp_update",
                    "finish"]

    def set_kernel_timing(self, enable: bool):
        """Eager mode with CUDA events around every kernel (bench.py's per-kernel breakdown)."""
        _lib.check(_lib.lib().cer_plan_set_kernel_timing(self._plan, int(enable)), "set_kernel_timing")

    def kernel_times(self):
        """{kernel class: (total ms, launches)} since the last call; synchronises the current stream."""
        n = len(self.KERNEL_KINDS)
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cer_plan_kernel_times(self._plan, ms, cnt, n, _lib.stream_ptr()), "kernel_times")
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.KERNEL_KINDS)}

    @property
    def last_launch_count(self):
        return int(_lib.lib().cer_plan_last_launch_count(self._plan))

    @property
    def workspace_bytes(self):
        return int(_lib.lib().cer_plan_workspace_bytes(self._plan))
