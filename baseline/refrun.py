"""Run the UNMODIFIED reference (``baseline/_ref``, or ``/root/reference`` in the build container) as the checker
and as the timed baseline.  TEST / BASELINE INFRASTRUCTURE ONLY -- nothing in ``cer_mvs_b200/`` imports this.

Two ways of serving the reference's one native dependency, ``alt_cuda_corr`` (core/corr.py:3):

* ``mode="gpu"``: the reference's own kernel, compiled from its two source files by ``oracle/build_ref.py``
  (``oracle/_ref/alt_cuda_corr_ref*.so``).  Everything on the path is then reference code: core/raft.py,
  core/corr.py, core/update.py, utils/*, correlation_kernel.cu, under the real ``torch.cuda.amp.autocast``.
* ``mode="cpu"``: ``alt_cuda_corr.forward`` is CUDA-only, so the CPU run uses the oracle's restatement of that one
  function (``oracle/cer_oracle.corr_forward``, pinned on the GPU against the compiled kernel) and turns the
  reference's unconditional ``.cuda()`` calls (core/corr.py:60, core/raft.py:108) into the identity.

The three pure-Python dependencies that are absent here (gin, fastcore, opt_einsum; no network) are the 5-line import
shims of ``oracle/shims``.  The encoders (core/extractor.py) are outside the hot path: the runs
below replace ``fnet`` / ``cnet`` by stubs that hand back given feature / context maps, so the timed / compared
region is exactly core/raft.py:75-108.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ORACLE = os.path.join(ROOT, "oracle")
_state = {"mode": None, "ns": None, "tensor_cuda": torch.Tensor.cuda, "module_cuda": torch.nn.Module.cuda}


def ref_dir():
    for d in (os.path.join(HERE, "_ref"), os.environ.get("CER_REFERENCE_DIR", "/root/reference")):
        if os.path.isfile(os.path.join(d, "core", "raft.py")):
            return d
    return None


def available(mode="gpu") -> bool:
    if ref_dir() is None:
        return False
    if mode == "gpu":
        sys.path.insert(0, ORACLE) if ORACLE not in sys.path else None
        import build_ref
        return os.path.isfile(build_ref.so_path()) and torch.cuda.is_available()
    return True


def _cpu_shim():
    import cer_oracle
    shim = types.ModuleType("alt_cuda_corr")
    shim.forward = cer_oracle.corr_forward

    def _bwd(*a, **k):
        raise NotImplementedError("inference only")
    shim.backward = _bwd
    return shim


def import_reference(mode="gpu"):
    """Returns a namespace with the reference's modules (raft, corr, update, pops), switched to ``mode``.
    The modules are imported once; switching the mode swaps the ``alt_cuda_corr`` binding of core/corr.py and the
    ``.cuda()`` patch."""
    d = ref_dir()
    if d is None:
        raise RuntimeError("no reference tree (run baseline/install_ref.py in the build container)")
    for p in (ORACLE, os.path.join(ORACLE, "shims"), d):
        if p not in sys.path:
            sys.path.insert(0, p)
    if mode == "gpu":
        import build_ref
        ext = _state.get("ext") or build_ref.load()      # the reference's kernel, its own sources
        _state["ext"] = ext
        torch.Tensor.cuda, torch.nn.Module.cuda = _state["tensor_cuda"], _state["module_cuda"]
    else:
        ext = _state.get("shim") or _cpu_shim()
        _state["shim"] = ext
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.modules["alt_cuda_corr"] = ext
    ns = _state["ns"]
    if ns is None:
        import core.corr
        import core.raft
        import core.update
        import utils.projective_ops
        ns = types.SimpleNamespace(corr=core.corr, update=core.update, raft=core.raft, pops=utils.projective_ops,
                                   originals=dict(CorrBlock=core.corr.CorrBlock, UpdateBlock=core.update.UpdateBlock,
                                                  ConvGRU=core.update.ConvGRU))
    ns.corr.alt_cuda_corr = ext                           # `import alt_cuda_corr` binding at core/corr.py:3
    ns.originals["alt_cuda_corr"] = ext
    _state.update(mode=mode, ns=ns)
    return ns


def restore_reference_classes():
    """Undo cer_mvs_b200.install.install(): put the reference's own classes / extension back."""
    ns = _state["ns"]
    if ns is None:
        return
    o = ns.originals
    ns.corr.CorrBlock = o["CorrBlock"]
    ns.corr.alt_cuda_corr = o["alt_cuda_corr"]
    ns.update.UpdateBlock, ns.update.ConvGRU = o["UpdateBlock"], o["ConvGRU"]
    ns.raft.CorrBlock, ns.raft.UpdateBlock = o["CorrBlock"], o["UpdateBlock"]
    sys.modules["alt_cuda_corr"] = o["alt_cuda_corr"]


class _Stub(torch.nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


def make_model(ref, sd, cascade, fmaps, ctx_pre, device):
    """The reference's RAFT (test_mode) with UpdateBlock weights ``sd`` and stub encoders.

    fmaps   [1,V+1,64,h1,w1]  what fnet emits per image (fp16 on the GPU: autocast, core/raft.py:55,66-69)
    ctx_pre [1,1,128,h1,w1]   what cnet emits before the tanh / relu split (core/raft.py:57-60)
    """
    model = ref.raft.RAFT(cascade=cascade, test_mode=True)
    model.update_block.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
    model = model.to(device).eval()
    calls = {"i": 0}
    n_img = fmaps.shape[1]

    def fnet(x):
        i = calls["i"] % n_img
        calls["i"] += 1
        return fmaps[:, [i]]

    model.fnet, model.cnet = _Stub(fnet), _Stub(lambda x: ctx_pre)
    return model


def run_forward(model, images, poses, intrinsics, scale):
    """RAFT.forward as inference.py:54 calls it (poses / intrinsics are modified in place by the reference, so copies
    go in; ``images`` is only read for its shape by the stub encoders but normalised in place, core/raft.py:40-41)."""
    # `scale`: the DataLoader collates the dataset's Python float into a float64 tensor of shape [1]
    # (datasets/dtu.py:276); inference.py:33-34 wraps the model in nn.DataParallel for the released checkpoints,
    # whose scatter moves it to the GPU with the other arguments
    with torch.no_grad():
        return model(images, poses.clone(), intrinsics.clone(),
                     scale=torch.tensor([scale], dtype=torch.float64, device=poses.device))


def time_forward(model, images, poses, intrinsics, scale, steps=3, warmup=1):
    """CUDA-event time of the hot path of the reference on the current device: ms per depth map (mean)."""
    for _ in range(warmup):
        run_forward(model, images, poses, intrinsics, scale)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = run_forward(model, images, poses, intrinsics, scale)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out
