"""Import the read-only reference (/root/reference) on CPU through the three shims.
TEST INFRASTRUCTURE ONLY; works only in the build container (the GPU box has no /root/reference).

* gin / fastcore / opt_einsum -> oracle/shims
* alt_cuda_corr (CUDA-only extension) -> oracle.cer_oracle.corr_forward
* ``Tensor.cuda()`` -> identity (core/corr.py:60, core/raft.py:108 call it unconditionally)
"""
import os
import sys
import types

import torch

REF = os.environ.get("CER_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "core", "corr.py"))


def import_reference():
    """Returns the reference modules (core.corr, core.update, core.raft, utils.projective_ops)."""
    if not available():
        raise RuntimeError("reference tree not present")
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, REF)
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import cer_oracle
    shim = types.ModuleType("alt_cuda_corr")
    shim.forward = cer_oracle.corr_forward

    def _bwd(*a, **k):
        raise NotImplementedError("inference-only oracle")
    shim.backward = _bwd
    sys.modules["alt_cuda_corr"] = shim
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import core.corr
    import core.raft
    import core.update
    import utils.projective_ops
    return types.SimpleNamespace(corr=core.corr, update=core.update, raft=core.raft,
                                 pops=utils.projective_ops)
