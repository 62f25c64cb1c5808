"""Make the reference's own ``inference.py`` / ``core/raft.py`` run on the B200-native ops without
editing them: module substitution.

    import cer_mvs_b200.install as I; I.install()        # before `from core.raft import RAFT`
    from core.raft import RAFT                            # reference code, unchanged

* ``alt_cuda_corr`` (imported at core/corr.py:3)          -> cer_mvs_b200.alt_cuda_corr
* ``core.corr.CorrBlock``                                 -> cer_mvs_b200.corr.CorrBlock
* ``core.update.ConvGRU`` / ``UpdateBlock``               -> cer_mvs_b200.update.*
and, because core/raft.py binds the names at import time (``from core.corr import CorrBlock``),
``core.raft.CorrBlock`` / ``core.raft.UpdateBlock`` are rebound too when that module is loaded.
"""
import importlib
import sys


def install(patch_loaded: bool = True):
    from . import alt_cuda_corr, corr, update
    sys.modules["alt_cuda_corr"] = alt_cuda_corr
    if not patch_loaded:
        return
    for modname, names in (("core.corr", {"CorrBlock": corr.CorrBlock}),
                           ("core.update", {"ConvGRU": update.ConvGRU, "UpdateBlock": update.UpdateBlock}),
                           ("core.raft", {"CorrBlock": corr.CorrBlock, "UpdateBlock": update.UpdateBlock})):
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:  # noqa: BLE001  reference not on sys.path: only alt_cuda_corr is substituted
                continue
        for k, v in names.items():
            setattr(mod, k, v)
