"""Make the reference's own ``inference.py`` / ``core/raft.py`` run on the B200-native ops without
editing them: module substitution.

    import cer_mvs_b200.install as I; I.install()        # before `from core.raft import RAFT`
    from core.raft import RAFT                            # reference code, unchanged

* ``alt_cuda_corr`` (imported at core/corr.py:3)          -> cer_mvs_b200.alt_cuda_corr
* ``core.corr.CorrBlock``                                 -> cer_mvs_b200.corr.CorrBlock
* ``core.update.ConvGRU`` / ``UpdateBlock``               -> cer_mvs_b200.update.*
* ``core.extractor.BasicEncoder`` (``encoders=True``)     -> cer_mvs_b200.extractor.BasicEncoder
and, because core/raft.py binds the names at import time (``from core.corr import CorrBlock``),
``core.raft.CorrBlock`` / ``core.raft.UpdateBlock`` are rebound too when that module is loaded.
"""
import importlib
import sys


def install(patch_loaded: bool = True, encoders: bool = False):
    """encoders=True also substitutes the feature / context encoders (core/raft.py:28-29 builds them from
    ``BasicEncoder``); RAFT.forward then runs without a cuDNN call.  (For the fully fused pipeline -- encoders writing
    straight into the plan's buffers -- use ``cer_mvs_b200.raft.RAFT`` in place of ``core.raft.RAFT``.)"""
    from . import alt_cuda_corr, corr, update
    sys.modules["alt_cuda_corr"] = alt_cuda_corr
    if not patch_loaded:
        return
    targets = [("core.corr", {"CorrBlock": corr.CorrBlock}),
               ("core.update", {"ConvGRU": update.ConvGRU, "UpdateBlock": update.UpdateBlock}),
               ("core.raft", {"CorrBlock": corr.CorrBlock, "UpdateBlock": update.UpdateBlock})]
    if encoders:
        from . import extractor
        targets += [("core.extractor", {"BasicEncoder": extractor.BasicEncoder}),
                    ("core.raft", {"BasicEncoder": extractor.BasicEncoder})]
    for modname, names in targets:
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:  # noqa: BLE001  reference not on sys.path: only alt_cuda_corr is substituted
                continue
        for k, v in names.items():
            setattr(mod, k, v)
