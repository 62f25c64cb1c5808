"""cer_mvs_b200 -- B200-native (sm_100a) implementation of the CER-MVS inference hot path.

Python surface = the reference's own operator surface (princeton-vl/CER-MVS):

    cer_mvs_b200.alt_cuda_corr.forward(fmap1, fmap2, coords, radius)   alt_cuda_corr/correlation.cpp:52
    cer_mvs_b200.corr.CorrBlock                                         core/corr.py:45
    cer_mvs_b200.update.ConvGRU / UpdateBlock                           core/update.py:9,29
    cer_mvs_b200.hotpath.DepthHotPath                                   core/raft.py:75-108 as one native plan
    cer_mvs_b200.install.install()                                      module substitution for inference.py

Everything computes in libcer_mvs_b200.so (hand-written CUDA, C ABI in include/cer_mvs_b200.h);
there is no CPU or PyTorch fallback: importing the op modules without the library raises.
"""
__version__ = "0.1.0"
