"""Drop-in for the reference's ``alt_cuda_corr`` extension module
(alt_cuda_corr/correlation.cpp:52-53): ``forward`` and ``backward`` with the same signatures.

``forward`` runs ``cer_corr_forward_f32`` (csrc/corr_ops.cu).  ``backward`` is training-only
(correlation_kernel.cu:122-256) and outside the inference hot path: it raises.
"""
import torch

from . import _lib


def forward(fmap1: torch.Tensor, fmap2: torch.Tensor, coords: torch.Tensor, radius: int):
    """fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2] fp32 CUDA contiguous
    -> [corr [B,N,(2r+1)^2,H1,W1]] (a one-element list, like the reference)."""
    for name, t in (("fmap1", fmap1), ("fmap2", fmap2), ("coords", coords)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")          # CHECK_CUDA, correlation.cpp:19
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")             # CHECK_CONTIGUOUS, :20
        if t.dtype != torch.float32:
            raise RuntimeError(f"expected scalar type Float but found {t.dtype} for {name}")
    if fmap1.dim() != 4 or fmap2.dim() != 4 or coords.dim() != 5 or coords.shape[-1] != 2:
        raise RuntimeError("alt_cuda_corr.forward: expected fmap [B,H,W,C] and coords [B,N,H,W,2]")
    B, H1, W1, C = fmap1.shape
    _, H2, W2, C2 = fmap2.shape
    Bc, N, Hc, Wc, _ = coords.shape
    if C2 != C or fmap2.shape[0] != B or Bc != B or Hc != H1 or Wc != W1:
        raise RuntimeError("alt_cuda_corr.forward: inconsistent shapes")
    r = int(radius)
    rd = 2 * r + 1
    with torch.cuda.device(fmap1.device):
        corr = torch.empty(B, N, rd * rd, H1, W1, device=fmap1.device, dtype=torch.float32)
        _lib.check(_lib.lib().cer_corr_forward_f32(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(),
                                                   corr.data_ptr(), B, H1, W1, H2, W2, C, N, r,
                                                   _lib.stream_ptr()), "alt_cuda_corr.forward")
    return [corr]


def backward(fmap1, fmap2, coords, corr_grad, radius):
    raise NotImplementedError("cer_mvs_b200.alt_cuda_corr.backward: training is outside the inference hot path "
                              "(reference: alt_cuda_corr/correlation_kernel.cu:122-256)")
