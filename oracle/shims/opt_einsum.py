"""Import shim for opt_einsum.contract (absent here) -> torch.einsum. Test infrastructure only."""
import torch


def contract(eq, *ops):
    return torch.einsum(eq, *ops)
