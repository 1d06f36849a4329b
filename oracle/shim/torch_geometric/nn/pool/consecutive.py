import torch

def consecutive_cluster(src):
    """Appendix A.5: CPU scatter_ keeps the highest member index per cluster."""
    unique, inv = torch.unique(src, sorted=True, return_inverse=True)
    perm = torch.arange(inv.size(0), dtype=inv.dtype, device=inv.device)
    perm = inv.new_empty(unique.size(0)).scatter_(0, inv, perm)
    return inv, perm
