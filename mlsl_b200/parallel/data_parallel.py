import torch

from .. import comm


def broadcast_parameters(module, root=0, distribution=None):
    """Make every data-parallel replica start from rank `root`'s weights and buffers (Distribution::Bcast)."""
    works, keep = [], []
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            if t.numel() == 0:
                continue
            c = t.data if t.is_contiguous() else t.data.contiguous()
            works.append((comm.bcast(c, root=root, async_op=True, distribution=distribution), t, c))
        for w, t, c in works:
            w.wait()
            if c.data_ptr() != t.data_ptr():
                t.data.copy_(c)
    return module


class DistributedDataParallel(torch.nn.Module):
    """Thin wrapper: synchronises the initial weights; gradient exchange is done by mlsl_b200.DistributedOptimizer
    (which owns the gradient buckets), so forward/backward are the wrapped module's own."""

    def __init__(self, module, distribution=None):
        super().__init__()
        self.module = broadcast_parameters(module, 0, distribution)

    def forward(self, *a, **kw):
        return self.module(*a, **kw)
