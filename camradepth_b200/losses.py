"""Fused CUDA losses with the reference's class names and call signatures
(src/utils/loss_funcs.py:14-34 MaskedFocalLoss, :36-46 MaskedMSELoss, :77-91 MaskedSmoothL1Loss).
One masked reduction + analytic backward each; no boolean-index compaction, no host sync."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


# Data-parallel loss normalisation.  The reference computes every loss ONCE over the whole gathered batch
# (nn.DataParallel gathers the outputs on GPU 0, runner.py:193-203): a masked mean over all valid pixels of all
# samples.  With one process per GPU each rank only sees its shard, so when a world is configured the per-rank
# partial sums (sum, valid count) are all-reduced before the division: the loss VALUE is the global masked mean on
# every rank, and the gradient each rank back-propagates is world * d(global loss)/d(local prediction), so that
# the gradient all-reduce(AVG) of parallel.DataParallel yields exactly the single-process batch gradient
# (SURVEY.md §8e caveat 1).  world == 1: nothing changes.
_DP = {"world": 1, "group": None}


def set_data_parallel(world: int = 1, group=None):
    """Called by parallel.DataParallel; `set_data_parallel(1)` restores per-process losses."""
    _DP["world"], _DP["group"] = int(world), group


def _global_sums(acc):
    """all-reduce(SUM) of the loss accumulators over the data-parallel world (no-op for one process)."""
    if _DP["world"] > 1:
        import torch.distributed as dist
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=_DP["group"])
    return float(_DP["world"])


def _need_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("camradepth_b200 losses run on CUDA devices only (no CPU fallback by design)")


class _MaskedSmoothL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        _need_cuda(pred)
        p = pred.detach().contiguous().float()
        t = target.detach().contiguous().float()
        acc = torch.zeros(3, dtype=torch.float32, device=p.device)
        out = torch.empty(2, dtype=torch.float32, device=p.device)
        ops.masked_l1_fwd(p, t, acc)
        ctx.gscale = _global_sums(acc)
        ops.loss_finalize(acc, out, 0)
        ctx.save_for_backward(p, t, acc)
        ctx.shape = pred.shape
        ctx.rmse = out[1]
        return out[0].clone()

    @staticmethod
    def backward(ctx, gout):
        p, t, acc = ctx.saved_tensors
        dpred = torch.empty_like(p)
        ops.masked_l1_bwd(p, t, acc, (gout.detach().float() * ctx.gscale).contiguous().view(1), dpred)
        return dpred.view(ctx.shape), None


class MaskedSmoothL1Loss(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.eps = 1e-8

    def forward(self, pred, target):
        assert pred.dim() == target.dim(), "inconsistent dimensions"
        return _MaskedSmoothL1.apply(pred, target)


class MaskedMSELoss(nn.Module):
    """Metric only in the reference (runner.py:208); no gradient."""

    def forward(self, pred, target):
        assert pred.dim() == target.dim(), "inconsistent dimensions"
        _need_cuda(pred)
        p = pred.detach().contiguous().float()
        t = target.detach().contiguous().float()
        acc = torch.zeros(3, dtype=torch.float32, device=p.device)
        out = torch.empty(2, dtype=torch.float32, device=p.device)
        ops.masked_l1_fwd(p, t, acc)
        _global_sums(acc)
        ops.loss_finalize(acc, out, 0)
        self.loss = out[1] * out[1]
        return self.loss


class _MaskedFocal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, gamma):
        _need_cuda(logits)
        lg = logits.detach().contiguous().float()
        tg = target.detach().contiguous().long()
        acc = torch.zeros(2, dtype=torch.float32, device=lg.device)
        out = torch.empty(2, dtype=torch.float32, device=lg.device)
        ops.ce_fwd(lg, tg, acc)
        ctx.gscale = _global_sums(acc)
        ops.loss_finalize(acc, out, 1, float(gamma))
        ctx.save_for_backward(lg, tg, acc)
        ctx.gamma = float(gamma)
        return out[0].clone()

    @staticmethod
    def backward(ctx, gout):
        lg, tg, acc = ctx.saved_tensors
        d = torch.empty_like(lg)
        ops.ce_bwd(lg, tg, acc, (gout.detach().float() * ctx.gscale).contiguous().view(1), ctx.gamma, d)
        return d, None, None


class MaskedFocalLoss(nn.Module):
    ''' Focal transform of the scalar mean cross-entropy (ignore_index=255), as the reference computes it '''

    def __init__(self, weight=None, gamma=2, reduction='mean'):
        super().__init__()
        self.gamma = gamma
        self.reduction = reduction

    def forward(self, inputs, target):
        return _MaskedFocal.apply(inputs, target, self.gamma)
