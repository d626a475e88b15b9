"""Test-mode metrics of the reference's evaluation loop (src/main/runner.py:442-492) as fused on-device
reductions: masked RMSE / MAE / REL within 100 m and within 50 m, and mean IoU of the segmentation logits.

Reference semantics kept: prediction clipped to [0,1], both maps scaled by `max_depth`, valid pixels are
`0 < gt <= max_distances[0]` (runner.py:455-457); the 50 m subset is taken in inverse-depth space (`gt >= 50` after scaling, runner.py:473-477).
The reference's IoU call constructs `torchmetrics.JaccardIndex(ignore_index=255)` inside a try/except that
swallows the resulting ValueError (runner.py:433-439), i.e. it reports NaN; here IoU is the standard mean over
the classes present in prediction or label, ignoring label 255.
"""
from __future__ import annotations

import torch

from ._lib import K
from .ops import P, stream


def depth_metrics(pred_depth: torch.Tensor, gt_depth: torch.Tensor, max_depth: float = 100.0,
                  max_distances=(100, 50)) -> dict:
    """-> {"rmse_100","mae_100","rel_100","rmse_50","mae_50","rel_50"} as 0-d CUDA tensors (no host sync)."""
    if not pred_depth.is_cuda:
        raise RuntimeError("camradepth_b200 metrics run on CUDA devices only (no CPU fallback by design)")
    p = pred_depth.detach().contiguous().float()
    g = gt_depth.detach().contiguous().float()
    assert p.numel() == g.numel()
    acc = torch.empty(8, dtype=torch.float32, device=p.device)
    out = torch.empty(6, dtype=torch.float32, device=p.device)
    K.crd_depth_metrics(P(p), P(g), P(acc), P(out), p.numel(), float(max_depth), float(max_distances[0]),
                        float(max_distances[1]), stream())
    d0, d1 = int(max_distances[0]), int(max_distances[1])
    return {f"rmse_{d0}": out[0], f"mae_{d0}": out[1], f"rel_{d0}": out[2],
            f"rmse_{d1}": out[3], f"mae_{d1}": out[4], f"rel_{d1}": out[5]}


def confusion_matrix(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = 255) -> torch.Tensor:
    lg = logits.detach().contiguous().float()
    tg = target.detach().contiguous().long()
    B, C = lg.shape[0], lg.shape[1]
    conf = torch.zeros(C, C, dtype=torch.float32, device=lg.device)
    K.crd_confusion(P(lg), P(tg), P(conf), B, C, lg.numel() // (B * C), ignore_index, stream())
    return conf


def mean_iou(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = 255) -> torch.Tensor:
    conf = confusion_matrix(logits, target, ignore_index)
    inter = conf.diag()
    union = conf.sum(0) + conf.sum(1) - inter
    present = union > 0
    return (inter[present] / union[present]).mean()
