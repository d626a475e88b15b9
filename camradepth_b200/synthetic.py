"""Seeded synthetic RGB+radar batches with the tensor contract of the reference dataloader.

Follows SURVEY.md §8(d) "Synthetic inputs"; contract source: src/data/dataloader.py:202-333
(channel order RGB(ImageNet-normalised) | radar depth/100 | radar u | radar v | radar velocity,
inverse-normalised lidar GT with a 3-level zero-ignoring min-pool pyramid, mseg labels with 255 = ignore).
Everything is generated on the HOST (CPU generator) like the reference's DataLoader does.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def minpool_gt(t: torch.Tensor) -> torch.Tensor:
    """Zero-ignoring 3x3 stride-2 min-pool of a (B,1,H,W) GT map (dataloader.py:213-222)."""
    x = t.clone()
    x[t == 0] = 255
    x = -F.max_pool2d(-x, kernel_size=3, stride=2, padding=1)
    x[x == 255] = 0
    return x


def make_batch(B: int, H: int, W: int, seed: int = 0, input_channels: int = 7, num_classes: int = 21,
               pin: bool = False):
    """Returns dict(image (B,C,H,W) f32, gt_final (B,1,H,W), gt_s4 (B,1,H/2,W/2), gt_s3 (B,1,H/4,W/4),
    gt_seg (B,H,W) int64)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(B, 7, H, W)
    x[:, 0:3] = torch.randn(B, 3, H, W, generator=g).clamp_(-2.12, 2.64)
    mask = (torch.rand(B, 1, H, W, generator=g) < 0.002).float()
    x[:, 3:4] = mask * (0.01 + 0.99 * torch.rand(B, 1, H, W, generator=g))
    x[:, 4:5] = mask * (-0.8 + 1.6 * torch.rand(B, 1, H, W, generator=g))
    x[:, 5:6] = mask * (-0.4 + 0.9 * torch.rand(B, 1, H, W, generator=g))
    x[:, 6:7] = mask * (torch.rand(B, 1, H, W, generator=g) < 0.3).float()
    x = x[:, :input_channels].contiguous()
    gmask = (torch.rand(B, 1, H, W, generator=g) < 0.15).float()
    gt = gmask * (0.01 + 0.98 * torch.rand(B, 1, H, W, generator=g))
    gt_s4 = minpool_gt(gt)
    gt_s3 = minpool_gt(gt_s4)
    seg = torch.randint(0, num_classes, (B, H, W), generator=g)
    seg[torch.rand(B, H, W, generator=g) < 0.1] = 255
    out = {"image": x, "gt_final": gt, "gt_s4": gt_s4, "gt_s3": gt_s3, "gt_seg": seg}
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return out
