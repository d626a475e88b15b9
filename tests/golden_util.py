"""Shared helpers: load golden fixtures (generated from the real reference by oracle/make_golden.py)
and run the oracle on the same seeded inputs."""
import glob
import os

import torch

from oracle import camradepth_oracle as O
from camradepth_b200.synthetic import make_batch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLD, "ref_*.pt")))


def load_case(path):
    g = torch.load(path, weights_only=False)
    cfg = O.Cfg(g["variant"])
    sd = O.init_state_dict(cfg, seed=1, perturb=0.05)
    batch = make_batch(g["B"], g["H"], g["W"], seed=3, input_channels=cfg.cin)
    masks = O.make_masks(cfg, g["B"], seed=11) if g["train"] else (None, None)
    return g, cfg, sd, batch, masks


def samp(t):
    return t.detach().flatten()[::max(1, t.numel() // 4096)]


def relerr(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def oracle_run(cfg, sd, batch, masks, dtype=torch.float32):
    sd = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    dps, d2s = masks
    pred = O.forward(sd, cfg, batch["image"].to(dtype), dps, d2s)
    loss, parts = O.training_loss(pred, batch["gt_final"].to(dtype), batch["gt_s4"].to(dtype),
                                  batch["gt_s3"].to(dtype), batch["gt_seg"], cfg)
    loss.backward()
    grads = {k: v.grad for k, v in sd.items()}
    return pred, loss, parts, grads, sd
