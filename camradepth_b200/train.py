"""Training-step driver with the semantics of `Trainer.train_one_epoch` (src/main/runner.py:166-270) minus
its four per-batch `.item()` host syncs: loss mix (1,1,1,0.2,0.2)/3.4 divided by `update_interval`, gradient
accumulation over `update_interval` micro-batches, optimizer step + `zero_grad(set_to_none=True)` on the
boundary (or on the last batch of the epoch), and the reference's scheduler rule (one `scheduler.step()`
per micro-batch once `i + 1 > update_interval`, runner.py:269-270).  Running statistics stay on the device.
"""
from __future__ import annotations

import torch

from .args import args
from .losses import MaskedFocalLoss, MaskedMSELoss, MaskedSmoothL1Loss

LOSS_WEIGHTS = [1, 1, 1, 0.2, 0.2]          # runner.py:212


class TrainStep:
    def __init__(self, model, optimizer, scheduler=None, update_interval=None, supervised_seg=None,
                 batches_per_epoch=None):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.update_interval = int(update_interval if update_interval is not None else args.get("update_interval", 1))
        self.supervised_seg = bool(args.get("supervised_seg", False) if supervised_seg is None else supervised_seg)
        self.batches_per_epoch = batches_per_epoch
        self.criterion = {"depth": MaskedSmoothL1Loss(), "seg": MaskedFocalLoss()}
        self.mse = MaskedMSELoss()
        self.i = 0                              # micro-batch index within the epoch
        self.training_steps = 0
        self._stats = None
        self._stats_n = 0                       # micro-batches accumulated into _stats since the last stats()

    def start_epoch(self):
        self.i = 0
        self._stats, self._stats_n = None, 0
        self.optimizer.zero_grad()

    def loss(self, pred, batch):
        """runner.py:193-218; returns (scaled loss, dict of unscaled parts)."""
        depth_full, inter = pred["depth"]["final_depth"], pred["depth"]["intermediate_depths"]
        final_seg = pred["seg"]["final_seg"]
        l_seg = (self.criterion["seg"](final_seg, batch["gt_seg"]) if final_seg is not None else 0) * self.supervised_seg
        l4 = self.criterion["depth"](inter[-1].squeeze(1), batch["gt_s4"].squeeze(1))
        l3 = self.criterion["depth"](inter[-2].squeeze(1), batch["gt_s3"].squeeze(1))
        lf = self.criterion["depth"](depth_full, batch["gt_final"])
        w = LOSS_WEIGHTS
        loss = (w[0] * lf + w[1] * l4 + w[2] * l3 + w[3] * l_seg + w[4] * 0) / sum(w)
        return loss / self.update_interval, {"final": lf, "s4": l4, "s3": l3, "seg": l_seg}

    def __call__(self, batch):
        """One micro-batch; batch: dict(image, gt_final, gt_s4, gt_s3, gt_seg) of CUDA tensors."""
        pred = self.model(batch["image"])
        loss, parts = self.loss(pred, batch)
        with torch.no_grad():
            rmse = torch.sqrt(self.mse(pred["depth"]["final_depth"], batch["gt_final"])) * args.get("max_depth", 100)
            seg = parts["seg"].detach() if torch.is_tensor(parts["seg"]) else torch.zeros_like(rmse)
            s = torch.stack([parts["final"].detach(), parts["s4"].detach(), rmse, seg])
            self._stats = s if self._stats is None else self._stats + s
            self._stats_n += 1
        loss.backward()
        last = self.batches_per_epoch is not None and (self.i + 1) == self.batches_per_epoch
        stepped = False
        if (self.i + 1) % self.update_interval == 0 or last:
            self.training_steps += 1
            self.optimizer.step()
            self.optimizer.zero_grad(set_to_none=True)
            stepped = True
        if self.scheduler is not None and (self.i + 1) > self.update_interval:
            self.scheduler.step()
        self.i += 1
        return loss, stepped

    def stats(self):
        """Means since the last call (ONE device->host read): depth_final, depth_stage4, RMSE [m], seg."""
        if self._stats is None:
            return None
        out = (self._stats / max(1, self._stats_n)).tolist()
        self._stats, self._stats_n = None, 0
        return {"loss_depth_final": out[0], "loss_depth_stage_4": out[1], "RMSE": out[2], "loss_seg": out[3]}


def evaluate(model, batches, update_interval=None, max_depth=None, cutoff=600):
    """`Trainer.eval` (src/main/runner.py:273-350) over an iterable of batch dicts
    (image, gt_final, gt_s4, gt_seg as CUDA tensors) -> (val_loss, RMSE) with the reference's aggregation:
    every `update_interval` batches (and on the last one) a window record [mean final-depth loss, mean stage-4 loss,
    nan-mean of the last `cutoff` per-batch RMSEs, mean seg loss] is appended; the result is the nan-mean of the
    records' first / third column.

    Fixed here: the reference slices the input with `args.num_features`, a key its args never
    define, so its eval loop raises AttributeError on the first batch (runner.py:297); the test loop uses
    `args.input_channels` for the same purpose (:420) and so does this function.  The per-batch `.item()` host
    syncs (:302-309) are gone: the four scalars of every batch stay on the device and are read back once.
    The reference runs this forward under fp16 autocast; the engine's bf16 mode plays that role."""
    import numpy as np
    ui = int(update_interval if update_interval is not None else args.get("update_interval", 1))
    md = float(args.get("max_depth", 100) if max_depth is None else max_depth)
    crit_d, crit_s, mse = MaskedSmoothL1Loss(), MaskedFocalLoss(), MaskedMSELoss()
    was_training = model.training
    model.eval()
    rows = []
    with torch.no_grad():
        for batch in batches:
            x = batch["image"][:, :args.input_channels]
            pred = model(x)
            depth_full, inter = pred["depth"]["final_depth"], pred["depth"]["intermediate_depths"]
            final_seg = pred["seg"]["final_seg"]
            zero = torch.zeros((), dtype=torch.float32, device=depth_full.device)
            l_seg = crit_s(final_seg, batch["gt_seg"]) if final_seg is not None else zero        # runner.py:301
            l4 = crit_d(inter[-1].squeeze(1), batch["gt_s4"].squeeze(1))
            lf = crit_d(depth_full, batch["gt_final"])
            rmse = torch.sqrt(mse(depth_full, batch["gt_final"])) * md
            rows.append(torch.stack([lf, l4, rmse, l_seg.reshape(())]))
    model.train(was_training)
    if not rows:
        return float("nan"), float("nan")
    vals = torch.stack(rows).double().cpu().numpy()             # ONE device -> host read
    n = vals.shape[0]
    eval_losses, rmse_arr, win = [], [], []
    for i in range(n):
        rmse_arr.append(vals[i, 2])
        win.append(vals[i])
        if (i + 1) % ui == 0 or (i + 1) == n:
            w = np.array(win)
            eval_losses.append([np.nanmean(w[:, 0]), np.nanmean(w[:, 1]), np.nanmean(rmse_arr[-cutoff:]),
                                np.nanmean(w[:, 3])])
            win = []
    eval_losses = np.array(eval_losses)
    return float(np.nanmean(eval_losses[:, 0])), float(np.nanmean(eval_losses[:, 2]))


def save_checkpoint(path, model, optimizer, lr, steps):
    """Reference checkpoint layout (runner.py:369-371): {'state_dict','optimizer','lr','steps'}."""
    osd = optimizer.state_dict()
    for st in osd["state"].values():
        st.pop("_slot", None)
        if torch.is_tensor(st.get("exp_grad_norm")):
            st["exp_grad_norm"] = st["exp_grad_norm"].clone()
    torch.save({"state_dict": model.state_dict(), "optimizer": osd, "lr": lr, "steps": steps}, path)
