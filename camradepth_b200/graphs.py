"""Whole-step CUDA-graph capture.

The hot path is ~3000 small-to-medium kernel launches per training step; launched eagerly from Python the
host becomes the bottleneck long before the GPU does.  `GraphedTrainStep` captures forward + losses +
backward + diffGradNorm step (all stream-ordered, no host sync anywhere on the path) into ONE CUDA graph
and replays it; inputs are copied into static device buffers first.  `GraphedInference` does the same for
the eval forward (batch-1 latency).

The optimizer's bias-correction factors and learning rate change every step, so they are not baked into the
graph: the captured update reads its step size from a device scalar, and `GraphedTrainStep` (which owns the
optimizer handle) calls `optimizer.advance_for_replay()` before every replay -- step counters advance, the
scalar is refreshed from `group['lr']` (so LR schedulers keep working) -- and steps the scheduler after it.
"""
from __future__ import annotations

import math

import torch


class GraphedInference:
    def __init__(self, model, example, warmup=3):
        self.model = model
        self.static_in = example.clone()
        model.eval()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                model(self.static_in)
        torch.cuda.current_stream().wait_stream(s)
        # The packed bf16 / K-major weight copies are refreshed INSIDE the graph (one batched launch), so a replay
        # after further training or load_state_dict() never evaluates stale weights.
        self._engines = [e for e in getattr(model, "_engines", {}).values() if e.device is not None]
        for e in self._engines:
            e.prepack()                      # builds the device-side job table (needs a host->device copy)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            for e in self._engines:
                e.prepack()
            self.static_out = model(self.static_in)

    def __call__(self, x):
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedTrainStep:
    """step_fn(batch_dict) must run forward, loss, backward, optimizer.step(), zero_grad(set_to_none=True) and
    return the loss tensor, using only stream-ordered work.

    optimizer: the `diffGradNorm` instance step_fn steps (its host-side bookkeeping runs before every replay), or
    None if step_fn does not contain an optimizer step.  scheduler (optional): stepped once after every replay,
    like `Trainer.train_one_epoch` does per batch (runner.py:269-270)."""

    def __init__(self, step_fn, example_batch, warmup=3, optimizer=None, scheduler=None):
        self.optimizer, self.scheduler = optimizer, scheduler
        if optimizer is not None and not hasattr(optimizer, "advance_for_replay"):
            raise TypeError("GraphedTrainStep needs an optimizer with advance_for_replay() (camradepth_b200.diffGradNorm)")
        self.static = {k: v.clone() for k, v in example_batch.items()}
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                step_fn(self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(self.static)

    def _replay(self):
        if self.optimizer is not None:
            self.optimizer.advance_for_replay()
        self.graph.replay()
        if self.scheduler is not None:
            self.scheduler.step()
        return self.loss

    def __call__(self, batch=None):
        if batch is not None:
            for k, v in batch.items():
                self.static[k].copy_(v, non_blocking=True)
        return self._replay()

    # ---- double-buffered input feed: the host->device copy of the NEXT batch runs on a copy stream while the
    # graph of the current step executes; the step itself starts with a device-to-device copy into the static
    # inputs (the reference hides its H2D copies behind DataLoader workers + .to(device), runner.py:176-190)
    def prefetch(self, host_batch):
        """Start copying a (pinned) host batch to the device; returns immediately."""
        if not hasattr(self, "_stage"):
            self._stage = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._copy_stream = torch.cuda.Stream()
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        self._copy_stream.wait_event(self._consumed)          # the previous staging contents have been taken
        with torch.cuda.stream(self._copy_stream):
            for k, v in host_batch.items():
                self._stage[k].copy_(v, non_blocking=True)
            self._staged.record()

    def wait_prefetch(self):
        """Make the current stream wait for the copy started by the last prefetch()."""
        torch.cuda.current_stream().wait_event(self._staged)

    def run_prefetched(self):
        """Run one step on the batch passed to the last prefetch()."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        for k, v in self._stage.items():
            self.static[k].copy_(v, non_blocking=True)
        self._consumed.record()
        return self._replay()
