"""Data-parallel training across the GPUs of one box: one process per GPU, replicas kept in sync by
bucketed NCCL all-reduce of the flat fp32 gradient buffer, issued from inside the backward program as
soon as a bucket's gradients are final (overlapped with the remaining backward kernels).

Replaces the reference's single-process `nn.DataParallel(model)` (runner.py:135-136), which re-broadcasts
all parameters every iteration and reduces gradients onto GPU 0.  The wrapper keeps
the `.module` attribute and the 'module.'-prefixed state_dict of nn.DataParallel so checkpoints stay
interchangeable.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn


def bucket_ranges(names, offsets, tag):
    """Contiguous [start, end) ranges of the flat gradient buffer whose gradients are final once the
    backward program reports `tag` (pure host logic; unit-tested on CPU)."""
    def span(pred):
        sel = [offsets[n] for n in names if pred(n)]
        if not sel:
            return None
        return (min(o for o, _ in sel), max(o + s for o, s in sel))

    if tag == "decoder":          # decoder pyramid + every head (registered after the encoder)
        r = span(lambda n: not n.startswith("dest_encoder."))
        return [r] if r else []
    if tag in ("stage3", "stage2", "stage1"):
        r = span(lambda n: n.startswith(f"dest_encoder.block{int(tag[-1]) + 1}."))
        return [r] if r else []
    if tag == "stage0":           # first stage + all patch embeddings (contiguous at the start of the buffer)
        r = span(lambda n: n.startswith("dest_encoder.block1.") or n.startswith("dest_encoder.patch_embed"))
        return [r] if r else []
    return []


class DataParallel(nn.Module):
    """global_loss_mean=True (default): the camradepth_b200 losses become masked means over the valid pixels of the
    WHOLE data-parallel batch, like the reference's single-process nn.DataParallel computes them on the gathered
    outputs (runner.py:193-203); False keeps per-rank means (the torch DDP convention)."""

    def __init__(self, module, process_group=None, global_loss_mean=True):
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._works = []
        self.require_backward_grad_sync = True
        self.global_loss_mean = bool(global_loss_mean)
        if hasattr(module, "_engines"):              # a camradepth_b200.CamRaDepth (the CPU unit tests wrap a stand-in)
            from . import losses
            losses.set_data_parallel(self.world if self.global_loss_mean else 1, process_group)
        if self.world > 1:
            with torch.no_grad():
                for p in module.parameters():          # one initial broadcast (not one per step)
                    dist.broadcast(p.data, src=0, group=process_group)
            module._grad_bucket_hook = self._on_bucket
            module._post_backward_hook = self._finish

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def _reduce(self, t):
        backend = dist.get_backend(self.process_group)
        if backend == "nccl":
            self._works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.process_group, async_op=True))
        else:
            w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
            self._works.append((w, t))

    def _on_bucket(self, eng, tag):
        if self.world == 1 or not self.require_backward_grad_sync:
            return
        for (a, b) in bucket_ranges(eng.names, eng.pg_offsets, tag):
            self._reduce(eng.flat_grad[a:b])

    def _finish(self, eng, grads):
        for w in self._works:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.world)
            else:
                w.wait()               # stream-level wait on the NCCL stream; does not block the host
        self._works = []
