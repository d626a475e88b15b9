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

import os

import torch

from . import ops


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


class GraphedDataParallelStep(GraphedTrainStep):
    """The reference's training iteration (runner.py:176-232: forward, loss mix (lf + l4 + l3 + 0.2 * lseg) / 3.4 /
    update_interval, backward, optimizer step) driven WITHOUT autograd, as explicit phases of the engine's forward /
    backward programs, so that the step can be cut into several CUDA graphs at the gradient-bucket boundaries:

        graph 0   forward + per-rank loss sums
        (NCCL)    all-reduce of the 11 loss accumulators -> global masked means (SURVEY.md §8e caveat 1)
        graph 1   loss finalize + loss gradients + backward of the heads, the decoder pyramid and encoder stages 4..2
        (NCCL)    all-reduce(AVG) of those buckets of the flat gradient buffer (81 of 88 MB), on NCCL's stream,
        graph 2   backward of encoder stage 1 and the first patch embedding      ... overlapped with this graph
        (NCCL)    all-reduce of the last bucket (runner.py:135-136 replacement)
        graph 3   diffGradNorm step (waits for the buckets)

    The cut points are the gradient-bucket boundaries of the backward program (`CAMRADEPTH_DP_CUTS`, default one cut
    after stage 2: every additional graph boundary costs ~0.1 ms, see profiles/r2_dp_overhead_2gpu.txt).
    With one process (world == 1) all phases are captured into ONE graph.  The public surface is the one of
    `GraphedTrainStep`: `__call__(batch)`, `prefetch(host_batch)`, `run_prefetched()`; the loss returned is the
    global (whole data-parallel batch) loss.  `model` may be a `parallel.DataParallel` wrapper or the bare module.
    """

    TAGS = ("decoder", "stage3", "stage2", "stage1", "stage0")

    @classmethod
    def plan_segments(cls, one_graph, cuts="stage1"):
        """Phases of one step grouped into CUDA graphs.  `cuts`: comma-separated bucket tags after which a new graph
        starts.  The number of cuts does not change the step time measurably (profiles/r2_dp_overhead_2gpu.txt); every
        bucket left to the last backward graph is exposed, so the default cuts once, after encoder stage 2: 81 of the
        88 MB travel while the stage-1 backward (~4 ms) still runs.  The forward always ends its own graph when there
        are several (the loss exchange sits behind it)."""
        if one_graph:
            return [["fwd", "loss"] + list(cls.TAGS[1:]) + ["opt"]]
        cutset = set(filter(None, (cuts or "").split(",")))
        unknown = cutset - set(cls.TAGS[1:])
        if unknown:
            raise ValueError(f"CAMRADEPTH_DP_CUTS: unknown bucket tag(s) {sorted(unknown)}; known: {cls.TAGS[1:]}")
        segments, cur = [["fwd"]], ["loss"]
        for t in cls.TAGS[1:]:
            cur.append(t)
            if t in cutset:
                segments.append(cur)
                cur = []
        if cur:
            segments.append(cur)
        segments.append(["opt"])
        return segments

    def __init__(self, model, optimizer, example_batch, scheduler=None, update_interval=1, process_group=None,
                 global_loss_mean=None, warmup=2):
        import torch.distributed as dist
        from .parallel import bucket_ranges
        self.net = model
        self.model = model.module if hasattr(model, "module") else model
        self.optimizer, self.scheduler = optimizer, scheduler
        if not hasattr(optimizer, "advance_for_replay"):
            raise TypeError("GraphedDataParallelStep needs camradepth_b200.diffGradNorm")
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        if global_loss_mean is None:
            global_loss_mean = getattr(model, "global_loss_mean", True)
        self.global_mean = bool(global_loss_mean) and self.world > 1
        # experiment switches (tools/dp_overhead.py): "noloss" / "nobucket" skip the respective exchange (results are
        # then per-rank, for timing only), "onegraph" additionally captures the whole step into one graph
        self._debug = set(filter(None, os.environ.get("CAMRADEPTH_DP_DEBUG", "").split(",")))
        self.ui = float(update_interval)
        self.static = {k: v.clone() for k, v in example_batch.items()}
        m = self.model
        dev = self.static["image"].device
        self.eng = m._engine_for(self.static["image"])
        eng = self.eng
        self.sup = bool(m.supervised_seg)
        f32 = torch.float32
        self.lacc = torch.zeros(16, dtype=f32, device=dev)        # [l1 final | l1 s4 | l1 s3 | ce] partial sums
        self.lout = torch.zeros(8, dtype=f32, device=dev)
        gs = (self.world if self.global_mean else 1.0) / (3.4 * self.ui)
        self.gout = torch.tensor([gs, 0.2 * gs], dtype=f32, device=dev)
        total = sum(p.numel() for p in m.parameters())
        self.flat = torch.zeros(total, dtype=f32, device=dev)
        self._S = self._outs = self._gen = None
        self._dist = dist
        self.loss = torch.zeros((), dtype=f32, device=dev)

        # ---- eager warm-up (also sizes the arenas, builds the weight-pack table and the optimizer state)
        for _ in range(max(1, warmup)):
            self._eager_step()
        torch.cuda.synchronize()
        self._ranges = {t: bucket_ranges(eng.names, eng.pg_offsets, t) for t in self.TAGS}
        # ---- capture
        segments = self.plan_segments(self.world == 1 or "onegraph" in self._debug,
                                      os.environ.get("CAMRADEPTH_DP_CUTS", "stage1"))
        pool = torch.cuda.graph_pool_handle()
        self.graphs = []
        for seg in segments:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                for ph in seg:
                    self._phase(ph)
            self.graphs.append((g, seg))
        self._S = self._outs = self._gen = None

    # ------------------------------------------------------------------ phases
    def _phase(self, ph):
        eng, b, m = self.eng, self.static, self.model
        if ph == "fwd":
            self._outs, self._S = eng.forward(b["image"], m.training, m._pop_masks(), save=True)
            o = self._outs
            self.lacc.zero_()
            ops.masked_l1_fwd(o["final_depth"], b["gt_final"], self.lacc[0:3])
            ops.masked_l1_fwd(o["inter4"], b["gt_s4"], self.lacc[3:6])
            ops.masked_l1_fwd(o["inter3"], b["gt_s3"], self.lacc[6:9])
            if self.sup:
                ops.ce_fwd(o["final_seg"], b["gt_seg"], self.lacc[9:11])
        elif ph == "loss":
            o = self._outs
            ops.loss_finalize(self.lacc[0:3], self.lout[0:2], 0)
            ops.loss_finalize(self.lacc[3:6], self.lout[2:4], 0)
            ops.loss_finalize(self.lacc[6:9], self.lout[4:6], 0)
            tot = self.lout[0] + self.lout[2] + self.lout[4]
            if self.sup:
                ops.loss_finalize(self.lacc[9:11], self.lout[6:8], 1, 2.0)
                tot = tot + 0.2 * self.lout[6]
            self.loss.copy_(tot / (3.4 * self.ui))
            d_f, d_4, d_3 = (torch.empty_like(o[k]) for k in ("final_depth", "inter4", "inter3"))
            ops.masked_l1_bwd(o["final_depth"], b["gt_final"], self.lacc[0:3], self.gout[0:1], d_f)
            ops.masked_l1_bwd(o["inter4"], b["gt_s4"], self.lacc[3:6], self.gout[0:1], d_4)
            ops.masked_l1_bwd(o["inter3"], b["gt_s3"], self.lacc[6:9], self.gout[0:1], d_3)
            d_s = None
            if self.sup:
                d_s = torch.empty_like(o["final_seg"])
                ops.ce_bwd(o["final_seg"], b["gt_seg"], self.lacc[9:11], self.gout[1:2], 2.0, d_s)
            self.flat.zero_()
            self._gen = eng.backward_steps(self._S, d_f, d_3, d_4, d_s, flat_grad=self.flat)
            self._advance("decoder")
        elif ph == "opt":
            self._bind_grads()
            self.optimizer.step()
        else:
            self._advance(ph)

    def _advance(self, tag):
        for t in self._gen:
            if t == tag:
                break
        if tag == "stage0":
            for _ in self._gen:           # exhaust (releases the saved activations held by the generator)
                pass
            self._S = self._gen = None

    def _bind_grads(self):
        """p.grad = view of the persistent flat buffer (None for parameters without a gradient path, F9)."""
        eng = self.eng
        skip = set(eng.no_grad_names(self.sup))
        for n, p in eng.P.items():
            if n in skip:
                p.grad = None
            elif p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * eng.pg_offsets[n][0]:
                o, sz = eng.pg_offsets[n]
                p.grad = self.flat[o:o + sz].view(p.shape)

    def _allreduce_losses(self):
        if self.global_mean and "noloss" not in self._debug:
            self._dist.all_reduce(self.lacc, op=self._dist.ReduceOp.SUM, group=self.group)

    def _allreduce_bucket(self, tag, works):
        if "nobucket" in self._debug:
            return
        for (a, c) in self._ranges[tag]:
            t = self.flat[a:c]
            if self._dist.get_backend(self.group) == "nccl":
                works.append(self._dist.all_reduce(t, op=self._dist.ReduceOp.AVG, group=self.group, async_op=True))
            else:
                works.append((self._dist.all_reduce(t, group=self.group, async_op=True), t))

    def _wait(self, works):
        for w in works:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.world)
            else:
                w.wait()

    def _eager_step(self):
        from .parallel import bucket_ranges
        works = []
        self._phase("fwd")
        if self.world > 1:
            self._allreduce_losses()
        self._phase("loss")
        self._ranges = {t: bucket_ranges(self.eng.names, self.eng.pg_offsets, t) for t in self.TAGS}
        if self.world > 1:
            self._allreduce_bucket("decoder", works)
        for t in self.TAGS[1:]:
            self._phase(t)
            if self.world > 1:
                self._allreduce_bucket(t, works)
        self._wait(works)
        self._phase("opt")

    def _replay(self):
        self.optimizer.advance_for_replay()
        works = []
        for g, seg in self.graphs:
            if seg[0] == "opt":
                self._wait(works)
            g.replay()
            if self.world > 1:
                for ph in seg:                       # exchanges of everything this graph completed
                    if ph == "fwd":
                        self._allreduce_losses()
                    elif ph == "loss":
                        self._allreduce_bucket("decoder", works)
                    elif ph in self.TAGS:
                        self._allreduce_bucket(ph, works)
        if self.scheduler is not None:
            self.scheduler.step()
        return self.loss
