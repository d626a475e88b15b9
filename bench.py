#!/usr/bin/env python
"""Headline benchmark: CamRaDepth base bf16 training step (forward + 3 masked losses + backward +
diffGradNorm), batch 32 per GPU at 192x416 (the runnable stand-in for the nominal 192x400),
synthetic RGB+radar data, random-init weights.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in HBM; `e2e` = the same
step through the public API with HOST (pinned) inputs, H2D copies and a D2H loss read inside the timed region.
`--impl reference` times the CPU oracle port of the reference on the host cores (the reference itself is a
Python package that does not exist on the GPU box).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W = 192, 416
FLOPS_FWD_PER_SAMPLE = {"base": 150.5e9}          # SURVEY.md §8(d), conv/GEMM 2*MACs as written in the reference
FLOPS_TRAIN_PER_SAMPLE = {"base": 451.6e9}
DOMINANT = "depth_upsample.4.conv.layers.2.model.0.weight"   # 3x3, Cin 296 -> 128 at full resolution
DOMINANT_FLOPS_PER_SAMPLE = 2.0 * 192 * 416 * 128 * 9 * 296  # as written in the reference (54.5 GF)
# dram__bytes_read.sum + dram__bytes_write.sum of that launch at batch 32, from the committed `ncu --set full`
# capture profiles/r1_ncu_full_summary.md (algorithmic bytes: 2.168e9)
DOMINANT_TRAFFIC_B32 = 2.141e9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons every 200 ms during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap", 0x80: "hw_power_brake"}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.2)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons)}


def run_reference(a):
    """CPU arm: the oracle port of the reference (fp32, eager torch CPU ops) on all host cores; one step =
    forward + losses + backward + diffGradNorm on ONE sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import camradepth_oracle as O
    from camradepth_b200.synthetic import make_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.Cfg(a.variant)
    sd = {k: v.requires_grad_(True) for k, v in O.init_state_dict(cfg, seed=0).items()}
    states = {k: {} for k in sd}
    bs = 1
    batch = make_batch(bs, H, W, seed=0)

    def step():
        dps, d2s = O.make_masks(cfg, bs, seed=1)
        pred = O.forward(sd, cfg, batch["image"], dps, d2s)
        loss, _ = O.training_loss(pred, batch["gt_final"], batch["gt_s4"], batch["gt_s3"], batch["gt_seg"], cfg)
        loss.backward()
        with torch.no_grad():
            for k, p in sd.items():
                if p.grad is not None:
                    O.diffgradnorm_step(p, p.grad, states[k], lr=6e-5)
                    p.grad = None
        return float(loss)

    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    v = bs * a.steps / dt
    sample = f"{a.steps} steps of 1 sample (of the batch-32 workload), fp32, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "train samples/s", "value": v, "unit": "samples/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CamRaDepth {a.variant} training step, 192x416 (nominal 192x400), batch 1 sample per step on CPU"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline(variant):
    from oracle import camradepth_oracle as O
    from camradepth_b200.synthetic import make_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.Cfg(variant)
    sd = {k: v.requires_grad_(True) for k, v in O.init_state_dict(cfg, seed=0).items()}
    states = {k: {} for k in sd}
    batch = make_batch(1, H, W, seed=0)
    times = []
    for it in range(4):
        t0 = time.perf_counter()
        pred = O.forward(sd, cfg, batch["image"])
        loss, _ = O.training_loss(pred, batch["gt_final"], batch["gt_s4"], batch["gt_s3"], batch["gt_seg"], cfg)
        loss.backward()
        with torch.no_grad():
            for k, p in sd.items():
                if p.grad is not None:
                    O.diffgradnorm_step(p, p.grad, states[k], lr=6e-5)
                    p.grad = None
        times.append(time.perf_counter() - t0)
    t = sorted(times[1:])[1]
    return {"value": 1.0 / t, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "3 timed train steps (fwd+losses+bwd+diffGradNorm) of 1 sample at 192x416, fp32, median, after 1 warm-up"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch")
    ap.add_argument("--variant", default="base")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--quick", action="store_true", help="resident timing only (for ncu launch lists): no e2e / roofline passes")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    if not a.quick:
        a.warmup = max(a.warmup, 3)          # timing rule: at least three warm-up steps (ncu launch lists excepted)

    import torch.distributed as dist
    import camradepth_b200 as C
    from camradepth_b200 import ops
    from camradepth_b200.parallel import DataParallel
    from camradepth_b200.synthetic import make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    C.set_model(a.variant)
    torch.manual_seed(0)
    model = C.CamRaDepth(precision=a.precision).to(dev)
    net = DataParallel(model) if world > 1 else model
    model.train()
    crit_d, crit_s = C.MaskedSmoothL1Loss(), C.MaskedFocalLoss()
    opt = C.diffGradNorm(model.parameters(), lr=6e-5)
    B = a.batch
    host = make_batch(B, H, W, seed=100 + rank, input_channels=C.args.input_channels, pin=True)
    devb = {k: v.to(dev) for k, v in host.items()}
    loss_host = torch.zeros(1).pin_memory()

    def fwd_bwd(b):
        pred = net(b["image"])
        inter = pred["depth"]["intermediate_depths"]
        lf = crit_d(pred["depth"]["final_depth"], b["gt_final"])
        l4 = crit_d(inter[-1], b["gt_s4"])
        l3 = crit_d(inter[-2], b["gt_s3"])
        fs = pred["seg"]["final_seg"]
        ls = crit_s(fs, b["gt_seg"]) if fs is not None else 0
        loss = (lf + l4 + l3 + 0.2 * ls) / 3.4
        loss.backward()
        return loss

    def opt_step():
        opt.step()
        opt.zero_grad(set_to_none=True)

    def step(b):
        loss = fwd_bwd(b)
        opt_step()
        return loss

    def e2e_step():
        b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss = step(b)
        loss_host.copy_(loss.detach().view(1), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from camradepth_b200.graphs import GraphedTrainStep
    use_graph = not a.no_graph
    for _ in range(a.warmup):
        step(devb)
    n0 = ops.launch_count()
    step(devb)
    launches_per_step = ops.launch_count() - n0
    eng = model._engines[a.precision]
    if use_graph and world == 1:
        # the whole step (fwd + losses + bwd + optimizer) is captured once and replayed; the optimizer's
        # host-side bookkeeping (step counters, device-resident step size) runs before each replay
        gstep = GraphedTrainStep(step, devb, warmup=0)

        def run_resident():
            opt.advance_for_replay()
            return gstep()

        def run_e2e():
            # double-buffered feed: consume the batch whose H2D copy was started during the previous step, start
            # the copy of the next one (copy stream), read the loss back.  One H2D copy of a full pinned host
            # batch and one D2H loss read are issued per step inside the timed region.
            opt.advance_for_replay()
            loss = gstep.run_prefetched()
            gstep.prefetch(host)
            loss_host.copy_(loss.detach().view(1), non_blocking=True)

        gstep.prefetch(host)
    elif use_graph:
        # N > 1: graph A = fwd + losses + bwd, then ONE NCCL all-reduce of the flat gradient buffer (launched
        # eagerly: capturing NCCL work inside the graph hangs with this torch/NCCL pair), then graph B = optimizer.
        # The all-reduce is ~0.3 % of the step, so not overlapping it costs less than eager launch overhead.
        net.require_backward_grad_sync = False
        fwd_bwd(devb)                                # p.grad now lives in the engine's persistent flat buffer
        dist.all_reduce(eng._flat_own, op=dist.ReduceOp.AVG)
        opt_step()
        g_fb = GraphedTrainStep(fwd_bwd, devb, warmup=0)
        torch.cuda.synchronize()
        g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_opt):
            opt_step()

        def _replay(prefetched):
            opt.advance_for_replay()
            loss = g_fb.run_prefetched() if prefetched else g_fb(None)
            dist.all_reduce(eng._flat_own, op=dist.ReduceOp.AVG)
            g_opt.replay()
            return loss

        def run_resident():
            return _replay(False)

        def run_e2e():
            loss = _replay(True)                     # same double-buffered feed as the single-GPU arm
            g_fb.prefetch(host)
            loss_host.copy_(loss.detach().view(1), non_blocking=True)

        g_fb.prefetch(host)
    else:
        def run_resident():
            return step(devb)
        run_e2e = e2e_step
    for _ in range(a.warmup):
        run_resident()
    # ---- device-resident timing
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        run_resident()
    e1.record()
    barrier()
    sampler.stop_flag = True
    launches = launches_per_step * a.steps
    ms = e0.elapsed_time(e1)
    if a.quick:
        if rank == 0:
            print(json.dumps({"metric": "train samples/s", "value": world * B * a.steps / (ms / 1e3),
                              "ms_per_step": ms / a.steps, "quick": True}))
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- end-to-end timing (host buffers)
    for _ in range(2):
        run_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(a.steps):
        run_e2e()
    if use_graph:
        # the copy started during the last step belongs to the timed region too: K full H2D copies for K steps
        (gstep if world == 1 else g_fb).wait_prefetch()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    # ---- dominant kernel, timed live with CUDA events on its launch stream over the same steps (eager launches:
    # events cannot be placed inside a replayed graph)
    eng.timed = {("fwd", DOMINANT): []}
    for _ in range(a.steps):
        step(devb)
    torch.cuda.synchronize()
    dom = [x.elapsed_time(y) for (x, y) in eng.timed[("fwd", DOMINANT)]]
    eng.timed = None
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    sampler.join(timeout=2)
    if rank == 0:
        pk, pk_src = peaks()
        value = world * B * a.steps / (ms / 1e3)
        e2e = world * B * a.steps / (ms_e2e / 1e3)
        dom_ms = sum(dom) / max(1, len(dom))
        dom_tflops = DOMINANT_FLOPS_PER_SAMPLE * B / (dom_ms * 1e-3) / 1e12 if dom else None
        peak = pk["bf16_tflops_sustained"]
        h2d = world * sum(v.numel() * v.element_size() for v in host.values())      # whole job, like `value`
        out = {
            "metric": "train samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
            "config": {"workload": f"CamRaDepth {a.variant} {a.precision} training step (fwd + masked SmoothL1 x3 + bwd + "
                                   f"diffGradNorm, DropPath/Dropout2d on), batch {B}/GPU, 192x416 (nominal 192x400: the "
                                   f"reference cannot run 400-wide inputs), RGB+radar 7ch",
                       "global_batch": world * B, "parallelism": f"dp{world}",
                       "launch": ("one CUDA graph per step" if world == 1 else "two CUDA graphs per step around one NCCL all-reduce")
                       if use_graph else "eager kernel launches",
                       "e2e_feed": ("pinned host batch -> device staging buffers on a copy stream, overlapped with the "
                                    "previous step; D2D into the graph's static inputs; D2H loss read every step"
                                    if use_graph else "H2D copies on the compute stream every step"),
                       "l2": "no flush: per-step working set (several GB of activations) is far larger than the 126 MB L2",
                       "step_flops": FLOPS_TRAIN_PER_SAMPLE.get(a.variant, 0) * B,
                       "step_tensor_frac_of_" + pk_src: (value / world) * FLOPS_TRAIN_PER_SAMPLE.get(a.variant, 0) / 1e12 / peak},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * world},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": "conv_tc_halo_kernel: tcgen05 implicit-GEMM 3x3 conv fwd, depth_upsample[4] layer 2 "
                                   "(M=B*79872 pixels, N=128, K=9*296)",
                         "achieved": dom_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": (dom_tflops / peak) if dom_tflops else None,
                         "traffic": DOMINANT_TRAFFIC_B32 * B / 32.0,
                         "algorithmic_bytes": (192 * 416 * (296 + 128) * 2.0) * B + 128 * 9 * 296 * 2.0,
                         "peak_source": pk_src + " (bf16_tflops_sustained)", "launch_ms": dom_ms,
                         "launches_timed": len(dom)},
        }
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(a.variant)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
