#!/usr/bin/env python
"""Headline benchmark: CamRaDepth base bf16 training step (forward + 3 masked losses + backward +
diffGradNorm), batch 32 per GPU at 192x416 (the runnable stand-in for the nominal 192x400),
synthetic RGB+radar data, random-init weights.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in HBM; `e2e` = the same
step through the public API (`graphs.GraphedDataParallelStep`) with HOST (pinned) inputs, H2D copies and a D2H
loss read inside the timed region.  `configs` carries the other BASELINE.json configurations measured in the same
run (supervised_seg training at this N; on one GPU also the inference batch sweep, the full-resolution latencies
and the eager ATen/cuDNN execution of the same step on this GPU).  `--impl reference` times the CPU oracle port
of the reference on the host cores (the reference itself is a Python package that does not exist on the GPU box).
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
# conv / GEMM 2*MACs as written in the reference, forward + data gradient + weight gradient
FLOPS_TRAIN_PER_SAMPLE = {"base": 451.6e9, "supervised_seg": 816.8e9, "unsupervised_seg": 573.4e9}
FLOPS_FWD_PER_SAMPLE = {"base": 150.5e9}
DOMINANT = "depth_upsample.4.conv.layers.2.model.0.weight"   # 3x3, Cin 296 -> 128 at full resolution
DOMINANT_FLOPS_PER_SAMPLE = 2.0 * 192 * 416 * 128 * 9 * 296  # as written in the reference (54.5 GF)
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch come from a TRACKED export of an
# `ncu --set full` capture (tools/ncu_export.py writes it; the raw-page CSV it was parsed from sits next to it)
DOMINANT_NCU = os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def dominant_traffic(batch):
    """-> (bytes per launch at `batch` samples or None, source string)"""
    try:
        d = json.load(open(DOMINANT_NCU))
        per_sample = (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["batch"]
        return per_sample * batch, os.path.relpath(DOMINANT_NCU, ROOT) + " <- " + d.get("source", "?")
    except Exception:
        return None, None


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


# ---------------------------------------------------------------------------------------------- CPU arms
def _oracle_cpu_step(variant):
    from oracle import camradepth_oracle as O
    from camradepth_b200.synthetic import make_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.Cfg(variant)
    sd = {k: v.requires_grad_(True) for k, v in O.init_state_dict(cfg, seed=0).items()}
    states = {k: {} for k in sd}
    batch = make_batch(1, H, W, seed=0)

    def step():
        dps, d2s = O.make_masks(cfg, 1, seed=1)
        pred = O.forward(sd, cfg, batch["image"], dps, d2s)
        loss, _ = O.training_loss(pred, batch["gt_final"], batch["gt_s4"], batch["gt_s3"], batch["gt_seg"], cfg)
        loss.backward()
        with torch.no_grad():
            for k, p in sd.items():
                if p.grad is not None:
                    O.diffgradnorm_step(p, p.grad, states[k], lr=6e-5)
                    p.grad = None
        return float(loss)
    return step


def run_reference(a):
    """CPU arm: the oracle port of the reference (fp32, eager torch CPU ops) on all host cores; one step =
    forward + losses + backward + diffGradNorm on ONE sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step = _oracle_cpu_step(a.variant)
    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    v = a.steps / dt
    sample = f"{a.steps} steps of 1 sample (of the batch-32 workload), fp32, {torch.get_num_threads()} threads"
    print(json.dumps({
        "impl": "reference", "metric": "train samples/s", "value": v, "unit": "samples/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CamRaDepth {a.variant} training step, 192x416 (nominal 192x400), batch 1 sample per step on CPU"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline(variant):
    step = _oracle_cpu_step(variant)
    times = []
    for _ in range(4):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    t = sorted(times[1:])[1]
    return {"value": 1.0 / t, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "3 timed train steps (fwd+losses+bwd+diffGradNorm) of 1 sample at 192x416, fp32, median, after 1 warm-up"}


def eager_gpu_baseline(variant, B, dev):
    """The eager ATen/cuDNN execution model of the reference on THIS GPU (SURVEY.md §8d "the real bar"): the oracle
    restatement moved to cuda:0 under bf16 autocast, same step minus the optimizer (which favours this arm: the
    reference's diffGradNorm adds ~13k launches and 881 host syncs).  Checker-side code, timed as a baseline."""
    from oracle import camradepth_oracle as O
    from camradepth_b200.synthetic import make_batch
    cfg = O.Cfg(variant)
    sd = {k: v.to(dev).requires_grad_(True) for k, v in O.init_state_dict(cfg, seed=0).items()}
    batch = {k: v.to(dev) for k, v in make_batch(B, H, W, seed=0, input_channels=cfg.cin).items()}
    dps, d2s = O.make_masks(cfg, B, seed=1)
    dps, d2s = [t.to(dev) for t in dps], [t.to(dev) for t in d2s]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        ts = []
        for _ in range(4):
            for v in sd.values():
                v.grad = None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                pred = O.forward(sd, cfg, batch["image"], dps, d2s)
            pf = {"depth": {"final_depth": pred["depth"]["final_depth"].float(),
                            "intermediate_depths": tuple(None if t is None else t.float()
                                                         for t in pred["depth"]["intermediate_depths"])},
                  "seg": pred["seg"]}
            loss, _ = O.training_loss(pf, batch["gt_final"], batch["gt_s4"], batch["gt_s3"], batch["gt_seg"], cfg)
            loss.backward()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ms = 1e3 * sorted(ts[1:])[1]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    del sd, batch, pred, pf, loss
    torch.cuda.empty_cache()
    return {"value": B * 1e3 / ms, "unit": "samples/s", "ms_per_step": ms, "kind": "oracle port on cuda:0, eager ATen/cuDNN, bf16 autocast",
            "sample": f"fwd + losses + bwd (no optimizer) of batch {B} at 192x416, median of 3 after 1 warm-up"}


# ---------------------------------------------------------------------------------------------- GPU arm
def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def timed_region(fn, steps, world, after=None):
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    if after is not None:
        after()
    e1.record()
    barrier(world)
    return e0.elapsed_time(e1)


def build_train(variant, precision, B, dev, rank, world, use_graph):
    """-> dict with the model, the graphed step and its feeds (public API objects only)."""
    import camradepth_b200 as C
    from camradepth_b200.graphs import GraphedDataParallelStep
    from camradepth_b200.parallel import DataParallel
    from camradepth_b200.synthetic import make_batch
    C.set_model(variant)
    torch.manual_seed(0)
    model = C.CamRaDepth(precision=precision).to(dev)
    net = DataParallel(model) if world > 1 else model
    model.train()
    opt = C.diffGradNorm(model.parameters(), lr=6e-5)
    host = make_batch(B, H, W, seed=100 + rank, input_channels=C.args.input_channels, pin=True)
    devb = {k: v.to(dev) for k, v in host.items()}
    out = dict(model=model, net=net, opt=opt, host=host, devb=devb)
    if use_graph:
        out["gstep"] = GraphedDataParallelStep(net, opt, devb, warmup=2)
    else:
        ts = C.TrainStep(net, opt, update_interval=1)
        out["eager"] = lambda b: ts(b)[0]
    return out


def measure_train(variant, a, dev, rank, world, local, headline):
    from camradepth_b200 import ops
    B = a.batch
    use_graph = not a.no_graph
    T = build_train(variant, a.precision, B, dev, rank, world, use_graph)
    host, devb = T["host"], T["devb"]
    loss_host = torch.zeros(1).pin_memory()
    if use_graph:
        g = T["gstep"]

        def run_resident():
            return g()                       # static inputs already hold the batch

        def run_e2e():
            # double-buffered feed: consume the batch whose H2D copy was started during the previous step, start
            # the copy of the next one (copy stream), read the loss back.  One H2D copy of a full pinned host
            # batch and one D2H loss read are issued per step inside the timed region.
            loss = g.run_prefetched()
            g.prefetch(host)
            loss_host.copy_(loss.detach().view(1), non_blocking=True)
        after = g.wait_prefetch
    else:
        step = T["eager"]

        def run_resident():
            return step(devb)

        def run_e2e():
            b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            loss_host.copy_(step(b).detach().view(1), non_blocking=True)
        after = None
    steps = a.steps if headline else max(3, min(5, a.steps))
    for _ in range(a.warmup if headline else 3):
        run_resident()
    sampler = ClockSampler(local) if headline else None
    barrier(world)
    if sampler:
        sampler.start()
    # CAMRADEPTH_PROFILE_TIMED=1 (with `ncu --profile-from-start off`): only the timed steps are instrumented, so a
    # launch list of one step costs one step's worth of replays instead of the warm-up's as well
    prof = os.environ.get("CAMRADEPTH_PROFILE_TIMED", "0") == "1"
    if prof:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    ms = timed_region(run_resident, steps, world)
    if prof:
        torch.cuda.cudart().cudaProfilerStop()
    if sampler:
        sampler.stop_flag = True
    res = {"ms": ms, "steps": steps, "T": T, "sampler": sampler, "graphs": len(g.graphs) if use_graph else 0}
    if a.quick or not headline:
        return res
    # ---- end-to-end timing (host buffers)
    if use_graph:
        g.prefetch(host)
    for _ in range(2):
        run_e2e()
    res["ms_e2e"] = timed_region(run_e2e, steps, world, after)
    # ---- launches per step and the dominant kernel, timed live with CUDA events on its launch stream over the same
    # number of steps (eager launches through the autograd path: events cannot be placed inside a replayed graph)
    import camradepth_b200 as C
    model, opt = T["model"], T["opt"]
    eng = model._engines[a.precision]
    ts = C.TrainStep(T["net"], opt, update_interval=1)
    ts(devb)
    n0 = ops.launch_count()
    ts(devb)
    res["launches_per_step"] = ops.launch_count() - n0
    eng.timed = {("fwd", DOMINANT): []}
    for _ in range(steps):
        ts(devb)
    torch.cuda.synchronize()
    res["dom"] = [x.elapsed_time(y) for (x, y) in eng.timed[("fwd", DOMINANT)]]
    eng.timed = None
    return res


def measure_inference(dev, sweep, shapes):
    """BASELINE config 5: eval forward (bf16, CUDA graph) ms/img over the batch sweep at 192x416 and batch-1
    latency at 416x800 / 896x1600 (the reference cannot run 900x1600: H, W must be multiples of 32)."""
    import camradepth_b200 as C
    from camradepth_b200.graphs import GraphedInference
    from camradepth_b200.synthetic import make_batch
    C.set_model("base")
    torch.manual_seed(0)
    model = C.CamRaDepth(precision="bf16").to(dev).eval()
    out = {}
    for (B, h, w) in [(b, H, W) for b in sweep] + [(1, h, w) for (h, w) in shapes]:
        x = make_batch(B, h, w, seed=2)["image"].to(dev)
        g = GraphedInference(model, x, warmup=2)
        n = 20 if B <= 8 else (6 if B <= 64 else 3)
        for _ in range(2):
            g(x)
        ms = timed_region(lambda: g(x), n, 1) / n
        key = f"b{B}" if (h, w) == (H, W) else f"b{B}_{h}x{w}"
        out[key] = round(ms / B, 4)
        del g, x
        torch.cuda.empty_cache()
    del model
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations (`configs`)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--quick", action="store_true", help="resident timing only (for ncu launch lists): no e2e / roofline passes")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    if not a.quick:
        a.warmup = max(a.warmup, 3)          # timing rule: at least three warm-up steps (ncu launch lists excepted)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = a.batch
    use_graph = not a.no_graph

    res = measure_train(a.variant, a, dev, rank, world, local, headline=True)
    ms, steps = res["ms"], res["steps"]
    if a.quick:
        if rank == 0:
            print(json.dumps({"metric": "train samples/s", "value": world * B * steps / (ms / 1e3),
                              "ms_per_step": ms / steps, "quick": True}))
        if world > 1:
            dist.destroy_process_group()
        return
    t = torch.tensor([ms, res["ms_e2e"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    host = res["T"]["host"]
    h2d = world * sum(v.numel() * v.element_size() for v in host.values())      # whole job, like `value`
    n_graphs = res["graphs"]
    sampler = res["sampler"]
    sampler.join(timeout=2)
    dom = res["dom"]
    launches = res["launches_per_step"] * steps
    del res
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations, measured in the same run
    configs = {}
    if not a.no_configs:
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            r3 = measure_train("supervised_seg", a, dev, rank, world, local, headline=False)
            t3 = torch.tensor([r3["ms"]], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            v3 = world * B * r3["steps"] / (float(t3[0]) / 1e3)
            configs["supervised_seg_train"] = {
                "value": v3, "unit": "samples/s", "ms_per_step": float(t3[0]) / r3["steps"], "steps": r3["steps"],
                "n_gpus": world, "batch_per_gpu": B,
                "tensor_frac": (v3 / world) * FLOPS_TRAIN_PER_SAMPLE["supervised_seg"] / 1e12 / peaks()[0]["bf16_tflops_sustained"],
                "note": "BASELINE config 3: supervised semantic-segmentation branch (+ focal CE loss), bf16, data-parallel, "
                        "global masked-mean losses, gradient buckets all-reduced while the next stage's backward runs"}
            del r3
        except Exception as e:  # pragma: no cover
            configs["supervised_seg_train"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        gc.collect()
        torch.cuda.empty_cache()
        if world == 1:
            try:
                configs["inference_ms_per_img"] = measure_inference(dev, (1, 2, 4, 8, 16, 32, 64, 128, 256),
                                                                    ((416, 800), (896, 1600)))
                configs["inference_note"] = ("BASELINE config 5: base, eval, bf16, CUDA-graph replay, 192x416 batch sweep; "
                                             "b1_416x800 / b1_896x1600 = batch-1 latency at the native / full resolution")
            except Exception as e:  # pragma: no cover
                configs["inference_ms_per_img"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            try:
                configs["eager_gpu_baseline"] = eager_gpu_baseline(a.variant, B, dev)
            except Exception as e:  # pragma: no cover
                configs["eager_gpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        pk, pk_src = peaks()
        value = world * B * steps / (ms / 1e3)
        e2e = world * B * steps / (ms_e2e / 1e3)
        dom_ms = sum(dom) / max(1, len(dom))
        dom_tflops = DOMINANT_FLOPS_PER_SAMPLE * B / (dom_ms * 1e-3) / 1e12 if dom else None
        peak = pk["bf16_tflops_sustained"]
        traffic, traffic_src = dominant_traffic(B)
        if not use_graph:
            launch = "eager kernel launches"
        elif world == 1:
            launch = "one CUDA graph per step"
        else:
            launch = (f"{n_graphs} CUDA graphs per step (forward | losses + backward down to encoder stage 2 | rest of the "
                      "backward | optimizer); the NCCL all-reduce of the gradient buckets a graph completed runs on NCCL's "
                      "stream while the next graph executes; loss sums all-reduced for global masked means")
        out = {
            "metric": "train samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps,
            "warmup": a.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
            "config": {"workload": f"CamRaDepth {a.variant} {a.precision} training step (fwd + masked SmoothL1 x3 + bwd + "
                                   f"diffGradNorm, DropPath/Dropout2d on), batch {B}/GPU, 192x416 (nominal 192x400: the "
                                   f"reference cannot run 400-wide inputs), RGB+radar 7ch",
                       "global_batch": world * B, "parallelism": f"dp{world}", "launch": launch,
                       "e2e_feed": ("pinned host batch -> device staging buffers on a copy stream, overlapped with the "
                                    "previous step; D2D into the graph's static inputs; D2H loss read every step"
                                    if use_graph else "H2D copies on the compute stream every step"),
                       "l2": "no flush: per-step working set (several GB of activations) is far larger than the 126 MB L2",
                       "step_flops": FLOPS_TRAIN_PER_SAMPLE.get(a.variant, 0) * B,
                       "step_tensor_frac_of_" + pk_src: (value / world) * FLOPS_TRAIN_PER_SAMPLE.get(a.variant, 0) / 1e12 / peak},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * world},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": "conv_tc_halo_kernel: tcgen05 implicit-GEMM 3x3 conv fwd with GroupNorm sums in the "
                                   "read-out, depth_upsample[4] layer 2 (M=B*79872 pixels, N=128, K=9*296)",
                         "achieved": dom_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": (dom_tflops / peak) if dom_tflops else None,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": (192 * 416 * (296 + 128) * 2.0) * B + 128 * 9 * 296 * 2.0,
                         "peak_source": pk_src + " (bf16_tflops_sustained)", "launch_ms": dom_ms,
                         "launches_timed": len(dom)},
            "configs": configs,
        }
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(a.variant)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
