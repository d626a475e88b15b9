"""BASELINE.json configs 3-5 on one GPU (config 2 is bench.py's headline line):
  * training samples/s of the supervised / unsupervised / sup+unsup segmentation variants (batch 32, 192x416);
  * inference (eval, no_grad, bf16) ms/img for batch 1..256 at 192x416 and batch 1 at 416x800 / 896x1600
    (the reference cannot run 192x400 / 900x1600: H, W must be multiples of 32), CUDA-graph replay.
Prints one JSON line per measurement."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import camradepth_b200 as C  # noqa: E402
from camradepth_b200.graphs import GraphedInference, GraphedTrainStep  # noqa: E402
from camradepth_b200.synthetic import make_batch  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def train_variant(variant, B=32, steps=5):
    C.set_model(variant)
    torch.manual_seed(0)
    model = C.CamRaDepth(input_channels=C.args.input_channels, precision="bf16").to(dev).train()
    opt = C.diffGradNorm(model.parameters(), lr=6e-5)
    ts = C.TrainStep(model, opt, None, update_interval=1)
    b = {k: v.to(dev) for k, v in make_batch(B, 192, 416, seed=1, input_channels=C.args.input_channels).items()}

    def step(bb):
        return ts(bb)[0]

    for _ in range(3):
        step(b)
    g = GraphedTrainStep(step, b, warmup=0)

    def run():
        opt.advance_for_replay()
        g()
    ms = timed(run, steps)
    print(json.dumps({"config": f"train {variant} bf16 batch {B} 192x416, 1 GPU, CUDA graph", "ms_per_step": ms,
                      "samples_per_s": B * 1e3 / ms}), flush=True)
    del g, model, opt
    torch.cuda.empty_cache()


def infer(B, H, W, steps=10):
    C.set_model("base")
    torch.manual_seed(0)
    model = C.CamRaDepth(precision="bf16").to(dev).eval()
    x = make_batch(B, H, W, seed=2)["image"].to(dev)
    g = GraphedInference(model, x)
    ms = timed(lambda: g(x), steps)
    with torch.no_grad():
        ms_eager = timed(lambda: model(x), max(2, steps // 3))
    print(json.dumps({"config": f"inference base bf16 batch {B} {H}x{W}", "ms_per_batch_graph": ms,
                      "ms_per_img_graph": ms / B, "ms_per_img_eager": ms_eager / B, "img_per_s": B * 1e3 / ms}),
          flush=True)
    del g, model
    torch.cuda.empty_cache()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "infer"):
        for B in (1, 2, 4, 8, 16, 32, 64, 128, 256):
            infer(B, 192, 416, steps=10 if B <= 32 else 4)
        infer(1, 416, 800)
        infer(1, 896, 1600)
    if what in ("all", "train"):
        for v in ("supervised_seg", "unsupervised_seg", "sup_unsup_seg"):
            train_variant(v)
