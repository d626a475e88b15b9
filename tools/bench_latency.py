"""Eval-forward latency (bf16, CUDA-graph replay) at a few batch sizes; one JSON line.  Used for A/B runs of engine
switches (environment variables), e.g. CAMRADEPTH_PDL=0 python tools/bench_latency.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import camradepth_b200 as C  # noqa: E402
from camradepth_b200.graphs import GraphedInference  # noqa: E402
from camradepth_b200.synthetic import make_batch  # noqa: E402

dev = torch.device("cuda:0")
C.set_model("base")
torch.manual_seed(0)
model = C.CamRaDepth(precision="bf16").to(dev).eval()
out = {}
for B in (1, 8, 32):
    x = make_batch(B, 192, 416, seed=2)["image"].to(dev)
    g = GraphedInference(model, x, warmup=2)
    for _ in range(3):
        g(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30 if B == 1 else 10
    e0.record()
    for _ in range(n):
        g(x)
    e1.record()
    torch.cuda.synchronize()
    out[f"b{B}_ms_per_img"] = round(e0.elapsed_time(e1) / n / B, 4)
    del g
print(json.dumps(out))
