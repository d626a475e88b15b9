"""Runs the encoder's 1x1 contractions in isolation at the BASELINE shapes (B=32, 192x416) so `ncu --set full`
can capture them: Mix-FFN fc1 / fc2 and the q projection of every stage, forward / data gradient / weight
gradient (simplified_attention.py:34-43, 90-109).  Prints CUDA-event times per launch as well (not under ncu)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from camradepth_b200 import ops  # noqa: E402

d = torch.device("cuda:0")
BF = torch.bfloat16
B = int(os.environ.get("ENC_B", "32"))
ITERS = int(os.environ.get("ENC_ITERS", "1"))
STAGES = [(48, 104, 64, 512), (24, 52, 128, 1024), (12, 26, 160, 640), (6, 13, 256, 1024)]   # H, W, C, rC


def timed(fn, n=20):
    """us per launch; the launches are replayed from a CUDA graph so the host (ctypes + descriptor setup, ~10 us per
    call) does not bound the small shapes."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def gemm_set(tag, H, W, K, N, gn):
    x = torch.randn(B, H, W, K, device=d).to(BF)
    w = (torch.randn(N, K, device=d) / math.sqrt(K)).to(BF)
    wt = w.t().contiguous()
    bias = torch.randn(N, device=d)
    y = torch.empty(B, H, W, N, dtype=BF, device=d)
    dx = torch.empty(B, H, W, K, dtype=BF, device=d)
    dw = torch.zeros(N, K, device=d)
    db = torch.zeros(N, device=d)
    sums = torch.zeros(B, N, 2, device=d)
    fwd = lambda: ops.conv_fwd(ops.make_desc(x, y, K, N, 1, 1, 1, 0), x, w, bias, y, use_tc=True, gn_sums=sums if gn else None)
    dgr = lambda: ops.conv_fwd(ops.make_desc(y, dx, N, K, 1, 1, 1, 0, transposed=1), y, wt, None, dx, use_tc=True)
    wgr = lambda: ops.conv_wgrad(ops.make_desc(x, y, K, N, 1, 1, 1, 0), x, y, dw, use_tc=True, db=db)
    for _ in range(ITERS):
        fwd(); dgr(); wgr()
    if os.environ.get("ENC_TIME", "0") == "1":
        P = B * H * W
        byt_f = (P * K + P * N + N * K) * 2.0
        for nm, fn, byt in (("fwd", fwd, byt_f), ("dgrad", dgr, byt_f), ("wgrad", wgr, (P * K + P * N) * 2.0 + N * K * 4)):
            us = timed(fn)
            print(f"{tag:14s} {nm:6s} M={P:6d} K={K:4d} N={N:4d}: {us:7.1f} us  {2.0 * P * K * N / us / 1e6:7.1f} TF/s  "
                  f"{byt / us / 1e3:6.0f} GB/s algorithmic")


SEL = [int(t) for t in os.environ.get("ENC_STAGES", "1,2,3,4").split(",")]
for s, (H, W, C, rC) in enumerate(STAGES):
    if s + 1 not in SEL:
        continue
    gemm_set(f"s{s + 1}.fc1", H, W, C, rC, True)
    gemm_set(f"s{s + 1}.fc2", H, W, rC, C, False)
    gemm_set(f"s{s + 1}.q", H, W, C, C, False)
torch.cuda.synchronize()
print("done")
