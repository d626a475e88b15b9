"""Achieved HBM bandwidth of the memory-bound kernels at the BASELINE shapes (B=32, 192x416)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from camradepth_b200 import ops  # noqa: E402

d = torch.device("cuda:0")
BF = torch.bfloat16


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def report(name, ms, nbytes):
    print(f"{name:44s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.1f} GB/s  ({nbytes/1e6:.0f} MB)")


SHAPES = [(32, 192 * 416, 128, 320), (32, 192 * 416, 96, 320), (32, 96 * 208, 128, 320), (32, 48 * 104, 512, 512),
          (32, 24 * 52, 1024, 1024), (32, 12 * 26, 640, 640), (32, 48 * 104, 64, 64)]
for (B, N, C, ld) in SHAPES:
    x = torch.randn(B, N, C, device=d).to(BF)
    big = torch.randn(B, N, ld, device=d).to(BF)
    dy = big[..., :C]
    out = torch.empty(B, N, C, dtype=BF, device=d)
    ab = torch.randn(B, C, 2, device=d)
    coef = torch.randn(B, C, 3, device=d)
    pq = torch.zeros(B, C, 2, device=d)
    sums = torch.zeros(B, C, 2, device=d)
    post = torch.ones(B, C, device=d)
    e = B * N * C * 2
    tag = f"[B{B} N{N} C{C} ld{ld}]"
    report("copy (torch) " + tag, timeit(lambda: out.copy_(x)), 2 * e)
    report("chan_stats " + tag, timeit(lambda: ops.chan_stats(x, sums)), e)
    report("affine_act gelu " + tag, timeit(lambda: ops.affine_act(x, big[..., :C], ab, post, ops.ACT_GELU)), 2 * e)
    report("gnact_bwd_reduce gelu " + tag, timeit(lambda: ops.gnact_bwd_reduce(dy, x, ab, post, None, ops.ACT_GELU, pq)), 2 * e)
    report("gnact_bwd_reduce gelu + dz " + tag, timeit(lambda: ops.gnact_bwd_reduce(dy, x, ab, post, None, ops.ACT_GELU, pq, dy)), 3 * e)
    report("gnact_bwd_apply gelu " + tag, timeit(lambda: ops.gnact_bwd_apply(dy, x, ab, post, None, ops.ACT_GELU, coef, out, False)), 3 * e)
    report("gnact_bwd_apply none " + tag, timeit(lambda: ops.gnact_bwd_apply(dy, x, ab, None, None, ops.ACT_NONE, coef, out, False)), 3 * e)
    if C >= 512:
        H = {48 * 104: 48, 24 * 52: 24, 12 * 26: 12}[N]
        W = N // H
        x4, o4 = x.view(B, H, W, C), out.view(B, H, W, C)
        w = torch.randn(C, 9, device=d)
        bias = torch.randn(C, device=d)
        dw, db = torch.zeros(C, 9, device=d), torch.zeros(C, device=d)
        report("dwconv_fwd " + tag, timeit(lambda: ops.dwconv_fwd(x4, ab, w, bias, o4)), 2 * e)
        report("dwconv_bwd (fused) " + tag, timeit(lambda: ops.dwconv_bwd(x4, x4, ab, w, o4, dw, db)), 3 * e)
    del x, big, dy, out

# bicubic x2 into / out of a channel slice of the concat buffer (decoder stage 4: 96x208 -> 192x416, 136 channels)
Bq, Hq, Wq, Cq, ldq = 32, 96, 208, 136, 320
src = torch.randn(Bq, Hq, Wq, Cq, device=d).to(BF)
cat = torch.randn(Bq, 2 * Hq, 2 * Wq, ldq, device=d).to(BF)
dsrc = torch.empty_like(src)
eb = Bq * Hq * Wq * Cq * 2
report("bicubic2x_fwd [B32 96x208 C136 -> ld320]", timeit(lambda: ops.bicubic2x_fwd(src, cat[..., :Cq])), 5 * eb)
report("bicubic2x_bwd [B32 192x416 C136 ld320 ->]", timeit(lambda: ops.bicubic2x_bwd(cat[..., :Cq], dsrc, False)), 5 * eb)
