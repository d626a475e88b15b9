"""Runs the dominant kernels in isolation at the BASELINE shapes so `ncu --set full` can capture them:
the depth_upsample[4] layer-2 convolution (fwd / dgrad / wgrad, B=32 at 192x416) and one stage-2 Mix-FFN
depthwise conv + GroupNorm backward."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from camradepth_b200 import ops  # noqa: E402

d = torch.device("cuda:0")
BF = torch.bfloat16
B, H, W, Cin, Cout = 32, 192, 416, 296, 128
x = (torch.randn(B, H, W, Cin, device=d) * 0.5).to(BF)
w = (torch.randn(Cout, 9 * Cin, device=d) / math.sqrt(9 * Cin)).to(BF)
wd = (torch.randn(Cin, 9 * Cout, device=d) / math.sqrt(9 * Cout)).to(BF)
y = torch.empty(B, H, W, Cout, dtype=BF, device=d)
dx = torch.empty(B, H, W, Cin, dtype=BF, device=d)
dw = torch.zeros(Cout, 9 * Cin, device=d)
for _ in range(int(os.environ.get("CONV_ITERS", "3"))):
    ops.conv_fwd(ops.make_desc(x, y, Cin, Cout, 3, 3, 1, 1), x, w, None, y, use_tc=True)
    ops.conv_fwd(ops.make_desc(y, dx, Cout, Cin, 3, 3, 1, 1, transposed=1), y, wd, None, dx, use_tc=True)
    ops.conv_wgrad(ops.make_desc(x, y, Cin, Cout, 3, 3, 1, 1), x, y, dw, use_tc=True)
# stage-2 Mix-FFN pieces
Bh, Hh, Wh, C = 32, 24, 52, 1024
h = torch.randn(Bh, Hh, Wh, C, device=d).to(BF)
g = torch.randn(Bh, Hh, Wh, C, device=d).to(BF)
o = torch.empty_like(h)
ab = torch.randn(Bh, C, 2, device=d)
wdw = torch.randn(C, 9, device=d)
bias = torch.randn(C, device=d)
coef = torch.randn(Bh, C, 3, device=d)
pq = torch.zeros(Bh, C, 2, device=d)
dwg, dbg = torch.zeros(C, 9, device=d), torch.zeros(C, device=d)
for _ in range(3):
    ops.dwconv_fwd(h, ab, wdw, bias, o)
    ops.dwconv_bwd_input(g, wdw, o)
    ops.dwconv_bwd_weight(g, h, ab, dwg, dbg)
    ops.gnact_bwd_reduce(g, h, ab, None, None, ops.ACT_GELU, pq)
    ops.gnact_bwd_apply(g, h, ab, None, None, ops.ACT_GELU, coef, o, False)
    ops.affine_act(h, o, ab, None, ops.ACT_GELU)
torch.cuda.synchronize()
print("done")
