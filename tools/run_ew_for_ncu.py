import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from camradepth_b200 import ops
d = torch.device("cuda:0"); BF = torch.bfloat16
B, N, C = 32, 192 * 416, 128
x = torch.randn(B, N, C, device=d).to(BF)
out = torch.empty_like(x)
sums = torch.zeros(B, C, 2, device=d); ab = torch.randn(B, C, 2, device=d)
for _ in range(2):
    ops.chan_stats(x, sums)
    ops.affine_act(x, out, ab, None, ops.ACT_NONE)
    out.copy_(x)
torch.cuda.synchronize()
