"""tcgen05/TMA implicit-GEMM kernels (GPU): parity against F.conv2d on bf16-rounded operands and against
the CUDA-core path, over the decoder / encoder shapes incl. ragged channel counts, sliced buffers,
partial spatial tiles, accumulate and the sigmoid / fp32-output epilogues."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.test_gpu_ops import dev, rel, nhwc, nchw, rnd, r8

pytestmark = pytest.mark.gpu
BF = torch.bfloat16

TC_CASES = [
    # B, Cin, H, W, Cout, k
    (2, 136, 16, 32, 96, 3),
    (1, 232, 24, 48, 64, 3),
    (2, 296, 8, 16, 128, 3),
    (2, 129, 12, 20, 96, 3),      # Cin padded 129 -> 136, partial spatial tiles
    (1, 416, 6, 13, 96, 3),       # tiny image: one partial tile
    (2, 128, 10, 14, 21, 3),      # ragged N (21 -> UMMA N 32)
    (2, 128, 16, 16, 32, 3),
    (3, 160, 6, 13, 160, 1),      # 1x1, N = 128 + 32
    (2, 64, 48, 104, 512, 1),     # fc1-like
    (2, 1024, 6, 13, 256, 1),     # fc2-like, long K
    (1, 256, 3, 5, 256, 1),
    # H % 16 == 0: the halo-reuse kernel (two accumulators per CTA, 18-row activation boxes)
    (2, 136, 32, 48, 96, 3),
    (1, 232, 16, 32, 64, 3),      # dgrad output 232 -> two N tiles of 128
    (2, 128, 16, 24, 21, 3),      # ragged N, partial tile in W
    (1, 129, 32, 16, 128, 3),     # dgrad output 136 -> one N tile of 144 (accumulators 256 TMEM columns apart)
    (1, 296, 16, 16, 128, 3),     # dgrad output 296 -> two N tiles of 160
    (3, 64, 48, 104, 32, 3),
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_fwd_and_dgrad(case):
    from camradepth_b200 import ops
    B, Cin, H, W, Cout, k = case
    pad = k // 2
    torch.manual_seed(0)
    d = dev()
    x = torch.randn(B, Cin, H, W, device=d)
    w = torch.randn(Cout, Cin, k, k, device=d) / math.sqrt(Cin * k * k)
    bias = torch.randn(Cout, device=d)
    xr, wr = rnd(x, BF).requires_grad_(True), rnd(w, BF).requires_grad_(True)
    yr = F.conv2d(xr, wr, bias, padding=pad)
    cin_p, cout_p = r8(Cin), r8(Cout)
    xb = nhwc(x, BF, cin_p + 8)[..., :cin_p]
    wp = torch.zeros(Cout, k * k * cin_p, dtype=BF, device=d)
    ops.weight_pack(w, wp, None, Cout, Cin, k * k, cin_p, cout_p, 0)
    ybuf = torch.zeros(B, H, W, cout_p + 16, dtype=BF, device=d)
    y = ybuf[..., 8:8 + cout_p]
    desc = ops.make_desc(xb, y, cin_p, Cout, k, k, 1, pad)
    ops.conv_fwd(desc, xb, wp, bias, y, use_tc=True)
    torch.cuda.synchronize()
    assert rel(nchw(y, Cout), yr) < 5e-3, rel(nchw(y, Cout), yr)
    assert float(ybuf[..., :8].abs().max()) == 0 and float(ybuf[..., 8 + cout_p:].abs().max()) == 0
    # GroupNorm statistics from the accumulator read-out: per-(sample, channel) sum and sum of squares of the fp32
    # results (bias included), accumulated on top of the buffer's contents; 1x1 tiles may straddle samples
    sums = torch.full((B, Cout, 2), 0.25, dtype=torch.float32, device=d)
    y_gn = torch.zeros(B, H, W, cout_p, dtype=BF, device=d)
    ops.conv_fwd(ops.make_desc(xb, y_gn, cin_p, Cout, k, k, 1, pad), xb, wp, bias, y_gn, use_tc=True, gn_sums=sums)
    torch.cuda.synchronize()
    assert torch.equal(y_gn[..., :Cout], y[..., :Cout])
    yd = yr.detach().double()
    assert rel(sums[..., 0] - 0.25, yd.sum((2, 3))) < 2e-4, "sum"
    assert rel(sums[..., 1] - 0.25, (yd * yd).sum((2, 3))) < 1e-4, "sum of squares"
    # same thing on the CUDA-core path: the two must agree to accumulation-order noise
    y2 = torch.zeros(B, H, W, cout_p, dtype=BF, device=d)
    desc2 = ops.make_desc(xb, y2, cin_p, Cout, k, k, 1, pad)
    ops.conv_fwd(desc2, xb, wp, bias, y2)
    assert rel(y[..., :Cout], y2[..., :Cout]) < 4e-3
    # fp32 output + sigmoid epilogue
    y3 = torch.zeros(B, H, W, cout_p, dtype=torch.float32, device=d)
    desc3 = ops.make_desc(xb, y3, cin_p, Cout, k, k, 1, pad, act=ops.ACT_SIGMOID)
    ops.conv_fwd(desc3, xb, wp, bias, y3, use_tc=True)
    assert rel(nchw(y3, Cout), torch.sigmoid(yr)) < 1e-4
    # dgrad = transposed mode, accumulated on top of ones
    gy = torch.randn_like(yr)
    gx_ref, gw_ref = torch.autograd.grad(yr, (xr, wr), rnd(gy, BF))
    dy = nhwc(gy, BF, cout_p)
    wd = torch.zeros(cin_p, k * k * cout_p, dtype=BF, device=d)
    ops.weight_pack(w, wd, None, Cout, Cin, k * k, cin_p, cout_p, 1)
    dx = torch.ones(B, H, W, cin_p, dtype=BF, device=d)
    desc4 = ops.make_desc(dy, dx, cout_p, cin_p, k, k, 1, pad, transposed=1, accumulate=1)
    ops.conv_fwd(desc4, dy, wd, None, dx, use_tc=True)
    assert rel(nchw(dx, Cin) - 1.0, gx_ref) < 2e-2
    # wgrad (both operands MN-major, split-K with fp32 atomics); dy read from a channel slice
    dyb = torch.zeros(B, H, W, cout_p + 8, dtype=BF, device=d)
    dyb[..., 8:] = dy
    dys = dyb[..., 8:]
    dwp = torch.zeros(Cout, k * k * cin_p, dtype=torch.float32, device=d)
    ops.conv_wgrad(ops.make_desc(xb, dys, cin_p, Cout, k, k, 1, pad), xb, dys, dwp, use_tc=True)
    gw = torch.empty_like(w)
    ops.weight_unpack_grad(dwp, gw, None, Cout, Cin, k * k, cin_p, False)
    torch.cuda.synchronize()
    assert rel(gw, gw_ref) < 1e-3, rel(gw, gw_ref)
    if cin_p > Cin:      # padded input channels must receive exactly zero
        pad_cols = dwp.view(Cout, k * k, cin_p)[:, :, Cin:]
        assert float(pad_cols.abs().max()) == 0
    if k == 1:
        # 1x1: weight and bias gradients from one pass over dy (an extra MMA against a tile of ones), accumulating
        dwp2 = torch.zeros_like(dwp)
        db = torch.full((Cout,), 0.5, dtype=torch.float32, device=d)
        ops.conv_wgrad(ops.make_desc(xb, dys, cin_p, Cout, k, k, 1, pad), xb, dys, dwp2, use_tc=True, db=db)
        assert rel(dwp2, dwp) < 1e-5
        assert rel(db - 0.5, rnd(gy, BF).sum((0, 2, 3))) < 1e-4


def test_conv_tc_large_matches_simt():
    """BASELINE-sized layer: depth_upsample[4] layer 2 (Cin 296 -> 128 at 192x416), B=2."""
    from camradepth_b200 import ops
    d = dev()
    torch.manual_seed(1)
    B, Cin, H, W, Cout = 2, 296, 192, 416, 128
    xb = (torch.randn(B, H, W, Cin, device=d) * 0.5).to(BF)
    wp = (torch.randn(Cout, 9 * Cin, device=d) / math.sqrt(9 * Cin)).to(BF)
    y1 = torch.empty(B, H, W, Cout, dtype=BF, device=d)
    y2 = torch.empty(B, H, W, Cout, dtype=BF, device=d)
    ops.conv_fwd(ops.make_desc(xb, y1, Cin, Cout, 3, 3, 1, 1), xb, wp, None, y1, use_tc=True)
    ops.conv_fwd(ops.make_desc(xb, y2, Cin, Cout, 3, 3, 1, 1), xb, wp, None, y2)
    torch.cuda.synchronize()
    assert rel(y1, y2) < 4e-3
    # linearity (size-independent property): conv(2x) == 2 conv(x) exactly in bf16 (power-of-two scale)
    y3 = torch.empty_like(y1)
    x2 = (xb.float() * 2).to(BF)
    ops.conv_fwd(ops.make_desc(x2, y3, Cin, Cout, 3, 3, 1, 1), x2, wp, None, y3, use_tc=True)
    assert torch.equal(y3.float(), y1.float() * 2)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        ops.conv_fwd(ops.make_desc(xb, y1, Cin, Cout, 3, 3, 1, 1), xb, wp, None, y1, use_tc=True)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print(f"tc conv 296->128 @192x416 B=2: {ms:.3f} ms, {2*B*H*W*Cout*9*Cin/ms/1e9:.1f} TFLOP/s")
    dy = (torch.randn(B, H, W, Cout, device=d) * 0.5).to(BF)
    dw1 = torch.zeros(Cout, 9 * Cin, device=d)
    dw2 = torch.zeros(Cout, 9 * Cin, device=d)
    ops.conv_wgrad(ops.make_desc(xb, dy, Cin, Cout, 3, 3, 1, 1), xb, dy, dw1, use_tc=True)
    ops.conv_wgrad(ops.make_desc(xb, dy, Cin, Cout, 3, 3, 1, 1), xb, dy, dw2)
    torch.cuda.synchronize()
    assert rel(dw1, dw2) < 1e-4, rel(dw1, dw2)
    t0.record()
    for _ in range(5):
        ops.conv_wgrad(ops.make_desc(xb, dy, Cin, Cout, 3, 3, 1, 1), xb, dy, dw1, use_tc=True)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print(f"tc wgrad 296->128 @192x416 B=2: {ms:.3f} ms, {2*B*H*W*Cout*9*Cin/ms/1e9:.1f} TFLOP/s")


@pytest.mark.parametrize("case", [(2, 32, 48, 21), (1, 16, 32, 19), (2, 12, 20, 21)])   # halo kernel x2, plain kernel
def test_conv_argmax_fused(case):
    """Seg_Block head fused into the conv read-out: argmax over the fp32 accumulators, logits never stored."""
    from camradepth_b200 import ops
    B, H, W, ncls = case
    Cin = 128
    d = dev()
    torch.manual_seed(3)
    x = torch.randn(B, Cin, H, W, device=d)
    w = torch.randn(ncls, Cin, 3, 3, device=d) / math.sqrt(9 * Cin)
    bias = torch.randn(ncls, device=d) * 0.1
    lg = F.conv2d(rnd(x, BF), rnd(w, BF), bias, padding=1)
    xb = nhwc(x, BF, Cin + 8)[..., :Cin]
    wp = torch.zeros(ncls, 9 * Cin, dtype=BF, device=d)
    ops.weight_pack(w, wp, None, ncls, Cin, 9, Cin, r8(ncls), 0)
    buf0 = torch.full((B, H, W, 136), 7.0, dtype=BF, device=d)
    buf1 = torch.full((B, H, W, 8), 7.0, dtype=BF, device=d)
    mf = torch.full((B, 1, H, W), 7.0, device=d)
    desc = ops.make_desc(xb, xb, Cin, ncls, 3, 3, 1, 1, out_dtype=ops.BF16)
    ops.conv_argmax(desc, xb, wp, bias, ncls, buf0[..., 128:129], buf1[..., 3:4], mf)
    torch.cuda.synchronize()
    idx = torch.round(mf * ncls).long().squeeze(1)
    ref = lg.argmax(1)
    top2 = lg.topk(2, dim=1).values
    gap = top2[:, 0] - top2[:, 1]
    bad = idx != ref
    assert float(bad.float().mean()) < 2e-3
    assert not bool((bad & (gap > 1e-3)).any())          # disagreements only at near-ties (accumulation order)
    want = (idx.float() / ncls).to(BF)
    assert torch.equal(buf0[..., 128], want) and torch.equal(buf1[..., 3], want)
    assert float((buf0[..., :128] - 7).abs().max()) == 0 and float((buf0[..., 129:] - 7).abs().max()) == 0
    assert float((buf1[..., :3] - 7).abs().max()) == 0 and float((buf1[..., 4:] - 7).abs().max()) == 0


STRIDED_CASES = [
    # B, Cin, H, W, Cout, k, stride, pad
    (2, 64, 48, 104, 64, 8, 8, 0),       # stage-1 spatial-reduction conv (simplified_attention.py:68)
    (2, 128, 24, 52, 128, 4, 4, 0),      # stage 2
    (3, 160, 12, 26, 160, 2, 2, 0),      # stage 3: ragged last 64-channel chunk, tiles cross sample boundaries
    (2, 64, 48, 104, 128, 3, 2, 1),      # patch_embed2 (k3 s2 p1, :158-160)
    (1, 128, 24, 52, 160, 3, 2, 1),      # patch_embed3
    (2, 64, 16, 16, 96, 2, 2, 0),
]


@pytest.mark.parametrize("case", STRIDED_CASES)
def test_conv_tc_strided_implicit_gemm(case):
    """Strided convolutions straight from the NHWC tensor through 5-D tensor maps (no im2col buffer): forward with
    bias and GroupNorm sums, and the weight gradient, against F.conv2d / autograd on the bf16-rounded operands."""
    from camradepth_b200 import ops
    B, Cin, H, W, Cout, k, st, pad = case
    torch.manual_seed(0)
    d = dev()
    x = torch.randn(B, Cin, H, W, device=d)
    w = torch.randn(Cout, Cin, k, k, device=d) / math.sqrt(Cin * k * k)
    bias = torch.randn(Cout, device=d)
    xr, wr = rnd(x, BF).requires_grad_(True), rnd(w, BF).requires_grad_(True)
    yr = F.conv2d(xr, wr, bias, stride=st, padding=pad)
    Ho, Wo = yr.shape[2], yr.shape[3]
    xb = nhwc(x, BF)
    wp = torch.zeros(Cout, k * k * Cin, dtype=BF, device=d)
    ops.weight_pack(w, wp, None, Cout, Cin, k * k, Cin, r8(Cout), 0)
    ybuf = torch.zeros(B, Ho, Wo, Cout + 16, dtype=BF, device=d)
    y = ybuf[..., 8:8 + Cout]
    sums = torch.zeros(B, Cout, 2, device=d)
    ops.conv_fwd(ops.make_desc(xb, y, Cin, Cout, k, k, st, pad), xb, wp, bias, y, use_tc=True, gn_sums=sums)
    torch.cuda.synchronize()
    assert rel(nchw(y, Cout), yr) < 5e-3, rel(nchw(y, Cout), yr)
    assert float(ybuf[..., :8].abs().max()) == 0 and float(ybuf[..., 8 + Cout:].abs().max()) == 0
    yd = yr.detach().double()
    assert rel(sums[..., 0], yd.sum((2, 3))) < 2e-4 and rel(sums[..., 1], (yd * yd).sum((2, 3))) < 1e-4
    # weight gradient
    gy = torch.randn_like(yr)
    (gw_ref,) = torch.autograd.grad(yr, (wr,), rnd(gy, BF))
    dy = nhwc(gy, BF)
    dwp = torch.zeros(Cout, k * k * Cin, dtype=torch.float32, device=d)
    ops.conv_wgrad(ops.make_desc(xb, dy, Cin, Cout, k, k, st, pad), xb, dy, dwp, use_tc=True)
    gw = torch.empty_like(w)
    ops.weight_unpack_grad(dwp, gw, None, Cout, Cin, k * k, Cin, False)
    torch.cuda.synchronize()
    assert rel(gw, gw_ref) < 1e-3, rel(gw, gw_ref)


@pytest.mark.parametrize("case", [(2, 64, 48, 104, 64, 8), (2, 128, 24, 52, 128, 4), (3, 160, 12, 26, 160, 2), (1, 64, 16, 16, 96, 2)])
def test_conv_tc_strided_dgrad_depth_to_space(case):
    """Data gradient of a k == stride conv as a 1x1 GEMM with a depth-to-space read-out (with and without
    accumulation into an existing gradient), against autograd on the bf16-rounded operands."""
    from camradepth_b200 import ops
    B, Cin, H, W, Cout, st = case
    torch.manual_seed(0)
    d = dev()
    x = torch.randn(B, Cin, H, W, device=d)
    w = torch.randn(Cout, Cin, st, st, device=d) / math.sqrt(Cin * st * st)
    xr, wr = rnd(x, BF).requires_grad_(True), rnd(w, BF)
    yr = F.conv2d(xr, wr, None, stride=st)
    gy = torch.randn_like(yr)
    (gx_ref,) = torch.autograd.grad(yr, (xr,), rnd(gy, BF))
    dy = nhwc(gy, BF)
    w2 = torch.zeros(st * st * Cin, r8(Cout), dtype=BF, device=d)
    ops.weight_pack(w, w2, None, Cout, Cin, st * st, Cin, r8(Cout), 2)
    for acc in (0, 1):
        dx = torch.full((B, H, W, Cin), 1.0 if acc else 7.0, dtype=BF, device=d)
        desc = ops.make_desc(dy, dx, r8(Cout), Cin, st, st, st, 0, transposed=1, accumulate=acc)
        ops.conv_fwd(desc, dy, w2, None, dx, use_tc=True)
        torch.cuda.synchronize()
        assert rel(nchw(dx, Cin) - (1.0 if acc else 0.0), gx_ref) < (2e-2 if acc else 5e-3)
