"""Per-kernel parity (GPU): every C-ABI entry point against the ATen op it replaces, fp32 and bf16.

Tolerances: fp32 kernels 2e-5 relative L2 (summation order only); bf16 kernels are compared with the
same fp32 reference evaluated on bf16-rounded operands, 1e-2 relative L2 (output rounding)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DT = [torch.float32, torch.bfloat16]
TOL = {torch.float32: 2e-5, torch.bfloat16: 1e-2}


def dev():
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(t, dtype, cpad=None):
    """NCHW fp32 -> NHWC `dtype`, channels zero-padded to cpad."""
    B, C, H, W = t.shape
    cp = C if cpad is None else cpad
    out = torch.zeros(B, H, W, cp, dtype=dtype, device=t.device)
    out[..., :C] = t.permute(0, 2, 3, 1).to(dtype)
    return out


def nchw(t, C=None):
    t = t.float().permute(0, 3, 1, 2)
    return t if C is None else t[:, :C]


def rnd(t, dtype):
    return t.to(dtype).float()


def r8(c):
    return (c + 7) // 8 * 8


CONV_CASES = [
    # B, Cin, H, W, Cout, k, stride, pad
    (2, 7, 32, 64, 64, 7, 4, 3),
    (2, 64, 16, 24, 128, 3, 2, 1),
    (1, 136, 24, 40, 96, 3, 1, 1),
    (2, 129, 12, 20, 64, 3, 1, 1),
    (3, 160, 6, 13, 160, 1, 1, 0),
    (2, 64, 16, 32, 64, 8, 8, 0),
    (2, 128, 10, 14, 21, 3, 1, 1),
    (1, 128, 9, 11, 32, 3, 1, 1),
]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(case, dtype):
    from camradepth_b200 import ops
    B, Cin, H, W, Cout, k, stride, pad = case
    torch.manual_seed(0)
    d = dev()
    x = torch.randn(B, Cin, H, W, device=d)
    w = torch.randn(Cout, Cin, k, k, device=d) / math.sqrt(Cin * k * k)
    bias = torch.randn(Cout, device=d)
    xr, wr = rnd(x, dtype).requires_grad_(True), rnd(w, dtype).requires_grad_(True)
    yr = F.conv2d(xr, wr, bias, stride=stride, padding=pad)
    Ho, Wo = yr.shape[2], yr.shape[3]
    cin_p, cout_p = r8(Cin), r8(Cout)
    # forward, written into a channel slice of a wider buffer
    xb = nhwc(x, dtype, cin_p + 8)[..., :cin_p]
    wp = torch.zeros(Cout, k * k * cin_p, dtype=dtype, device=d)
    ops.weight_pack(w, wp, None, Cout, Cin, k * k, cin_p, cout_p, 0)
    ybuf = torch.zeros(B, Ho, Wo, cout_p + 16, dtype=dtype, device=d)
    y = ybuf[..., 8:8 + cout_p]
    desc = ops.make_desc(xb, y, cin_p, Cout, k, k, stride, pad)
    ops.conv_fwd(desc, xb, wp, bias, y)
    torch.cuda.synchronize()
    assert rel(nchw(y, Cout), yr) < TOL[dtype]
    assert float(ybuf[..., :8].abs().max()) == 0 and float(ybuf[..., 8 + cout_p:].abs().max()) == 0
    # NCHW fp32 output + sigmoid epilogue
    if dtype == torch.float32 or Cout == 21:
        yn = torch.empty(B, Cout, Ho, Wo, dtype=torch.float32, device=d)
        desc = ops.make_desc(xb, yn, cin_p, Cout, k, k, stride, pad, act=ops.ACT_SIGMOID, out_nchw=1)
        ops.conv_fwd(desc, xb, wp, bias, yn)
        assert rel(yn, torch.sigmoid(yr)) < TOL[dtype]
    # backward reference
    gy = torch.randn_like(yr)
    gx_ref, gw_ref = torch.autograd.grad(yr, (xr, wr), rnd(gy, dtype))
    dy = nhwc(gy, dtype, cout_p)
    # dgrad (transposed mode), accumulate on top of ones
    wd = torch.zeros(cin_p, k * k * cout_p, dtype=dtype, device=d)
    ops.weight_pack(w, wd, None, Cout, Cin, k * k, cin_p, cout_p, 1)
    dx = torch.ones(B, H, W, cin_p, dtype=dtype, device=d)
    desc = ops.make_desc(dy, dx, cout_p, cin_p, k, k, stride, pad, transposed=1, accumulate=1)
    ops.conv_fwd(desc, dy, wd, None, dx)
    assert rel(nchw(dx, Cin) - 1.0, gx_ref) < 2 * TOL[dtype]
    # wgrad
    dwp = torch.zeros(Cout, k * k * cin_p, dtype=torch.float32, device=d)
    desc = ops.make_desc(xb, dy, cin_p, Cout, k, k, stride, pad)
    ops.conv_wgrad(desc, xb, dy, dwp)
    gw = torch.empty_like(w)
    ops.weight_unpack_grad(dwp, gw, None, Cout, Cin, k * k, cin_p, False)
    assert rel(gw, gw_ref) < 5e-5 if dtype == torch.float32 else rel(gw, gw_ref) < 1e-3
    db = torch.zeros(Cout, device=d)
    ops.col_sum(dy, db, Cout)
    assert rel(db, rnd(gy, dtype).sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("dtype", DT)
def test_weight_pack_with_channel_map(dtype):
    from camradepth_b200 import ops
    d = dev()
    Cout, Cin, taps, cin_p = 5, 11, 9, 24
    w = torch.randn(Cout, Cin, 3, 3, device=d)
    cmap = torch.tensor([c if c < 6 else 8 + c for c in range(Cin)], dtype=torch.int32, device=d)
    wp = torch.zeros(Cout, taps * cin_p, dtype=dtype, device=d)
    ops.weight_pack(w, wp, cmap, Cout, Cin, taps, cin_p, 8, 0)
    ref = torch.zeros(Cout, taps, cin_p, device=d)
    ref[:, :, cmap.long()] = w.reshape(Cout, Cin, taps).permute(0, 2, 1)
    assert torch.equal(wp.float().view(Cout, taps, cin_p), ref.to(dtype).float())
    wd = torch.zeros(cin_p, taps * 8, dtype=dtype, device=d)
    ops.weight_pack(w, wd, cmap, Cout, Cin, taps, cin_p, 8, 1)
    refd = torch.zeros(cin_p, taps, 8, device=d)
    refd[cmap.long(), :, :Cout] = w.reshape(Cout, Cin, taps).permute(1, 2, 0)
    assert torch.equal(wd.float().view(cin_p, taps, 8), refd.to(dtype).float())
    g = torch.ones_like(w)
    ops.weight_unpack_grad(ref.contiguous().view(Cout, -1), g, cmap, Cout, Cin, taps, cin_p, True)
    assert torch.allclose(g, w + 1)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(2, 64, 12, 20, 4), (2, 512, 6, 10, 32), (1, 96, 24, 40, 6), (3, 1024, 3, 5, 16),
                                   # large enough for the TMA-staged bf16 kernels (ragged pixel counts, channel
                                   # tiles of 128 / 96 / 160 / 64 channels, inputs that are slices of wider buffers)
                                   (2, 128, 63, 71, 8), (2, 96, 80, 72, 6), (1, 1024, 33, 35, 16), (4, 160, 41, 43, 10),
                                   (3, 192, 50, 40, 12)])
@pytest.mark.parametrize("act", [0, 1])
def test_groupnorm_protocol(shape, act, dtype):
    from camradepth_b200 import ops
    B, C, H, W, G = shape
    d = dev()
    torch.manual_seed(1)
    x = (torch.randn(B, C, H, W, device=d) * 1.7 + 0.3)
    gamma = torch.randn(C, device=d) * 0.3 + 1
    beta = torch.randn(C, device=d) * 0.2
    post = (torch.rand(B, C, device=d) > 0.2).float() / 0.8
    xr = rnd(x, dtype).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.group_norm(xr, G, gr, br, 1e-5)
    yr = (F.gelu(z) if act else z) * post.view(B, C, 1, 1)
    xb = nhwc(x, dtype)
    N = H * W
    if B * N * C >= (1 << 20):
        wide = torch.zeros(B, H, W, C + 24, dtype=dtype, device=d)
        wide[..., 8:8 + C] = xb
        xb = wide[..., 8:8 + C]
    sums = torch.zeros(B, C, 2, device=d)
    ops.chan_stats(xb, sums)
    ab = torch.empty(B, C, 2, device=d)
    mr = torch.empty(B, G, 2, device=d)
    xbar = torch.empty(B, C, device=d)
    ops.gn_finalize(sums, gamma, beta, ab, mr, xbar, B, C, G, N)
    ybuf = torch.zeros(B, H, W, C + 8, dtype=dtype, device=d)
    y = ybuf[..., 8:]
    ops.affine_act(xb, y, ab, post, act)
    assert rel(nchw(y), yr) < TOL[dtype]
    assert rel(xbar, z.mean((2, 3))) < 1e-4
    # backward, with the broadcast add-in used by the attention token-mean path
    gy = torch.randn_like(yr)
    addbc = torch.randn(B, C, device=d) * 0.1
    gye = rnd(gy, dtype) + addbc.view(B, C, 1, 1)
    gx_ref, gg_ref, gb_ref = torch.autograd.grad(yr, (xr, gr, br), gye)
    dy = nhwc(gy, dtype)
    pq = torch.zeros(B, C, 2, device=d)
    ops.gnact_bwd_reduce(dy, xb, ab, post, addbc, act, pq)
    coef = torch.empty(B, C, 3, device=d)
    dg, db = torch.zeros(C, device=d), torch.zeros(C, device=d)
    ops.gn_bwd_finalize(pq, mr, gamma, coef, dg, db, B, C, G, N)
    dx = torch.ones(B, H, W, C, dtype=torch.float32, device=d)
    ops.gnact_bwd_apply(dy, xb, ab, post, addbc, act, coef, dx, True)
    t = 5e-4 if dtype == torch.float32 else 2e-2
    assert rel(nchw(dx) - 1, gx_ref) < t
    assert rel(dg, gg_ref) < t and rel(db, gb_ref) < t
    dx2 = torch.empty(B, H, W, C, dtype=dtype, device=d)
    ops.gnact_bwd_apply(dy, xb, ab, post, addbc, act, coef, dx2, False)
    assert rel(nchw(dx2), gx_ref) < max(t, TOL[dtype])
    # variant used by the engine: the reduce pass materialises dz in place, the apply pass is a pure stream
    dz = dy.clone()
    pq2 = torch.zeros(B, C, 2, device=d)
    ops.gnact_bwd_reduce(dz, xb, ab, post, addbc, act, pq2, dz)
    assert rel(pq2, pq) < 1e-5
    dx3 = torch.empty(B, H, W, C, dtype=dtype, device=d)
    ops.gnact_bwd_apply(dz, xb, ab, None, None, ops.ACT_NONE, coef, dx3, False)
    assert rel(nchw(dx3), gx_ref) < max(t, 2 * TOL[dtype])
    # ---- the one-launch variants (gn_fused_small.cu): same results from one kernel per direction
    assert ops.gn_fused_supported(B, N, C, G)
    ab_f, mr_f, xbar_f = torch.empty(B, C, 2, device=d), torch.empty(B, G, 2, device=d), torch.empty(B, C, device=d)
    yfb = torch.zeros(B, H, W, C + 8, dtype=dtype, device=d)
    yf = yfb[..., 8:]
    ops.gn_fused_fwd(xb, yf, gamma, beta, G, None, post, act, ab_f, mr_f, xbar_f)
    assert rel(nchw(yf), yr) < TOL[dtype] and float(yfb[..., :8].abs().max()) == 0
    assert rel(ab_f, ab) < 1e-5 and rel(mr_f, mr) < 1e-5 and rel(xbar_f, xbar) < 1e-5
    # statistics handed in (conv read-out) and finalize-only form
    ab_g, mr_g = torch.empty_like(ab), torch.empty_like(mr)
    ops.gn_fused_fwd(xb, None, gamma, beta, G, sums, None, 0, ab_g, mr_g, None)
    assert rel(ab_g, ab) < 1e-6 and rel(mr_g, mr) < 1e-6
    yg = torch.empty(B, H, W, C, dtype=torch.float32, device=d)
    ops.gn_fused_fwd(xb, yg, gamma, beta, G, sums, post, act, None, None, None)
    assert rel(nchw(yg), yr) < (1e-5 if dtype == torch.float32 else TOL[dtype])
    dgf, dbf = torch.zeros(C, device=d), torch.zeros(C, device=d)
    dyf = dy.clone()
    dxf = torch.ones(B, H, W, C, dtype=torch.float32, device=d)
    ops.gn_fused_bwd(dyf, xb, ab, mr, gamma, G, post, addbc, act, dxf, True, dgf, dbf)
    assert rel(nchw(dxf) - 1, gx_ref) < t
    assert rel(dgf, gg_ref) < t and rel(dbf, gb_ref) < t
    if act:
        assert rel(dyf, dz) < 1e-3            # dz left in place of dy, like the reduce pass does (bf16 rounding of dz)
    else:
        assert torch.equal(dyf, dy)
    dyf = dy.clone()
    dxh = torch.empty(B, H, W, C + 16, dtype=dtype, device=d)[..., 8:8 + C]
    ops.gn_fused_bwd(dyf, xb, ab, mr, gamma, G, post, addbc, act, dxh, False, None, None)
    assert rel(nchw(dxh), gx_ref) < max(t, TOL[dtype])


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(2, 160, 9, 14),
                                   # TMA-staged bf16 path: C % 64 == 0 and >= 2^19 elements; ragged last strip
                                   # (W = 100), one-strip maps (W = 13), split row ranges, H not a multiple of 6
                                   (4, 128, 24, 52), (2, 64, 50, 100), (8, 1024, 6, 13), (2, 192, 47, 31)])
def test_dwconv(shape, dtype):
    from camradepth_b200 import ops
    d = dev()
    torch.manual_seed(2)
    B, C, H, W = shape
    x = torch.randn(B, C, H, W, device=d)
    w = torch.randn(C, 1, 3, 3, device=d) * 0.3
    bias = torch.randn(C, device=d)
    a = torch.randn(B, C, device=d) * 0.5 + 1
    sh = torch.randn(B, C, device=d) * 0.2
    ab = torch.stack([a, sh], -1).contiguous()
    xr = rnd(x, dtype).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    xn = xr * a.view(B, C, 1, 1) + sh.view(B, C, 1, 1)
    xn.retain_grad()
    yr = F.conv2d(xn, wr, bias, padding=1, groups=C)
    xb = nhwc(x, dtype)
    y = torch.empty(B, H, W, C, dtype=dtype, device=d)
    ops.dwconv_fwd(xb, ab, w, bias, y)
    assert rel(nchw(y), yr) < TOL[dtype]
    gy = torch.randn_like(yr)
    yr.backward(rnd(gy, dtype))
    dy = nhwc(gy, dtype)
    dxn = torch.empty(B, H, W, C, dtype=dtype, device=d)
    ops.dwconv_bwd_input(dy, w, dxn)
    assert rel(nchw(dxn), xn.grad) < TOL[dtype]
    dw, db = torch.zeros(C, 9, device=d), torch.zeros(C, device=d)
    ops.dwconv_bwd_weight(dy, xb, ab, dw, db)
    assert rel(dw, wr.grad.view(C, 9)) < 1e-3 and rel(db, rnd(gy, dtype).sum((0, 2, 3))) < 1e-4
    # fused backward (one pass over dy)
    dxn2 = torch.empty(B, H, W, C, dtype=dtype, device=d)
    dw2, db2 = torch.zeros(C, 9, device=d), torch.zeros(C, device=d)
    ops.dwconv_bwd(dy, xb, ab, w, dxn2, dw2, db2)
    assert rel(nchw(dxn2), xn.grad) < TOL[dtype]
    assert rel(dw2, wr.grad.view(C, 9)) < 1e-3 and rel(db2, rnd(gy, dtype).sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("cfg", [(2, 200, 78, 64, 1), (2, 130, 70, 160, 4), (1, 78, 78, 256, 8), (2, 64, 4, 128, 2),
                                 # several 128-token tiles; 200 keys (two 256-column accumulators); 300 and 1400
                                 # keys (> 256: chunks of 256 keys with a running max)
                                 (3, 1300, 78, 128, 2), (2, 500, 200, 160, 4), (1, 300, 300, 64, 2), (1, 2000, 1400, 128, 2)])
def test_attention_score(cfg, dtype):
    from camradepth_b200 import ops
    B, N, M, C, heads = cfg
    d = dev()
    torch.manual_seed(3)
    q = torch.randn(B, N, C, device=d)
    k = torch.randn(B, M, C, device=d)
    qr, kr = rnd(q, dtype).requires_grad_(True), rnd(k, dtype).requires_grad_(True)
    hd = C // heads
    scale = hd ** -0.5
    att = torch.einsum("bnhd,bmhd->bhnm", qr.view(B, N, heads, hd), kr.view(B, M, heads, hd)) * scale
    mx, am = att.max(-1)
    sr = mx.sum(1)
    s = torch.empty(B, N, device=d)
    idx = torch.empty(B, heads, N, dtype=torch.int16, device=d)
    ops.attn_qkmax_fwd(q.to(dtype), k.to(dtype), s, idx, heads, scale)
    assert rel(s, sr) < 1e-4
    assert float((idx.long() != am).float().mean()) < 2e-3
    ds = torch.randn(B, N, device=d)
    gq, gk = torch.autograd.grad(sr, (qr, kr), ds)
    dq = torch.empty(B, N, C, dtype=dtype, device=d)
    dk = torch.zeros(B, M, C, device=d)
    ops.attn_qkmax_bwd(ds, q.to(dtype), k.to(dtype), idx, dq, dk, heads, scale)
    assert rel(dq, gq) < 3 * TOL[dtype] and rel(dk, gk) < 3 * TOL[dtype]


def test_attention_out_and_pv():
    from camradepth_b200 import ops
    d = dev()
    torch.manual_seed(4)
    B, N, C = 3, 130, 160
    x = torch.randn(B, N, C, device=d)
    xbar = torch.randn(B, C, device=d, requires_grad=True)
    Wp = (torch.randn(C, C, device=d) * 0.1).requires_grad_(True)
    bp = torch.randn(C, device=d, requires_grad=True)
    s = torch.randn(B, N, device=d, requires_grad=True)
    dp = torch.tensor([1.0, 0.0, 1.25], device=d)
    pvr = xbar @ Wp.t()
    outr = x + dp.view(B, 1, 1) * (pvr.unsqueeze(1) * s.unsqueeze(-1) + bp)
    pv = torch.empty(B, C, device=d)
    ops.attn_pv_fwd(xbar.detach(), Wp.detach(), pv)
    assert rel(pv, pvr) < 1e-5
    out = torch.empty_like(x)
    ops.attn_out_residual(x, pv, s.detach(), bp.detach(), dp, out)
    assert rel(out, outr) < 1e-6
    g = torch.randn_like(outr)
    gs, gxbar, gW, gb = torch.autograd.grad(outr, (s, xbar, Wp, bp), g)
    ds = torch.empty(B, N, device=d)
    dpv = torch.empty(B, C, device=d)
    dbp = torch.zeros(C, device=d)
    tmp = torch.empty(B, C, 2, device=d)
    ops.attn_out_bwd(g, pv, s.detach(), dp, ds, dpv, dbp, tmp)
    assert rel(ds, gs) < 1e-4 and rel(dbp, gb) < 1e-4
    dW = torch.zeros(C, C, device=d)
    dxbar = torch.empty(B, C, device=d)
    ops.attn_pv_bwd(dpv, xbar.detach(), Wp.detach(), dW, dxbar, 0.5)
    assert rel(dW, gW) < 1e-4 and rel(dxbar, gxbar * 0.5) < 1e-4
    # residual add / scale-cast / add
    y = torch.randn(B, N, C, device=d)
    o = torch.empty_like(x)
    ops.residual_add(x, y.bfloat16(), dp, o)
    assert rel(o, x + dp.view(B, 1, 1) * y.bfloat16().float()) < 1e-6
    c = torch.empty(B, N, C, dtype=torch.bfloat16, device=d)
    ops.scale_cast(x, dp, c)
    assert rel(c, (x * dp.view(B, 1, 1)).bfloat16()) < 1e-6
    acc = x.clone()
    ops.add_f32(acc, y.bfloat16())
    assert rel(acc, x + y.bfloat16().float()) < 1e-6


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(2, 136, 6, 13), (1, 256, 3, 4), (2, 8, 12, 20),
                                   # TMA-staged bf16 path (>= 2^18 elements): a short third channel tile (136),
                                   # ragged strips (W = 37), split row ranges, H not a multiple of the box height
                                   (2, 136, 24, 52), (4, 128, 48, 104), (4, 256, 12, 26), (3, 64, 50, 37),
                                   (16, 72, 7, 33)])
def test_bicubic(shape, dtype):
    from camradepth_b200 import ops
    B, C, H, W = shape
    d = dev()
    torch.manual_seed(5)
    x = torch.randn(B, C, H, W, device=d)
    xr = rnd(x, dtype).requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=2, mode="bicubic")
    xb = nhwc(x, dtype)
    ybuf = torch.zeros(B, 2 * H, 2 * W, C + 16, dtype=dtype, device=d)
    ops.bicubic2x_fwd(xb, ybuf[..., :C])
    assert rel(nchw(ybuf[..., :C]), yr) < TOL[dtype]
    gy = torch.randn_like(yr)
    (gx,) = torch.autograd.grad(yr, xr, rnd(gy, dtype))
    dyb = torch.zeros(B, 2 * H, 2 * W, C + 16, dtype=dtype, device=d)
    dyb[..., :C] = nhwc(gy, dtype)
    dx = torch.ones(B, H, W, C, dtype=dtype, device=d)
    ops.bicubic2x_bwd(dyb[..., :C], dx, True)
    assert rel(nchw(dx) - 1, gx) < 2 * TOL[dtype]
    dx2 = torch.full((B, H, W, C + 8), 7.0, dtype=dtype, device=d)
    ops.bicubic2x_bwd(dyb[..., :C], dx2[..., :C], False)
    assert rel(nchw(dx2[..., :C]), gx) < TOL[dtype]
    assert float((dx2[..., C:].float() - 7).abs().max()) == 0 and float(ybuf[..., C:].float().abs().max()) == 0


@pytest.mark.parametrize("dtype", DT)
def test_depth_head_stencil_sigmoid_argmax(dtype):
    from camradepth_b200 import ops
    d = dev()
    torch.manual_seed(6)
    B, C, H, W = 2, 32, 11, 17
    x = torch.rand(B, C, H, W, device=d)
    w = torch.randn(1, C, 3, 3, device=d) * 0.2
    bias = torch.randn(1, device=d)
    xr, wr = rnd(x, dtype).requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, bias, padding=1)
    xb = nhwc(x, dtype)
    wp = torch.zeros(1, 9 * C, device=d)
    ops.weight_pack(w, wp, None, 1, C, 9, C, 8, 0)
    y = torch.empty(B, 1, H, W, device=d)
    ops.conv3x3_c1_fwd(xb, wp, bias, y)
    assert rel(y, yr) < 1e-5
    gy = torch.randn_like(yr)
    gx, gw = torch.autograd.grad(yr, (xr, wr), gy)
    dx = torch.empty(B, H, W, C, dtype=dtype, device=d)
    dw = torch.zeros(1, 9 * C, device=d)
    db = torch.zeros(1, device=d)
    ops.conv3x3_c1_bwd(gy, xb, wp, dx, dw, db)
    gwu = torch.empty_like(w)
    ops.weight_unpack_grad(dw, gwu, None, 1, C, 9, C, False)
    assert rel(nchw(dx), gx) < TOL[dtype] and rel(gwu, gw) < 1e-4 and rel(db, gy.sum().view(1)) < 1e-4
    # fused form used by Depth_Activation: x is a sigmoid output, dx the gradient wrt the sigmoid's input
    dxs = torch.empty(B, H, W, C, dtype=dtype, device=d)
    dw_s, db_s = torch.zeros(1, 9 * C, device=d), torch.zeros(1, device=d)
    ops.conv3x3_c1_bwd_sigmoid(gy, xb, wp, dxs, dw_s, db_s)
    xq = rnd(x, dtype)
    assert rel(nchw(dxs), gx * xq * (1 - xq)) < TOL[dtype]
    assert rel(dw_s, dw) < 1e-6 and rel(db_s, db) < 1e-6
    # a wide, ragged image (W % 4 != 0) through the four-pixels-per-thread forward kernel
    x2 = torch.rand(1, C, 5, 23, device=d)
    y2 = torch.empty(1, 1, 5, 23, device=d)
    ops.conv3x3_c1_fwd(nhwc(x2, dtype), wp, bias, y2)
    assert rel(y2, F.conv2d(rnd(x2, dtype), w, bias, padding=1)) < 1e-5
    # sigmoid backward
    sg = torch.sigmoid(x)
    g = torch.randn_like(x)
    o = torch.empty(B, H, W, C, dtype=dtype, device=d)
    ops.sigmoid_bwd(nhwc(g, dtype), nhwc(sg, dtype), o)
    assert rel(nchw(o), rnd(g, dtype) * rnd(sg, dtype) * (1 - rnd(sg, dtype))) < TOL[dtype]
    # argmax map into a channel of a wider buffer + fp32 plane
    lg = torch.randn(B, 21, H, W, device=d)
    lb = nhwc(lg, dtype, 24)
    buf = torch.zeros(B, H, W, 136, dtype=dtype, device=d)
    plane = torch.empty(B, 1, H, W, device=d)
    ops.argmax_map(lb, 21, buf[..., 129:130], plane)
    ref = torch.argmax(rnd(lg, dtype), 1, keepdim=True) / 21
    assert float((plane != ref).float().mean()) < 1e-3
    assert rel(buf[..., 129].float(), ref[:, 0].to(dtype).float()) < 1e-3
    assert float(buf[..., :129].abs().max()) == 0


def test_weight_pack_batch_matches_per_tensor_pack():
    """One-launch packing (bricks through shared memory) == the per-tensor kernel, all three layouts, with and without
    a channel map, ragged bricks, 7x7 taps."""
    from camradepth_b200 import ops
    d = dev()
    torch.manual_seed(11)
    cases = [(96, 136, 9, 0, False), (136, 96, 9, 1, False), (128, 296, 9, 0, True), (64, 7, 49, 0, False),
             (160, 160, 4, 2, False), (1024, 128, 1, 1, False), (21, 128, 9, 0, False), (33, 70, 9, 1, True)]
    rows, blk, keep, owner = [], 0, [], []
    for (cout, cin, taps, mode, use_map) in cases:
        w = torch.randn(cout, cin, taps, device=d)
        cin_p, cout_p = r8(cin) + (8 if use_map else 0), r8(cout)
        cmap = (torch.randperm(cin_p, device=d)[:cin].to(torch.int32).contiguous()) if use_map else None
        shape = (cout, taps * cin_p) if mode == 0 else ((cin_p, taps * cout_p) if mode == 1 else (taps * cin_p, cout_p))
        for dt in (torch.bfloat16, torch.float32):
            ref = torch.zeros(shape, dtype=dt, device=d)
            ops.weight_pack(w, ref, cmap, cout, cin, taps, cin_p, cout_p, mode)
            dst = torch.zeros(shape, dtype=dt, device=d)
            rows.append([w.data_ptr(), dst.data_ptr(), 0 if cmap is None else cmap.data_ptr(), blk, cout, cin, taps,
                         cin_p, cout_p, mode, ops.dcode(dst), cout * cin * taps])
            nb = ops.weight_pack_blocks(cout, cin, taps)
            owner += [len(rows) - 1] * nb
            blk += nb
            keep.append((w, cmap, ref, dst))
    rows.append([0, 0, 0, blk] + [0] * 8)
    table = torch.tensor([v for r in rows for v in r] + owner, dtype=torch.int64).to(d)
    ops.weight_pack_batch(table, len(rows) - 1, blk)
    torch.cuda.synchronize()
    for (_, _, ref, dst) in keep:
        assert torch.equal(ref, dst)


def test_zero_channels():
    from camradepth_b200 import ops
    d = dev()
    for dt in DT:
        for (ld, a, b) in ((136, 128, 136), (136, 129, 136), (24, 21, 24), (320, 296, 320)):
            t = torch.full((3, 5, 7, ld), 3.0, dtype=dt, device=d)
            ops.zero_channels(t[..., a:b])
            assert float(t[..., a:b].abs().max()) == 0 and float((t[..., :a] - 3).abs().max()) == 0
            if b < ld:
                assert float((t[..., b:] - 3).abs().max()) == 0


def test_layout_roundtrip():
    from camradepth_b200 import ops
    d = dev()
    x = torch.randn(2, 7, 12, 20, device=d)
    for dtype in DT:
        buf = torch.zeros(2, 12, 20, 16, dtype=dtype, device=d)
        ops.nchw_to_nhwc(x, buf[..., 4:11])
        assert torch.equal(buf[..., 4:11].float(), x.permute(0, 2, 3, 1).to(dtype).float())
        assert float(buf[..., :4].abs().max()) == 0 and float(buf[..., 11:].abs().max()) == 0
        back = torch.empty_like(x)
        ops.nhwc_to_nchw(buf[..., 4:11], back)
        assert torch.equal(back, x.to(dtype).float())


def test_losses_match_oracle():
    from camradepth_b200 import MaskedSmoothL1Loss, MaskedFocalLoss, MaskedMSELoss
    from oracle import camradepth_oracle as O
    d = dev()
    torch.manual_seed(7)
    pred = (torch.randn(2, 1, 24, 40, device=d) * 1.5).requires_grad_(True)
    tgt = torch.rand(2, 1, 24, 40, device=d) * (torch.rand(2, 1, 24, 40, device=d) < 0.3)
    l = MaskedSmoothL1Loss()(pred, tgt)
    pr = pred.detach().clone().requires_grad_(True)
    lr = O.masked_smooth_l1(pr, tgt)
    assert abs(float(l) - float(lr)) < 1e-6 * max(1, abs(float(lr)))
    (l * 1.7).backward()
    (lr * 1.7).backward()
    assert rel(pred.grad, pr.grad) < 1e-5
    assert abs(float(MaskedMSELoss()(pred, tgt)) - float(O.masked_mse(pr, tgt))) < 1e-5
    lg = (torch.randn(2, 21, 24, 40, device=d) * 2).requires_grad_(True)
    t = torch.randint(0, 21, (2, 24, 40), device=d)
    t[torch.rand(2, 24, 40, device=d) < 0.1] = 255
    f = MaskedFocalLoss()(lg, t)
    lgr = lg.detach().clone().requires_grad_(True)
    fr = O.masked_focal(lgr, t)
    assert abs(float(f) - float(fr)) < 1e-5 * max(1, abs(float(fr)))
    (f * 0.2).backward()
    (fr * 0.2).backward()
    assert rel(lg.grad, lgr.grad) < 1e-4


def test_diffgradnorm_matches_oracle():
    from camradepth_b200 import diffGradNorm
    from oracle import camradepth_oracle as O
    d = dev()
    torch.manual_seed(8)
    shapes = [(64, 64, 1), (40000,), (3, 5, 3, 3), (1,), (128, 296, 3, 3)]
    ps = [torch.nn.Parameter(torch.randn(s, device=d)) for s in shapes]
    ref = [p.detach().cpu().clone() for p in ps]
    states = [dict() for _ in ps]
    opt = diffGradNorm(ps, lr=6e-5)
    for it in range(4):
        scale = [1.0, 0.3, 2.0, 0.1][it]          # shrinking grads exercise the norm-correction branch
        for p, r, st in zip(ps, ref, states):
            g = torch.randn(p.shape) * scale
            p.grad = g.to(d)
            O.diffgradnorm_step(r, g, st)
        opt.step()
    torch.cuda.synchronize()
    for p, r, r0 in zip(ps, ref, shapes):
        assert rel(p.detach().cpu(), r) < 1e-6
    for p, st in zip(ps, states):
        assert abs(float(opt.state[p]["exp_grad_norm"]) - float(st["exp_grad_norm"])) < 1e-4 * float(st["exp_grad_norm"])


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("case", [(2, 8, 32, 64, 7, 4, 3), (2, 64, 16, 24, 3, 2, 1), (2, 64, 16, 32, 8, 8, 0), (1, 160, 12, 26, 2, 2, 0)])
def test_im2col_col2im(case, dtype):
    """im2col is a pure gather; col2im must be its exact adjoint (<col2im(g), x> == <g, im2col(x)>)."""
    from camradepth_b200 import ops
    B, C, H, W, k, stride, pad = case
    d = dev()
    torch.manual_seed(9)
    x = torch.randn(B, C, H, W, device=d)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    xb = nhwc(x, dtype)
    col = torch.empty(B, Ho, Wo, k * k * C, dtype=dtype, device=d)
    ops.im2col(xb, col, C, k, k, stride, pad)
    ref = F.unfold(rnd(x, dtype), k, padding=pad, stride=stride)            # (B, C*k*k, L) with (c, kh, kw) order
    ref = ref.view(B, C, k * k, Ho, Wo).permute(0, 3, 4, 2, 1).reshape(B, Ho, Wo, k * k * C)
    assert torch.equal(col.float(), ref)
    g = torch.randn(B, Ho, Wo, k * k * C, device=d).to(dtype)
    dx = torch.ones(B, H, W, C, dtype=dtype, device=d)
    ops.col2im(g, dx, C, k, k, stride, pad, True)
    gr = g.float().view(B, Ho * Wo, k * k, C).permute(0, 3, 2, 1).reshape(B, C * k * k, Ho * Wo)
    fold = F.fold(gr, (H, W), k, padding=pad, stride=stride)
    assert rel(nchw(dx) - 1, fold) < (1e-5 if dtype == torch.float32 else 2e-2)
