"""Typed Python wrappers over the C-ABI kernels (one function per entry point).

Tensors are torch CUDA tensors used purely as device memory: the wrappers pass `data_ptr()`s, sizes and
the current CUDA stream to the library.  Activations are NHWC views; `ld` is taken from the view's pixel
stride so channel slices of concat buffers can be passed directly.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import K, ConvDesc, load

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_SIGMOID = 0, 1, 2
OPT_CHUNK = 16384


def dcode(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def P(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def _ld(t: torch.Tensor) -> int:
    """pixel stride of an NHWC / (B,N,C) view; checks the view is pixel-contiguous."""
    assert t.stride(-1) == 1, "channel dim must be contiguous"
    ld = t.stride(-2)
    if t.dim() == 4:
        B, H, W, _ = t.shape
        assert (W == 1 or t.stride(2) == ld) and (H == 1 or t.stride(1) == W * ld) and \
               (B == 1 or t.stride(0) == H * W * ld), f"not an NHWC view: {t.shape} {t.stride()}"
    elif t.dim() == 3:
        B, N, _ = t.shape
        assert B == 1 or t.stride(0) == N * ld, f"not a token view: {t.shape} {t.stride()}"
    return ld


def launch_count() -> int:
    return int(load().crd_launch_count())


# ------------------------------------------------------------------ layout
def nchw_to_nhwc(src, dst):
    """src f32 (B,C,H,W) contiguous -> dst[..., :C] of an NHWC view."""
    B, C, H, W = src.shape
    assert src.is_contiguous() and src.dtype == torch.float32
    K.crd_nchw_to_nhwc(P(src), P(dst), dcode(dst), B, C, H, W, _ld(dst), stream())


def nhwc_to_nchw(src, dst):
    B, C, H, W = dst.shape
    assert dst.is_contiguous() and dst.dtype == torch.float32
    K.crd_nhwc_to_nchw(P(src), dcode(src), P(dst), B, C, H, W, _ld(src), stream())


def copy_channels(src, dst):
    """dst[..., :C] = src[..., :C] for channel-slice views of NHWC buffers (C = src.shape[-1], any alignment)."""
    assert src.dtype == dst.dtype and src.shape == dst.shape
    C = src.shape[-1]
    K.crd_copy_channels(P(src), _ld(src), P(dst), _ld(dst), dcode(src), C, src.numel() // C, stream())


def make_masks(out, keep_dp, n_dp, B, n_d2, C2, keep_d2, state):
    K.crd_make_masks(P(out), P(keep_dp), n_dp, B, n_d2, C2, float(keep_d2), P(state), stream())


# ------------------------------------------------------------------ conv
def make_desc(x, y, Cin, Cout, KH, KW, stride, pad, transposed=0, act=0, accumulate=0, out_nchw=0,
              out_dtype=None, in_dtype=None, w_tap_stride=0, w_koff=0):
    d = ConvDesc()
    d.B, d.H, d.W = x.shape[0], x.shape[1], x.shape[2]
    d.Cin, d.ldx = Cin, _ld(x)
    if out_nchw:
        d.Ho, d.Wo = y.shape[2], y.shape[3]
        d.ldy = 0
    else:
        d.Ho, d.Wo = y.shape[1], y.shape[2]
        d.ldy = _ld(y)
    d.Cout = Cout
    d.KH, d.KW, d.stride, d.pad = KH, KW, stride, pad
    d.transposed = transposed
    d.in_dtype = dcode(x) if in_dtype is None else in_dtype
    d.out_dtype = dcode(y) if out_dtype is None else out_dtype
    d.act, d.accumulate, d.out_nchw = act, accumulate, out_nchw
    d.w_tap_stride, d.w_koff = w_tap_stride, w_koff
    return d


def conv_fwd(desc, x, w, bias, y, use_tc=False, gn_sums=None):
    if use_tc:
        K.crd_conv_fwd_tc(ctypes.byref(desc), P(x), P(w), P(bias), P(y), P(gn_sums), stream())
    else:
        K.crd_conv_fwd(ctypes.byref(desc), P(x), P(w), P(bias), P(y), stream())


def conv_argmax(desc, x, w, bias, ncls, map0=None, map1=None, map_f32=None):
    """tcgen05 conv whose read-out writes argmax_c / ncls into channel views map0 / map1 (B,H,W,1 slices of NHWC
    bf16 buffers) and / or an fp32 (B,1,H,W) tensor; the logits are not stored."""
    for m in (map0, map1):
        assert m is None or m.dtype == torch.bfloat16
    K.crd_conv_argmax_tc(ctypes.byref(desc), P(x), P(w), P(bias), ncls, P(map0), 0 if map0 is None else _ld(map0),
                         P(map1), 0 if map1 is None else _ld(map1), P(map_f32), stream())


def conv_wgrad(desc, x, dy, dw, use_tc=False, db=None):
    """db (tensor-core 1x1 path only): the bias gradient is accumulated by the same kernel."""
    if use_tc and db is not None:
        K.crd_conv_wgrad_bias_tc(ctypes.byref(desc), P(x), P(dy), P(dw), P(db), stream())
    elif use_tc:
        K.crd_conv_wgrad_tc(ctypes.byref(desc), P(x), P(dy), P(dw), stream())
    else:
        K.crd_conv_wgrad(ctypes.byref(desc), P(x), P(dy), P(dw), stream())


def weight_pack(w, dst, cmap, Cout, Cin, taps, Cin_p, Cout_p, mode):
    K.crd_weight_pack(P(w), P(dst), dcode(dst), P(cmap), Cout, Cin, taps, Cin_p, Cout_p, mode, stream())


def zero_channels(view):
    """Zero a channel slice view[..., a:b] of an NHWC buffer (last dim contiguous)."""
    C = view.shape[-1]
    K.crd_zero_channels(P(view), _ld(view), dcode(view), C, view.numel() // max(C, 1), stream())


def weight_pack_blocks(cout, cin, taps):
    """Blocks one item of a weight_pack_batch table owns."""
    n = load().crd_weight_pack_blocks(cout, cin, taps)
    if n <= 0:
        raise ValueError(f"weight_pack_batch cannot take Cout={cout} Cin={cin} taps={taps}")
    return n


def weight_pack_batch(table, n_items, n_blocks):
    K.crd_weight_pack_batch(P(table), n_items, n_blocks, stream())


def weight_unpack_grad(dwp, grad, cmap, Cout, Cin, taps, Cin_p, accumulate):
    K.crd_weight_unpack_grad(P(dwp), P(grad), P(cmap), Cout, Cin, taps, Cin_p, int(accumulate), stream())


def im2col(x, col, Cin, KH, KW, stride, pad):
    """x NHWC view (B,H,W,>=Cin) -> col (B,Ho,Wo,KH*KW*Cin) contiguous"""
    B, H, W, _ = x.shape
    _, Ho, Wo, _ = col.shape
    assert col.is_contiguous() and col.shape[-1] == KH * KW * Cin
    K.crd_im2col(P(x), P(col), dcode(x), B, H, W, Cin, _ld(x), Ho, Wo, KH, KW, stride, pad, stream())


def col2im(dcol, dx, Cin, KH, KW, stride, pad, accumulate):
    B, H, W, _ = dx.shape
    _, Ho, Wo, _ = dcol.shape
    assert dcol.is_contiguous()
    K.crd_col2im(P(dcol), P(dx), dcode(dx), int(accumulate), B, H, W, Cin, _ld(dx), Ho, Wo, KH, KW, stride, pad,
                 stream())


def col_sum(dy, db, N):
    """db[:N] += column sums of the (M, ld) matrix behind the NHWC view dy."""
    M = dy.numel() // dy.shape[-1]
    K.crd_col_sum(P(dy), dcode(dy), P(db), M, N, _ld(dy), stream())


# ------------------------------------------------------------------ groupnorm protocol
def _bnc(x):
    B = x.shape[0]
    C = x.shape[-1]
    N = x.numel() // (B * C)
    return B, N, C


def chan_stats(x, sums):
    B, N, C = _bnc(x)
    K.crd_chan_stats(P(x), dcode(x), P(sums), B, N, C, _ld(x), stream())


def gn_finalize(sums, gamma, beta, ab, mean_rstd, xbar, B, C, G, N, eps=1e-5):
    K.crd_gn_finalize(P(sums), P(gamma), P(beta), P(ab), P(mean_rstd), P(xbar), B, C, G, N, eps, stream())


def affine_act(x, y, ab, post, act):
    B, N, C = _bnc(x)
    K.crd_affine_act(P(x), dcode(x), P(y), dcode(y), P(ab), P(post), act, B, N, C, _ld(x), _ld(y), stream())


def gnact_bwd_reduce(dy, x, ab, post, addbc, act, pq, dz_out=None):
    B, N, C = _bnc(x)
    if dz_out is not None:
        assert dz_out.dtype == dy.dtype and _ld(dz_out) == _ld(dy)
    K.crd_gnact_bwd_reduce(P(dy), dcode(dy), P(x), dcode(x), P(ab), P(post), P(addbc), act, P(pq), P(dz_out), B, N,
                           C, _ld(dy), _ld(x), stream())


def gn_bwd_finalize(pq, mean_rstd, gamma, coef, dgamma, dbeta, B, C, G, N):
    K.crd_gn_bwd_finalize(P(pq), P(mean_rstd), P(gamma), P(coef), P(dgamma), P(dbeta), B, C, G, N, stream())


def gnact_bwd_apply(dy, x, ab, post, addbc, act, coef, dx, accumulate):
    B, N, C = _bnc(x)
    K.crd_gnact_bwd_apply(P(dy), dcode(dy), P(x), dcode(x), P(ab), P(post), P(addbc), act, P(coef), P(dx),
                          dcode(dx), int(accumulate), B, N, C, _ld(dy), _ld(x), _ld(dx), stream())


def gn_fused_supported(B, N, C, G):
    return bool(load().crd_gn_fused_supported(B, N, C, G))


def gn_fused_fwd(x, y, gamma, beta, G, sums, post, act, ab, mean_rstd, xbar, eps=1e-5):
    """One-launch GroupNorm forward (statistics + finalize + apply); y may be None (finalize only)."""
    B, N, C = _bnc(x)
    K.crd_gn_fused_fwd(P(x), dcode(x), P(y), dcode(y) if y is not None else 0, P(gamma), P(beta), P(sums), P(post), act,
                       P(ab), P(mean_rstd), P(xbar), B, N, C, G, _ld(x), _ld(y) if y is not None else 8, eps, stream())


def gn_fused_bwd(dy, x, ab, mean_rstd, gamma, G, post, addbc, act, dx, accumulate, dgamma, dbeta):
    B, N, C = _bnc(x)
    K.crd_gn_fused_bwd(P(dy), dcode(dy), P(x), dcode(x), P(ab), P(mean_rstd), P(gamma), P(post), P(addbc), act, P(dx),
                       dcode(dx), int(accumulate), P(dgamma), P(dbeta), B, N, C, G, _ld(dy), _ld(x), _ld(dx), stream())


# ------------------------------------------------------------------ encoder pieces
def dwconv_fwd(x, ab, w, bias, y):
    B, H, W, C = x.shape
    assert x.is_contiguous() and y.is_contiguous()
    K.crd_dwconv3x3_fwd(P(x), dcode(x), P(ab), P(w), P(bias), P(y), B, H, W, C, stream())


def dwconv_bwd_input(dy, w, dxn):
    B, H, W, C = dy.shape
    assert dy.is_contiguous() and dxn.is_contiguous()
    K.crd_dwconv3x3_bwd_input(P(dy), dcode(dy), P(w), P(dxn), B, H, W, C, stream())


def dwconv_bwd_weight(dy, x, ab, dw, db):
    B, H, W, C = dy.shape
    assert dy.is_contiguous() and x.is_contiguous()
    K.crd_dwconv3x3_bwd_weight(P(dy), dcode(dy), P(x), P(ab), P(dw), P(db), B, H, W, C, stream())


def dwconv_bwd(dy, x, ab, w, dxn, dw, db):
    B, H, W, C = dy.shape
    assert dy.is_contiguous() and x.is_contiguous() and dxn.is_contiguous()
    K.crd_dwconv3x3_bwd(P(dy), dcode(dy), P(x), P(ab), P(w), P(dxn), P(dw), P(db), B, H, W, C, stream())


def attn_qkmax_fwd(q, k, s, idx, heads, scale):
    B, N, C = q.shape
    M = k.shape[1]
    assert q.is_contiguous() and k.is_contiguous()
    K.crd_attn_qkmax_fwd(P(q), P(k), dcode(q), P(s), P(idx), B, N, M, C, heads, scale, stream())


def attn_qkmax_bwd(ds, q, k, idx, dq, dk, heads, scale):
    B, N, C = q.shape
    M = k.shape[1]
    K.crd_attn_qkmax_bwd(P(ds), P(q), P(k), dcode(q), P(idx), P(dq), P(dk), B, N, M, C, heads, scale, stream())


def attn_pv_fwd(xbar, Wp, pv):
    B, C = xbar.shape
    K.crd_attn_pv_fwd(P(xbar), P(Wp), P(pv), B, C, stream())


def attn_pv_bwd(dpv, xbar, Wp, dWp, dxbar, scale):
    B, C = xbar.shape
    K.crd_attn_pv_bwd(P(dpv), P(xbar), P(Wp), P(dWp), P(dxbar), scale, B, C, stream())


def attn_out_residual(x, pv, s, bp, dp, xout):
    B, N, C = x.shape
    K.crd_attn_out_residual(P(x), P(pv), P(s), P(bp), P(dp), P(xout), B, N, C, stream())


def attn_out_bwd(dx, pv, s, dp, ds, dpv, dbp, tmp):
    B, N, C = dx.shape
    K.crd_attn_out_bwd(P(dx), P(pv), P(s), P(dp), P(ds), P(dpv), P(dbp), P(tmp), B, N, C, stream())


def residual_add(x, y, dp, xout):
    B, N, C = x.shape
    assert y.is_contiguous()
    K.crd_residual_add(P(x), P(y), dcode(y), P(dp), P(xout), B, N, C, stream())


def scale_cast(dx, dp, dy):
    B = dx.shape[0]
    C = dx.shape[-1]
    N = dx.numel() // (B * C)
    assert dx.is_contiguous() and dy.is_contiguous() and dx.dtype == torch.float32
    K.crd_scale_cast(P(dx), P(dp), P(dy), dcode(dy), B, N, C, stream())


def add_f32(dst, src):
    assert dst.is_contiguous() and src.is_contiguous() and dst.dtype == torch.float32
    K.crd_add_f32(P(dst), P(src), dcode(src), dst.numel(), stream())


# ------------------------------------------------------------------ decoder pieces
def bicubic2x_fwd(x, y):
    B, H, W, C = x.shape
    K.crd_bicubic2x_fwd(P(x), P(y), dcode(x), B, H, W, C, _ld(x), _ld(y), stream())


def bicubic2x_bwd(dy, dx, accumulate):
    B, H, W, C = dx.shape
    K.crd_bicubic2x_bwd(P(dy), P(dx), dcode(dx), int(accumulate), B, H, W, C, _ld(dy), _ld(dx), stream())


def conv3x3_c1_fwd(x, w, bias, y):
    B, H, W, C = x.shape
    K.crd_conv3x3_c1_fwd(P(x), dcode(x), P(w), P(bias), P(y), B, H, W, C, _ld(x), stream())


def conv3x3_c1_bwd(dy, x, w, dx, dw, db):
    B, H, W, C = x.shape
    K.crd_conv3x3_c1_bwd(P(dy), P(x), dcode(x), P(w), P(dx), P(dw), P(db), B, H, W, C, _ld(x),
                         _ld(dx) if dx is not None else 8, stream())


def conv3x3_c1_bwd_sigmoid(dy, x, w, dx, dw, db):
    """conv3x3_c1_bwd whose dx is the gradient wrt the INPUT of the sigmoid that produced x (dgrad * x * (1 - x))."""
    B, H, W, C = x.shape
    K.crd_conv3x3_c1_bwd_sigmoid(P(dy), P(x), dcode(x), P(w), P(dx), P(dw), P(db), B, H, W, C, _ld(x), _ld(dx),
                                 stream())


def sigmoid_bwd(dy, y, dx):
    assert dy.is_contiguous() and y.is_contiguous() and dx.is_contiguous()
    K.crd_sigmoid_bwd(P(dy), P(y), P(dx), dcode(y), y.numel(), stream())


def argmax_map(logits, ncls, dst, dst_f32):
    npix = logits.numel() // logits.shape[-1]
    K.crd_argmax_map(P(logits), dcode(logits), _ld(logits), ncls, P(dst), dcode(dst) if dst is not None else 0,
                     _ld(dst) if dst is not None else 0, P(dst_f32), npix, stream())


# ------------------------------------------------------------------ losses / optimizer
def masked_l1_fwd(pred, target, acc):
    K.crd_masked_l1_fwd(P(pred), P(target), P(acc), pred.numel(), stream())


def masked_l1_bwd(pred, target, acc, gout, dpred):
    K.crd_masked_l1_bwd(P(pred), P(target), P(acc), P(gout), P(dpred), pred.numel(), stream())


def ce_fwd(logits, target, acc, ignore_index=255):
    B, C = logits.shape[0], logits.shape[1]
    HW = logits.numel() // (B * C)
    K.crd_ce_fwd(P(logits), P(target), P(acc), B, C, HW, ignore_index, stream())


def ce_bwd(logits, target, acc, gout, gamma, dlogits, ignore_index=255):
    B, C = logits.shape[0], logits.shape[1]
    HW = logits.numel() // (B * C)
    K.crd_ce_bwd(P(logits), P(target), P(acc), P(gout), gamma, P(dlogits), B, C, HW, ignore_index, stream())


def loss_finalize(acc, out, kind, gamma=0.0):
    K.crd_loss_finalize(P(acc), P(out), kind, gamma, stream())


def mt_sumsq(table, chunks, nchunks, sumsq):
    K.crd_mt_sumsq(P(table), P(chunks), nchunks, P(sumsq), stream())


def diffgradnorm_update(table, chunks, nchunks, sumsq, egn_in, egn_out, step_size, beta1, beta2, eps):
    K.crd_diffgradnorm_update(P(table), P(chunks), nchunks, P(sumsq), P(egn_in), P(egn_out), P(step_size), beta1,
                              beta2, eps, stream())
