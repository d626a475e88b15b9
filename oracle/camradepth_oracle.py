"""CPU oracle for the CamRaDepth hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain, functional fp32/fp64 restatement (torch CPU ops used as the array library)
of the reference algorithm on the path BASELINE.json names.  Each function cites the
reference file:line (relative to /root/reference) it follows.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file; the product package `camradepth_b200` never does.

Parity pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference module itself, imported in the build
container by `oracle/make_golden.py` and committed under `tests/golden/`
(`tests/test_oracle_golden.py` checks them).

Layout: everything here is the reference layout (NCHW / (B,C,N)), state_dict keys and
shapes are the reference's.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- config

DEPTHS = (3, 10, 16, 5)          # CamRaDepth.py:27
DIMS = (64, 128, 160, 256)       # CamRaDepth.py:28
HEADS = (1, 2, 4, 8)             # CamRaDepth.py:24
FF = (8, 8, 4, 4)                # CamRaDepth.py:25
SR = (8, 4, 2, 1)                # CamRaDepth.py:26
MID = 128                        # CamRaDepth.py:37
GN_DIV = 16                      # args.py:37
DROP_PATH_RATE = 0.1             # CamRaDepth.py:57
DROPOUT2D_P = 0.2                # CamRaDepth.py:96

VARIANTS = {
    # name: (supervised_seg, unsupervised_seg, input_channels)      args.py:156-168
    "base": (False, False, 7),
    "base (rgb)": (False, False, 3),
    "supervised_seg": (True, False, 7),
    "unsupervised_seg": (False, True, 7),
    "sup_unsup_seg": (True, True, 7),
    "sup_unsup_seg (rgb)": (True, True, 3),
}


class Cfg:
    def __init__(self, variant="base", depths=DEPTHS, dims=DIMS, heads=HEADS, ff=FF, sr=SR,
                 num_classes=21, input_channels=None):
        sup, unsup, cin = VARIANTS[variant]
        self.variant = variant
        self.sup, self.unsup = sup, unsup
        self.cin = cin if input_channels is None else input_channels
        self.depths, self.dims, self.heads, self.ff, self.sr = depths, dims, heads, ff, sr
        self.num_classes = num_classes


# ----------------------------------------------------------------------------- parameters

def param_spec(cfg: Cfg) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """name -> (shape, init kind), in the reference's registration order.

    Order/shape source: simplified_attention.py:190-246 (encoder), CamRaDepth.py:53-94,
    utils.py:103-124,201-221,274-283.  Init kinds (SURVEY.md §8b "Init"):
      tn02   trunc_normal_(std=.02)                    simplified_attention.py:28-32,85-88
      fanout N(0, sqrt(2/fan_out)), fan_out=kh*kw*Cout/groups   :79-84,134-139,176-181
      kaiming kaiming_normal_(fan_out, relu)           utils.py:309-313
      default torch Conv2d default (kaiming_uniform a=sqrt(5)); bias U(+-1/sqrt(fan_in))
      ones / zeros
    """
    sp: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()
    dims, depths, ff, sr = cfg.dims, cfg.depths, cfg.ff, cfg.sr
    cin = cfg.cin
    # patch embeds are registered first (simplified_attention.py:204-211)
    pe_in = (cin, dims[0], dims[1], dims[2])
    pe_k = (7, 3, 3, 3)
    for s in range(4):
        p = f"dest_encoder.patch_embed{s + 1}"
        sp[p + ".proj.weight"] = ((dims[s], pe_in[s], pe_k[s], pe_k[s]), "fanout")
        sp[p + ".proj.bias"] = ((dims[s],), "zeros")
        sp[p + ".norm.weight"] = ((dims[s],), "ones")
        sp[p + ".norm.bias"] = ((dims[s],), "zeros")
    for s in range(4):
        C, rC = dims[s], dims[s] * ff[s]
        for i in range(depths[s]):
            p = f"dest_encoder.block{s + 1}.{i}"
            sp[p + ".norm1.weight"] = ((C,), "ones")
            sp[p + ".norm1.bias"] = ((C,), "zeros")
            sp[p + ".norm2.weight"] = ((C,), "ones")
            sp[p + ".norm2.bias"] = ((C,), "zeros")
            sp[p + ".attn.q.weight"] = ((C, C, 1), "tn02")
            sp[p + ".attn.q.bias"] = ((C,), "zeros")
            sp[p + ".attn.k.weight"] = ((C, C, 1), "tn02")
            sp[p + ".attn.k.bias"] = ((C,), "zeros")
            sp[p + ".attn.proj.weight"] = ((C, C, 1), "tn02")
            sp[p + ".attn.proj.bias"] = ((C,), "zeros")
            if sr[s] > 1:
                sp[p + ".attn.sr.weight"] = ((C, C, sr[s], sr[s]), "fanout")
                sp[p + ".attn.sr.bias"] = ((C,), "zeros")
                sp[p + ".attn.norm.weight"] = ((C,), "ones")
                sp[p + ".attn.norm.bias"] = ((C,), "zeros")
            sp[p + ".mlp1.fc1.weight"] = ((rC, C, 1), "tn02")
            sp[p + ".mlp1.fc1.bias"] = ((rC,), "zeros")
            sp[p + ".mlp1.dwconv.dwconv.weight"] = ((rC, 1, 3, 3), "fanout_dw")
            sp[p + ".mlp1.dwconv.dwconv.bias"] = ((rC,), "zeros")
            sp[p + ".mlp1.fc2.weight"] = ((C, rC, 1), "tn02")
            sp[p + ".mlp1.fc2.bias"] = ((C,), "zeros")
            sp[p + ".mlp1.norm1.weight"] = ((rC,), "ones")
            sp[p + ".mlp1.norm1.bias"] = ((rC,), "zeros")
            sp[p + ".mlp1.norm2.weight"] = ((rC,), "ones")
            sp[p + ".mlp1.norm2.bias"] = ((rC,), "zeros")
    for j, C in enumerate((dims[3], dims[2], dims[1], dims[0])):
        p = f"from_encoder_{j + 1}.model"
        sp[p + ".0.weight"] = ((C, C, 1, 1), "kaiming")
        sp[p + ".1.weight"] = ((C,), "ones")
        sp[p + ".1.bias"] = ((C,), "zeros")

    def short_res(prefix, cin_):
        # utils.py:107-124: out 96, 64, out_channels(=128); dense concatenation
        inp = cin_
        for li, out in enumerate((int(MID * 0.75), int(MID * 0.5), MID)):
            sp[f"{prefix}.conv.layers.{li}.model.0.weight"] = ((out, inp, 3, 3), "kaiming")
            sp[f"{prefix}.conv.layers.{li}.model.1.weight"] = ((out,), "ones")
            sp[f"{prefix}.conv.layers.{li}.model.1.bias"] = ((out,), "zeros")
            inp += out

    dec_in = (dims[3] + dims[2], MID + dims[1], MID + dims[0], MID + 1, MID + 1 + cin)
    for d in range(5):
        short_res(f"depth_upsample.{d}", dec_in[d])
    nseg = int(cfg.sup) + int(cfg.unsup)
    for name, c in (("depth_activation_3", MID), ("depth_activation_4", MID + nseg),
                    ("depth_activation_5", MID + nseg)):
        sp[name + ".conv_1.weight"] = ((32, c, 3, 3), "default")
        sp[name + ".conv_1.bias"] = ((32,), "default_bias")
        sp[name + ".conv_2.weight"] = ((1, 32, 3, 3), "default")
        sp[name + ".conv_2.bias"] = ((1,), "default_bias")
    if cfg.sup or cfg.unsup:
        short_res("seg_upsample.0", MID + 1)
        short_res("seg_upsample.1", MID + 1 + cin)
    if cfg.sup:
        for n in ("seg_conv_stage_4", "seg_conv_final"):
            sp[n + ".weight"] = ((cfg.num_classes, MID, 3, 3), "default")
            sp[n + ".bias"] = ((cfg.num_classes,), "default_bias")
    if cfg.unsup:
        for n in ("unsup_stage_4", "unsup_final"):
            sp[n + ".weight"] = ((19, MID, 3, 3), "default")
            sp[n + ".bias"] = ((19,), "default_bias")
    return sp


def init_state_dict(cfg: Cfg, seed: int = 0, dtype=torch.float32, perturb: float = 0.0):
    """Seeded state_dict with the reference's init distributions.

    `perturb` > 0 adds N(0, perturb) noise to biases / affine params so parity tests exercise
    them (the reference initialises them to exactly 0 / 1).
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    spec = param_spec(cfg)
    for name, (shape, kind) in spec.items():
        if kind == "zeros":
            t = torch.zeros(shape)
        elif kind == "ones":
            t = torch.ones(shape)
        elif kind == "tn02":
            t = torch.empty(shape)
            # trunc_normal_(std=.02, a=-2, b=2): bounds are +-100 sigma, i.e. plain normal
            t.normal_(0, 0.02, generator=g).clamp_(-2, 2)
        elif kind in ("fanout", "fanout_dw"):
            fan_out = shape[2] * shape[3] * shape[0]
            if kind == "fanout_dw":
                fan_out //= shape[0]
            t = torch.empty(shape).normal_(0, math.sqrt(2.0 / fan_out), generator=g)
        elif kind == "kaiming":
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.empty(shape).normal_(0, math.sqrt(2.0 / fan_out), generator=g)
        elif kind == "default":
            fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / math.sqrt(fan_in)      # kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "default_bias":
            wshape = spec[name.replace(".bias", ".weight")][0]
            bound = 1.0 / math.sqrt(wshape[1] * wshape[2] * wshape[3])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            raise ValueError(kind)
        if perturb > 0 and kind in ("zeros", "ones"):
            t = t + torch.randn(shape, generator=g) * perturb
        sd[name] = t.to(dtype)
    return sd


# ----------------------------------------------------------------------------- encoder

def _gn(x, sd, prefix, groups):
    return F.group_norm(x, groups, sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def patch_embed(sd, p, x, k, stride):
    """OverlapPatchEmbed.forward (simplified_attention.py:183-188)."""
    x = F.conv2d(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"], stride=stride, padding=k // 2)
    _, C, H, W = x.shape
    x = _gn(x, sd, p + ".norm", C // GN_DIV)
    return x.flatten(2), H, W


def attention_maxpool(sd, p, x, H, W, heads, sr):
    """Attention_MaxPool.forward (simplified_attention.py:90-109)."""
    B, C, N = x.shape
    q = F.conv1d(x, sd[p + ".q.weight"], sd[p + ".q.bias"])
    q = q.reshape(B, heads, C // heads, N).permute(0, 1, 3, 2)
    if sr > 1:
        x_ = x.reshape(B, C, H, W)
        x_ = F.conv2d(x_, sd[p + ".sr.weight"], sd[p + ".sr.bias"], stride=sr).reshape(B, C, -1)
        x_ = _gn(x_, sd, p + ".norm", C // GN_DIV)
        k = F.conv1d(x_, sd[p + ".k.weight"], sd[p + ".k.bias"]).reshape(B, heads, C // heads, -1)
    else:
        k = F.conv1d(x, sd[p + ".k.weight"], sd[p + ".k.bias"]).reshape(B, heads, C // heads, -1)
    v = torch.mean(x, 2, True).repeat(1, 1, heads).transpose(-2, -1)      # (B, heads, C)
    scale = (C // heads) ** -0.5
    attn = (q @ k) * scale                                                # (B, heads, N, M)
    attn, _ = torch.max(attn, -1)                                         # (B, heads, N)
    out = attn.transpose(-2, -1) @ v                                      # (B, N, C)
    out = out.transpose(-2, -1)
    return F.conv1d(out, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def mlp(sd, p, x, H, W, C):
    """Mlp.forward (simplified_attention.py:34-43) + DWConv.forward (:318-323)."""
    x = F.conv1d(x, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"])
    B, rC, N = x.shape
    x = _gn(x, sd, p + ".norm1", rC // GN_DIV)
    x = F.conv2d(x.reshape(B, rC, H, W), sd[p + ".dwconv.dwconv.weight"], sd[p + ".dwconv.dwconv.bias"],
                 padding=1, groups=rC).flatten(2)
    x = _gn(x, sd, p + ".norm2", C // GN_DIV)        # groups from OUT features (:24)
    x = F.gelu(x)
    return F.conv1d(x, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])


def block(sd, p, x_orig, H, W, C, heads, sr, dp_attn=None, dp_mlp=None):
    """Block.forward (simplified_attention.py:141-145).  `self.drop_path` is called twice (:143,:144) and timm's
    drop_path draws an independent per-sample Bernoulli mask on every call, so the two residual branches get
    their own (B,) DropPath scales: dp_attn for the attention branch, dp_mlp for the Mix-FFN branch."""
    def dp(t, sc):
        return t if sc is None else t * sc.view(-1, 1, 1).to(t.dtype)
    x = _gn(x_orig, sd, p + ".norm1", C // GN_DIV)
    x = x_orig + dp(attention_maxpool(sd, p + ".attn", x, H, W, heads, sr), dp_attn)
    x = x + dp(mlp(sd, p + ".mlp1", _gn(x, sd, p + ".norm2", C // GN_DIV), H, W, C), dp_mlp)
    return x


def encoder(sd, cfg: Cfg, x, drop_path_scales: Optional[Sequence] = None):
    """SimplifiedTransformer.forward_features (simplified_attention.py:265-306)."""
    B = x.shape[0]
    outs = []
    pe_k, pe_s = (7, 3, 3, 3), (4, 2, 2, 2)
    bi = 0
    for s in range(4):
        x, H, W = patch_embed(sd, f"dest_encoder.patch_embed{s + 1}", x, pe_k[s], pe_s[s])
        for i in range(cfg.depths[s]):
            # two scales per block, in call order: attention branch, then Mix-FFN branch
            sa = None if drop_path_scales is None else drop_path_scales[2 * bi]
            sm = None if drop_path_scales is None else drop_path_scales[2 * bi + 1]
            x = block(sd, f"dest_encoder.block{s + 1}.{i}", x, H, W, cfg.dims[s], cfg.heads[s], cfg.sr[s], sa, sm)
            bi += 1
        x = x.reshape(B, -1, H, W).contiguous()
        outs.append(x)
    return outs


# ----------------------------------------------------------------------------- decoder

def conv_layer(sd, p, x, pad):
    """ConvLayer.forward (utils.py:223-228): conv(no bias) -> GroupNorm(Cout/16) -> GELU."""
    w = sd[p + ".model.0.weight"]
    x = F.conv2d(x, w, None, padding=pad)
    x = _gn(x, sd, p + ".model.1", w.shape[0] // GN_DIV)
    return F.gelu(x)


def short_res_block(sd, p, x):
    """ShortResBlock.forward (utils.py:127-135)."""
    for li in range(2):
        out = conv_layer(sd, f"{p}.layers.{li}", x, 1)
        x = torch.cat((x, out), dim=1)
    return conv_layer(sd, f"{p}.layers.2", x, 1)


def decoder_block(sd, p, x, skip=None):
    """Decoder.forward (utils.py:249-257): bicubic x2 -> cat(skip) -> ShortResBlock."""
    x = F.interpolate(x, scale_factor=2, mode="bicubic")
    if skip is not None:
        x = torch.cat((x, skip), dim=1)
    return short_res_block(sd, p + ".conv", x)


def depth_activation(sd, p, x):
    """Depth_Activation.forward (utils.py:285-289)."""
    x = F.conv2d(x, sd[p + ".conv_1.weight"], sd[p + ".conv_1.bias"], padding=1)
    x = torch.sigmoid(x)
    return F.conv2d(x, sd[p + ".conv_2.weight"], sd[p + ".conv_2.bias"], padding=1)


def seg_block(logits, num_classes):
    """Seg_Block.forward (utils.py:95-100)."""
    return torch.argmax(logits, dim=1, keepdim=True) / num_classes


def decoder(sd, cfg: Cfg, lay_out, x, dropout_scales: Optional[Sequence] = None):
    """CamRaDepth.dest_decoder (CamRaDepth.py:99-170).

    dropout_scales: list of (B, C) Dropout2d scale tensors (0 or 1/0.8) in application order,
    or None for eval mode.
    """
    it = iter(dropout_scales) if dropout_scales is not None else None

    def drop(t):
        if it is None:
            return t
        m = next(it)
        return t * m.view(m.shape[0], m.shape[1], 1, 1).to(t.dtype)

    unsup_map = sup_seg_map = seg_logits_final = seg_map = seg_features = None
    e1 = conv_layer(sd, "from_encoder_1", lay_out[-1], 0)
    e2 = conv_layer(sd, "from_encoder_2", lay_out[-2], 0)
    e3 = conv_layer(sd, "from_encoder_3", lay_out[-3], 0)
    e4 = conv_layer(sd, "from_encoder_4", lay_out[-4], 0)
    d1 = drop(decoder_block(sd, "depth_upsample.0", e1, e2))
    d2 = drop(decoder_block(sd, "depth_upsample.1", d1, e3))
    d3 = drop(decoder_block(sd, "depth_upsample.2", d2, e4))
    inter3 = depth_activation(sd, "depth_activation_3", d3)
    d3 = torch.cat([d3, inter3], 1)
    d4 = drop(decoder_block(sd, "depth_upsample.3", d3))
    if cfg.sup or cfg.unsup:
        seg_features = drop(decoder_block(sd, "seg_upsample.0", d3))
    if cfg.sup:
        lg = F.conv2d(seg_features, sd["seg_conv_stage_4.weight"], sd["seg_conv_stage_4.bias"], padding=1)
        sup_seg_map = seg_block(lg, cfg.num_classes)
        seg_map = sup_seg_map
    if cfg.unsup:
        um = F.conv2d(seg_features, sd["unsup_stage_4.weight"], sd["unsup_stage_4.bias"], padding=1)
        unsup_map = seg_block(um, 19)
        seg_map = unsup_map if sup_seg_map is None else torch.cat([sup_seg_map, unsup_map], 1)
    if cfg.sup:
        seg_features = torch.cat((seg_features, sup_seg_map.to(seg_features.dtype)), dim=1)
    elif cfg.unsup:
        seg_features = torch.cat((seg_features, unsup_map.to(seg_features.dtype)), dim=1)
    tmp = torch.cat((d4, seg_map.to(d4.dtype)), dim=1) if seg_map is not None else d4
    inter4 = depth_activation(sd, "depth_activation_4", tmp)
    d4 = torch.cat([d4, inter4], 1)
    d5 = drop(decoder_block(sd, "depth_upsample.4", d4, x))
    if cfg.sup or cfg.unsup:
        seg_features = drop(decoder_block(sd, "seg_upsample.1", seg_features, x))
    if cfg.sup:
        seg_logits_final = F.conv2d(seg_features, sd["seg_conv_final.weight"], sd["seg_conv_final.bias"], padding=1)
        sup_seg_map = seg_block(seg_logits_final, cfg.num_classes)
        seg_map = sup_seg_map
    if cfg.unsup:
        um = F.conv2d(seg_features, sd["unsup_final.weight"], sd["unsup_final.bias"], padding=1)
        unsup_map = seg_block(um, 19)
        seg_map = unsup_map if sup_seg_map is None else torch.cat([sup_seg_map, unsup_map], 1)
    tmp = torch.cat((d5, seg_map.to(d5.dtype)), dim=1) if seg_map is not None else d5
    final_depth = depth_activation(sd, "depth_activation_5", tmp)
    return {"depth": {"intermediate_depths": (None, None, inter3, inter4), "final_depth": final_depth},
            "seg": {"final_seg": seg_logits_final, "intermediate_seg": None, "unsup_map": unsup_map}}


def forward(sd, cfg: Cfg, x, drop_path_scales=None, dropout_scales=None):
    """CamRaDepth.forward (CamRaDepth.py:173-176)."""
    outs = encoder(sd, cfg, x, drop_path_scales)
    return decoder(sd, cfg, outs, x, dropout_scales)


# ----------------------------------------------------------------------------- losses

def masked_smooth_l1(pred, target):
    """MaskedSmoothL1Loss.forward (loss_funcs.py:83-91), SmoothL1 beta=1, mean over target>0."""
    assert pred.dim() == target.dim()
    m = target > 0
    d = (pred[m] - target[m])
    a = d.abs()
    return torch.where(a < 1, 0.5 * d * d, a - 0.5).mean()


def masked_mse(pred, target):
    """MaskedMSELoss.forward (loss_funcs.py:40-46)."""
    m = target > 0
    return ((target - pred)[m] ** 2).mean()


def masked_focal(logits, target, gamma=2):
    """MaskedFocalLoss.forward (loss_funcs.py:25-34): focal transform of the SCALAR mean CE."""
    ce = F.cross_entropy(logits, target, ignore_index=255)
    pt = torch.exp(-ce)
    return (1 - pt) ** gamma * ce


def training_loss(pred, gt_final, gt_s4, gt_s3, gt_seg, cfg: Cfg, update_interval=1):
    """Loss mix of Trainer.train_one_epoch (runner.py:197-218)."""
    l_seg = 0.0
    if cfg.sup and pred["seg"]["final_seg"] is not None:
        l_seg = masked_focal(pred["seg"]["final_seg"], gt_seg)
    inter = pred["depth"]["intermediate_depths"]
    l4 = masked_smooth_l1(inter[-1].squeeze(1), gt_s4.squeeze(1))
    l3 = masked_smooth_l1(inter[-2].squeeze(1), gt_s3.squeeze(1))
    lf = masked_smooth_l1(pred["depth"]["final_depth"], gt_final)
    w = [1, 1, 1, 0.2, 0.2]
    loss = (w[0] * lf + w[1] * l4 + w[2] * l3 + w[3] * l_seg + w[4] * 0.0) / sum(w)
    return loss / update_interval, {"final": lf, "s4": l4, "s3": l3, "seg": l_seg}


def minpool(t):
    """NuscenesDataset.__getitem__.minpool (dataloader.py:213-222): zero-ignoring 3x3 s2 min-pool."""
    x = t.clone()
    x[t == 0] = 255
    x = -F.max_pool2d(-x, kernel_size=3, stride=2, padding=1)
    x[x == 255] = 0
    return x


# ----------------------------------------------------------------------------- optimizer

def diffgradnorm_step(p, g, state, lr=6e-5, betas=(0.9, 0.999), eps=1e-8):
    """One diffGradNorm.step for one tensor (diffGradNorm.py:41-113). Mutates p and state.

    state: dict(step, exp_avg, exp_avg_sq, previous_grad, exp_grad_norm)
    """
    if not state:
        state.update(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p),
                     previous_grad=torch.zeros_like(p), exp_grad_norm=torch.zeros((), dtype=p.dtype))
    b1, b2 = betas
    state["step"] += 1
    gn = torch.linalg.norm(g)
    egn = 0.95 * state["exp_grad_norm"] + 0.05 * gn
    g1 = g * egn / (gn + 1e-8) if bool(egn > gn) else g
    state["exp_grad_norm"] = egn.clone()
    state["exp_avg"].mul_(b1).add_(g1, alpha=1 - b1)
    state["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = state["exp_avg_sq"].sqrt().add_(eps)
    bc1 = 1 - b1 ** state["step"]
    bc2 = 1 - b2 ** state["step"]
    dfc = 1.0 / (1.0 + torch.exp(-(state["previous_grad"] - g).abs()))
    state["previous_grad"] = g.clone()
    step_size = lr * math.sqrt(bc2) / (bc1 + 1e-8)
    p.addcdiv_(state["exp_avg"] * dfc, denom, value=-step_size)


# ----------------------------------------------------------------------------- stochastic masks

def n_dropout_sites(cfg: Cfg) -> int:
    return 7 if (cfg.sup or cfg.unsup) else 5


def make_masks(cfg: Cfg, B: int, seed: int):
    """Seeded DropPath scales (2 x 34 x (B,): one per `drop_path` CALL, attention branch then Mix-FFN branch of
    each block, simplified_attention.py:143-144) and Dropout2d scales (5 or 7 x (B,128)).

    DropPath: timm drop_path, rate linspace(0, .1, sum(depths)) (simplified_attention.py:214);
    block 0 has rate 0 -> Identity (:123).  Dropout2d(0.2): whole (b, channel) planes (CamRaDepth.py:96).
    """
    g = torch.Generator().manual_seed(seed)
    nb = sum(cfg.depths)
    rates = torch.linspace(0, DROP_PATH_RATE, nb).tolist()
    dps = []
    for r in rates:
        keep = 1.0 - r
        for _ in range(2):
            dps.append((torch.rand(B, generator=g) < keep).float() / keep)
    d2 = [(torch.rand(B, MID, generator=g) >= DROPOUT2D_P).float() / (1 - DROPOUT2D_P)
          for _ in range(n_dropout_sites(cfg))]
    return dps, d2
