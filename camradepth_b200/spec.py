"""Parameter schema of the drop-in CamRaDepth module: names, shapes and init distributions.

Mirrors the reference's registration order so `state_dict()` / `parameters()` line up with reference
checkpoints and saved optimizer state (simplified_attention.py:190-246,
CamRaDepth.py:53-94, utils.py:103-124,201-221,274-283).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

MID = 128            # CamRaDepth.py:37
DROP_PATH_RATE = 0.1  # CamRaDepth.py:57
DROPOUT2D_P = 0.2     # CamRaDepth.py:96


class ModelCfg:
    def __init__(self, dims, heads, ff, sr, depths, cin, sup, unsup, num_classes, gn_div):
        self.dims, self.heads, self.ff, self.sr, self.depths = tuple(dims), tuple(heads), tuple(ff), tuple(sr), tuple(depths)
        self.cin, self.sup, self.unsup = int(cin), bool(sup), bool(unsup)
        self.num_classes, self.gn_div = int(num_classes), int(gn_div)

    @property
    def nseg(self):
        return int(self.sup) + int(self.unsup)

    @property
    def n_dropout_sites(self):
        return 7 if (self.sup or self.unsup) else 5


def param_spec(cfg: ModelCfg):
    """OrderedDict name -> (shape, init kind)."""
    sp = OrderedDict()
    dims, depths, ff, sr, cin = cfg.dims, cfg.depths, cfg.ff, cfg.sr, cfg.cin
    pe_in = (cin, dims[0], dims[1], dims[2])
    pe_k = (7, 3, 3, 3)
    for s in range(4):
        p = f"dest_encoder.patch_embed{s + 1}"
        sp[p + ".proj.weight"] = ((dims[s], pe_in[s], pe_k[s], pe_k[s]), "fanout")
        sp[p + ".proj.bias"] = ((dims[s],), "zeros")
        sp[p + ".norm.weight"] = ((dims[s],), "ones")
        sp[p + ".norm.bias"] = ((dims[s],), "zeros")
    for s in range(4):
        C, rC = dims[s], int(dims[s] * ff[s])
        for i in range(depths[s]):
            p = f"dest_encoder.block{s + 1}.{i}"
            for n in ("norm1", "norm2"):
                sp[f"{p}.{n}.weight"] = ((C,), "ones")
                sp[f"{p}.{n}.bias"] = ((C,), "zeros")
            for n in ("q", "k", "proj"):
                sp[f"{p}.attn.{n}.weight"] = ((C, C, 1), "tn02")
                sp[f"{p}.attn.{n}.bias"] = ((C,), "zeros")
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
            for n in ("norm1", "norm2"):
                sp[f"{p}.mlp1.{n}.weight"] = ((rC,), "ones")
                sp[f"{p}.mlp1.{n}.bias"] = ((rC,), "zeros")
    for j, C in enumerate((dims[3], dims[2], dims[1], dims[0])):
        p = f"from_encoder_{j + 1}.model"
        sp[p + ".0.weight"] = ((C, C, 1, 1), "kaiming")
        sp[p + ".1.weight"] = ((C,), "ones")
        sp[p + ".1.bias"] = ((C,), "zeros")

    def short_res(prefix, cin_):
        inp = cin_
        for li, out in enumerate((int(MID * 0.75), int(MID * 0.5), MID)):
            sp[f"{prefix}.conv.layers.{li}.model.0.weight"] = ((out, inp, 3, 3), "kaiming")
            sp[f"{prefix}.conv.layers.{li}.model.1.weight"] = ((out,), "ones")
            sp[f"{prefix}.conv.layers.{li}.model.1.bias"] = ((out,), "zeros")
            inp += out

    dec_in = (dims[3] + dims[2], MID + dims[1], MID + dims[0], MID + 1, MID + 1 + cin)
    for d in range(5):
        short_res(f"depth_upsample.{d}", dec_in[d])
    for name, c in (("depth_activation_3", MID), ("depth_activation_4", MID + cfg.nseg),
                    ("depth_activation_5", MID + cfg.nseg)):
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


def init_tensor(shape, kind, wshape=None):
    """Fresh-construction init (SURVEY.md §8b "Init"); draws from torch's global RNG like the reference."""
    if kind == "zeros":
        return torch.zeros(shape)
    if kind == "ones":
        return torch.ones(shape)
    if kind == "tn02":
        return torch.nn.init.trunc_normal_(torch.empty(shape), std=.02)
    if kind in ("fanout", "fanout_dw", "kaiming"):
        fan_out = shape[2] * shape[3] * shape[0]
        if kind == "fanout_dw":
            fan_out //= shape[0]
        return torch.empty(shape).normal_(0, math.sqrt(2.0 / fan_out))
    if kind == "default":
        bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
        return torch.empty(shape).uniform_(-bound, bound)
    if kind == "default_bias":
        bound = 1.0 / math.sqrt(wshape[1] * wshape[2] * wshape[3])
        return torch.empty(shape).uniform_(-bound, bound)
    raise ValueError(kind)
