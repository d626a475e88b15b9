"""Drop-in `CamRaDepth` nn.Module: the reference's constructor, `forward()` signature, output
dictionary and `state_dict` layout (src/models/CamRaDepth.py:20-176), executed by the B200 engine.
"""
from __future__ import annotations

import os
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from .args import args
from .engine import Engine
from .spec import ModelCfg, param_spec, init_tensor


def cast_tuple(val, depth):
    return val if isinstance(val, tuple) else (val,) * depth


class _Node(nn.Module):
    """Parameter container reproducing the reference's module tree (names only; no forward)."""


class _Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        eng = model._engine_for(x)
        masks = model._pop_masks()
        outs, saved = eng.forward(x, model.training, masks, save=True, packed=model._packed_call)
        ctx.model, ctx.eng, ctx.saved = model, eng, saved
        ret = [outs["final_depth"], outs["inter3"], outs["inter4"]]
        ctx.has_seg = outs["final_seg"] is not None
        ctx.has_unsup = outs["unsup_map"] is not None
        if ctx.has_seg:
            ret.append(outs["final_seg"])
        if ctx.has_unsup:
            ret.append(outs["unsup_map"])
            ctx.mark_non_differentiable(outs["unsup_map"])
        return tuple(ret)

    @staticmethod
    def backward(ctx, g_final, g3, g4, *rest):
        g_seg = rest[0] if ctx.has_seg else None
        model, eng = ctx.model, ctx.eng
        if ctx.saved is None:
            raise RuntimeError("backward through CamRaDepth called twice (activations were released)")
        hook = model._grad_bucket_hook
        grads = eng.backward(ctx.saved, g_final, g3, g4, g_seg,
                             on_bucket=None if hook is None else partial(hook, eng))
        ctx.saved = None
        if model._post_backward_hook is not None:
            model._post_backward_hook(eng, grads)
        return (None, None) + tuple(grads.get(n) for n in eng.names)


class CamRaDepth(nn.Module):
    def __init__(
            self,
            img_size=(416, 800),
            heads=(1, 2, 4, 8),
            ff_expansion=(8, 8, 4, 4),
            reduction_ratio=(8, 4, 2, 1),
            depths=(3, 10, 16, 5),
            dims=(64, 128, 160, 256),
            input_channels=None,
            **kwargs
    ):
        super().__init__()
        self.depths = depths
        self.mid_channels = 128
        self.num_classes = args.num_classes
        self.dense = True
        self.dims = dims
        self.as_final_block = False
        self.unsupervised_seg = args.get("unsupervised_seg", False)
        self.supervised_seg = args.get("supervised_seg", False)
        self.img_size = np.array(img_size)
        input_channels = input_channels if input_channels is not None else args.input_channels

        dims, heads, ff_expansion, reduction_ratio, self.depths = map(
            partial(cast_tuple, depth=4), (dims, heads, ff_expansion, reduction_ratio, self.depths))
        assert all([*map(lambda t: len(t) == 4, (dims, heads, ff_expansion, reduction_ratio, self.depths))]), \
            'only four stages are allowed, all keyword arguments must be either a single value or a tuple of 4 values'
        assert input_channels > 0, 'input_channels must be > 0'

        self.cfg = ModelCfg(dims, heads, ff_expansion, reduction_ratio, self.depths, input_channels,
                            self.supervised_seg, self.unsupervised_seg, self.num_classes,
                            args.get("groupnorm_divisor", 16))
        # "bf16" (training path, tensor cores) or "fp32" (exact-parity mode on CUDA cores)
        self.precision = kwargs.get("precision", os.environ.get("CAMRADEPTH_PRECISION", "bf16"))
        assert self.precision in ("bf16", "fp32")
        # bit-reproducible forward (fixed-order GroupNorm reductions instead of fp32 atomics); slower on the large
        # decoder tensors, meant for debugging / regression runs
        self.deterministic = bool(kwargs.get("deterministic", os.environ.get("CAMRADEPTH_DETERMINISTIC", "0") == "1"))

        spec = param_spec(self.cfg)
        for name, (shape, kind) in spec.items():
            wshape = spec[name.replace(".bias", ".weight")][0] if kind == "default_bias" else None
            self._register(name, nn.Parameter(init_tensor(shape, kind, wshape)))
        self.dropout = nn.Dropout2d(0.2)

        self._engines = {}
        self._masks = None
        self._grad_bucket_hook = None
        self._post_backward_hook = None
        self._packed_call = False

    def _register(self, name, param):
        node = self
        parts = name.split(".")
        for part in parts[:-1]:
            child = node._modules.get(part)
            if child is None:
                child = _Node()
                node.add_module(part, child)
            node = child
        node.register_parameter(parts[-1], param)

    # -- engine plumbing -------------------------------------------------------------------------
    def _engine_for(self, x):
        key = self.precision
        eng = self._engines.get(key)
        if eng is None:
            eng = Engine(self, self.cfg, precision=key, deterministic=self.deterministic)
            self._engines[key] = eng
        else:
            eng.P = dict(self.named_parameters())
        for p in eng.P.values():
            if p.device != x.device:
                raise RuntimeError(f"input is on {x.device} but parameters are on {p.device}")
            break
        return eng

    def set_stochastic_masks(self, drop_path_scales, dropout2d_scales):
        """Inject the DropPath / Dropout2d scale tensors used by the NEXT training-mode forward
        (parity tests share masks with the reference).  drop_path_scales: 2 * sum(depths) entries
        ((B,) tensors or None), one per `drop_path` call in the reference's call order -- attention branch then
        Mix-FFN branch of every block (simplified_attention.py:143-144)."""
        self._masks = (list(drop_path_scales), list(dropout2d_scales))

    def _pop_masks(self):
        m, self._masks = self._masks, None
        return m

    def forward_packed(self, x_nhwc):
        """Same as forward(x) for an input that is already in the engine's layout: (B,H,W,8) bf16 NHWC
        [RGB | radar planes | zero pad] as produced by `preprocess.pack_input_nhwc` (the
        GPU input pipeline feeds the network directly, no NCHW fp32 -> NHWC bf16 pack inside the step)."""
        self._packed_call = True
        try:
            return self.forward(x_nhwc)
        finally:
            self._packed_call = False

    # -- reference surface -----------------------------------------------------------------------
    def forward(self, x):
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if need_grad:
            res = _Fn.apply(self, x, *self.parameters())
            final_depth, inter3, inter4 = res[0], res[1], res[2]
            i = 3
            final_seg = unsup_map = None
            if self.supervised_seg:
                final_seg = res[i]
                i += 1
            if self.unsupervised_seg:
                unsup_map = res[i]
        else:
            eng = self._engine_for(x)
            outs, _ = eng.forward(x, self.training, self._pop_masks(), save=False, packed=self._packed_call)
            final_depth, inter3, inter4 = outs["final_depth"], outs["inter3"], outs["inter4"]
            final_seg, unsup_map = outs["final_seg"], outs["unsup_map"]
        return {"depth": {"intermediate_depths": (None, None, inter3, inter4), "final_depth": final_depth},
                "seg": {"final_seg": final_seg, "intermediate_seg": None, "unsup_map": unsup_map}}


def load_checkpoint_with_shape_match(model, checkpoint_dict):
    """Same contract as the reference loader (utils.py:352-370): `module.` prefixes of DataParallel checkpoints are
    ignored, a tensor is taken from the checkpoint only when both its name and its shape match, everything else keeps
    the model's current (freshly initialised) value and is reported on stdout."""
    have = {name.replace('module.', ''): t for name, t in checkpoint_dict.items()}
    current = model.state_dict()
    merged = dict(current)
    for name, cur in current.items():
        src = have.get(name)
        if src is None:
            print(f"{args.hashtags_prefix} Key not in checkpoint: ", name)
        elif src.shape != cur.shape:
            print(f"{args.hashtags_prefix} Shape mismatch: ", name, src.shape, cur.shape)
        else:
            merged[name] = src
    model.load_state_dict(merged, strict=True)
