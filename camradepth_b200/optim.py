"""`diffGradNorm` with the reference's constructor signature (src/models/diffGradNorm.py:26) as a
multi-tensor CUDA step: one sum-of-squares launch + one update launch for ALL parameter tensors, the
norm-correction branch (diffGradNorm.py:84) evaluated on the device (no per-tensor host sync)."""
from __future__ import annotations

import math

import numpy as np
import torch
from torch.optim.optimizer import Optimizer

from . import ops
from .engine import bump_weight_epoch


class diffGradNorm(Optimizer):
    def __init__(self, args, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super(diffGradNorm, self).__init__(args, defaults)
        self._tables = {}

    @staticmethod
    def build_tables(ptrs, numels, device):
        """Host-side table construction (pure integer logic; unit-tested on CPU).
        ptrs: list of (p, g, m, v, prev) device addresses; -> (table int64[T,6], chunks int64[K,2])"""
        table = np.zeros((len(ptrs), 6), dtype=np.int64)
        chunks = []
        for t, (pp, n) in enumerate(zip(ptrs, numels)):
            table[t, :5] = pp
            table[t, 5] = n
            for start in range(0, n, ops.OPT_CHUNK):
                # crd_opt_chunk {int tensor; int pad; long long start}
                chunks.append((t, start))
        ck = np.zeros((len(chunks), 2), dtype=np.int64)
        for i, (t, start) in enumerate(chunks):
            ck[i, 0] = t            # low 32 bits = tensor, high 32 bits = pad (little endian)
            ck[i, 1] = start
        return table, ck

    @staticmethod
    def _step_size(group, step_no):
        beta1, beta2 = group['betas']
        bc1 = 1 - beta1 ** step_no
        bc2 = 1 - beta2 ** step_no
        return group['lr'] * math.sqrt(bc2) / (bc1 + 1e-8)       # diffGradNorm.py:97-108

    @torch.no_grad()
    def advance_for_replay(self):
        """Host-side part of a step whose device work is replayed from a CUDA graph: bump the step counters
        and refresh the device-resident step size (lr schedule + bias corrections)."""
        for gi, group in enumerate(self.param_groups):
            ent = self._tables.get(gi)
            if ent is None:
                continue
            step_no = None
            for p in group['params']:
                st = self.state.get(p)
                if st:
                    st['step'] += 1
                    step_no = st['step']
            if step_no is not None:
                ent[7].fill_(self._step_size(group, step_no))
        bump_weight_epoch()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if group['weight_decay'] != 0:
                # the reference's weight-decay branch calls a removed add_(scalar, tensor) overload and raises
                raise RuntimeError("diffGradNorm: weight_decay != 0 is not supported (it raises in the reference too)")
            params = [p for p in group['params'] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            if not params[0].is_cuda:
                raise RuntimeError("camradepth_b200.diffGradNorm runs on CUDA devices only (no CPU fallback)")
            by_step = {}
            capturing = torch.cuda.is_current_stream_capturing()
            for p in params:
                if p.grad.is_sparse:
                    raise RuntimeError('diffGradNorm does not support sparse gradients')
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p.data)
                    st['exp_avg_sq'] = torch.zeros_like(p.data)
                    st['previous_grad'] = torch.zeros_like(p.data)
                    st['exp_grad_norm'] = torch.zeros((), dtype=torch.float32, device=dev)
                    st['_slot'] = None
                if not capturing:          # a capture records the device work only; advance_for_replay() counts steps
                    st['step'] += 1
                by_step.setdefault(st['step'], []).append(p)
            beta1, beta2 = group['betas']
            for step_no, plist in by_step.items():
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in plist]
                key = (gi, tuple((p.data_ptr(), g.data_ptr(), self.state[p]['previous_grad'].data_ptr())
                                 for p, g in zip(plist, grads)))
                ent = self._tables.get(gi)
                if ent is None or ent[0] != key:
                    ptrs = [(p.data_ptr(), g.data_ptr(), self.state[p]['exp_avg'].data_ptr(),
                             self.state[p]['exp_avg_sq'].data_ptr(), self.state[p]['previous_grad'].data_ptr())
                            for p, g in zip(plist, grads)]
                    table, ck = self.build_tables(ptrs, [p.numel() for p in plist], dev)
                    T = len(plist)
                    # per-tensor exp_grad_norm lives in one persistent ping-pong buffer; state entries are views
                    egn = torch.stack([self.state[p]['exp_grad_norm'].reshape(()) for p in plist]).contiguous()
                    pp = torch.stack([egn, torch.zeros_like(egn)]).contiguous()
                    ent = [key, torch.from_numpy(table).to(dev), torch.from_numpy(ck).to(dev), ck.shape[0], pp, 0,
                           torch.zeros(ck.shape[0], dtype=torch.float32, device=dev),     # per-chunk partials
                           torch.zeros(1, dtype=torch.float32, device=dev)]
                    self._tables[gi] = ent
                _, table_d, ck_d, nchunks, pp, cur, sumsq, hyper = ent
                egn_in, egn_out = pp[cur], pp[1 - cur]
                ops.mt_sumsq(table_d, ck_d, nchunks, sumsq)
                if not torch.cuda.is_current_stream_capturing():
                    # the step size lives in a device scalar so a captured step can be replayed with fresh
                    # bias corrections (see advance_for_replay)
                    hyper.fill_(self._step_size(group, step_no))
                ops.diffgradnorm_update(table_d, ck_d, nchunks, sumsq, egn_in, egn_out, hyper,
                                        float(beta1), float(beta2), float(group['eps']))
                if torch.cuda.is_current_stream_capturing():
                    # a replayed graph must read and write the same exp_grad_norm buffer every time
                    pp[cur].copy_(pp[1 - cur])
                    egn_out = pp[cur]
                else:
                    ent[5] = 1 - cur
                for i, p in enumerate(plist):
                    self.state[p]['exp_grad_norm'] = egn_out[i]
        bump_weight_epoch()
        return loss
