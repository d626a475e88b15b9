"""Forward / backward programs of the CamRaDepth hot path on the C-ABI kernels.

The engine is a hand-scheduled tape: `forward` launches the kernels of the encoder
(simplified_attention.py:265-306) and decoder (CamRaDepth.py:99-170) in stream order and records
what backward needs; `backward` replays the reverse program and fills one flat fp32 gradient
buffer (views per parameter, reference parameter order) so data-parallel all-reduce works on
contiguous buckets.  torch is used only for device memory, the RNG of the stochastic masks and
stream handles.

Data layout in HBM (bf16 mode; fp32 mode stores everything as fp32):
  activations NHWC bf16, channel counts padded to multiples of 8 with zero channels;
  the encoder residual stream, GroupNorm statistics and parameter gradients are fp32;
  each decoder ShortResBlock owns ONE concat buffer [up(src) | skip | o1(96) | o2(64)] that the
  producers write slice-wise (no torch.cat copies; utils.py:127-135,249-257).
"""
from __future__ import annotations

import math
import os

import torch

from . import ops
from .spec import MID, DROP_PATH_RATE, DROPOUT2D_P

# bumped by optimizers that update parameters through raw pointers (no torch version bump)
WEIGHT_EPOCH = [0]


def bump_weight_epoch():
    WEIGHT_EPOCH[0] += 1


def r8(c):
    return (c + 7) // 8 * 8


def r64(c):
    """Pixel stride (in channels) of the wide decoder buffers: a multiple of 64 bf16 = 128 bytes, so every
    64-channel TMA row of a tap/chunk starts on a 128-byte line (an unaligned row costs two L2 requests)."""
    return (c + 63) // 64 * 64 if os.environ.get("CAMRADEPTH_PAD128", "1") == "1" else c


class ZeroArena:
    """Bump allocator over one zero-filled fp32 buffer (one memset per pass instead of hundreds)."""

    def __init__(self, device):
        self.device = device
        self.buf = None
        self.off = 0
        self.need = 0

    def reset(self):
        cap = 0 if self.buf is None else self.buf.numel()
        if self.need > cap:
            self.buf = torch.empty(int(self.need * 1.1) + 1024, dtype=torch.float32, device=self.device)
        if self.buf is not None:
            self.buf.zero_()
        self.off = 0
        self.need = 0

    def take(self, *shape):
        n = 1
        for s in shape:
            n *= s
        n_al = (n + 63) // 64 * 64
        self.need += n_al
        if self.buf is not None and self.off + n_al <= self.buf.numel():
            t = self.buf[self.off:self.off + n].view(*shape)
            self.off += n_al
            return t
        return torch.zeros(*shape, dtype=torch.float32, device=self.device)


class Engine:
    def __init__(self, model, cfg, precision="bf16", deterministic=False):
        self.model = model
        self.cfg = cfg
        self.precision = precision
        self.tdtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.P = dict(model.named_parameters())
        self.names = list(self.P.keys())
        self.device = None
        self._packs = {}
        self._pack_jobs = {}        # key -> ([(param name, dst view, cmap, Cout, Cin, taps, Cin_p, Cout_p, mode)], ver fn)
        self._pack_table = None     # (signature, device table, n items, n blocks)
        self._maps = {}
        self.use_tc = precision == "bf16" and os.environ.get("CAMRADEPTH_TC", "1") == "1"
        self.use_tc_wgrad = self.use_tc and os.environ.get("CAMRADEPTH_TC_WGRAD", "1") == "1"
        # GroupNorm statistics from the accumulator read-out of the producing tcgen05 conv / GEMM
        self.fuse_gn_stats = self.use_tc and os.environ.get("CAMRADEPTH_GN_EPILOGUE", "1") == "1"
        # GroupNorm passes whose tensors total at most this many bytes run as ONE launch (two sweeps of an
        # L2-resident slab per CTA) instead of statistics / finalize / apply launches; 0 disables
        self.gn_fused_bytes = int(float(os.environ.get("CAMRADEPTH_GN_FUSED_MB", "16")) * (1 << 20))
        # 1x1 GEMMs with wide outputs (Mix-FFN fc1: 512 .. 1024 channels) are bound by their accumulator read-out;
        # the GroupNorm reduction in that read-out costs more (+33 us at stage 2) than a separate statistics pass
        # over the bf16 output (~20 us), so it is fused only up to this many output channels
        self.gn_epilogue_max_cout = int(os.environ.get("CAMRADEPTH_GN_EPILOGUE_MAXN", "256"))
        self.strided_igemm = os.environ.get("CAMRADEPTH_STRIDED_IGEMM", "1") == "1"
        # Deterministic forward: every GroupNorm statistic comes from the one-launch kernel, whose reduction order is
        # fixed (per-thread strides, shared-memory tree, no atomics), at any tensor size; the conv read-out sums
        # (fp32 atomics across CTAs) are not used.  The forward pass is then bit-reproducible run to run and
        # independent of the batch a sample sits in.  (Parameter gradients still meet in fp32 atomics.)
        self.deterministic = bool(deterministic) or os.environ.get("CAMRADEPTH_DETERMINISTIC", "0") == "1"
        if self.deterministic:
            self.fuse_gn_stats = False
            self.gn_fused_bytes = 1 << 62
        self._build_layers()
        self.fwd_arena = None
        self.bwd_arena = None
        # optional CUDA-event timing of selected conv launches: {(kind, weight name): [(ev0, ev1), ...]}
        self.timed = None
        self.timed_flops = {}        # (kind, weight name) -> MACs*2 of the launch actually timed, where it differs from the layer's own
        self._grad_epoch = 0
        self._flat_own = None
        self._grads_out = None
        self._mask_state = None

    def _timed(self, kind, name):
        if self.timed is None or (kind, name) not in self.timed:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        self.timed[(kind, name)].append((e0, e1))
        e0.record()
        return e1

    # ------------------------------------------------------------------ layer table
    def _build_layers(self):
        cfg = self.cfg
        L = {}

        def add(name, k, stride, pad, cmap=None, cin_p=None):
            shp = self.P[name].shape
            cout, cin = shp[0], shp[1]
            cin_p = r8(cin) if cin_p is None else cin_p
            L[name] = dict(k=k, stride=stride, pad=pad, cmap=cmap, cin=cin, cin_p=cin_p, cout=cout,
                           cout_p=r8(cout), taps=k * k)

        pe_k, pe_s = (7, 3, 3, 3), (4, 2, 2, 2)
        for s in range(4):
            add(f"dest_encoder.patch_embed{s + 1}.proj.weight", pe_k[s], pe_s[s], pe_k[s] // 2)
            for i in range(cfg.depths[s]):
                p = f"dest_encoder.block{s + 1}.{i}"
                add(p + ".attn.q.weight", 1, 1, 0)
                add(p + ".attn.k.weight", 1, 1, 0)
                if cfg.sr[s] > 1:
                    add(p + ".attn.sr.weight", cfg.sr[s], cfg.sr[s], 0)
                add(p + ".mlp1.fc1.weight", 1, 1, 0)
                add(p + ".mlp1.fc2.weight", 1, 1, 0)
        for j in range(4):
            add(f"from_encoder_{j + 1}.model.0.weight", 1, 1, 0)

        def short_res(prefix, cx):
            cxp = r8(cx)
            for li, extra in enumerate((0, 96, 160)):
                cin = cx + extra
                cmap = None if cx == cxp else [c if c < cx else cxp + (c - cx) for c in range(cin)]
                add(f"{prefix}.conv.layers.{li}.model.0.weight", 3, 1, 1, cmap, cxp + extra)

        d = cfg.dims
        self.dec_cx = {"depth_upsample.0": d[3] + d[2], "depth_upsample.1": MID + d[1],
                       "depth_upsample.2": MID + d[0], "depth_upsample.3": MID + 1,
                       "depth_upsample.4": MID + 1 + cfg.cin}
        if cfg.sup or cfg.unsup:
            self.dec_cx["seg_upsample.0"] = MID + 1
            self.dec_cx["seg_upsample.1"] = MID + 1 + cfg.cin
        for k, cx in self.dec_cx.items():
            short_res(k, cx)
        add("depth_activation_3.conv_1.weight", 3, 1, 1)
        for n in ("depth_activation_4", "depth_activation_5"):
            if cfg.nseg:
                cmap = list(range(MID)) + [MID + 1 + j for j in range(cfg.nseg)]
                add(n + ".conv_1.weight", 3, 1, 1, cmap, r8(MID + 1 + cfg.nseg))
            else:
                add(n + ".conv_1.weight", 3, 1, 1)
        for n in ("depth_activation_3", "depth_activation_4", "depth_activation_5"):
            add(n + ".conv_2.weight", 3, 1, 1)
        if cfg.sup:
            add("seg_conv_stage_4.weight", 3, 1, 1)
            add("seg_conv_final.weight", 3, 1, 1)
        if cfg.unsup:
            add("unsup_stage_4.weight", 3, 1, 1)
            add("unsup_final.weight", 3, 1, 1)
        self.L = L

    # ------------------------------------------------------------------ small helpers
    def _empty(self, *shape, dtype=None):
        return torch.empty(*shape, dtype=self.tdtype if dtype is None else dtype, device=self.device)

    def _feat(self, B, H, W, FW):
        """Feature buffer [feat 128 | depth | seg maps | pad]: the 128 feature channels are fully written by the
        decoder block, so only the tail channels are zero-filled (a full fill of the 192x416 buffer is 0.7 GB)."""
        t = self._empty(B, H, W, FW)
        ops.zero_channels(t[..., MID:])
        return t

    def _zeros(self, *shape, dtype=None):
        return torch.zeros(*shape, dtype=self.tdtype if dtype is None else dtype, device=self.device)

    def _cmap(self, name):
        cm = self.L[name]["cmap"]
        if cm is None:
            return None
        t = self._maps.get(name)
        if t is None or t.device != self.device:
            t = torch.tensor(cm, dtype=torch.int32, device=self.device)
            self._maps[name] = t
        return t

    def wpack(self, name, mode, dtype=None):
        """Packed K-major copy of a conv weight (mode 0: forward / wgrad layout, 1: dgrad layout)."""
        dtype = self.tdtype if dtype is None else dtype
        p = self.P[name]
        L = self.L[name]
        if mode == 0 and dtype == torch.float32 and L["taps"] == 1 and L["cmap"] is None and L["cin_p"] == L["cin"]:
            return p.detach()
        key = (name, mode, dtype)
        ver = (p._version, p.data_ptr(), WEIGHT_EPOCH[0], self._grad_epoch)
        ent = self._packs.get(key)
        if ent is not None and ent[1] == ver:
            return ent[0]
        if ent is not None and ent[0].device == self.device:
            dst = ent[0]
        elif mode == 0:
            dst = torch.zeros(L["cout"], L["taps"] * L["cin_p"], dtype=dtype, device=self.device)
        elif mode == 1:
            dst = torch.zeros(L["cin_p"], L["taps"] * L["cout_p"], dtype=dtype, device=self.device)
        else:
            dst = torch.zeros(L["taps"] * L["cin_p"], L["cout_p"], dtype=dtype, device=self.device)
        ops.weight_pack(p.detach(), dst, self._cmap(name), L["cout"], L["cin"], L["taps"], L["cin_p"], L["cout_p"],
                        mode)
        self._packs[key] = (dst, ver)
        self._pack_jobs[key] = ([(name, dst, self._cmap(name), L["cout"], L["cin"], L["taps"], L["cin_p"],
                                  L["cout_p"], mode)],
                                lambda: (p._version, p.data_ptr(), WEIGHT_EPOCH[0], self._grad_epoch))
        return dst

    def prepack(self):
        """Refresh every packed weight copy seen so far with ONE kernel launch (instead of ~380 per training
        step).  Called when the packed copies go stale: at the start of each grad-enabled forward."""
        if not self._pack_jobs:
            return
        jobs = [j for js, _ in self._pack_jobs.values() for j in js]
        sig = tuple((self.P[j[0]].data_ptr(), j[1].data_ptr()) for j in jobs)
        if (self._pack_table is None or self._pack_table[0] != sig) and torch.cuda.is_current_stream_capturing():
            # building the device table needs a host->device copy, which a capturing stream cannot take
            for (name, dst, cmap, cout, cin, taps, cin_p, cout_p, mode) in jobs:
                ops.weight_pack(self.P[name].detach(), dst, cmap, cout, cin, taps, cin_p, cout_p, mode)
        elif self._pack_table is None or self._pack_table[0] != sig:
            rows, blk, owner = [], 0, []
            for it, (name, dst, cmap, cout, cin, taps, cin_p, cout_p, mode) in enumerate(jobs):
                total = cout * cin * taps
                rows.append([self.P[name].data_ptr(), dst.data_ptr(), 0 if cmap is None else cmap.data_ptr(), blk,
                             cout, cin, taps, cin_p, cout_p, mode, ops.dcode(dst), total])
                nb = ops.weight_pack_blocks(cout, cin, taps)
                owner += [it] * nb
                blk += nb
            rows.append([0, 0, 0, blk] + [0] * 8)
            table = torch.tensor([v for r in rows for v in r] + owner, dtype=torch.int64).to(self.device)
            self._pack_table = (sig, table, len(jobs), blk)
        if self._pack_table is not None and self._pack_table[0] == sig:
            _, table, n, blk = self._pack_table
            ops.weight_pack_batch(table, n, blk)
        for key, (js, verfn) in self._pack_jobs.items():
            self._packs[key] = (self._packs[key][0], verfn())

    def _tc_ok(self, L, x, y_or_dy):
        """tcgen05 path: bf16 operands, stride-1 'same' KxK or 1x1 contractions."""
        return (self.tdtype == torch.bfloat16 and L["stride"] == 1 and L["k"] in (1, 3) and
                2 * L["pad"] == L["k"] - 1 and x.dtype == torch.bfloat16)

    def _gemm_route(self, L, x):
        """strided convs without an implicit-GEMM form on the tensor-core path (pe1: k7 s4 on 8 channels; Cin not a
        multiple of 64 for k3 s2) and every strided data gradient: im2col + GEMM (+ col2im)"""
        return self.use_tc and self.tdtype == torch.bfloat16 and L["stride"] > 1 and x.dtype == torch.bfloat16

    def _strided_tc(self, L, x):
        """Strided conv as implicit GEMM: the 5-D tensor maps of crd_conv_fwd_tc / crd_conv_wgrad_tc address the taps
        of a strided window directly (k == stride spatial-reduction convs; k3 s2 p1 patch embeddings)."""
        if not (self.use_tc and self.strided_igemm and self.tdtype == torch.bfloat16 and x.dtype == torch.bfloat16):
            return False
        k, st, pad = L["k"], L["stride"], L["pad"]
        if st <= 1 or x.shape[1] % st or x.shape[2] % st or ops._ld(x) != L["cin_p"]:
            return False
        return (k == st and pad == 0) or (k == 3 and st == 2 and pad == 1 and L["cin_p"] % 64 == 0)

    def _im2col(self, L, x, Ho, Wo):
        col = self._empty(x.shape[0], Ho, Wo, L["taps"] * L["cin_p"])
        ops.im2col(x, col, L["cin_p"], L["k"], L["k"], L["stride"], L["pad"])
        return col

    def conv(self, x, name, y, bias=None, act=0, accumulate=0, out_nchw=0, gn=False):
        """gn=True: the conv is followed by a GroupNorm.  On the tensor-core path the per-(sample, channel) sums are
        produced by the accumulator read-out of the conv itself and returned (zero-initialised arena slice);
        otherwise None is returned and the caller runs the separate statistics pass."""
        L = self.L[name]
        w = self.wpack(name, 0)
        fuse_gn = gn and self.fuse_gn_stats and act == 0 and not accumulate and \
            (L["k"] > 1 or L["cout"] <= self.gn_epilogue_max_cout)
        if self._gemm_route(L, x) and not out_nchw and not self._strided_tc(L, x):
            col = self._im2col(L, x, y.shape[1], y.shape[2])
            d = ops.make_desc(col, y, col.shape[-1], L["cout"], 1, 1, 1, 0, 0, act, accumulate, 0)
            sums = self.fwd_arena.take(x.shape[0], L["cout"], 2) if fuse_gn else None
            ops.conv_fwd(d, col, w, None if bias is None else self.P[bias].detach(), y, use_tc=True, gn_sums=sums)
            return sums
        d = ops.make_desc(x, y, L["cin_p"], L["cout"], L["k"], L["k"], L["stride"], L["pad"], 0, act, accumulate,
                          out_nchw)
        b = None if bias is None else self.P[bias].detach()
        ev = self._timed("fwd", name)
        tc = self.use_tc and (self._tc_ok(L, x, y) or self._strided_tc(L, x))
        sums = self.fwd_arena.take(x.shape[0], L["cout"], 2) if (fuse_gn and tc) else None
        ops.conv_fwd(d, x, w, b, y, use_tc=tc, gn_sums=sums)
        if ev is not None:
            ev.record()
        return sums

    def conv_dgrad(self, dy, name, dx, accumulate):
        L = self.L[name]
        if (self._gemm_route(L, dy) and self.strided_igemm and L["k"] == L["stride"] and L["pad"] == 0 and
                L["cin_p"] == L["cin"] and L["cin"] % 16 == 0 and ops._ld(dx) % 16 == 0 and ops._ld(dy) == L["cout_p"]
                and dx.shape[1] == dy.shape[1] * L["stride"] and dx.shape[2] == dy.shape[2] * L["stride"]):
            # k == stride: the windows do not overlap, so the data gradient is a 1x1 GEMM [pixels of dy] x [(tap, ci)]
            # whose read-out stores each 16-channel chunk at its (oh*s + kh, ow*s + kw) pixel -- no dcol buffer
            w = self.wpack(name, 2)
            d = ops.make_desc(dy, dx, L["cout_p"], L["cin_p"], L["k"], L["k"], L["stride"], 0, transposed=1,
                              accumulate=int(accumulate))
            ops.conv_fwd(d, dy, w, None, dx, use_tc=True)
            return
        if self._gemm_route(L, dy):
            w = self.wpack(name, 2)
            K = L["taps"] * L["cin_p"]
            dcol = self._empty(dy.shape[0], dy.shape[1], dy.shape[2], K)
            d = ops.make_desc(dy, dcol, L["cout_p"], K, 1, 1, 1, 0)
            ops.conv_fwd(d, dy, w, None, dcol, use_tc=True)
            ops.col2im(dcol, dx, L["cin_p"], L["k"], L["k"], L["stride"], L["pad"], accumulate)
            return
        w = self.wpack(name, 1)
        d = ops.make_desc(dy, dx, L["cout_p"], L["cin_p"], L["k"], L["k"], L["stride"], L["pad"], 1, 0,
                          int(accumulate), 0)
        ev = self._timed("dgrad", name)
        ops.conv_fwd(d, dy, w, None, dx, use_tc=self.use_tc and self._tc_ok(L, dy, dx))
        if ev is not None:
            ev.record()

    def conv_wgrad(self, x, dy, name, bias=None):
        L = self.L[name]
        direct = L["taps"] == 1 and L["cmap"] is None and L["cin_p"] == L["cin"]
        g = self.pg[name]
        dwp = g.view(L["cout"], L["cin"]) if direct else self.bwd_arena.take(L["cout"], L["taps"] * L["cin_p"])
        db = self.pg[bias] if bias is not None else None
        if self._gemm_route(L, x) and self.use_tc_wgrad and not self._strided_tc(L, x):
            col = self._im2col(L, x, dy.shape[1], dy.shape[2])      # recomputed: cheaper than keeping it alive
            d = ops.make_desc(col, dy, col.shape[-1], L["cout"], 1, 1, 1, 0)
            ops.conv_wgrad(d, col, dy, dwp, use_tc=True, db=db)     # bias gradient from the same pass over dy
            db = None
        else:
            d = ops.make_desc(x, dy, L["cin_p"], L["cout"], L["k"], L["k"], L["stride"], L["pad"], 0, 0, 0, 0)
            ev = self._timed("wgrad", name)
            tc = self.use_tc_wgrad and (self._tc_ok(L, x, dy) or self._strided_tc(L, x))
            fused = tc and L["k"] == 1 and db is not None
            ops.conv_wgrad(d, x, dy, dwp, use_tc=tc, db=db if fused else None)
            if fused:
                db = None
            if ev is not None:
                ev.record()
        if not direct:
            ops.weight_unpack_grad(dwp, g, self._cmap(name), L["cout"], L["cin"], L["taps"], L["cin_p"], False)
        if db is not None:
            ops.col_sum(dy, db, L["cout"])

    def _gn_fused_ok(self, x, G, other=None):
        """One-launch GroupNorm (gn_fused_small.cu) when the tensors of the pass fit the L2 (second sweep = L2 hits)."""
        nbytes = x.numel() * x.element_size() + (0 if other is None else other.numel() * other.element_size())
        if self.gn_fused_bytes <= 0 or nbytes > self.gn_fused_bytes:
            return False
        B, N, C = ops._bnc(x)
        return ops.gn_fused_supported(B, N, C, G)

    def gn_apply_fwd(self, x, y, prefix, G, post=None, act=ops.ACT_NONE, want_xbar=False, sums=None):
        """GroupNorm forward: y = act(GN(x)) * post (y None: statistics / affine only, the consumer applies it).
        sums: per-(b,c) sum / sum of squares already produced by the conv that wrote x (Engine.conv(gn=True)).
        -> (ab [B][C][2], mean_rstd [B][G][2], xbar [B][C] or None), all needed again by the backward pass."""
        B, N, C = ops._bnc(x)
        gamma, beta = self.P[prefix + ".weight"].detach(), self.P[prefix + ".bias"].detach()
        ab = self._empty(B, C, 2, dtype=torch.float32)
        mr = self._empty(B, G, 2, dtype=torch.float32)
        xbar = self._empty(B, C, dtype=torch.float32) if want_xbar else None
        if self._gn_fused_ok(x, G):
            ops.gn_fused_fwd(x, y, gamma, beta, G, sums, post, act, ab, mr, xbar)
            return ab, mr, xbar
        if sums is None:
            sums = self.fwd_arena.take(B, C, 2)
            ops.chan_stats(x, sums)
        ops.gn_finalize(sums, gamma, beta, ab, mr, xbar, B, C, G, N)
        if y is not None:
            ops.affine_act(x, y, ab, post, act)
        return ab, mr, xbar

    def gn_bwd(self, dy, x, ab, mr, prefix, G, act, post, addbc, dx, accumulate):
        B, N, C = ops._bnc(x)
        if self._gn_fused_ok(x, G, dy):
            ops.gn_fused_bwd(dy, x, ab, mr, self.P[prefix + ".weight"].detach(), G, post, addbc, act, dx, accumulate,
                             self.pg[prefix + ".weight"], self.pg[prefix + ".bias"])
            return
        pq = self.bwd_arena.take(B, C, 2)
        # with an activation, dz = dy * post * act'(.) is written back over dy by the reduce pass (every such dy is
        # consumed only here), so the derivative is evaluated once and the apply pass is a pure stream
        inplace = act != ops.ACT_NONE
        ops.gnact_bwd_reduce(dy, x, ab, post, addbc, act, pq, dy if inplace else None)
        coef = self._empty(B, C, 3, dtype=torch.float32)
        ops.gn_bwd_finalize(pq, mr, self.P[prefix + ".weight"].detach(), coef, self.pg[prefix + ".weight"],
                            self.pg[prefix + ".bias"], B, C, G, N)
        if inplace:
            ops.gnact_bwd_apply(dy, x, ab, None, None, ops.ACT_NONE, coef, dx, accumulate)
        else:
            ops.gnact_bwd_apply(dy, x, ab, post, addbc, act, coef, dx, accumulate)

    # ------------------------------------------------------------------ masks
    def make_masks(self, B):
        """DropPath scales (two per block, in call order: attention branch, Mix-FFN branch; rate linspace(0, .1,
        sum(depths)), simplified_attention.py:143-144,214) and Dropout2d(0.2) scale planes (CamRaDepth.py:96), all from
        ONE Philox launch whose step counter lives on the device (fresh masks on every CUDA-graph replay)."""
        cfg = self.cfg
        nb = sum(cfg.depths)
        rates = torch.linspace(0, DROP_PATH_RATE, nb).tolist()
        n_dp, n_d2 = 2 * nb, cfg.n_dropout_sites
        st = self._mask_state
        if st is None or st["B"] != B or st["dev"] != self.device:
            keep = torch.tensor([1.0 - r for r in rates for _ in range(2)], dtype=torch.float32).to(self.device)
            seed = torch.initial_seed() & 0x7fffffffffffffff
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    seed = (seed + 0x9E3779B97F4A7C15 * (dist.get_rank() + 1)) & 0x7fffffffffffffff   # per-rank masks
            except Exception:
                pass
            state = torch.tensor([seed, 0], dtype=torch.int64).to(self.device)
            buf = torch.empty(n_dp * B + n_d2 * B * MID, dtype=torch.float32, device=self.device)
            st = self._mask_state = dict(B=B, dev=self.device, keep=keep, state=state, buf=buf)
        ops.make_masks(st["buf"], st["keep"], n_dp, B, n_d2, MID, 1.0 - DROPOUT2D_P, st["state"])
        buf = st["buf"]
        dps = [None if rates[i // 2] == 0.0 else buf[i * B:(i + 1) * B] for i in range(n_dp)]     # Identity (:123)
        off = n_dp * B
        d2s = [buf[off + j * B * MID: off + (j + 1) * B * MID].view(B, MID) for j in range(n_d2)]
        return dps, d2s

    # ------------------------------------------------------------------ encoder forward
    def pe_fwd(self, s, xin, save):
        cfg = self.cfg
        name = f"dest_encoder.patch_embed{s + 1}"
        L = self.L[name + ".proj.weight"]
        B, H, W, _ = xin.shape
        Ho = (H + 2 * L["pad"] - L["k"]) // L["stride"] + 1
        Wo = (W + 2 * L["pad"] - L["k"]) // L["stride"] + 1
        C = cfg.dims[s]
        y = self._empty(B, Ho, Wo, C)
        sums = self.conv(xin, name + ".proj.weight", y, bias=name + ".proj.bias", gn=True)
        x0 = self._empty(B, Ho, Wo, C, dtype=torch.float32)
        ab, mr, _ = self.gn_apply_fwd(y, x0, name + ".norm", C // cfg.gn_div, sums=sums)
        rec = dict(xin=xin, y=y, ab=ab, mr=mr) if save else None
        return x0, rec

    def pe_bwd(self, s, rec, dx, dxin, accumulate):
        cfg = self.cfg
        name = f"dest_encoder.patch_embed{s + 1}"
        C = cfg.dims[s]
        dy = self._empty(*rec["y"].shape)
        self.gn_bwd(dx, rec["y"], rec["ab"], rec["mr"], name + ".norm", C // cfg.gn_div, ops.ACT_NONE, None, None,
                    dy, False)
        self.conv_wgrad(rec["xin"], dy, name + ".proj.weight", bias=name + ".proj.bias")
        if dxin is not None:
            self.conv_dgrad(dy, name + ".proj.weight", dxin, accumulate)

    def block_fwd(self, s, i, x, dp, dp_mlp, save):
        cfg = self.cfg
        p = f"dest_encoder.block{s + 1}.{i}"
        B, H, W, C = x.shape
        N = H * W
        heads, sr = cfg.heads[s], cfg.sr[s]
        rC = int(C * cfg.ff[s])
        G = C // cfg.gn_div
        f32 = torch.float32
        x1 = self._empty(B, H, W, C)
        ab1, mr1, xbar1 = self.gn_apply_fwd(x, x1, p + ".norm1", G, want_xbar=True)
        q = self._empty(B, H, W, C)
        self.conv(x1, p + ".attn.q.weight", q, bias=p + ".attn.q.bias")
        rec = {}
        if sr > 1:
            Hs, Ws = H // sr, W // sr
            xs = self._empty(B, Hs, Ws, C)
            sums_s = self.conv(x1, p + ".attn.sr.weight", xs, bias=p + ".attn.sr.bias", gn=True)
            xsn = self._empty(B, Hs, Ws, C)
            ab_s, mr_s, _ = self.gn_apply_fwd(xs, xsn, p + ".attn.norm", G, sums=sums_s)
            kin = xsn
            rec.update(xs=xs, ab_s=ab_s, mr_s=mr_s, xsn=xsn)
        else:
            Hs, Ws = H, W
            kin = x1
        M = Hs * Ws
        k = self._empty(B, Hs, Ws, C)
        self.conv(kin, p + ".attn.k.weight", k, bias=p + ".attn.k.bias")
        sc = self._empty(B, N, dtype=f32)
        idx = torch.empty(B, heads, N, dtype=torch.int16, device=self.device)
        scale = float((C // heads) ** -0.5)
        ops.attn_qkmax_fwd(q.view(B, N, C), k.view(B, M, C), sc, idx, heads, scale)
        pv = self._empty(B, C, dtype=f32)
        ops.attn_pv_fwd(xbar1, self.P[p + ".attn.proj.weight"].detach(), pv)
        x_mid = self._empty(B, H, W, C, dtype=f32)
        ops.attn_out_residual(x.view(B, N, C), pv, sc, self.P[p + ".attn.proj.bias"].detach(), dp,
                              x_mid.view(B, N, C))
        # Mix-FFN
        x2 = self._empty(B, H, W, C)
        ab2, mr2, _ = self.gn_apply_fwd(x_mid, x2, p + ".norm2", G)
        h1 = self._empty(B, H, W, rC)
        sums_m1 = self.conv(x2, p + ".mlp1.fc1.weight", h1, bias=p + ".mlp1.fc1.bias", gn=True)
        ab_m1, mr_m1, _ = self.gn_apply_fwd(h1, None, p + ".mlp1.norm1", rC // cfg.gn_div, sums=sums_m1)  # applied by the dwconv
        h2 = self._empty(B, H, W, rC)
        ops.dwconv_fwd(h1, ab_m1, self.P[p + ".mlp1.dwconv.dwconv.weight"].detach(),
                       self.P[p + ".mlp1.dwconv.dwconv.bias"].detach(), h2)
        h3 = self._empty(B, H, W, rC)
        ab_m2, mr_m2, _ = self.gn_apply_fwd(h2, h3, p + ".mlp1.norm2", G, act=ops.ACT_GELU)
        y2 = self._empty(B, H, W, C)
        self.conv(h3, p + ".mlp1.fc2.weight", y2, bias=p + ".mlp1.fc2.bias")
        x_out = self._empty(B, H, W, C, dtype=f32)
        ops.residual_add(x_mid.view(B, N, C), y2.view(B, N, C), dp_mlp, x_out.view(B, N, C))
        if not save:
            return x_out, None
        rec.update(x=x, ab1=ab1, mr1=mr1, xbar1=xbar1, x1=x1, q=q, k=k, sc=sc, idx=idx, pv=pv, x_mid=x_mid,
                   ab2=ab2, mr2=mr2, x2=x2, h1=h1, ab_m1=ab_m1, mr_m1=mr_m1, h2=h2, ab_m2=ab_m2, mr_m2=mr_m2,
                   h3=h3, dp=dp, dp_mlp=dp_mlp, M=M, scale=scale)
        return x_out, rec

    def block_bwd(self, s, i, rec, dx):
        """dx: fp32 grad wrt the block output; updated IN PLACE to the grad wrt the block input."""
        cfg = self.cfg
        p = f"dest_encoder.block{s + 1}.{i}"
        B, H, W, C = dx.shape
        N = H * W
        heads, sr = cfg.heads[s], cfg.sr[s]
        rC = int(C * cfg.ff[s])
        G = C // cfg.gn_div
        f32 = torch.float32
        dp = rec["dp"]
        # ---- Mix-FFN branch
        dy2 = self._empty(B, H, W, C)
        ops.scale_cast(dx, rec["dp_mlp"], dy2)
        self.conv_wgrad(rec["h3"], dy2, p + ".mlp1.fc2.weight", bias=p + ".mlp1.fc2.bias")
        dh3 = self._empty(B, H, W, rC)
        self.conv_dgrad(dy2, p + ".mlp1.fc2.weight", dh3, False)
        dh2 = self._empty(B, H, W, rC)
        self.gn_bwd(dh3, rec["h2"], rec["ab_m2"], rec["mr_m2"], p + ".mlp1.norm2", G, ops.ACT_GELU, None, None, dh2,
                    False)
        del dh3
        dh1n = self._empty(B, H, W, rC)
        ops.dwconv_bwd(dh2, rec["h1"], rec["ab_m1"], self.P[p + ".mlp1.dwconv.dwconv.weight"].detach(), dh1n,
                       self.pg[p + ".mlp1.dwconv.dwconv.weight"], self.pg[p + ".mlp1.dwconv.dwconv.bias"])
        dh1 = dh2   # reuse
        self.gn_bwd(dh1n, rec["h1"], rec["ab_m1"], rec["mr_m1"], p + ".mlp1.norm1", rC // cfg.gn_div, ops.ACT_NONE,
                    None, None, dh1, False)
        del dh1n
        self.conv_wgrad(rec["x2"], dh1, p + ".mlp1.fc1.weight", bias=p + ".mlp1.fc1.bias")
        dx2 = dy2   # reuse
        self.conv_dgrad(dh1, p + ".mlp1.fc1.weight", dx2, False)
        self.gn_bwd(dx2, rec["x_mid"], rec["ab2"], rec["mr2"], p + ".norm2", G, ops.ACT_NONE, None, None, dx, True)
        # ---- attention branch (dx is now the grad wrt x_mid)
        ds = self._empty(B, N, dtype=f32)
        dpv = self._empty(B, C, dtype=f32)
        tmp = self._empty(B, C, 2, dtype=f32)
        ops.attn_out_bwd(dx.view(B, N, C), rec["pv"], rec["sc"], dp, ds, dpv, self.pg[p + ".attn.proj.bias"], tmp)
        dxbar = self._empty(B, C, dtype=f32)
        ops.attn_pv_bwd(dpv, rec["xbar1"], self.P[p + ".attn.proj.weight"].detach(),
                        self.pg[p + ".attn.proj.weight"], dxbar, 1.0 / N)
        M = rec["M"]
        dq = self._empty(B, H, W, C)
        dk32 = self.bwd_arena.take(B, M, C)
        ops.attn_qkmax_bwd(ds, rec["q"].view(B, N, C), rec["k"].view(B, M, C), rec["idx"], dq.view(B, N, C), dk32,
                           heads, rec["scale"])
        self.conv_wgrad(rec["x1"], dq, p + ".attn.q.weight", bias=p + ".attn.q.bias")
        dx1 = dx2   # reuse (T, B,H,W,C)
        self.conv_dgrad(dq, p + ".attn.q.weight", dx1, False)
        kshape = rec["k"].shape
        dk = self._empty(*kshape)
        ops.scale_cast(dk32.view(*kshape), None, dk)
        if sr > 1:
            self.conv_wgrad(rec["xsn"], dk, p + ".attn.k.weight", bias=p + ".attn.k.bias")
            dxsn = self._empty(*kshape)
            self.conv_dgrad(dk, p + ".attn.k.weight", dxsn, False)
            dxs = dk   # reuse
            self.gn_bwd(dxsn, rec["xs"], rec["ab_s"], rec["mr_s"], p + ".attn.norm", G, ops.ACT_NONE, None, None,
                        dxs, False)
            self.conv_wgrad(rec["x1"], dxs, p + ".attn.sr.weight", bias=p + ".attn.sr.bias")
            self.conv_dgrad(dxs, p + ".attn.sr.weight", dx1, True)
        else:
            self.conv_wgrad(rec["x1"], dk, p + ".attn.k.weight", bias=p + ".attn.k.bias")
            self.conv_dgrad(dk, p + ".attn.k.weight", dx1, True)
        self.gn_bwd(dx1, rec["x"], rec["ab1"], rec["mr1"], p + ".norm1", G, ops.ACT_NONE, None, dxbar, dx, True)

    # ------------------------------------------------------------------ decoder pieces
    def convlayer_fwd(self, x, prefix, dest, post, save, act=ops.ACT_GELU):
        """ConvLayer (utils.py:223-228): conv(no bias) -> GN(Cout/16) -> GELU, output written into `dest`."""
        name = prefix + ".model.0.weight"
        L = self.L[name]
        B, H, W, _ = x.shape
        y = self._empty(B, H, W, L["cout"])
        sums = self.conv(x, name, y, gn=True)
        ab, mr, _ = self.gn_apply_fwd(y, dest, prefix + ".model.1", L["cout"] // self.cfg.gn_div, post=post, act=act,
                                      sums=sums)
        return dict(x=x, y=y, ab=ab, mr=mr, post=post) if save else None

    def convlayer_bwd(self, prefix, rec, ddest, dx, accumulate):
        name = prefix + ".model.0.weight"
        L = self.L[name]
        dy = self._empty(*rec["y"].shape)
        self.gn_bwd(ddest, rec["y"], rec["ab"], rec["mr"], prefix + ".model.1", L["cout"] // self.cfg.gn_div,
                    ops.ACT_GELU, rec["post"], None, dy, False)
        self.conv_wgrad(rec["x"], dy, name)
        if dx is not None:
            self.conv_dgrad(dy, name, dx, accumulate)

    def dec_fwd(self, prefix, src, skip_fn, dest, post, save):
        """Decoder (utils.py:249-257): bicubic x2 -> cat(skip) -> ShortResBlock (utils.py:127-135).

        src: NHWC buffer whose channels (all, zero padded) are upsampled; skip_fn(view) writes the
        skip tensor into its slice of the concat buffer; dest: NHWC view receiving the 128 outputs.
        """
        cx = self.dec_cx[prefix]
        cxp = r8(cx)
        B, h, w, cs = src.shape
        ct = cxp + 160
        cat = self._empty(B, 2 * h, 2 * w, r64(ct))[..., :ct]
        ops.bicubic2x_fwd(src, cat[..., :cs])
        if skip_fn is not None:
            skip_fn(cat)
        recs = []
        recs.append(self.convlayer_fwd(cat[..., :cxp], f"{prefix}.conv.layers.0", cat[..., cxp:cxp + 96], None, save))
        recs.append(self.convlayer_fwd(cat[..., :cxp + 96], f"{prefix}.conv.layers.1", cat[..., cxp + 96:ct], None,
                                       save))
        recs.append(self.convlayer_fwd(cat, f"{prefix}.conv.layers.2", dest, post, save))
        return dict(cat=cat, recs=recs, cs=cs, cxp=cxp) if save else None

    DY_OFF = (0, 96, 160, 288)      # channel offsets of the three layers' dy inside the stacked dy buffer

    def _stacked_dgrad_w(self, prefix):
        """[ct][9][288] bf16: row c = concat-buffer channel, column (tap, k) = weight of the layer whose output
        gradient sits at channel k of the stacked dy buffer (zero where a layer does not read channel c)."""
        names = [f"{prefix}.conv.layers.{li}.model.0.weight" for li in range(3)]
        ps = [self.P[n] for n in names]
        key = ("stack", prefix)
        ver = tuple((p._version, p.data_ptr()) for p in ps) + (WEIGHT_EPOCH[0], self._grad_epoch)
        ent = self._packs.get(key)
        if ent is not None and ent[1] == ver:
            return ent[0]
        ct = self.L[names[2]]["cin_p"]
        ktot = self.DY_OFF[3]
        dst = ent[0] if ent is not None and ent[0].device == self.device else \
            torch.zeros(ct, 9 * ktot, dtype=self.tdtype, device=self.device)
        flat = dst.view(-1)
        jobs = []
        for li, n in enumerate(names):
            L = self.L[n]
            ops.weight_pack(ps[li].detach(), flat[self.DY_OFF[li]:], self._cmap(n), L["cout"], L["cin"], 9,
                            L["cin_p"], ktot, 1)
            jobs.append((n, flat[self.DY_OFF[li]:], self._cmap(n), L["cout"], L["cin"], 9, L["cin_p"], ktot, 1))
        self._packs[key] = (dst, ver)
        self._pack_jobs[key] = (jobs, lambda: tuple((p._version, p.data_ptr()) for p in ps) +
                                (WEIGHT_EPOCH[0], self._grad_epoch))
        return dst

    def dec_bwd(self, prefix, rec, ddest, dsrc, src_accumulate):
        """Returns the grad of the concat buffer (callers slice the skip part out of it)."""
        cat, recs, cs, cxp = rec["cat"], rec["recs"], rec["cs"], rec["cxp"]
        ct = cxp + 160
        dcat = self._empty(cat.shape[0], cat.shape[1], cat.shape[2], r64(ct))[..., :ct]
        if not (self.use_tc and self.tdtype == torch.bfloat16):
            self.convlayer_bwd(f"{prefix}.conv.layers.2", recs[2], ddest, dcat, False)
            self.convlayer_bwd(f"{prefix}.conv.layers.1", recs[1], dcat[..., cxp + 96:ct], dcat[..., :cxp + 96], True)
            self.convlayer_bwd(f"{prefix}.conv.layers.0", recs[0], dcat[..., cxp:cxp + 96], dcat[..., :cxp], True)
        else:
            # Data gradients by OUTPUT-channel block with concatenated K: the block of concat channels that only
            # layer 2 reads gets one GEMM over dy2, the block layers 1+2 read gets one GEMM over [dy1|dy2], the block
            # all three read one GEMM over [dy0|dy1|dy2].  Same MACs as three per-layer dgrads, but every output
            # element is written once (no read-modify-write epilogue) and K is 2-3x longer.
            B_, H_, W_, _ = cat.shape
            off = self.DY_OFF
            dycat = self._empty(B_, H_, W_, r64(off[3]))[..., :off[3]]
            wb = self._stacked_dgrad_w(prefix)
            row_lo = (0, cxp, cxp + 96)              # first concat channel of the block finished after layer li
            row_hi = (cxp, cxp + 96, ct)
            ddests = (dcat[..., cxp:cxp + 96], dcat[..., cxp + 96:ct], ddest)
            for li in (2, 1, 0):
                name = f"{prefix}.conv.layers.{li}"
                r = recs[li]
                L = self.L[name + ".model.0.weight"]
                dy = dycat[..., off[li]:off[li + 1]]
                self.gn_bwd(ddests[li], r["y"], r["ab"], r["mr"], name + ".model.1", L["cout"] // self.cfg.gn_div,
                            ops.ACT_GELU, r["post"], None, dy, False)
                self.conv_wgrad(r["x"], dy, name + ".model.0.weight")
                a = dycat[..., off[li]:]
                out = dcat[..., row_lo[li]:row_hi[li]]
                d = ops.make_desc(a, out, a.shape[-1], row_hi[li] - row_lo[li], 3, 3, 1, 1, transposed=1,
                                  w_tap_stride=off[3], w_koff=off[li])
                ev = self._timed("dgrad", name + ".model.0.weight")
                if ev is not None:
                    # this launch is the OUTPUT-channel block finished after layer li: rows row_lo..row_hi of the
                    # concat gradient from the stacked [dy_li | .. | dy_2] (K = 9 * its width), not layer li's own dgrad
                    self.timed_flops[("dgrad", name + ".model.0.weight")] = \
                        2.0 * B_ * H_ * W_ * (row_hi[li] - row_lo[li]) * 9 * a.shape[-1]
                ops.conv_fwd(d, a, wb[row_lo[li]:row_hi[li]], None, out, use_tc=True)
                if ev is not None:
                    ev.record()
        if dsrc is not None:
            ops.bicubic2x_bwd(dcat[..., :cs], dsrc, src_accumulate)
        return dcat

    def da_fwd(self, name, F, save):
        """Depth_Activation (utils.py:285-289): conv3x3(+bias) -> sigmoid -> conv3x3(32->1, +bias)."""
        L = self.L[name + ".conv_1.weight"]
        B, H, W, _ = F.shape
        sg = self._empty(B, H, W, 32)
        self.conv(F[..., :L["cin_p"]], name + ".conv_1.weight", sg, bias=name + ".conv_1.bias", act=ops.ACT_SIGMOID)
        w2 = self.wpack(name + ".conv_2.weight", 0, dtype=torch.float32)
        depth = self._empty(B, 1, H, W, dtype=torch.float32)
        ops.conv3x3_c1_fwd(sg, w2, self.P[name + ".conv_2.bias"].detach(), depth)
        return depth, (dict(F=F, sg=sg) if save else None)

    def da_bwd(self, name, rec, ddepth, dF, accumulate):
        L = self.L[name + ".conv_1.weight"]
        sg, F = rec["sg"], rec["F"]
        w2 = self.wpack(name + ".conv_2.weight", 0, dtype=torch.float32)
        dpre = self._empty(*sg.shape)
        dw2 = self.bwd_arena.take(1, 9 * 32)
        # data gradient of conv_2 times the sigmoid derivative in one pass (no dsg buffer)
        ops.conv3x3_c1_bwd_sigmoid(ddepth, sg, w2, dpre, dw2, self.pg[name + ".conv_2.bias"])
        ops.weight_unpack_grad(dw2, self.pg[name + ".conv_2.weight"], None, 1, 32, 9, 32, False)
        self.conv_wgrad(F[..., :L["cin_p"]], dpre, name + ".conv_1.weight", bias=name + ".conv_1.bias")
        self.conv_dgrad(dpre, name + ".conv_1.weight", dF[..., :L["cin_p"]], accumulate)

    def seg_logits(self, F, name, out_dtype):
        """3x3 conv 128 -> ncls (+bias) (CamRaDepth.py:129,133,155,159); NHWC logits padded to 24 channels."""
        L = self.L[name + ".weight"]
        B, H, W, _ = F.shape
        lg = self._zeros(B, H, W, L["cout_p"], dtype=out_dtype)
        self.conv(F[..., :MID], name + ".weight", lg, bias=name + ".bias")
        return lg

    def seg_map(self, F, name, ncls, map0=None, map1=None, map_f32=None):
        """Seg_Block (utils.py:95-100) for heads whose logits feed nothing but the argmax: map = argmax_c(conv) / ncls
        written into the given channel views / fp32 map.  Tensor-core path: fused into the conv's accumulator
        read-out (fp32 logits, never materialised); otherwise conv + argmax_map."""
        L = self.L[name + ".weight"]
        x = F[..., :MID]
        if self.use_tc and self._tc_ok(L, x, None):
            B, H, W, _ = F.shape
            d = ops.make_desc(x, x, L["cin_p"], L["cout"], 3, 3, 1, 1, out_dtype=ops.BF16)
            ops.conv_argmax(d, x, self.wpack(name + ".weight", 0), self.P[name + ".bias"].detach(), ncls, map0, map1,
                            map_f32)
            return
        lg = self.seg_logits(F, name, torch.float32)
        for m in (map0, map1):
            if m is not None:
                ops.argmax_map(lg, ncls, m, None)
        if map_f32 is not None:
            ops.argmax_map(lg, ncls, None, map_f32)

    # ------------------------------------------------------------------ whole forward
    def forward(self, x, train, masks=None, save=True, packed=False):
        """x: (B,cin,H,W) fp32 NCHW (the nn.Module surface), or with packed=True the engine's own input layout
        (B,H,W,r8(cin)) bf16 NHWC as written by preprocess.pack_input_nhwc (no layout pack in the step)."""
        cfg = self.cfg
        if not x.is_cuda:
            raise RuntimeError("camradepth_b200 runs on CUDA devices only (no CPU fallback by design)")
        if packed:
            if self.tdtype != torch.bfloat16 or x.dtype != torch.bfloat16 or x.dim() != 4 or x.shape[-1] != r8(cfg.cin):
                raise RuntimeError(f"packed input must be (B,H,W,{r8(cfg.cin)}) bf16 NHWC in bf16 mode")
            x = x.detach().contiguous()
        if self.device != x.device:
            self.device = x.device
            self.fwd_arena = ZeroArena(self.device)
            self.bwd_arena = ZeroArena(self.device)
        if packed:
            (B, H, W, _), cin = x.shape, cfg.cin
        else:
            B, cin, H, W = x.shape
        if cin != cfg.cin:
            raise RuntimeError(f"expected {cfg.cin} input channels, got {cin}")
        if H % 32 or W % 32:
            raise RuntimeError(f"Sizes of tensors must match: H and W must be multiples of 32, got {H}x{W}")
        if not packed:
            x = x.detach().contiguous().float()
        if save:
            # training forward: parameters may have been updated through `.data` (no version bump, e.g. the
            # reference's own optimizer), so packed copies are rebuilt every grad-enabled forward
            self._grad_epoch += 1
            self.prepack()
        self.fwd_arena.reset()
        if train:
            dps, d2s = masks if masks is not None else self.make_masks(B)
            if len(dps) != 2 * sum(cfg.depths) or len(d2s) != cfg.n_dropout_sites:
                raise RuntimeError(f"expected {2 * sum(cfg.depths)} DropPath scales (two per block: attention branch, "
                                   f"Mix-FFN branch) and {cfg.n_dropout_sites} Dropout2d scales, got {len(dps)} / {len(d2s)}")
        else:
            dps, d2s = [None] * (2 * sum(cfg.depths)), [None] * cfg.n_dropout_sites
        S = dict(B=B, H=H, W=W)
        f32 = torch.float32

        # ---- encoder (SimplifiedTransformer.forward_features, simplified_attention.py:265-306)
        if packed:
            X0 = x
        else:
            X0 = self._zeros(B, H, W, r8(cin))
            ops.nchw_to_nhwc(x, X0)
        cur = X0
        stage_T, pe_recs, blk_recs = [], [], []
        bi = 0
        for s in range(4):
            xr, prec = self.pe_fwd(s, cur, save)
            pe_recs.append(prec)
            brs = []
            for i in range(cfg.depths[s]):
                dp, dp_mlp = (None if m is None else m.to(self.device, f32).contiguous()
                              for m in (dps[2 * bi], dps[2 * bi + 1]))
                xr, brec = self.block_fwd(s, i, xr, dp, dp_mlp, save)
                brs.append(brec)
                bi += 1
            st = self._empty(*xr.shape)
            ops.scale_cast(xr, None, st)
            blk_recs.append(brs)
            stage_T.append(st)
            cur = st
        S.update(pe=pe_recs, blk=blk_recs, stage_T=stage_T)

        # ---- decoder (CamRaDepth.dest_decoder, CamRaDepth.py:99-170)
        d2 = [None if m is None else m.to(self.device, f32).contiguous() for m in d2s]
        h0, w0 = H // 32, W // 32
        E1 = self._empty(B, h0, w0, cfg.dims[3])
        S["fe1"] = self.convlayer_fwd(stage_T[3], "from_encoder_1", E1, None, save)
        fe = {}

        def skip_from(j, stage, c0):
            def fn(cat):
                C = stage.shape[-1]
                fe[j] = self.convlayer_fwd(stage, f"from_encoder_{j}", cat[..., c0:c0 + C], None, save)
            return fn

        F1 = self._empty(B, 2 * h0, 2 * w0, MID)
        S["D0"] = self.dec_fwd("depth_upsample.0", E1, skip_from(2, stage_T[2], cfg.dims[3]), F1, d2[0], save)
        F2 = self._empty(B, 4 * h0, 4 * w0, MID)
        S["D1"] = self.dec_fwd("depth_upsample.1", F1, skip_from(3, stage_T[1], MID), F2, d2[1], save)
        FW = r8(MID + 1 + cfg.nseg)                    # [feat 128 | depth | seg maps | zero pad]
        F3 = self._feat(B, 8 * h0, 8 * w0, FW)
        S["D2"] = self.dec_fwd("depth_upsample.2", F2, skip_from(4, stage_T[0], MID), F3[..., :MID], d2[2], save)
        S["fe"] = fe
        inter3, S["DA3"] = self.da_fwd("depth_activation_3", F3, save)
        ops.nchw_to_nhwc(inter3, F3[..., MID:MID + 1])
        F4 = self._feat(B, 16 * h0, 16 * w0, FW)
        S["D3"] = self.dec_fwd("depth_upsample.3", F3, None, F4[..., :MID], d2[3], save)
        seg = cfg.sup or cfg.unsup
        FS4 = None
        if seg:
            FS4 = self._feat(B, 16 * h0, 16 * w0, FW)
            S["S0"] = self.dec_fwd("seg_upsample.0", F3, None, FS4[..., :MID], d2[4], save)
            if cfg.sup:
                self.seg_map(FS4, "seg_conv_stage_4", cfg.num_classes, FS4[..., MID:MID + 1], F4[..., MID + 1:MID + 2])
            if cfg.unsup:
                j = MID + 1 + int(cfg.sup)
                self.seg_map(FS4, "unsup_stage_4", 19, F4[..., j:j + 1], None if cfg.sup else FS4[..., MID:MID + 1])
        inter4, S["DA4"] = self.da_fwd("depth_activation_4", F4, save)
        ops.nchw_to_nhwc(inter4, F4[..., MID:MID + 1])

        def skip_input(cat):
            if packed:
                ops.copy_channels(X0[..., :cin], cat[..., MID + 1:MID + 1 + cin])
            else:
                ops.nchw_to_nhwc(x, cat[..., MID + 1:MID + 1 + cin])

        # the last feature map feeds only the heads: without seg maps nothing reads its tail channels
        da5_w = self.L["depth_activation_5.conv_1.weight"]["cin_p"]
        F5 = self._empty(B, H, W, MID) if (not seg and da5_w == MID) else self._feat(B, H, W, FW)
        S["D4"] = self.dec_fwd("depth_upsample.4", F4, skip_input, F5[..., :MID], d2[5 if seg else 4], save)
        final_seg = unsup_map = None
        if seg:
            FS5 = self._empty(B, H, W, MID)
            S["S1"] = self.dec_fwd("seg_upsample.1", FS4, skip_input, FS5, d2[6], save)
            S["FS5"] = FS5
            if cfg.sup:
                lg = self.seg_logits(FS5, "seg_conv_final", f32)
                final_seg = self._empty(B, cfg.num_classes, H, W, dtype=f32)
                ops.nhwc_to_nchw(lg, final_seg)
                ops.argmax_map(lg, cfg.num_classes, F5[..., MID + 1:MID + 2], None)
            if cfg.unsup:
                unsup_map = self._empty(B, 1, H, W, dtype=f32)
                j = MID + 1 + int(cfg.sup)
                self.seg_map(FS5, "unsup_final", 19, F5[..., j:j + 1], None, unsup_map)
        final_depth, S["DA5"] = self.da_fwd("depth_activation_5", F5, save)
        outs = dict(final_depth=final_depth, inter3=inter3, inter4=inter4, final_seg=final_seg, unsup_map=unsup_map)
        return outs, (S if save else None)

    # ------------------------------------------------------------------ whole backward
    def backward(self, S, g_final, g3, g4, g_seg, flat_grad=None, on_bucket=None):
        """-> dict name -> fp32 grad view (params without a gradient path are absent)."""
        for tag in self.backward_steps(S, g_final, g3, g4, g_seg, flat_grad):
            if on_bucket:
                on_bucket(tag)
        grads, self._grads_out = self._grads_out, None
        return grads

    def backward_steps(self, S, g_final, g3, g4, g_seg, flat_grad=None):
        """The backward program as a generator: runs up to the next point where a bucket of the flat gradient buffer
        is final, yields its tag ("heads", "decoder", "stage3" .. "stage0") and continues on the next `next()`.
        Callers can start the bucket's all-reduce (parallel.py) or close one CUDA-graph capture and open the next
        (graphs.GraphedDataParallelStep) at every yield.  The gradient dict is left in `self._grads_out`."""
        cfg = self.cfg
        B, H, W = S["B"], S["H"], S["W"]
        f32 = torch.float32
        self.bwd_arena.reset()
        sizes = [self.P[n].numel() for n in self.names]
        total = sum(sizes)
        if flat_grad is not None:
            flat = flat_grad
        else:
            # One persistent flat buffer (stable addresses: optimizer tables and CUDA graphs stay valid).  If the
            # caller still holds gradients that alias it (no zero_grad(set_to_none=True) since the last
            # backward, i.e. gradient accumulation), a fresh buffer is used and autograd accumulates into theirs.
            own = self._flat_own
            aliased = own is None or own.numel() != total or own.device != self.device
            if not aliased:
                lo, hi = own.data_ptr(), own.data_ptr() + own.numel() * 4
                for p_ in (self.P[self.names[0]], self.P[self.names[-1]], self.P[self.names[len(self.names) // 2]]):
                    if p_.grad is not None and lo <= p_.grad.data_ptr() < hi:
                        aliased = True
                        break
                if aliased:
                    own = None
            if own is None or own.numel() != total or own.device != self.device:
                flat = torch.zeros(total, dtype=f32, device=self.device)
                if self._flat_own is None or self._flat_own.numel() != total or self._flat_own.device != self.device:
                    self._flat_own = flat
            else:
                flat = own
                flat.zero_()
        self.pg = {}
        off = 0
        self.pg_offsets = {}
        for n, sz in zip(self.names, sizes):
            self.pg[n] = flat[off:off + sz].view(self.P[n].shape)
            self.pg_offsets[n] = (off, sz)
            off += sz
        self.flat_grad = flat
        h0, w0 = H // 32, W // 32
        seg = cfg.sup or cfg.unsup
        FW = r8(MID + 1 + cfg.nseg)

        def gz(g, shape):
            if g is None:
                return torch.zeros(*shape, dtype=f32, device=self.device)
            return g.detach().contiguous().float()

        # ---- heads at full resolution
        # the depth head's data gradient (accumulate=False) overwrites every channel it reads; a zero fill is only
        # needed when the buffer is wider than that
        da5_w = self.L["depth_activation_5.conv_1.weight"]["cin_p"]
        # only dF5[..., :MID] is consumed below (the decoder's output channels): no tail, no fill
        dF5 = self._empty(B, H, W, max(da5_w, MID))
        if da5_w < MID:
            ops.zero_channels(dF5[..., da5_w:])
        self.da_bwd("depth_activation_5", S["DA5"], gz(g_final, (B, 1, H, W)), dF5, False)
        sup_grad = cfg.sup and g_seg is not None
        dFS4 = None
        if sup_grad:
            dlg = self._zeros(B, H, W, r8(cfg.num_classes))
            ops.nchw_to_nhwc(gz(g_seg, ()), dlg)
            FS5 = S["FS5"]
            self.conv_wgrad(FS5, dlg, "seg_conv_final.weight", bias="seg_conv_final.bias")
            dFS5 = self._empty(B, H, W, MID)
            self.conv_dgrad(dlg, "seg_conv_final.weight", dFS5, False)
            dFS4 = self._empty(B, 16 * h0, 16 * w0, FW)
            self.dec_bwd("seg_upsample.1", S["S1"], dFS5, dFS4, False)
            del dFS5, dlg
        dF4 = self._empty(B, 16 * h0, 16 * w0, FW)
        self.dec_bwd("depth_upsample.4", S["D4"], dF5[..., :MID], dF4, False)
        del dF5
        # inter_depth_4 receives grad from the loss and from the upsampled concat (CamRaDepth.py:144)
        d4 = gz(g4, (B, 1, 16 * h0, 16 * w0)).clone() if g4 is not None else self._zeros(B, 1, 16 * h0, 16 * w0, dtype=f32)
        t4 = self._empty(B, 1, 16 * h0, 16 * w0, dtype=f32)
        ops.nhwc_to_nchw(dF4[..., MID:MID + 1], t4)
        ops.add_f32(d4, t4)
        self.da_bwd("depth_activation_4", S["DA4"], d4, dF4, True)
        dF3 = self._empty(B, 8 * h0, 8 * w0, FW)
        acc3 = False
        if sup_grad:
            self.dec_bwd("seg_upsample.0", S["S0"], dFS4[..., :MID], dF3, False)
            acc3 = True
            del dFS4
        self.dec_bwd("depth_upsample.3", S["D3"], dF4[..., :MID], dF3, acc3)
        del dF4
        d3 = gz(g3, (B, 1, 8 * h0, 8 * w0)).clone() if g3 is not None else self._zeros(B, 1, 8 * h0, 8 * w0, dtype=f32)
        t3 = self._empty(B, 1, 8 * h0, 8 * w0, dtype=f32)
        ops.nhwc_to_nchw(dF3[..., MID:MID + 1], t3)
        ops.add_f32(d3, t3)
        self.da_bwd("depth_activation_3", S["DA3"], d3, dF3, True)
        yield "heads"
        # ---- pyramid
        stage_T = S["stage_T"]
        dstage = [self._empty(*t.shape) for t in stage_T]
        dF2 = self._empty(B, 4 * h0, 4 * w0, MID)
        dcat = self.dec_bwd("depth_upsample.2", S["D2"], dF3[..., :MID], dF2, False)
        self.convlayer_bwd("from_encoder_4", S["fe"][4], dcat[..., MID:MID + cfg.dims[0]], dstage[0], False)
        del dF3, dcat
        dF1 = self._empty(B, 2 * h0, 2 * w0, MID)
        dcat = self.dec_bwd("depth_upsample.1", S["D1"], dF2, dF1, False)
        self.convlayer_bwd("from_encoder_3", S["fe"][3], dcat[..., MID:MID + cfg.dims[1]], dstage[1], False)
        del dF2, dcat
        dE1 = self._empty(B, h0, w0, cfg.dims[3])
        dcat = self.dec_bwd("depth_upsample.0", S["D0"], dF1, dE1, False)
        self.convlayer_bwd("from_encoder_2", S["fe"][2], dcat[..., cfg.dims[3]:cfg.dims[3] + cfg.dims[2]], dstage[2],
                           False)
        del dF1, dcat
        self.convlayer_bwd("from_encoder_1", S["fe1"], dE1, dstage[3], False)
        yield "decoder"
        # ---- encoder, last stage first
        for s in (3, 2, 1, 0):
            dx = torch.zeros(*stage_T[s].shape, dtype=f32, device=self.device)
            ops.add_f32(dx, dstage[s])
            for i in reversed(range(cfg.depths[s])):
                self.block_bwd(s, i, S["blk"][s][i], dx)
            self.pe_bwd(s, S["pe"][s], dx, dstage[s - 1] if s > 0 else None, True)
            if s == 0:
                grads = dict(self.pg)
                for n in self.no_grad_names(sup_grad):
                    grads.pop(n, None)
                self.pg = None
                self._grads_out = grads
            yield f"stage{s}"

    def no_grad_names(self, sup_grad=True):
        """Parameters with no gradient path: argmax-fed heads; the whole seg branch if unsupervised."""
        cfg = self.cfg
        out = []
        if cfg.sup:
            out += ["seg_conv_stage_4.weight", "seg_conv_stage_4.bias"]
        if cfg.unsup:
            out += ["unsup_stage_4.weight", "unsup_stage_4.bias", "unsup_final.weight", "unsup_final.bias"]
        if (cfg.sup or cfg.unsup) and not sup_grad:
            out += [n for n in self.names if n.startswith("seg_upsample.") or n.startswith("seg_conv_final")]
        return out
