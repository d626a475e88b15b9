"""Generate golden fixtures by running the REAL reference (imported from /root/reference).

Test infrastructure.  Runs only in the build container (the reference does not exist on the
GPU box); writes small tensors to tests/golden/*.pt which travel with the repo.

    python oracle/make_golden.py            # regenerates every fixture

Recipe: PYTHONPATH = oracle/shims + /root/reference/src, argv carries a valid
--split / --output_dir / --model BEFORE `utils.args` is imported (args.py:65,148-153).  `args` is a
process global, so each variant is generated in its own subprocess.
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("CAMRADEPTH_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")

CASES = [
    # (variant, B, H, W, train_masks)
    ("base", 2, 64, 64, False),
    ("base", 2, 64, 96, True),
    ("supervised_seg", 1, 64, 64, False),
    ("unsupervised_seg", 1, 64, 64, False),
    ("sup_unsup_seg", 2, 64, 64, True),
    ("base (rgb)", 1, 64, 64, False),
]
FULL_GRADS = [  # tensors whose complete gradient is stored
    "dest_encoder.patch_embed1.proj.weight",
    "dest_encoder.block1.0.attn.q.weight",
    "dest_encoder.block1.0.attn.sr.weight",
    "dest_encoder.block2.3.mlp1.dwconv.dwconv.weight",
    "dest_encoder.block3.5.attn.k.weight",
    "dest_encoder.block4.4.mlp1.fc2.weight",
    "dest_encoder.block4.4.norm2.weight",
    "from_encoder_1.model.0.weight",
    "depth_upsample.0.conv.layers.0.model.1.weight",
    "depth_upsample.4.conv.layers.2.model.0.weight",
    "depth_activation_5.conv_1.weight",
    "depth_activation_3.conv_2.weight",
]


def worker(variant, B, H, W, train, out_path):
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, os.path.join(REF, "src"))
    sys.path.insert(0, REPO)
    tmp = tempfile.mkdtemp()
    sys.argv = [sys.argv[0], "--split", os.path.join(REF, "src/data/new_split.npy"),
                "--output_dir", tmp, "--model", variant]
    import torch
    import torch.nn as nn
    torch.manual_seed(0)
    torch.set_num_threads(8)
    from utils.args import args                      # noqa  (reference)
    from models.CamRaDepth import CamRaDepth          # noqa  (reference)
    from models.diffGradNorm import diffGradNorm      # noqa  (reference)
    from utils.loss_funcs import MaskedSmoothL1Loss, MaskedFocalLoss, MaskedMSELoss  # noqa
    from oracle import camradepth_oracle as O
    from camradepth_b200.synthetic import make_batch

    cfg = O.Cfg(variant)
    model = CamRaDepth(input_channels=args.input_channels)
    ref_sd = model.state_dict()
    spec = O.param_spec(cfg)
    assert list(ref_sd.keys()) == list(spec.keys()), "oracle param order != reference"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == spec[k][0], (k, v.shape, spec[k][0])
    sd = O.init_state_dict(cfg, seed=1, perturb=0.05)
    model.load_state_dict(sd, strict=True)

    batch = make_batch(B, H, W, seed=3, input_channels=cfg.cin)
    x = batch["image"]
    dps = d2s = None
    if train:
        dps, d2s = O.make_masks(cfg, B, seed=11)
        model.train()

        class Inject(nn.Module):
            def __init__(self, q):
                super().__init__()
                self.q = q

            def forward(self, t):
                m = self.q.pop(0)
                return t * m.view(*m.shape, *([1] * (t.dim() - m.dim())))

        # timm's DropPath draws a fresh per-sample mask on EVERY call and Block.forward calls it twice
        # (simplified_attention.py:143-144): the injected module pops one mask per call, in call order
        dq = list(dps)
        bi = 0
        for s in range(4):
            for blk in getattr(model.dest_encoder, f"block{s + 1}"):
                blk.drop_path = Inject([dq[2 * bi], dq[2 * bi + 1]])
                bi += 1
        model.dropout = Inject(list(d2s))
    else:
        model.eval()

    pred = model(x)
    crit_d, crit_s = MaskedSmoothL1Loss(), MaskedFocalLoss()
    inter = pred["depth"]["intermediate_depths"]
    fs = pred["seg"]["final_seg"]
    l_seg = (crit_s(fs, batch["gt_seg"]) if fs is not None else 0) * args.supervised_seg
    l4 = crit_d(inter[-1].squeeze(1), batch["gt_s4"].squeeze(1))
    l3 = crit_d(inter[-2].squeeze(1), batch["gt_s3"].squeeze(1))
    lf = crit_d(pred["depth"]["final_depth"], batch["gt_final"])
    w = [1, 1, 1, 0.2, 0.2]
    loss = (w[0] * lf + w[1] * l4 + w[2] * l3 + w[3] * l_seg + w[4] * 0) / sum(w)
    loss.backward()
    rmse = torch.sqrt(MaskedMSELoss()(pred["depth"]["final_depth"], batch["gt_final"]))

    gnames, gvals = [], []
    none_grads = []
    for n, p in model.named_parameters():
        if p.grad is None:
            none_grads.append(n)
        else:
            gnames.append(n)
            gvals.append(torch.stack([p.grad.double().sum(), p.grad.double().norm()]).float())
    gstats = {"names": gnames, "sum_norm": torch.stack(gvals)}
    # strided samples (<= ~4096 values) of selected gradients: flatten()[::max(1, numel // 4096)]
    full = {n: p.grad.flatten()[::max(1, p.grad.numel() // 4096)].clone()
            for n, p in model.named_parameters() if n in FULL_GRADS and p.grad is not None}

    # one (two) reference optimizer steps on a few tensors
    opt_names = ["dest_encoder.block1.0.attn.q.weight", "depth_activation_5.conv_1.weight",
                 "dest_encoder.block4.4.norm2.weight"]
    named = dict(model.named_parameters())
    opt = diffGradNorm([named[n] for n in opt_names], lr=6e-5)
    opt.step()
    def samp(t):
        return t.detach().flatten()[::max(1, t.numel() // 4096)].clone()
    after1 = {n: samp(named[n]) for n in opt_names}
    for n in opt_names:          # second step with a deterministic different gradient
        named[n].grad = named[n].grad * 0.5 + 0.01
    opt.step()
    after2 = {n: samp(named[n]) for n in opt_names}

    out = {
        "variant": variant, "B": B, "H": H, "W": W, "train": train,
        "final_depth": pred["depth"]["final_depth"].detach(),
        "inter3": inter[-2].detach(), "inter4": inter[-1].detach(),
        "final_seg_sample": None if fs is None else fs.detach()[:, :, ::4, ::4].contiguous(),
        "final_seg_sum": None if fs is None else fs.detach().double().sum().float(),
        "unsup_map": None if pred["seg"]["unsup_map"] is None else pred["seg"]["unsup_map"].detach().to(torch.float32),
        "losses": torch.tensor([float(v) for v in (lf.detach(), l4.detach(), l3.detach(), torch.as_tensor(l_seg).detach(), loss.detach(), rmse.detach())]),
        "grad_stats": gstats, "none_grads": none_grads, "full_grads": full,
        "opt_after1": after1, "opt_after2": after2,
    }
    torch.save(out, out_path)
    print("wrote", out_path, "loss", float(loss), "none_grads", len(none_grads))


def main():
    os.makedirs(GOLD, exist_ok=True)
    for (variant, B, H, W, train) in CASES:
        tag = variant.replace(" ", "").replace("(", "_").replace(")", "")
        out = os.path.join(GOLD, f"ref_{tag}_{B}x{H}x{W}_{'train' if train else 'eval'}.pt")
        subprocess.check_call([sys.executable, __file__, "--worker", variant, str(B), str(H), str(W),
                               str(int(train)), out])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        _, _, variant, B, H, W, train, out = sys.argv
        worker(variant, int(B), int(H), int(W), bool(int(train)), out)
    else:
        main()
