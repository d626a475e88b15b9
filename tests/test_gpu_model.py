"""Whole-path parity (GPU): the drop-in CamRaDepth module + fused losses + diffGradNorm against
(1) golden outputs/gradients of the REAL reference (tests/golden) and
(2) the CPU oracle at a BASELINE-sized sample.

Tolerances (SURVEY.md §8c calibration):
  fp32 mode : final depth rel-L2 <= 1e-4; per-tensor grad rel-L2 <= 2e-3 (argmax-routed attn tensors 2e-2)
  bf16 mode : final depth rel-L2 <= 1e-2; global grad rel-L2 <= 3e-2, cosine >= 0.999;
              attn q/k/sr/norm tensors (argmax flips under bf16 rounding) cosine >= 0.95 reported separately
"""
import json
import os

import pytest
import torch

from tests.golden_util import golden_files, load_case, relerr, samp, oracle_run, depth_tol

pytestmark = pytest.mark.gpu
FILES = golden_files()
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def build_model(variant, sd, precision):
    import camradepth_b200 as C
    C.set_model(variant)
    m = C.CamRaDepth(input_channels=C.args.input_channels, precision=precision)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def run_product(variant, cfg, sd, batch, masks, precision, train):
    import camradepth_b200 as C
    m = build_model(variant, sd, precision)
    m.train(train)
    if train:
        m.set_stochastic_masks(masks[0], masks[1])
    dev = torch.device("cuda:0")
    x = batch["image"].to(dev)
    pred = m(x)
    crit_d, crit_s = C.MaskedSmoothL1Loss(), C.MaskedFocalLoss()
    inter = pred["depth"]["intermediate_depths"]
    fs = pred["seg"]["final_seg"]
    l_seg = (crit_s(fs, batch["gt_seg"].to(dev)) if fs is not None else 0) * cfg.sup
    l4 = crit_d(inter[-1].squeeze(1), batch["gt_s4"].to(dev).squeeze(1))
    l3 = crit_d(inter[-2].squeeze(1), batch["gt_s3"].to(dev).squeeze(1))
    lf = crit_d(pred["depth"]["final_depth"], batch["gt_final"].to(dev))
    w = [1, 1, 1, 0.2, 0.2]
    loss = (w[0] * lf + w[1] * l4 + w[2] * l3 + w[3] * l_seg + w[4] * 0) / sum(w)
    loss.backward()
    torch.cuda.synchronize()
    return m, pred, loss, (lf, l4, l3, l_seg)


ATTN_KEYS = (".attn.q.", ".attn.k.", ".attn.sr.", ".attn.norm.")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.basename(p))
def test_model_matches_reference_golden(path, precision):
    g, cfg, sd, batch, masks = load_case(path)
    m, pred, loss, parts = run_product(g["variant"], cfg, sd, batch, masks, precision, g["train"])
    tol = depth_tol(precision, os.path.basename(path))
    e_final = relerr(pred["depth"]["final_depth"], g["final_depth"])
    e3 = relerr(pred["depth"]["intermediate_depths"][2], g["inter3"])
    e4 = relerr(pred["depth"]["intermediate_depths"][3], g["inter4"])
    # north_star states the bf16 bound for the (final) forward depth; the small intermediate maps get 2x
    tol_i = tol if precision == "fp32" else 2 * tol
    assert e_final < tol and e3 < tol_i and e4 < tol_i, (e_final, e3, e4)
    if g["final_seg_sample"] is not None:
        assert relerr(pred["seg"]["final_seg"][:, :, ::4, ::4], g["final_seg_sample"]) < tol_i
    else:
        assert pred["seg"]["final_seg"] is None
    if g["unsup_map"] is not None:
        um = pred["seg"]["unsup_map"]
        assert um.shape == g["unsup_map"].shape and um.dtype == torch.float32
        flips = float((um.cpu() != g["unsup_map"]).float().mean())
        assert flips < (1e-3 if precision == "fp32" else 5e-2), flips
    assert abs(float(loss) - float(g["losses"][4])) < tol * max(1.0, abs(float(g["losses"][4])))
    # gradient structure and values
    named = dict(m.named_parameters())
    none = sorted(n for n, p in named.items() if p.grad is None)
    assert none == sorted(g["none_grads"])
    names, ref = g["grad_stats"]["names"], g["grad_stats"]["sum_norm"]
    worst, worst_attn = ("", 0.0), ("", 0.0)
    for i, n in enumerate(names):
        gn = float(named[n].grad.double().norm())
        e = abs(gn - float(ref[i, 1])) / (float(ref[i, 1]) + 1e-12)
        if any(k in n for k in ATTN_KEYS):
            if e > worst_attn[1]:
                worst_attn = (n, e)
        elif e > worst[1]:
            worst = (n, e)
    report = {"case": os.path.basename(path), "precision": precision, "final_depth_rel_l2": e_final,
              "worst_grad_norm_err": worst, "worst_attn_grad_norm_err": worst_attn, "full_grads": {}}
    gt = 2e-3 if precision == "fp32" else 8e-2
    for n, gs in g["full_grads"].items():
        mine = samp(named[n].grad).cpu()
        e = relerr(mine, gs)
        cos = float(torch.nn.functional.cosine_similarity(mine.double(), gs.double(), dim=0))
        report["full_grads"][n] = (e, cos)
        if any(k in n for k in ATTN_KEYS):
            # bf16: observed 0.983 .. 0.9996 over the six fixtures (profiles/r2_parity_report.jsonl); a wrong tap
            # order / transposition in the sr or q / k gradients gives |cos| < 0.5.  The bf16-only kernels behind these
            # tensors (strided implicit GEMM, depth-to-space data gradient, tcgen05 score) are pinned one by one in
            # test_gpu_tc.py / test_gpu_ops.py; fp32 mode pins the program around them at 0.999.
            assert cos > (0.999 if precision == "fp32" else 0.95), (n, e, cos)
        else:
            assert e < gt, (n, e, cos)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity_report.jsonl"), "a") as fh:
        fh.write(json.dumps(report) + "\n")
    assert worst[1] < (2e-3 if precision == "fp32" else 8e-2), worst
    assert worst_attn[1] < (2e-2 if precision == "fp32" else 0.3), worst_attn      # observed <= 0.11


@pytest.mark.parametrize("path", FILES[:1], ids=lambda p: os.path.basename(p))
def test_optimizer_step_matches_reference_golden(path):
    import camradepth_b200 as C
    g, cfg, sd, batch, masks = load_case(path)
    m, pred, loss, parts = run_product(g["variant"], cfg, sd, batch, masks, "fp32", g["train"])
    named = dict(m.named_parameters())
    names = list(g["opt_after1"].keys())
    opt = C.diffGradNorm([named[n] for n in names], lr=6e-5)
    opt.step()
    for n in names:
        assert relerr(samp(named[n]).cpu(), g["opt_after1"][n]) < 1e-5, n
    for n in names:
        named[n].grad = named[n].grad * 0.5 + 0.01
    opt.step()
    for n in names:
        assert relerr(samp(named[n]).cpu(), g["opt_after2"][n]) < 1e-5, n
    # packed weight copies must follow parameter updates: the raw-pointer step of our optimizer, an in-place
    # torch update, and (in grad mode) even a `.data` update that bumps no version counter
    x = batch["image"].cuda()
    m.eval()
    with torch.no_grad():
        a = m(x)["depth"]["final_depth"].clone()
        for n in names:
            named[n].grad = torch.ones_like(named[n]) * (1.0 if "norm" in n else 100.0)
        C.diffGradNorm([named[n] for n in names], lr=1e-2).step()
        b = m(x)["depth"]["final_depth"].clone()
        named[names[1]].mul_(1.5)
        c = m(x)["depth"]["final_depth"].clone()
    assert relerr(a, b) > 1e-4 and relerr(b, c) > 1e-4
    named[names[1]].data.mul_(1.5)
    d = m(x)["depth"]["final_depth"]
    assert relerr(c, d) > 1e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_baseline_size_vs_oracle(precision):
    """192x416 (the runnable stand-in for BASELINE's 192x400), base model, eval, B=1."""
    from oracle import camradepth_oracle as O
    from camradepth_b200.synthetic import make_batch
    torch.set_num_threads(os.cpu_count() or 8)
    cfg = O.Cfg("base")
    sd = O.init_state_dict(cfg, seed=5, perturb=0.02)
    batch = make_batch(1, 192, 416, seed=9)
    pred_o, loss_o, parts_o, grads_o, _ = oracle_run(cfg, sd, batch, (None, None))
    m, pred, loss, parts = run_product("base", cfg, sd, batch, (None, None), precision, False)
    tol = depth_tol(precision)
    e = relerr(pred["depth"]["final_depth"], pred_o["depth"]["final_depth"])
    assert e < tol, e
    assert abs(float(loss) - float(loss_o)) < tol * max(1.0, abs(float(loss_o)))
    num = den = dot = n1 = n2 = 0.0
    for n, p in m.named_parameters():
        a, b = p.grad.double().cpu().flatten(), grads_o[n].double().flatten()
        num += float((a - b).pow(2).sum()); den += float(b.pow(2).sum())
        dot += float(a @ b); n1 += float(a @ a); n2 += float(b @ b)
    grel, gcos = (num / den) ** 0.5, dot / (n1 * n2) ** 0.5
    with open(os.path.join(OUT, "parity_report.jsonl"), "a") as fh:
        fh.write(json.dumps({"case": "base_1x192x416_eval_vs_oracle", "precision": precision,
                             "final_depth_rel_l2": e, "global_grad_rel_l2": grel, "global_grad_cos": gcos}) + "\n")
    assert grel < (2e-3 if precision == "fp32" else 3e-2), grel
    assert gcos > (0.99999 if precision == "fp32" else 0.999), gcos


def test_properties_full_batch():
    """Size-independent properties at a larger batch: per-sample independence (GroupNorm has no cross-sample
    coupling) and eval determinism."""
    import camradepth_b200 as C
    from camradepth_b200.synthetic import make_batch
    C.set_model("base")
    torch.manual_seed(0)
    m = C.CamRaDepth(precision="bf16").cuda().eval()
    batch = make_batch(4, 192, 416, seed=2)
    x = batch["image"].cuda()
    with torch.no_grad():
        full = m(x)["depth"]["final_depth"]
        again = m(x)["depth"]["final_depth"]
        one = m(x[2:3])["depth"]["final_depth"]
    # GroupNorm statistics are reduced with fp32 atomics (order varies run to run); in bf16 a 1e-7 change of a
    # statistic can flip roundings downstream, so repeat runs agree to rounding noise, not bit-exactly
    assert relerr(full, again) < 5e-3
    assert relerr(full[2:3], one) < 5e-3
    assert full.shape == (4, 1, 192, 416)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 7, 192, 400, device="cuda"))          # reference also fails on 400-wide input (F2)


def test_deterministic_forward_mode():
    """deterministic=True: GroupNorm statistics from fixed-order reductions (no atomics) -> the forward pass is
    bit-identical run to run, and a sample's result does not depend on its batch mates (bitwise)."""
    import camradepth_b200 as C
    from camradepth_b200.synthetic import make_batch
    C.set_model("sup_unsup_seg")
    try:
        torch.manual_seed(0)
        m = C.CamRaDepth(precision="bf16", deterministic=True).cuda().eval()
        x = make_batch(3, 192, 416, seed=2)["image"].cuda()
        with torch.no_grad():
            a = m(x)
            b = m(x)
            one = m(x[1:2])
        for k in ("final_depth",):
            assert torch.equal(a["depth"][k], b["depth"][k])
            assert torch.equal(a["depth"][k][1:2], one["depth"][k])
        assert torch.equal(a["seg"]["final_seg"], b["seg"]["final_seg"])
        assert torch.equal(a["seg"]["unsup_map"], b["seg"]["unsup_map"])
        assert torch.equal(a["depth"]["intermediate_depths"][3][1:2], one["depth"]["intermediate_depths"][3])
        # and it is the same function as the default mode up to rounding noise
        m2 = C.CamRaDepth(precision="bf16").cuda().eval()
        m2.load_state_dict(m.state_dict())
        with torch.no_grad():
            c = m2(x)
        assert relerr(c["depth"]["final_depth"], a["depth"]["final_depth"]) < 2.5e-2     # argmax-fed heads, see golden_util
    finally:
        C.set_model("base")


def test_edge_cases_native_resolution_and_empty_mask():
    """Reference-native 416x800 input (args.py:19), odd batch, and the loss on an empty valid mask (NaN, like the
    reference's mean over an empty selection, loss_funcs.py:83-91)."""
    import camradepth_b200 as C
    from camradepth_b200.synthetic import make_batch
    C.set_model("base")
    torch.manual_seed(0)
    m = C.CamRaDepth(precision="bf16").cuda().eval()
    b = {k: v.cuda() for k, v in make_batch(3, 416, 800, seed=6).items()}
    pred = m(b["image"])
    assert pred["depth"]["final_depth"].shape == (3, 1, 416, 800)
    assert pred["depth"]["intermediate_depths"][2].shape == (3, 1, 104, 200)
    assert pred["depth"]["intermediate_depths"][3].shape == (3, 1, 208, 400)
    loss = C.MaskedSmoothL1Loss()(pred["depth"]["final_depth"], b["gt_final"])
    loss.backward()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(loss))
    for n, p in m.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
    empty = C.MaskedSmoothL1Loss()(pred["depth"]["final_depth"].detach(), torch.zeros_like(b["gt_final"]))
    assert bool(torch.isnan(empty))
    # eval-mode forward of one sample is independent of its batch mates at this size too
    with torch.no_grad():
        one = m(b["image"][1:2])["depth"]["final_depth"]
    assert relerr(pred["depth"]["final_depth"][1:2], one) < 5e-3
