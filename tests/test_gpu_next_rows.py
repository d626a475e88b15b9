"""test-mode metrics, input pipeline, training-step driver, checkpoint layout --
each against the reference semantics restated with plain torch ops / the CPU oracle."""
import os

import pytest
import torch
import torch.nn.functional as F

from tests.test_gpu_ops import rel

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def test_depth_metrics_match_reference_loop():
    """Restates runner.py:442-492 with torch ops on the same tensors."""
    from camradepth_b200 import depth_metrics
    torch.manual_seed(0)
    pred = torch.rand(1, 1, 64, 96, device=dev) * 1.2 - 0.1
    gt = torch.rand(1, 1, 64, 96, device=dev) * (torch.rand(1, 1, 64, 96, device=dev) < 0.3)
    m = depth_metrics(pred, gt)
    p = torch.clip(pred.squeeze(), 0, 1) * 100
    g = gt.squeeze() * 100
    g = g.clone(); g[g > 100] = 0
    idx = torch.where(g > 0)
    err = p[idx] - g[idx]
    ref = {"mae_100": err.abs().mean(), "rmse_100": err.pow(2).mean().sqrt(), "rel_100": (err.abs() / g[idx]).mean()}
    g[g < 50] = 0
    idx = torch.where(g > 0)
    err = p[idx] - g[idx]
    ref.update({"mae_50": err.abs().mean(), "rmse_50": err.pow(2).mean().sqrt(), "rel_50": (err.abs() / g[idx]).mean()})
    for k, v in ref.items():
        assert abs(float(m[k]) - float(v)) < 1e-4 * max(1.0, abs(float(v))), k
    # max_distances[0] below max_depth: pixels with gt > max_distances[0] are dropped from both sets (runner.py:455-457)
    m = depth_metrics(pred, gt, max_depth=100.0, max_distances=(80, 50))
    g = gt.squeeze() * 100
    g = g.clone(); g[g > 80] = 0
    idx = torch.where(g > 0)
    err = p[idx] - g[idx]
    ref = {"mae_80": err.abs().mean(), "rmse_80": err.pow(2).mean().sqrt(), "rel_80": (err.abs() / g[idx]).mean()}
    g[g < 50] = 0
    idx = torch.where(g > 0)
    err = p[idx] - g[idx]
    ref.update({"mae_50": err.abs().mean(), "rmse_50": err.pow(2).mean().sqrt(), "rel_50": (err.abs() / g[idx]).mean()})
    for k, v in ref.items():
        assert abs(float(m[k]) - float(v)) < 1e-4 * max(1.0, abs(float(v))), k


def test_mean_iou():
    from camradepth_b200.metrics import confusion_matrix, mean_iou
    torch.manual_seed(1)
    lg = torch.randn(2, 21, 32, 48, device=dev)
    t = torch.randint(0, 21, (2, 32, 48), device=dev)
    t[torch.rand(2, 32, 48, device=dev) < 0.1] = 255
    conf = confusion_matrix(lg, t)
    pr = lg.argmax(1)
    valid = t != 255
    ref = torch.zeros(21, 21, device=dev)
    ref.index_put_((t[valid], pr[valid]), torch.ones(int(valid.sum()), device=dev), accumulate=True)
    assert torch.equal(conf, ref)
    inter = ref.diag(); union = ref.sum(0) + ref.sum(1) - inter
    assert abs(float(mean_iou(lg, t)) - float((inter / union).mean())) < 1e-6


def test_input_pipeline_matches_dataloader_contract():
    from camradepth_b200.preprocess import normalize_image, gt_pyramid, IMAGENET_MEAN, IMAGENET_STD
    from oracle import camradepth_oracle as O
    torch.manual_seed(2)
    img = torch.randint(0, 256, (2, 64, 96, 3), dtype=torch.uint8, device=dev)
    x = torch.zeros(2, 7, 64, 96, device=dev)
    normalize_image(img, x)
    ref = (img.float() / 255 - torch.tensor(IMAGENET_MEAN, device=dev)) / torch.tensor(IMAGENET_STD, device=dev)
    assert rel(x[:, :3], ref.permute(0, 3, 1, 2)) < 1e-6 and float(x[:, 3:].abs().max()) == 0
    depth = torch.rand(2, 1, 64, 96, device=dev) * 130 * (torch.rand(2, 1, 64, 96, device=dev) < 0.2)
    g0, g1, g2 = gt_pyramid(depth)
    d = depth.cpu().clamp(0, 100)
    gn = torch.where(d > 0, (100 - d) / 100, torch.zeros_like(d))        # dataloader.py:240-245
    assert rel(g0.cpu(), gn) < 1e-6
    m1 = O.minpool(gn)                                                   # dataloader.py:213-222
    assert torch.allclose(g1.cpu(), m1) and torch.allclose(g2.cpu(), O.minpool(m1))


def test_train_step_driver_matches_reference_loop():
    """Three micro-batches with update_interval=2 and OneCycleLR, fp32 mode, against the oracle driven by the
    reference's loop logic (runner.py:166-270)."""
    import camradepth_b200 as C
    from camradepth_b200.synthetic import make_batch
    from oracle import camradepth_oracle as O
    torch.set_num_threads(os.cpu_count() or 8)
    cfg = O.Cfg("supervised_seg")
    sd = O.init_state_dict(cfg, seed=2, perturb=0.02)
    C.set_model("supervised_seg")
    model = C.CamRaDepth(precision="fp32")
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    UI, NB, LR = 2, 3, 1e-3
    opt = C.diffGradNorm(model.parameters(), lr=LR)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=LR, steps_per_epoch=NB, epochs=2, div_factor=2, pct_start=0.15)
    ts = C.TrainStep(model, opt, sched, update_interval=UI, supervised_seg=True, batches_per_epoch=NB)
    batches = [make_batch(1, 64, 64, seed=20 + i) for i in range(NB)]
    ts.start_epoch()
    stepped = []
    for b in batches:
        _, st = ts({k: v.to(dev) for k, v in b.items()})
        stepped.append(st)
    assert stepped == [False, True, True]            # boundary at i+1 == 2 and at the last batch of the epoch
    stats = ts.stats()
    # ---- reference loop on the oracle
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    dummy = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=LR)
    rs = torch.optim.lr_scheduler.OneCycleLR(dummy, max_lr=LR, steps_per_epoch=NB, epochs=2, div_factor=2, pct_start=0.15)
    states = {k: {} for k in p}
    finals = []
    for i, b in enumerate(batches):
        pred = O.forward(p, cfg, b["image"])
        loss, parts = O.training_loss(pred, b["gt_final"], b["gt_s4"], b["gt_s3"], b["gt_seg"], cfg, update_interval=UI)
        finals.append(float(parts["final"]))
        loss.backward()
        if (i + 1) % UI == 0 or (i + 1) == NB:
            lr = dummy.param_groups[0]["lr"]
            with torch.no_grad():
                for k, v in p.items():
                    if v.grad is not None:
                        O.diffgradnorm_step(v, v.grad, states[k], lr=lr)
                        v.grad = None
        if (i + 1) > UI:
            rs.step()
    assert abs(opt.param_groups[0]["lr"] - dummy.param_groups[0]["lr"]) < 1e-12
    # Biases that feed straight into a GroupNorm have an analytically zero gradient; Adam-style normalisation turns
    # the numerical noise of either implementation into +-lr updates of random sign, so they are not comparable.
    zero_grad_by_construction = (".proj.bias", ".attn.sr.bias", ".mlp1.fc1.bias", ".dwconv.dwconv.bias")
    # The same normalisation makes near-zero gradient entries (e.g. a class bias with |g| ~ 1e-8) move by a
    # noise-dependent fraction of one lr unit, so small tensors get an absolute bound (< 0.5 lr per entry) and
    # the weight tensors the relative one.
    # What the driver controls is the UPDATE: compare the parameter deltas of all weight tensors (a missing
    # accumulation step, a wrong lr or a wrong step boundary changes them by O(1)); entries whose gradient is
    # near zero or depends on an argmax near-tie (seg-map input channel) keep this from being 1e-5-tight.
    num = den = 0.0
    for n, q in model.named_parameters():
        if n.endswith(zero_grad_by_construction):
            continue
        a, b_, p0 = q.detach().cpu(), p[n].detach(), sd[n]
        assert float((a - b_).abs().max()) < 1.01 * LR, n
        if a.numel() >= 1024:
            num += float(((a - p0) - (b_ - p0)).double().pow(2).sum())
            den += float((b_ - p0).double().pow(2).sum())
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5
    assert abs(stats["loss_depth_final"] - sum(finals) / NB) < 1e-5
    C.set_model("base")


def test_checkpoint_layout_roundtrip(tmp_path):
    import camradepth_b200 as C
    from camradepth_b200.synthetic import make_batch
    C.set_model("base")
    torch.manual_seed(3)
    m = C.CamRaDepth(precision="bf16").to(dev).eval()
    opt = C.diffGradNorm(m.parameters(), lr=6e-5)
    b = {k: v.to(dev) for k, v in make_batch(1, 64, 64, seed=1).items()}
    C.MaskedSmoothL1Loss()(m(b["image"])["depth"]["final_depth"], b["gt_final"]).backward()
    opt.step()
    path = str(tmp_path / "ck.pth")
    C.save_checkpoint(path, torch.nn.DataParallel(m) if False else m, opt, 6e-5, 1)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert sorted(ck.keys()) == ["lr", "optimizer", "state_dict", "steps"]
    st0 = ck["optimizer"]["state"][0]
    assert sorted(st0.keys()) == ["exp_avg", "exp_avg_sq", "exp_grad_norm", "previous_grad", "step"]
    assert len(ck["optimizer"]["state"]) == 881 and st0["exp_avg"].shape == m.state_dict()["dest_encoder.patch_embed1.proj.weight"].shape
    m2 = C.CamRaDepth(precision="bf16")
    C.load_checkpoint_with_shape_match(m2, {"module." + k: v for k, v in ck["state_dict"].items()})
    assert all(torch.equal(a.cpu(), b2) for a, b2 in zip(m.state_dict().values(), m2.state_dict().values()))


def test_seg_label_resize_matches_skimage_rule():
    """dataloader.py:262-268: skimage.transform.resize(order=0, anti_aliasing=False) = scipy.ndimage.zoom(order=0,
    grid_mode=True) (skimage is not in this image; scipy is what it calls)."""
    import numpy as np
    from scipy import ndimage
    from camradepth_b200.preprocess import seg_resize_nearest
    rng = np.random.default_rng(0)
    for (hi, wi, ho, wo) in [(416, 800, 208, 400), (416, 800, 416, 800), (450, 800, 192, 416), (37, 53, 64, 96)]:
        lab = rng.integers(0, 21, size=(2, hi, wi)).astype(np.uint8)
        lab[rng.random(lab.shape) < 0.1] = 255
        ref = np.stack([ndimage.zoom(l, (ho / hi, wo / wi), order=0, mode="reflect", grid_mode=True) for l in lab])
        for t in (torch.from_numpy(lab), torch.from_numpy(lab.astype(np.int64))):
            out = seg_resize_nearest(t.to(dev), (ho, wo))
            assert out.dtype == torch.int64 and tuple(out.shape) == (2, ho, wo)
            assert np.array_equal(out.cpu().numpy(), ref.astype(np.int64)), (hi, wi, ho, wo)


def test_packed_input_feed_matches_module_surface():
    """GPU input pipeline feeding the engine layout directly: forward_packed(pack_input_nhwc(...))
    == forward(NCHW fp32 tensor built by normalize_image + radar planes)."""
    import camradepth_b200 as C
    from camradepth_b200.preprocess import normalize_image, pack_input_nhwc
    C.set_model("base")
    torch.manual_seed(0)
    m = C.CamRaDepth(precision="bf16").cuda().eval()
    B, H, W = 2, 64, 96
    img = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    radar = torch.rand(B, 4, H, W, device=dev) * (torch.rand(B, 1, H, W, device=dev) < 0.05)
    x = torch.zeros(B, 7, H, W, device=dev)
    normalize_image(img, x)
    x[:, 3:] = radar
    xp = pack_input_nhwc(img, radar.contiguous())
    assert xp.shape == (B, H, W, 8) and xp.dtype == torch.bfloat16
    assert rel(xp[..., :7].float().permute(0, 3, 1, 2), x) < 4e-3 and float(xp[..., 7].abs().max()) == 0
    with torch.no_grad():
        a = m(x)["depth"]["final_depth"]
        b = m.forward_packed(xp)["depth"]["final_depth"]
    assert rel(b, a) < 5e-3          # same bf16 input values; GroupNorm sums are atomics-ordered
    # gradients flow through the packed entry too
    m.train(False)
    loss = C.MaskedSmoothL1Loss()(m.forward_packed(xp)["depth"]["final_depth"], torch.rand(B, 1, H, W, device=dev))
    loss.backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
    with pytest.raises(RuntimeError):
        m.forward_packed(xp.float())


def test_stochastic_masks_kernel():
    """One Philox launch for all DropPath / Dropout2d scales: values in {0, 1/keep}, keep rates honoured, rate-0
    calls are identity, and the device-side step counter gives fresh masks per call (CUDA-graph replays included)."""
    from camradepth_b200 import ops
    n_dp, B, n_d2, C2 = 68, 64, 7, 128
    rates = torch.linspace(0, 0.1, n_dp // 2).repeat_interleave(2)
    keep = (1 - rates).to(dev)
    out = torch.empty(n_dp * B + n_d2 * B * C2, device=dev)
    state = torch.tensor([1234, 0], dtype=torch.int64, device=dev)
    draws = []
    for it in range(200):
        ops.make_masks(out, keep, n_dp, B, n_d2, C2, 0.8, state)
        draws.append(out.clone())
    torch.cuda.synchronize()
    assert int(state[1]) == 200
    d = torch.stack(draws)
    dp = d[:, :n_dp * B].view(200, n_dp, B)
    d2 = d[:, n_dp * B:]
    assert bool((dp[:, :2] == 1).all())                                   # rate 0: never dropped
    for r in (10, 40, 67):
        k = float(keep[r])
        vals = dp[:, r].unique()
        assert all(abs(float(v)) < 1e-6 or abs(float(v) - 1 / k) < 1e-5 for v in vals)
        frac = float((dp[:, r] > 0).float().mean())
        assert abs(frac - k) < 0.02, (r, frac, k)
    assert abs(float((d2 > 0).float().mean()) - 0.8) < 5e-3
    assert set(d2.unique().tolist()) <= {0.0, 1.25}
    assert not torch.equal(draws[0], draws[1])
    # same seed and counter -> same masks (reproducible)
    state2 = torch.tensor([1234, 0], dtype=torch.int64, device=dev)
    out2 = torch.empty_like(out)
    ops.make_masks(out2, keep, n_dp, B, n_d2, C2, 0.8, state2)
    assert torch.equal(out2, draws[0])


def test_eval_loop_matches_reference_semantics():
    """`evaluate` against the reference's eval loop (runner.py:273-350) restated on the CPU oracle, with the F12 fix
    (input sliced with args.input_channels): five batches, update_interval 2, fp32 mode."""
    import numpy as np
    import camradepth_b200 as C
    from camradepth_b200.synthetic import make_batch
    from oracle import camradepth_oracle as O
    torch.set_num_threads(os.cpu_count() or 8)
    cfg = O.Cfg("supervised_seg")
    sd = O.init_state_dict(cfg, seed=3, perturb=0.02)
    C.set_model("supervised_seg")
    try:
        model = C.CamRaDepth(precision="fp32")
        model.load_state_dict(sd)
        model = model.to(dev).train()             # evaluate() must switch to eval and restore the mode
        batches = [make_batch(1, 64, 64, seed=40 + i) for i in range(5)]
        batches[3]["gt_final"] = torch.zeros_like(batches[3]["gt_final"])       # empty mask -> NaN losses (nanmean)
        val_loss, rmse = C.evaluate(model, [{k: v.to(dev) for k, v in b.items()} for b in batches], update_interval=2)
        assert model.training
        # reference loop
        eval_losses, rmse_arr, fin, s4, seg = [], [], [], [], []
        with torch.no_grad():
            for i, b in enumerate(batches):
                pred = O.forward(sd, cfg, b["image"][:, :cfg.cin])
                seg.append(float(O.masked_focal(pred["seg"]["final_seg"], b["gt_seg"])))
                s4.append(float(O.masked_smooth_l1(pred["depth"]["intermediate_depths"][-1].squeeze(1), b["gt_s4"].squeeze(1))))
                fin.append(float(O.masked_smooth_l1(pred["depth"]["final_depth"], b["gt_final"])))
                rmse_arr.append(float(torch.sqrt(O.masked_mse(pred["depth"]["final_depth"], b["gt_final"]))) * 100)
                if (i + 1) % 2 == 0 or (i + 1) == len(batches):
                    eval_losses.append([np.nanmean(fin), np.nanmean(s4), np.nanmean(rmse_arr[-600:]), np.nanmean(seg)])
                    fin, s4, seg = [], [], []
        eval_losses = np.array(eval_losses)
        want_loss, want_rmse = np.nanmean(eval_losses[:, 0]), np.nanmean(eval_losses[:, 2])
        assert abs(val_loss - want_loss) < 1e-4 * abs(want_loss), (val_loss, want_loss)
        assert abs(rmse - want_rmse) < 1e-4 * abs(want_rmse), (rmse, want_rmse)
    finally:
        C.set_model("base")
