"""Pins the CPU oracle (oracle/camradepth_oracle.py) against outputs of the REAL reference
(tests/golden/*.pt, produced by oracle/make_golden.py in the build container)."""
import pytest
import torch

from oracle import camradepth_oracle as O
from tests.golden_util import golden_files, load_case, oracle_run, relerr, samp

FILES = golden_files()


def test_fixtures_present():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1])
def test_oracle_matches_reference(path):
    torch.set_num_threads(8)
    g, cfg, sd, batch, masks = load_case(path)
    pred, loss, parts, grads, sd_g = oracle_run(cfg, sd, batch, masks)
    assert relerr(pred["depth"]["final_depth"], g["final_depth"]) < 2e-5
    assert relerr(pred["depth"]["intermediate_depths"][2], g["inter3"]) < 2e-5
    assert relerr(pred["depth"]["intermediate_depths"][3], g["inter4"]) < 2e-5
    if g["final_seg_sample"] is not None:
        assert relerr(pred["seg"]["final_seg"][:, :, ::4, ::4], g["final_seg_sample"]) < 2e-5
    if g["unsup_map"] is not None:
        # argmax maps: allow a handful of near-tie flips
        neq = (pred["seg"]["unsup_map"].float() != g["unsup_map"]).float().mean()
        assert float(neq) < 1e-3
    assert abs(float(loss) - float(g["losses"][4])) < 1e-5 * max(1.0, abs(float(g["losses"][4])))
    # which params get no gradient
    none = sorted(k for k, v in grads.items() if v is None)
    assert none == sorted(g["none_grads"])
    names = g["grad_stats"]["names"]
    ref = g["grad_stats"]["sum_norm"]
    worst = 0.0
    for i, n in enumerate(names):
        gn = float(grads[n].double().norm())
        worst = max(worst, abs(gn - float(ref[i, 1])) / (float(ref[i, 1]) + 1e-12))
    assert worst < 5e-3, worst          # norms incl. argmax-routed attn tensors
    for n, gs in g["full_grads"].items():
        assert relerr(samp(grads[n]), gs) < 5e-3, n


@pytest.mark.parametrize("path", FILES[:1], ids=lambda p: p.split("/")[-1])
def test_oracle_optimizer_matches_reference(path):
    g, cfg, sd, batch, masks = load_case(path)
    pred, loss, parts, grads, sd_g = oracle_run(cfg, sd, batch, masks)
    for n in g["opt_after1"]:
        p = sd[n].clone()
        st = {}
        gr = grads[n].clone()
        O.diffgradnorm_step(p, gr, st)
        assert relerr(samp(p), g["opt_after1"][n]) < 1e-6
        O.diffgradnorm_step(p, gr * 0.5 + 0.01, st)
        assert relerr(samp(p), g["opt_after2"][n]) < 1e-6
        # the step must have moved the parameter
        assert relerr(samp(p), samp(sd[n])) > 1e-7
