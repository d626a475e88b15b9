"""CPU-side checks: the C-ABI library builds, loads, exports every symbol include/*.h declares; the product
fails loudly without CUDA; schema / host logic matches the oracle."""
import ctypes
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_all_symbols():
    from camradepth_b200.build import build
    from camradepth_b200 import _lib
    path = build()
    assert os.path.exists(path)
    protos = _lib.parse_header()
    assert len(protos) >= 40
    lib = ctypes.CDLL(path)
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/camradepth_b200.h but not exported"
    assert _lib.load().crd_version() >= 1


def test_conv_desc_layout_matches_header():
    from camradepth_b200 import _lib
    src = open(_lib.HEADER).read()
    body = src[src.index("typedef struct {"):src.index("} crd_conv_desc;")]
    import re
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for line in body.splitlines():
        line = line.strip()
        if line.startswith("int "):
            names += [n.strip() for n in line[4:].rstrip(";").split(",")]
    assert names == [f[0] for f in _lib.ConvDesc._fields_]


def test_no_cpu_fallback():
    import camradepth_b200 as C
    C.set_model("base")
    m = C.CamRaDepth()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 7, 64, 64))
    with pytest.raises(RuntimeError):
        C.MaskedSmoothL1Loss()(torch.zeros(1, 1, 4, 4), torch.ones(1, 1, 4, 4))
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        C.diffGradNorm([p]).step()


def test_product_never_imports_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "camradepth_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle." not in src, f


@pytest.mark.parametrize("variant", ["base", "base (rgb)", "supervised_seg", "unsupervised_seg", "sup_unsup_seg"])
def test_schema_matches_oracle(variant):
    import camradepth_b200 as C
    from oracle import camradepth_oracle as O
    C.set_model(variant)
    m = C.CamRaDepth(input_channels=C.args.input_channels)
    spec = O.param_spec(O.Cfg(variant))
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == spec[k][0]
    counts = {"base": 21966595, "supervised_seg": 23182861, "unsupervised_seg": 23178249,
              "sup_unsup_seg": 23227251, "base (rgb)": 21943683}
    assert sum(p.numel() for p in m.parameters()) == counts[variant]
    # DataParallel-style 'module.' checkpoints round-trip through the shape-matching loader (utils.py:352-370)
    ck = {"module." + k: torch.full_like(v, 0.5) for k, v in sd.items()}
    ck["module.depth_activation_4.conv_1.weight"] = torch.zeros(32, 999, 3, 3)
    C.load_checkpoint_with_shape_match(m, ck)
    assert float(m.state_dict()["dest_encoder.block1.0.attn.q.weight"].mean()) == 0.5
    assert m.state_dict()["depth_activation_4.conv_1.weight"].shape[1] != 999
    C.set_model("base")


def test_engine_layer_table_channel_maps():
    import camradepth_b200 as C
    from camradepth_b200.engine import Engine
    C.set_model("sup_unsup_seg")
    m = C.CamRaDepth()
    e = Engine(m, m.cfg)
    L = e.L["depth_upsample.3.conv.layers.1.model.0.weight"]
    assert L["cin"] == 225 and L["cin_p"] == 232 and L["cmap"][128] == 128 and L["cmap"][129] == 136
    L = e.L["depth_upsample.4.conv.layers.2.model.0.weight"]
    assert L["cin"] == 296 and L["cmap"] is None and L["cin_p"] == 296
    L = e.L["depth_activation_5.conv_1.weight"]
    assert L["cin"] == 130 and L["cin_p"] == 136 and L["cmap"][128:] == [129, 130]
    assert e.no_grad_names(True) == ["seg_conv_stage_4.weight", "seg_conv_stage_4.bias", "unsup_stage_4.weight",
                                     "unsup_stage_4.bias", "unsup_final.weight", "unsup_final.bias"]
    C.set_model("base")


def test_optimizer_tables():
    from camradepth_b200.optim import diffGradNorm
    from camradepth_b200.ops import OPT_CHUNK
    table, ck = diffGradNorm.build_tables([(1, 2, 3, 4, 5), (6, 7, 8, 9, 10)], [OPT_CHUNK * 2 + 5, 3], "cpu")
    assert table.shape == (2, 6) and table[0, 5] == OPT_CHUNK * 2 + 5
    assert ck.tolist() == [[0, 0], [0, OPT_CHUNK], [0, 2 * OPT_CHUNK], [1, 0]]


def test_synthetic_batch_contract():
    from camradepth_b200.synthetic import make_batch, minpool_gt
    from oracle import camradepth_oracle as O
    b = make_batch(2, 64, 96, seed=0)
    assert b["image"].shape == (2, 7, 64, 96) and b["gt_s4"].shape == (2, 1, 32, 48) and b["gt_s3"].shape == (2, 1, 16, 24)
    assert b["gt_seg"].dtype == torch.int64 and int(b["gt_seg"].max()) == 255
    assert torch.equal(minpool_gt(b["gt_final"]), O.minpool(b["gt_final"]))
    b2 = make_batch(2, 64, 96, seed=0)
    assert torch.equal(b["image"], b2["image"])


def test_weight_pack_brick_count_is_a_pure_host_function():
    """crd_weight_pack_blocks (how many blocks one item of a crd_weight_pack_batch table owns) needs no GPU: one brick
    per (<= 64 output channels) x (<= 64 input channels) x all taps with at most 8192 elements."""
    from camradepth_b200._lib import load
    lib = load()
    f = lib.crd_weight_pack_blocks
    assert f(128, 296, 9) == ((128 + 13) // 14) * 5          # 64 ci x 9 taps = 576 per row -> 14 output channels per brick
    assert f(1024, 128, 1) == 16 * 2                          # 1x1: 64 x 64 bricks
    assert f(64, 7, 49) == 3                                  # 7 ci x 49 taps = 343 per row -> 23 output channels
    assert f(1, 32, 9) == 1
    assert f(0, 32, 9) < 0 and f(8, 8, 129) < 0               # rejected: empty / more taps than a brick row can hold
