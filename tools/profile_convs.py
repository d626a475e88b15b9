"""Per-layer CUDA-event timing of every convolution launch (fwd / dgrad / wgrad) in one training step."""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import camradepth_b200 as C  # noqa: E402
from camradepth_b200.synthetic import make_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
C.set_model("base")
model = C.CamRaDepth(precision="bf16").to(dev).train()
crit = C.MaskedSmoothL1Loss()
b = {k: v.to(dev) for k, v in make_batch(B, 192, 416, seed=1).items()}


def step():
    pred = model(b["image"])
    inter = pred["depth"]["intermediate_depths"]
    loss = crit(pred["depth"]["final_depth"], b["gt_final"]) + crit(inter[-1], b["gt_s4"]) + crit(inter[-2], b["gt_s3"])
    loss.backward()
    model.zero_grad(set_to_none=True)


for _ in range(3):
    step()
eng = model._engines["bf16"]
eng.timed = {(k, n): [] for n in eng.L for k in ("fwd", "dgrad", "wgrad")}
step()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
for (kind, name), evs in eng.timed.items():
    L = eng.L[name]
    g = re.sub(r"\.\d+\.", ".N.", name)
    g = re.sub(r"block\d", "blockS", g) if False else g
    for e0, e1 in evs:
        ms = e0.elapsed_time(e1)
        agg[(kind, g)][0] += ms
        agg[(kind, g)][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"B={B}: conv launches total {tot:.2f} ms")
for (kind, g), (ms, n, _) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f"{ms:8.3f} ms {n:4d}  {kind:6s} {g}")

# decoder layers one by one, with the useful tensor throughput (2 * pixels * Cout * Cin * taps per launch)
print("\ndecoder convolutions, per layer:")
for (kind, name), evs in sorted(eng.timed.items(), key=lambda kv: kv[0][1]):
    if "depth_upsample" not in name or not evs:
        continue
    L = eng.L[name]
    ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    st = int(name.split(".")[1])
    px = B * (12 * 26) * 4 ** st
    fl = 2.0 * px * L["cout"] * L["cin"] * L["taps"]
    note = ""
    if (kind, name) in eng.timed_flops:
        # stacked-K data gradients: the timed launch computes one output-channel block of the concat gradient from
        # the dy of SEVERAL layers; credit the MACs that launch really does (summing over the three launches of a
        # dense block gives the same total as the three per-layer dgrads)
        fl = eng.timed_flops[(kind, name)] * len(evs)
        note = "  (output-channel block, stacked K)"
    else:
        fl *= len(evs)
    print(f"{ms:8.3f} ms x{len(evs)}  {kind:6s} {name:52s} Cin {L['cin']:4d} Cout {L['cout']:4d}  {fl / ms / 1e9:7.0f} TF/s{note}")
