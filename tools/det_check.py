"""Finds the first tensor of the forward program that differs between two runs on identical inputs
(deterministic mode should give none).   python tools/det_check.py [variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import camradepth_b200 as C  # noqa: E402
from camradepth_b200.synthetic import make_batch  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "base"
C.set_model(variant)
torch.manual_seed(0)
m = C.CamRaDepth(precision="bf16", deterministic=True).cuda().eval()
x = make_batch(2, 192, 416, seed=2, input_channels=C.args.input_channels)["image"].cuda()
eng = m._engine_for(x)


def flat(prefix, obj, out):
    if torch.is_tensor(obj):
        out.append((prefix, obj))
    elif isinstance(obj, dict):
        for k, v in obj.items():
            flat(f"{prefix}.{k}", v, out)
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            flat(f"{prefix}[{i}]", v, out)


def run():
    outs, S = eng.forward(x, False, None, save=True)
    torch.cuda.synchronize()
    lst = []
    flat("S", S, lst)
    flat("out", outs, lst)
    return [(n, t.clone()) for n, t in lst]


a = run()
b = run()
bad = 0
for (n1, t1), (n2, t2) in zip(a, b):
    assert n1 == n2
    if t1.shape != t2.shape or not torch.equal(t1, t2):
        d = (t1.float() - t2.float()).abs().max().item() if t1.shape == t2.shape else float("nan")
        print("DIFF", n1, tuple(t1.shape), t1.dtype, "max abs diff", d)
        bad += 1
        if bad > 12:
            break
print("tensors compared:", len(a), "differing:", bad)
