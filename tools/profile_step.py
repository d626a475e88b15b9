"""Per-kernel device-time breakdown of one training step (torch.profiler / CUPTI, no replay).
   python tools/profile_step.py [--batch 32] [--variant base] > gpurun_out/step_profile.txt"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

import camradepth_b200 as C  # noqa: E402
from camradepth_b200.synthetic import make_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--variant", default="base")
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    C.set_model(a.variant)
    torch.manual_seed(0)
    model = C.CamRaDepth(precision="bf16").to(dev).train()
    crit_d, crit_s = C.MaskedSmoothL1Loss(), C.MaskedFocalLoss()
    opt = C.diffGradNorm(model.parameters(), lr=6e-5)
    b = {k: v.to(dev) for k, v in make_batch(a.batch, 192, 416, seed=1, input_channels=C.args.input_channels).items()}

    def step():
        pred = model(b["image"])
        inter = pred["depth"]["intermediate_depths"]
        loss = crit_d(pred["depth"]["final_depth"], b["gt_final"]) + crit_d(inter[-1], b["gt_s4"]) + crit_d(inter[-2], b["gt_s3"])
        fs = pred["seg"]["final_seg"]
        if fs is not None:
            loss = loss + 0.2 * crit_s(fs, b["gt_seg"])
        (loss / 3.4).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    step()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_total = time.perf_counter() - t0
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    agg, cnt = collections.Counter(), collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:90]
            agg[name] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            cnt[name] += 1
    tot = sum(agg.values())
    print(f"batch {a.batch} {a.variant}: host enqueue {t_host*1e3:.1f} ms/step, wall {t_total*1e3:.1f} ms/step; "
          f"device busy {tot/1e3/a.steps:.1f} ms/step over {sum(cnt.values())//a.steps} kernels/step")
    for n, v in agg.most_common(45):
        print("%6.2f%% %9.3f ms/step %6d  %s" % (100 * v / tot, v / 1e3 / a.steps, cnt[n] // a.steps, n))


if __name__ == "__main__":
    main()
