"""Data-parallel parity (run under torchrun, one rank per GPU):
grads after the bucketed NCCL all-reduce on N ranks (per-rank batch b) == mean over ranks of the single-GPU
grads of each rank's shard (exact up to summation order).  Also checks replicas stay bit-identical after
an optimizer step.   torchrun --nproc-per-node 2 tools/dp_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import camradepth_b200 as C  # noqa: E402
from camradepth_b200.parallel import DataParallel  # noqa: E402
from camradepth_b200.synthetic import make_batch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    variant = sys.argv[1] if len(sys.argv) > 1 else "supervised_seg"
    precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    C.set_model(variant)
    b = 2
    crit_d, crit_s = C.MaskedSmoothL1Loss(), C.MaskedFocalLoss()

    def loss_of(model, batch):
        pred = model(batch["image"])
        inter = pred["depth"]["intermediate_depths"]
        loss = crit_d(pred["depth"]["final_depth"], batch["gt_final"]) + crit_d(inter[-1], batch["gt_s4"]) + \
            crit_d(inter[-2], batch["gt_s3"])
        if pred["seg"]["final_seg"] is not None:
            loss = loss + 0.2 * crit_s(pred["seg"]["final_seg"], batch["gt_seg"])
        return loss / 3.4

    torch.manual_seed(123 + rank)            # deliberately different init per rank: the wrapper must broadcast
    model = C.CamRaDepth(input_channels=C.args.input_channels, precision=precision).to(dev).eval()
    net = DataParallel(model)
    shards = [{k: v.to(dev) for k, v in make_batch(b, 64, 96, seed=50 + r, input_channels=C.args.input_channels).items()}
              for r in range(world)]
    loss_of(net, shards[rank]).backward()
    torch.cuda.synchronize()
    dp_grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    # reference on this rank alone: every shard through an unwrapped copy of the (broadcast) parameters
    ref = C.CamRaDepth(input_channels=C.args.input_channels, precision=precision).to(dev).eval()
    ref.load_state_dict(model.state_dict())
    acc = {}
    for r in range(world):
        ref.zero_grad(set_to_none=True)
        loss_of(ref, shards[r]).backward()
        for n, p in ref.named_parameters():
            if p.grad is not None:
                acc[n] = acc.get(n, 0) + p.grad / world
    num = sum(float((dp_grads[n] - acc[n]).double().pow(2).sum()) for n in acc)
    den = sum(float(acc[n].double().pow(2).sum()) for n in acc)
    rel = (num / den) ** 0.5
    assert set(dp_grads) == set(acc)
    # replicas stay in sync through an optimizer step
    opt = C.diffGradNorm(model.parameters(), lr=1e-3)
    opt.step()
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    tol = 1e-5 if precision == "fp32" else 2e-2
    if rank == 0:
        print(f"dp_check {variant} {precision} world={world}: grad rel-L2 vs mean-of-shards {rel:.3e} (tol {tol}), "
              f"replicas identical after step: {same}, tensors with grad: {len(acc)}")
    assert rel < tol and same
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
