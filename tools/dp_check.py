"""Data-parallel parity (run under torchrun, one rank per GPU):

  1. eager path: gradients after the bucketed NCCL all-reduce on N ranks (per-rank batch b, losses normalised as
     GLOBAL masked means) == single-process gradients of the concatenated batch of N*b samples -- what the
     reference's nn.DataParallel computes on its gathered outputs (runner.py:135-136,193-203);
  2. replicas stay bit-identical through an optimizer step;
  3. graphed path (graphs.GraphedDataParallelStep: one CUDA graph per gradient bucket, all-reduce of a bucket
     overlapped with the backward of the next one): loss and parameters after several steps == a single process
     stepping on the concatenated batch.

    torchrun --nproc-per-node 2 tools/dp_check.py [variant] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import camradepth_b200 as C  # noqa: E402
from camradepth_b200 import losses  # noqa: E402
from camradepth_b200.graphs import GraphedDataParallelStep  # noqa: E402
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

    def build():
        return C.CamRaDepth(input_channels=C.args.input_channels, precision=precision).to(dev).eval()

    torch.manual_seed(123 + rank)            # deliberately different init per rank: the wrapper must broadcast
    model = build()
    net = DataParallel(model)                # global masked-mean losses from here on
    # shards with deliberately different valid-pixel counts (the case where per-rank means differ from the global one)
    shards = []
    for r in range(world):
        s = make_batch(b, 64, 96, seed=50 + r, input_channels=C.args.input_channels)
        keep = (torch.rand(s["gt_final"].shape, generator=torch.Generator().manual_seed(r)) < (0.3 + 0.6 * r / max(1, world - 1)))
        s["gt_final"] = s["gt_final"] * keep
        shards.append({k: v.to(dev) for k, v in s.items()})
    full = {k: torch.cat([s[k] for s in shards]) for k in shards[0]}
    dp_loss = loss_of(net, shards[rank])
    dp_loss.backward()
    torch.cuda.synchronize()
    dp_grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    # reference on this rank alone: the concatenated batch through an unwrapped copy of the (broadcast) parameters
    ref = build()
    ref.load_state_dict(model.state_dict())
    losses.set_data_parallel(1)
    ref_loss = loss_of(ref, full)
    ref_loss.backward()
    ref_grads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    losses.set_data_parallel(world)
    num = sum(float((dp_grads[n] - ref_grads[n]).double().pow(2).sum()) for n in ref_grads)
    den = sum(float(ref_grads[n].double().pow(2).sum()) for n in ref_grads)
    rel = (num / den) ** 0.5
    assert set(dp_grads) == set(ref_grads)
    dl = abs(float(dp_loss) - float(ref_loss)) / abs(float(ref_loss))
    # replicas stay in sync through an optimizer step
    opt = C.diffGradNorm(model.parameters(), lr=1e-3)
    opt.step()

    def checksum(m):
        flat = torch.cat([p.detach().flatten() for p in m.parameters()])
        chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        return all(torch.equal(gathered[0], g) for g in gathered)

    same = checksum(model)
    tol = 1e-5 if precision == "fp32" else 2e-2
    if rank == 0:
        print(f"dp_check {variant} {precision} world={world}: loss rel diff vs single-process batch {dl:.2e}, grad rel-L2 "
              f"{rel:.3e} (tol {tol}), replicas identical after step: {same}, tensors with grad: {len(ref_grads)}")
    assert rel < tol and same and dl < tol

    # ---- graphed, bucket-overlapped step vs a single process on the concatenated batch
    STEPS = 4
    torch.manual_seed(7)
    m_dp = build()
    net2 = DataParallel(m_dp)
    o_dp = C.diffGradNorm(m_dp.parameters(), lr=1e-3)
    g = GraphedDataParallelStep(net2, o_dp, shards[rank], warmup=2)
    got = [float(g(shards[rank])) for _ in range(STEPS - 2)]
    torch.cuda.synchronize()
    # single-process reference from the same initial parameters (every rank seeds 7, so rank 0's == everyone's)
    torch.manual_seed(7)
    m_sp = build()
    o_sp = C.diffGradNorm(m_sp.parameters(), lr=1e-3)
    losses.set_data_parallel(1)
    want = []
    for _ in range(STEPS):
        l = loss_of(m_sp, full)
        l.backward()
        o_sp.step()
        o_sp.zero_grad(set_to_none=True)
        want.append(float(l))
    losses.set_data_parallel(world)
    p_dp = torch.cat([p.detach().flatten() for p in m_dp.parameters()])
    p_sp = torch.cat([p.detach().flatten() for p in m_sp.parameters()])
    prel = float((p_dp - p_sp).double().norm() / p_sp.double().norm())
    lrel = max(abs(a - c) / abs(a) for a, c in zip(want[2:], got))
    same2 = checksum(m_dp)
    if rank == 0:
        print(f"dp_check graphed step: {len(g.graphs)} graphs/step, loss rel diff {lrel:.2e}, params rel-L2 vs single process "
              f"after {STEPS} steps {prel:.3e}, replicas identical: {same2}")
    ptol = 1e-5 if precision == "fp32" else 5e-3
    assert lrel < 10 * tol and prel < ptol and same2
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
