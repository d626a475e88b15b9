"""Where does the data-parallel step lose time against N independent single-GPU steps?

Run under torchrun (one rank per GPU).  The same batch-32 base training step is timed (CUDA events, max over ranks)
with the exchanges switched off one at a time (graphs.GraphedDataParallelStep, CAMRADEPTH_DP_DEBUG):

    onegraph   one CUDA graph, no exchange at all: N processes that merely run at the same time on one box
    graphs     seven graphs, no exchange: the cost of cutting the step at the bucket boundaries
    buckets    + the gradient all-reduce of every bucket (no loss exchange)
    full       + the all-reduce of the loss accumulators between the forward and the loss graph (the shipped path)

    torchrun --nproc-per-node 2 tools/dp_overhead.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import camradepth_b200 as C  # noqa: E402
from camradepth_b200.graphs import GraphedDataParallelStep  # noqa: E402
from camradepth_b200.parallel import DataParallel  # noqa: E402
from camradepth_b200.synthetic import make_batch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    C.set_model("base")
    torch.manual_seed(0)
    model = C.CamRaDepth(input_channels=C.args.input_channels, precision="bf16").to(dev).train()
    net = DataParallel(model)
    batch = {k: v.to(dev) for k, v in make_batch(32, 192, 416, seed=rank, input_channels=C.args.input_channels).items()}
    variants = [("onegraph", "onegraph,noloss,nobucket"), ("graphs", "noloss,nobucket"), ("buckets", "noloss"),
                ("full", ""), ("onegraph", "onegraph,noloss,nobucket"), ("full", "")]
    if os.environ.get("DPO_CUTS"):      # e.g. DPO_CUTS="stage3,stage2,stage1,stage0;stage1;stage2;": one full run per setting
        variants = [("cuts=" + c, c) for c in os.environ["DPO_CUTS"].split(";")] * 2
    for name, dbg in variants:
        if name.startswith("cuts="):
            os.environ["CAMRADEPTH_DP_CUTS"], dbg = dbg, ""
        os.environ["CAMRADEPTH_DP_DEBUG"] = dbg
        opt = C.diffGradNorm(model.parameters(), lr=1e-4)
        g = GraphedDataParallelStep(net, opt, batch)
        for _ in range(5):
            g(batch)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            g(batch)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        lo = ms.clone()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{name:9s} world {world}: {float(ms):7.3f} ms/step (slowest rank), {float(lo):7.3f} (fastest)", flush=True)
        del g, opt
        torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
