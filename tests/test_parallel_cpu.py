"""Host-side logic of the data-parallel path on CPU: bucket planning and a world_size-2 gloo run of the
bucketed all-reduce sequence the backward program issues."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _spec_offsets(variant):
    import camradepth_b200 as C
    from camradepth_b200.spec import param_spec
    C.set_model(variant)
    m = C.CamRaDepth(input_channels=C.args.input_channels)
    names, offsets, off = [], {}, 0
    for n, p in m.named_parameters():
        names.append(n)
        offsets[n] = (off, p.numel())
        off += p.numel()
    C.set_model("base")
    return names, offsets, off


@pytest.mark.parametrize("variant", ["base", "sup_unsup_seg"])
def test_bucket_ranges_cover_every_parameter_once(variant):
    from camradepth_b200.parallel import bucket_ranges
    names, offsets, total = _spec_offsets(variant)
    covered = torch.zeros(total, dtype=torch.int32)
    order = ["decoder", "stage3", "stage2", "stage1", "stage0"]      # order in which backward reports them
    for tag in order:
        for a, b in bucket_ranges(names, offsets, tag):
            assert 0 <= a < b <= total
            covered[a:b] += 1
    assert int(covered.min()) == 1 and int(covered.max()) == 1
    # the decoder bucket is everything registered after the encoder, i.e. a suffix of the flat buffer
    (a, b), = bucket_ranges(names, offsets, "decoder")
    assert b == total and names[[i for i, n in enumerate(names) if offsets[n][0] == a][0]].startswith("from_encoder_1")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from camradepth_b200.parallel import DataParallel, bucket_ranges

    class FakeEngine:
        pass

    names, offsets, total = _spec_offsets("base")
    eng = FakeEngine()
    eng.names, eng.pg_offsets = names, offsets
    eng.flat_grad = torch.full((total,), float(rank + 1))
    net = torch.nn.Linear(2, 2)                      # stands in for the module (parameters get broadcast)
    with torch.no_grad():
        net.weight.fill_(float(rank))
    dp = DataParallel(net)
    assert float(net.weight.sum()) == 0.0            # rank 0's parameters everywhere
    for tag in ["heads", "decoder", "stage3", "stage2", "stage1", "stage0"]:
        dp._on_bucket(eng, tag)
    dp._finish(eng, None)
    ok = bool(torch.allclose(eng.flat_grad, torch.full((total,), (1 + world) / 2.0)))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _loss_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from camradepth_b200 import losses
    # per-rank partial sums of a masked-mean loss: [sum of per-pixel losses, valid count, sum of squares]
    g = torch.Generator().manual_seed(5)
    per_pixel = [torch.rand(1000, generator=g), 2.0 * torch.rand(300, generator=g)]   # rank 1: fewer valid pixels, larger losses
    acc = torch.tensor([float(per_pixel[rank].sum()), float(per_pixel[rank].numel()), 0.0])
    losses.set_data_parallel(world)
    scale = losses._global_sums(acc)
    losses.set_data_parallel(1)
    want = torch.cat(per_pixel)
    ok = abs(float(acc[0] / acc[1]) - float(want.double().mean())) < 1e-5 and scale == float(world) and float(acc[1]) == 1300.0
    # gradient convention: every rank back-propagates world * d(global mean)/d(local pixel) = world / global count, so
    # the all-reduce(AVG) of parameter gradients equals the single-process gradient of the mean over all pixels
    local_grad = torch.full((1,), scale / float(acc[1]) * per_pixel[rank].numel())  # sum over this rank's pixels
    dist.all_reduce(local_grad)
    ok = ok and abs(float(local_grad) / world - 1.0) < 1e-6
    per_rank_mean_avg = (per_pixel[0].mean() + per_pixel[1].mean()) / 2                # the DDP convention differs:
    ok = ok and abs(float(per_rank_mean_avg) - float(want.mean())) > 1e-4
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_global_masked_mean_convention_gloo_world2():
    """losses._global_sums: the loss accumulators of all ranks are summed before the division (reference semantics:
    one masked mean over the gathered batch, runner.py:193-203), unlike an average of per-rank means."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loss_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_graph_segments_of_the_data_parallel_step():
    """Host logic of graphs.GraphedDataParallelStep: which phases share a CUDA graph for a given set of cut points."""
    from camradepth_b200.graphs import GraphedDataParallelStep as G
    assert G.plan_segments(True) == [["fwd", "loss", "stage3", "stage2", "stage1", "stage0", "opt"]]
    assert G.plan_segments(False) == [["fwd"], ["loss", "stage3", "stage2", "stage1"], ["stage0"], ["opt"]]
    assert G.plan_segments(False, "") == [["fwd"], ["loss", "stage3", "stage2", "stage1", "stage0"], ["opt"]]
    assert G.plan_segments(False, "stage3,stage2,stage1,stage0") == \
        [["fwd"], ["loss", "stage3"], ["stage2"], ["stage1"], ["stage0"], ["opt"]]
    for segs in (G.plan_segments(False), G.plan_segments(False, "stage2,stage0")):
        flat = [p for s in segs for p in s]
        assert flat == ["fwd", "loss", "stage3", "stage2", "stage1", "stage0", "opt"]      # every phase once, in order
    import pytest
    with pytest.raises(ValueError):
        G.plan_segments(False, "stage7")
