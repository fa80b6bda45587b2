"""md2_b200.trainer on the CPU: the flat gradient buffer, its buckets and the hook-driven all-reduce,
single process and two gloo ranks against DistributedDataParallel.  (The CUDA-graph path is covered by
tests/test_gpu_trainer.py on the B200.)"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def make_nets(seed):
    torch.manual_seed(seed)
    enc = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1))
    dec = nn.Sequential(nn.Conv2d(8, 4, 3, padding=1), nn.ELU(), nn.Conv2d(4, 1, 3, padding=1), nn.Sigmoid())
    return {"encoder": enc, "decoder": dec}


def batch_process_of(models):
    def batch_process(inputs):
        d = models["decoder"](models["encoder"](inputs["color"]))
        return {"loss": ((d - inputs["target"]) ** 2).mean()}
    return batch_process


def data(seed, n=4):
    g = torch.Generator().manual_seed(seed)
    return {"color": torch.rand(n, 3, 16, 24, generator=g), "target": torch.rand(n, 1, 16, 24, generator=g)}


def test_flat_gradients_single_process_matches_plain_autograd():
    from md2_b200.trainer import GraphedTrainStep
    a, b = make_nets(0), make_nets(0)
    pa = [p for m in a.values() for p in m.parameters()]
    pb = [p for m in b.values() for p in m.parameters()]
    oa = torch.optim.Adam(pa, 1e-2)
    ob = torch.optim.Adam(pb, 1e-2)
    step = GraphedTrainStep(a, batch_process_of(a), oa, data(0), graph=False, buckets=3)
    br = step.grads.bucket_range
    assert 2 <= len(br) <= 3 and br[0][0] == 0 and br[-1][1] == step.flat_gradients.numel()
    assert all(br[i][1] == br[i + 1][0] for i in range(len(br) - 1))  # contiguous segments of the flat buffer
    for i in range(4):
        la = step(data(i))
        ob.zero_grad()
        lb = batch_process_of(b)(data(i))["loss"]
        lb.backward()
        ob.step()
        assert float(la) == pytest.approx(float(lb), rel=1e-6)
    for x, y in zip(pa, pb):
        assert torch.allclose(x, y, rtol=1e-5, atol=1e-7)
        assert x.grad.data_ptr() >= step.flat_gradients.data_ptr()  # still a view of the flat buffer


def test_unused_parameters_are_found_once_and_a_changing_step_is_reported():
    from md2_b200.trainer import GraphedTrainStep
    a = make_nets(1)
    extra = nn.Linear(3, 3)  # never used by batch_process (like the fc head of a torchvision encoder)
    models = dict(a, unused=extra)
    opt = torch.optim.SGD([p for m in models.values() for p in m.parameters()], 1e-2)
    flag = {"skip_decoder": False}

    def bp(inputs):
        f = a["encoder"](inputs["color"])
        d = f[:, :1] if flag["skip_decoder"] else a["decoder"](f)
        return {"loss": ((d - inputs["target"]) ** 2).mean()}
    step = GraphedTrainStep(models, bp, opt, data(0), graph=False, buckets=2)
    step(data(0))                       # calibration step: the unused head is dropped from the bucket counts
    assert {id(p) for p in step.grads.unused} == {id(p) for p in extra.parameters()}
    step(data(1))                       # static step: fine
    flag["skip_decoder"] = True         # the step changes shape: parameters that used to get gradients no longer do
    with pytest.raises(RuntimeError, match="no gradient"):
        step(data(2))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        from md2_b200.trainer import GraphedTrainStep
        ours = make_nets(10 + rank)          # different initial weights per rank: the ctor must broadcast rank 0's
        ref = make_nets(10)
        ddp = {k: nn.parallel.DistributedDataParallel(m) for k, m in ref.items()}
        oo = torch.optim.Adam([p for m in ours.values() for p in m.parameters()], 1e-2)
        orf = torch.optim.Adam([p for m in ref.values() for p in m.parameters()], 1e-2)
        step = GraphedTrainStep(ours, batch_process_of(ours), oo, data(0), graph=False, buckets=3)
        losses = []
        for i in range(3):
            batch = data(100 * i + rank)      # every rank its own shard
            lo = step(batch)
            orf.zero_grad()
            lr = batch_process_of(ddp)(batch)["loss"]
            lr.backward()
            orf.step()
            losses.append((float(lo), float(lr)))
        err = max(float((x - y).abs().max()) for x, y in zip((p for m in ours.values() for p in m.parameters()),
                                                              (p for m in ref.values() for p in m.parameters())))
        bn = float((ours["encoder"][1].running_mean - ref["encoder"][1].running_mean).abs().max())
        if rank == 0:
            ret["losses"] = losses
            ret["param_err"] = err
            ret["bn_err"] = bn
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_match_distributed_data_parallel():
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    for a, b in ret["losses"]:
        assert a == pytest.approx(b, rel=1e-5)
    assert ret["param_err"] <= 1e-5
    assert ret["bn_err"] <= 1e-6


def test_branch_streams_without_cuda_run_in_order():
    """md2_b200.trainer.BranchStreams on a host without CUDA (or disabled): the callables run sequentially, results in
    argument order."""
    from md2_b200.trainer import BranchStreams
    log = []
    br = BranchStreams(enabled=torch.cuda.is_available())
    out = br(lambda: log.append("main") or {"a": torch.ones(2)}, lambda: log.append("side1") or [torch.zeros(1)],
             lambda: log.append("side2") or 3)
    assert isinstance(out, list) and len(out) == 3 and out[2] == 3 and torch.equal(out[0]["a"], torch.ones(2))
    assert sorted(log) == ["main", "side1", "side2"]
    assert BranchStreams(enabled=False)(lambda: 1, lambda: 2) == [1, 2]
    assert BranchStreams()(lambda: 7) == [7]
