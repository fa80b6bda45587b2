"""N>1 host logic on CPU: two gloo ranks, each running the kernel logic (host emulation) on its
shard of the batch.  Equal shards + mean-of-means must reproduce the single-process loss, and the
per-rank gradients divided by the world size must equal the global-batch gradients of the shard's
images (what DDP's gradient averaging relies on)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from host_emu import emu
        import md2_b200.distributed as D
        from test_kernel_logic_emu import synth_args
        torch.set_num_threads(1)
        full = synth_args(4, 32, 64, [0, -1, 1], True, "smooth", 30)
        mine = D.shard_batch(full, D.rank(), D.world())
        out = emu.forward_backward(mine)
        loss = D.mean_over_ranks(out["loss"])
        tmax = D.max_over_ranks(float(rank + 1))
        gathered = [torch.zeros_like(out["grad_disp"][0]) for _ in range(world)]
        dist.all_gather(gathered, out["grad_disp"][0].contiguous())
        if rank == 0:
            ref = emu.forward_backward(full)
            ret["loss"] = (float(loss), float(ref["loss"]))
            ret["tmax"] = tmax
            ret["grad"] = float((torch.cat(gathered) / world - ref["grad_disp"][0]).abs().max() /
                                ref["grad_disp"][0].abs().max())
            ret["argmin_equal"] = bool(torch.equal(out["argmin"], ref["argmin"][:, :2]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    a, b = ret["loss"]
    assert a == pytest.approx(b, rel=2e-6)
    assert ret["tmax"] == 2.0
    assert ret["grad"] <= 1e-5
    assert ret["argmin_equal"]


def test_shard_range_rejects_ragged_batches():
    import md2_b200.distributed as D
    assert D.shard_range(12, 1, 4) == (3, 6)
    with pytest.raises(ValueError):
        D.shard_range(10, 0, 4)
