"""Can this stack capture NCCL all-reduces issued from autograd hooks inside a CUDA graph?  (torchrun, 2 ranks)
Prints one line per variant; every variant is bounded by the caller's `timeout`."""
import os, sys, time
import torch, torch.distributed as dist, torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
variant = sys.argv[1]
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank()
torch.manual_seed(0)
net = nn.Sequential(nn.Conv2d(3, 32, 3, padding=1), nn.BatchNorm2d(32), nn.ReLU(), nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(),
                    nn.Conv2d(32, 1, 3, padding=1)).to(dev)
x = torch.rand(4, 3, 64, 64, device=dev) + rank
t0 = time.time()
if variant == "main_thread":
    # all-reduce issued from the capturing thread itself, after backward
    flat = torch.zeros(1 << 20, device=dev)
    dist.all_reduce(flat)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            w = dist.all_reduce(flat, async_op=True); w.wait()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        flat.add_(1.0)
        w = dist.all_reduce(flat, async_op=True)
        w.wait()
        flat.mul_(0.5)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    print(f"rank {rank} main_thread ok {time.time() - t0:.1f}s value {float(flat[0]):.3f}", flush=True)
else:
    from md2_b200.trainer import GraphedTrainStep
    opt = torch.optim.Adam(net.parameters(), 1e-3)
    bp = lambda inputs: {"loss": net(inputs["x"]).square().mean()}
    step = GraphedTrainStep(net, bp, opt, {"x": x}, graph=True, comm=variant, buckets=3,
                            broadcast_buffers=(os.environ.get("PROBE_BUFFERS", "1") == "1"))
    for i in range(5):
        loss = step({"x": x})
    torch.cuda.synchronize()
    w0 = next(net.parameters()).detach().flatten()[:4].clone()
    g = [torch.zeros_like(w0) for _ in range(dist.get_world_size())]
    dist.all_gather(g, w0)
    print(f"rank {rank} {variant} ok {time.time() - t0:.1f}s loss {float(loss):.5f} weights equal across ranks: "
          f"{bool(torch.equal(g[0], g[1]))}", flush=True)
    if os.environ.get("PROBE_CLOSE", "1") == "1":
        step.close()
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank} exited cleanly after {time.time() - t0:.1f}s", flush=True)
