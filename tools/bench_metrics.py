"""Times the depth-metrics drop-in (md2_b200.metrics, 9 launches, no sync) against the eager PyTorch restatement
of model_metric.py:70-106 on the same GPU (oracle, ~40 ATen launches + 2 sorts).  Batch 12, 192x640 prediction,
375x1242 ground truth at LiDAR-like 5 % density.  CUDA events, 20 iterations after 5 warm-ups.  One JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import md2_b200.metrics as M  # noqa: E402
from oracle import oracle_torch as O  # noqa: E402  (baseline leg only)


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    out = {}
    for density in (0.05, 1.0):
        depth = 0.5 + 60 * torch.rand(12, 1, 192, 640, generator=g, device=dev)
        gt = torch.zeros(12, 1, 375, 1242, device=dev)
        hit = torch.rand(12, 1, 375, 1242, generator=g, device=dev) < density
        gt[hit] = 1.0 + 79 * torch.rand(int(hit.sum()), generator=g, device=dev)
        fused = timed(lambda: M.depth_metrics(depth, gt))
        eager = timed(lambda: O.depth_metrics(depth, gt))
        # the reference's logger additionally moves each metric to the host on its own (logger.py:31-35)
        eager_sync = timed(lambda: [v.cpu() for v in O.depth_metrics(depth, gt)])
        fused_sync = timed(lambda: M.depth_metrics(depth, gt).cpu())
        out[f"density_{density}"] = {"md2_ms": round(fused, 4), "eager_ms": round(eager, 4),
                                     "md2_with_host_read_ms": round(fused_sync, 4),
                                     "eager_with_per_metric_cpu_ms": round(eager_sync, 4),
                                     "masked_pixels": int(M.depth_metrics(depth, gt)[7])}
    print(json.dumps({"workload": "depth_metrics batch12 192x640 -> 375x1242", **out}))


if __name__ == "__main__":
    main()
