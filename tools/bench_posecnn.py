"""Times the posecnn branch of the drop-in (composed from the symbol-level kernels, compute.py) against the fused
path and against the eager oracle, batch 12, 192x640, forward + backward.  One JSON line."""
import json
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import md2_b200.synthetic as syn  # noqa: E402
from md2_b200 import functional as F_  # noqa: E402
from md2_b200.compute import compute  # noqa: E402
from oracle import oracle_torch as O  # noqa: E402  (baseline leg only)


def timed(fn, iters=10, warm=3):
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
    dev = torch.device("cuda:0")
    B, H, W, fids = 12, 192, 640, [0, -1, 1]
    inputs, outputs = syn.make_batch(B, H, W, fids, 4, 0, "iid", device=dev, requires_grad=True)
    noise = syn.make_noise(B, 2, H, W, 4, 0, device=dev)
    for f in fids[1:]:
        outputs[("R", f, 0)] = outputs[("axisangle", f)].detach()[:, None].clone().requires_grad_(True)
        outputs[("T", f, 0)] = outputs[("translation", f)].detach()[:, None].clone().requires_grad_(True)
    mk = lambda pt: SimpleNamespace(frame_ids=fids, scales=range(4), height=H, width=W, min_depth=0.1, max_depth=100.0,
                                    pose_type=pt, use_automasking=True, disp_smoothness=1e-3)
    leaves = [outputs[("disp", s)] for s in range(4)] + [outputs[("R", f, 0)] for f in fids[1:]] + [outputs[("T", f, 0)] for f in fids[1:]]

    def run(pt):
        c = compute(mk(pt), dev)
        o = dict(outputs)
        if pt != "posecnn":
            for f in fids[1:]:
                o[("c2c", f, 0)] = F_.param2matrix(o[("R", f, 0)][:, 0], o[("T", f, 0)][:, 0], invert=(f < 0))
        c.image2warping(inputs, o, None, noise=noise)
        loss = c.compute_loss(inputs, o, None)["loss"]
        torch.autograd.grad(loss, leaves, allow_unused=True)

    def run_oracle():
        res = O.view_synthesis_loss(inputs[("color", 0, 0)], [inputs[("color", f, 0)] for f in fids[1:]],
                                    [outputs[("disp", s)] for s in range(4)], [inputs[("color", 0, s)] for s in range(4)],
                                    inputs[("K", 0)], inputs[("inv_K", 0)], None, noise=noise,
                                    posecnn=[(outputs[("R", f, 0)][:, 0], outputs[("T", f, 0)][:, 0], f < 0) for f in fids[1:]])
        torch.autograd.grad(res["loss"], leaves, allow_unused=True)

    print(json.dumps({"workload": "loss fwd+bwd batch 12 192x640 S=2 through md2_b200.compute",
                      "fused_separate_ms": round(timed(lambda: run("separate")), 3),
                      "composed_posecnn_ms": round(timed(lambda: run("posecnn")), 3),
                      "eager_oracle_posecnn_ms": round(timed(run_oracle), 3)}))


if __name__ == "__main__":
    main()
