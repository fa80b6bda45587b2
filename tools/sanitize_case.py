"""Small fused forward / forward+backward / stand-alone backward calls (partial tiles, S=1..4) for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import md2_b200.cabi as cabi
from test_gpu_parity import synth_args
cl = cabi.CLoss()
for B, H, W, fids, am in ((2, 40, 72, [0, -1, 1], True), (1, 32, 64, [0, 1], False), (2, 64, 96, [0, -1, 1, "s", 2], True)):
    a = synth_args(B, H, W, fids, am, "smooth", 3)
    f = cl.forward(a)
    o = cl.forward_backward(a)
    b = cl.backward(a, o["argmin"], 1.0)
    a["noise"] = None
    o2 = cl.forward_backward(a)
    torch.cuda.synchronize()
    print(B, H, W, len(fids) - 1, float(o["loss"]), float(f["loss"]), float(o2["loss"]))
print("sanitize case done")
