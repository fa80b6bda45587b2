"""Kernel-time table of one eager training step (tools/train_step.py harness, fused loss, channels-last) from
torch.profiler: which stock PyTorch / cuDNN kernels the step spends its time in.  python tools/profile_train_step.py [out.txt]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import train_step as ts  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
layers = 50 if "--config3" in sys.argv else 18
cfg = (8, 320, 1024, [0, -1, 1, "s"]) if layers == 50 else (12, 192, 640, [0, -1, 1])
step, imgs = ts.make_step("fused", *cfg, dev, ddp=False, graph=False, channels_last=True, layers=layers)
for _ in range(5):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes="--shapes" in sys.argv) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
# kernels first (rows whose name is a kernel), then the ATen operators that launched them (self CUDA time)
avg = prof.key_averages(group_by_input_shape="--shapes" in sys.argv)
table = avg.table(sort_by="self_cuda_time_total", row_limit=70, max_name_column_width=150)
print(table)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if args:
    open(args[0], "w").write(table)
