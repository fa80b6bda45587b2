"""Time every build/variants/*.so on the benchmark configuration (one subprocess per library)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, ctypes as C
sys.path.insert(0, %r)
import torch
import md2_b200.cabi as cabi
sys.argv = ["bench"]
import bench
cl = cabi.CLoss()
inputs, outputs = bench.make_host_batch(0)
a = bench.to_args(inputs, outputs, "cuda"); a["seed"] = 1
out = cl.forward_backward(a)
torch.cuda.synchronize()
k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
k0.record(); k1.record(); torch.cuda.synchronize()
ts = []
for i in range(12):
    cl.forward_backward(a, events=(k0, k1)); torch.cuda.synchronize(); ts.append(k0.elapsed_time(k1))
ts = sorted(ts[2:])
print(json.dumps({"lib": os.path.basename(os.environ["MD2_LIB"]), "kernel_ms_med": ts[len(ts)//2], "kernel_ms_min": ts[0],
                  "loss": float(out["loss"]), "gsum": float(out["grad_disp"][0].double().abs().sum())}))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))):
    env = dict(os.environ, MD2_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-500:], flush=True)
