"""BASELINE.json configs[3] / configs[4]: fused-loss throughput over resolution x source count on one GPU,
against the HBM roofline (algorithmic bytes B*H*W*(183.8125 + 112*S)).
Usage: python tools/sweep.py [out.json]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import md2_b200.cabi as cabi
from test_gpu_parity import synth_args

pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
cl = cabi.CLoss()
FR = {1: [0, 1], 2: [0, -1, 1], 3: [0, -1, 1, "s"], 4: [0, -1, 1, "s", 2]}
cases = [(12, h, w, s) for (h, w) in ((96, 320), (192, 640), (288, 960), (320, 1024), (384, 1280)) for s in (1, 2, 3, 4)]
cases.append((8, 320, 1024, 3))  # configs[3]: mono+stereo high-res, per-GPU batch 8
if os.environ.get("MD2_SWEEP_CASES"):  # "B,H,W,S;B,H,W,S;..." for quick A/B runs
    cases = [tuple(int(x) for x in c.split(",")) for c in os.environ["MD2_SWEEP_CASES"].split(";")]
rows = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, H, W, S in cases:
    a = synth_args(B, H, W, FR[S], True, "iid", 0)
    a["noise"] = None
    ts = []
    for i in range(8):
        flush.zero_()  # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cl.forward_backward(a); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    px = B * S * 4 * H * W
    algo = B * H * W * (183.8125 + 112.0 * S)
    rows.append({"B": B, "H": H, "W": W, "S": S, "ms_fwd_bwd": ms, "warped_px_per_s": px / (ms * 1e-3),
                 "algorithmic_MB": algo / 1e6, "achieved_GBps": algo / (ms * 1e-3) / 1e9,
                 "roofline_frac": algo / (ms * 1e-3) / 1e9 / peak})
    print(json.dumps(rows[-1]), flush=True)
    del a
    torch.cuda.empty_cache()
if len(sys.argv) > 1:
    json.dump({"peak_GBps": peak, "l2": "256 MB buffer zeroed between timed iterations", "rows": rows},
              open(sys.argv[1], "w"), indent=1)
