"""Times the colour-pyramid kernels (md2_b200.pipeline) for one training batch - batch 12 x 3 frames = 36 decoded
375x1242 frames -> 4 levels of 192x640 - against Pillow on the host (what the reference's loader workers run,
kitti_mono.py:352-355), and reports the HBM roofline fraction of the byte work.  One JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import md2_b200.pipeline as P  # noqa: E402


def main():
    N, Hin, Win, H, W, S = 36, 375, 1242, 192, 640, 4
    dev = "cuda:0"
    rng = np.random.default_rng(0)
    host = rng.integers(0, 256, (N, Hin, Win, 3), dtype=np.uint8)
    sets = [torch.from_numpy(np.roll(host, i, axis=0)).to(dev) for i in range(3)]  # 3 x 50 MB > L2
    flip = torch.tensor([i % 2 for i in range(N)], dtype=torch.uint8, device=dev)
    pyr = P.ColorPyramid(N, Hin, Win, H, W, S, device=dev)
    for i in range(5):
        pyr(sets[i % 3], flip)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    iters = 30
    a.record()
    for i in range(iters):
        pyr(sets[i % 3], flip)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    bytes_alg = N * (Hin * Win * 3 + sum(3 * (H >> s) * (W >> s) * 4 for s in range(S)))
    peak = 6539.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    # colour jitter of level 0 (the only jittered level the networks read), all 36 frames
    jit = P.ColorJitter([2, 0, 3, 1], 1.13, 0.87, 1.19, -0.07)
    lv0 = pyr(sets[0], flip)[0]
    for _ in range(3):
        jit(lv0)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        jit(lv0)
    b.record()
    torch.cuda.synchronize()
    jit_ms = a.elapsed_time(b) / iters
    # ToTensor alone (md2_to_tensor): the headline batch's image groups - target pyramid (4 levels) and two source
    # frames, batch 12 - as bytes -> float planes, one launch; rotating sets so that nothing is served from L2
    shapes = [(12, H >> s, W >> s) for s in range(S)] + [(12, H, W)] * 2
    tsets = [[torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).to(dev) for n, h, w in shapes]
             for _ in range(3)]
    touts = [[torch.empty(n, 3, h, w, device=dev) for n, h, w in shapes] for _ in range(3)]   # 3 x 73.6 MB > L2
    for i in range(3):
        P.to_tensor(tsets[i], out=touts[i])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()   # the launch is shorter than its Python call: replay three of them from a graph
    with torch.cuda.graph(graph):
        for i in range(3):
            P.to_tensor(tsets[i], out=touts[i])
    graph.replay()
    torch.cuda.synchronize()
    a.record()
    for i in range(10):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    tt_ms = a.elapsed_time(b) / 30
    tt_bytes = sum(n * h * w * 3 * 5 for n, h, w in shapes)
    from PIL import Image
    t0 = time.perf_counter()
    n_cpu = 4
    for n in range(n_cpu):
        pil = Image.fromarray(host[n])
        for s in range(S):
            np.asarray(pil.resize((W >> s, H >> s), Image.LANCZOS), dtype=np.float32).transpose(2, 0, 1) / 255
    cpu_ms_per_image = (time.perf_counter() - t0) / n_cpu * 1e3
    print(json.dumps({"workload": f"colour pyramid, {N} frames {Hin}x{Win} u8 -> {S} levels of {H}x{W} f32",
                      "ms_per_batch": round(ms, 4), "images_per_s": round(N / (ms * 1e-3)),
                      "algorithmic_MB": round(bytes_alg / 1e6, 2), "achieved_GBps": round(bytes_alg / (ms * 1e-3) / 1e9, 1),
                      "peak_GBps": peak, "roofline_frac": round(bytes_alg / (ms * 1e-3) / 1e9 / peak, 4),
                      "gpu_launches_per_batch": 1 + 2 * S,
                      "color_jitter_level0_ms_per_batch": round(jit_ms, 4),
                      "to_tensor": {"workload": "ToTensor of the headline batch's image groups (6 groups, one launch)",
                                    "ms_per_batch": round(tt_ms, 5), "algorithmic_MB": round(tt_bytes / 1e6, 2),
                                    "achieved_GBps": round(tt_bytes / (tt_ms * 1e-3) / 1e9, 1),
                                    "roofline_frac": round(tt_bytes / (tt_ms * 1e-3) / 1e9 / peak, 4)},
                      "pillow_host_ms_per_image_1_thread": round(cpu_ms_per_image, 2),
                      "pillow_host_ms_per_batch_1_thread": round(cpu_ms_per_image * N, 1)}))


if __name__ == "__main__":
    main()
