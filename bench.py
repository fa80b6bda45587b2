#!/usr/bin/env python
"""bench.py - throughput of the fused view-synthesis loss (forward + backward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): warped pixels per second of the fused loss forward+backward, where a
warped pixel is one (image, scale, source frame, y, x) sample: B * S * scales * H * W per step.
Workload at every N: BASELINE.json configs[1] per GPU - batch 12, 192x640, frame_ids [0,-1,1]
(S=2), 4 scales, fp32, auto-mask on, synthetic i.i.d. U(0,1) images (weak scaling: every rank
runs its own batch; the path has no data-path collective).

A step is one pass of the hot path over one batch:
  value    the three launches of md2_loss_forward_backward through the C ABI, inputs resident in
           HBM.  The step rotates over several independent input/output sets whose total size
           exceeds the 126 MB L2, so no step finds its inputs cached.
  e2e      the same step through the public Python API (md2_b200.compute: image2warping +
           compute_loss + loss.backward()) with HOST inputs: every step copies its batch from
           pinned host memory (on a copy stream, overlapping the previous step's kernels) and
           reads the loss back.
  roofline the tile kernel alone (CUDA events recorded around it on its stream), algorithmic
           bytes N*(183.8125 + 112*S) against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle port of the reference's PyTorch path on the host cores, bounded sample.

  gpu_reference  the reference's own loss step (oracle/_ref when staged, else the port) on the SAME GPU at the
           same configuration, forward + backward, CUDA-event timed: as shipped (host torch.randn + H2D per scale,
           processor.py:195) and with the noise drawn on the device (SURVEY.md 8d ii).

--impl reference times the reference's CPU implementation of the path on the host cores: the unmodified
reference staged by oracle/stage_ref.py into the git-ignored oracle/_ref/ (it travels to the GPU box like a built
.so), else the port oracle/oracle_torch.py, which is pinned to the reference's outputs by tests/test_oracle_golden.py.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B, H, W, FRAME_IDS, NUM_SCALES = 12, 192, 640, [0, -1, 1], 4
S = len(FRAME_IDS) - 1
# "iid": every disparity value is an independent U(0,1) draw, i.e. neighbouring pixels sample the source frames up to
# 37 pixels apart (the worst case for the gather); "smooth": band-limited disparities like a network's
DISP_KIND = os.environ.get("MD2_BENCH_DISP", "iid")
METRIC = "warped_pixels_per_sec_fused_loss_fwd_bwd"
UNIT = "warped_px/s"
WORKLOAD = "fused view-synthesis loss fwd+bwd, batch 12 per GPU, 192x640, frame_ids [0,-1,1], 4 scales, automask, fp32"


def algorithmic_bytes(b, h, w, s):
    return b * h * w * (183.8125 + 112.0 * s)


def cross_scale_bytes(b, h, w, s):
    """SURVEY.md 8d cross-scale lower bound (every distinct input read once per pass for all four scales, the loss
    reduced in-kernel, only depth(0,0) written): 17.5 bytes per warped pixel, 206 MB at batch 12, 192x640, S=2."""
    return 17.5 * b * s * NUM_SCALES * h * w


def gpu_reference(dev, ours_ms):
    """The reference's loss step on this GPU at the benchmark configuration (SURVEY.md 8d ii), forward + backward,
    CUDA events, median of 20 after 3 warm-up steps: as shipped and with the auto-mask noise drawn on the device."""
    sys.path.insert(0, ROOT)
    res = {}
    try:
        from oracle import ref_loader as RL
        if RL.available():
            mk = lambda host_noise: RL.make_step(B, H, W, FRAME_IDS, dev, NUM_SCALES, 0, "iid", host_noise)
            res["kind"] = "reference (oracle/_ref: model_tool/processor.py image2warping + compute_loss + backward)"
        else:
            raise RuntimeError("not staged")
    except Exception:
        from oracle import oracle_torch as O
        import md2_b200.synthetic as syn
        inputs, outputs = syn.make_batch(B, H, W, FRAME_IDS, NUM_SCALES, 0, "iid", requires_grad=False)
        g = lambda t: t.to(dev)
        srcs = FRAME_IDS[1:]
        Ts = [O.pose_matrix(g(outputs[("axisangle", f)]), g(outputs[("translation", f)]), invert=(f < 0)).detach()
              for f in srcs]
        base = dict(target=g(inputs[("color", 0, 0)]), sources=[g(inputs[("color", f, 0)]) for f in srcs],
                    color_pyr=[g(inputs[("color", 0, s)]) for s in range(NUM_SCALES)], K=g(inputs[("K", 0)]),
                    inv_K=g(inputs[("inv_K", 0)]), automask=True)
        disp0 = [g(outputs[("disp", s)]) for s in range(NUM_SCALES)]

        def mk(host_noise):
            def step():
                disps = [d.clone().requires_grad_(True) for d in disp0]
                T = [t.clone().requires_grad_(True) for t in Ts]
                nz = None if host_noise else [torch.randn(B, S, H, W, device=dev) for _ in range(NUM_SCALES)]
                out = O.view_synthesis_loss(disps=disps, Ts=T, noise=nz, **base)
                out["loss"].backward()
                return out["loss"].detach()
            return step
        res["kind"] = "port (oracle/oracle_torch.py on cuda)"

    def timed(step, n=20):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], ts[0]
    a_med, a_min = timed(mk(True))
    b_med, b_min = timed(mk(False))
    res.update({"ms_as_shipped": a_med, "ms_as_shipped_min": a_min, "ms_noise_on_device": b_med,
                "ms_noise_on_device_min": b_min, "ours_ms_per_step": ours_ms,
                "config": "batch 12, 192x640, frame_ids [0,-1,1], 4 scales, fp32, same GPU, eager PyTorch ops"})
    return res


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region by an `nvidia-smi -lms` subprocess
    (the profiling recipe's clocks line); samples are kept when their timestamp falls inside the region."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.t0, self.t1 = index, None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)  # let it start sampling before the region begins
        except Exception:
            self.proc = None

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *a):
        self.t1 = time.time()
        time.sleep(0.03)
        if self.proc:
            self.proc.terminate()
            try:
                self.out = self.proc.communicate(timeout=5)[0]
            except Exception:
                self.out = ""
        else:
            self.out = ""

    def summary(self):
        import datetime
        sm, mx, reasons, all_sm = [], 0.0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.out.splitlines():
            r = [x.strip() for x in line.split(",")]
            if len(r) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(r[1]), float(r[2])
            except ValueError:
                continue
            mx = max(mx, cmax)
            all_sm.append(clk)
            if self.t0 - 0.02 <= ts <= self.t1 + 0.02:
                sm.append(clk)
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if not sm:  # region shorter than the sampling period: take the nearest samples
            sm = all_sm[-3:]
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one tile-kernel launch, from the newest committed
    `ncu --set full` summary of this workload under profiles/ (a profiler pass is never part of a bench run)."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_tile_kernel_ncu_full.json"))):
        best = path
    if best is None:
        return None, None
    try:
        d = json.load(open(best))
        to_bytes = lambda v: float(v.split()[0]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[v.split()[1]]
        if not re.search(r"Tile<2, 1, 32, 16", d.get("Kernel Name", "")):
            return None, None
        return to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"]), os.path.relpath(best, ROOT)
    except Exception:
        return None, None


def make_host_batch(seed, dev="cuda"):
    """One synthetic batch on the host, shaped like the reference loaders' output (pose matrices from this
    package's own param2matrix kernel)."""
    import md2_b200.synthetic as syn
    from md2_b200.functional import param2matrix
    pose_matrix = lambda a, t, invert: param2matrix(a.to(dev), t.to(dev), invert).cpu()
    inputs, outputs = syn.make_batch(B, H, W, FRAME_IDS, NUM_SCALES, seed, "iid", requires_grad=False)
    if DISP_KIND == "smooth":
        # network-like disparities: band-limited fields instead of i.i.d. values (see config["gather"])
        smooth = syn.make_batch(B, H, W, FRAME_IDS, NUM_SCALES, seed, "smooth", requires_grad=False)[1]
        for s in range(NUM_SCALES):
            outputs[("disp", s)] = smooth[("disp", s)]
    for f in FRAME_IDS[1:]:
        outputs[("c2c", f, 0)] = pose_matrix(outputs[("axisangle", f)], outputs[("translation", f)],
                                             invert=(f < 0)).detach()
    return inputs, outputs


def to_args(inputs, outputs, dev):
    g = lambda t: t.to(dev, non_blocking=True)
    srcs = FRAME_IDS[1:]
    return dict(target=g(inputs[("color", 0, 0)]), sources=[g(inputs[("color", f, 0)]) for f in srcs],
                disps=[g(outputs[("disp", s)]) for s in range(NUM_SCALES)],
                color_pyr=[g(inputs[("color", 0, s)]) for s in range(NUM_SCALES)],
                K=g(inputs[("K", 0)]), inv_K=g(inputs[("inv_K", 0)]),
                Ts=[g(outputs[("c2c", f, 0)]) for f in srcs], automask=True, noise=None)


# ------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import ctypes as C
    import torch.distributed as dist
    import md2_b200.cabi as cabi
    from md2_b200.compute import compute as FusedCompute
    from types import SimpleNamespace

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours): no CUDA device - the fused loss has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cl = cabi.CLoss()
    lib = cl.lib
    # ---- device-resident leg: rotating sets larger than L2 -------------------------------------
    n_sets = 3
    host = [make_host_batch(100 * rank + i, dev) for i in range(n_sets)]
    sets = []
    set_bytes = 0
    for inputs, outputs in host:
        a = to_args(inputs, outputs, dev)
        a["seed"] = 1
        prep, cfg, inp = cl._prep(a)
        o = cl._alloc_out(cfg, dev)
        gd = [torch.empty_like(d) for d in prep["disps"]]
        gT = [torch.empty(B, 4, 4, device=dev) for _ in prep["Ts"]]
        ws = cl._ws(cfg, dev)
        out = cabi.make_outputs(o["loss"], None, o["argmin"], o["depth"])
        g = cabi.make_grads(gd, gT)
        sets.append((prep, cfg, inp, out, g, ws, o, gd, gT))
        tens = [prep["target"]] + prep["sources"] + prep["disps"] + prep["color_pyr"][1:] + \
               [o["argmin"], o["depth"]] + gd
        set_bytes = sum(t.numel() * t.element_size() for t in tens)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step(i):
        prep, cfg, inp, out, g, ws, *_ = sets[i % n_sets]
        rc = lib.md2_loss_forward_backward(C.byref(cfg), C.byref(inp), C.byref(out), C.byref(g), C.c_float(1.0),
                                           C.c_void_p(ws.data_ptr()), stream)
        if rc != 0:
            raise RuntimeError(f"md2_loss_forward_backward returned {rc}")

    for i in range(args.warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = clk.summary()

    # ---- tile kernel alone (roofline leg): events recorded around the launch on its stream ------
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(); k1.record()  # materialise the cudaEvent_t handles
    torch.cuda.synchronize()
    kms = []
    for i in range(min(args.steps, 20)):
        prep, cfg, inp, out, g, ws, *_ = sets[i % n_sets]
        rc = lib.md2_loss_forward_backward_timed(C.byref(cfg), C.byref(inp), C.byref(out), C.byref(g), C.c_float(1.0),
                                                 C.c_void_p(ws.data_ptr()), stream, C.c_void_p(k0.cuda_event),
                                                 C.c_void_p(k1.cuda_event))
        if rc != 0:
            raise RuntimeError(f"md2_loss_forward_backward_timed returned {rc}")
        torch.cuda.synchronize()
        kms.append(k0.elapsed_time(k1))
    kernel_ms = sum(kms) / len(kms)

    # ---- the two split entry points (validation forward; stand-alone backward), for reference ----------
    def time_calls(fn, n=20):
        fn(0)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(n):
            fn(i)
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / n

    def fwd_only(i):
        prep, cfg, inp, out, g, ws, *_ = sets[i % n_sets]
        assert lib.md2_loss_forward(C.byref(cfg), C.byref(inp), C.byref(out), C.c_void_p(ws.data_ptr()), stream) == 0

    gl_dev = torch.ones(1, device=dev)

    def bwd_only(i):
        prep, cfg, inp, out, g, ws, o, *_ = sets[i % n_sets]
        assert lib.md2_loss_backward(C.byref(cfg), C.byref(inp), C.c_void_p(o["argmin"].data_ptr()),
                                     C.c_void_p(gl_dev.data_ptr()), C.byref(g), C.c_void_p(ws.data_ptr()), stream) == 0

    fwd_ms, bwd_ms = time_calls(fwd_only), time_calls(bwd_only)

    # ---- end-to-end leg: public API, host inputs in pinned memory -------------------------------
    opt = SimpleNamespace(frame_ids=FRAME_IDS, scales=range(NUM_SCALES), height=H, width=W, min_depth=0.1,
                          max_depth=100.0, pose_type="separate", use_automasking=True, disp_smoothness=1e-3)
    comp = FusedCompute(opt, dev)
    # only what the path consumes travels: target pyramid, scale-0 sources, K / inv_K, disparities, poses
    def consumed(k):
        return (k[0] == "color" and (k[2] == 0 or k[1] == 0)) or k[0] in ("K", "inv_K")
    # What a training step receives from the HOST is the loader's batch (images, intrinsics); the disparities and the
    # poses are network outputs and already live on the device (model_train.py:90-96), so they are not uploaded.
    # Two upload formats, one pinned staging buffer per batch and a single H2D copy per step in both:
    #   "u8"   the loader keeps the resized PIL images as bytes [B,H,W,3]; transforms.ToTensor() (kitti_mono.py:283)
    #          runs on the device (md2_b200.pipeline.to_tensor, one launch) - the headline e2e number;
    #   "f32"  the loader's float tensors as the reference's DataLoader yields them (ToTensor in the workers).
    import md2_b200.pipeline as pipeline
    dev_outputs = [{k: v.to(dev) for k, v in outputs.items() if k[0] in ("disp", "c2c")} for _, outputs in host]
    loss_host = [torch.empty((), pin_memory=True) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.Stream(device=dev)

    def e2e_leg(mode):
        pinned = []
        for inputs, _ in host:
            items = [(k, v) for k, v in inputs.items() if consumed(k)]
            layout, off, chunks = [], 0, []
            for key, v in items:
                if mode == "u8" and key[0] == "color":
                    v = (v * 255.0).round().clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
                raw = v.contiguous().view(-1).view(torch.uint8)
                layout.append((key, off, raw.numel(), tuple(v.shape), v.dtype))
                chunks.append((off, raw))
                off += (raw.numel() + 255) // 256 * 256
            flat = torch.zeros(off, dtype=torch.uint8).pin_memory()
            for o, raw in chunks:
                flat[o:o + raw.numel()].copy_(raw)
            pinned.append((flat, layout))
        h2d_bytes = pinned[0][0].numel()
        # two-stage pipeline: a copy stream uploads batch i+1 while the compute stream runs step i (every step
        # still pays its own H2D copy and loss read-back inside the timed region; they overlap with compute)
        staged = {}
        dev_bufs = [torch.empty(h2d_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
        buf_free = [None, None]  # event recorded on the compute stream when the step using the buffer is enqueued

        def upload(i):
            flat, layout = pinned[i % n_sets]
            j = i & 1
            with torch.cuda.stream(copy_stream):
                if buf_free[j] is not None:
                    copy_stream.wait_event(buf_free[j])  # the step that last read this buffer has finished
                dev_bufs[j].copy_(flat, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            staged[i] = (dev_bufs[j], layout, ev)

        def e2e_step(i, last):
            if i not in staged:
                upload(i)
            dflat, layout, ev = staged.pop(i)
            if not last:
                upload(i + 1)
            with torch.cuda.stream(main_stream):
                main_stream.wait_event(ev)
                views = {key: dflat[off:off + n].view(dt).view(shape) for key, off, n, shape, dt in layout}
                if mode == "u8":
                    keys = [k for k in views if k[0] == "color"]
                    inputs = dict(zip(keys, pipeline.to_tensor([views[k] for k in keys])))   # one launch
                    inputs.update({k: v for k, v in views.items() if k[0] != "color"})
                else:
                    inputs = views
                outputs = {k: v.detach().requires_grad_(True) for k, v in dev_outputs[i % n_sets].items()}
                comp.image2warping(inputs, outputs, None)
                comp.compute_loss(inputs, outputs, None)
                outputs["loss"].backward()
                loss_host[i & 1].copy_(outputs["loss"].detach(), non_blocking=True)
                buf_free[i & 1] = torch.cuda.Event()
                buf_free[i & 1].record(main_stream)
            return outputs

        for i in range(args.warmup):
            e2e_step(i, i == args.warmup - 1)
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(main_stream)
        t_host = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i, i == args.steps - 1)
        x1.record(main_stream)
        enqueue_ms = (time.perf_counter() - t_host) * 1e3  # host time to issue the steps (no synchronisation inside)
        barrier()
        ms = x0.elapsed_time(x1)
        # the copies alone: is the end-to-end number bound by the host link or by this code?  Every rank uploads its
        # pinned batch `steps` times with nothing else running (all ranks at once, like the e2e leg)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy_stream):
            c0.record(copy_stream)
            for i in range(args.steps):
                dev_bufs[i & 1].copy_(pinned[i % n_sets][0], non_blocking=True)
            c1.record(copy_stream)
        barrier()
        return ms, c0.elapsed_time(c1), h2d_bytes, enqueue_ms

    e2e_ms_total, h2d_ms_total, h2d_bytes, e2e_enqueue_ms = e2e_leg("u8")
    f32_ms_total, f32_h2d_ms_total, f32_bytes, _ = e2e_leg("f32")

    # ---- max over ranks -------------------------------------------------------------------------
    t = torch.tensor([ms_total, e2e_ms_total, kernel_ms, h2d_ms_total, f32_ms_total, f32_h2d_ms_total], device=dev,
                     dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total, kernel_ms, h2d_ms_total, f32_ms_total, f32_h2d_ms_total = [float(x) for x in t.tolist()]
    px_step = B * S * NUM_SCALES * H * W * world
    value = px_step * args.steps / (ms_total * 1e-3)
    e2e_value = px_step * args.steps / (e2e_ms_total * 1e-3)
    peak, peak_src = hbm_peak()
    algo = algorithmic_bytes(B, H, W, S)
    achieved = algo / (kernel_ms * 1e-3) / 1e9

    traffic, traffic_src = ncu_traffic()
    # bytes the timed launch really moves (no per-pixel loss map written, no noise tensor read: bench passes neither)
    # and the cross-scale lower bound (every distinct input once per step)
    n_px = B * H * W
    algo_matched = algo - NUM_SCALES * n_px * (4 + 4 * S)
    algo_cross = cross_scale_bytes(B, H, W, S)
    gpu_ref = gpu_reference(dev, ms_total / args.steps) if rank == 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "height": H, "width": W, "sources": S,
                   "scales": NUM_SCALES, "warped_px_per_step": px_step,
                   "l2": f"rotating {n_sets} independent input/output sets of {set_bytes / 1e6:.0f} MB each "
                         f"({n_sets * set_bytes / 1e6:.0f} MB > 126 MB L2)",
                   "noise": "auto-mask noise generated on the device (statistically equivalent to the "
                            "reference's host torch.randn)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms_total / args.steps,
                "h2d_only_ms_per_step": h2d_ms_total / args.steps,
                "h2d_only_gbps_per_rank": h2d_bytes * args.steps / (h2d_ms_total * 1e-3) / 1e9,
                "h2d_note": "the same pinned uploads with no kernels running, all ranks at once: when this is close to "
                            "ms_per_step the end-to-end number is bound by the host link, not by the loss",
                "input_format": "uint8 [B,H,W,3] images as the loader's PIL resize leaves them; transforms.ToTensor() "
                                "(kitti_mono.py:283) on the device by md2_b200.pipeline.to_tensor - bit-identical values",
                "gpu_launches_per_step": 6,
                "host_enqueue_ms_per_step": e2e_enqueue_ms / args.steps,
                "f32_upload": {"value": px_step * args.steps / (f32_ms_total * 1e-3), "unit": UNIT,
                               "h2d_bytes_per_step": f32_bytes, "ms_per_step": f32_ms_total / args.steps,
                               "h2d_only_ms_per_step": f32_h2d_ms_total / args.steps,
                               "h2d_only_gbps_per_rank": f32_bytes * args.steps / (f32_h2d_ms_total * 1e-3) / 1e9,
                               "note": "the same step fed the loader's float tensors (ToTensor in the workers, as the "
                                       "reference's DataLoader yields them): four times the bytes, host-link bound"},
                "api": "md2_b200.pipeline.to_tensor + md2_b200.compute.compute.image2warping + compute_loss + "
                       "loss.backward(); the loader's batch (target pyramid, source frames, K, inv_K) from pinned host "
                       "memory on a copy stream (step i+1 uploads while step i computes), disparities / poses on the "
                       "device like network outputs, loss read back"},
        "gpu_launches": 5 * args.steps,
        "split_calls_ms": {"md2_loss_forward": fwd_ms, "md2_loss_backward": bwd_ms,
                           "md2_loss_forward_backward": ms_total / args.steps},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "md2::tile_kernel<Tile<2,true,32,16,320>>",
                     "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": algo, "peak_source": peak_src,
                     "frac_matched": algo_matched / (kernel_ms * 1e-3) / 1e9 / peak,
                     "matched_bytes_per_launch": algo_matched,
                     "frac_cross_scale": algo_cross / (kernel_ms * 1e-3) / 1e9 / peak,
                     "cross_scale_bytes_per_launch": algo_cross,
                     "note": "frac uses SURVEY 8d's per-scale bytes (601 MB); frac_matched drops the per-pixel loss "
                             "write and the noise read this call does not perform; frac_cross_scale counts every "
                             "distinct input once per step"},
    }
    if gpu_ref is not None:
        line["gpu_reference"] = gpu_ref
        line["speedup_vs_gpu_reference"] = {
            "as_shipped_host_randn": gpu_ref["ms_as_shipped"] / (ms_total / args.steps),
            "noise_on_device": gpu_ref["ms_noise_on_device"] / (ms_total / args.steps)}
    # ---- the second half of BASELINE.json's metric: training images/s at this N ------------------------------
    if not args.no_train_step:
        try:
            line["train_step"] = train_step_leg(dev, world, local)
        except Exception as e:  # the headline line must survive a failure of the auxiliary leg
            line["train_step"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        line["cpu_baseline"] = cpu_baseline(steps=3)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def train_step_leg(dev, world, local, steps=30, warmup=8):
    """BASELINE.json configs[2] at this N: full mono training step (ResNet-18 depth + separate pose network on stock
    PyTorch / cuDNN, channels-last, fused loss, Adam), batch 12 per GPU, replayed as one CUDA graph by
    md2_b200.trainer.GraphedTrainStep with the bucket all-reduces captured inside it.  Every rank returns the dict."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import train_step as ts
    torch.manual_seed(0)
    step, imgs = ts.make_step("fused", B, H, W, FRAME_IDS, dev, ddp=world > 1, graph=True, channels_last=True,
                               comm="captured", buckets=6)
    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    final = float(loss.detach())
    step.close()
    return {"metric": "train_images_per_sec", "value": imgs * world / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
            "steps": steps, "warmup": warmup, "n_gpus": world, "final_loss": final,
            "mode": "fused loss, channels-last networks with md2_b200.modules.ReflectionPad2d (NHWC-preserving, "
                    "bit-identical) in the decoder's Conv3x3 and md2_b200.modules.MaxPool2d in the encoder stems, cuDNN "
                    "autotuner on (as the reference sets it, "
                    "model_utility.py:327), fused Adam, whole step one CUDA graph, 6 gradient buckets all-reduced inside "
                    "the graph (md2_b200.trainer.GraphedTrainStep)",
            "config": "ResNet-18 depth + separate ResNet-18 pose net, batch 12 per GPU, 192x640, frame_ids [0,-1,1], "
                      "4 scales, Adam, fp32, synthetic data, random-init weights"}


# ------------------------------------------------------------------------------------ CPU reference
CPU_B = 4  # cpu_baseline of the ours arm: bounded sample, one third of the batch (same image size, sources, scales)


def cpu_step_fn(batch):
    """-> (step, kind).  The reference's own code from oracle/_ref when it is staged, else the port."""
    torch.set_num_threads(os.cpu_count() or 1)
    try:
        from oracle import ref_loader as RL
        if RL.available():
            step = RL.make_step(batch, H, W, FRAME_IDS, "cpu", NUM_SCALES, 0, "iid", True)
            return (lambda: float(step())), "reference"
    except Exception:
        pass
    from oracle import oracle_torch as O
    import md2_b200.synthetic as syn
    inputs, outputs = syn.make_batch(batch, H, W, FRAME_IDS, NUM_SCALES, 0, "iid", requires_grad=False)
    srcs = FRAME_IDS[1:]
    Ts = [O.pose_matrix(outputs[("axisangle", f)], outputs[("translation", f)], invert=(f < 0)).detach()
          for f in srcs]
    base = dict(target=inputs[("color", 0, 0)], sources=[inputs[("color", f, 0)] for f in srcs],
                color_pyr=[inputs[("color", 0, s)] for s in range(NUM_SCALES)], K=inputs[("K", 0)],
                inv_K=inputs[("inv_K", 0)], automask=True, noise=None)

    def step():
        disps = [outputs[("disp", s)].clone().requires_grad_(True) for s in range(NUM_SCALES)]
        T = [t.clone().requires_grad_(True) for t in Ts]
        out = O.view_synthesis_loss(disps=disps, Ts=T, **base)
        out["loss"].backward()
        return float(out["loss"].detach())
    return step, "port"


def cpu_baseline(steps=3):
    step, kind = cpu_step_fn(CPU_B)
    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    px = CPU_B * S * NUM_SCALES * H * W
    what = "the unmodified reference (oracle/_ref)" if kind == "reference" else "oracle/oracle_torch.py"
    return {"value": px / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{steps} steps of batch {CPU_B} (of 12) at 192x640, S=2, 4 scales, fwd+bwd, "
                      f"{what} on the host CPU, {dt:.3f} s/step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, kind = cpu_step_fn(B)  # the benchmark configuration itself: batch 12
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    px = B * S * NUM_SCALES * H * W
    value = px * args.steps / dt
    cores = torch.get_num_threads()
    what = "the unmodified reference staged in oracle/_ref (compute.image2warping + compute_loss + backward)" \
        if kind == "reference" else "the port oracle/oracle_torch.py (oracle/_ref not staged)"
    sample = f"each step = the full batch 12 at 192x640, S=2, 4 scales, fwd+bwd on {cores} host threads, {what}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "height": H, "width": W, "sources": S,
                   "scales": NUM_SCALES, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_train_step(args):
    """BASELINE.json configs[2]: full mono training step, ResNet-18 depth + separate pose network (stock
    PyTorch/cuDNN), batch 12 per GPU, DDP.  --loss fused | eager (eager = PyTorch-op restatement of the
    reference loss, i.e. the reference's own training step on the same GPU)."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import train_step as ts
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    tb, th, tw, tf, layers = (8, 320, 1024, [0, -1, 1, "s"], 50) if args.config3 else (B, H, W, FRAME_IDS, 18)
    step, imgs = ts.make_step(args.loss, tb, th, tw, tf, dev, ddp=world > 1, graph=args.graph,
                               channels_last=args.channels_last, device_pipeline=args.device_pipeline,
                               layers=layers, comm=args.comm, buckets=args.buckets)
    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    final_loss = float(loss.detach())
    if hasattr(step, "close"):
        step.close()
    if rank == 0:
        print(json.dumps({"metric": "train_images_per_sec", "value": imgs * world * args.steps / (ms * 1e-3),
                          "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                          "dtype": "f32", "data": "synthetic", "loss_impl": args.loss, "cuda_graph": bool(args.graph), "channels_last": bool(args.channels_last), "device_pipeline": bool(args.device_pipeline), "final_loss": final_loss,
                          "comm": args.comm if (args.graph and world > 1) else ("ddp" if world > 1 else None),
                          "config": {"workload": ("mono+stereo training step (BASELINE configs[3]): ResNet-50 depth + separate "
                                                  "ResNet-50 pose net, batch 8 per GPU, 320x1024, frame_ids [0,-1,1,'s'], "
                                                  "4 scales, Adam, fp32") if args.config3 else
                                                 ("mono training step: ResNet-18 depth + separate ResNet-18 pose net, "
                                                  "batch 12 per GPU, 192x640, frame_ids [0,-1,1], 4 scales, Adam, fp32"),
                                     "parallelism": f"ddp{world}"},
                          "clocks": clk.summary()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="loss", choices=["loss", "train_step"])
    ap.add_argument("--loss", default="fused", choices=["fused", "eager"])
    ap.add_argument("--channels-last", action="store_true", help="train_step: NHWC activations in the cuDNN networks")
    ap.add_argument("--device-pipeline", action="store_true",
                    help="train_step: start from uint8 frames on the device (md2_b200.pipeline) and end with md2_b200.metrics")
    ap.add_argument("--graph", action="store_true", help="train_step: replay the whole step as one CUDA graph")
    ap.add_argument("--comm", default="captured", choices=["captured", "eager"],
                    help="train_step --graph on N > 1 GPUs: bucket all-reduces inside the graph (overlapped with "
                         "backward) or eagerly between a forward+backward graph and an optimizer graph")
    ap.add_argument("--buckets", type=int, default=6, help="train_step --graph: gradient buckets")
    ap.add_argument("--no-train-step", action="store_true",
                    help="loss workload: skip the training-step leg (the `train_step` key of the JSON line)")
    ap.add_argument("--config3", action="store_true",
                    help="train_step: BASELINE configs[3] (ResNet-50, frame_ids [0,-1,1,'s'], 320x1024, batch 8 per GPU)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.workload == "train_step":
        run_train_step(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
