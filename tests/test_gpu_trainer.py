"""md2_b200.trainer.GraphedTrainStep on the B200: N replays of the captured step == N eager steps of the same modules
(/root/reference/model_train.py:54-96).  The second test wraps the REFERENCE's own objects - its
ResNet encoder / depth decoder / pose networks and its compute.forward_depth / forward_pose (staged in oracle/_ref) -
with this package's compute.image2warping / compute_loss, i.e. exactly the drop-in INTEGRATION.md describes."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _run(make, n_steps, graph):
    torch.manual_seed(0)
    models, batch_process, params, batches = make()
    # plain SGD with momentum: the update is linear in the gradient, so run-to-run differences stay at the level of
    # the kernels' floating-point atomics (Adam would turn the sign of a noise-level gradient into a full lr step)
    opt = torch.optim.SGD(params, 1e-3, momentum=0.9)
    from md2_b200.trainer import GraphedTrainStep
    warm = 3
    step = GraphedTrainStep(models, batch_process, opt, batches[0], graph=graph, warmup=warm)
    losses = []
    if not graph:
        for _ in range(warm):  # the graphed run spends its warm-up steps on the example batch
            step(batches[0])
    for i in range(n_steps):
        losses.append(float(step(batches[i % len(batches)])))
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().flatten() for p in params]).clone()
    return losses, flat


def _tiny_factory(branch_streams=False):
    import train_step as ts
    from md2_b200 import functional as F_
    from md2_b200.compute import compute
    B, H, W, fids = 2, 64, 96, [0, -1, 1]
    nets = ts.MonoNets(branch_streams=branch_streams).to(DEV)
    cfg = SimpleNamespace(frame_ids=fids, scales=range(4), height=H, width=W, min_depth=0.1, max_depth=100.0,
                          pose_type="separate", use_automasking=True, disp_smoothness=1e-3)
    comp = compute(cfg, DEV)
    comp.base_seed = 1234  # same auto-mask noise sequence in both runs

    def batch_process(inputs):
        outputs = nets(inputs, fids)
        for f in fids[1:]:
            outputs[("c2c", f, 0)] = F_.param2matrix(outputs[("axisangle", f)], outputs[("translation", f)], invert=(f < 0))
        comp.image2warping(inputs, outputs, None)
        return comp.compute_loss(inputs, outputs, None)
    batches = [ts.synthetic_batch(B, H, W, fids, s, torch.device(DEV)) for s in range(2)]
    return nets, batch_process, list(nets.parameters()), batches


def test_graph_replays_equal_eager_steps():
    le, pe = _run(_tiny_factory, 4, graph=False)
    lg, pg = _run(_tiny_factory, 4, graph=True)
    for a, b in zip(le, lg):
        assert a == pytest.approx(b, rel=2e-4), (le, lg)
    assert float((pe - pg).abs().max()) <= 2e-5


def test_branch_streams_in_the_graph_equal_the_single_stream_eager_step():
    """md2_b200.trainer.BranchStreams: depth and pose branches captured as parallel branches of the step's graph (forward
    and backward on two streams) give the losses and parameters of the plain single-stream eager loop."""
    le, pe = _run(_tiny_factory, 4, graph=False)
    lg, pg = _run(lambda: _tiny_factory(branch_streams=True), 4, graph=True)
    for a, b in zip(le, lg):
        assert a == pytest.approx(b, rel=2e-4), (le, lg)
    assert float((pe - pg).abs().max()) <= 2e-5
    # and eagerly, without a graph
    ls, ps = _run(lambda: _tiny_factory(branch_streams=True), 4, graph=False)
    for a, b in zip(le, ls):
        assert a == pytest.approx(b, rel=2e-4), (le, ls)
    assert float((pe - ps).abs().max()) <= 2e-5


def _reference_factory(two_streams=False):
    from oracle import ref_loader as RL
    import train_step as ts
    from md2_b200.compute import compute as FusedCompute
    R = RL.load()
    sys.path.insert(0, RL.REF)
    from model_layer import DepthDecoder, PoseDecoder, ResnetEncoder
    B, H, W, fids = 2, 64, 96, [0, -1, 1]
    opt = SimpleNamespace(frame_ids=fids, scales=range(4), height=H, width=W, min_depth=0.1, max_depth=100.0,
                          pose_type="separate", pose_frames=2, use_automasking=True, disp_smoothness=1e-3, batch=B,
                          num_layers=18, weight_init=False)
    model = {"encoder": ResnetEncoder(18, False)}
    model["decoder"] = DepthDecoder(model["encoder"].num_ch_enc, opt.scales)
    model["pose_encoder"] = ResnetEncoder(18, False, 2)
    model["pose_decoder"] = PoseDecoder(model["pose_encoder"].num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)
    model = {k: m.to(DEV).train() for k, m in model.items()}
    setting = SimpleNamespace(model=model)
    # The reference's param2matrix builds its matrices on the host and copies them over (model_layer/warp.py:50,111),
    # which a CUDA graph cannot capture: the symbol-level drop-in replaces it.
    import model_tool.processor as ref_processor
    from md2_b200 import functional as F_
    ref_processor.param2matrix = F_.param2matrix
    ref_compute = R.compute(opt, DEV)          # the reference's forward_depth / forward_pose ...
    fused = FusedCompute(opt, DEV)             # ... and this package's image2warping / compute_loss
    fused.base_seed = 99

    from md2_b200.trainer import BranchStreams
    branches = BranchStreams(enabled=two_streams)

    def batch_process(inputs):                 # model_train.py:90-96 with the two loss calls swapped
        if two_streams:                        # depth and pose networks on two streams
            depth_out, pose_out = branches(lambda: ref_compute.forward_depth(inputs, {}, setting)[1],
                                           lambda: ref_compute.forward_pose(inputs, {}, setting)[1])
            outputs = {**depth_out, **pose_out}
        else:
            outputs = {}
            inputs, outputs = ref_compute.forward_depth(inputs, outputs, setting)
            inputs, outputs = ref_compute.forward_pose(inputs, outputs, setting)
        inputs, outputs = fused.image2warping(inputs, outputs, setting)
        return fused.compute_loss(inputs, outputs, setting)
    batches = [ts.synthetic_batch(B, H, W, fids, s, torch.device(DEV)) for s in range(2)]
    params = [p for m in model.values() for p in m.parameters()]
    return model, batch_process, params, batches


def test_graphed_step_over_the_reference_trainer_objects():
    from oracle import ref_loader as RL
    if not RL.available():
        pytest.skip("oracle/_ref not staged (python oracle/stage_ref.py where /root/reference exists)")
    le, pe = _run(_reference_factory, 3, graph=False)
    lg, pg = _run(_reference_factory, 3, graph=True)
    for a, b in zip(le, lg):
        assert a == pytest.approx(b, rel=2e-4), (le, lg)
    assert float((pe - pg).abs().max()) <= 2e-5
    # the same objects with forward_depth / forward_pose on two streams inside the captured step
    l2, p2 = _run(lambda: _reference_factory(two_streams=True), 3, graph=True)
    for a, b in zip(le, l2):
        assert a == pytest.approx(b, rel=2e-4), (le, l2)
    assert float((pe - p2).abs().max()) <= 2e-5


def test_lr_schedule_reaches_the_captured_optimizer():
    """model_train.py:81 steps a StepLR between epochs (model_tool/loader.py:107-108: Adam + StepLR).  A captured Adam
    step must follow it: same parameters as the eager loop after the learning rate has dropped three times."""
    import copy
    from md2_b200.trainer import GraphedTrainStep
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ELU(), torch.nn.Linear(32, 4)).to(DEV)
    ref = copy.deepcopy(net)
    xs = [torch.randn(64, 16, device=DEV) for _ in range(6)]
    ys = [torch.randn(64, 4, device=DEV) for _ in range(6)]

    def process(m):
        return lambda inputs: {"loss": ((m(inputs["x"]) - inputs["y"]) ** 2).mean()}

    opt = torch.optim.Adam(net.parameters(), 1e-2)
    sched = torch.optim.lr_scheduler.StepLR(opt, 1, gamma=0.1)
    step = GraphedTrainStep(net, process(net), opt, {"x": xs[0], "y": ys[0]}, warmup=1)
    # the warm-up + capture steps have moved `net`: restart both from the same point, with fresh optimizer state
    net.load_state_dict(ref.state_dict())
    for st in opt.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()
    ref_opt = torch.optim.Adam(ref.parameters(), 1e-2)
    ref_sched = torch.optim.lr_scheduler.StepLR(ref_opt, 1, gamma=0.1)
    for i in range(6):
        step({"x": xs[i], "y": ys[i]})
        ref_opt.zero_grad(set_to_none=True)
        process(ref)({"x": xs[i], "y": ys[i]})["loss"].backward()
        ref_opt.step()
        if i % 2 == 1:                      # an "epoch" of two steps
            sched.step()
            ref_sched.step()
    assert torch.is_tensor(opt.param_groups[0]["lr"])
    assert float(opt.param_groups[0]["lr"]) == pytest.approx(1e-5, rel=1e-5)
    assert float(ref_opt.param_groups[0]["lr"]) == pytest.approx(1e-5, rel=1e-5)
    for a, b in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), float((a - b).abs().max())
    step.close()


def test_channels_last_padding_in_the_reference_decoder():
    """md2_b200.modules.use_channels_last_padding on the REFERENCE's DepthDecoder (model_layer/depth_decoder.py:36-50):
    same disparities and the same parameter gradients as with nn.ReflectionPad2d, tensors stay channels-last."""
    from oracle import ref_loader as RL
    if not RL.available():
        pytest.skip("oracle/_ref not staged (python oracle/stage_ref.py where /root/reference exists)")
    import copy
    import md2_b200.modules as M
    RL.load()
    sys.path.insert(0, RL.REF)
    from model_layer import DepthDecoder, ResnetEncoder
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False     # compare the operators, not cuDNN's TF32 algorithm choice per layout
    try:
        torch.manual_seed(1)
        enc = ResnetEncoder(18, False).to(DEV).to(memory_format=torch.channels_last).eval()
        dec = DepthDecoder(enc.num_ch_enc, range(4)).to(DEV).to(memory_format=torch.channels_last)
        dec2 = copy.deepcopy(dec)
        n = M.use_channels_last_padding(dec2)
        assert n == sum(isinstance(m, M.ReflectionPad2d) for m in dec2.modules()) and n >= 10
        assert not any(type(m) is torch.nn.ReflectionPad2d for m in dec2.modules())
        x = torch.rand(2, 3, 64, 96, device=DEV).contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            feats = enc(x)
        outs = []
        for d in (dec, dec2):
            o = d([f.clone() for f in feats])
            disp = [o[("disp", s)] for s in range(4)]
            sum((t * t).mean() for t in disp).backward()
            outs.append((disp, [p.grad.clone() for p in d.parameters()]))
        for a, b in zip(outs[0][0], outs[1][0]):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
        for a, b in zip(outs[0][1], outs[1][1]):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-7), float((a - b).abs().max())
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
