"""Full mono training step (BASELINE.json configs[2]): ResNet-18 depth network + separate ResNet-18 pose
network on stock PyTorch / cuDNN, view-synthesis loss either fused (md2_b200) or eager PyTorch ops
(the oracle restatement of the reference's loss), DDP over N GPUs (NCCL all-reduce of the network
gradients only; the loss needs no collective).

Driven by bench.py (--workload train_step); not part of the product package.  The networks follow the
monodepth2 architecture (ResNet encoder, U-Net decoder with reflection-padded 3x3 convolutions + ELU and
nearest x2 up-sampling, sigmoid disparities at 4 scales; pose decoder on the last encoder feature with a
0.01 output scale) and are written here from that description, with random initial weights.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision


class Conv3x3(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.pad = nn.ReflectionPad2d(1)
        self.conv = nn.Conv2d(cin, cout, 3)

    def forward(self, x):
        return self.conv(self.pad(x))


class Encoder(nn.Module):
    def __init__(self, n_images=1, layers=18):
        super().__init__()
        net = {18: torchvision.models.resnet18, 50: torchvision.models.resnet50}[layers](weights=None)
        net.fc = nn.Identity()  # unused head: no parameters without gradients under DDP
        if n_images > 1:
            net.conv1 = nn.Conv2d(3 * n_images, 64, 7, 2, 3, bias=False)
        self.net = net
        self.ch = [64, 64, 128, 256, 512] if layers == 18 else [64, 256, 512, 1024, 2048]

    def forward(self, x):
        n = self.net
        x = (x - 0.45) / 0.225
        f0 = n.relu(n.bn1(n.conv1(x)))
        f1 = n.layer1(n.maxpool(f0))
        f2 = n.layer2(f1)
        f3 = n.layer3(f2)
        f4 = n.layer4(f3)
        return [f0, f1, f2, f3, f4]


class DepthDecoder(nn.Module):
    def __init__(self, enc_ch, scales=range(4)):
        super().__init__()
        dec = [16, 32, 64, 128, 256]
        self.scales = list(scales)
        self.up0, self.up1, self.head = nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict()
        for i in range(4, -1, -1):
            cin = enc_ch[-1] if i == 4 else dec[i + 1]
            self.up0[str(i)] = Conv3x3(cin, dec[i])
            self.up1[str(i)] = Conv3x3(dec[i] + (enc_ch[i - 1] if i > 0 else 0), dec[i])
        for s in self.scales:
            self.head[str(s)] = Conv3x3(dec[s], 1)

    def forward(self, feats):
        out = {}
        x = feats[-1]
        for i in range(4, -1, -1):
            x = F.elu(self.up0[str(i)](x))
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            if i > 0:
                x = torch.cat([x, feats[i - 1]], 1)
            x = F.elu(self.up1[str(i)](x))
            if i in self.scales:
                out[("disp", i)] = torch.sigmoid(self.head[str(i)](x))
        return out


class PoseDecoder(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.squeeze = nn.Conv2d(cin, 256, 1)
        self.c0 = nn.Conv2d(256, 256, 3, 1, 1)
        self.c1 = nn.Conv2d(256, 256, 3, 1, 1)
        self.c2 = nn.Conv2d(256, 12, 1)

    def forward(self, f):
        x = F.relu(self.squeeze(f))
        x = F.relu(self.c0(x))
        x = F.relu(self.c1(x))
        x = self.c2(x).mean(3).mean(2)
        x = 0.01 * x.view(-1, 2, 1, 6)
        return x[..., :3], x[..., 3:]


class MonoNets(nn.Module):
    """The four networks behind one module so that a single DDP wrapper covers them."""

    def __init__(self, layers=18, branch_streams=None):
        super().__init__()
        self.enc = Encoder(1, layers)
        self.dec = DepthDecoder(self.enc.ch)
        self.pose_enc = Encoder(2, layers)
        self.pose_dec = PoseDecoder(self.pose_enc.ch[-1])
        from md2_b200.trainer import BranchStreams
        self.branch_streams = (os.environ.get("MD2_BRANCH_STREAMS", "1") == "1") if branch_streams is None else bool(branch_streams)
        self._branches = BranchStreams()

    def _pose(self, inputs, f):
        pair = [inputs[("color_aug", f, 0)], inputs[("color_aug", 0, 0)]] if f < 0 else \
               [inputs[("color_aug", 0, 0)], inputs[("color_aug", f, 0)]]
        aa, tr = self.pose_dec(self.pose_enc(torch.cat(pair, 1))[-1])
        return aa[:, 0].contiguous(), tr[:, 0].contiguous()

    def forward(self, inputs, frame_ids):
        x0 = inputs[("color_aug", 0, 0)]
        mono = [f for f in frame_ids[1:] if f != "s"]   # the stereo baseline is data (inputs["stereo"]), no pose network
        if not (self.branch_streams and x0.is_cuda and mono):
            outputs = self.dec(self.enc(x0))
            for f in mono:
                outputs[("axisangle", f)], outputs[("translation", f)] = self._pose(inputs, f)
            return outputs
        # The depth network and the pose network do not depend on each other until the loss: md2_b200.trainer.
        # BranchStreams runs the pose branch on a side stream (a parallel branch of the captured graph, forward and
        # backward).  The pairs stay sequential inside that branch: they share the pose network's BatchNorm buffers.
        outputs, poses = self._branches(lambda: self.dec(self.enc(x0)),
                                        lambda: {f: self._pose(inputs, f) for f in mono})
        for f, (aa, tr) in poses.items():
            outputs[("axisangle", f)], outputs[("translation", f)] = aa, tr
        return outputs


def synthetic_batch(B, H, W, frame_ids, seed, device):
    import md2_b200.synthetic as syn
    inputs, _ = syn.make_batch(B, H, W, frame_ids, 4, seed, "iid", device=device, requires_grad=False)
    for f in frame_ids:
        inputs[("color_aug", f, 0)] = inputs[("color", f, 0)]
    return inputs


def make_step(loss_impl, B, H, W, frame_ids, device, ddp, graph=False, channels_last=False, device_pipeline=False,
              layers=18, comm="captured", buckets=6):
    """Returns (step_fn, images_per_step).  loss_impl: 'fused' (md2_b200) or 'eager' (PyTorch ops).
    graph: capture forward + loss + backward + Adam in one CUDA graph and replay it per step;
    the batch is copied into static input buffers on the device before every replay.
    device_pipeline: every step starts from decoded uint8 frames [3B,375,1242,3] resident on the device, builds the
    colour pyramid with md2_b200.pipeline and ends with md2_b200.metrics on a sparse 375x1242
    ground truth (N2), i.e. the whole md2_b200 surface around the stock networks."""
    from types import SimpleNamespace
    # the reference switches cuDNN's autotuner on (model_utility.py:327-328); shapes are static, so it pays once
    torch.backends.cudnn.benchmark = os.environ.get("MD2_CUDNN_BENCHMARK", "1") == "1"
    # depth and pose branches on two streams (md2_b200.trainer.BranchStreams) in the graphed step; the eager /
    # DistributedDataParallel arms keep the reference's single-stream order
    nets = MonoNets(layers, branch_streams=bool(graph) and os.environ.get("MD2_BRANCH_STREAMS", "1") == "1").to(device)
    if channels_last:   # NHWC activations for the cuDNN convolutions (precision-neutral); the loss inputs stay NCHW
        nets = nets.to(memory_format=torch.channels_last)
        if loss_impl == "fused" and os.environ.get("MD2_STOCK_PAD", "0") != "1":
            # ATen's reflection pad only knows NCHW: every Conv3x3 would pay layout copies around it, forward and
            # backward.  md2_b200.modules.ReflectionPad2d is the same operator (bit-identical) that keeps NHWC tensors NHWC.
            from md2_b200.modules import use_channels_last_padding, use_channels_last_pooling
            use_channels_last_padding(nets)
            if os.environ.get("MD2_STOCK_POOL", "0") != "1":   # same for the encoders' max-pool (gather backward)
                use_channels_last_pooling(nets)
    # With --graph the DDP wrapper is not used: md2_b200.trainer.GraphedTrainStep keeps the gradients in one flat,
    # bucketed buffer and overlaps the bucket all-reduces with the backward pass inside the captured step (comm =
    # "captured"), or runs them eagerly between a forward+backward graph and an optimizer graph (comm = "eager").
    model = nn.parallel.DistributedDataParallel(nets, device_ids=[device.index]) if (ddp and not graph) else nets
    # same Adam as model_tool/loader.py:107; the fused multi-tensor implementation makes one pass over the parameters
    # (None = torch's default, the multi-kernel foreach implementation)
    fused = True if (device.type == "cuda" and os.environ.get("MD2_FUSED_ADAM", "1") == "1") else None
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=graph, fused=fused)
    batches = [synthetic_batch(B, H, W, frame_ids, s, device) for s in range(2)]
    if channels_last:
        for b in batches:
            for f in frame_ids:
                b[("color_aug", f, 0)] = b[("color", f, 0)].contiguous(memory_format=torch.channels_last)
    cfg = SimpleNamespace(frame_ids=frame_ids, scales=range(4), height=H, width=W, min_depth=0.1, max_depth=100.0,
                          pose_type="separate", use_automasking=True, disp_smoothness=1e-3)
    if loss_impl == "fused":
        from md2_b200.compute import compute
        from md2_b200 import functional as F_
        comp = compute(cfg, device)

        def loss_fn(inputs, outputs):
            for f in frame_ids[1:]:
                if f != "s":
                    outputs[("c2c", f, 0)] = F_.param2matrix(outputs[("axisangle", f)],
                                                             outputs[("translation", f)], invert=(f < 0))
            # the auto-mask seed lives in a device tensor the step advances itself, so a replayed graph draws fresh noise
            comp.image2warping(inputs, outputs, None)
            return comp.compute_loss(inputs, outputs, None)["loss"]
    else:
        from oracle import oracle_torch as O  # eager PyTorch restatement of the reference loss (baseline leg)

        def loss_fn(inputs, outputs):
            Ts = [inputs["stereo"] if f == "s" else
                  O.pose_matrix(outputs[("axisangle", f)], outputs[("translation", f)], invert=(f < 0))
                  for f in frame_ids[1:]]
            return O.view_synthesis_loss(inputs[("color", 0, 0)], [inputs[("color", f, 0)] for f in frame_ids[1:]],
                                         [outputs[("disp", s)] for s in range(4)],
                                         [inputs[("color", 0, s)] for s in range(4)], inputs[("K", 0)],
                                         inputs[("inv_K", 0)], Ts,
                                         noise=[torch.randn(B, len(frame_ids) - 1, H, W, device=device)
                                                for _ in range(4)] if graph else None)["loss"]

    state = {"i": 0}

    if device_pipeline:
        if graph or loss_impl != "fused":
            raise NotImplementedError("--device-pipeline: fused loss, no --graph")
        import md2_b200.metrics as MM
        import md2_b200.pipeline as MP
        nf = len(frame_ids)
        pyr = MP.ColorPyramid(B * nf, 375, 1242, H, W, 4, device=device)
        gen = torch.Generator(device=device).manual_seed(1)
        frames = [torch.randint(0, 256, (B * nf, 375, 1242, 3), generator=gen, device=device, dtype=torch.uint8)
                  for _ in range(2)]
        flip = (torch.rand(B, generator=gen, device=device) > 0.5).to(torch.uint8).repeat(nf)   # one draw per sample
        gt = torch.zeros(B, 1, 375, 1242, device=device)
        hit = torch.rand(B, 1, 375, 1242, generator=gen, device=device) < 0.05
        gt[hit] = 1.0 + 79 * torch.rand(int(hit.sum()), generator=gen, device=device)
        static_in = {k: v for k, v in batches[0].items() if k[0] in ("K", "inv_K")}

        def step():
            levels = pyr(frames[state["i"] % 2], flip)            # [nf*B, 3, h, w] per level, frame-major
            state["i"] += 1
            inputs = dict(static_in)
            for i, f in enumerate(frame_ids):
                for s in range(4):
                    inputs[("color", f, s)] = levels[s][i * B:(i + 1) * B]
                c0 = inputs[("color", f, 0)]
                inputs[("color_aug", f, 0)] = c0.contiguous(memory_format=torch.channels_last) if channels_last else c0
            outputs = model(inputs, frame_ids)
            loss = loss_fn(inputs, outputs)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            state["metrics"] = MM.depth_metrics(outputs[("depth", 0, 0)], gt)   # stays on the device
            return loss

        return step, B

    def step():
        inputs = batches[state["i"] % len(batches)]
        state["i"] += 1
        outputs = model(inputs, frame_ids)
        loss = loss_fn(inputs, outputs)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    if not graph:
        return step, B

    from md2_b200.trainer import GraphedTrainStep

    def batch_process(inputs):
        return {"loss": loss_fn(inputs, model(inputs, frame_ids))}

    gstep = GraphedTrainStep(nets, batch_process, opt, batches[0], graph=True, comm=comm, buckets=buckets)

    def graphed_step():
        batch = batches[state["i"] % len(batches)]
        state["i"] += 1
        return gstep(batch)

    graphed_step.close = gstep.close   # release the graphs before the process group is torn down
    return graphed_step, B
