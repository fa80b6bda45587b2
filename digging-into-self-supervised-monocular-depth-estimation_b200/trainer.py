"""Caller-side glue of the training step: ``trainer.batch_process`` + ``zero_grad`` + ``backward`` +
``optimizer.step`` of /root/reference/model_train.py:54-96 replayed as ONE CUDA graph, with the gradient all-reduce of
data-parallel training overlapped with the backward pass inside that graph.

    step = GraphedTrainStep(models, batch_process, optimizer, example_inputs)      # models: setting.model (a dict)
    for inputs in loader:
        loss = step(inputs)                                                        # replay; loss is a device scalar

``batch_process(inputs) -> outputs`` is the reference's own method (model_train.py:90-96: forward_depth, forward_pose,
image2warping, compute_loss) or any callable built from its objects; the modules, the dict protocol and the optimizer
are the caller's.  What this class adds:

* **Flat gradient buffer.**  Every parameter's ``.grad`` is a view (same strides as the parameter, so channels-last
  weights keep autograd's layout contract) of one contiguous fp32 buffer, cut into ``buckets`` contiguous segments in
  reverse registration order - roughly the order in which backward produces them.
* **Overlapped all-reduce.**  A post-accumulate hook on every parameter counts its bucket down; the hook that completes a
  bucket enqueues ``all_reduce(bucket, AVG)`` asynchronously (NCCL over NVLink / NVSwitch on its own stream).  Under
  capture these collectives become parallel branches of the step's graph, so they overlap the remaining backward
  kernels exactly like DistributedDataParallel's reducer - without its per-step Python, bucket rebuilding or unused-
  parameter search, none of which a static step needs.  ``comm="eager"`` keeps the collectives out of the graphs (one
  graph for forward + backward, eager bucket all-reduces, one graph for the optimizer) for stacks whose NCCL cannot be
  captured.
* **Buffers.**  Floating-point module buffers (BatchNorm running statistics) are views of one flat tensor that is
  re-broadcast from rank 0 once per step (DistributedDataParallel(broadcast_buffers=True) does it with one broadcast
  per buffer at the start of the forward pass; here it is one collective, issued after the forward pass and
  overlapped with backward).
* **Learning-rate schedules survive capture.**  For optimizers with a ``capturable`` mode (Adam, AdamW, ...) the
  learning rate of every parameter group becomes a device tensor, which ``torch.optim.lr_scheduler`` updates in place
  (the reference steps a StepLR per epoch, model_train.py:81; model_tool/loader.py:107-108).  Other optimizers bake
  their hyper-parameters into the graph: build a new GraphedTrainStep after changing them.
* **Fused optimizer.**  An untouched ``torch.optim.Adam(params, lr)`` (model_tool/loader.py:107) runs as ~10 multi-tensor
  kernels per step; with ``fuse_optimizer=True`` (default) a fresh optimizer whose groups leave ``fused`` / ``foreach``
  at their defaults is switched to torch's fused implementation before capture (same update rule, one pass).
* **One graph.**  Static input tensors are refilled by ``copy_`` before each replay; the loss tensor is static.  The
  auto-mask noise of md2_b200.compute stays fresh across replays because its seed lives in a device tensor the captured
  step advances itself (functional.view_synthesis_loss(seed_tensor=...)).

``graph=False`` runs the same hooks and buckets eagerly (any device, any backend): that is what the CPU / gloo tests
exercise, and a drop-in for DistributedDataParallel when graphs are not wanted.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Union

import torch
import torch.distributed as dist
import torch.nn as nn


def _modules_of(models) -> List[nn.Module]:
    if isinstance(models, nn.Module):
        return [models]
    if isinstance(models, dict):
        return list(models.values())
    return list(models)


class FlatGradients:
    """Parameters' gradients as views of one flat buffer, bucketed, with hook-driven asynchronous all-reduce."""

    def __init__(self, params: Iterable[nn.Parameter], buckets: int = 6, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        # reverse registration order: the last layers' gradients arrive first and fill bucket 0
        order = list(reversed(self.params))
        n_b = max(1, min(int(buckets), len(order)))
        target = (total + n_b - 1) // n_b
        self.bucket_of: Dict[int, int] = {}
        self.bucket_range: List[List[int]] = []
        off, b, b_start, b_count = 0, 0, 0, 0
        self.bucket_params: List[int] = []
        for p in order:
            n = p.numel()
            # same strides as the parameter (dense, possibly permuted): autograd accumulates into it in place
            p.grad = self.flat[off:off + n].as_strided(p.size(), p.stride())
            self.bucket_of[id(p)] = b
            b_count += 1
            off += n
            if off - b_start >= target and b < n_b - 1:
                self.bucket_range.append([b_start, off])
                self.bucket_params.append(b_count)
                b, b_start, b_count = b + 1, off, 0
        self.bucket_range.append([b_start, off])
        self.bucket_params.append(b_count)
        self._pending = list(self.bucket_params)
        self._works = []
        self._streams_of = [set() for _ in self.bucket_params]   # CUDA streams that wrote gradients of a bucket
        self.hook_comm = True    # bucket all-reduces are issued from the hooks (False: reduce_now() does them)
        # The first step calibrates: parameters that never receive a gradient (an unused classifier head of a
        # torchvision encoder, a decoder branch of a scale that is not trained) are dropped from the bucket
        # counts - statically, once; the step is static, so DistributedDataParallel's per-step search is not needed.
        self._calibrating = True
        self._fired = set()
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    # ---- per-step protocol: begin() ... backward ... finish()
    def begin(self):
        self.flat.zero_()
        self._pending = list(self.bucket_params)
        self._works = []
        for st in self._streams_of:
            st.clear()

    def _hook(self, p):
        if self._calibrating:
            self._fired.add(id(p))
            return
        b = self.bucket_of[id(p)]
        self._pending[b] -= 1
        reduce_here = self.world > 1 and self.hook_comm
        if reduce_here and self.flat.is_cuda:
            # backward may run on several streams (BranchStreams): remember which ones wrote into this bucket
            self._streams_of[b].add(torch.cuda.current_stream())
        if self._pending[b] == 0 and reduce_here:
            if self.flat.is_cuda:
                # The collective is ordered after the CURRENT stream only.  Gradients of this bucket written on other
                # streams were enqueued earlier by this same autograd thread: an event recorded on each of them now
                # covers those writes.
                cur = torch.cuda.current_stream()
                for st in self._streams_of[b]:
                    if st != cur:
                        ev = torch.cuda.Event()
                        ev.record(st)
                        cur.wait_event(ev)
            lo, hi = self.bucket_range[b]
            self._works.append(dist.all_reduce(self.flat[lo:hi],
                                               op=dist.ReduceOp.AVG if self.flat.is_cuda else dist.ReduceOp.SUM,
                                               group=self.group, async_op=True))

    def finish(self):
        """Join the outstanding collectives (and average on backends without ReduceOp.AVG)."""
        for w in self._works:
            w.wait()
        self._works = []
        if self._calibrating:
            # end of the calibration step: fix the bucket counts, then reduce this step's gradients in one go
            self._calibrating = False
            self.unused = [p for p in self.params if id(p) not in self._fired]
            for p in self.unused:
                self.bucket_params[self.bucket_of[id(p)]] -= 1
            if self.hook_comm:
                self.reduce_now()
            return
        if self.world > 1 and not self.flat.is_cuda and self.hook_comm:
            self.flat.div_(self.world)
        if any(n != 0 for n in self._pending):
            raise RuntimeError("FlatGradients: some parameters received no gradient in this step (buckets "
                               f"{[i for i, n in enumerate(self._pending) if n]} incomplete); the step is not static")

    def reduce_now(self):
        """All-reduce every bucket here and now (comm='eager': between the two graphs)."""
        if self.world > 1:
            for lo, hi in self.bucket_range:
                dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.AVG if self.flat.is_cuda else dist.ReduceOp.SUM,
                                group=self.group)
            if not self.flat.is_cuda:
                self.flat.div_(self.world)

    def remove(self):
        for h in self._handles:
            h.remove()


def _record_stream(obj, stream):
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_stream(v, stream)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_stream(v, stream)


class BranchStreams:
    """Run independent parts of ``batch_process`` concurrently:  ``depth_out, pose_out = branches(depth_fn, pose_fn)``.

    The depth network and the pose network of the reference's step (compute.forward_depth / compute.forward_pose,
    model_train.py:92-93) do not depend on each other until image2warping.  The first callable runs on the current
    stream, every other one on its own side stream that is forked from the current stream before and joined to it
    after; tensors returned by the side callables (nested dicts / lists / tuples) are recorded on the current stream.
    Inside GraphedTrainStep's capture the branches become parallel branches of the step's CUDA graph, and because
    autograd runs every backward node on the stream of its forward, the backward pass forks the same way: the small
    kernels of the deep, low-resolution layers of one network fill the SMs the other leaves idle (ResNet-18
    configuration: 14.9 -> 12.8 ms per step).  Callables that share state - e.g. two passes through one network's
    BatchNorm buffers - belong in the SAME callable.  Without CUDA the callables simply run in order."""

    def __init__(self, enabled: bool = True):
        self.enabled = bool(enabled)
        self._streams = {}

    def __call__(self, main_fn, *side_fns):
        if not (self.enabled and side_fns and torch.cuda.is_available()):
            return [main_fn()] + [fn() for fn in side_fns]
        dev = torch.cuda.current_device()
        cur = torch.cuda.current_stream()
        pool = self._streams.setdefault(dev, [])
        while len(pool) < len(side_fns):
            pool.append(torch.cuda.Stream(device=dev))
        results = [None] * (1 + len(side_fns))
        for i, fn in enumerate(side_fns):
            pool[i].wait_stream(cur)
            with torch.cuda.stream(pool[i]):
                results[i + 1] = fn()
        results[0] = main_fn()
        for i in range(len(side_fns)):
            cur.wait_stream(pool[i])
            _record_stream(results[i + 1], cur)
        return results


class GraphedTrainStep:
    """One training step of the reference's trainer as a replayable CUDA graph (see the module docstring)."""

    def __init__(self, models: Union[nn.Module, Dict[str, nn.Module], Iterable[nn.Module]],
                 batch_process: Callable[[dict], dict], optimizer: torch.optim.Optimizer, example_inputs: dict, *,
                 graph: bool = True, comm: str = "captured", buckets: int = 6, broadcast_buffers: bool = True,
                 process_group=None, warmup: int = 3, loss_key: str = "loss", fuse_optimizer: bool = True):
        if comm not in ("captured", "eager"):
            raise ValueError("comm must be 'captured' or 'eager'")
        self.modules = _modules_of(models)
        self.batch_process, self.optimizer, self.loss_key = batch_process, optimizer, loss_key
        self.group = process_group
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.graph, self.comm = bool(graph), comm
        params, seen = [], set()
        for m in self.modules:
            for p in m.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        if self.distributed:  # identical initial weights on every rank (what DistributedDataParallel's ctor does)
            for p in params:
                dist.broadcast(p.data, 0, group=process_group)
        self.grads = FlatGradients(params, buckets, process_group)
        self.grads.hook_comm = not (self.graph and comm == "eager")
        # floating-point buffers (BatchNorm running statistics) become views of one flat tensor, so that the per-step
        # re-broadcast from rank 0 (DistributedDataParallel's broadcast_buffers) is ONE collective instead of ~120
        self._buffer_flat = None
        if broadcast_buffers and self.distributed:
            slots, seen_b = [], set()
            for m in self.modules:
                for sub in m.modules():
                    for name, b in sub._buffers.items():
                        if b is not None and b.is_floating_point() and id(b) not in seen_b:
                            seen_b.add(id(b))
                            slots.append((sub, name, b))
            if slots:
                flat = torch.empty(sum(b.numel() for _, _, b in slots), device=slots[0][2].device, dtype=slots[0][2].dtype)
                off = 0
                for sub, name, b in slots:
                    view = flat[off:off + b.numel()].view(b.shape)
                    view.copy_(b)
                    sub._buffers[name] = view
                    off += b.numel()
                self._buffer_flat = flat
        if self.graph:
            if not self.grads.flat.is_cuda:
                raise RuntimeError("GraphedTrainStep(graph=True) needs CUDA modules")
            for g in optimizer.param_groups:  # Adam & co. keep their step counter on the device when capturable
                if "capturable" in g:
                    g["capturable"] = True
                    # ... and read a TENSOR learning rate at replay time: torch's schedulers update such a tensor in
                    # place (model_train.py:81 steps a StepLR every epoch), whereas a Python float would be frozen
                    # into the captured kernels' arguments and the schedule silently ignored
                    if not torch.is_tensor(g["lr"]):
                        g["lr"] = torch.tensor(float(g["lr"]), dtype=torch.float32, device=self.grads.flat.device)
                    # A fresh Adam-family optimizer left at torch's default (foreach: ~10 multi-tensor kernels per
                    # step) is switched to its fused implementation - one pass over parameters and moments (measured:
                    # 21.5 -> 20.4 ms per step of the ResNet-18 configuration).  Same update rule; an optimizer that
                    # already has state, or whose caller chose foreach / fused explicitly, is left alone.
                    if (fuse_optimizer and "fused" in g and g.get("fused") is None and g.get("foreach") is None
                            and len(optimizer.state) == 0
                            and all(p.is_cuda and torch.is_floating_point(p) for p in g["params"])):
                        g["fused"] = True
        self.static_inputs = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_inputs.items()}
        self._loss = None
        self._graphs = None
        if self.graph:
            self._capture(max(int(warmup), 1))

    # ---- the step body (captured, or run eagerly when graph=False)
    def _forward_backward(self, inputs):
        self.grads.begin()
        outputs = self.batch_process(inputs)
        loss = outputs[self.loss_key]
        # The running statistics are final once the forward pass is done (training-mode BatchNorm normalises with the
        # batch statistics and only UPDATES the buffers), so their re-broadcast from rank 0 travels beside the
        # backward pass instead of in front of the next forward pass, where DistributedDataParallel puts it.
        if self._buffer_flat is not None:
            self.grads._works.append(dist.broadcast(self._buffer_flat, 0, group=self.group, async_op=True))
        loss.backward()
        return loss

    def _body(self, inputs):
        loss = self._forward_backward(inputs)
        self.grads.finish()
        if not self.grads.hook_comm:
            self.grads.reduce_now()
        self.optimizer.step()
        return loss.detach()

    def _capture(self, warmup):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # also creates the optimizer state, cuDNN plans and the NCCL communicator
                self._body(self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.comm == "captured" or not self.distributed:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._loss = self._body(self.static_inputs)
            self._graphs = (g,)
        else:
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self._loss = self._forward_backward(self.static_inputs).detach()
                self.grads.finish()
            with torch.cuda.graph(g2, pool=g1.pool()):
                self.optimizer.step()
            self._graphs = (g1, g2)

    def __call__(self, inputs: dict):
        if not self.graph:
            return self._body(inputs)
        for k, v in inputs.items():
            if torch.is_tensor(v):
                self.static_inputs[k].copy_(v, non_blocking=True)
        if len(self._graphs) == 1:
            self._graphs[0].replay()
        else:
            self._graphs[0].replay()
            self.grads.reduce_now()
            self._graphs[1].replay()
        return self._loss

    @property
    def flat_gradients(self) -> torch.Tensor:
        return self.grads.flat

    def close(self):
        """Release the captured graphs (they hold NCCL work when comm='captured': destroy_process_group() waits for
        them) and the gradient hooks.  Call before tearing the process group down."""
        if self._graphs is not None:
            torch.cuda.synchronize()
            for g in self._graphs:
                g.reset()
            self._graphs = None
        self.grads.remove()
