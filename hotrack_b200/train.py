"""The training step of the data-parallel path as one object: forward, loss, backward, the single
gradient all-reduce (when world_size > 1) and the flat Adam update -- optionally captured ONCE into a
CUDA graph and replayed, which removes the per-launch host cost of the ~400 kernels of a step
(shapes are static: B clouds of N points per rank).

The reference's equivalent is Trainer.update (network/trainer.py:278-302): zero_grad, model forward,
compute_loss, backward, optimizer.step, each a Python-dispatched op sequence.
"""
import os

import torch
import torch.distributed as dist

from .flat import FlatAdam, FlatParams


class TrainStep:
    def __init__(self, model, loss_fn, lr=1e-4, weight_decay=1e-4, graph=True):
        self.model, self.loss_fn = model, loss_fn
        from . import fused
        self._fused = fused
        self.weights = fused.WeightPlan()
        self.use_plan = os.environ.get("PN2_NO_WPLAN", "") == ""
        self.flat = FlatParams(model, tail=(1 << 18) if self.use_plan else 0)
        self.arena = fused.ZeroArena(buf=self.flat.tail) if self.use_plan else None  # zeroed by flat.zero_grad()
        self.flat.broadcast(0)
        self.opt = FlatAdam(self.flat, lr=lr, weight_decay=weight_decay)
        self.use_graph = graph
        self._graph = None
        self._static_in = None
        self._loss = None
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    @staticmethod
    def flat_device(model):
        return next(model.parameters()).device

    def _fwd_bwd(self, inputs):
        self.flat.zero_grad()
        f = self._fused
        saved = (f.SPARSE_GRAD_SINK, f.DIRECT_PARAM_GRADS, f.WGRAD_SIDE_STREAM)
        # whole-step backward owned by this object: gather gradients travel in row form, parameter gradients are
        # accumulated straight into the flat .grad buffer and the weight-gradient kernels run on a side stream beside
        # the backward chain -- all only for the duration of the step
        f.SPARSE_GRAD_SINK = f.DIRECT_PARAM_GRADS = True
        f.WGRAD_SIDE_STREAM = os.environ.get("PN2_WGRAD_STREAM", "1") != "0"
        if self.arena is not None:
            self.arena.begin()
            f.ACTIVE_ARENA = self.arena
        try:
            return self._fwd_bwd_inner(inputs)
        finally:
            f.join_wgrad()
            f.ACTIVE_ARENA = None
            f.SPARSE_GRAD_SINK, f.DIRECT_PARAM_GRADS, f.WGRAD_SIDE_STREAM = saved

    def _fwd_bwd_inner(self, inputs):
        if self.use_plan:
            self.weights.prepare()  # fp16 weight copies of every layer, on a side stream, while the step starts
            self._fused.ACTIVE_PLAN = self.weights
        try:
            loss = self.loss_fn(self.model(*inputs))
        finally:
            self._fused.ACTIVE_PLAN = None
            self.weights.finish()
        loss.backward()
        self._fused.join_wgrad()  # the weight-gradient stream rejoins before anything reads .grad
        return loss.detach()

    def _finish(self):
        self.opt.step(self.flat.allreduce_grads())

    def _eager(self, inputs):
        loss = self._fwd_bwd(inputs)
        self._finish()
        return loss

    def _capture(self, inputs):
        self._static_in = [t.clone() for t in inputs]
        # The warm-up steps must leave no trace: the first call applies ONE optimiser update, like the eager path and the
        # reference's Trainer.update.  Weights, Adam moments, the step counter and the BatchNorm buffers are restored
        # afterwards (the centring constants the warm-up leaves on the BatchNorm modules stay: they do not enter the
        # mathematics, and the captured step must be the steady-state one that reads them).
        snap = [t.clone() for t in self._state_tensors()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream: lazy inits, allocator, cuDNN/cuBLAS handles
            for _ in range(3):
                self._eager(self._static_in)
            for t, s0 in zip(self._state_tensors(), snap):
                t.copy_(s0)
        torch.cuda.current_stream().wait_stream(side)
        # world > 1: the NCCL all-reduce of the flat gradient and the Adam kernel are launched eagerly right behind the
        # replay (two launches).  Capturing them too (PN2_GRAPH_ALLREDUCE=1) is EXPERIMENTAL: with this torch / NCCL
        # pair the 2-GPU bench stopped making progress inside the captured exchange (round 2), so it is off by default.
        self._finish_in_graph = self.world == 1 or os.environ.get("PN2_GRAPH_ALLREDUCE", "0") == "1"
        self._graph = torch.cuda.CUDAGraph()
        # PN2_PRIO=1: capture the chain of dependent kernels on a HIGH-priority stream (graph kernel nodes keep the priority
        # of the stream they were captured on; the weight-gradient side stream has default priority, so the latency-critical
        # chain gets free SMs first).  Measured (B=32, N=4096): 3.65 -> 3.47 ms alone, but nothing on top of programmatic
        # dependent launch (3.43 ms without, 3.49 ms with), so it is off by default.
        cap = torch.cuda.Stream(priority=-1) if os.environ.get("PN2_PRIO", "0") == "1" else None
        with torch.cuda.graph(self._graph, stream=cap):
            self._loss = self._fwd_bwd(self._static_in)
            if self._finish_in_graph:
                self._finish()
        return 0

    def _state_tensors(self):
        return [self.flat.data, self.opt.exp_avg, self.opt.exp_avg_sq, self.opt.step_dev] + list(self.model.buffers())

    def __call__(self, *inputs):
        """inputs: CUDA tensors of the (static) shapes of the first call -> detached scalar loss tensor."""
        if not self.use_graph:
            return self._eager(inputs)
        if self._graph is None:
            self._capture(inputs)
        for dst, src in zip(self._static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        if not self._finish_in_graph:
            self._finish()  # NCCL all-reduce + Adam outside the captured region
        return self._loss


class GraphedForward:
    """Inference forward (eval mode, no_grad) of a module on static shapes, captured once and replayed: the
    per-frame tracking loop of the reference (network/models/track_network.py:159-224 calls the network once per
    frame, frame t+1 depending on frame t) is launch-latency bound at B = 1, so one graph launch per frame
    replaces ~150 kernel launches."""

    def __init__(self, model, graph=True):
        self.model, self.use_graph = model, graph
        self._graph = self._static_in = self._out = None

    @torch.no_grad()
    def __call__(self, *inputs):
        if not self.use_graph:
            return self.model(*inputs)
        if self._graph is None:
            self._static_in = [t.clone() for t in inputs]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self.model(*self._static_in)
            torch.cuda.current_stream().wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._out = self.model(*self._static_in)
        for dst, src in zip(self._static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._out
