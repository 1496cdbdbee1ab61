"""Data-parallel training of a whole flow: one process per GPU, batch sharded, gradients all-reduced.

What the reference's multi-GPU scripts do with nn.DataParallel (fastflow_cifar_multi_gpu.py:392-451,
train/experiment.py:219-277: scatter the batch, replicate the model every step, gather outputs on GPU 0,
reduce gradients to GPU 0, step there) done the torch.distributed way:

  * every rank holds a replica and its shard of the batch; the loss is the reference's
    `-(log p).sum() / B_global` (train/experiment.py:198-206) so that summed gradients equal the
    single-process gradient;
  * all gradients live in ONE flat fp32 buffer (each `p.grad` is a view into it), cut into buckets in
    reverse parameter order; a bucket's NCCL all-reduce is launched from the autograd hook of its last
    parameter, so the collective of the late layers overlaps the backward of the early ones
    (the CIFAR-10 flow has 22.9 M parameters = 91 MB of gradients);
  * the FInC gradient mask is applied inside the weight-gradient kernel (mask_in_backward=True), i.e.
    before the all-reduce -- masked entries are exactly zero on every rank;
  * ActNorm's data-dependent initialisation uses the statistics of the GLOBAL batch (flows.ActNorm
    all-reduces count / sum / sum of squares), so replicas start identical (the reference initialises on
    replica 0's shard: layers/actnorm.py:17-23);
  * sampling and likelihood evaluation use no collective;
  * `use_graph=True`: after `graph_warmup` eager steps the whole step -- forward, backward, the bucketed NCCL
    all-reduces launched from the autograd hooks, optimizer -- is captured into ONE CUDA graph and replayed (an
    eager step of the CIFAR-10 flow issues ~6000 kernel launches; the host, not the GPU, was the limit: 61 ms of
    CPU for 52 ms of GPU work).  Every rank must call step() the same number of times.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _native


class FlatAdam:
    """Adam over ONE flat fp32 parameter buffer and ONE flat gradient buffer: a single finc_adam_step_f32 launch
    (torch's fused multi-tensor Adam takes 16 launches and 1.1 ms for the 22.9 M parameters of the CIFAR-10 flow,
    ten times the time the 640 MB it moves need).  The step counter lives on the device, so the update is
    CUDA-graph capturable; `lr` is a launch argument (a trainer that replays a graph re-captures when it changes).
    Every parameter is a view into `flat_param`; after an update their autograd version counters are bumped so that
    caches keyed on them (prepared coupling weights, glue constants) see the change."""

    def __init__(self, params, flat_param, flat_grad, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.params, self.flat_param, self.flat_grad = list(params), flat_param, flat_grad
        self.param_groups = [{"params": self.params, "lr": lr, "betas": betas, "eps": eps}]   # lr schedulers edit this
        self.exp_avg = torch.zeros_like(flat_param)
        self.exp_avg_sq = torch.zeros_like(flat_param)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=flat_param.device)

    def step(self):
        g = self.param_groups[0]
        _native.adam_step_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_t,
                           lr=g["lr"], betas=g["betas"], eps=g["eps"])
        bump_versions(self.params)

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()

    def state_dict(self):
        g = self.param_groups[0]
        return {"flat_adam": True, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_t,
                "lr": g["lr"], "betas": g["betas"], "eps": g["eps"]}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_t.copy_(sd["step"])
        self.param_groups[0].update(lr=sd["lr"], betas=tuple(sd["betas"]), eps=sd["eps"])


def bump_versions(params):
    """mark parameters as modified in place (no kernel): a raw CUDA kernel or a CUDA-graph replay updates their
    memory without touching the autograd version counters that the layers' weight caches are keyed on"""
    for p in params:
        torch.autograd.graph.increment_version(p)


class FlowTrainer:
    def __init__(self, model, lr=1e-3, process_group=None, bucket_mb=25.0, optimizer=None, grad_clip_norm=None,
                 use_graph=False, graph_warmup=3, flat_adam=True):
        self.model = model
        self.flat_adam = flat_adam and optimizer is None
        self.use_graph, self.graph_warmup = use_graph, graph_warmup
        self._graph = None
        self._steps = 0
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) else 1
        if self.world > 1 and self.pg is None:
            self.pg = dist.group.WORLD
        self.grad_clip_norm = grad_clip_norm
        self.params = [p for p in model.parameters() if p.requires_grad]
        self._build_buckets(bucket_mb)
        self.flat_adam = self.flat_adam and self.params[0].is_cuda and all(p.dtype == torch.float32 for p in self.params)
        self.optimizer = optimizer if optimizer is not None else (self._make_flat_adam(lr) if self.flat_adam else self._make_adam(lr))
        if self.world > 1:
            self.broadcast_parameters()

    # ---- flat gradient buffer + buckets -----------------------------------------------------------
    def _build_buckets(self, bucket_mb):
        dev, dt = self.params[0].device, self.params[0].dtype
        # offsets in REVERSE parameter order: the last layers' gradients are ready first.  Every parameter starts on
        # a 64-byte boundary (the flat parameter buffer uses the same layout, and the kernels' vector loads / TMA
        # copies need 16-byte aligned weights)
        ALIGN = 16
        self.offsets = {}
        off = 0
        order = list(reversed(self.params))
        for p in order:
            self.offsets[p] = off
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        total = off
        self.flat_grad = torch.zeros(total, dtype=dt, device=dev)
        for p in order:
            off = self.offsets[p]
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
        limit = max(int(bucket_mb * 1024 * 1024 / self.flat_grad.element_size()), 1)
        self.buckets, start, members = [], 0, []
        for p in order:
            members.append(p)
            end = self.offsets[p] + p.numel()
            if end - start >= limit:
                self.buckets.append((start, end, members))
                start, members = end, []
        if members:
            self.buckets.append((start, total, members))
        self._bucket_of = {p: i for i, (_, _, ms) in enumerate(self.buckets) for p in ms}
        self._pending = [0] * len(self.buckets)
        self._works = []
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad_ready)

    def _make_flat_adam(self, lr):
        """parameters become views into ONE flat buffer laid out like the flat gradient buffer"""
        self.flat_param = torch.zeros_like(self.flat_grad)
        for p in self.params:
            o = self.offsets[p]
            view = self.flat_param[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        return FlatAdam(self.params, self.flat_param, self.flat_grad, lr=lr)

    def _make_adam(self, lr):
        cuda = self.params[0].is_cuda
        try:
            return torch.optim.Adam(self.params, lr=lr, fused=cuda, capturable=cuda and self.use_graph)
        except (TypeError, RuntimeError):
            return torch.optim.Adam(self.params, lr=lr)

    def _on_grad_ready(self, p):
        i = self._bucket_of[p]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            a, b, _ = self.buckets[i]
            self._works.append(dist.all_reduce(self.flat_grad[a:b], group=self.pg, async_op=True))

    def broadcast_parameters(self, src=0):
        """one-time sync: replicas start from rank `src`'s parameters and buffers"""
        for t in list(self.model.parameters()) + list(self.model.buffers()):
            dist.broadcast(t.data, src=src, group=self.pg)

    # ---- one optimisation step ----------------------------------------------------------------------
    def step(self, x):
        """x = this rank's shard.  Returns the global mean negative log-likelihood (a 0-dim tensor)."""
        if not (self.use_graph and x.is_cuda):
            return self._step_eager(x)
        if self._graph is not None and tuple(x.shape) == tuple(self._static_x.shape) and self._graph_lr == self._lr():
            self._static_x.copy_(x)
            self._graph.replay()
            bump_versions(self.params)         # the replay changed the weights without any autograd bookkeeping
            return self._static_loss
        self._steps += 1
        if self._steps <= self.graph_warmup:
            return self._step_eager(x)         # ActNorm initialisation, lazy kernel attributes, Adam state
        self._static_x = x.clone()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.pg)        # every rank captures the same sequence of collectives
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._static_loss = self._step_eager(self._static_x)
        self._graph, self._graph_lr = g, self._lr()
        g.replay()                             # capture executed nothing: run the captured step once for this call
        bump_versions(self.params)
        return self._static_loss

    def _lr(self):
        return tuple(g["lr"] for g in self.optimizer.param_groups)

    def _step_eager(self, x):
        self.flat_grad.zero_()
        for p in self.params:  # autograd must accumulate into the bucket views
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + self.offsets[p] * self.flat_grad.element_size():
                p.grad = self.flat_grad[self.offsets[p]:self.offsets[p] + p.numel()].view_as(p)
        self._pending = [len(ms) for _, _, ms in self.buckets]
        self._works = []
        _, logp = self.model(x)
        logp = torch.nan_to_num(logp, nan=0.0)           # train/experiment.py:201-204
        loss = -logp.sum() / (x.shape[0] * self.world)
        loss.backward()
        for w in self._works:
            w.wait()
        if self.world > 1 and any(n != 0 for n in self._pending):
            # parameters that received no gradient this step: reduce their buckets now (zeros stay zeros)
            for i, n in enumerate(self._pending):
                if n != 0:
                    a, b, _ = self.buckets[i]
                    dist.all_reduce(self.flat_grad[a:b], group=self.pg)
        if self.grad_clip_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.grad_clip_norm)
        self.optimizer.step()
        if not isinstance(self.optimizer, FlatAdam):
            bump_versions(self.params)         # torch's fused optimizers update in place WITHOUT bumping `_version`
        if self.world > 1:
            loss = loss.detach().clone()
            dist.all_reduce(loss, group=self.pg)
        return loss.detach()

    def close(self):
        """drop the captured graph (it references the process group's NCCL communicator: destroy the graph BEFORE
        the process group, or communicator teardown waits for it forever)"""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None
            self._static_x = self._static_loss = None

    # ---- checks ---------------------------------------------------------------------------------------
    def replica_max_diff(self):
        """max |parameter - rank 0's parameter| over all parameters and ranks (must be 0)"""
        if self.world == 1:
            return 0.0
        worst = torch.zeros((), dtype=torch.float64, device=self.params[0].device)
        for p in self.params:
            ref = p.detach().clone()
            dist.broadcast(ref, src=0, group=self.pg)
            worst = torch.maximum(worst, (p.detach() - ref).abs().max().double())
        dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=self.pg)
        return float(worst.item())
