"""CUDA-graph replay of the forward+backward of a model step (launch-bound inner loop -> one graph launch).

A training step of the unfused pipeline issues a few thousand small kernels; at ~7 us of host time each the host,
not the GPU, bounds the step.  Everything after neighbour search is free of host synchronisation
(``GraphAttentionTransformer.forward_edges``), so it is captured once per input signature ``(atoms, edges, graphs)``
and replayed: inputs are copied into static buffers, ``graph.replay()`` runs forward, loss and backward, gradients
land in the flat bucket of :class:`equiformer_b200.parallel.FlatGradAllReduce`.  A new signature triggers a new
capture (cached), so variable-size batches still work - they just pay the capture when a size is first seen.
Neighbour search, the gradient all-reduce and the optimiser stay outside the graph.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch

from . import ops
from .graph import radius_graph_csr


class _Captured:
    __slots__ = ("graph", "pos", "batch", "z", "target", "src", "dst", "row_ptr", "csr", "loss")


class GraphedForwardBackward:
    def __init__(self, model: torch.nn.Module, loss_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                 bucket, max_radius: float, warmup: int = 3, max_cached: int = 8):
        self.model, self.loss_fn, self.bucket = model, loss_fn, bucket
        self.max_radius, self.warmup, self.max_cached = max_radius, warmup, max_cached
        self._cache: Dict[Tuple[int, int, int], _Captured] = {}
        self.captures = 0

    def _fwd_bwd(self, c: _Captured) -> torch.Tensor:
        # the source-sorted (CSC) view of the edge list is rebuilt inside the captured region from the static `src`
        # buffer (sort + scatter-add + scan, no host synchronisation): nothing of it is left on the host's critical path
        c.csr._src_perm = c.csr._src_row_ptr = None
        out = self.model.forward_edges(c.pos, c.batch, c.z, c.src, c.dst, graph=c.csr, n_graphs=c.target.shape[0])
        loss = self.loss_fn(out, c.target)
        # gradients as a list + one multi-tensor copy into the flat bucket (autograd's per-parameter accumulation into
        # the bucket views would be ~290 tiny `grad += g` launches per step)
        self.bucket.store(torch.autograd.grad(loss, self.bucket.params, allow_unused=True))
        return loss.detach()

    def _capture(self, pos, batch, z, target, src, dst, row_ptr) -> _Captured:
        c = _Captured()
        c.pos, c.batch, c.z, c.target = pos.clone(), batch.clone(), z.clone(), target.clone()
        c.src, c.dst, c.row_ptr = src.clone(), dst.clone(), row_ptr.clone()
        csr = ops.Graph.__new__(ops.Graph)
        csr.n_nodes, csr.n_edges, csr.perm = int(pos.shape[0]), int(src.numel()), None
        csr.src, csr.dst, csr.row_ptr = c.src, c.dst, c.row_ptr
        csr._src_perm = csr._src_row_ptr = None
        c.csr = csr
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._fwd_bwd(c)
        torch.cuda.current_stream().wait_stream(side)
        c.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(c.graph):
            c.loss = self._fwd_bwd(c)
        self.captures += 1
        return c

    def __call__(self, pos, batch, z, target) -> torch.Tensor:
        """Neighbour search (eager) + replay of the captured forward/backward; returns the (static) loss tensor."""
        edge, row_ptr = radius_graph_csr(pos, self.max_radius, batch, max_num_neighbors=1000)
        src, dst = edge[0], edge[1]
        key = (int(pos.shape[0]), int(src.numel()), int(target.shape[0]))
        c = self._cache.get(key)
        if c is None:
            if len(self._cache) >= self.max_cached:
                self._cache.pop(next(iter(self._cache)))
            c = self._capture(pos, batch, z, target, src, dst, row_ptr)
            self._cache[key] = c
        # inputs -> static buffers: one multi-tensor copy per dtype instead of seven small launches
        torch._foreach_copy_([c.pos, c.target], [pos, target])
        torch._foreach_copy_([c.batch, c.z, c.src, c.dst, c.row_ptr], [batch, z, src, dst, row_ptr])
        c.graph.replay()
        return c.loss
