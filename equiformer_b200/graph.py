"""Neighbour-list construction (stand-in for ``torch_cluster.radius_graph``).

The ordering contract the hot path relies on - ``edge_index[1]`` (the centre atom, ``edge_dst``) ascending, neighbours
in index order, no self loops, ``d < r``, at most ``max_num_neighbors`` per centre - as used at
``nets/graph_attention_transformer.py:866-867``.  On CUDA fp32 inputs two small kernels (count, fill) around one
prefix sum do it (``eqf_radius_graph_count / _fill``); the torch brute force below is the statement they are tested
against and the path for CPU tensors (host-side tests, oracle inputs).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def radius_graph(pos: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None,
                 max_num_neighbors: int = 32, loop: bool = False, chunk: int = 4096) -> torch.Tensor:
    """Return ``edge_index [2, E]`` with row 0 = neighbour (source) and row 1 = centre (destination, ascending)."""
    n = pos.shape[0]
    if pos.is_cuda and pos.dtype == torch.float32 and n > 0 and (batch is None or batch.dtype == torch.long):
        return _radius_graph_cuda(pos, r, batch, max_num_neighbors, loop)
    return radius_graph_torch(pos, r, batch, max_num_neighbors, loop, chunk)


def radius_graph_csr(pos: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None, max_num_neighbors: int = 32,
                     loop: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(edge_index [2, E], row_ptr [N + 1])``: the neighbour list and the CSR offsets of its (sorted) centres."""
    n = pos.shape[0]
    if pos.is_cuda and pos.dtype == torch.float32 and n > 0 and (batch is None or batch.dtype == torch.long):
        return _radius_graph_cuda(pos, r, batch, max_num_neighbors, loop, with_row_ptr=True)
    edge = radius_graph_torch(pos, r, batch, max_num_neighbors, loop)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=pos.device)
    torch.cumsum(torch.bincount(edge[1], minlength=n), 0, out=row_ptr[1:])
    return edge, row_ptr


def _radius_graph_cuda(pos, r, batch, max_num_neighbors, loop, with_row_ptr: bool = False):
    import ctypes

    from . import _lib
    lib = _lib.load()
    n = pos.shape[0]
    p = pos.detach().contiguous()
    b = batch.contiguous() if batch is not None else None
    cap = int(max_num_neighbors) if max_num_neighbors is not None else n
    stream = ctypes.c_void_p(torch.cuda.current_stream(pos.device).cuda_stream)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=pos.device)
    deg = torch.empty(n, dtype=torch.int64, device=pos.device)
    with torch.cuda.device(pos.device):
        rc = lib.eqf_radius_graph_count(p.data_ptr(), b.data_ptr() if b is not None else None, n, float(r) * float(r),
                                        1 if loop else 0, cap, deg.data_ptr(), stream)
        _lib.check(rc, "eqf_radius_graph_count")
        torch.cumsum(deg, 0, out=row_ptr[1:])
        n_edges = int(row_ptr[-1].item())          # the one host synchronisation of the neighbour search
        edge = torch.empty((2, n_edges), dtype=torch.int64, device=pos.device)
        if n_edges > 0:
            rc = lib.eqf_radius_graph_fill(p.data_ptr(), b.data_ptr() if b is not None else None, n, float(r) * float(r),
                                           1 if loop else 0, cap, row_ptr.data_ptr(), edge[0].data_ptr(), edge[1].data_ptr(),
                                           stream)
            _lib.check(rc, "eqf_radius_graph_fill")
    return (edge, row_ptr) if with_row_ptr else edge


def radius_graph_torch(pos: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None,
                       max_num_neighbors: int = 32, loop: bool = False, chunk: int = 4096) -> torch.Tensor:
    """The plain torch brute force over same-graph pairs (reference statement of the contract above)."""
    n = pos.shape[0]
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long, device=pos.device)
    pos_d = pos.detach()
    srcs, dsts = [], []
    for start in range(0, n, chunk):
        stop = min(start + chunk, n)
        d2 = (pos_d[start:stop, None, :] - pos_d[None, :, :]).pow(2).sum(-1)
        mask = (d2 < r * r) & (batch[start:stop, None] == batch[None, :])
        if not loop:
            idx = torch.arange(start, stop, device=pos.device)
            mask[idx - start, idx] = False
        if max_num_neighbors is not None and max_num_neighbors < n:
            rank = mask.cumsum(dim=1)
            mask &= rank <= max_num_neighbors
        dst, src = mask.nonzero(as_tuple=True)
        srcs.append(src)
        dsts.append(dst + start)
    return torch.stack([torch.cat(srcs), torch.cat(dsts)], dim=0)


def scatter_sum(x: torch.Tensor, index: torch.Tensor, dim: int = 0, dim_size: Optional[int] = None) -> torch.Tensor:
    """``torch_scatter.scatter(x, index, dim=0, dim_size=..)`` with the default ``sum`` reduce (node/graph level)."""
    if dim != 0:
        raise NotImplementedError("scatter_sum: only dim=0 is used by the reference")
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    out = torch.zeros((dim_size,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    return out.index_add(0, index, x)
