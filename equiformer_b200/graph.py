"""Neighbour-list construction for synthetic inputs (stand-in for ``torch_cluster.radius_graph``).

Out of the kernel scope for this round: a plain torch brute force over same-graph pairs that
reproduces the ordering contract the hot path relies on - ``edge_index[1]`` (the centre atom, ``edge_dst``)
ascending, neighbours in index order, no self loops, ``d < r`` - as used at
``nets/graph_attention_transformer.py:866-867``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def radius_graph(pos: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None,
                 max_num_neighbors: int = 32, loop: bool = False, chunk: int = 4096) -> torch.Tensor:
    """Return ``edge_index [2, E]`` with row 0 = neighbour (source) and row 1 = centre (destination, ascending)."""
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
