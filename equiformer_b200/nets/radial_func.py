"""Radial profile MLP producing the per-edge DTP weights (drop-in for ``nets/radial_func.py``).

``Linear -> LayerNorm -> SiLU`` per hidden width, then ``Linear(no bias) + offset`` (``radial_func.py:9-50``);
``state_dict`` keys ``net.{0,1,3,4,6}.*`` and ``offset`` as in the reference.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class RadialProfile(nn.Module):
    def __init__(self, ch_list, use_layer_norm: bool = True, use_offset: bool = True):
        super().__init__()
        layers = []
        last = len(ch_list) - 1
        for i in range(1, len(ch_list)):
            is_last = i == last
            layers.append(nn.Linear(ch_list[i - 1], ch_list[i], bias=not (is_last and use_offset)))
            if is_last:
                break
            if use_layer_norm:
                layers.append(nn.LayerNorm(ch_list[i]))
            layers.append(nn.SiLU())
        self.net = nn.Sequential(*layers)
        self.offset = None
        if use_offset:
            self.offset = nn.Parameter(torch.zeros(ch_list[-1]))
            fan_in = ch_list[-2]
            bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
            nn.init.uniform_(self.offset, -bound, bound)

    def forward(self, f_in, add_offset: bool = True):
        """``add_offset=False`` returns the MLP output without ``offset``: the caller hands ``self.offset`` to the
        tensor-product kernel, which adds it while loading the weights (no extra pass over ``[E, weight_numel]``)."""
        from .. import ops
        out = f_in
        mods = list(self.net)    # same modules / state_dict keys as nn.Sequential; executed with fused kernels on CUDA
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            nxt2 = mods[i + 2] if i + 2 < len(mods) else None
            fuse_ln = (isinstance(nxt, nn.LayerNorm) and isinstance(nxt2, nn.SiLU) and nxt.elementwise_affine
                       and len(nxt.normalized_shape) == 1)
            if isinstance(m, nn.Linear) and fuse_ln:
                # Linear -> LayerNorm -> SiLU: GEMM without bias, then ONE kernel for bias + LayerNorm + SiLU
                out = ops.ln_silu(ops.linear_f32(out, m.weight, None), nxt.weight, nxt.bias, nxt.eps, bias=m.bias)
                i += 3
            elif isinstance(m, nn.Linear):
                out = ops.linear_f32(out, m.weight, m.bias)
                i += 1
            elif isinstance(m, nn.LayerNorm) and isinstance(nxt, nn.SiLU) and m.elementwise_affine \
                    and len(m.normalized_shape) == 1:
                out = ops.ln_silu(out, m.weight, m.bias, m.eps)      # LayerNorm + SiLU in one pass
                i += 2
            else:
                out = m(out)
                i += 1
        if self.offset is not None and add_offset:
            out = ops.add_bias(out, self.offset)
        return out
