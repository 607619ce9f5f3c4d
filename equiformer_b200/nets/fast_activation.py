"""Scalar activations and gates (drop-in for the reference's ``nets/fast_activation.py``).

``Activation`` wraps every scalar non-linearity in ``normalize2mom`` (``fast_activation.py:25``) - the
Monte-Carlo second-moment constant is part of the forward arithmetic - and ``Gate`` splits a row into
``[scalars | gates | gated]`` (``:132-148``): scalars -> act, gates -> act, gated irreps multiplied channel-wise
by their gate (``ElementwiseTensorProduct`` against ``0e`` is a plain product).  Both also provide the planar
form used on edge tensors inside ``GraphAttention``.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from ..math import normalize2mom
from ..o3 import Irreps


def _act_parity(act) -> int:
    x = torch.linspace(0, 10, 256)
    a1, a2 = act(x), act(-x)
    if (a1 - a2).abs().max() < 1e-5:
        return 1
    if (a1 + a2).abs().max() < 1e-5:
        return -1
    return 0


class Activation(torch.nn.Module):
    """Apply one (second-moment normalised) scalar function per irreps entry; non-scalars pass through."""

    def __init__(self, irreps_in, acts):
        super().__init__()
        irreps_in = Irreps(irreps_in)
        if len(irreps_in) != len(acts):
            raise AssertionError((irreps_in, acts))
        acts = [normalize2mom(a) if a is not None else None for a in acts]
        out = []
        for (mul, ir), act in zip(irreps_in, acts):
            if act is None:
                out.append((mul, ir))
                continue
            if ir.l != 0:
                raise ValueError("Activation: cannot apply an activation function to a non-scalar input.")
            p_act = _act_parity(act)
            p_out = p_act if ir.p == -1 else ir.p
            if p_out == 0:
                raise ValueError("Activation: the parity is violated! The input scalar is odd but the "
                                 "activation is neither even nor odd.")
            out.append((mul, (0, p_out)))
        self.irreps_in = irreps_in
        self.irreps_out = Irreps(out)
        self.acts = torch.nn.ModuleList(acts)

    def extra_repr(self) -> str:
        return f"{self.irreps_in} -> {self.irreps_out}, "

    def forward(self, features, dim: int = -1):
        if len(self.acts) == 1:
            return self.acts[0](features)
        pieces, index = [], 0
        for (mul, ir), act in zip(self.irreps_in, self.acts):
            width = mul if act is not None else mul * ir.dim
            chunk = features.narrow(dim, index, width)
            pieces.append(act(chunk) if act is not None else chunk)
            index += mul * ir.dim
        if not pieces:
            return torch.zeros_like(features)
        return torch.cat(pieces, dim=dim) if len(pieces) > 1 else pieces[0]


class Gate(torch.nn.Module):
    def __init__(self, irreps_scalars, act_scalars, irreps_gates, act_gates, irreps_gated):
        super().__init__()
        irreps_scalars, irreps_gates, irreps_gated = Irreps(irreps_scalars), Irreps(irreps_gates), Irreps(irreps_gated)
        if len(irreps_gates) > 0 and irreps_gates.lmax > 0:
            raise ValueError(f"Gate scalars must be scalars, instead got irreps_gates = {irreps_gates}")
        if len(irreps_scalars) > 0 and irreps_scalars.lmax > 0:
            raise ValueError(f"Scalars must be scalars, instead got irreps_scalars = {irreps_scalars}")
        if irreps_gates.num_irreps != irreps_gated.num_irreps:
            raise ValueError(f"There are {irreps_gated.num_irreps} irreps in irreps_gated, but a different number "
                             f"({irreps_gates.num_irreps}) of gate scalars in irreps_gates")
        self.irreps_scalars, self.irreps_gates, self.irreps_gated = irreps_scalars, irreps_gates, irreps_gated
        self._irreps_in = (irreps_scalars + irreps_gates + irreps_gated).simplify()
        self.act_scalars = Activation(irreps_scalars, act_scalars)
        self.act_gates = Activation(irreps_gates, act_gates)
        self._irreps_out = self.act_scalars.irreps_out + irreps_gated

    def __repr__(self) -> str:
        return f"{self.__class__.__name__} ({self.irreps_in} -> {self.irreps_out})"

    @property
    def irreps_in(self):
        return self._irreps_in

    @property
    def irreps_out(self):
        return self._irreps_out

    def forward(self, features):
        ns, ng = self.irreps_scalars.dim, self.irreps_gates.dim
        scalars = self.act_scalars(features.narrow(-1, 0, ns))
        if ng == 0:
            return scalars
        gates = self.act_gates(features.narrow(-1, ns, ng))
        pieces, off, goff = [scalars], ns + ng, 0
        lead = features.shape[:-1]
        for mul, ir in self.irreps_gated:
            blk = features.narrow(-1, off, mul * ir.dim).reshape(*lead, mul, ir.dim)
            pieces.append((blk * gates.narrow(-1, goff, mul).unsqueeze(-1)).reshape(*lead, mul * ir.dim))
            off += mul * ir.dim
            goff += mul
        return torch.cat(pieces, dim=-1)

    def planar(self, blocks: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """Planar gate.  ``blocks`` follow ``irreps_in`` entries (first entry = scalars+gates merged as 0e)."""
        ns, ng = self.irreps_scalars.dim, self.irreps_gates.dim
        entries = list(self._irreps_in)
        if ng == 0:
            return [self.act_scalars(b) for b in blocks]
        has_scalar_entry = entries[0][1].l == 0 and entries[0][1].p == 1
        if not has_scalar_entry or entries[0][0] != ns + ng or len(entries) != 1 + len(self.irreps_gated):
            raise NotImplementedError("planar Gate expects [scalars+gates as one 0e entry | gated entries]")
        first = blocks[0]
        out = []
        if ns > 0:
            out.append(self.act_scalars(first.narrow(-1, 0, ns)))
        gates = self.act_gates(first.narrow(-1, ns, ng))  # [R, 1, ng]
        goff = 0
        for blk, (mul, _ir) in zip(blocks[1:], self.irreps_gated):
            out.append(blk * gates.narrow(-1, goff, mul))
            goff += mul
        return out
