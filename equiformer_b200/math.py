"""Second-moment normalisation of scalar activations (``e3nn.math.normalize2mom``).

The reference wraps every activation with it (``nets/fast_activation.py:25``), so the constants
are part of the forward arithmetic.  e3nn 0.4.4 estimates the second moment by Monte-Carlo on
``torch.randn(1_000_000, generator=Generator().manual_seed(0), dtype=float64)`` - not
analytically - and this file reproduces that recipe.
"""
from __future__ import annotations

import torch

_CACHE = {}


def moment(f, n: int, dtype=torch.float64, device="cpu") -> torch.Tensor:
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64).to(dtype=dtype, device=device)
    return f(z).pow(n).mean()


class normalize2mom(torch.nn.Module):
    """``x -> cst * f(x)`` with ``cst = E[f(z)^2]^-1/2``; identity wrap if ``|cst-1| < 1e-4``."""

    _is_id: bool

    def __init__(self, f, dtype=torch.float64, device="cpu"):
        super().__init__()
        key = None
        if isinstance(f, torch.nn.Module):
            key = (type(f).__name__, repr(f))
        elif hasattr(f, "__name__"):
            key = ("fn", getattr(f, "__module__", ""), f.__name__)
        if key is not None and key in _CACHE:
            cst = _CACHE[key]
        else:
            with torch.no_grad():
                cst = moment(f, 2, dtype=torch.float64, device="cpu").pow(-0.5).item()
            if key is not None:
                _CACHE[key] = cst
        self._is_id = abs(cst - 1) < 1e-4
        self.f = f
        self.cst = cst

    def forward(self, x):
        if self._is_id:
            return self.f(x)
        return self.f(x).mul(self.cst)

    def extra_repr(self) -> str:
        return f"cst={self.cst:.10f}"
