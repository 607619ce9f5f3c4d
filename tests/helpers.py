"""Shared synthetic inputs for the parity tests (seeded)."""
from __future__ import annotations

import torch


def molecules(sizes, seed=0, spread=1.6, species=(1, 6, 7, 8, 9), dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    pos = torch.cat([torch.randn(n, 3, generator=g, dtype=torch.float64) * spread for n in sizes]).to(dtype)
    batch = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(sizes)])
    z = torch.tensor(species)[torch.randint(0, len(species), (sum(sizes),), generator=g)]
    return pos, batch, z


def qm9_like_batch(n_graphs=128, seed=0, dtype=torch.float32):
    """128 molecules x ~18 atoms, positions ~ N(0, 1.6^2) with a 0.9 A rejection radius."""
    g = torch.Generator().manual_seed(seed)
    sizes = torch.clamp(torch.round(torch.randn(n_graphs, generator=g) * 3 + 18), 4, 29).long().tolist()
    probs = torch.tensor([0.51, 0.35, 0.06, 0.08, 0.002])
    species = torch.tensor([1, 6, 7, 8, 9])
    pos_all, z_all, batch_all = [], [], []
    for i, n in enumerate(sizes):
        pts = []
        while len(pts) < n:
            c = torch.randn(3, generator=g, dtype=torch.float64) * 1.6
            if all((c - p).norm() > 0.9 for p in pts):
                pts.append(c)
        pos_all.append(torch.stack(pts))
        z_all.append(species[torch.multinomial(probs, n, replacement=True, generator=g)])
        batch_all.append(torch.full((n,), i, dtype=torch.long))
    return torch.cat(pos_all).to(dtype), torch.cat(batch_all), torch.cat(z_all)


def aspirin_like(seed=0, dtype=torch.float32):
    """21 atoms (9 C, 4 O, 8 H) on a jittered lattice, min distance ~0.95 A."""
    g = torch.Generator().manual_seed(seed)
    grid = torch.stack(torch.meshgrid(torch.arange(3.), torch.arange(3.), torch.arange(3.), indexing="ij"), -1).reshape(-1, 3)
    sel = torch.randperm(27, generator=g)[:21]
    pos = grid[sel].double() * 1.45 + (torch.rand(21, 3, generator=g, dtype=torch.float64) - 0.5) * 0.4
    z = torch.tensor([6] * 9 + [8] * 4 + [1] * 8)
    return pos.to(dtype), torch.zeros(21, dtype=torch.long), z


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def closed_form_tensor(name: str, shape, mean: float, std: float) -> torch.Tensor:
    """A float32 tensor that both the golden generator and the tests can rebuild from (name, shape, mean, std) alone -
    ``mean + std * sqrt(2) * sin(phase(name) + 0.618 i)`` over the flat index - so that full-size models need only two
    numbers per large tensor in a fixture (tests/golden/make_reference_golden.py, "headline" case)."""
    import zlib
    n = 1
    for d in shape:
        n *= int(d)
    phase = (zlib.crc32(name.encode()) % 10007) * 0.001
    i = torch.arange(n, dtype=torch.float64)
    return (mean + std * (2.0 ** 0.5) * torch.sin(phase + 0.6180339887 * i)).float().reshape(tuple(int(d) for d in shape))
