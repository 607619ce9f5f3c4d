"""Seeded synthetic inputs of the BASELINE.json shapes: used by ``bench.py``, ``smoke()`` and the
parity tests alike, so the bench and the tests run on the same geometry generators."""
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


def oc20_like_frames(n_frames=16, seed=0, mean_atoms=73, neighbours=50, radius=5.0, dtype=torch.float32):
    """OC20-IS2RE-like frames: ``n_atoms ~ clip(N(73, 20), 20, 200)`` uniformly placed (0.9 A rejection
    on a jittered lattice is unnecessary at this density) in a cubic cell whose density puts ``neighbours`` atoms within
    ``radius``; atomic numbers in 1..83, tags in {0, 1, 2}.  Returns (pos, batch, z, tags, cell) with ``cell`` the
    per-frame cubic cell edge ``[n_frames]`` (for the periodic neighbour list)."""
    g = torch.Generator().manual_seed(seed)
    sizes = torch.clamp(torch.round(torch.randn(n_frames, generator=g) * 20 + mean_atoms), 20, 200).long().tolist()
    rho = neighbours / (4.0 / 3.0 * 3.141592653589793 * radius ** 3)
    pos, batch, cells = [], [], []
    for i, n in enumerate(sizes):
        side = (n / rho) ** (1.0 / 3.0)
        pos.append(torch.rand(n, 3, generator=g, dtype=torch.float64) * side)
        batch.append(torch.full((n,), i, dtype=torch.long))
        cells.append(side)
    n_all = sum(sizes)
    z = torch.randint(1, 84, (n_all,), generator=g)
    tags = torch.randint(0, 3, (n_all,), generator=g)
    return torch.cat(pos).to(dtype), torch.cat(batch), z, tags, torch.tensor(cells, dtype=torch.float64)


def stress_cell(n_atoms=10000, seed=0, neighbours=50, radius=5.0, dtype=torch.float32):
    """One periodic cubic cell of ``n_atoms`` atoms (QM9 species) at the density that gives ``neighbours`` atoms within
    ``radius`` (BASELINE.json configs[4]).  Returns (pos, batch, z, cell_edge)."""
    g = torch.Generator().manual_seed(seed)
    rho = neighbours / (4.0 / 3.0 * 3.141592653589793 * radius ** 3)
    side = (n_atoms / rho) ** (1.0 / 3.0)
    pos = torch.rand(n_atoms, 3, generator=g, dtype=torch.float64) * side
    z = torch.tensor([1, 6, 7, 8, 9])[torch.randint(0, 5, (n_atoms,), generator=g)]
    return pos.to(dtype), torch.zeros(n_atoms, dtype=torch.long), z, side
