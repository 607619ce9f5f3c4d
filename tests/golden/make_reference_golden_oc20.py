"""Golden vector from the reference's OWN OC20 model file (``nets/graph_attention_transformer_oc20.py``), small
configuration with periodic boundary conditions -> ``tests/golden/reference_model_oc20_small.npz``.

Same method as ``make_reference_golden.py`` (whose stubs it reuses): the reference file is imported from where it lies and
run end to end in float64; the third-party calls it makes are served by stand-ins.  In addition to e3nn / torch_scatter /
PyG this file needs ``ocpmodels.common.registry`` (a decorator) and ``ocpmodels.common.utils``: ``radius_graph_pbc`` and
``get_pbc_distances`` (ocpmodels 0.0.3 @ d2aaaeb, absent from the image) are restated here - every (centre i, atom j,
image) of a frame with ``1e-4 < d^2 <= r^2``, images ``[-rep, rep]`` per lattice vector with ``rep = ceil(r * |b x c| /
volume)``, ordered by (i, j, image), the nearest ``max_neighbors`` kept; ``edge_index = (j, i)``; distance vector
``pos[j] - pos[i] + cell_offsets @ cell``.  What the fixture pins is the reference file's own wiring (tag embedding, OC20
statistics, head, PBC edge vectors); the third-party numerics under the stand-ins stay "parity unpinned".

Run in the build container only: ``python tests/golden/make_reference_golden_oc20.py``.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_reference_golden as G  # noqa: E402


def _stub_ocpmodels():
    from equiformer_b200.graph import radius_graph_pbc_torch
    reg = types.ModuleType("ocpmodels.common.registry")

    class _Registry:
        @staticmethod
        def register_model(name):
            return lambda cls: cls

    reg.registry = _Registry()
    utils = types.ModuleType("ocpmodels.common.utils")
    utils.conditional_grad = lambda dec: (lambda fn: fn)

    def radius_graph_pbc(data, radius, max_num_neighbors_threshold):
        edge_index, cell_offsets, _d2 = radius_graph_pbc_torch(data.pos, data.batch, data.cell, radius, max_num_neighbors_threshold)
        neighbors = torch.bincount(data.batch[edge_index[1]], minlength=data.cell.shape[0])
        return edge_index, cell_offsets, neighbors

    def get_pbc_distances(pos, edge_index, cell, cell_offsets, neighbors, return_offsets=False, return_distance_vec=False):
        row, col = edge_index
        vec = pos[row] - pos[col]
        cells = torch.repeat_interleave(cell, neighbors, dim=0)
        offsets = cell_offsets.to(pos.dtype).view(-1, 1, 3).bmm(cells.to(pos.dtype)).view(-1, 3)
        vec = vec + offsets
        dist = vec.norm(dim=-1)
        keep = torch.arange(len(dist))[dist != 0]
        out = {"edge_index": edge_index[:, keep], "distances": dist[keep]}
        if return_distance_vec:
            out["distance_vec"] = vec[keep]
        if return_offsets:
            out["offsets"] = offsets[keep]
        return out

    utils.radius_graph_pbc, utils.get_pbc_distances = radius_graph_pbc, get_pbc_distances
    common = types.ModuleType("ocpmodels.common")
    common.__path__ = []
    for name, mod in {"ocpmodels.common": common, "ocpmodels.common.registry": reg, "ocpmodels.common.utils": utils}.items():
        sys.modules[name] = mod


def main():
    G._stub_e3nn()
    G._stub_third_party()
    _stub_ocpmodels()
    oc = G._reference_module("graph_attention_transformer_oc20")
    cfg = dict(irreps_node_embedding="16x0e+8x1e", num_layers=2, irreps_node_attr="1x0e", use_node_attr=False,
               irreps_sh="1x0e+1x1e", max_radius=5.0, number_of_basis=16, fc_neurons=[16, 16], use_atom_edge_attr=False,
               irreps_atom_edge_attr="1x0e", irreps_feature="32x0e", irreps_head="8x0e+4x1e", num_heads=2,
               irreps_pre_attn="16x0e+8x1e", rescale_degree=False, nonlinear_message=True, irreps_mlp_mid="48x0e+24x1e",
               norm_layer="layer", alpha_drop=0.0, proj_drop=0.0, out_drop=0.0, drop_path_rate=0.0, otf_graph=True,
               use_pbc=True, max_neighbors=500)
    torch.manual_seed(21)
    model = oc.GraphAttentionTransformerOC20(None, None, 1, **cfg)
    gen = torch.Generator().manual_seed(99)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if prm.abs().max() == 0 or "bias" in name or "offset" in name:
                prm.add_(0.1 * torch.randn(prm.shape, generator=gen))
    torch.set_default_dtype(torch.float64)
    model = model.double().eval()
    n_atoms = [7, 9]
    batch = torch.repeat_interleave(torch.arange(2), torch.tensor(n_atoms))
    # two triclinic cells ~6-7 A across: with r = 5 every atom sees periodic images, including its own
    cell = torch.tensor([[[6.2, 0.0, 0.0], [0.7, 6.6, 0.0], [0.3, -0.5, 7.1]],
                         [[7.0, 0.4, 0.0], [0.0, 6.1, 0.6], [0.5, 0.0, 6.4]]], dtype=torch.float64)
    frac = torch.rand(sum(n_atoms), 3, generator=gen, dtype=torch.float64)
    pos = G._f32(torch.einsum("nk,nkd->nd", frac, cell[batch]))
    cell = G._f32(cell)
    z = torch.tensor([6, 1, 8, 29, 29, 78, 1, 13, 13, 8, 1, 6, 47, 47, 47, 7])
    tags = torch.tensor([2, 2, 2, 1, 0, 0, 2, 1, 1, 2, 2, 2, 0, 1, 0, 2])
    data = types.SimpleNamespace(pos=pos, batch=batch, cell=cell, atomic_numbers=z, tags=tags, natoms=torch.tensor(n_atoms))
    energy = model(data)
    model.zero_grad()
    (energy * torch.tensor([[1.0], [-0.7]])).sum().backward()
    out = {f"state/{k}": (v.detach().float().numpy() if v.is_floating_point() else v.numpy())
           for k, v in model.state_dict().items() if v is not None}
    for k, v in model.state_dict().items():
        if v is not None and v.is_floating_point():
            assert torch.equal(v.float().double(), v), k
    out.update({f"grad/{k}": p.grad.detach().numpy() for k, p in model.named_parameters() if p.grad is not None})
    out.update({"pos": pos.float().numpy(), "batch": batch.numpy(), "cell": cell.float().numpy(), "z": z.numpy(),
                "tags": tags.numpy(), "energy": energy.detach().numpy(), "edge_index": data.edge_index.numpy(),
                "cell_offsets": data.cell_offsets.numpy()})
    for k, v in cfg.items():
        if v is not None:
            out[f"cfg/{k}"] = np.asarray(v)
    path = os.path.join(HERE, "reference_model_oc20_small.npz")
    np.savez_compressed(path, **out)
    torch.set_default_dtype(torch.float32)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB; energy {energy.flatten().tolist()}; "
          f"{data.edge_index.shape[1]} edges")


if __name__ == "__main__":
    main()
