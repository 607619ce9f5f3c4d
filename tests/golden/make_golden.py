"""Generate the committed golden fixtures from the fp64 oracle (run from the repo root: python tests/golden/make_golden.py).

PARITY UNPINNED: the reference has no golden vectors and its third-party numerics (e3nn 0.4.4, torch_scatter, PyG)
cannot be imported in this image, so these vectors come from ``oracle/`` (the CPU restatement), not from the reference
itself.  They pin the oracle against regressions and give the GPU tests size-independent fixed points.
Module weights are stored inside the fixtures (small irreps) so nothing depends on RNG streams of a torch version.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import e3nn_ref as e3  # noqa: E402
from oracle import equiformer_ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dtp_fixture(name, irreps, sh, E, seed):
    g = torch.Generator().manual_seed(seed)
    ir, shi = e3.parse_irreps(irreps), e3.parse_irreps(sh)
    out_ir, ins = R.dtp_instructions(ir, shi, ir)
    wn = sum(ir[i][0] for i, _, _, _ in ins)
    x = torch.randn(E, e3.irreps_dim(ir), generator=g).double()
    y = torch.randn(E, e3.irreps_dim(shi), generator=g).double()
    w = torch.randn(E, wn, generator=g).double()
    ws = torch.randn(wn, generator=g).double()
    out = e3.tensor_product(x, y, w, ir, shi, out_ir, ins, False)
    out_shared = e3.tensor_product(x, y, ws, ir, shi, out_ir, ins, True)
    np.savez_compressed(os.path.join(HERE, name), irreps=irreps, sh=sh, x=x.float().numpy(), y=y.float().numpy(),
                        w=w.float().numpy(), w_shared=ws.float().numpy(), out=out.numpy(), out_shared=out_shared.numpy())


def ga_fixture():
    from equiformer_b200.nets import GraphAttention
    torch.manual_seed(0)
    irreps, sh, head, H, nb = "16x0e+8x1e+4x2e", "1x0e+1x1e+1x2e", "4x0e+2x1e+1x2e", 4, 8
    ga = GraphAttention(irreps, "1x0e", sh, irreps, [nb, 16, 16], head, H, nonlinear_message=True, alpha_drop=0.0,
                        proj_drop=0.0)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in ga.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    sd = {k: v.detach().float() for k, v in ga.state_dict().items() if v.numel() > 0 and "output_mask" not in k}
    params = {"ga." + k: v.double() for k, v in sd.items()}
    pos = torch.randn(9, 3, generator=g).double() * 1.3
    src, dst = R.radius_graph(pos, 3.0, torch.zeros(9, dtype=torch.long))
    vec = (pos[src] - pos[dst]).float().double()
    edge_sh = e3.spherical_harmonics([0, 1, 2], vec, True, "component").float().double()
    x = torch.randn(9, 16 + 24 + 20, generator=g).double().requires_grad_(True)
    rbf = torch.randn(src.numel(), nb, generator=g).double()
    ir = e3.parse_irreps(irreps)
    out = R.graph_attention(params, "ga", ir, e3.parse_irreps(sh), e3.parse_irreps(head), H, ir, True, x, src, dst,
                            edge_sh, rbf)
    (gx,) = torch.autograd.grad(out.pow(2).sum(), x)
    np.savez_compressed(os.path.join(HERE, "graph_attention_small"), irreps=irreps, sh=sh, head=head, heads=H, nb=nb,
                        x=x.detach().float().numpy(), edge_src=src.numpy(), edge_dst=dst.numpy(),
                        edge_sh=edge_sh.float().numpy(), rbf=rbf.float().numpy(), out=out.detach().numpy(),
                        grad_x=gx.numpy(), **{"p:" + k: v.numpy() for k, v in sd.items()})


def md17_fixture():
    from equiformer_b200.nets.graph_attention_transformer_md17 import GraphAttentionTransformerMD17
    torch.manual_seed(0)
    cfg = dict(irreps_in="64x0e", irreps_node_embedding="16x0e+8x1e+4x2e", num_layers=2, irreps_sh="1x0e+1x1e+1x2e",
               max_radius=5.0, number_of_basis=8, basis_type="exp", fc_neurons=[16, 16], irreps_feature="32x0e",
               irreps_head="4x0e+2x1e+1x2e", num_heads=4, nonlinear_message=True, irreps_mlp_mid="48x0e+24x1e+12x2e",
               alpha_drop=0.0)
    model = GraphAttentionTransformerMD17(**cfg)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    sd = {k: v.detach().float() for k, v in model.state_dict().items() if v.numel() > 0 and "output_mask" not in k}
    params = {k: v.double() for k, v in sd.items()}
    sys.path.insert(0, ROOT)
    from tests.helpers import aspirin_like
    pos, batch, z = aspirin_like(seed=0)
    rcfg = R.Config(irreps_node_embedding=cfg["irreps_node_embedding"], irreps_head=cfg["irreps_head"],
                    irreps_mlp_mid=cfg["irreps_mlp_mid"], irreps_feature=cfg["irreps_feature"], num_layers=2,
                    number_of_basis=8, basis_type="exp", max_atom_type=64, qm9_atom_remap=False)
    energy, forces = R.energy_and_forces(params, rcfg, pos.double(), batch, z, 1)
    np.savez_compressed(os.path.join(HERE, "md17_small"), pos=pos.numpy(), z=z.numpy(), energy=energy.detach().numpy(),
                        forces=forces.detach().numpy(), **{"p:" + k: v.numpy() for k, v in sd.items()})


if __name__ == "__main__":
    dtp_fixture("dtp_qm9_l2", "128x0e+64x1e+32x2e", "1x0e+1x1e+1x2e", 5, 0)
    dtp_fixture("dtp_md17_l3", "128x0e+64x1e+64x2e+32x3e", "1x0e+1x1e+1x2e+1x3e", 3, 1)
    dtp_fixture("dtp_oc20_l1", "256x0e+128x1e", "1x0e+1x1e", 5, 2)
    ga_fixture()
    md17_fixture()
    print("golden fixtures written to", HERE)
