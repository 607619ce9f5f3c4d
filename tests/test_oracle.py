"""CPU: self-consistency of the oracle's third-party numerics (PARITY UNPINNED for those: e3nn cannot run here and the
reference ships no golden vectors - SURVEY.md section 4; what IS pinned to the reference's own code lives in
tests/test_reference_golden.py).

What can be pinned without the absent third-party packages: e3nn's normalize2mom recipe reproduces the constants the
survey recorded, the tensor product is O(3)-equivariant under Wigner matrices derived independently from the SH,
PyG softmax semantics (rows sum to one, empty segments, 1e-16), and F = -dE/dpos by finite differences.
"""
import math

import pytest
import torch

from oracle import e3nn_ref as e3
from oracle import equiformer_ref as R


def test_normalize2mom_constants():
    assert abs(e3.normalize2mom_const(torch.nn.functional.silu) - e3.NORMALIZE2MOM["silu"]) < 1e-12
    assert abs(e3.normalize2mom_const(torch.sigmoid) - e3.NORMALIZE2MOM["sigmoid"]) < 1e-12
    assert abs(e3.normalize2mom_const(lambda x: e3.smooth_leaky_relu(x, 0.2)) - e3.NORMALIZE2MOM["smooth_leaky_relu_0.2"]) < 1e-12
    assert abs(e3.NORMALIZE2MOM["silu"] - 1.67653) > 1e-3      # Monte-Carlo value, not the analytic one


def test_product_uses_the_same_constants():
    from equiformer_b200.math import normalize2mom
    from equiformer_b200.nets.graph_attention_transformer import SmoothLeakyReLU
    assert abs(normalize2mom(torch.nn.SiLU()).cst - e3.NORMALIZE2MOM["silu"]) < 1e-12
    assert abs(normalize2mom(torch.sigmoid).cst - e3.NORMALIZE2MOM["sigmoid"]) < 1e-12
    assert abs(normalize2mom(SmoothLeakyReLU(0.2)).cst - e3.NORMALIZE2MOM["smooth_leaky_relu_0.2"]) < 1e-12


def test_dtp_path_table_matches_survey_appendix_b():
    irr = e3.parse_irreps("128x0e+64x1e+32x2e")
    sh = e3.parse_irreps("1x0e+1x1e+1x2e")
    out, ins = R.dtp_instructions(irr, sh, irr)
    assert len(ins) == 15 and e3.irreps_dim(out) == 3136
    assert e3.simplify(out) == [(224, 0, 1), (384, 1, 1), (352, 2, 1)]
    # creation order -> sorted position (Appendix B): k=4 (1,1,0) lands right after k=0 in the 0e region
    assert [io for _, _, io, _ in ins][:5] == [0, 3, 9, 4, 1]
    l3 = e3.parse_irreps("128x0e+64x1e+64x2e+32x3e")
    out3, ins3 = R.dtp_instructions(l3, e3.parse_irreps("1x0e+1x1e+1x2e+1x3e"), l3)
    assert len(ins3) == 34 and e3.simplify(out3) == [(288, 0, 1), (576, 1, 1), (672, 2, 1), (576, 3, 1)]
    oc = e3.parse_irreps("256x0e+128x1e")
    outo, inso = R.dtp_instructions(oc, e3.parse_irreps("1x0e+1x1e"), oc)
    assert len(inso) == 5 and e3.simplify(outo) == [(384, 0, 1), (512, 1, 1)]


def _D(l, Rm):
    g = torch.Generator().manual_seed(l)
    x = torch.randn(200, 3, generator=g, dtype=torch.float64)
    return torch.linalg.lstsq(e3.spherical_harmonics([l], x, True, "norm"),
                              e3.spherical_harmonics([l], x @ Rm.T, True, "norm")).solution.T


def _block_rot(irreps, Ds, x):
    out, off = [], 0
    for mul, l, _ in irreps:
        d = 2 * l + 1
        out.append(torch.einsum("ij,zuj->zui", Ds[l], x[:, off:off + mul * d].reshape(-1, mul, d)).reshape(-1, mul * d))
        off += mul * d
    return torch.cat(out, dim=1)


def test_graph_attention_layer_is_equivariant():
    """Rotate inputs by D(R): outputs rotate by D(R) (pattern of nets/layer_norm.py:328-350)."""
    from equiformer_b200.nets import GraphAttention
    torch.manual_seed(0)
    irreps = "16x0e+8x1e+4x2e"
    ga = GraphAttention(irreps, "1x0e", "1x0e+1x1e+1x2e", irreps, [8, 16, 16], "4x0e+2x1e+1x2e", 4,
                        nonlinear_message=True, alpha_drop=0.0, proj_drop=0.0).double()
    params = {"ga." + k: v for k, v in R.cast_params(ga.state_dict(), torch.float64).items()}
    ir, sh_ir, head = e3.parse_irreps(irreps), e3.parse_irreps("1x0e+1x1e+1x2e"), e3.parse_irreps("4x0e+2x1e+1x2e")
    g = torch.Generator().manual_seed(1)
    pos = torch.randn(7, 3, generator=g, dtype=torch.float64)
    src, dst = R.radius_graph(pos, 10.0, torch.zeros(7, dtype=torch.long))
    x = torch.randn(7, e3.irreps_dim(ir), generator=g, dtype=torch.float64)
    rbf = torch.randn(src.numel(), 8, generator=g, dtype=torch.float64)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    Ds = [_D(l, q) for l in range(3)]

    def run(p, feats):
        vec = p[src] - p[dst]
        sh = e3.spherical_harmonics([0, 1, 2], vec, True, "component")
        return R.graph_attention(params, "ga", ir, sh_ir, head, 4, ir, True, feats, src, dst, sh, rbf)

    out = run(pos, x)
    out_rot = run(pos @ q.T, _block_rot(ir, Ds, x))
    assert (out_rot - _block_rot(ir, Ds, out)).abs().max() < 1e-9 * max(1.0, out.abs().max().item())


def test_pyg_softmax_semantics():
    z = torch.tensor([[1.0, -2.0], [3.0, 0.5], [0.0, 0.0], [100.0, -100.0]], dtype=torch.float64)
    idx = torch.tensor([0, 0, 2, 2])
    a = R.pyg_softmax(z, idx, 4)
    ref0 = torch.softmax(z[:2], dim=0)
    assert torch.allclose(a[:2], ref0) and torch.allclose(a[2:].sum(0), torch.ones(2, dtype=torch.float64))
    assert torch.isfinite(a).all()
    assert torch.allclose(R.scatter_sum(a, idx, 4)[[1, 3]], torch.zeros(2, 2, dtype=torch.float64))


def test_forces_are_minus_energy_gradient():
    """F = -dE/dpos (nets/graph_attention_transformer_md17.py:318-325) checked by central differences on the oracle."""
    from equiformer_b200.nets.graph_attention_transformer_md17 import GraphAttentionTransformerMD17
    torch.manual_seed(0)
    model = GraphAttentionTransformerMD17(irreps_in="64x0e", num_layers=2, max_radius=5.0, number_of_basis=16,
                                          basis_type="exp", irreps_head="32x0e+16x1e+8x2e", nonlinear_message=True,
                                          irreps_mlp_mid="384x0e+192x1e+96x2e", alpha_drop=0.0)
    params = R.cast_params(model.state_dict(), torch.float64)
    cfg = R.Config(basis_type="exp", number_of_basis=16, max_atom_type=64, qm9_atom_remap=False, num_layers=2)
    params = {k: v for k, v in params.items()}
    g = torch.Generator().manual_seed(2)
    pos = torch.randn(5, 3, generator=g, dtype=torch.float64) * 1.2
    batch = torch.zeros(5, dtype=torch.long)
    z = torch.tensor([6, 1, 8, 1, 6])
    _, f = R.energy_and_forces(params, cfg, pos, batch, z, 1)
    h = 1e-5
    for atom, axis in [(0, 0), (3, 2), (4, 1)]:
        p1, p2 = pos.clone(), pos.clone()
        p1[atom, axis] += h
        p2[atom, axis] -= h
        with torch.no_grad():
            e1 = R.model_forward(params, cfg, p1, batch, z, 1)
            e2 = R.model_forward(params, cfg, p2, batch, z, 1)
        fd = -(e1 - e2).item() / (2 * h)
        assert abs(fd - f[atom, axis].item()) < 1e-6 * max(1.0, abs(fd)), (atom, axis, fd, f[atom, axis].item())


def test_normalize2mom_constants_against_exact_second_moments():
    """The one e3nn number that is a Monte-Carlo estimate (1e6 samples, seed 0): how far can it be from what e3nn itself
    computed?  At most as far as it is from the exact value (adaptive quadrature of f(z)^2 against the normal density) -
    0.16 % for SiLU, 0.03 % for the sigmoid, 0.15 % for the smooth leaky ReLU - which bounds the one remaining
    scale ambiguity of the activations."""
    integrate = pytest.importorskip("scipy.integrate")

    def exact(f):
        val, _ = integrate.quad(lambda z: f(z) ** 2 * math.exp(-z * z / 2) / math.sqrt(2 * math.pi), -12, 12,
                                epsabs=1e-14, epsrel=1e-14)
        return val ** -0.5

    sig = lambda z: 1 / (1 + math.exp(-z))  # noqa: E731
    fs = {"silu": lambda z: z * sig(z), "sigmoid": sig,
          "smooth_leaky_relu_0.2": lambda z: 0.6 * z + 0.4 * z * (2 * sig(z) - 1)}
    for name, f in fs.items():
        assert abs(e3.NORMALIZE2MOM[name] - exact(f)) / exact(f) < 2e-3, name
