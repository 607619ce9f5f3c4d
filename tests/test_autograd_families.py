"""CPU: the hand-derived gradient families are closed under differentiation (first and second order).

``gradcheck`` / ``gradgradcheck`` in fp64 on tiny graphs, with the raw kernel calls swapped for the table-walking
stand-ins of tests/_emulation.py - this validates ops.py's Function wiring (which kernel computes which partial
derivative, argument order, saved tensors), exactly the part that the GPU numerics tests cannot localise.
"""
import pytest
import torch

from equiformer_b200 import ops
from tests._emulation import emulated_kernels


def _plan():
    from equiformer_b200.nets.graph_attention_transformer import DepthwiseTensorProduct
    return DepthwiseTensorProduct("3x0e+2x1e+2x2e", "1x0e+1x1e+1x2e", "3x0e+2x1e+2x2e", internal_weights=False, bias=False).tp.plan


@pytest.mark.parametrize("shared", [False, True])
def test_dtp_gradcheck_and_gradgradcheck(shared):
    plan = _plan()
    E = 3
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(E, 2 * l + 1, m, generator=g, dtype=torch.float64, requires_grad=True) for l, m in plan.in1_blocks]
    y = torch.randn(E, plan.d_y, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn((plan.weight_numel,) if shared else (E, plan.weight_numel), generator=g, dtype=torch.float64,
                    requires_grad=True)

    def f(y, w, *xs):
        return tuple(ops.DtpOut.apply(plan, y, w, *xs))

    with emulated_kernels():
        assert torch.autograd.gradcheck(f, (y, w, *xs), atol=1e-7)
        assert torch.autograd.gradgradcheck(f, (y, w, *xs), atol=1e-6)


def test_attention_family_gradcheck_and_gradgradcheck():
    n_nodes, E, H = 4, 9, 2
    dst = torch.tensor([0, 0, 0, 1, 1, 3, 3, 3, 3])
    src = torch.tensor([1, 2, 3, 0, 2, 0, 1, 2, 0])
    lay = ops.HeadLayout([1, 3], [4, 2], H)
    g = torch.Generator().manual_seed(1)
    z = torch.randn(E, H, generator=g, dtype=torch.float64, requires_grad=True)
    Vs = [torch.randn(E, d, c, generator=g, dtype=torch.float64, requires_grad=True) for d, c in zip(lay.ds, lay.Cs)]
    with emulated_kernels():
        graph = ops.Graph(src, dst, n_nodes)

        def f(z, *Vs):
            alpha = ops.segment_softmax(z, graph)
            return tuple(ops.attention_aggregate(lay, graph, alpha, Vs))

        assert torch.autograd.gradcheck(f, (z, *Vs), atol=1e-7)
        assert torch.autograd.gradgradcheck(f, (z, *Vs), atol=1e-6)

        def f_sum(*Vs):   # alpha=None: plain segment sum (EdgeDegreeEmbeddingNetwork)
            return tuple(ops.attention_aggregate(lay, graph, None, Vs))

        assert torch.autograd.gradcheck(f_sum, tuple(Vs), atol=1e-7)


def test_unsorted_edge_list_gives_same_layer_output():
    """GraphAttention accepts any edge order (sorted once by destination); result equals the pre-sorted call."""
    from equiformer_b200.nets import GraphAttention
    torch.manual_seed(0)
    irreps = "8x0e+4x1e+4x2e"
    ga = GraphAttention(irreps, "1x0e", "1x0e+1x1e+1x2e", irreps, [8, 16, 16], "2x0e+1x1e+1x2e", 4,
                        nonlinear_message=True, alpha_drop=0.0, proj_drop=0.0).double().eval()
    g = torch.Generator().manual_seed(1)
    n, E = 6, 20
    dst = torch.sort(torch.randint(0, n, (E,), generator=g)).values
    src = torch.randint(0, n, (E,), generator=g)
    x = torch.randn(n, 8 + 12 + 20, generator=g, dtype=torch.float64)
    sh = torch.randn(E, 9, generator=g, dtype=torch.float64)
    rbf = torch.randn(E, 8, generator=g, dtype=torch.float64)
    perm = torch.randperm(E, generator=g)
    with emulated_kernels():
        a = ga(x, None, src, dst, sh, rbf, None)
        b = ga(x, None, src[perm], dst[perm], sh[perm], rbf[perm], None)
    assert (a - b).abs().max() < 1e-12


def test_dtp_gathered_offset_gradcheck_with_dependent_inputs():
    """DtpOutGatheredOffset: first and second order, including the case that broke a nested-autograd backward once:
    the node tables are themselves functions of the edge harmonics (edge-degree embedding), so a partial derivative
    taken with autograd.grad on the saved tensors must not follow that ancestry."""
    plan = _plan()
    n_nodes, E = 4, 9
    dst = torch.tensor([0, 0, 0, 1, 1, 3, 3, 3, 3])
    src = torch.tensor([1, 2, 3, 0, 2, 0, 1, 2, 0])
    g = torch.Generator().manual_seed(4)
    nb = len(plan.in1_blocks)
    A0 = [torch.randn(n_nodes, 2 * l + 1, m, generator=g, dtype=torch.float64, requires_grad=True) for l, m in plan.in1_blocks]
    B0 = [torch.randn(n_nodes, 2 * l + 1, m, generator=g, dtype=torch.float64, requires_grad=True) for l, m in plan.in1_blocks]
    y = torch.randn(E, plan.d_y, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(E, plan.weight_numel, generator=g, dtype=torch.float64, requires_grad=True)
    off = torch.randn(plan.weight_numel, generator=g, dtype=torch.float64, requires_grad=True)

    with emulated_kernels():
        graph = ops.Graph(src, dst, n_nodes, check_sorted=False)

        def f(y, w, off, *AB):
            # node tables depend on y (like node features built from the edge-degree embedding)
            s = torch.zeros(n_nodes, dtype=y.dtype).index_add(0, dst, y.sum(1))
            AB = [t * (1.0 + 0.1 * s.view(-1, 1, 1)) for t in AB]
            return tuple(ops.DtpOutGatheredOffset.apply(plan, graph, nb, y, w * (1.0 + y[:, :1]), off, *AB))

        assert torch.autograd.gradcheck(f, (y, w, off, *A0, *B0), atol=1e-7)
        assert torch.autograd.gradgradcheck(f, (y, w, off, *A0, *B0), atol=1e-6)


def test_planar_linear_family_gradcheck_and_gradgradcheck():
    """``PlanarLinearFwd / Dgrad / Wgrad`` (all degrees of a linear in one grouped launch): first and second order in fp64,
    with a path constant != 1 and a weight offset != 0 for every path."""
    spec = ops.LinearSpec([(0, 0, 0, 4, 8, 0.5), (1, 1, 32, 4, 4, 1.0), (2, 2, 48, 8, 4, 1.7)], 80)
    assert spec.aligned()
    R = 3
    g = torch.Generator().manual_seed(0)
    w = torch.randn(80, generator=g, dtype=torch.float64, requires_grad=True)
    xs = [torch.randn(R, d, m, generator=g, dtype=torch.float64, requires_grad=True) for d, m in ((1, 4), (3, 4), (5, 8))]

    def f(w, *xs):
        return ops.PlanarLinearFwd.apply(spec, w, *xs)

    with emulated_kernels():
        outs = f(w, *xs)
        for p, x, o in zip(spec.paths, xs, outs):
            W = w[p[2]:p[2] + p[3] * p[4]].view(p[3], p[4])
            assert torch.allclose(o, p[5] * torch.einsum("rdu,uw->rdw", x, W), atol=1e-12)
        assert torch.autograd.gradcheck(f, (w, *xs), atol=1e-7)
        assert torch.autograd.gradgradcheck(f, (w, *xs), atol=1e-6)
        # the first-order backward takes the single-launch route when no graph is being built: same numbers
        cots = [torch.randn_like(o) for o in outs]
        a = torch.autograd.grad(f(w, *xs), (w, *xs), cots)
        b = torch.autograd.grad(f(w, *xs), (w, *xs), cots, create_graph=True)
        for u, v in zip(a, b):
            assert torch.allclose(u, v, atol=1e-12)


def test_linear_spec_matches_weight_views():
    """``TensorProduct._linear_spec`` (offsets, shapes, path constants of the grouped launch) against the per-path views and
    constants of ``linear_weight_blocks`` that the per-degree route uses."""
    from equiformer_b200.nets.tensor_product_rescale import LinearRS
    lin = LinearRS("8x0e+4x1e+4x2e", "4x0e+8x1e+4x2e", bias=True)
    spec = lin.tp._linear_spec()
    assert spec is not None and spec.w_numel == lin.tp.weight_numel
    w = lin.tp.weight.detach()
    blocks = lin.tp.linear_weight_blocks(w)
    assert len(blocks) == len(spec.paths)
    for (i1, _i2, io, W, c), p in zip(blocks, spec.paths):
        assert (i1, io) == (p[0], p[1]) and abs(c - p[5]) < 1e-12
        assert torch.equal(W.reshape(p[3], p[4]), w[p[2]:p[2] + p[3] * p[4]].view(p[3], p[4]))
