"""CPU: Wigner-3j / spherical harmonics of the product vs the oracle and vs the anchors of SURVEY.md 8c."""
import math

import numpy as np
import pytest
import torch

from equiformer_b200 import o3
from equiformer_b200.o3.wigner import wigner_3j_np
from oracle import e3nn_ref as e3

TRIPLES = [(l1, l2, l3) for l1 in range(4) for l2 in range(4) for l3 in range(abs(l1 - l2), min(l1 + l2, 3) + 1)]


@pytest.mark.parametrize("l1,l2,l3", TRIPLES)
def test_wigner_product_equals_oracle(l1, l2, l3):
    a = wigner_3j_np(l1, l2, l3)
    b = e3.wigner_3j(l1, l2, l3).numpy()
    assert np.abs(a - b).max() < 1e-12
    assert abs(np.linalg.norm(a) - 1) < 1e-12
    k = np.einsum("ijk,ijl->kl", a, a)          # orthogonality: sum_ij C_ijk C_ijl = delta / (2 l3 + 1)
    assert np.abs(k - np.eye(2 * l3 + 1) / (2 * l3 + 1)).max() < 1e-12


def test_wigner_anchors():
    for l in range(4):
        d = 2 * l + 1
        assert np.allclose(wigner_3j_np(l, l, 0)[:, :, 0], np.eye(d) / math.sqrt(d))
        assert np.allclose(wigner_3j_np(l, 0, l)[:, 0, :], np.eye(d) / math.sqrt(d))
        assert np.allclose(wigner_3j_np(0, l, l)[0], np.eye(d) / math.sqrt(d))
    eps = np.zeros((3, 3, 3))
    for i, j, k in [(0, 1, 2), (1, 2, 0), (2, 0, 1)]:
        eps[i, j, k], eps[i, k, j] = 1, -1
    assert np.allclose(wigner_3j_np(1, 1, 1), eps / math.sqrt(6))   # 1 x 1 -> 1 is the cross product


def test_wigner_against_sympy():
    sympy = pytest.importorskip("sympy")
    from sympy.physics.wigner import clebsch_gordan
    from equiformer_b200.o3.wigner import su2_cg
    for j1, j2, j3 in [(1, 1, 2), (2, 1, 1), (2, 2, 2), (3, 2, 1), (3, 3, 0)]:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                m3 = m1 + m2
                if abs(m3) <= j3:
                    assert abs(su2_cg(j1, m1, j2, m2, j3, m3) - float(clebsch_gordan(j1, j2, j3, m1, m2, m3))) < 1e-12


def test_spherical_harmonics_product_equals_oracle_and_anchors():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(64, 3, generator=g, dtype=torch.float64)
    for norm in ("component", "norm", "integral"):
        a = o3.spherical_harmonics("1x0e+1x1e+1x2e+1x3e", v, True, norm)
        b = e3.spherical_harmonics([0, 1, 2, 3], v, True, norm)
        assert (a - b).abs().max() < 1e-12
    a = o3.spherical_harmonics([0, 1, 2, 3], v, False, "component")      # homogeneous polynomials when not normalised
    b = e3.spherical_harmonics([0, 1, 2, 3], v, False, "component")
    assert (a - b).abs().max() < 1e-10
    y = o3.spherical_harmonics("1x0e+1x1e+1x2e+1x3e", v, True, "component")
    for l, sl in zip(range(4), [slice(0, 1), slice(1, 4), slice(4, 9), slice(9, 16)]):
        assert torch.allclose(y[:, sl].pow(2).sum(-1), torch.full((64,), 2.0 * l + 1, dtype=torch.float64))
    u = torch.nn.functional.normalize(v, dim=-1)
    assert torch.allclose(y[:, 1:4], math.sqrt(3) * u)                     # l=1 is sqrt(3) (x, y, z)
    pole = o3.spherical_harmonics([2, 3], torch.tensor([[0.0, 1.0, 0.0]], dtype=torch.float64), True, "component")
    assert torch.allclose(pole[0, 2], torch.tensor(math.sqrt(5.0), dtype=torch.float64))   # y is the polar axis
    assert torch.allclose(pole[0, 5 + 3], torch.tensor(math.sqrt(7.0), dtype=torch.float64))


def _wigner_D_from_sh(l, R):
    """D_l(R) defined by Y_l(R x) = D_l(R) Y_l(x), solved by least squares on random directions (oracle SH)."""
    g = torch.Generator().manual_seed(l)
    x = torch.randn(200, 3, generator=g, dtype=torch.float64)
    a = e3.spherical_harmonics([l], x, True, "norm")
    b = e3.spherical_harmonics([l], x @ R.T, True, "norm")
    return torch.linalg.lstsq(a, b).solution.T


def test_wigner_3j_is_equivariant():
    g = torch.Generator().manual_seed(3)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    D = [_wigner_D_from_sh(l, q) for l in range(4)]
    for l in range(4):
        assert (D[l] @ D[l].T - torch.eye(2 * l + 1, dtype=torch.float64)).abs().max() < 1e-9
    for l1, l2, l3 in TRIPLES:
        c = torch.from_numpy(wigner_3j_np(l1, l2, l3).copy())
        rot = torch.einsum("ijk,ai,bj,ck->abc", c, D[l1], D[l2], D[l3])
        assert (rot - c).abs().max() < 1e-9, (l1, l2, l3)


def test_sh_is_twice_differentiable():
    v = torch.randn(5, 3, dtype=torch.float64).requires_grad_(True)
    f = lambda t: o3.spherical_harmonics("1x0e+1x1e+1x2e+1x3e", t, True, "component")
    assert torch.autograd.gradcheck(f, (v,), atol=1e-7)
    assert torch.autograd.gradgradcheck(f, (v,), atol=1e-6)


def test_spherical_harmonics_against_scipy():
    """Third-party anchor for the SH numerics: 'component'-normalised e3nn harmonics = sqrt(4 pi) x the standard real
    spherical harmonics (sqrt(2) (-1)^m Re / Im of scipy's complex Y_l^m, no extra phase) evaluated in the frame whose
    polar axis is e3nn's y:  (x', y', z') = (z, x, y).  Every (l, m) up to l = 3, oracle and product."""
    sp = pytest.importorskip("scipy.special")
    import numpy as np
    g = torch.Generator().manual_seed(0)
    v = torch.randn(300, 3, generator=g, dtype=torch.float64)
    v = v / v.norm(dim=1, keepdim=True)
    x, y, z = (v[:, i].numpy() for i in range(3))
    polar, azimuth = np.arccos(np.clip(y, -1, 1)), np.arctan2(x, z)

    def complex_sh(l, m):
        return sp.sph_harm_y(l, m, polar, azimuth) if hasattr(sp, "sph_harm_y") else sp.sph_harm(m, l, azimuth, polar)

    cols = []
    for l in range(4):
        for m in range(-l, l + 1):
            c = complex_sh(l, abs(m))
            real = c.real if m == 0 else math.sqrt(2) * (-1) ** m * (c.real if m > 0 else c.imag)
            cols.append(math.sqrt(4 * math.pi) * real)
    ref = torch.from_numpy(np.stack(cols, axis=1))
    assert (e3.spherical_harmonics([0, 1, 2, 3], v, True, "component") - ref).abs().max() < 1e-12
    assert (o3.spherical_harmonics("1x0e+1x1e+1x2e+1x3e", v, True, "component") - ref).abs().max() < 1e-12


def test_wigner_3j_equals_gaunt_tensor_of_the_harmonics():
    """Third-party anchor for the real Wigner 3j: for every even l1 + l2 + l3 (up to 3), the sphere average of
    Y^{l1}_a Y^{l2}_b Y^{l3}_c - harmonics pinned to scipy above, Gauss-Legendre x uniform quadrature, exact at this
    degree - equals  s sqrt((2l1+1)(2l2+1)(2l3+1)) (l1 l2 l3; 0 0 0) C_abc  with sympy's 3j symbol and ONE sign s = +-1
    per triple.  That fixes every component of the real-basis tensor; only that global sign per (l1, l2, l3) remains an
    e3nn convention that cannot be checked without e3nn (odd sums vanish here; (1,1,1) is anchored on epsilon / sqrt 6)."""
    sympy_wigner = pytest.importorskip("sympy.physics.wigner")
    ct, wt = np.polynomial.legendre.leggauss(16)
    phi = 2 * np.pi * np.arange(33) / 33
    CT, PH = np.meshgrid(ct, phi, indexing="ij")
    st = np.sqrt(1 - CT ** 2)
    pts = torch.from_numpy(np.stack([st * np.cos(PH), st * np.sin(PH), CT], -1).reshape(-1, 3))
    w = torch.from_numpy((wt[:, None] * np.full((1, 33), 2 * np.pi / 33)).reshape(-1)) / (4 * np.pi)
    Y = e3.spherical_harmonics([0, 1, 2, 3], pts, True, "component")
    sl = [slice(0, 1), slice(1, 4), slice(4, 9), slice(9, 16)]
    for l1, l2, l3 in TRIPLES:
        if (l1 + l2 + l3) % 2:
            continue
        gaunt = torch.einsum("n,na,nb,nc->abc", w, Y[:, sl[l1]], Y[:, sl[l2]], Y[:, sl[l3]])
        pref = math.sqrt((2 * l1 + 1) * (2 * l2 + 1) * (2 * l3 + 1)) * float(sympy_wigner.wigner_3j(l1, l2, l3, 0, 0, 0))
        for C in (e3.wigner_3j(l1, l2, l3), torch.from_numpy(wigner_3j_np(l1, l2, l3))):
            s = float((gaunt * C).sum() / (C * C).sum()) / pref
            assert abs(abs(s) - 1.0) < 1e-12, (l1, l2, l3, s)
            assert (gaunt - s * pref * C).abs().max() < 1e-12, (l1, l2, l3)
