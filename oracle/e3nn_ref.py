"""ORACLE (test infrastructure, never on the product path): CPU restatement of the e3nn-0.4.4 numerics that the
reference's hot path calls.

PARITY UNPINNED: ``e3nn==0.4.4`` (``/root/reference/env/env_equiformer.yml:358``) is an un-vendored third-party
dependency, absent from this image and not installable (no network); the reference ships no tests or golden vectors
(SURVEY.md section 4).  This file therefore restates e3nn's *published* algorithms and is anchored on the
reference's call sites and on mathematical invariants (tests/test_oracle.py), not on outputs of the real library.
Third-party anchors that do exist (tests/test_o3.py): the spherical harmonics against scipy's, the SU(2) coefficients against
sympy's, the real Wigner 3j against the Gaunt tensors of the harmonics (all components, up to one sign per triple).
(What IS pinned to the reference's own code - its module and model files executed on top of these restatements - is
listed in ``oracle/equiformer_ref.py``.)

Call sites restated:
  * ``o3.TensorProduct(..., path_normalization='none')``  - nets/tensor_product_rescale.py:33-37
  * ``o3.spherical_harmonics(l, x, normalize=True, normalization='component')`` - nets/graph_attention_transformer.py:869-870
  * ``e3nn.math.normalize2mom``                            - nets/fast_activation.py:25
  * ``o3.Irreps`` string grammar                           - everywhere

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs import this.
Deliberately shares no code with ``equiformer_b200``: Wigner symbols come from e3nn's own Racah-sum form, spherical
harmonics from e3nn's generated closed-form polynomials (the product uses a coupling recurrence instead).
"""
from __future__ import annotations

import functools
import math
from fractions import Fraction
from math import factorial
from typing import List, Sequence, Tuple

import torch

# ----------------------------------------------------------------------------------------------------------------
# Irreps strings


def parse_irreps(s) -> List[Tuple[int, int, int]]:
    """'128x0e+64x1e' -> [(128, 0, +1), (64, 1, +1)]  (mul, l, parity)."""
    if not isinstance(s, str):
        s = str(s)
    out = []
    for chunk in s.split("+"):
        chunk = chunk.strip()
        if not chunk:
            continue
        mul, ir = chunk.split("x") if "x" in chunk else ("1", chunk)
        out.append((int(mul), int(ir[:-1]), 1 if ir[-1] == "e" else -1))
    return out


def irreps_dim(irreps) -> int:
    return sum(mul * (2 * l + 1) for mul, l, _ in irreps)


def irreps_slices(irreps) -> List[slice]:
    out, start = [], 0
    for mul, l, _ in irreps:
        out.append(slice(start, start + mul * (2 * l + 1)))
        start += mul * (2 * l + 1)
    return out


def simplify(irreps):
    out = []
    for mul, l, p in irreps:
        if out and out[-1][1:] == (l, p):
            out[-1] = (out[-1][0] + mul, l, p)
        elif mul > 0:
            out.append((mul, l, p))
    return out


def product_irreps(l1, p1, l2, p2):
    return [(l, p1 * p2) for l in range(abs(l1 - l2), l1 + l2 + 1)]


# ----------------------------------------------------------------------------------------------------------------
# Wigner 3j in e3nn's real basis (e3nn/o3/_wigner.py: _su2_clebsch_gordan_coeff, change_basis_real_to_complex,
# _so3_clebsch_gordan; wigner_3j == that tensor, Frobenius-normalised)


def _su2_cg_coeff(j1, m1, j2, m2, j3, m3) -> float:
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))

    def f(n):
        return factorial(round(n))

    c = ((2.0 * j3 + 1.0) * Fraction(
        f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
        f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2))) ** 0.5
    s = 0
    for v in range(vmin, vmax + 1):
        s += (-1) ** int(v + j2 + m2) * Fraction(
            f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
            f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3))
    return float(c * s)


def _su2_cg(j1, j2, j3) -> torch.Tensor:
    mat = torch.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1), dtype=torch.float64)
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l) -> torch.Tensor:
    q = torch.zeros((2 * l + 1, 2 * l + 1), dtype=torch.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / 2 ** 0.5
        q[l + m, l - abs(m)] = -1j / 2 ** 0.5
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / 2 ** 0.5
        q[l + m, l - abs(m)] = 1j * (-1) ** m / 2 ** 0.5
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def wigner_3j(l1: int, l2: int, l3: int) -> torch.Tensor:
    q1, q2, q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    c = _su2_cg(l1, l2, l3).to(torch.complex128)
    c = torch.einsum("ij,kl,mn,ikn->jlm", q1, q2, torch.conj(q3.T), c)
    assert torch.all(torch.abs(torch.imag(c)) < 1e-5)
    c = torch.real(c)
    return c / torch.linalg.norm(c)


# ----------------------------------------------------------------------------------------------------------------
# spherical harmonics: e3nn's generated closed forms (e3nn/o3/_spherical_harmonics.py, 'component' pre-factors
# pulled out), y = polar axis


def _sh_norm(l: int, x, y, z):
    """'norm'-normalised real SH of degree l for (already normalised or raw) coordinates; homogeneous of degree l."""
    if l == 0:
        return torch.ones_like(x).unsqueeze(-1)
    if l == 1:
        return torch.stack([x, y, z], dim=-1)
    x2, y2, z2 = x * x, y * y, z * z
    if l == 2:
        s3 = math.sqrt(3.0)
        return torch.stack([s3 * x * z, s3 * x * y, y2 - 0.5 * (x2 + z2), s3 * y * z, 0.5 * s3 * (z2 - x2)], dim=-1)
    if l == 3:
        # e3nn: sh_3_* expressed through the degree-2 polynomials; divided here by sqrt(7) ('component' -> 'norm')
        s15 = math.sqrt(15.0)
        sh20 = s15 * x * z
        sh24 = 0.5 * s15 * (z2 - x2)
        x2z2 = x2 + z2
        c = 1.0 / math.sqrt(7.0)
        return c * torch.stack([
            (1 / 6) * math.sqrt(42) * (sh20 * z + sh24 * x),
            math.sqrt(7) * sh20 * y,
            (1 / 8) * math.sqrt(168) * (4 * y2 - x2z2) * x,
            0.5 * math.sqrt(7) * y * (2 * y2 - 3 * x2z2),
            (1 / 8) * math.sqrt(168) * z * (4 * y2 - x2z2),
            math.sqrt(7) * sh24 * y,
            (1 / 6) * math.sqrt(42) * (sh24 * z - sh20 * x)], dim=-1)
    raise NotImplementedError("oracle spherical harmonics are restated up to l = 3 (Equiformer uses Lmax <= 3)")


def spherical_harmonics(ls: Sequence[int], vec: torch.Tensor, normalize: bool = True,
                        normalization: str = "component") -> torch.Tensor:
    if normalize:
        vec = torch.nn.functional.normalize(vec, dim=-1)
    x, y, z = vec[..., 0], vec[..., 1], vec[..., 2]
    out = []
    for l in ls:
        sh = _sh_norm(l, x, y, z)
        if normalization == "component":
            sh = sh * math.sqrt(2 * l + 1)
        elif normalization == "integral":
            sh = sh * math.sqrt((2 * l + 1) / (4 * math.pi))
        out.append(sh)
    return torch.cat(out, dim=-1)


# ----------------------------------------------------------------------------------------------------------------
# normalize2mom (e3nn/math/_normalize_activation.py): Monte-Carlo second moment on a fixed-seed normal sample

# evaluated with torch 2.11 on CPU; recomputed by tests/test_oracle.py
NORMALIZE2MOM = {"silu": 1.6791767923989418, "sigmoid": 1.8467055342154763, "smooth_leaky_relu_0.2": 1.531320475574866}


def normalize2mom_const(f) -> float:
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    return f(z).pow(2).mean().pow(-0.5).item()


def smooth_leaky_relu(x, alpha: float = 0.2):
    """nets/graph_attention_transformer.py:54-63"""
    return ((1 + alpha) / 2) * x + ((1 - alpha) / 2) * x * (2 * torch.sigmoid(x) - 1)


# ----------------------------------------------------------------------------------------------------------------
# TensorProduct forward as e3nn's code generator emits it: one einsum per instruction, outputs concatenated.
# irrep_normalization='component', path_normalization='none' -> path_weight = sqrt(2 l_out + 1).


def tensor_product(x: torch.Tensor, y: torch.Tensor, weight: torch.Tensor, irreps_in1, irreps_in2, irreps_out,
                   instructions: Sequence[Tuple[int, int, int, str]], shared_weights: bool) -> torch.Tensor:
    sl1, sl2 = irreps_slices(irreps_in1), irreps_slices(irreps_in2)
    z = x.shape[0]
    outs = [None] * len(irreps_out)
    woff = 0
    for i1, i2, io, mode in instructions:
        mul1, l1, _ = irreps_in1[i1]
        mul2, l2, _ = irreps_in2[i2]
        mulo, lo, _ = irreps_out[io]
        x1 = x[:, sl1[i1]].reshape(z, mul1, 2 * l1 + 1)
        x2 = y[:, sl2[i2]].reshape(z, mul2, 2 * l2 + 1)
        w3j = wigner_3j(l1, l2, lo).to(device=x.device, dtype=x.dtype)    # device-following: bench.py's reference-gpu arm
        pw = math.sqrt(2 * lo + 1)
        xx = torch.einsum("zui,zvj->zuvij", x1, x2)
        if mode == "uvu":
            n = mul1 * mul2
            w = weight[..., woff:woff + n]
            if shared_weights:
                res = torch.einsum("uv,ijk,zuvij->zuk", w.reshape(mul1, mul2), w3j, xx)
            else:
                res = torch.einsum("zuv,ijk,zuvij->zuk", w.reshape(z, mul1, mul2), w3j, xx)
        elif mode == "uvw":
            n = mul1 * mul2 * mulo
            w = weight[..., woff:woff + n]
            if shared_weights:
                res = torch.einsum("uvw,ijk,zuvij->zwk", w.reshape(mul1, mul2, mulo), w3j, xx)
            else:
                res = torch.einsum("zuvw,ijk,zuvij->zwk", w.reshape(z, mul1, mul2, mulo), w3j, xx)
        else:
            raise NotImplementedError(mode)
        woff += n
        res = pw * res.reshape(z, mulo * (2 * lo + 1))
        outs[io] = res if outs[io] is None else outs[io] + res
    for io, (mulo, lo, _) in enumerate(irreps_out):
        if outs[io] is None:
            outs[io] = x.new_zeros((z, mulo * (2 * lo + 1)))
    return torch.cat(outs, dim=1)
