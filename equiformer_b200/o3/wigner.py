"""Real-basis Wigner-3j symbols (the Clebsch-Gordan tables the edge kernels contract with).

Product-side generator: exact rational su(2) Clebsch-Gordan coefficients via the
van-der-Waerden form of the Racah sum, rotated into e3nn's real spherical basis and
Frobenius-normalised.  The result is what ``e3nn.o3.wigner_3j`` (0.4.4) returns, which
is the tensor ``o3.TensorProduct`` contracts with inside the reference's
``TensorProductRescale`` (``nets/tensor_product_rescale.py:33-37``).

The values are consumed on the host only: `equiformer_b200/plan.py` bakes them (times the
e3nn path weight) into the dense per-path tables that `eqf_plan_create` uploads once.
An independent restatement lives in ``oracle/e3nn_ref.py`` and the two are compared in
``tests/test_o3.py``; nothing here imports the oracle.
"""
from __future__ import annotations

import functools
from fractions import Fraction
from math import factorial, isqrt

import numpy as np


def _tri(a: int, b: int, c: int) -> Fraction:
    return Fraction(
        factorial(a + b - c) * factorial(a - b + c) * factorial(-a + b + c),
        factorial(a + b + c + 1),
    )


def _sqrt_fraction(q: Fraction) -> float:
    # exact when both parts are perfect squares, otherwise one correctly rounded sqrt
    n, d = q.numerator, q.denominator
    rn, rd = isqrt(n), isqrt(d)
    if rn * rn == n and rd * rd == d:
        return rn / rd
    return float(np.sqrt(np.float64(float(q))))


def su2_cg(j1: int, m1: int, j2: int, m2: int, j3: int, m3: int) -> float:
    """<j1 m1 j2 m2 | j3 m3> for integer spins (Condon-Shortley phase)."""
    if m1 + m2 != m3 or not (abs(j1 - j2) <= j3 <= j1 + j2):
        return 0.0
    if abs(m1) > j1 or abs(m2) > j2 or abs(m3) > j3:
        return 0.0
    pref = Fraction(2 * j3 + 1) * _tri(j1, j2, j3)
    pref *= (
        factorial(j1 + m1) * factorial(j1 - m1) * factorial(j2 + m2)
        * factorial(j2 - m2) * factorial(j3 + m3) * factorial(j3 - m3)
    )
    total = Fraction(0)
    for k in range(0, j1 + j2 - j3 + 1):
        args = (k, j1 + j2 - j3 - k, j1 - m1 - k, j2 + m2 - k, j3 - j2 + m1 + k, j3 - j1 - m2 + k)
        if min(args) < 0:
            continue
        den = 1
        for a in args:
            den *= factorial(a)
        total += Fraction((-1) ** k, den)
    sign = 1 if total >= 0 else -1
    return sign * _sqrt_fraction(pref * total * total)


def real_to_complex(l: int) -> np.ndarray:
    """Unitary Q_l with (complex SH)_m = sum_k Q[m, k] (real SH)_k, e3nn phase ``(-i)^l``."""
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    s = 1.0 / np.sqrt(2.0)
    for m in range(-l, 0):
        q[l + m, l - m] = s
        q[l + m, l + m] = -1j * s
    q[l, l] = 1.0
    for m in range(1, l + 1):
        q[l + m, l + m] = (-1) ** m * s
        q[l + m, l - m] = 1j * (-1) ** m * s
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def _wigner_3j_cached(l1: int, l2: int, l3: int) -> np.ndarray:
    c = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1), dtype=np.float64)
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            m3 = m1 + m2
            if abs(m3) <= l3:
                c[l1 + m1, l2 + m2, l3 + m3] = su2_cg(l1, m1, l2, m2, l3, m3)
    q1, q2, q3 = real_to_complex(l1), real_to_complex(l2), real_to_complex(l3)
    t = np.einsum("ij,kl,mn,ikn->jlm", q1, q2, np.conj(q3.T), c.astype(np.complex128))
    if np.abs(t.imag).max() > 1e-9:
        raise AssertionError(f"wigner_3j({l1},{l2},{l3}) is not real in the e3nn basis")
    t = np.ascontiguousarray(t.real)
    t /= np.linalg.norm(t)
    t[np.abs(t) < 1e-14] = 0.0
    t.setflags(write=False)
    return t


def wigner_3j_np(l1: int, l2: int, l3: int) -> np.ndarray:
    """float64 array ``[2l1+1, 2l2+1, 2l3+1]``, Frobenius norm 1; zeros if the triangle rule fails."""
    if not (abs(l1 - l2) <= l3 <= l1 + l2):
        raise ValueError(f"({l1},{l2},{l3}) violates the triangle inequality")
    return _wigner_3j_cached(int(l1), int(l2), int(l3))


def wigner_3j(l1: int, l2: int, l3: int, dtype=None, device=None):
    """torch view of :func:`wigner_3j_np` (same call shape as ``e3nn.o3.wigner_3j``)."""
    import torch

    out = torch.from_numpy(wigner_3j_np(l1, l2, l3).copy())
    return out.to(dtype=dtype or torch.get_default_dtype(), device=device)
