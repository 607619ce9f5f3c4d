"""Irreducible-representation bookkeeping for O(3) (``Irrep`` / ``Irreps``).

Host-side only: pure Python, no tensors.  It mirrors the subset of the
``e3nn.o3.Irrep`` / ``e3nn.o3.Irreps`` surface (e3nn 0.4.4, pinned by the
reference at ``env/env_equiformer.yml:358``) that the reference's hot path
touches:

* ``nets/graph_attention_transformer.py:166-177`` - ``ir_in * ir_edge``,
  ``ir in irreps``, ``Irreps(list)``, ``sort`` replacement,
* ``nets/graph_attention_transformer.py:434-439`` - ``irreps * int``,
  ``simplify``, ``+``,
* ``nets/tensor_product_rescale.py:60-71`` - ``str(irreps)`` parsing,
  ``slices()``,
* ``nets/fast_activation.py:33-51`` - unpacking ``(mul, (l, p))``.

e3nn itself is not installable in this image, so this file is a from-scratch
implementation of the published behaviour (``'128x0e+64x1e'`` grammar, the
``(l, p)`` ordering rules, ``Irreps * int`` == tuple repetition, ``simplify``
merging *adjacent* equal irreps only, stable ``sort`` by ``(l, p)`` with odd
before even).
"""
from __future__ import annotations

import collections
import re
from typing import Iterator, List, Tuple, Union

_IRREP_RE = re.compile(r"^\s*(\d+)\s*([eoy])\s*$")


class Irrep(tuple):
    """One irreducible representation of O(3): degree ``l`` and parity ``p`` (+1 'e', -1 'o')."""

    __slots__ = ()

    def __new__(cls, l: Union[int, str, "Irrep", tuple], p: Union[int, None] = None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                m = _IRREP_RE.match(l)
                if m is None:
                    raise ValueError(f"unable to convert string '{l}' into an Irrep")
                deg = int(m.group(1))
                tag = m.group(2)
                par = {"e": 1, "o": -1, "y": (-1) ** deg}[tag]
                return tuple.__new__(cls, (deg, par))
            if isinstance(l, tuple) and len(l) == 2:
                l, p = l
            elif hasattr(l, "l") and hasattr(l, "p"):  # foreign (real e3nn) Irrep
                l, p = l.l, l.p
            else:
                raise ValueError(f"unable to convert {l!r} into an Irrep")
        if not isinstance(l, int) or l < 0:
            raise ValueError(f"l must be a non-negative integer, got {l!r}")
        if p not in (-1, 1):
            raise ValueError(f"parity must be +1 or -1, got {p!r}")
        return tuple.__new__(cls, (l, p))

    @property
    def l(self) -> int:  # noqa: E743
        return self[0]

    @property
    def p(self) -> int:
        return self[1]

    @property
    def dim(self) -> int:
        return 2 * self[0] + 1

    def is_scalar(self) -> bool:
        return self[0] == 0 and self[1] == 1

    def __repr__(self) -> str:
        return f"{self[0]}{'e' if self[1] == 1 else 'o'}"

    __str__ = __repr__

    def __mul__(self, other):  # selection rule |l1-l2| .. l1+l2, parity product
        if isinstance(other, int):
            raise TypeError("use `mul * ir` (int on the left) to build Irreps")
        other = Irrep(other)
        p = self.p * other.p
        return [Irrep(l, p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]

    def __rmul__(self, mul: int) -> "Irreps":
        return Irreps([(int(mul), self)])

    def __add__(self, other) -> "Irreps":
        return Irreps(self) + Irreps(other)

    def count(self, _value):  # tuple API we do not want to leak
        raise NotImplementedError

    def index(self, _value):
        raise NotImplementedError

    @classmethod
    def iterator(cls, lmax: Union[int, None] = None) -> Iterator["Irrep"]:
        l = 0
        while lmax is None or l <= lmax:
            yield Irrep(l, (-1) ** l)
            yield Irrep(l, -((-1) ** l))
            l += 1


class _MulIr(tuple):
    """``(mul, ir)`` pair; unpacks as a 2-tuple like e3nn's ``_MulIr``."""

    __slots__ = ()

    def __new__(cls, mul: int, ir=None):
        if ir is None:
            mul, ir = mul
        return tuple.__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self) -> int:
        return self[0]

    @property
    def ir(self) -> Irrep:
        return self[1]

    @property
    def dim(self) -> int:
        return self[0] * self[1].dim

    def __repr__(self) -> str:
        return f"{self[0]}x{self[1]}"


_SortResult = collections.namedtuple("sort", ["irreps", "p", "inv"])


def _inverse_perm(p: Tuple[int, ...]) -> Tuple[int, ...]:
    out = [0] * len(p)
    for i, j in enumerate(p):
        out[j] = i
    return tuple(out)


class Irreps(tuple):
    """Direct sum of irreps with multiplicities, e.g. ``Irreps('128x0e+64x1e+32x2e')``."""

    __slots__ = ()

    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return irreps
        out: List[_MulIr] = []
        if irreps is None:
            pass
        elif isinstance(irreps, Irrep):
            out.append(_MulIr(1, irreps))
        elif isinstance(irreps, _MulIr):
            out.append(irreps)
        elif isinstance(irreps, str):
            text = irreps.strip()
            if text:
                for chunk in text.split("+"):
                    chunk = chunk.strip()
                    if "x" in chunk:
                        mul_s, ir_s = chunk.split("x")
                        out.append(_MulIr(int(mul_s), Irrep(ir_s.strip())))
                    else:
                        out.append(_MulIr(1, Irrep(chunk)))
        elif hasattr(irreps, "__iter__"):
            for item in irreps:
                if isinstance(item, _MulIr):
                    out.append(item)
                elif isinstance(item, Irrep):
                    out.append(_MulIr(1, item))
                elif isinstance(item, str):
                    out.append(_MulIr(1, Irrep(item)))
                elif hasattr(item, "mul") and hasattr(item, "ir"):  # real e3nn _MulIr
                    out.append(_MulIr(item.mul, Irrep(item.ir.l, item.ir.p)))
                elif len(item) == 2:
                    mul, ir = item
                    if not isinstance(mul, int) or mul < 0:
                        raise ValueError(f"unable to interpret {item!r} as (mul, ir)")
                    out.append(_MulIr(mul, Irrep(ir)))
                else:
                    raise ValueError(f"unable to interpret {item!r} as an irrep")
        else:  # foreign Irreps object (real e3nn): go through its string form
            return cls.__new__(cls, str(irreps))
        return tuple.__new__(cls, out)

    # ------------------------------------------------------------------ views
    @staticmethod
    def spherical_harmonics(lmax: int, p: int = -1) -> "Irreps":
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    @property
    def dim(self) -> int:
        return sum(mul * ir.dim for mul, ir in self)

    @property
    def num_irreps(self) -> int:
        return sum(mul for mul, _ in self)

    @property
    def ls(self) -> List[int]:
        return [ir.l for mul, ir in self for _ in range(mul)]

    @property
    def lmax(self) -> int:
        if len(self) == 0:
            raise ValueError("Cannot get lmax of empty Irreps")
        return max(ir.l for _, ir in self)

    def slices(self) -> List[slice]:
        out, start = [], 0
        for mul_ir in self:
            out.append(slice(start, start + mul_ir.dim))
            start += mul_ir.dim
        return out

    def simplify(self) -> "Irreps":
        out: List[Tuple[int, Irrep]] = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def remove_zero_multiplicities(self) -> "Irreps":
        return Irreps([(mul, ir) for mul, ir in self if mul > 0])

    def sort(self):
        """Stable sort by ``(l, p)`` (odd before even), like ``e3nn.o3.Irreps.sort``."""
        keyed = sorted((ir, i, mul) for i, (mul, ir) in enumerate(self))
        inv = tuple(i for _, i, _ in keyed)
        p = _inverse_perm(inv)
        return _SortResult(Irreps([(mul, ir) for ir, _, mul in keyed]), p, inv)

    def count(self, ir) -> int:
        ir = Irrep(ir)
        return sum(mul for mul, ir2 in self if ir2 == ir)

    def index(self, _value):
        raise NotImplementedError

    # -------------------------------------------------------------- operators
    def __getitem__(self, i):
        x = tuple.__getitem__(self, i)
        if isinstance(i, slice):
            return Irreps(x)
        return x

    def __contains__(self, ir) -> bool:
        try:
            ir = Irrep(ir)
        except ValueError:
            return False
        return any(ir == ir2 for _, ir2 in self)

    def __add__(self, other) -> "Irreps":
        return Irreps(tuple.__add__(self, Irreps(other)))

    def __radd__(self, other) -> "Irreps":
        return Irreps(other) + self

    def __mul__(self, n: int) -> "Irreps":
        if not isinstance(n, int):
            raise NotImplementedError("Irreps can only be repeated by an int")
        return Irreps(tuple.__mul__(self, n))

    __rmul__ = __mul__

    def __eq__(self, other) -> bool:
        try:
            other = Irreps(other)
        except (ValueError, TypeError):
            return False
        return tuple.__eq__(self, other)

    def __ne__(self, other) -> bool:
        return not self.__eq__(other)

    def __hash__(self) -> int:
        return tuple.__hash__(self)

    def __repr__(self) -> str:
        return "+".join(f"{mul_ir}" for mul_ir in self)

    __str__ = __repr__
