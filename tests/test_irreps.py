"""CPU: Irreps bookkeeping follows the e3nn grammar the reference relies on."""
import pytest

from equiformer_b200.o3 import Irrep, Irreps


def test_parse_and_dims():
    ir = Irreps("128x0e+64x1e+32x2e")
    assert ir.dim == 480 and ir.num_irreps == 224 and ir.lmax == 2 and len(ir) == 3
    assert [(m, i.l, i.p) for m, i in ir] == [(128, 0, 1), (64, 1, 1), (32, 2, 1)]
    assert str(ir) == "128x0e+64x1e+32x2e"
    assert ir.slices() == [slice(0, 128), slice(128, 320), slice(320, 480)]
    assert Irreps("1o").dim == 3 and Irreps("0e+1e") == Irreps("1x0e+1x1e")
    assert Irreps([(3, (1, -1)), (2, "0e")]) == Irreps("3x1o+2x0e")
    assert Irreps(Irreps("4x2o")) == "4x2o"
    mul, (l, p) = Irreps("5x3o")[0]           # nets/fast_activation.py:33 unpacks like this
    assert (mul, l, p) == (5, 3, -1)
    with pytest.raises(ValueError):
        Irrep("2x")


def test_products_and_membership():
    assert Irrep("1e") * Irrep("2e") == [Irrep("1e"), Irrep("2e"), Irrep("3e")]
    assert Irrep("1o") * Irrep("1o") == [Irrep("0e"), Irrep("1e"), Irrep("2e")]
    assert Irrep(2, 1) in Irreps("3x0e+2x2e") and Irrep(2, -1) not in Irreps("3x0e+2x2e")
    assert Irrep(0, 1) == Irrep("0e") and Irrep(0, 1).is_scalar() and Irrep("1o").dim == 3


def test_mul_simplify_sort():
    head = Irreps("32x0e+16x1e+8x2e")
    rep = head * 4                               # tuple repetition (graph_attention_transformer.py:434)
    assert len(rep) == 12 and rep.dim == 4 * head.dim
    assert str(rep.simplify()) == str(rep)       # only ADJACENT equal irreps merge
    assert str(Irreps("2x0e+3x0e+1x1e+4x0e").simplify()) == "5x0e+1x1e+4x0e"
    s = Irreps("1x1e+2x0o+3x0e+1x1o").sort()
    assert str(s.irreps) == "2x0o+3x0e+1x1o+1x1e"  # e3nn: odd before even at equal l
    assert s.inv == (1, 2, 3, 0) and s.p == (3, 0, 1, 2)
    assert (Irreps("2x0e") + Irreps("1x1o")) == "2x0e+1x1o"
    assert Irreps.spherical_harmonics(2) == "1x0e+1x1o+1x2e" and Irreps.spherical_harmonics(2, p=1) == "1x0e+1x1e+1x2e"


def test_sort_even_first_matches_reference_helper():
    from equiformer_b200.nets.tensor_product_rescale import sort_irreps_even_first
    r = sort_irreps_even_first(Irreps("4x1o+2x0o+3x0e+1x1e"))
    assert str(r.irreps) == "3x0e+2x0o+1x1e+4x1o" and r.inv == (2, 1, 3, 0)
    assert [r.p[i] for i in r.inv] == [0, 1, 2, 3]
