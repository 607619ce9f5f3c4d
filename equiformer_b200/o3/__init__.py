"""Minimal ``e3nn.o3`` surface used by the Equiformer hot path (Irreps, Wigner-3j, SH, TensorProduct)."""
from .irreps import Irrep, Irreps  # noqa: F401
from .sh import spherical_harmonics  # noqa: F401
from .tensor_product import Instruction, TensorProduct  # noqa: F401
from .wigner import wigner_3j, wigner_3j_np  # noqa: F401
