"""Drop-in for the reference's `op` package (op/__init__.py:1-2): `upfirdn2d`, `fused_leaky_relu`,
`FusedLeakyReLU`, backed by libst_b200 instead of the two JIT-built torch extensions
(op/upfirdn2d.py:10-16, op/fused_act.py:11-17)."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu  # noqa: F401
from .upfirdn2d import upfirdn2d  # noqa: F401
