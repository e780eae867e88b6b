"""qgrad_b200 -- B200-native implementation of qgrad's differentiable quantum-operator hot path.

``qgrad_b200.qgrad_qutip`` mirrors ``qgrad.qgrad_qutip``; the kernels live in
``libqgrad_b200.so`` (C ABI: ``include/qgrad_b200.h``).  See DESIGN.md.
"""
from ._lib import QgradB200Error, launch_count, load as load_library  # noqa: F401

__version__ = "0.1.0"
