"""``qgrad.qgrad_qutip`` drop-in: re-exports the B200 implementation (qgrad_b200.qgrad_qutip)."""
from qgrad_b200.qgrad_qutip import *  # noqa: F401,F403
from qgrad_b200.qgrad_qutip import _make_rot, grad  # noqa: F401
