"""Import shim so that code written against the reference (``from qgrad.qgrad_qutip import ...``)
runs unchanged on the B200 implementation."""
