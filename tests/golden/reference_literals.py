"""Golden values the reference's own tests and notebooks hold for the hot path, restated as
data (numbers only) with the file:line each comes from (relative to /root/reference).
These pin the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path."""
import numpy as np

# qgrad/tests/test_qgrad_qutip.py:299-332 -- squeeze(4, 0.1+0.1j); np.allclose default tol.
SQUEEZE_4_Z = 0.1 + 0.1j
SQUEEZE_4 = np.array([
    [0.99500417, 0.0, 0.07059289 - 0.07059289j, 0.0],
    [0.0, 0.98503746, 0.0, 0.12186303 - 0.12186303j],
    [-0.07059289 - 0.07059289j, 0.0, 0.99500417, 0.0],
    [0.0, -0.12186303 - 0.12186303j, 0.0, 0.98503746],
], dtype=np.complex128)

# qgrad/tests/test_qgrad_qutip.py:338-370 -- Displace(4)(0.25); np.allclose default tol.
DISPLACE_4_ALPHA = 0.25
DISPLACE_4 = np.array([
    [0.96923323, -0.24230859, 0.04282883, -0.00626025],
    [0.24230859, 0.90866411, -0.33183303, 0.07418172],
    [0.04282883, 0.33183303, 0.84809499, -0.41083747],
    [0.00626025, 0.07418172, 0.41083747, 0.90866411],
], dtype=np.complex128)
# :371-373 -- Displace(N)(0) == eye(N) for N in range(2, 50, 5), 6 decimals.
DISPLACE_ZERO_DIMS = list(range(2, 50, 5))

# qgrad/tests/test_qgrad_qutip.py:126-134 / :145-153 -- 3x3 ladder operators.
DESTROY_3 = np.array([[0, 1.0, 0], [0, 0, 1.41421356], [0, 0, 0]], dtype=np.complex128)
CREATE_3 = np.array([[0, 0, 0], [1.0, 0, 0], [0, 1.41421356, 0]], dtype=np.complex128)

# qgrad/tests/test_qgrad_qutip.py:225-247 -- complex64 ket and its dagger (exact equality).
DAG_KET = np.array([[0.04896761 + 0.18014458j], [0.6698803 + 0.13728367j],
                    [-0.07598839 + 0.38113445j], [-0.00505985 + 0.10700243j],
                    [-0.18735261 + 0.5476768j]], dtype=np.complex64)
DAG_KET_DAG = np.array([[0.04896761 - 0.18014458j, 0.6698803 - 0.13728367j,
                         -0.07598839 - 0.38113445j, -0.00505985 - 0.10700243j,
                         -0.18735261 - 0.5476768j]], dtype=np.complex64)

# qgrad/tests/test_qgrad_qutip.py:375-400 -- (N, (theta, phi), (i, j)) cases whose _make_rot
# must be unitary to 6 decimals.
MAKE_ROT_CASES = [
    (2, (np.pi / 5, np.pi / 5), (1, 0)), (3, (0.0, 0.0), (2, 0)), (30, (0.0, 0.0), (1, 0)),
    (40, (np.pi / 8, np.pi / 7), (20, 15)), (10, (0.0, 2 * np.pi), (9, 8)),
    (5, (np.pi / 3, np.pi / 3), (3, 2)), (23, (np.pi / 2, np.pi / 2), (20, 0)),
    (50, (3 * np.pi, 5 * np.pi), (30, 20)), (64, (0.0, 0.0), (63, 62)),
    (75, (2 * np.pi, 2 * np.pi), (63, 62)), (84, (np.pi, np.pi), (2, 0)),
    (95, (np.pi / 4, np.pi / 4), (1, 0)),
]
# :418-423 -- Unitary(N) unitary to 6 decimals for N in range(2, 30, 6); :496-500 rand_unitary
UNITARY_DIMS = list(range(2, 30, 6))
RAND_UNITARY_DIMS = list(range(2, 43, 10))

# qgrad/tests/test_qgrad_qutip.py:213-218 -- |<a>_{coherent(10, 0.5)} - 0.5| < 1e-4 and
# coherent(N, 0) == basis(N, 0) for N in range(2, 30, 5).
COHERENT_N, COHERENT_ALPHA, COHERENT_TOL = 10, 0.5, 1e-4
COHERENT_ZERO_DIMS = list(range(2, 30, 5))

# examples/notebooks/Circuit-Learning.ipynb:195-202 -- losses printed at epoch 50..400
# (Adam 1e-2 from [0,0,0]; deterministic; pins expect/basis/sigmaz values + gradients + Adam).
CIRCUIT_LEARNING_LOSSES = [0.262702, -0.274620, -0.691323, -0.902352,
                           -0.976295, -0.995449, -0.999299, -0.999913]
