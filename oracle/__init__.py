"""CPU oracle for the qgrad hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package, and there only as the checker or the timed CPU
baseline, never as the product path.  Nothing under ``qgrad_b200/`` imports it.

What it restates (all citations relative to the reference checkout, ``/root/reference``):

* ``oracle.qgrad_ref``     -- ``qgrad/qgrad_qutip.py`` in full (torch-CPU, so that
  ``torch.autograd`` stands in for ``jax.grad``).
* ``oracle.examples_ref``  -- the example-local builders and cost functions the configs name
  (``examples/Qubit_Rotation.py``, ``SNAP_gates.py``, ``Unitary-Learning-qgrad.py``,
  ``Unitary-Learning-no-qgrad.py``, ``Circuit-Learning.py``) and the Adam optimiser
  (``jax.experimental.optimizers.adam`` defaults).
* ``oracle.expm_ref``      -- three mutually independent expm / Frechet oracles (SciPy Pade,
  Hermitian eigendecomposition + Daleckii-Krein, torch-CPU autograd) plus a NumPy
  restatement of the Higham-2005 scaling-and-squaring algorithm the CUDA kernels implement.

Parity status
-------------
* Pinned against the reference's own golden literals and known-answer values
  (``qgrad/tests/test_qgrad_qutip.py`` and the Circuit-Learning notebook trace): see
  ``tests/golden/reference_literals.py`` and ``tests/test_oracle_golden.py``.
* expm *values* are pinned only by ``test_squeeze`` (4x4, ~1e-8).  The expm **gradient**
  and the north-star's ``jax.scipy.linalg.expm`` have no call site, test or golden vector in
  the reference, and JAX is not installable here: for those rows the oracle is
  **parity unpinned** against the reference itself and is anchored instead on the mutual
  agreement of three independent third-party implementations (SciPy 1.18.1
  ``expm``/``expm_frechet``, eigendecomposition, torch 2.11 ``matrix_exp`` autograd).
"""
