"""Tolerance bookkeeping for the GPU parity tests.

The north star's bar is ONE number per dtype -- relative error <= 1e-12 for complex128 and <= 1e-5 for complex64, on
values and gradients -- for inputs of the named shapes (H / sqrt(n) Hamiltonians, Haar kets, angles in [0, 2 pi),
|alpha| ~ 1).  ``check`` asserts an error against that bar (or against an explicitly stated stress bound, with the
reason next to it) and records the measured value; at the end of a GPU session the records are written to
``gpurun_out/parity_errors.json`` so DESIGN.md's accuracy table quotes measurements, not bounds.
"""
import json
import os

import numpy as np
import torch

TOL = {torch.complex128: 1e-12, torch.complex64: 1e-5}
RECORDS = []


def relerr(x, ref, floor=0.0):
    """||x - ref||_F / max(||ref||_F, floor) over the whole array (scalars included)."""
    x = np.asarray(x, dtype=np.complex128).ravel()
    ref = np.asarray(ref, dtype=np.complex128).ravel()
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), floor, 1e-300))


def check(err, dt, what, stress=None, why=None):
    """assert ``err`` <= the dtype's bar -- or <= ``stress`` (a looser, explicitly justified bound for an input outside
    the named shapes: ``why`` is mandatory then) -- and log it."""
    bar = TOL[dt] if stress is None else stress
    if stress is not None:
        assert why, "a stress bound needs its reason"
    RECORDS.append({"what": what, "dtype": "c128" if dt == torch.complex128 else "c64", "err": float(err), "bar": bar,
                    "named_shape": stress is None, **({"why": why} if why else {})})
    assert err <= bar, f"{what}: {err:.3e} > {bar:.1e}"


def dump(path):
    if RECORDS:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(RECORDS, f, indent=0)
