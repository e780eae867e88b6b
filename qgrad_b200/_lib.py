"""ctypes binding of libqgrad_b200.so (the C ABI of include/qgrad_b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that
fails, or a compute entry point is called without a CUDA device, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QGRAD_B200_LIB") or os.path.join(HERE, "libqgrad_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "qgrad_b200.h")

QG_C64, QG_C128 = 0, 1
QG_ADJ_CONJ, QG_ADJ_TRANS, QG_ADJ_SKEW_BODY = 0, 1, 2
QG_EXPM_FWD, QG_EXPM_FWD_VJP = 0, 1
QG_RAND_HAAR, QG_RAND_REFERENCE = 0, 1


class QgradB200Error(RuntimeError):
    pass


_lib = None

_p, _i, _i64, _sz, _d, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double, C.c_uint64
_SIGS = {
    "qg_version": ([], _i),
    "qg_status_string": ([_i], C.c_char_p),
    "qg_launch_count": ([], _i64),
    "qg_release_scratch": ([], _i),
    "qg_set_plan_rigorous": ([_i], _i),
    "qg_set_c64_tensor": ([_i], _i),
    "qg_expm_workspace_bytes": ([_i, _i, _i64, _i, _i], _sz),
    "qg_expm": ([_i, _p, _p, _i64, _i, _p, _sz, _i, _p], _i),
    "qg_expm_fwd_vjp": ([_i, _p, _p, _p, _p, _i64, _i, _i, _p, _sz, _i, _p], _i),
    "qg_expm_fwd_save": ([_i, _p, _p, _i64, _i, _p, _sz, _i, _p], _i),
    "qg_expm_bwd_saved": ([_i, _p, _p, _p, _i64, _i, _i, _p, _sz, _i, _p], _i),
    "qg_expect_ket_fwd": ([_i, _p, _p, _p, _i64, _i, _i, _p], _i),
    "qg_expect_ket_bwd": ([_i, _p, _p, _p, _p, _p, _i64, _i, _i, _p], _i),
    "qg_expect_dm_fwd": ([_i, _p, _p, _p, _i64, _i, _i, _p], _i),
    "qg_expect_dm_bwd": ([_i, _p, _p, _p, _p, _p, _i64, _i, _i, _p], _i),
    "qg_fidelity_ket_fwd": ([_i, _p, _p, _p, _i64, _i, _p], _i),
    "qg_fidelity_ket_bwd": ([_i, _p, _p, _p, _p, _p, _i64, _i, _p], _i),
    "qg_fidelity_dm_fwd": ([_i, _p, _p, _p, _i64, _i, _p], _i),
    "qg_fidelity_dm_bwd": ([_i, _p, _p, _p, _p, _p, _i64, _i, _p], _i),
    "qg_displace_fwd": ([_i, _p, _p, _p, _p, _i64, _i, _p], _i),
    "qg_displace_bwd": ([_i, _p, _p, _p, _p, _p, _i64, _i, _p], _i),
    "qg_displace_apply_fwd": ([_i, _p, _p, _p, _p, _p, _i64, _i, _p], _i),
    "qg_displace_apply_bwd": ([_i, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _p], _i),
    "qg_givens_unitary_fwd": ([_i, _p, _p, _p, _p, _i64, _i, _p], _i),
    "qg_givens_unitary_bwd": ([_i, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _p], _i),
    "qg_make_rot_fwd": ([_i, _p, _p, _i64, _i, _i, _i, _p], _i),
    "qg_make_rot_bwd": ([_i, _p, _p, _p, _i64, _i, _i, _i, _p], _i),
    "qg_rot_zyz_fwd": ([_i, _p, _p, _i64, _p], _i),
    "qg_rot_zyz_bwd": ([_i, _p, _p, _p, _i64, _p], _i),
    "qg_rot_fidelity_fwd_grad": ([_i, _p, _p, _i, _p, _p, _i64, _p], _i),
    "qg_unitary_cost_fwd_grad": ([_i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i64, _i, _p], _i),
    "qg_snap_cost_fwd_grad": ([_i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _i64, _i, _p], _i),
    "qg_pauli_hamiltonian": ([_i, _p, _p, _p, _p, _i, _i, _i, _d, _p, _p], _i),
    "qg_pauli_project": ([_i, _p, _p, _p, _p, _i, _i, _i, _d, _p, _i, _p], _i),
    "qg_gemm": ([_i, _p, _p, _p, _i64, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _p], _i),
    "qg_overlap_loss": ([_i, _p, _p, _p, _p, _i64, _i, _i, _d, _p], _i),
    "qg_conj_transpose": ([_i, _p, _p, _i64, _i, _i, _p], _i),
    "qg_adam_step": ([_i, _p, _p, _p, _p, _i64, _d, _d, _d, _d, _i64, _p], _i),
    "qg_eigh": ([_i, _p, _p, _p, _i64, _i, _p], _i),
    "qg_sqrtm_psd": ([_i, _p, _p, _i64, _i, _p], _i),
    "qg_fidelity_uhlmann": ([_i, _p, _p, _p, _i64, _i, _p], _i),
    "qg_herm_stats": ([_i, _p, _p, _i64, _i, _p], _i),
    "qg_rand_ket": ([_i, _p, _i64, _i, _u64, _u64, _i, _p], _i),
    "qg_rand_hermitian": ([_i, _p, _i64, _i, _u64, _u64, _d, _p], _i),
    "qg_rand_uniform": ([_i, _p, _i64, _u64, _d, _d, _p], _i),
    "qg_adam_step_device": ([_i, _p, _p, _p, _p, _i64, _d, _d, _d, _d, _p, _p, _d, _p, _p], _i),
}


def declared_symbols():
    """Every function name include/qgrad_b200.h declares."""
    with open(HEADER) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(qg_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing=True):
    """Load (building first if needed) the shared library; raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise QgradB200Error(f"{LIB_PATH} is missing (run python -m qgrad_b200.build)")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (args, res) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().qg_status_string(int(status)).decode()
        raise QgradB200Error(f"{what or 'qgrad_b200'} failed: {msg} (status {status})")


def launch_count():
    return int(load().qg_launch_count())
