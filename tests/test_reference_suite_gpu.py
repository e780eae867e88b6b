"""The reference's own unit tests (qgrad/tests/test_qgrad_qutip.py), run against the CUDA path.

Same test names, cases and assertions; ``jnp`` arrays become torch CUDA tensors and QuTiP's random
states (``rand_ket``/``rand_dm``/``rand_herm``, used by the reference only to make inputs) become
seeded NumPy equivalents.  Line numbers refer to qgrad/tests/test_qgrad_qutip.py."""
import numpy as np
import pytest
import torch
from numpy.testing import (assert_almost_equal, assert_array_almost_equal, assert_array_equal, assert_equal,
                           assert_raises)

from tests.golden import reference_literals as G

pytestmark = pytest.mark.gpu

_rng = np.random.default_rng(2020)


def rand_ket(n):
    k = _rng.standard_normal((n, 1)) + 1j * _rng.standard_normal((n, 1))
    return k / np.linalg.norm(k)


def rand_dm(n):
    x = _rng.standard_normal((n, n)) + 1j * _rng.standard_normal((n, n))
    rho = x @ x.conj().T
    return rho / np.trace(rho).real


def rand_herm(n):
    x = _rng.standard_normal((n, n)) + 1j * _rng.standard_normal((n, n))
    return (x + x.conj().T) / 2


def npy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


@pytest.fixture(scope="module")
def q():
    import qgrad.qgrad_qutip as mod      # the import path reference users have
    return mod


def test_fidelity(q):                                                   # :44-64
    ket0 = np.array([[1.0], [0]]); ket1 = np.array([[0.0], [1]])
    ket_plus = 1 / np.sqrt(2) * (ket0 + ket1); ket_minus = 1 / np.sqrt(2) * (ket0 - ket1)
    ket_complx = rand_ket(2)
    assert q.fidelity(ket0, ket1) == 0.0
    assert q.fidelity(ket0, ket0) == 1.0
    assert q.fidelity(ket1, ket1) == 1.0
    assert_almost_equal(npy(q.fidelity(ket_plus, ket_minus)), 0.0)
    assert q.fidelity(rand_ket(4), rand_ket(4)) <= 1.0
    assert q.fidelity(rand_ket(10), rand_ket(10)) >= 0.0
    assert np.isclose(npy(q.fidelity(ket_complx, ket_complx)), 1.0)
    assert_almost_equal(npy(q.fidelity(ket_plus, ket0)), 1.0 / 2.0)
    assert_almost_equal(npy(q.fidelity(ket_plus, ket1)), 1.0 / 2.0)
    assert_almost_equal(npy(q.fidelity(ket0, ket_minus)), 1.0 / 2.0)
    assert_almost_equal(npy(q.fidelity(ket1, ket_minus)), 1.0 / 2.0)


def test_fidelity_max_dm(q):                                            # :67-73
    for _ in range(10):
        rho1, rho2 = rand_dm(25), rand_dm(25)
        assert_almost_equal(npy(q.fidelity(rho1, rho1)), 1.0, decimal=4)
        assert_almost_equal(npy(q.fidelity(rho2, rho2)), 1.0, decimal=4)


def test_fidelity_max_ket(q):                                           # :76-82
    for _ in range(10):
        ket1, ket2 = rand_ket(25), rand_ket(25)
        assert_almost_equal(npy(q.fidelity(ket1, ket1)), 1.0, decimal=6)
        assert_almost_equal(npy(q.fidelity(ket2, ket2)), 1.0, decimal=6)


def test_fidelity_bounded(q, tol=1e-7):                                 # :85-108 (mixed-mixed, pure-mixed, pure-pure)
    for _ in range(10):
        for a, b in ((rand_dm(25), rand_dm(25)), (rand_dm(25), rand_ket(25)), (rand_ket(25), rand_ket(25))):
            F = float(q.fidelity(a, b))
            assert -tol <= F <= 1 + tol


def test_basis(q):                                                      # :111-115
    np_arr = np.zeros((4, 1), dtype=np.complex64); np_arr[2, 0] = 1.0
    assert np.array_equal(npy(q.basis(4, 2)), np_arr)


def test_destroy(q):                                                    # :118-136
    lowered = q.destroy(10) @ q.basis(10, 9)
    assert_array_almost_equal(npy(lowered), 3.0 * npy(q.basis(10, 8)))
    assert_equal(np.allclose(G.DESTROY_3, npy(q.destroy(3))), True)
    assert_equal(np.allclose(npy(q.dag(q.destroy(3))), npy(q.create(3))), True)


def test_create(q):                                                     # :139-153
    raised = q.create(5) @ q.basis(5, 3)
    assert_equal(np.allclose(npy(raised), 2.0 * npy(q.basis(5, 4))), True)
    assert_equal(np.allclose(G.CREATE_3, npy(q.create(3))), True)


def test_sigmas(q):                                                     # :156-167
    assert_array_equal(npy(q.sigmax()), np.array([[0.0, 1.0], [1.0, 0.0]]))
    assert_array_equal(npy(q.sigmay()), np.array([[0.0, -1.0j], [1.0j, 0.0]]))
    assert_array_equal(npy(q.sigmaz()), np.array([[1.0, 0.0], [0.0, -1.0]]))


@pytest.mark.parametrize("op", ["sigmax", "sigmay", "sigmaz"])
@pytest.mark.parametrize("k", [0, 1])
def test_expect_sigmaxyz(q, op, k):                                     # :170-180
    oper, state = getattr(q, op)(), q.basis(2, k)
    if op != "sigmaz":
        assert q.expect(oper, state) == 0.0
    elif k == 0:
        assert q.expect(oper, state) == 1.0
    else:
        assert q.expect(oper, state) == -1.0


@pytest.mark.parametrize("n, k", [(2, 0), (4, 0), (20, 20)])
def test_expect_herm(q, n, k):                                          # :183-194 (basis(20, 20) is the zero vector)
    assert torch.imag(q.expect(rand_herm(n), q.basis(n, k))) == 0.0


@pytest.mark.parametrize("kind", ["herm", "dm"])
def test_expect_dag(q, kind):                                           # :197-210
    oper = rand_herm(5) if kind == "herm" else rand_dm(5)
    state = rand_ket(5)
    expected = (state.conj().T @ oper @ state)[0, 0]
    assert abs(complex(q.expect(oper, state)) - expected) < 1e-6


def test_coherent(q):                                                   # :213-218
    assert abs(complex(q.expect(q.destroy(10), q.coherent(10, 0.5))) - 0.5) < 1e-4
    for N in range(2, 30, 5):
        assert_array_almost_equal(npy(q.coherent(N, 0)), npy(q.basis(N, 0)))


def test_dag_ket(q):                                                    # :220-248
    assert_array_equal(npy(q.dag(q.basis(2, 0))), [[1.0, 0.0]])
    assert_array_equal(npy(q.dag(q.basis(2, 1))), [[0.0, 1.0]])
    assert_array_equal(npy(q.dag(G.DAG_KET)), G.DAG_KET_DAG)


@pytest.mark.repeat(10)
def test_dag_dot(q):                                                    # :251-256
    ket = rand_ket(int(_rng.integers(3, 10)))
    assert_almost_equal(npy(q.dag(ket) @ torch.as_tensor(ket).cuda()), 1.0)


def test_isket_isbra(q):                                                # :259-280
    for i in range(2, 6):
        assert q.isket(rand_ket(i)) and not q.isbra(rand_ket(i))
        assert not q.isket(npy(q.dag(rand_ket(i)))) and q.isbra(npy(q.dag(rand_ket(i))))
        assert not q.isket(rand_dm(i)) and not q.isbra(rand_dm(i))


def test_to_dm(q):                                                      # :283-296
    dm0 = np.array([[1.0, 0.0], [0.0, 0.0]], dtype=np.complex64)
    dm1 = np.array([[0.0, 0.0], [0.0, 1.0]], dtype=np.complex64)
    assert_array_equal(npy(q.to_dm(q.basis(2, 0))), dm0)
    assert_array_equal(npy(q.to_dm(q.basis(2, 1))), dm1)
    assert_array_equal(npy(q.to_dm(q.dag(q.basis(2, 0)))), dm0)
    assert_array_equal(npy(q.to_dm(q.dag(q.basis(2, 1)))), dm1)


def test_squeeze(q):                                                    # :299-332
    assert_equal(np.allclose(npy(q.squeeze(4, G.SQUEEZE_4_Z)), G.SQUEEZE_4), True)


class TestDisplace:                                                     # :335-373
    def test_displace(self, q):
        dp = q.Displace(4)
        assert_equal(np.allclose(npy(dp(G.DISPLACE_4_ALPHA)), G.DISPLACE_4), True)
        for N in range(2, 50, 5):
            assert_array_almost_equal(npy(q.Displace(N)(0)), np.eye(N))


@pytest.mark.parametrize("N, params, idx", G.MAKE_ROT_CASES)
def test_make_rot(q, N, params, idx):                                   # :375-400
    rotation = q._make_rot(N, list(params), idx)
    assert_array_almost_equal(npy(rotation @ q.dag(rotation)), np.eye(N))
    assert_array_almost_equal(npy(q.dag(rotation) @ rotation), np.eye(N))


class TestUnitary:                                                      # :403-423
    @staticmethod
    def generate_params(N, seed=0):
        rng = np.random.default_rng(seed)
        for _ in range(3):
            yield (rng.uniform(0, 2 * np.pi, N * (N - 1) // 2).astype(np.float32),
                   rng.uniform(0, 2 * np.pi, N * (N - 1) // 2).astype(np.float32),
                   rng.uniform(0, 2 * np.pi, N).astype(np.float32))

    def test_unitary(self, q):
        for N in range(2, 30, 6):
            for thetas, phis, omegas in TestUnitary.generate_params(N):
                unitary = q.Unitary(N)(thetas, phis, omegas)
                assert unitary.dtype == torch.complex64
                assert_array_almost_equal(npy(unitary @ q.dag(unitary)), np.eye(N))
                assert_array_almost_equal(npy(q.dag(unitary) @ unitary), np.eye(N))


def test_rand_ket_norm_and_seed(q):                                     # :426-444
    for N in range(2, 40, 6):
        assert_almost_equal(float(torch.linalg.norm(q.rand_ket(N))), 1.0, decimal=5)
    for N in range(2, 30, 6):
        for seed1, seed2 in zip(range(0, 1000, 100), range(1000, 2000, 100)):
            assert_raises(AssertionError, assert_array_equal, npy(q.rand_ket(N, seed1)), npy(q.rand_ket(N, seed2)))


@pytest.mark.parametrize("herm", [True, False])
def test_isherm(q, herm):                                               # :465-482
    if herm:
        for op in (q.sigmax(), q.sigmay(), q.sigmaz(), rand_herm(2), rand_herm(4), rand_herm(20)):
            assert q.isherm(op)
    else:
        assert not q.isherm(np.arange(9).reshape(3, 3)) and not q.isherm(np.arange(16).reshape(4, 4))


def test_isdm(q):                                                       # :485-494
    assert q.isdm(np.array([[1, 1], [-1, 1]])) == False
    assert q.isdm(q.to_dm(q.basis(2, 0))) == True
    assert q.isdm(q.sigmax() * q.sigmay()) == False
    assert q.isdm(np.eye(2) * 2) == False


def test_rand_unitary(q):                                               # :496-500
    for N in range(2, 43, 10):
        unitary = q.rand_unitary(N)
        assert_array_almost_equal(npy(unitary @ q.dag(unitary)), np.eye(N))
        assert_array_almost_equal(npy(q.dag(unitary) @ unitary), np.eye(N))


def test_circuit_learning_trace(q):
    """examples/Circuit-Learning.py:47-183 on the CUDA path: value + gradient + Adam known-answer test against
    the losses printed in examples/notebooks/Circuit-Learning.ipynb:195-202."""
    from oracle.examples_ref import Adam
    dt = torch.complex128

    def dev(x):
        return torch.as_tensor(x).to(dt).cuda()

    def rx(phi):
        c, s = torch.cos(phi / 2).to(dt), torch.sin(phi / 2).to(dt)
        return torch.stack([torch.stack([c, -1j * s]), torch.stack([-1j * s, c])])

    def ry(phi):
        c, s = torch.cos(phi / 2).to(dt), torch.sin(phi / 2).to(dt)
        return torch.stack([torch.stack([c, -s]), torch.stack([s, c])])

    def rz(phi):
        z = torch.zeros((), dtype=dt, device=phi.device)
        return torch.stack([torch.stack([torch.exp(-0.5j * phi.to(dt)), z]), torch.stack([z, torch.exp(0.5j * phi.to(dt))])])

    cnot = dev(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]))
    pi4 = torch.tensor(np.pi / 4, dtype=torch.float64, device="cuda")
    op = torch.kron(q.sigmaz(dt), torch.eye(2, dtype=dt, device="cuda"))

    def cost(tx, ty, tz):
        layer0 = torch.kron(q.basis(2, 0, dt), q.basis(2, 0, dt))
        unitary = (torch.kron(ry(pi4), ry(pi4)) @ cnot @ torch.kron(rx(tx), torch.eye(2, dtype=dt, device="cuda"))
                   @ cnot @ torch.kron(ry(ty), rz(tz)))
        return torch.real(q.expect(op, unitary @ layer0))

    opt = Adam([0.0, 0.0, 0.0], 1e-2)
    g = q.grad(cost, argnums=(0, 1, 2))
    printed = []
    for epoch in range(400):
        grads = g(*[torch.tensor(float(x), dtype=torch.float64) for x in opt.x])
        opt.update(epoch, [float(v) for v in grads])
        if epoch % 50 == 49:
            printed.append(float(cost(*[torch.tensor(float(x), dtype=torch.float64, device="cuda") for x in opt.x])))
    np.testing.assert_allclose(printed, G.CIRCUIT_LEARNING_LOSSES, atol=5e-6)
