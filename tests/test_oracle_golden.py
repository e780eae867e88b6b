"""Pins the CPU oracle against every golden vector / known-answer value the reference's own
tests and notebooks hold for the hot path, and the three independent expm
oracles against each other.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import examples_ref as ex
from oracle import expm_ref as E
from oracle import qgrad_ref as q
from tests.golden import reference_literals as G


def npy(x):
    return x.detach().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def test_squeeze_golden():
    assert np.allclose(npy(q.squeeze(4, G.SQUEEZE_4_Z)), G.SQUEEZE_4)
    # the same 4x4 through the restated Pade algorithm (the one the kernels implement)
    a, ad = npy(q.destroy(4)).astype(np.complex128), npy(q.create(4)).astype(np.complex128)
    op = 0.5 * (np.conj(G.SQUEEZE_4_Z) * a @ a - G.SQUEEZE_4_Z * ad @ ad)
    assert np.allclose(E.pade_expm(op), G.SQUEEZE_4)


def test_displace_golden():
    assert np.allclose(npy(q.Displace(4)(G.DISPLACE_4_ALPHA)), G.DISPLACE_4)
    for N in G.DISPLACE_ZERO_DIMS:
        np.testing.assert_array_almost_equal(npy(q.Displace(N)(0)), np.eye(N))


def test_displace_equals_expm():
    # D(alpha) = exp(alpha a^dag - conj(alpha) a) on the truncated space
    for n, alpha in ((10, 0.3 - 0.4j), (25, -1.1 + 0.2j), (40, 0.7j)):
        a = np.diag(np.sqrt(np.arange(1, n)), 1).astype(np.complex128)
        ref = E.expm_scipy(alpha * a.conj().T - np.conj(alpha) * a)
        assert E.rel_fro(npy(q.Displace(n)(alpha)), ref) < 1e-12


def test_ladder_golden():
    assert np.allclose(npy(q.destroy(3)), G.DESTROY_3)
    assert np.allclose(npy(q.create(3)), G.CREATE_3)
    assert np.allclose(npy(q.dag(q.destroy(3))), npy(q.create(3)))
    np.testing.assert_array_almost_equal(npy(q.destroy(10) @ q.basis(10, 9)), 3.0 * npy(q.basis(10, 8)))
    assert np.allclose(npy(q.create(5) @ q.basis(5, 3)), 2.0 * npy(q.basis(5, 4)))
    with pytest.raises(ValueError):
        q.destroy(2.5)
    with pytest.raises(ValueError):
        q.basis(-1)


def test_dag_golden():
    assert np.array_equal(npy(q.dag(G.DAG_KET)), G.DAG_KET_DAG)
    assert np.array_equal(npy(q.dag(q.basis(2, 0))), [[1.0, 0.0]])


def test_basis_and_to_dm():
    ref = np.zeros((4, 1), dtype=np.complex64)
    ref[2, 0] = 1.0
    assert np.array_equal(npy(q.basis(4, 2)), ref)
    assert npy(q.basis(4, 2)).dtype == np.complex64
    assert np.array_equal(npy(q.basis(20, 20)), np.zeros((20, 1)))      # out-of-range quirk
    dm0 = np.array([[1, 0], [0, 0]], dtype=np.complex64)
    dm1 = np.array([[0, 0], [0, 1]], dtype=np.complex64)
    assert np.array_equal(npy(q.to_dm(q.basis(2, 0))), dm0)
    assert np.array_equal(npy(q.to_dm(q.dag(q.basis(2, 1)))), dm1)
    with pytest.raises(TypeError):
        q.to_dm(np.eye(3))


def test_fidelity_known_answers():
    ket0, ket1 = np.array([[1.0], [0]]), np.array([[0.0], [1]])
    plus, minus = (ket0 + ket1) / np.sqrt(2), (ket0 - ket1) / np.sqrt(2)
    assert q.fidelity(ket0, ket1) == 0.0 and q.fidelity(ket0, ket0) == 1.0
    assert tuple(q.fidelity(ket0, ket0).shape) == (1, 1)
    np.testing.assert_almost_equal(npy(q.fidelity(plus, minus)), 0.0)
    for a, b in ((plus, ket0), (plus, ket1), (ket0, minus), (ket1, minus)):
        np.testing.assert_almost_equal(npy(q.fidelity(a, b)), 0.5)
    rng = np.random.default_rng(1)
    for _ in range(5):
        k = rng.standard_normal((25, 1)) + 1j * rng.standard_normal((25, 1))
        k /= np.linalg.norm(k)
        np.testing.assert_almost_equal(npy(q.fidelity(k, k)), 1.0, decimal=6)
        rho = npy(q.to_dm(k))
        np.testing.assert_almost_equal(float(q.fidelity(rho, rho)), 1.0, decimal=4)
        assert q.fidelity(rho, k).ndim == 0


def test_fidelity_dm_is_diagonal_surrogate():
    # elementwise |.|, so only the diagonals enter
    rng = np.random.default_rng(2)
    a = rng.standard_normal((5, 1)) + 1j * rng.standard_normal((5, 1)); a /= np.linalg.norm(a)
    b = rng.standard_normal((5, 1)) + 1j * rng.standard_normal((5, 1)); b /= np.linalg.norm(b)
    ra, rb = npy(q.to_dm(a)), npy(q.to_dm(b))
    expected = np.sqrt(1 - (0.5 * np.abs(np.diag(ra) - np.diag(rb)).sum()) ** 2)
    np.testing.assert_allclose(float(q.fidelity(ra, rb)), expected, rtol=1e-14)


def test_expect_known_answers():
    for op, vals in ((q.sigmax(), (0.0, 0.0)), (q.sigmay(), (0.0, 0.0)), (q.sigmaz(), (1.0, -1.0))):
        for k, v in zip((0, 1), vals):
            assert q.expect(op, q.basis(2, k)) == v
    rng = np.random.default_rng(3)
    h = rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5)); h = h + h.conj().T
    k = rng.standard_normal((5, 1)) + 1j * rng.standard_normal((5, 1)); k /= np.linalg.norm(k)
    assert abs(complex(q.expect(h, k)) - (k.conj().T @ h @ k)[0, 0]) < 1e-6
    assert np.imag(complex(q.expect(h, npy(q.basis(5, 0)).astype(np.complex128)))) == 0.0
    rho = npy(q.to_dm(k))
    np.testing.assert_allclose(complex(q.expect(h, rho)), np.trace(rho @ h), rtol=1e-13)


def test_coherent_known_answers():
    val = complex(q.expect(q.destroy(G.COHERENT_N), q.coherent(G.COHERENT_N, G.COHERENT_ALPHA)))
    assert abs(val - G.COHERENT_ALPHA) < G.COHERENT_TOL
    for N in G.COHERENT_ZERO_DIMS:
        np.testing.assert_array_almost_equal(npy(q.coherent(N, 0)), npy(q.basis(N, 0)))


@pytest.mark.parametrize("N, params, idx", G.MAKE_ROT_CASES)
def test_make_rot_unitary(N, params, idx):
    r = npy(q._make_rot(N, params, idx))
    np.testing.assert_array_almost_equal(r @ r.conj().T, np.eye(N))
    np.testing.assert_array_almost_equal(r.conj().T @ r, np.eye(N))
    i, j = idx
    th, ph = params
    np.testing.assert_allclose(r[i, j], -np.exp(1j * ph) * np.sin(th), atol=1e-6)
    np.testing.assert_allclose(r[j, i], np.sin(th), atol=1e-6)


def test_unitary_is_unitary_and_validates():
    rng = np.random.default_rng(0)
    for N in G.UNITARY_DIMS:
        h = N * (N - 1) // 2
        u = npy(q.Unitary(N)(rng.uniform(0, 2 * np.pi, h), rng.uniform(0, 2 * np.pi, h),
                             rng.uniform(0, 2 * np.pi, N)))
        np.testing.assert_array_almost_equal(u @ u.conj().T, np.eye(N))
    for N in G.RAND_UNITARY_DIMS:
        u = npy(q.rand_unitary(N, seed=N))
        np.testing.assert_array_almost_equal(u.conj().T @ u, np.eye(N))
    with pytest.raises(ValueError):
        q.Unitary(3)(np.zeros(3), np.zeros(3), np.zeros(2))
    with pytest.raises(ValueError):
        q.Unitary(3)(np.zeros(3), np.zeros(2), np.zeros(3))
    with pytest.raises(ValueError):
        q.Unitary(3)(np.zeros(2), np.zeros(2), np.zeros(3))


def test_predicates():
    assert q.isherm(q.sigmax()) and q.isherm(q.sigmay()) and q.isherm(q.sigmaz())
    assert not q.isherm(np.arange(9).reshape(3, 3))
    assert not q.isdm(np.array([[1, 1], [-1, 1]]))
    assert q.isdm(q.to_dm(q.basis(2, 0)))
    assert not q.isdm(npy(q.sigmax()) * npy(q.sigmay()))
    assert not q.isdm(np.eye(2) * 2)
    k = npy(q.rand_ket(7, seed=3))
    np.testing.assert_almost_equal(np.linalg.norm(k), 1.0, decimal=5)
    assert q.isket(k) and not q.isbra(k) and q.isbra(npy(q.dag(k)))


def test_circuit_learning_trace():
    """value + gradient + Adam known-answer test (Circuit-Learning.ipynb:195-202)."""
    got = ex.circuit_learning_trace()
    np.testing.assert_allclose(got, G.CIRCUIT_LEARNING_LOSSES, atol=5e-6)


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64])
@pytest.mark.parametrize("scaled", [True, False])
def test_expm_oracles_agree(n, scaled):
    rng = np.random.default_rng(n)
    H = E.rand_hermitian(rng, 2, n, scaled)
    A, Gc = -1j * H, E.rand_cotangent(rng, 2, n)
    y1 = E.expm_scipy(A)
    assert E.rel_fro(np.stack([E.expm_eig_skew(h) for h in H]), y1) < 1e-13
    assert E.rel_fro(E.expm_torch(A), y1) < 1e-13
    assert E.rel_fro(E.pade_expm(A), y1) < 1e-13
    l1 = E.expm_vjp(A, Gc)
    assert E.rel_fro(E.expm_vjp_torch(A, Gc), l1) < 1e-12
    assert E.rel_fro(np.stack([E.frechet_eig_skew(-h, g) for h, g in zip(H, Gc)]), l1) < 1e-13
    yf, lf = E.pade_expm_vjp(A, Gc)
    assert E.rel_fro(yf, y1) < 1e-13 and E.rel_fro(lf, l1) < 1e-13
    # single precision restatement within the north-star's 1e-5
    y32, l32 = E.pade_expm_vjp(A.astype(np.complex64), Gc.astype(np.complex64))
    assert E.rel_fro(y32, y1) < 1e-5 and E.rel_fro(l32, l1) < 1e-5
    assert E.rel_fro(E.pade_expm(A.astype(np.complex64)), y1) < 1e-5


def test_expm_general_nonnormal():
    rng = np.random.default_rng(7)
    A = 0.7 * E.rand_cotangent(rng, 3, 12)
    Gc = E.rand_cotangent(rng, 3, 12)
    assert E.rel_fro(E.pade_expm(A), E.expm_scipy(A)) < 1e-13
    y, l = E.pade_expm_vjp(A, Gc)
    assert E.rel_fro(l, E.expm_vjp(A, Gc)) < 1e-12
    assert E.rel_fro(l, E.expm_vjp_torch(A, Gc)) < 1e-12


def test_jax_style_grad_convention():
    # real loss of a complex argument: JAX returns the conjugate of torch's .grad
    z0 = torch.tensor(0.3 - 0.7j, dtype=torch.complex128)
    f = lambda z: torch.real(z * z * (1 + 2j))
    g = q.jax_style_grad(f)(z0)
    # d/dx - i d/dy of Re((x+iy)^2 (1+2i)) at z0
    eps = 1e-6
    dx = (f(z0 + eps) - f(z0 - eps)) / (2 * eps)
    dy = (f(z0 + 1j * eps) - f(z0 - 1j * eps)) / (2 * eps)
    assert abs(complex(g) - complex(dx, -dy)) < 1e-8


def test_example_costs_run_and_differentiate():
    rng = np.random.default_rng(5)
    # Qubit rotation: F = (1 + <sigma_z>)/2
    ket1 = q.basis(2, 1, torch.complex128)
    f = ex.qubit_rotation_cost(0.3, 1.1, -0.4, ket1)
    sz = ex.rot_expect_sigmaz(0.3, 1.1, -0.4, ket1)
    np.testing.assert_allclose(float(f), (1 + float(sz)) / 2, rtol=1e-13)
    # SNAP cost gradient exists for complex alphas
    N = 10
    init = q.basis(N, 0, torch.complex128)
    target = (np.sqrt(3) * q.basis(N, 3, torch.complex128) + q.basis(N, 9, torch.complex128)) / 2
    alphas = torch.tensor([1, 0.2, 0.1 - 1j], dtype=torch.complex128)
    thetas = torch.tensor(np.stack([ex.pad_thetas(N, p) for p in ([0.5], [0.0, 0.5], [0.1, 0.0, 0.0, 0.1])]))
    g = q.jax_style_grad(lambda a, t: ex.snap_cost((a, t), init, target), argnums=(0, 1))(alphas, thetas)
    assert g[0].shape == (3,) and g[1].shape == (3, N) and torch.isfinite(g[1]).all()
    # Hamiltonian-learning gradient: autograd vs the reference's finite differences
    d, NN = 2, 2
    A = E.rand_hermitian(rng, 1, d, False)[0]; B = E.rand_hermitian(rng, 1, d, False)[0]
    params = rng.uniform(0, 1, 2 * NN)
    kets = [(lambda k: k / np.linalg.norm(k))(rng.standard_normal((d, 1)) + 1j * rng.standard_normal((d, 1))) for _ in range(6)]
    Ut = E.expm_scipy(-1j * E.rand_hermitian(rng, 1, d, False)[0])
    outs = [Ut @ k for k in kets]
    cost = lambda p: ex.hamiltonian_learning_cost(p, kets, outs, NN, A, B)
    g_ad = npy(q.jax_style_grad(cost)(torch.tensor(params)))
    g_fd = ex.fd_gradient(cost, params, eps=1e-6)
    np.testing.assert_allclose(g_ad, g_fd, atol=1e-5)


# ---------------------------------------------------------------------------------- structured variants
@pytest.mark.parametrize("n,scale,s", [(6, 0.3, 0), (12, 2.0, 2), (24, 9.0, 5)])
def test_skew_body_pullback_identity(n, scale, s):
    """for skew-Hermitian A the skew-Hermitian part of L(A^H, Ybar) equals the body-frame recursion on
    Z = Ybar exp(A)^H that the CUDA path runs (QG_ADJ_SKEW_BODY); checked against O1 (SciPy expm_frechet)"""
    rng = np.random.default_rng(n)
    H = E.rand_hermitian(rng, 1, n, False)[0]
    H = scale * H / np.abs(np.linalg.eigvalsh(H)).max()
    A = -1j * H
    Yb = E.rand_cotangent(rng, 1, n)[0]
    Abar = E.expm_vjp(A[None], Yb[None])[0]
    want = 0.5 * (Abar - Abar.conj().T)
    U = E.expm_scipy(A[None])[0]
    S = E.skew_body_pullback(A, Yb @ U.conj().T, s)
    assert E.rel_fro(S, want) < 1e-13
    assert np.max(np.abs(S + S.conj().T)) < 1e-13 * np.max(np.abs(S))
    # and it carries everything a real parameter of a Hermitian generator sees: <Abar, dA> = <S, dA> for skew dA
    dA = -1j * E.rand_hermitian(rng, 1, n, False)[0]
    assert abs(np.real(np.vdot(Abar, dA)) - np.real(np.vdot(S, dA))) < 1e-12 * np.linalg.norm(Abar) * np.linalg.norm(dA)


def test_series_inverse_of_pade_denominator():
    """the Taylor coefficients of 1/q_7 the device uses, and one Newton-Schulz step: residual ~1e-7 -> ~1e-14 for
    every matrix the scaling admits (||A||_1 <= 0.783 general, rho <= 0.783 normal)"""
    from fractions import Fraction as F
    b = [F(int(x)) for x in E.PADE_B[7]]
    q = [b[k] / b[0] * (-1) ** k for k in range(8)]
    d = [F(1)]
    for k in range(1, 8):
        d.append(-sum(q[j] * d[k - j] for j in range(1, k + 1)))
    assert np.allclose([float(x) for x in d], E.INV_Q7_SERIES, rtol=1e-15, atol=0)
    rng = np.random.default_rng(3)
    n = 40
    X = E.rand_cotangent(rng, 1, n)[0]
    H = E.rand_hermitian(rng, 1, n, False)[0]
    cases = [0.783 * X / np.abs(X).sum(axis=0).max(), -0.783j * H / np.abs(np.linalg.eigvalsh(H)).max(), 1e-9 * X]
    for A in cases:
        Qi, Q = E.series_inverse_q7(A)
        assert np.linalg.norm(Qi @ Q - np.eye(n)) < 1e-13 * np.sqrt(n)
        assert E.rel_fro(Qi, np.linalg.inv(Q)) < 1e-13


def test_spectral_radius_estimate_is_a_tight_lower_bound():
    """8 power iterations with A^6 from the fixed start vector: never above rho, within 2.5 % of it -- on a
    nearest-neighbour Pauli Hamiltonian (the learning step's family), a GUE matrix, a clustered spectrum and an
    isolated top eigenvalue; the planner inflates the estimate by 3 %"""
    rng = np.random.default_rng(11)
    nq = 6
    n = 1 << nq
    xm, zm = [], []
    for b in range(nq - 1):
        m = (1 << b) | (1 << (b + 1))
        xm += [0, m]; zm += [m, 0]
    for s in range(nq):
        xm += [1 << s, 0]; zm += [0, 1 << s]
    Hp = sum(c * ex.pauli_dense(int(x), int(z), nq) for c, x, z in zip(rng.uniform(-1.5, 1.5, len(xm)), xm, zm))
    G = E.rand_hermitian(rng, 1, n, True)[0]
    Qm, _ = np.linalg.qr(E.rand_cotangent(rng, 1, n)[0])
    clustered = (Qm * np.linspace(0.95, 1.0, n)) @ Qm.conj().T
    isolated = (Qm * np.r_[np.linspace(0, 0.8, n - 1), 1.0]) @ Qm.conj().T
    for H in (Hp, G, clustered, isolated):
        A = -1j * H
        A2 = A @ A
        rho = np.abs(np.linalg.eigvalsh(H)).max()
        est = E.spectral_radius_estimate(A2 @ A2 @ A2, 6)
        assert all(e <= rho * (1 + 1e-12) for e in est)
        assert est[-1] >= 0.975 * rho and 1.03 * est[-1] >= rho


# ------------------------------------------------------------------------------------ generators (oracle/rand_ref.py)
def test_philox_restatement_matches_random123_known_answers():
    """the NumPy Philox4x32-10 the device generators are compared with reproduces Random123's kat_vectors"""
    from oracle import rand_ref as R
    for ctr, key, want in R.PHILOX4X32_10_KAT:
        got = tuple(int(w[0]) for w in R.philox4x32_10([np.array([c]) for c in ctr], key))
        assert got == want
    # vectorised evaluation = elementwise evaluation
    idx = np.array([0, 1, 2 ** 32 - 1, 2 ** 32, 2 ** 40 + 3], dtype=np.uint64)
    w = R.draw(7, R.STREAM_KET, idx)
    for i, ix in enumerate(idx):
        w1 = R.draw(7, R.STREAM_KET, np.array([ix], dtype=np.uint64))
        assert all(int(a[i]) == int(b[0]) for a, b in zip(w, w1))


def test_generator_oracle_distributions():
    from oracle import rand_ref as R
    k = R.rand_ket(64, 50, 3)
    assert np.allclose(np.linalg.norm(k, axis=1), 1.0)
    kr = R.rand_ket(4, 9, 3, kind=1)
    assert np.array_equal(kr.real, kr.imag) and np.all(kr.real > 0)
    H = R.rand_hermitian(3, 16, 11, scale=0.25)
    assert np.array_equal(H, np.conj(np.swapaxes(H, -1, -2)))
    assert np.array_equal(R.rand_hermitian(1, 16, 11, first=2, scale=0.25)[0], H[2])
    x = R.normal_pair(R.draw(5, R.STREAM_HERM, np.arange(200000)), False)
    assert abs(x.real.mean()) < 0.01 and abs(x.real.std() - 1) < 0.01 and abs(x.imag.std() - 1) < 0.01
    u = R.rand_uniform(100000, 1, 0.0, 1.0)
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
