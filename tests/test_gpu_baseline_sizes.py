"""GPU parity AT THE BASELINE SIZES (dim 512 and 1024), against the CPU oracle, through the C ABI.

Everything the headline learning-steps/s number runs -- the GEMM-chain expm at dim 1024 with s = 4-5, the
power-iteration planner, the blocked Gauss-Jordan inverse (batch >= 3), the series + Newton-Schulz inverse
(batch <= 2), the general and the body-frame skew-Hermitian pull-back, and one full ``HamiltonianLearner`` step
on a bench setting (dim 1024, 512 kets, 38 Pauli terms) -- is compared here with SciPy ``expm`` /
``expm_frechet`` (oracle O1) on the same inputs: <= 1e-12 relative Frobenius error for complex128, <= 1e-5 for
complex64, values and gradients.  One SciPy expm + Frechet pair at dim 1024 is ~2-3 s of CPU, so the oracle
results are computed once per input family and sliced for the batch sizes.
"""
import functools

import numpy as np
import pytest
import torch

import bench
from oracle import examples_ref as ex
from oracle import expm_ref as E
from tests._tol import check, relerr

pytestmark = pytest.mark.gpu

C128, C64 = torch.complex128, torch.complex64
TOL = {C128: 1e-12, C64: 1e-5}
NPDT = {C128: np.complex128, C64: np.complex64}
NMAX = 3            # matrices per input family: batch 1 and 2 take the series inverse, 3 the Gauss-Jordan sweep


@pytest.fixture(scope="module")
def q():
    import qgrad_b200.qgrad_qutip as mod
    assert torch.cuda.is_available()
    return mod


def dev(x, dt):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dt).cuda().contiguous()


def host(t):
    return t.detach().cpu().numpy()


def rel(x, ref):
    return E.rel_fro(np.asarray(x, dtype=np.complex128), np.asarray(ref, dtype=np.complex128))


def bench_pauli_generators(nq, count):
    """A_s = -i t H_s(theta0) of the first ``count`` control settings of bench.py's problem, at ``nq`` qubits
    (nq = 10: exactly the matrices the headline step exponentiates; ||A||_1 ~ 16-20)."""
    xm, zm = bench.pauli_terms(nq)
    rng = np.random.default_rng(0)
    P = len(xm)
    ctrl = rng.uniform(0.5, 1.5, (bench.N_SETTINGS, P))
    theta_true = rng.uniform(-1.0, 1.0, P)
    theta0 = theta_true + 0.2 * rng.standard_normal(P)
    n = 1 << nq
    out = np.zeros((count, n, n), dtype=np.complex128)
    ents = [ex.pauli_entries(int(x), int(z), nq) for x, z in zip(xm, zm)]
    for s in range(count):
        for p, (r, c, v) in enumerate(ents):
            out[s, r, c] += ctrl[s, p] * theta0[p] * v
    return -1j * bench.T_EVOLVE * out


@functools.lru_cache(maxsize=None)
def family(kind, n, dtname):
    """(A, Ybar, Y_ref, Abar_ref) for NMAX matrices; the references are computed in complex128 from the inputs as
    rounded to the test dtype."""
    rng = np.random.default_rng(7000 + n)
    if kind == "pauli":
        A = bench_pauli_generators(int(np.log2(n)), NMAX)
    else:                                                # "gue": SURVEY 8d's named input, H = (X + X^H) / (2 sqrt n), t = 1
        A = -1j * E.rand_hermitian(rng, NMAX, n, True)
    G = E.rand_cotangent(rng, NMAX, n)
    npdt = np.complex128 if dtname == "c128" else np.complex64
    A, G = A.astype(npdt), G.astype(npdt)
    A128, G128 = A.astype(np.complex128), G.astype(np.complex128)
    return A, G, E.expm_scipy(A128), E.expm_vjp(A128, G128)


@pytest.mark.parametrize("dt", [C128, C64])
@pytest.mark.parametrize("kind", ["pauli", "gue"])
@pytest.mark.parametrize("batch", [1, 2, 3])
@pytest.mark.parametrize("n", [512, 1024])
def test_expm_and_vjp_at_baseline_dims(q, n, batch, kind, dt):
    """qg_expm and qg_expm_fwd_vjp at dim 512 / 1024: batch 1 and 2 (series + Newton-Schulz inverse in complex128),
    batch 3 (blocked Gauss-Jordan), on the bench's Pauli Hamiltonians (s = 4-5 after the spectral-radius estimate)
    and on GUE / sqrt(n) (s = 1-2), both dtypes, values and pull-backs vs SciPy"""
    from qgrad_b200.qgrad_qutip import expm_fwd_vjp
    A, G, Yref, Lref = family(kind, n, "c128" if dt == C128 else "c64")
    Y = q.expm(dev(A[:batch], dt))
    check(rel(host(Y), Yref[:batch]), dt, f"expm dim {n} batch {batch} {kind}")
    Y2, Ab, plan = expm_fwd_vjp(dev(A[:batch], dt), dev(G[:batch], dt), return_plan=True)
    check(rel(host(Y2), Yref[:batch]), dt, f"expm_fwd_vjp value dim {n} batch {batch} {kind}")
    check(rel(host(Ab), Lref[:batch]), dt, f"expm_fwd_vjp pull-back dim {n} batch {batch} {kind}")
    s = host(plan)
    assert s.min() >= 0 and s.max() <= 6, s               # the planner's choice is sane (1-norm rule: 5-6 / 6-7)


@pytest.mark.parametrize("dt", [C128, C64])
@pytest.mark.parametrize("batch", [1, 3])
@pytest.mark.parametrize("n", [512, 1024])
def test_skew_body_pullback_at_baseline_dims(q, n, batch, dt):
    """QG_ADJ_SKEW_BODY at dim 512 / 1024 on the bench's Pauli Hamiltonians vs skew(expm_frechet): the exact path of
    the learning step (Z = Ybar U^H in, skew-Hermitian part of Abar out)"""
    from qgrad_b200 import _lib
    from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream
    A, G, Yref, Lref = family("pauli", n, "c128" if dt == C128 else "c64")
    A_d, G_d = dev(A[:batch], dt), dev(G[:batch], dt)
    lib = _lib.load()
    code, ms = _code(dt), 6
    ws = torch.empty(lib.qg_expm_workspace_bytes(code, n, batch, _lib.QG_EXPM_FWD_VJP, ms), dtype=torch.uint8, device="cuda")
    U, S = torch.empty_like(A_d), torch.empty_like(A_d)
    _call("qg_expm_fwd_save", code, _ptr(A_d), _ptr(U), batch, n, _ptr(ws), ws.numel(), ms, _stream())
    check(rel(host(U), Yref[:batch]), dt, f"expm_fwd_save dim {n} batch {batch} pauli")
    Z = (G_d @ U.mH).contiguous()
    _call("qg_expm_bwd_saved", code, _ptr(A_d), _ptr(Z), _ptr(S), batch, n, _lib.QG_ADJ_SKEW_BODY, _ptr(ws), ws.numel(), ms, _stream())
    want = 0.5 * (Lref[:batch] - np.conj(np.swapaxes(Lref[:batch], -1, -2)))
    check(rel(host(S), want), dt, f"skew-body pull-back dim {n} batch {batch} pauli")


def _bench_setting(nq, kets, seed=11):
    """One control setting of bench.py's problem with Haar-random kets and labels from the target Hamiltonian."""
    xm, zm, ctrl, theta_true, theta0 = bench.problem()
    if nq != bench.N_QUBITS:
        xm, zm = bench.pauli_terms(nq)
        ctrl, theta_true, theta0 = ctrl[:, : len(xm)], theta_true[: len(xm)], theta0[: len(xm)]
    n = 1 << nq
    rng = np.random.default_rng(seed)
    psi = rng.standard_normal((1, n, kets)) + 1j * rng.standard_normal((1, n, kets))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    return xm, zm, ctrl[:1], theta_true, theta0, psi


@functools.lru_cache(maxsize=None)
def _learner_case(nq, kets):
    import scipy.linalg
    xm, zm, ctrl, theta_true, theta0, psi = _bench_setting(nq, kets)
    n = 1 << nq
    H = np.zeros((n, n), dtype=np.complex128)
    for p in range(len(xm)):
        r, c, v = ex.pauli_entries(int(xm[p]), int(zm[p]), nq)
        H[r, c] += ctrl[0, p] * theta_true[p] * v
    out = (scipy.linalg.expm(-1j * bench.T_EVOLVE * H) @ psi[0])[None]
    l_ref, g_ref = ex.pauli_learning_loss_and_grad_frechet(theta0, ctrl, xm, zm, nq, bench.T_EVOLVE, psi, out,
                                                           m_total=bench.N_SETTINGS * kets)
    return xm, zm, ctrl, theta0, psi, out, l_ref, g_ref


@pytest.mark.parametrize("graphs", [False, True])
@pytest.mark.parametrize("skew", [True, False])
@pytest.mark.parametrize("nq", [9, 10])
def test_learner_step_at_baseline_size_matches_oracle(q, nq, skew, graphs):
    """one full HamiltonianLearner.loss_and_grad on a bench setting -- dim 1024 (and 512), 512 kets, 38 (34) Pauli
    terms, complex128 -- vs the SciPy oracle: the loss to 1e-12 and thetabar to 1e-12 relative, for the skew-Hermitian
    body-frame pull-back (what bench.py runs) and the general one, launched eagerly and replayed from a CUDA graph"""
    from qgrad_b200 import learning
    kets = bench.KETS_PER_SETTING
    xm, zm, ctrl, theta0, psi, out, l_ref, g_ref = _learner_case(nq, kets)
    L = learning.HamiltonianLearner(nq, xm, zm, ctrl, bench.T_EVOLVE, psi, out, bench.N_SETTINGS * kets,
                                    theta_bound=float(np.max(np.abs(theta0)) * 1.5), skew_pullback=skew)
    theta = torch.tensor(theta0, dtype=torch.float64, device="cuda")
    if graphs:
        L.enable_graphs(theta)
        L.loss_and_grad(theta)                      # capture
    loss, g = L.loss_and_grad(theta)
    # the learner's loss is 1 - (sum over ALL settings)/M; one setting of 8 is present here, as on an 8-rank shard
    tag = f"{'skew-body' if skew else 'general'} pull-back, {'graph replay' if graphs else 'eager'}, dim {1 << nq}"
    check(relerr(float(loss), l_ref), C128, f"learner loss ({tag})")
    check(relerr(host(g), g_ref), C128, f"learner thetabar ({tag})")
    if graphs:
        assert L.graph_kernel_launches > 0


def test_learner_step_dim1024_gauss_jordan_vs_series_inverse(q):
    """the same setting duplicated three times runs the Gauss-Jordan inverse (batch 3) where the single copy runs
    the series inverse: thetabar scales by 3 and agrees to round-off (the 2- and 4-rank shards vs the 8-rank one)"""
    from qgrad_b200 import learning
    nq, kets = 10, 64
    xm, zm, ctrl, theta0, psi, out, l_ref, g_ref = _learner_case(nq, kets)
    theta = torch.tensor(theta0, dtype=torch.float64, device="cuda")
    one = learning.HamiltonianLearner(nq, xm, zm, ctrl, bench.T_EVOLVE, psi, out, bench.N_SETTINGS * kets, theta_bound=2.0)
    l1, g1 = one.loss_and_grad(theta)
    l1, g1 = float(l1), host(g1).copy()
    three = learning.HamiltonianLearner(nq, xm, zm, np.repeat(ctrl, 3, 0), bench.T_EVOLVE, np.repeat(psi, 3, 0),
                                        np.repeat(out, 3, 0), bench.N_SETTINGS * kets, theta_bound=2.0)
    l3, g3 = three.loss_and_grad(theta)
    assert abs(l1 - l_ref) < 1e-12 and np.linalg.norm(g1 - g_ref) <= 1e-12 * np.linalg.norm(g_ref)
    assert abs((1 - float(l3)) - 3 * (1 - l1)) < 1e-12
    assert np.linalg.norm(host(g3) - 3 * g_ref) <= 1e-12 * np.linalg.norm(3 * g_ref)


# ------------------------------------------------------------------------------------ Unitary(H)(t): gradient in t
@pytest.mark.parametrize("dt", [C128, C64])
@pytest.mark.parametrize("n", [2, 8, 24, 48, 256, 1024])
def test_unitary_of_hamiltonian_gradient_in_t(q, n, dt):
    """d/dt of Re <Ybar, exp(-i H t)> through Unitary(H)(t) (the north star's signature) on every kernel family:
    exact derivative Re <Ybar, -i H exp(-i H t)>"""
    rng = np.random.default_rng(60 + n)
    H = E.rand_hermitian(rng, 1, n, True)[0].astype(NPDT[dt])
    Yb = (E.rand_cotangent(rng, 1, n)[0] / n).astype(NPDT[dt])
    t0 = 0.83
    Hd, Ybd = dev(H, dt), dev(Yb, dt)
    f = lambda t: torch.sum(torch.conj(Ybd) * q.Unitary(Hd)(t)).real
    g = q.grad(f)(torch.tensor(t0, dtype=torch.float64 if dt == C128 else torch.float32))
    H128, Y128 = H.astype(np.complex128), Yb.astype(np.complex128)
    U = E.expm_eig_skew(H128, t0)
    want = float(np.real(np.sum(np.conj(Y128) * (-1j * H128 @ U))))
    scale = float(np.linalg.norm(Y128) * np.linalg.norm(H128 @ U))      # the derivative is an inner product of these two
    check(abs(float(g) - want) / scale, dt, f"Unitary(H)(t) d/dt n={n}")
    val = float(f(torch.tensor(t0, dtype=torch.float64 if dt == C128 else torch.float32)))
    check(abs(val - float(np.real(np.sum(np.conj(Y128) * U)))) / float(np.linalg.norm(Y128) * np.sqrt(n)), dt, f"Unitary(H)(t) value n={n}")


def test_unitary_of_hamiltonian_noncontiguous_input_with_grad(q):
    """a transposed (non-contiguous) requires-grad input at n > 32 in complex128 takes the saved-workspace path with
    the Frechet squaring budget: ||A||_1 inside (0.783, 0.95] * 2^k -- the window where the value-only budget is
    one level short -- must give finite, correct gradients"""
    rng = np.random.default_rng(5)
    n = 48
    X = E.rand_cotangent(rng, 1, n)[0]
    for target in (0.9 * 4, 0.8 * 8, 0.94 * 2):
        A = X * (target / np.abs(X).sum(axis=0).max())
        G = E.rand_cotangent(rng, 1, n)[0]
        At = dev(A.T, C128).requires_grad_(True)        # the op sees At.mT == A as a transposed view
        Y = q.expm(At.mT)
        assert not At.mT.is_contiguous()
        (gT,) = torch.autograd.grad(Y, At, grad_outputs=dev(G, C128))
        want = E.expm_vjp(A, G)                          # dL/dA (torch convention); dL/dA^T is its transpose
        got = host(gT).T
        assert np.all(np.isfinite(got))
        assert rel(got, want) < 1e-12
        assert rel(host(Y), E.expm_scipy(A)) < 1e-12


# ------------------------------------------------------------------------------------ the reference's own expm workload
def test_layered_make_unitary_cost_and_gradient(q):
    """examples/Unitary-Learning-no-qgrad.py:112-141,165-191: U(t, tau) = prod_k e^{-i B tau_k} e^{-i A t_k} (d = 2,
    N = 4, the example's sizes), its cost over 80 training kets and the gradient in (t, tau) -- which the reference
    takes by forward differences (:213-241) -- vs the oracle's autograd gradient (1e-12) and its FD gradient (O(eps))"""
    rng = np.random.default_rng(21)
    d, N, M = 2, 4, 80
    A = E.rand_hermitian(rng, 1, d, False)[0]
    B = E.rand_hermitian(rng, 1, d, False)[0]
    params = rng.uniform(-1, 1, 2 * N)
    ins = [(lambda v: v / np.linalg.norm(v))(rng.standard_normal((d, 1)) + 1j * rng.standard_normal((d, 1))) for _ in range(M)]
    Ut = E.expm_scipy(-1j * E.rand_hermitian(rng, 1, d, False)[0])
    outs = [Ut @ k for k in ins]
    U = q.make_unitary(N, params, A, B)
    U_ref = ex.make_unitary(N, params, A, B).numpy()
    assert rel(host(U), U_ref) < 1e-12
    assert rel(host(q.make_unitary(N, params.reshape(2 * N, 1), A, B)), U_ref) < 1e-12       # the reference's (2N, 1) layout
    cost = lambda p: q.layered_unitary_cost(p, ins, outs, N, A, B)
    ref_cost = lambda p: ex.hamiltonian_learning_cost(p, ins, outs, N, A, B)
    assert abs(float(cost(params)) - float(ref_cost(params))) < 1e-12
    g = host(q.grad(cost)(torch.tensor(params)))
    pt = torch.tensor(params, requires_grad=True)
    ref_cost(pt).backward()
    assert np.max(np.abs(g - pt.grad.numpy())) < 1e-12
    fd = ex.fd_gradient(lambda p: float(ref_cost(p)), params, eps=1e-3)
    assert np.max(np.abs(g - fd)) < 5e-3
    with pytest.raises(ValueError):
        q.make_unitary(N, params[:-1], A, B)


@pytest.mark.parametrize("d", [8, 64])
def test_layered_make_unitary_larger_dims(q, d):
    """the same ansatz at d = 8 (shared-memory kernel) and d = 64 (GEMM chain), N = 3: value and gradient vs torch-CPU"""
    rng = np.random.default_rng(31 + d)
    N, M = 3, 16
    A = E.rand_hermitian(rng, 1, d, True)[0]
    B = E.rand_hermitian(rng, 1, d, True)[0]
    params = rng.uniform(-1, 1, 2 * N)
    ins = [(lambda v: v / np.linalg.norm(v))(rng.standard_normal((d, 1)) + 1j * rng.standard_normal((d, 1))) for _ in range(M)]
    Ut = E.expm_scipy(-1j * E.rand_hermitian(rng, 1, d, True)[0])
    outs = [Ut @ k for k in ins]
    cost = lambda p: q.layered_unitary_cost(p, ins, outs, N, A, B)
    ref_cost = lambda p: ex.hamiltonian_learning_cost(p, ins, outs, N, A, B)
    assert abs(float(cost(params)) - float(ref_cost(params))) < 1e-12
    g = host(q.grad(cost)(torch.tensor(params)))
    pt = torch.tensor(params, requires_grad=True)
    ref_cost(pt).backward()
    assert np.max(np.abs(g - pt.grad.numpy())) < 1e-12


def test_expect_batch_mismatch_raises(q):
    """a (B', n, n) operator batch against B != B' states is an error, not a silent read of oper[0]; one state against
    a batch of operators broadcasts"""
    rng = np.random.default_rng(1)
    ops = dev(E.rand_hermitian(rng, 3, 4, True), C128)
    kets = dev(rng.standard_normal((5, 4, 1)) + 0j, C128)
    with pytest.raises(ValueError):
        q.expect(ops, kets)
    one = q.expect(ops, kets[0])
    assert one.shape == (3,)
    for b in range(3):
        assert abs(complex(one[b]) - complex(q.expect(ops[b], kets[0]))) < 1e-14
