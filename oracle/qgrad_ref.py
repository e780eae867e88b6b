"""Torch-CPU restatement of ``qgrad/qgrad_qutip.py`` (reference, 620 lines, JAX).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the reference lines it
follows.  Differences from the reference that are deliberate and documented:

* arrays are ``torch`` CPU tensors (``torch.autograd`` stands in for ``jax.grad``; use
  :func:`jax_style_grad` to get JAX's complex-cotangent convention);
* precision: the reference hard-codes ``complex64`` in its builders and JAX (without
  ``jax_enable_x64``, which nothing in the reference sets) truncates every input to 32-bit.
  Builders here default to that dtype and take ``dtype=`` so the complex128 parity target of
  the north star can be checked with the same code; functions of user arrays compute in the
  precision of their inputs;
* RNG (``rand_ket`` / ``rand_dm`` / ``rand_unitary``): JAX's threefry stream cannot be
  reproduced without JAX, so these draw from ``numpy.random.default_rng(seed)`` while keeping
  the reference's structure (same uniform sample used for real and imaginary part, etc.).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg
import torch

C64, C128 = torch.complex64, torch.complex128


def _t(x, dtype=None):
    """``jnp.asarray`` analogue: accept numpy / python / torch, keep or set precision."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x


def _real_dtype(dtype):
    return torch.float32 if dtype == C64 else torch.float64


# ----------------------------------------------------------------------------- predicates
def isket(state):
    """qgrad_qutip.py:333-342 -- ``shape[1] == 1`` (1-D input raises IndexError)."""
    return state.shape[1] == 1


def isbra(state):
    """qgrad_qutip.py:345-354 -- ``shape[0] == 1``."""
    return state.shape[0] == 1


def dag(state):
    """qgrad_qutip.py:319-330 -- conjugate transpose."""
    state = _t(state)
    return torch.conj(state.transpose(-1, -2)).resolve_conj()


def isherm(oper):
    """qgrad_qutip.py:357-367 -- exact elementwise equality with the adjoint."""
    oper = _t(oper)
    return bool(torch.all(oper == dag(oper)))


def isdm(mat):
    """qgrad_qutip.py:370-396 -- not ket/bra, exactly Hermitian, |Re tr - 1| <= 1e-9,
    eigenvalues >= -1e-6."""
    mat = _t(mat)
    if isket(mat) or isbra(mat) or not isherm(mat):
        return False
    if not np.allclose(float(torch.real(torch.trace(mat.to(C128)))), 1.0, atol=1e-9):
        return False
    evals = np.linalg.eig(mat.to(C128).numpy())[0]
    for ev in evals:
        if ev.real < 0 and not np.allclose(ev, 0, atol=1e-6):
            return False
    return True


# ----------------------------------------------------------------------------- constants
def sigmax(dtype=C64):
    """qgrad_qutip.py:67-75."""
    return torch.tensor([[0.0, 1.0], [1.0, 0.0]], dtype=dtype)


def sigmay(dtype=C64):
    """qgrad_qutip.py:78-89."""
    return torch.tensor([[0.0, -1.0j], [1.0j, 0.0]], dtype=dtype)


def sigmaz(dtype=C64):
    """qgrad_qutip.py:92-101."""
    return torch.tensor([[1.0, 0.0], [0.0, -1.0]], dtype=dtype)


def _check_int_dim(N):
    if isinstance(N, bool) or not isinstance(N, (int, np.integer)):
        raise ValueError("Hilbert space dimension must be an integer value")


def destroy(N, dtype=C64):
    """qgrad_qutip.py:104-121 -- super-diagonal sqrt(1..N-1), evaluated in float32 first
    (``jnp.sqrt(jnp.arange(1, N, dtype=float32))``) as the reference does."""
    _check_int_dim(N)
    data = np.sqrt(np.arange(1, N, dtype=np.float32))
    mat = np.zeros((N, N))
    for k in range(N - 1):
        mat[k, k + 1] = data[k]
    return torch.as_tensor(mat).to(dtype)


def create(N, dtype=C64):
    """qgrad_qutip.py:137-152 -- sub-diagonal sqrt(1..N-1) (float32 sqrt)."""
    _check_int_dim(N)
    data = np.sqrt(np.arange(1, N, dtype=np.float32))
    mat = np.zeros((N, N))
    for k in range(N - 1):
        mat[k + 1, k] = data[k]
    return torch.as_tensor(mat).to(dtype)


def basis(N, n=0, dtype=C64):
    """qgrad_qutip.py:283-299 -- (N,1) Fock state.  ``n >= N`` silently yields zeros because
    JAX drops out-of-bounds scatters (relied on by test_qgrad_qutip.py:188)."""
    if isinstance(N, bool) or not isinstance(N, (int, np.integer)) or N < 0:
        raise ValueError("N must be integer N >= 0")
    out = torch.zeros((N, 1), dtype=dtype)
    if -N <= n < N:
        out[n, 0] = 1.0
    return out


# ----------------------------------------------------------------------------- state kernels
def to_dm(state):
    """qgrad_qutip.py:399-421 -- ket test first, then bra, else TypeError."""
    state = _t(state)
    if isket(state):
        return state @ dag(state)
    if isbra(state):
        return dag(state) @ state
    raise TypeError(
        "Input is neither a ket, nor a bra. First dimension of a bra should be 1. Eg: (1, 4)."
        " Second dimension of a ket should be 1. Eg: (4, 1)"
    )


def _promote(a, b):
    dt = torch.promote_types(a.dtype, b.dtype)
    if not dt.is_complex:
        dt = C128 if dt == torch.float64 else C64
    return a.to(dt), b.to(dt)


def _fidelity_ket(a, b):
    """qgrad_qutip.py:36-47 -- |a^dagger b|^2, shape (1,1), real."""
    a, b = _promote(_t(a), _t(b))
    return torch.abs(dag(a) @ b) ** 2


def _fidelity_dm(a, b):
    """qgrad_qutip.py:50-64 -- sqrt(1 - (0.5*trace(|rho1-rho2|))^2) with an ELEMENTWISE
    abs, i.e. only the two diagonals enter.  Reproduced as written."""
    a, b = _promote(_t(a), _t(b))
    tr_dist = 0.5 * torch.trace(torch.abs(a - b))
    return torch.sqrt(1 - tr_dist ** 2)


def fidelity(a, b):
    """qgrad_qutip.py:11-33 -- ket,ket -> (1,1); otherwise promote kets with to_dm."""
    a, b = _t(a), _t(b)
    if isket(a) and isket(b):
        return _fidelity_ket(a, b)
    if isket(a):
        a = to_dm(a)
    if isket(b):
        b = to_dm(b)
    return _fidelity_dm(a, b)


def _expect_dm(oper, state):
    """qgrad_qutip.py:188-195 -- trace(rho @ oper)."""
    oper, rho = _promote(_t(oper), _t(state))
    return torch.trace(rho @ oper)


def _expect_ket(oper, state):
    """qgrad_qutip.py:198-203 -- vdot(ket^T, oper @ ket): conjugates the first argument."""
    oper, ket = _promote(_t(oper), _t(state))
    return torch.sum(torch.conj(ket).reshape(-1) * (oper @ ket).reshape(-1))


def expect(oper, state):
    """qgrad_qutip.py:166-185 -- density-matrix branch iff ``state.shape[1] >= 2``."""
    state = _t(state)
    if state.shape[1] >= 2:
        return _expect_dm(oper, state)
    return _expect_ket(oper, state)


# ----------------------------------------------------------------------------- builders
class Displace:
    """qgrad_qutip.py:216-261 -- D(alpha) through the eigendecomposition of the real
    symmetric tridiagonal T (zero diagonal, off-diagonal (2(k%2)-1)*sqrt(k), k=1..n-1)."""

    def __init__(self, n, dtype=C128):
        k = np.arange(1, n)
        sym = (2.0 * (k % 2) - 1) * np.sqrt(k)
        mat = np.zeros((n, n), dtype=np.float64)
        for q in range(n - 1):
            mat[q + 1, q] = sym[q]
            mat[q, q + 1] = sym[q]
        evals, evecs = np.linalg.eigh(mat)
        self.n = n
        self.dtype = dtype
        rdt = _real_dtype(dtype)
        self.evals = torch.as_tensor(evals).to(rdt)
        self.evecs = torch.as_tensor(evecs).to(dtype)
        self.range = torch.arange(n)
        self.t_scale = torch.as_tensor(1j ** (np.arange(n) % 2)).to(dtype)

    def __call__(self, alpha):
        alpha = _t(alpha, self.dtype)
        r = torch.abs(alpha)
        # jnp.where(alpha == 0, t_scale, t_scale * (alpha/|alpha|) ** -range), qgrad_qutip.py:253-257
        if float(r) == 0.0:
            transform = self.t_scale
        else:
            transform = self.t_scale * (alpha / r) ** (-self.range.to(_real_dtype(self.dtype)))
        evecs = transform[:, None] * self.evecs
        diag = torch.exp(1j * r * self.evals)
        return torch.conj(evecs) @ (diag[:, None] * evecs.transpose(0, 1))


def coherent(N, alpha, dtype=C128):
    """qgrad_qutip.py:302-316 -- D(alpha)|0>."""
    return Displace(N, dtype)(alpha) @ basis(N, 0, dtype)


def squeeze(N, z):
    """qgrad_qutip.py:264-280 -- SciPy expm of 0.5*(conj(z) a^2 - z a^dagger^2), built from
    the complex64 ladder operators; returns a NumPy array, not differentiable."""
    a = destroy(N).numpy()
    ad = create(N).numpy()
    op = 0.5 * (np.conj(z) * np.linalg.matrix_power(a, 2) - z * np.linalg.matrix_power(ad, 2))
    return scipy.linalg.expm(op)


def _make_rot(N, params, idx, dtype=C64):
    """qgrad_qutip.py:424-459 -- identity with the (i,i),(i,j),(j,i),(j,j) entries replaced."""
    i, j = idx
    theta, phi = params
    rdt = _real_dtype(dtype)
    theta, phi = _t(theta, rdt), _t(phi, rdt)
    rot = torch.eye(N, dtype=dtype)
    e = torch.exp(1j * phi.to(dtype))
    # four successive .at[].set writes (later writes win); index_put keeps autograd intact
    rot[i, i] = e * torch.cos(theta)
    rot[i, j] = -e * torch.sin(theta)
    rot[j, i] = torch.sin(theta).to(dtype)
    rot[j, j] = torch.cos(theta).to(dtype)
    return rot


make_rot = _make_rot


class Unitary:
    """qgrad_qutip.py:462-555 -- diag(e^{i omega}) * prod_{i=2..N} prod_{j=1..i-1}
    R(-theta_k, -phi_k; (i-1, j-1)), k running row-major over (i, j); dense N x N products
    exactly as the reference forms them."""

    def __init__(self, N, dtype=C64):
        self.N = N
        self.dtype = dtype

    def __call__(self, thetas, phis, omegas):
        rdt = _real_dtype(self.dtype)
        thetas, phis, omegas = _t(thetas, rdt), _t(phis, rdt), _t(omegas, rdt)
        N = self.N
        if omegas.shape[0] != N:
            raise ValueError("The dimension of omegas should be the same as the unitary")
        if phis.shape[0] != thetas.shape[0]:
            raise ValueError("Number of phi and theta rotation parameters should be the same")
        if phis.shape[0] != N * (N - 1) / 2 or thetas.shape[0] != N * (N - 1) / 2:
            raise ValueError(
                "Size of each of the rotation parameters should be N * (N - 1) / 2, where N "
                "is the size of the unitary matrix")
        diagonal = torch.diag(torch.exp(1j * omegas.to(self.dtype)))
        rotation = torch.eye(N, dtype=self.dtype)
        k = 0
        for i in range(2, N + 1):
            for j in range(1, i):
                rotation = rotation @ _make_rot(N, (-thetas[k], -phis[k]), (i - 1, j - 1), self.dtype)
                k += 1
        return diagonal @ rotation


# ----------------------------------------------------------------------------- RNG stand-ins
def rand_ket(N, seed=None, dtype=C64):
    """qgrad_qutip.py:558-573 -- uniform(key)+1j*uniform(key) with the SAME key for both
    parts (so re == im), normalised.  numpy Generator replaces threefry."""
    if seed is None:
        seed = np.random.randint(1000)
    u = np.random.default_rng(seed).uniform(size=(N, 1))
    ket = u + 1j * u
    return torch.as_tensor(ket / np.linalg.norm(ket)).to(dtype)


def rand_dm(N, seed=None, dtype=C64):
    """qgrad_qutip.py:576-591 -- rank-1 |psi><psi| of rand_ket."""
    return to_dm(rand_ket(N, seed, dtype))


def rand_unitary(N, seed=None, dtype=C64):
    """qgrad_qutip.py:594-620 -- Unitary(N) on N^2 uniform(0, 2pi) angles split
    N(N-1)/2, N(N-1)/2, N."""
    if seed is None:
        seed = np.random.randint(1000)
    params = np.random.default_rng(seed).uniform(0.0, 2 * math.pi, size=(N * N,))
    h = N * (N - 1) // 2
    return Unitary(N, dtype)(params[:h], params[h:2 * h], params[2 * h:])


# ----------------------------------------------------------------------------- jax.grad stand-in
def jax_style_grad(fun, argnums=0):
    """``jax.grad`` analogue over torch-CPU autograd.

    For a real-valued ``fun`` JAX returns dL/dx - i dL/dy for a complex argument x + iy, the
    complex conjugate of torch's ``.grad``.  Real arguments are
    unaffected.  Arguments may be tensors or (nested) lists/tuples of tensors.
    """
    single = isinstance(argnums, int)
    nums = (argnums,) if single else tuple(argnums)

    def leafify(x):
        if isinstance(x, (list, tuple)):
            return type(x)(leafify(v) for v in x)
        x = _t(x)
        if not (x.dtype.is_floating_point or x.dtype.is_complex):
            x = x.to(torch.float64)
        return x.detach().clone().requires_grad_(True)

    def flat(x):
        if isinstance(x, (list, tuple)):
            return [l for v in x for l in flat(v)]
        return [x]

    def rebuild(x, it):
        if isinstance(x, (list, tuple)):
            return type(x)(rebuild(v, it) for v in x)
        g = next(it)
        return torch.conj(g).resolve_conj() if g.is_complex() else g

    def wrapped(*args):
        args = list(args)
        for k in nums:
            args[k] = leafify(args[k])
        out = fun(*args)
        leaves = [l for k in nums for l in flat(args[k])]
        grads = torch.autograd.grad(out, leaves, allow_unused=True)
        grads = [torch.zeros_like(l) if g is None else g for g, l in zip(grads, leaves)]
        it = iter(grads)
        res = tuple(rebuild(args[k], it) for k in nums)
        return res[0] if single else res

    return wrapped
