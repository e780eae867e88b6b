"""expm / Frechet-derivative oracles.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference's only matrix exponential is SciPy's (``qgrad/qgrad_qutip.py:7,280`` in
``squeeze``; ``examples/Unitary-Learning-no-qgrad.py:55,137`` in ``make_unitary``), a
third-party dependency that is not vendored (``requirements.txt:5`` ``scipy>=1.3``; CI pins
``scipy==1.3``, ``.github/workflows/main.yml:30``).  The north star's
``jax.scipy.linalg.expm`` has no call site in the reference and JAX is absent here, so parity to
it is *unpinned*; three independent implementations anchor the numbers instead:

* **O1** SciPy 1.18.1 ``scipy.linalg.expm`` / ``expm_frechet`` (Al-Mohy & Higham 2009).
* **O2** Hermitian eigendecomposition: exp(-iHt) = V e^{-i lambda t} V^H, and the
  Daleckii-Krein formula for the Frechet derivative of a normal matrix.
* **O3** ``torch.linalg.matrix_exp`` on CPU with autograd through the whole cost.

plus **P**, a NumPy restatement of the published Higham-2005 [m/m] Pade
scaling-and-squaring algorithm (the one ``jax.scipy.linalg.expm`` implements: orders
3/5/7/9/13 for 64-bit, 3/5/7 for 32-bit, ``solve(-U+V, U+V)`` with pivoted LU) together with
the Al-Mohy/Higham block-triangular Frechet recurrences.  P is what the CUDA kernels
implement step for step; tests pin P against O1-O3.

Conventions: ``expm_vjp(A, Ybar)`` is the reverse-mode pull-back in torch's convention,
``Abar = L(A^H, Ybar)``; ``jax`` convention is ``L(A^T, Ybar)``.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

# Pade numerator coefficients b_0..b_m (Higham 2005, Table 10.4 / "The scaling and squaring
# method for the matrix exponential revisited").
PADE_B = {
    3: (120.0, 60.0, 12.0, 1.0),
    5: (30240.0, 15120.0, 3360.0, 420.0, 30.0, 1.0),
    7: (17297280.0, 8648640.0, 1995840.0, 277200.0, 25200.0, 1512.0, 56.0, 1.0),
    9: (17643225600.0, 8821612800.0, 2075673600.0, 302702400.0, 30270240.0, 2162160.0,
        110880.0, 3960.0, 90.0, 1.0),
    13: (64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0,
         129060195264000.0, 10559470521600.0, 670442572800.0, 33522128640.0, 1323241920.0,
         40840800.0, 960960.0, 16380.0, 182.0, 1.0),
}
# theta_m: largest ||A||_1 for which the order-m approximant is accurate to unit round-off.
THETA_F64 = {3: 1.495585217958292e-2, 5: 2.539398330063230e-1, 7: 9.504178996162932e-1,
             9: 2.097847961257068e0, 13: 5.371920351148152e0}
THETA_F32 = {3: 4.258730016922831e-1, 5: 1.880152677804762e0, 7: 3.925724783138660e0}
# ell_m of Al-Mohy & Higham 2009 (Table 6.1): the tighter bounds under which BOTH exp and its
# Frechet derivative are accurate.  The fused value+VJP path uses these.
ELL_F64 = {3: 1.08e-2, 5: 2.00e-1, 7: 7.83e-1, 9: 1.78e0, 13: 4.74e0}
# No published single-precision ell table: use theta_m(f32) shrunk by the same ~0.8 ratio the
# double-precision tables show (ell_m/theta_m = 0.72..0.88).  Verified <= 1e-6 in tests.
ELL_F32 = {3: 3.4e-1, 5: 1.5e0, 7: 3.14e0}


def pade_plan(norm1: float, single: bool, frechet: bool):
    """(m, s) for a matrix of 1-norm ``norm1``: smallest order whose bound holds, else the
    top order with s = ceil(log2(norm/bound)) squarings."""
    tab = (ELL_F32 if frechet else THETA_F32) if single else (ELL_F64 if frechet else THETA_F64)
    orders = sorted(tab)
    for m in orders:
        if norm1 <= tab[m]:
            return m, 0
    top = orders[-1]
    s = max(0, int(np.ceil(np.log2(norm1 / tab[top]))))
    return top, s


# ------------------------------------------------------------------------------ O1: SciPy
def expm_scipy(A):
    A = np.asarray(A)
    if A.ndim == 2:
        return scipy.linalg.expm(A)
    return np.stack([scipy.linalg.expm(a) for a in A.reshape(-1, *A.shape[-2:])]).reshape(A.shape)


def frechet_scipy(A, E):
    """L(A, E) (and exp(A)) by Al-Mohy/Higham, one pair at a time."""
    A, E = np.asarray(A, dtype=np.complex128), np.asarray(E, dtype=np.complex128)
    if A.ndim == 2:
        return scipy.linalg.expm_frechet(A, E, compute_expm=True)
    ys, ls = [], []
    for a, e in zip(A.reshape(-1, *A.shape[-2:]), E.reshape(-1, *E.shape[-2:])):
        y, l = scipy.linalg.expm_frechet(a, e, compute_expm=True)
        ys.append(y)
        ls.append(l)
    return np.stack(ys).reshape(A.shape), np.stack(ls).reshape(A.shape)


def expm_vjp(A, Ybar, convention="torch"):
    """Reverse-mode pull-back of Y = exp(A).  torch: L(A^H, Ybar); jax: L(A^T, Ybar)."""
    A = np.asarray(A)
    At = np.swapaxes(A, -1, -2)
    X = np.conj(At) if convention == "torch" else At
    return frechet_scipy(X, Ybar)[1]


# ------------------------------------------------------------------------------ O2: eigendecomposition
def expm_eig_skew(H, t=1.0):
    """exp(-i H t) for Hermitian H via eigh."""
    lam, V = np.linalg.eigh(np.asarray(H, dtype=np.complex128))
    return (V * np.exp(-1j * lam * t)) @ V.conj().T


def frechet_eig_skew(H, E, t=1.0):
    """L(A, E) at A = -iHt (normal) via Daleckii-Krein:
    V [ (V^H E V) o Phi ] V^H, Phi_ij = (e^{mu_i} - e^{mu_j}) / (mu_i - mu_j)."""
    lam, V = np.linalg.eigh(np.asarray(H, dtype=np.complex128))
    mu = -1j * lam * t
    emu = np.exp(mu)
    dmu = mu[:, None] - mu[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        phi = (emu[:, None] - emu[None, :]) / dmu
    # (near-)degenerate pairs: divided difference -> e^{(mu_i+mu_j)/2} * sinhc((mu_i-mu_j)/2)
    small = np.abs(dmu) < 1e-8
    mean = 0.5 * (mu[:, None] + mu[None, :])
    phi = np.where(small, np.exp(mean) * (1 + dmu ** 2 / 24.0), phi)
    G = V.conj().T @ np.asarray(E, dtype=np.complex128) @ V
    return V @ (G * phi) @ V.conj().T


# ------------------------------------------------------------------------------ O3: torch autograd
def expm_torch(A):
    import torch
    return torch.linalg.matrix_exp(torch.as_tensor(np.asarray(A))).numpy()


def expm_vjp_torch(A, Ybar):
    """torch-CPU autograd pull-back (torch convention)."""
    import torch
    a = torch.as_tensor(np.asarray(A)).clone().requires_grad_(True)
    y = torch.linalg.matrix_exp(a)
    (g,) = torch.autograd.grad(y, a, grad_outputs=torch.as_tensor(np.asarray(Ybar)).to(y.dtype))
    return g.resolve_conj().numpy()


# ------------------------------------------------------------------------------ P: the algorithm the kernels implement
def _pade_uv(A, m, ident):
    """U (odd part) and V (even part) of the [m/m] Pade numerator, value only."""
    b = PADE_B[m]
    A2 = A @ A
    if m == 13:
        A4 = A2 @ A2
        A6 = A4 @ A2
        W1 = b[13] * A6 + b[11] * A4 + b[9] * A2
        W2 = b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * ident
        Z1 = b[12] * A6 + b[10] * A4 + b[8] * A2
        Z2 = b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * ident
        U = A @ (A6 @ W1 + W2)
        V = A6 @ Z1 + Z2
        return U, V
    pows = [ident, A2]
    for _ in range(2, m // 2 + 1):
        pows.append(pows[-1] @ A2)
    W = sum(b[2 * k + 1] * pows[k] for k in range(m // 2 + 1))
    V = sum(b[2 * k] * pows[k] for k in range(m // 2 + 1))
    return A @ W, V


def pade_expm(A, frechet_thresholds=False):
    """Higham-2005 scaling-and-squaring, computed in the precision of ``A`` (complex64 ->
    orders 3/5/7; complex128 -> 3/5/7/9/13).  A: (..., n, n)."""
    A = np.asarray(A)
    single = A.dtype == np.complex64
    out = np.empty_like(A)
    flat_in = A.reshape(-1, *A.shape[-2:])
    flat_out = out.reshape(-1, *A.shape[-2:])
    n = A.shape[-1]
    ident = np.eye(n, dtype=A.dtype)
    for q, a in enumerate(flat_in):
        norm1 = float(np.abs(a).sum(axis=0).max())
        m, s = pade_plan(norm1, single, frechet_thresholds)
        a = a * a.dtype.type(2.0 ** -s)
        U, V = _pade_uv(a, m, ident)
        R = np.linalg.solve(V - U, V + U)
        for _ in range(s):
            R = R @ R
        flat_out[q] = R
    return out


def pade_expm_frechet(A, E):
    """exp(A) and L(A, E) by differentiating the Pade recurrences (Al-Mohy & Higham 2009,
    Alg. 6.4 structure), block-triangular: every product [[P,Q],[0,P]]*[[R,S],[0,R]] costs
    three n x n products, the solve reuses one LU.  Precision follows ``A``."""
    A, E = np.asarray(A), np.asarray(E).astype(np.asarray(A).dtype)
    single = A.dtype == np.complex64
    n = A.shape[-1]
    ident = np.eye(n, dtype=A.dtype)
    Ys, Ls = np.empty_like(A), np.empty_like(A)
    fa, fe = A.reshape(-1, n, n), E.reshape(-1, n, n)
    fy, fl = Ys.reshape(-1, n, n), Ls.reshape(-1, n, n)
    for q, (a, e) in enumerate(zip(fa, fe)):
        norm1 = float(np.abs(a).sum(axis=0).max())
        m, s = pade_plan(norm1, single, True)
        sc = a.dtype.type(2.0 ** -s)
        a, e = a * sc, e * sc
        b = PADE_B[m]
        A2 = a @ a
        M2 = a @ e + e @ a
        if m == 13:
            A4 = A2 @ A2
            M4 = A2 @ M2 + M2 @ A2
            A6 = A4 @ A2
            M6 = A4 @ M2 + M4 @ A2
            W1 = b[13] * A6 + b[11] * A4 + b[9] * A2
            W2 = b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * ident
            Z1 = b[12] * A6 + b[10] * A4 + b[8] * A2
            Z2 = b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * ident
            W = A6 @ W1 + W2
            U = a @ W
            V = A6 @ Z1 + Z2
            Lw1 = b[13] * M6 + b[11] * M4 + b[9] * M2
            Lw2 = b[7] * M6 + b[5] * M4 + b[3] * M2
            Lz1 = b[12] * M6 + b[10] * M4 + b[8] * M2
            Lz2 = b[6] * M6 + b[4] * M4 + b[2] * M2
            Lw = A6 @ Lw1 + M6 @ W1 + Lw2
            Lu = a @ Lw + e @ W
            Lv = A6 @ Lz1 + M6 @ Z1 + Lz2
        else:
            pows, dpows = [ident, A2], [np.zeros_like(a), M2]
            for _ in range(2, m // 2 + 1):
                dpows.append(pows[-1] @ M2 + dpows[-1] @ A2)
                pows.append(pows[-1] @ A2)
            W = sum(b[2 * k + 1] * pows[k] for k in range(m // 2 + 1))
            V = sum(b[2 * k] * pows[k] for k in range(m // 2 + 1))
            Lw = sum(b[2 * k + 1] * dpows[k] for k in range(m // 2 + 1))
            Lv = sum(b[2 * k] * dpows[k] for k in range(m // 2 + 1))
            U = a @ W
            Lu = a @ Lw + e @ W
        lu_piv = scipy.linalg.lu_factor(V - U)
        R = scipy.linalg.lu_solve(lu_piv, V + U)
        L = scipy.linalg.lu_solve(lu_piv, Lu + Lv + (Lu - Lv) @ R)
        for _ in range(s):
            L = R @ L + L @ R
            R = R @ R
        fy[q], fl[q] = R, L
    return Ys, Ls


def pade_expm_vjp(A, Ybar):
    """Y = exp(A), Abar = L(A^H, Ybar) = L(A, Ybar^H)^H: the fused kernel's contract."""
    A, Ybar = np.asarray(A), np.asarray(Ybar)
    Y, L = pade_expm_frechet(A, np.conj(np.swapaxes(Ybar, -1, -2)))
    return Y, np.conj(np.swapaxes(L, -1, -2))


# ------------------------------------------------------------------------------ P': the structured variants the kernels add
def skew_body_pullback(A, Z, s):
    """Skew-Hermitian part of Abar = L(A^H, Ybar) for skew-Hermitian A, from the body-frame cotangent
    Z = Ybar exp(A)^H, by the recursion the CUDA path runs (expm_large.cu backward_skew):
        S_0 = R_0^H L(A/2^s, Zs/2^s),  Zs = (Z - Z^H)/2,   S_{k+1} = S_k + R_k^H S_k R_k,  R_k = exp(A 2^(k-s)).
    Exact pieces (SciPy expm / expm_frechet) stand in for the Pade ones: this pins the identity, not the kernels."""
    A, Z = np.asarray(A, dtype=np.complex128), np.asarray(Z, dtype=np.complex128)
    Zs = 0.5 * (Z - Z.conj().T)
    h = 2.0 ** -s
    R, L0 = scipy.linalg.expm_frechet(A * h, Zs * h, compute_expm=True)
    S = R.conj().T @ L0
    for _ in range(s):
        S = S + R.conj().T @ S @ R
        R = R @ R
    return S


# Taylor coefficients of 1 / q_7(z), q_7(z) = sum_k (b_k / b_0) (-z)^k   (expm_large.cu series_inverse)
INV_Q7_SERIES = (1.0, 0.5, 7.0 / 52.0, 1.0 / 39.0, 343.0 / 89232.0, 107.0 / 223080.0, 5383.0 / 104401440.0,
                 893.0 / 182702520.0)


def series_inverse_q7(A):
    """(V - U)^-1 for the [7/7] Pade denominator of ``A`` (already scaled: ||A|| <= 0.783) as the kernels form it for
    one or two matrices: X0 = sum_{k<=7} d_k A^k, then one Newton-Schulz step X0 (2I - Q X0).  Returns (Qi, Q)."""
    A = np.asarray(A, dtype=np.complex128)
    ident = np.eye(A.shape[-1], dtype=A.dtype)
    U, V = _pade_uv(A, 7, ident)
    b0 = PADE_B[7][0]
    Q = (V - U) / b0                                         # normalised: q(0) = 1, as on the device
    A2 = A @ A; A4 = A2 @ A2; A6 = A4 @ A2
    d = INV_Q7_SERIES
    X0 = A @ (d[1] * ident + d[3] * A2 + d[5] * A4 + d[7] * A6) + (d[0] * ident + d[2] * A2 + d[4] * A4 + d[6] * A6)
    return X0 @ (2.0 * ident - Q @ X0), Q


def planner_start_vector(which, n):
    """the two fixed pseudo-random start vectors of the device power iteration (expm_large.cu start_vec)"""
    c = np.arange(n, dtype=np.uint64)
    off = np.uint64(0x9E3779B9 if which else 0)

    def hash_unit(v):
        v = v & np.uint64(0xFFFFFFFF)
        v ^= v >> np.uint64(16); v = (v * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
        v ^= v >> np.uint64(15); v = (v * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
        v ^= v >> np.uint64(16)
        return (v & np.uint64(0xFFFF)).astype(np.float64) / 32768.0 - 1.0

    return hash_unit(np.uint64(2) * c + np.uint64(1) + off) + 1j * hash_unit(np.uint64(2) * c + np.uint64(2) + off)


def spectral_radius_estimate(B, k, iters=8):
    """rho(A) estimate the planner uses for normal A from the top even power B = A^k (Hermitian up to sign): ``iters``
    power iterations from each of the two fixed start vectors; returns, per iteration, the larger ||B x|| ^ (1/k) of the
    two (all lower bounds on rho)."""
    B = np.asarray(B, dtype=np.complex128)
    n = B.shape[-1]
    xs = [planner_start_vector(0, n), planner_start_vector(1, n)]
    out = []
    for _ in range(iters):
        ys = [B @ (x / np.linalg.norm(x)) for x in xs]
        out.append(max(np.linalg.norm(y) for y in ys) ** (1.0 / k))
        xs = ys
    return out


# ------------------------------------------------------------------------------ synthetic inputs
def rand_hermitian(rng, batch, n, scaled=True):
    """H = (X + X^H)/(2 sqrt n), X iid complex standard normal; ``scaled=False`` drops the
    1/sqrt(n) ("stiff" set, ||A||_1 ~ n)."""
    x = rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))
    h = (x + np.conj(np.swapaxes(x, -1, -2))) / 2.0
    return h / np.sqrt(n) if scaled else h


def rand_cotangent(rng, batch, n):
    return rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))


def rel_fro(x, ref):
    """max over the batch of ||x - ref||_F / ||ref||_F."""
    x, ref = np.asarray(x), np.asarray(ref)
    x = x.reshape(-1, *x.shape[-2:]) if x.ndim > 2 else x[None]
    ref = ref.reshape(x.shape)
    num = np.sqrt((np.abs(x - ref) ** 2).sum(axis=(-1, -2)))
    den = np.sqrt((np.abs(ref) ** 2).sum(axis=(-1, -2)))
    return float((num / np.maximum(den, 1e-300)).max())
