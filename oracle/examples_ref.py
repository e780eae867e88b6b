"""Torch-CPU restatement of the example-local builders, cost functions and optimiser that the
BASELINE configs name.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Citations are
relative to ``/root/reference``.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import qgrad_ref as q

C64, C128 = torch.complex64, torch.complex128


def _rt(x, dtype):
    return q._t(x, q._real_dtype(dtype))


# ----------------------------------------------------------------------------- Qubit_Rotation.py
def rot(phi, theta, omega, dtype=C128):
    """examples/Qubit_Rotation.py:40-69 -- RZ(omega) RY(theta) RZ(phi), 2x2."""
    phi, theta, omega = _rt(phi, dtype), _rt(theta, dtype), _rt(omega, dtype)
    cos, sin = torch.cos(theta / 2), torch.sin(theta / 2)
    e = lambda x: torch.exp(1j * x.to(dtype))
    row0 = torch.stack([e(-0.5 * (phi + omega)) * cos, -e(0.5 * (phi - omega)) * sin])
    row1 = torch.stack([e(-0.5 * (phi - omega)) * sin, e(0.5 * (phi + omega)) * cos])
    return torch.stack([row0, row1])


def qubit_rotation_cost(phi, theta, omega, ket, dtype=C128):
    """examples/Qubit_Rotation.py:72-92 -- fidelity(rot @ ket, |0>)[0][0]."""
    evolved = rot(phi, theta, omega, dtype) @ q._t(ket, dtype)
    return q.fidelity(evolved, q.basis(2, 0, dtype))[0][0]


def rot_expect_sigmaz(phi, theta, omega, ket, dtype=C128):
    """BASELINE config #2 wording: <sigma_z> of rot @ ket (F = (1 + <sigma_z>)/2)."""
    evolved = rot(phi, theta, omega, dtype) @ q._t(ket, dtype)
    return torch.real(q.expect(q.sigmaz(dtype), evolved))


# ----------------------------------------------------------------------------- SNAP_gates.py
def pad_thetas(hilbert_size, thetas):
    """examples/SNAP_gates.py:84-100 -- right-pad with zeros to the cut-off."""
    thetas = list(np.asarray(thetas, dtype=float).reshape(-1))
    return np.asarray(thetas + [0.0] * (hilbert_size - len(thetas)))


def snap(hilbert_size, thetas, dtype=C128):
    """examples/SNAP_gates.py:103-119 -- sum_i e^{i theta_i} |i><i| (a diagonal matrix)."""
    thetas = _rt(thetas, dtype)
    op = torch.zeros((hilbert_size, hilbert_size), dtype=dtype)
    for i in range(thetas.shape[0]):
        op = op + torch.exp(1j * thetas[i].to(dtype)) * q.to_dm(q.basis(hilbert_size, i, dtype))
    return op


def apply_blocks(alphas, thetas, initial_state, dtype=C128):
    """examples/SNAP_gates.py:150-182 -- T blocks D(alpha) S(theta) D(-alpha) on a ket."""
    if len(alphas) != len(thetas):
        raise ValueError("The number of alphas and theta vectors should be same")
    x = q._t(initial_state, dtype)
    N = x.shape[0]
    displace = q.Displace(N, dtype)
    for t in range(len(alphas)):
        x = displace(alphas[t]) @ x
        x = snap(N, thetas[t], dtype) @ x
        x = displace(-alphas[t]) @ x
    return x


def snap_cost(params, initial, target, dtype=C128):
    """examples/SNAP_gates.py:231-248 -- 1 - fidelity(target, evolved)[0][0]."""
    alphas, thetas = params[0], params[1]
    evo = apply_blocks(alphas, thetas, initial, dtype)
    return 1 - q.fidelity(q._t(target, dtype), evo)[0][0]


# ----------------------------------------------------------------------------- Unitary-Learning-qgrad.py
def unitary_learning_cost(params, inputs, outputs, N, dtype=C64):
    """examples/Unitary-Learning-qgrad.py:123-151 -- 1 - mean_k |Re <out_k| U |in_k>|."""
    thetas, phis, omegas = params
    unitary = q.Unitary(N, dtype)(thetas, phis, omegas)
    loss = 0.0
    for k in range(len(inputs)):
        pred = unitary @ q._t(inputs[k], dtype)
        loss = loss + torch.abs(torch.real(torch.conj(q._t(outputs[k], dtype)).transpose(0, 1) @ pred))
    loss = 1 - (1 / len(inputs)) * loss
    return loss[0][0]


def unitary_learning_score(params, inputs, outputs, N, dtype=C64):
    """examples/Unitary-Learning-qgrad.py:178-206 -- mean fidelity(pred, out)."""
    thetas, phis, omegas = params
    unitary = q.Unitary(N, dtype)(thetas, phis, omegas)
    fidel = 0
    for k in range(len(inputs)):
        fidel = fidel + q.fidelity(unitary @ q._t(inputs[k], dtype), q._t(outputs[k], dtype))
    return (fidel / len(inputs))[0][0]


# ----------------------------------------------------------------------------- Unitary-Learning-no-qgrad.py
def make_unitary(N, params, A, B):
    """examples/Unitary-Learning-no-qgrad.py:112-141 -- prod_k e^{-i B tau_k} e^{-i A t_k},
    params = (t_1..t_N, tau_1..tau_N); torch.linalg.matrix_exp stands in for SciPy expm so
    that the whole chain is differentiable (the reference uses finite differences here)."""
    A, B = q._t(A, C128), q._t(B, C128)
    params = q._t(params, torch.float64).reshape(-1)
    d = A.shape[0]
    unitary = torch.eye(d, dtype=C128)
    for i in range(N):
        step = torch.linalg.matrix_exp(-1j * B * params[i + N]) @ torch.linalg.matrix_exp(-1j * A * params[i])
        unitary = step @ unitary
    return unitary


def hamiltonian_learning_cost(params, inputs, outputs, N, A, B):
    """examples/Unitary-Learning-no-qgrad.py:165-191 -- 1 - mean_k |Re <out_k|U(params)|in_k>|.
    (The reference rebuilds U per ket; the value is the same.)"""
    U = make_unitary(N, params, A, B)
    loss = 0.0
    for k in range(len(inputs)):
        pred = U @ q._t(inputs[k], C128)
        loss = loss + torch.abs(torch.real(torch.conj(q._t(outputs[k], C128)).transpose(0, 1) @ pred))
    return (1 - loss / len(inputs))[0][0]


def fd_gradient(cost, params, eps=1e-3):
    """examples/Unitary-Learning-no-qgrad.py:213-241 -- forward differences, eps = 1e-3."""
    params = np.asarray(params, dtype=float).reshape(-1)
    base = float(cost(params))
    out = np.zeros_like(params)
    for i in range(params.shape[0]):
        p = params.copy()
        p[i] += eps
        out[i] = (float(cost(p)) - base) / eps
    return out


# ----------------------------------------------------------------------------- Circuit-Learning.py
def rx(phi, dtype=C128):
    """examples/Circuit-Learning.py:47-63."""
    phi = _rt(phi, dtype)
    c, s = torch.cos(phi / 2).to(dtype), torch.sin(phi / 2).to(dtype)
    return torch.stack([torch.stack([c, -1j * s]), torch.stack([-1j * s, c])])


def ry(phi, dtype=C128):
    """examples/Circuit-Learning.py:66-80."""
    phi = _rt(phi, dtype)
    c, s = torch.cos(phi / 2).to(dtype), torch.sin(phi / 2).to(dtype)
    return torch.stack([torch.stack([c, -s]), torch.stack([s, c])])


def rz(phi, dtype=C128):
    """examples/Circuit-Learning.py:83-96."""
    phi = _rt(phi, dtype).to(dtype)
    z = torch.zeros((), dtype=dtype)
    return torch.stack([torch.stack([torch.exp(-0.5j * phi), z]), torch.stack([z, torch.exp(0.5j * phi)])])


def cnot(dtype=C128):
    """examples/Circuit-Learning.py:99-101."""
    return torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=dtype)


def circuit(params, dtype=C128):
    """examples/Circuit-Learning.py:104-123."""
    thetax, thetay, thetaz = params
    layer0 = torch.kron(q.basis(2, 0, dtype), q.basis(2, 0, dtype))
    layer1 = torch.kron(ry(math.pi / 4, dtype), ry(math.pi / 4, dtype))
    layer2 = torch.kron(rx(thetax, dtype), torch.eye(2, dtype=dtype))
    layer3 = torch.kron(ry(thetay, dtype), rz(thetaz, dtype))
    unitary = layer1 @ cnot(dtype) @ layer2 @ cnot(dtype) @ layer3
    return unitary @ layer0


def circuit_cost(params, dtype=C128):
    """examples/Circuit-Learning.py:142-158 -- Re expect(sigma_z (x) I, circuit(params))."""
    op = torch.kron(q.sigmaz(dtype), torch.eye(2, dtype=dtype))
    return torch.real(q.expect(op, circuit(params, dtype)))


# ----------------------------------------------------------------------------- optimiser
class Adam:
    """``jax.experimental.optimizers.adam`` defaults (used at Unitary-Learning-qgrad.py:218,
    SNAP_gates.py:268, Circuit-Learning.py:163): b1=0.9, b2=0.999, eps=1e-8, bias-corrected,
    x <- x - lr * mhat / (sqrt(vhat) + eps).  Complex parameters: v accumulates g*g (JAX's
    adam squares without conjugation); only real parameters are pinned by the KAT."""

    def __init__(self, params, step_size, b1=0.9, b2=0.999, eps=1e-8):
        self.x = [np.array(p, dtype=np.result_type(np.asarray(p).dtype, np.float32)) for p in params]
        self.m = [np.zeros_like(x) for x in self.x]
        self.v = [np.zeros_like(x) for x in self.x]
        self.lr, self.b1, self.b2, self.eps = step_size, b1, b2, eps

    def update(self, i, grads):
        for k, g in enumerate(grads):
            g = np.asarray(g)
            self.m[k] = (1 - self.b1) * g + self.b1 * self.m[k]
            self.v[k] = (1 - self.b2) * (g * g) + self.b2 * self.v[k]
            mhat = self.m[k] / (1 - self.b1 ** (i + 1))
            vhat = self.v[k] / (1 - self.b2 ** (i + 1))
            self.x[k] = self.x[k] - self.lr * mhat / (np.sqrt(vhat) + self.eps)
        return self.x


def circuit_learning_trace(epochs=400, every=50, dtype=C128):
    """Reproduces the loop at examples/Circuit-Learning.py:162-183: Adam(1e-2) from
    [0, 0, 0]; returns the losses printed at epochs 50, 100, ..."""
    opt = Adam([0.0, 0.0, 0.0], 1e-2)
    g = q.jax_style_grad(lambda a, b, c: circuit_cost((a, b, c), dtype), argnums=(0, 1, 2))
    printed = []
    for epoch in range(epochs):
        grads = g(*[torch.tensor(float(x), dtype=torch.float64) for x in opt.x])
        opt.update(epoch, [float(v) for v in grads])
        if epoch % every == every - 1:
            printed.append(float(circuit_cost(opt.x, dtype)))
    return printed


# ------------------------------------------------------------------------------ Hamiltonian learning with Pauli terms
def pauli_dense(x, z, nq):
    """<r|P|c> = i^popcount(x&z) (-1)^popcount(c&z) [r == c^x]: the (xmask, zmask) convention of
    include/qgrad_b200.h (qubit 0 = least significant bit)."""
    n = 1 << nq
    M = np.zeros((n, n), dtype=np.complex128)
    for c in range(n):
        M[c ^ x, c] = (1j) ** bin(x & z).count("1") * (-1.0) ** bin(c & z).count("1")
    return M


def pauli_learning_loss_and_grad(theta, ctrl, xmask, zmask, nq, t, psi_in, psi_out):
    """L(theta) = 1 - mean_{s,k} |Re <out_sk| exp(-i t H_s(theta)) |in_sk>|, H_s = sum_p ctrl[s,p] theta_p P_p
    (the cost of examples/Unitary-Learning-qgrad.py:123-151 on a Pauli-parametrised Hamiltonian), and its
    gradient by torch-CPU autograd.  psi_in / psi_out: (n_set, n, m) with kets as columns."""
    dense = torch.tensor(np.stack([pauli_dense(int(x), int(z), nq) for x, z in zip(xmask, zmask)]))
    th = torch.tensor(np.asarray(theta, dtype=np.float64), requires_grad=True)
    n_set, _, m = np.asarray(psi_in).shape
    acc = 0.0
    for s in range(n_set):
        H = torch.tensordot((torch.tensor(np.asarray(ctrl)[s]) * th).to(torch.complex128), dense, 1)
        U = torch.linalg.matrix_exp(-1j * t * H)
        o = torch.sum(torch.conj(torch.tensor(np.asarray(psi_out)[s])) * (U @ torch.tensor(np.asarray(psi_in)[s])), dim=0)
        acc = acc + torch.sum(torch.abs(torch.real(o)))
    loss = 1 - acc / (n_set * m)
    loss.backward()
    return float(loss), th.grad.numpy()


def pauli_entries(x, z, nq):
    """Sparse form of :func:`pauli_dense`: (rows, cols, values) of the n non-zero entries of the Pauli string."""
    n = 1 << nq
    cols = np.arange(n)
    par = np.array([bin(int(c) & int(z)).count("1") & 1 for c in cols])
    vals = (1j) ** bin(int(x) & int(z)).count("1") * (1.0 - 2.0 * par)
    return cols ^ int(x), cols, vals


def pauli_learning_loss_and_grad_frechet(theta, ctrl, xmask, zmask, nq, t, psi_in, psi_out, m_total=None):
    """The same loss and gradient as :func:`pauli_learning_loss_and_grad` from oracle O1 (SciPy ``expm`` +
    ``expm_frechet``) instead of torch autograd: sparse Hamiltonian assembly, so it runs at the BASELINE size
    (10 qubits, dim 1024: ~3 s per control setting) where the dense 38 x 1024 x 1024 term stack does not pay.
    ``m_total``: number of kets the mean runs over (default: all kets given -- pass the global count to score one
    rank's shard).  Chain: o_k = <out_k|U|in_k>, L = 1 - sum_k |Re o_k| / M, Ubar = G Psi_in^H with
    G[:, k] = -sign(Re o_k) out_k / M, Abar = L_exp(A^H, Ubar), thetabar_p = Re <dA/dtheta_p, Abar>."""
    import scipy.linalg
    theta, ctrl = np.asarray(theta, dtype=np.float64), np.asarray(ctrl, dtype=np.float64)
    psi_in, psi_out = np.asarray(psi_in, dtype=np.complex128), np.asarray(psi_out, dtype=np.complex128)
    n_set, n, m = psi_in.shape
    M = float(m_total if m_total is not None else n_set * m)
    ents = [pauli_entries(int(x), int(z), nq) for x, z in zip(xmask, zmask)]
    acc, grad = 0.0, np.zeros_like(theta)
    for s in range(n_set):
        H = np.zeros((n, n), dtype=np.complex128)
        for p, (r, c, v) in enumerate(ents):
            H[r, c] += ctrl[s, p] * theta[p] * v
        A = -1j * t * H
        U = scipy.linalg.expm(A)
        o = np.sum(np.conj(psi_out[s]) * (U @ psi_in[s]), axis=0)
        acc += float(np.sum(np.abs(o.real)))
        G = (-np.sign(o.real) / M)[None, :] * psi_out[s]
        Ubar = G @ psi_in[s].conj().T
        Abar = scipy.linalg.expm_frechet(A.conj().T, Ubar, compute_expm=False)
        for p, (r, c, v) in enumerate(ents):
            grad[p] += ctrl[s, p] * float(np.real(np.sum(np.conj(Abar[r, c]) * (-1j * t * v))))
    return 1.0 - acc / M, grad
