"""Drop-in host side of ``qgrad.qgrad_qutip`` (reference: qgrad/qgrad_qutip.py, public API list
docs/source/api.rst:11-40) over the CUDA kernels of libqgrad_b200.so.

Same names, positional signatures, return shapes and exception behaviour as the reference.
Arrays are ``torch`` CUDA tensors (the reference uses JAX arrays; JAX is not installable in this
environment); NumPy / Python inputs are accepted everywhere, as the
reference accepts them through ``jnp.asarray``.  Every differentiable op is a
``torch.autograd.Function`` whose forward and backward are the ``*_fwd`` / ``*_bwd`` entry
points of the C ABI -- the analogue of the ``jax.custom_vjp`` pair the XLA-FFI shim registers.
:func:`grad` reproduces ``jax.grad`` including its complex-cotangent convention.

There is no CPU path: constructors of constants (``basis``, ``sigmax`` ...) work anywhere, but
every compute op raises :class:`QgradB200Error` without a CUDA device.

North-star extensions (additive): ``Unitary(H)`` with an array ``H`` is the callable
``t -> expm(-i H t)``; ``expm``, ``rot``, ``make_rot``, ``snap``, ``displace_apply``; leading
batch dimensions on every op; ``dtype=`` on the builders (the reference hard-codes complex64).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import QG_ADJ_CONJ, QG_C64, QG_C128, QG_EXPM_FWD, QG_EXPM_FWD_VJP, QgradB200Error

__all__ = [
    "unitary_learning_cost",
    "basis", "coherent", "create", "dag", "destroy", "expect", "fidelity", "isbra", "isdm", "isherm",
    "isket", "rand_unitary", "rand_ket", "rand_dm", "sigmax", "sigmay", "sigmaz", "squeeze", "to_dm",
    "Displace", "Unitary", "_make_rot", "make_rot", "rot", "expm", "snap", "displace_apply", "grad",
    "rot_fidelity", "set_default_dtype", "get_default_dtype", "make_unitary", "layered_unitary_cost", "apply_blocks", "snap_cost", "rand_kets", "rand_herm", "rand_uniform", "eigh", "sqrtm", "fidelity_uhlmann",
]

C64, C128 = torch.complex64, torch.complex128
_DEFAULT = {"dtype": C64}     # the reference's builders are hard complex64 (qgrad_qutip.py:75,...)


def set_default_dtype(dtype):
    """complex64 (reference behaviour) or complex128 (the ``jax_enable_x64`` analogue)."""
    if dtype not in (C64, C128):
        raise ValueError("dtype must be torch.complex64 or torch.complex128")
    _DEFAULT["dtype"] = dtype


def get_default_dtype():
    return _DEFAULT["dtype"]


# ----------------------------------------------------------------------------- plumbing
def _device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def _require_cuda(what):
    if not torch.cuda.is_available():
        raise QgradB200Error(f"{what}: no CUDA device -- qgrad_b200 has no CPU path")


def _real_of(dtype):
    return torch.float32 if dtype == C64 else torch.float64


def _code(dtype):
    return QG_C64 if dtype == C64 else QG_C128


def _asarray(x, dtype=None, complex_=False):
    """``jnp.asarray`` analogue: NumPy / Python / torch in, tensor on the compute device out."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    if x.device != _device():
        x = x.to(_device())
    if dtype is not None:
        if x.dtype != dtype:
            x = x.to(dtype)
    elif complex_ and not x.is_complex():
        x = x.to(C128 if x.dtype == torch.float64 else C64)
    return x


def _cdtype(*xs):
    """complex compute dtype of a set of tensors (complex128 if any input is 64-bit)."""
    for x in xs:
        if x.dtype in (C128, torch.float64):
            return C128
    return C64


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _call(name, *args):
    _lib.check(getattr(_lib.load(), name)(*args), name)


# ----------------------------------------------------------------------------- predicates
def isket(state):
    """qgrad_qutip.py:333-342."""
    return state.shape[1] == 1


def isbra(state):
    """qgrad_qutip.py:345-354."""
    return state.shape[0] == 1


def dag(state):
    """qgrad_qutip.py:319-330 -- conjugate transpose (of the last two dims)."""
    state = _asarray(state)
    return torch.conj(state.transpose(-1, -2)).resolve_conj()


def _herm_stats(x):
    """(batch, 4) reals per matrix: exact-mismatch count, Re tr, Im tr, max |a_rc - conj(a_cr)| (``qg_herm_stats``)"""
    _require_cuda("isherm / isdm")
    x = _asarray(x, complex_=True)
    xb = x.reshape(-1, x.shape[-2], x.shape[-1]).contiguous()
    stats = torch.empty((xb.shape[0], 4), dtype=_real_of(xb.dtype), device=xb.device)
    _call("qg_herm_stats", _code(xb.dtype), _ptr(xb), _ptr(stats), xb.shape[0], xb.shape[-1], _stream())
    return stats


def isherm(oper):
    """qgrad_qutip.py:357-367 -- exact equality with the adjoint, tested on the device (one pass, one 16-byte read-back)."""
    oper = _asarray(oper, complex_=True)
    if oper.shape[-1] != oper.shape[-2]:
        return False
    return bool(float(_herm_stats(oper)[:, 0].sum()) == 0.0)


def eigh(A, eigenvectors=True):
    """Eigenvalues (ascending) and, optionally, eigenvectors (columns) of Hermitian matrices (..., n, n), n <= 64, by
    the batched Jacobi kernel (``qg_eigh``) -- what ``jnp.linalg.eigh`` / ``eig`` do for ``Displace.__init__``
    (qgrad_qutip.py:237) and ``isdm`` (:392).  Not differentiable."""
    _require_cuda("eigh")
    A = _asarray(A, complex_=True).detach()
    Ab = A.reshape(-1, A.shape[-2], A.shape[-1]).contiguous()
    B, n = Ab.shape[0], Ab.shape[-1]
    w = torch.empty((B, n), dtype=_real_of(Ab.dtype), device=Ab.device)
    V = torch.empty_like(Ab) if eigenvectors else None
    _call("qg_eigh", _code(Ab.dtype), _ptr(Ab), _ptr(w), _ptr(V), B, n, _stream())
    w = w.reshape(A.shape[:-1])
    return (w, V.reshape(A.shape)) if eigenvectors else w


def sqrtm(A):
    """Principal square root of Hermitian positive semi-definite matrices (..., n, n), n <= 64 (``qg_sqrtm_psd``) -- the
    use the reference imports ``scipy.linalg.sqrtm`` for (qgrad_qutip.py:7) and never makes.  Not differentiable."""
    _require_cuda("sqrtm")
    A = _asarray(A, complex_=True).detach()
    Ab = A.reshape(-1, A.shape[-2], A.shape[-1]).contiguous()
    out = torch.empty_like(Ab)
    _call("qg_sqrtm_psd", _code(Ab.dtype), _ptr(Ab), _ptr(out), Ab.shape[0], Ab.shape[-1], _stream())
    return out.reshape(A.shape)


def fidelity_uhlmann(a, b, squared=True):
    r"""True mixed-state fidelity :math:`F(\rho,\sigma) = (\mathrm{tr}\sqrt{\sqrt\rho\,\sigma\sqrt\rho})^2` of density
    matrices (kets / bras are promoted with ``to_dm``); ``squared=False`` gives QuTiP's ``fidelity`` (the root).  For two
    pure states it equals the reference's ket fidelity :math:`|\langle a|b\rangle|^2`; the reference's own density-matrix
    branch (``fidelity`` -> ``_fidelity_dm``, qgrad_qutip.py:50-64) is a diagonal-only surrogate, reproduced as written by
    :func:`fidelity`.  (..., n, n) with n <= 64 -> (...,) reals.  Value only (not differentiable)."""
    _require_cuda("fidelity_uhlmann")
    a, b = _asarray(a, complex_=True).detach(), _asarray(b, complex_=True).detach()
    dt = _cdtype(a, b)
    a, b = a.to(dt), b.to(dt)
    a = to_dm(a) if (a.shape[-1] == 1 or a.shape[-2] == 1) else a
    b = to_dm(b) if (b.shape[-1] == 1 or b.shape[-2] == 1) else b
    lead = torch.broadcast_shapes(a.shape[:-2], b.shape[:-2])
    n = a.shape[-1]
    ab = a.expand(lead + (n, n)).reshape(-1, n, n).contiguous()
    bb = b.expand(lead + (n, n)).reshape(-1, n, n).contiguous()
    out = torch.empty(ab.shape[0], dtype=_real_of(dt), device=ab.device)
    _call("qg_fidelity_uhlmann", _code(dt), _ptr(ab), _ptr(bb), _ptr(out), ab.shape[0], n, _stream())
    out = out if squared else torch.sqrt(out)
    return out.reshape(lead)


def isdm(mat):
    """qgrad_qutip.py:370-396 -- not a ket / bra, exactly Hermitian, ``allclose(Re tr, 1, atol=1e-9)``, and no eigenvalue
    below ``-1e-6``; Hermiticity and trace from one device pass (``qg_herm_stats``), the spectrum from the Jacobi kernel
    (``qg_eigh``, values only).  n <= 64."""
    mat = _asarray(mat, complex_=True)
    if isket(mat) or isbra(mat) or mat.dim() != 2 or mat.shape[0] != mat.shape[1]:
        return False
    st = _herm_stats(mat)[0].tolist()
    if st[0] != 0.0 or not abs(st[1] - 1.0) <= 1e-9 + 1e-5:           # jnp.allclose(x, 1, atol=1e-9): rtol keeps its default 1e-5
        return False
    w = eigh(mat.to(C128), eigenvectors=False)
    return bool(float(w.min()) >= -1e-6)


# ----------------------------------------------------------------------------- constants
def sigmax(dtype=None):
    """qgrad_qutip.py:67-75."""
    return torch.tensor([[0.0, 1.0], [1.0, 0.0]], dtype=dtype or _DEFAULT["dtype"], device=_device())


def sigmay(dtype=None):
    """qgrad_qutip.py:78-89."""
    return torch.tensor([[0.0, -1.0j], [1.0j, 0.0]], dtype=dtype or _DEFAULT["dtype"], device=_device())


def sigmaz(dtype=None):
    """qgrad_qutip.py:92-101."""
    return torch.tensor([[1.0, 0.0], [0.0, -1.0]], dtype=dtype or _DEFAULT["dtype"], device=_device())


def _ladder(N, k, dtype):
    if isinstance(N, bool) or not isinstance(N, (int, np.integer)):
        raise ValueError("Hilbert space dimension must be an integer value")
    data = np.sqrt(np.arange(1, N, dtype=np.float32))       # float32 sqrt as in the reference
    mat = np.diag(data.astype(np.float64), k=k) if N > 1 else np.zeros((N, N))
    return torch.as_tensor(mat).to(dtype or _DEFAULT["dtype"]).to(_device())


def destroy(N, dtype=None):
    """qgrad_qutip.py:104-121."""
    return _ladder(N, 1, dtype)


def create(N, dtype=None):
    """qgrad_qutip.py:137-152."""
    return _ladder(N, -1, dtype)


def basis(N, n=0, dtype=None):
    """qgrad_qutip.py:283-299.  ``n >= N`` yields zeros (JAX drops out-of-bounds scatters; the
    reference's own test relies on it, test_qgrad_qutip.py:188)."""
    if isinstance(N, bool) or not isinstance(N, (int, np.integer)) or N < 0:
        raise ValueError("N must be integer N >= 0")
    out = torch.zeros((N, 1), dtype=dtype or _DEFAULT["dtype"], device=_device())
    if -N <= n < N:
        out[n, 0] = 1.0
    return out


def to_dm(state):
    """qgrad_qutip.py:399-421 -- ket test first, then bra, else TypeError."""
    state = _asarray(state, complex_=True)
    if isket(state) or isbra(state):
        # the outer product, made Hermitian BIT FOR BIT: a fused multiply-add in the product kernel rounds x_i conj(x_j)
        # and x_j conj(x_i) differently, and isherm / isdm (exact equality, :367) would then reject a pure state's own
        # density matrix; (m + m^H) / 2 is exact-symmetric by construction and changes nothing above round-off
        m = state @ dag(state) if isket(state) else dag(state) @ state
        return 0.5 * (m + dag(m))
    raise TypeError(
        "Input is neither a ket, nor a bra. First dimension of a bra should be 1. Eg: (1, 4)."
        " Second dimension of a ket should be 1. Eg: (4, 1)")


# ----------------------------------------------------------------------------- expm
def _max_squarings(A, frechet):
    """Host-side bound on the number of squarings (one device->host read of the largest
    1-norm); used only by the large-dimension path, whose launch count depends on it."""
    norms = A.detach().abs().sum(dim=-2).amax(dim=-1).reshape(-1)
    finite = norms[torch.isfinite(norms)]             # a NaN / Inf matrix must not shrink the bound of its neighbours
    if finite.numel() == 0:
        return 1
    norm = float(finite.max())
    single = A.dtype == C64
    bound = (1.3 if single else (0.783 if frechet else 0.9504178996162932))
    # 1e-3 slack: the kernel sums the column norms in the matrix precision
    return max(1, int(math.ceil(math.log2(max(norm * (1 + 1e-3) / bound, 1.0)))))


class _Expm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, max_squarings):
        _require_cuda("expm")
        # read BEFORE .contiguous(): forward runs under no-grad, so the copy made for a transposed / conjugated view
        # (expm(dag(H)), -1j * H.mT * t) has requires_grad = False and the saved-workspace path -- and the Frechet
        # bound for the squaring budget -- would be skipped
        needs_grad = bool(ctx.needs_input_grad[0])
        A = A.contiguous()
        batch, n = (int(np.prod(A.shape[:-2])) if A.dim() > 2 else 1), A.shape[-1]
        Y = torch.empty_like(A)
        code = _code(A.dtype)
        lib = _lib.load()
        ctx.large = n > 32
        ctx.ms = 0
        ws = None
        if ctx.large:
            ctx.ms = max_squarings if max_squarings else _max_squarings(A, needs_grad)
            mode = QG_EXPM_FWD_VJP if needs_grad else QG_EXPM_FWD
            nbytes = lib.qg_expm_workspace_bytes(code, n, batch, mode, ctx.ms)
            free, _ = torch.cuda.mem_get_info()
            nbytes = min(nbytes, max(int(free * 0.8), lib.qg_expm_workspace_bytes(code, n, 1, mode, ctx.ms)))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
            full = nbytes >= lib.qg_expm_workspace_bytes(code, n, batch, mode, ctx.ms)
            if needs_grad and full:
                _call("qg_expm_fwd_save", code, _ptr(A), _ptr(Y), batch, n, _ptr(ws), ws.numel(), ctx.ms, _stream())
                ctx.ws = ws
            else:
                _call("qg_expm", code, _ptr(A), _ptr(Y), batch, n, _ptr(ws), ws.numel(), ctx.ms, _stream())
                ctx.ws = None
        else:
            _call("qg_expm", code, _ptr(A), _ptr(Y), batch, n, None, 0, 0, _stream())
            ctx.ws = None
        ctx.save_for_backward(A)
        ctx.batch, ctx.n = batch, n
        return Y

    @staticmethod
    def backward(ctx, Ybar):
        (A,) = ctx.saved_tensors
        Ybar = Ybar.contiguous().to(A.dtype)
        Abar = torch.empty_like(A)
        code = _code(A.dtype)
        if ctx.large and ctx.ws is not None:
            _call("qg_expm_bwd_saved", code, _ptr(A), _ptr(Ybar), _ptr(Abar), ctx.batch, ctx.n, QG_ADJ_CONJ,
                  _ptr(ctx.ws), ctx.ws.numel(), ctx.ms, _stream())
        else:
            ws, nb = None, 0
            if ctx.large:
                lib = _lib.load()
                nb = lib.qg_expm_workspace_bytes(code, ctx.n, ctx.batch, QG_EXPM_FWD_VJP, ctx.ms)
                free, _ = torch.cuda.mem_get_info()
                nb = min(nb, max(int(free * 0.8), lib.qg_expm_workspace_bytes(code, ctx.n, 1, QG_EXPM_FWD_VJP, ctx.ms)))
                ws = torch.empty(nb, dtype=torch.uint8, device=A.device)
            _call("qg_expm_fwd_vjp", code, _ptr(A), _ptr(Ybar), None, _ptr(Abar), ctx.batch, ctx.n, QG_ADJ_CONJ,
                  _ptr(ws), nb, ctx.ms, _stream())
        return Abar, None


def expm(A, max_squarings=None):
    """Batched complex matrix exponential of ``A`` (..., n, n), differentiable (reverse mode =
    Frechet derivative L(A^H, Ybar)).  Replaces ``scipy.linalg.expm`` at qgrad_qutip.py:280."""
    A = _asarray(A, complex_=True)
    if A.dim() < 2 or A.shape[-1] != A.shape[-2]:
        raise ValueError("expm expects square matrices (..., n, n)")
    return _Expm.apply(A, max_squarings)


def expm_plan(ws, batch):
    """Per-matrix number of squarings the large-dimension path chose in its last forward call over
    workspace ``ws`` (include/qgrad_b200.h: the first ``batch`` int32 of the 256-byte aligned workspace)."""
    off = (-ws.data_ptr()) % 256
    return ws[off: off + 4 * batch].view(torch.int32).clone()


def expm_fwd_vjp(A, Ybar, adjoint=QG_ADJ_CONJ, max_squarings=None, return_plan=False):
    """Fused value + pull-back in one pass (the kernel the expm+VJP throughput metric times).
    ``return_plan`` (n > 32 only, whole batch in one workspace chunk) also returns the squarings used."""
    _require_cuda("expm_fwd_vjp")
    A = _asarray(A, complex_=True).contiguous()
    Ybar = _asarray(Ybar, A.dtype).contiguous()
    batch, n = (int(np.prod(A.shape[:-2])) if A.dim() > 2 else 1), A.shape[-1]
    Y, Abar = torch.empty_like(A), torch.empty_like(A)
    code = _code(A.dtype)
    ws, nb, ms = None, 0, 0
    if n > 32:
        lib = _lib.load()
        ms = max_squarings or _max_squarings(A, True)
        nb = lib.qg_expm_workspace_bytes(code, n, batch, QG_EXPM_FWD_VJP, ms)
        free, _ = torch.cuda.mem_get_info()
        nb = min(nb, max(int(free * 0.8), lib.qg_expm_workspace_bytes(code, n, 1, QG_EXPM_FWD_VJP, ms)))
        ws = torch.empty(nb, dtype=torch.uint8, device=A.device)
    _call("qg_expm_fwd_vjp", code, _ptr(A), _ptr(Ybar), _ptr(Y), _ptr(Abar), batch, n, adjoint, _ptr(ws), nb, ms,
          _stream())
    if return_plan:
        return Y, Abar, (expm_plan(ws, batch) if ws is not None else None)
    return Y, Abar


def squeeze(N, z, dtype=None):
    """qgrad_qutip.py:264-280 -- exp(0.5 (conj(z) a^2 - z a^dag^2)).  The reference calls SciPy
    here (not differentiable, ``# TODO:gradients of squeezing``); this one runs on the expm kernel
    and is differentiable in ``z``."""
    dt = dtype or _DEFAULT["dtype"]
    z = _asarray(z, dt)
    a, ad = destroy(N, dt), create(N, dt)
    op = 0.5 * (torch.conj(z) * (a @ a) - z * (ad @ ad))
    return expm(op)


def make_unitary(N, params, A, B, dtype=C128):
    r"""examples/Unitary-Learning-no-qgrad.py:112-141 -- the layered ansatz
    :math:`U(\vec t,\vec\tau)=e^{-iB\tau_N}e^{-iAt_N}\cdots e^{-iB\tau_1}e^{-iAt_1}`, ``params = (t_1..t_N,
    tau_1..tau_N)`` (flat or (2N, 1) as in the reference).  The 2N exponentials are ONE batched launch of the
    expm kernel (the reference calls SciPy 2N times per cost evaluation and differentiates by finite
    differences, :213-241); differentiable in ``params`` through the Frechet pull-back."""
    A, B = _asarray(A, dtype), _asarray(B, dtype)
    params = _asarray(params, _real_of(dtype)).reshape(-1)
    if params.shape[0] != 2 * N:
        raise ValueError("params must hold N values of t followed by N values of tau")
    gens = torch.cat([(-1j * A).unsqueeze(0).expand(N, -1, -1), (-1j * B).unsqueeze(0).expand(N, -1, -1)])
    steps = expm(gens * params.to(dtype).reshape(2 * N, 1, 1))          # e^{-iA t_k} (k < N), e^{-iB tau_k}
    unitary = torch.eye(A.shape[-1], dtype=dtype, device=A.device)
    for k in range(N):
        unitary = steps[N + k] @ (steps[k] @ unitary)
    return unitary


def layered_unitary_cost(params, inputs, outputs, N, A, B, dtype=C128):
    """examples/Unitary-Learning-no-qgrad.py:165-191 -- ``1 - mean_k |Re <out_k| U(params) |in_k>|`` with
    :func:`make_unitary`; inputs / outputs: lists of (d, 1) kets or (M, d) arrays."""
    U = make_unitary(N, params, A, B, dtype)
    d = U.shape[-1]

    def kets(x):
        if isinstance(x, (list, tuple)):
            x = torch.stack([_asarray(v, dtype).reshape(-1) for v in x])
        return _asarray(x, dtype).reshape(-1, d)

    kin, kout = kets(inputs), kets(outputs)
    o = torch.sum(torch.conj(kout) * (kin @ U.transpose(0, 1)), dim=-1)
    return 1 - torch.mean(torch.abs(torch.real(o)))


# ----------------------------------------------------------------------------- expect / fidelity
def _batchify(x, core_dims):
    """(tensor with a leading batch dim, had_batch)"""
    if x.dim() == core_dims:
        return x.unsqueeze(0), False
    return x, True


def _oper_batched(oper, B):
    """1 if every sample has its own operator, 0 if one operator is shared; anything else is an error (a
    (B', n, n) operator batch against B states would silently read oper[0] -- or out of bounds)."""
    if oper.shape[0] == B and B > 1:
        return 1
    if oper.shape[0] == 1:
        return 0
    raise ValueError(f"expect: operator batch {oper.shape[0]} does not match state batch {B} (must be 1 or equal)")


class _ExpectKet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, oper, ket):          # oper (B|1, n, n), ket (B, n)
        _require_cuda("expect")
        oper, ket = oper.contiguous(), ket.contiguous()
        B, n = ket.shape
        ob = _oper_batched(oper, B)
        out = torch.empty(B, dtype=ket.dtype, device=ket.device)
        _call("qg_expect_ket_fwd", _code(ket.dtype), _ptr(oper), _ptr(ket), _ptr(out), B, n, ob, _stream())
        ctx.save_for_backward(oper, ket)
        ctx.ob = ob
        return out

    @staticmethod
    def backward(ctx, g):
        oper, ket = ctx.saved_tensors
        B, n = ket.shape
        g = g.contiguous().to(ket.dtype)
        kb = torch.empty_like(ket) if ctx.needs_input_grad[1] else None
        ob_ = torch.empty_like(oper) if ctx.needs_input_grad[0] else None
        _call("qg_expect_ket_bwd", _code(ket.dtype), _ptr(oper), _ptr(ket), _ptr(g), _ptr(kb), _ptr(ob_), B, n,
              ctx.ob, _stream())
        return ob_, kb


class _ExpectDm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, oper, rho):          # oper (B|1, n, n), rho (B, n, n)
        _require_cuda("expect")
        oper, rho = oper.contiguous(), rho.contiguous()
        B, n = rho.shape[0], rho.shape[-1]
        ob = _oper_batched(oper, B)
        out = torch.empty(B, dtype=rho.dtype, device=rho.device)
        _call("qg_expect_dm_fwd", _code(rho.dtype), _ptr(oper), _ptr(rho), _ptr(out), B, n, ob, _stream())
        ctx.save_for_backward(oper, rho)
        ctx.ob = ob
        return out

    @staticmethod
    def backward(ctx, g):
        oper, rho = ctx.saved_tensors
        B, n = rho.shape[0], rho.shape[-1]
        g = g.contiguous().to(rho.dtype)
        rb = torch.empty_like(rho) if ctx.needs_input_grad[1] else None
        ob_ = torch.empty_like(oper) if ctx.needs_input_grad[0] else None
        _call("qg_expect_dm_bwd", _code(rho.dtype), _ptr(oper), _ptr(rho), _ptr(g), _ptr(rb), _ptr(ob_), B, n,
              ctx.ob, _stream())
        return ob_, rb


def expect(oper, state):
    """qgrad_qutip.py:166-203 -- density-matrix branch iff ``state.shape[-1] >= 2``; 0-d complex
    (or (B,) for batched states (B, n, 1) / (B, n, n))."""
    oper, state = _asarray(oper, complex_=True), _asarray(state, complex_=True)
    dt = _cdtype(oper, state)
    oper, state = oper.to(dt), state.to(dt)
    state_b, batched = _batchify(state, 2)
    oper_b, oper_batched = _batchify(oper, 2)
    if oper_batched and state_b.shape[0] == 1 and oper_b.shape[0] > 1:      # one state against a batch of operators
        state_b, batched = state_b.expand(oper_b.shape[0], -1, -1), True
    if state_b.shape[-1] >= 2:
        out = _ExpectDm.apply(oper_b, state_b)
    else:
        out = _ExpectKet.apply(oper_b, state_b.squeeze(-1))
    return out if batched else out[0]


class _FidelityKet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):               # (B, n) each
        _require_cuda("fidelity")
        a, b = a.contiguous(), b.contiguous()
        B, n = a.shape
        out = torch.empty(B, dtype=_real_of(a.dtype), device=a.device)
        _call("qg_fidelity_ket_fwd", _code(a.dtype), _ptr(a), _ptr(b), _ptr(out), B, n, _stream())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        B, n = a.shape
        g = g.contiguous().to(_real_of(a.dtype))
        ab = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        bb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        _call("qg_fidelity_ket_bwd", _code(a.dtype), _ptr(a), _ptr(b), _ptr(g), _ptr(ab), _ptr(bb), B, n, _stream())
        return ab, bb


class _FidelityDm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r1, r2):             # (B, n, n) each
        _require_cuda("fidelity")
        r1, r2 = r1.contiguous(), r2.contiguous()
        B, n = r1.shape[0], r1.shape[-1]
        out = torch.empty(B, dtype=_real_of(r1.dtype), device=r1.device)
        _call("qg_fidelity_dm_fwd", _code(r1.dtype), _ptr(r1), _ptr(r2), _ptr(out), B, n, _stream())
        ctx.save_for_backward(r1, r2)
        return out

    @staticmethod
    def backward(ctx, g):
        r1, r2 = ctx.saved_tensors
        B, n = r1.shape[0], r1.shape[-1]
        g = g.contiguous().to(_real_of(r1.dtype))
        b1 = torch.empty_like(r1) if ctx.needs_input_grad[0] else None
        b2 = torch.empty_like(r2) if ctx.needs_input_grad[1] else None
        _call("qg_fidelity_dm_bwd", _code(r1.dtype), _ptr(r1), _ptr(r2), _ptr(g), _ptr(b1), _ptr(b2), B, n, _stream())
        return b1, b2


def fidelity(a, b):
    """qgrad_qutip.py:11-64.  ket,ket -> (1,1) real ((B,1,1) batched); otherwise kets are promoted
    with ``to_dm`` and the reference's elementwise-abs surrogate is returned as a 0-d real."""
    a, b = _asarray(a, complex_=True), _asarray(b, complex_=True)
    dt = _cdtype(a, b)
    a, b = a.to(dt), b.to(dt)
    if a.shape[-1] == 1 and b.shape[-1] == 1:
        ab, batched_a = _batchify(a, 2)
        bb, batched_b = _batchify(b, 2)
        B = max(ab.shape[0], bb.shape[0])
        ab, bb = ab.expand(B, -1, -1), bb.expand(B, -1, -1)
        out = _FidelityKet.apply(ab.squeeze(-1), bb.squeeze(-1)).reshape(B, 1, 1)
        return out if (batched_a or batched_b) else out[0]
    if a.shape[-1] == 1:
        a = a @ dag(a)
    if b.shape[-1] == 1:
        b = b @ dag(b)
    ab, batched_a = _batchify(a, 2)
    bb, batched_b = _batchify(b, 2)
    B = max(ab.shape[0], bb.shape[0])
    out = _FidelityDm.apply(ab.expand(B, -1, -1), bb.expand(B, -1, -1))
    return out if (batched_a or batched_b) else out[0]


# ----------------------------------------------------------------------------- Displace / coherent
_DISPLACE_CACHE = {}


def _displace_constants(n, dtype):
    """Eigensystem of the tridiagonal T of qgrad_qutip.py:225-239, computed once per (n, dtype,
    device) on the host (LAPACK) and cached on the device."""
    key = (n, dtype, str(_device()))
    if key not in _DISPLACE_CACHE:
        k = np.arange(1, n)
        sym = (2.0 * (k % 2) - 1) * np.sqrt(k)
        mat = np.diag(sym, 1) + np.diag(sym, -1) if n > 1 else np.zeros((n, n))
        rdt = _real_of(dtype)
        if n <= 64 and torch.cuda.is_available():
            # jnp.linalg.eigh of the reference (:237) on the device: the Jacobi kernel in complex128; T is real symmetric,
            # so every eigenvector is a real vector times a phase -- rotate each column to its real form
            w, V = eigh(torch.as_tensor(mat).to(C128).to(_device()))
            big = V.abs().argmax(dim=0)
            ph = V[big, torch.arange(n, device=V.device)]
            V = (V * (ph.conj() / ph.abs())).real
            evals_d, evecs_d = w.to(rdt).contiguous(), V.to(rdt).contiguous()
        else:                                        # larger cut-offs: constants from LAPACK, computed once per n
            evals, evecs = np.linalg.eigh(mat)
            evals_d = torch.as_tensor(evals).to(rdt).to(_device()).contiguous()
            evecs_d = torch.as_tensor(np.ascontiguousarray(evecs)).to(rdt).to(_device()).contiguous()
        _DISPLACE_CACHE[key] = (evecs_d, evals_d)
    return _DISPLACE_CACHE[key]


class _DisplaceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, evecs, evals):        # alpha (B,) complex
        _require_cuda("Displace")
        alpha = alpha.contiguous()
        B, n = alpha.shape[0], evals.shape[0]
        D = torch.empty((B, n, n), dtype=alpha.dtype, device=alpha.device)
        _call("qg_displace_fwd", _code(alpha.dtype), _ptr(evecs), _ptr(evals), _ptr(alpha), _ptr(D), B, n, _stream())
        ctx.save_for_backward(alpha, evecs, evals)
        return D

    @staticmethod
    def backward(ctx, Dbar):
        alpha, evecs, evals = ctx.saved_tensors
        B, n = alpha.shape[0], evals.shape[0]
        Dbar = Dbar.contiguous().to(alpha.dtype)
        abar = torch.empty_like(alpha)
        _call("qg_displace_bwd", _code(alpha.dtype), _ptr(evecs), _ptr(evals), _ptr(alpha), _ptr(Dbar), _ptr(abar), B,
              n, _stream())
        return abar, None, None


class _DisplaceApplyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, x, evecs, evals):     # alpha (B,), x (B, n)
        _require_cuda("displace_apply")
        alpha, x = alpha.contiguous(), x.contiguous()
        B, n = x.shape
        y = torch.empty_like(x)
        _call("qg_displace_apply_fwd", _code(x.dtype), _ptr(evecs), _ptr(evals), _ptr(alpha), _ptr(x), _ptr(y), B, n,
              _stream())
        ctx.save_for_backward(alpha, x, evecs, evals)
        return y

    @staticmethod
    def backward(ctx, ybar):
        alpha, x, evecs, evals = ctx.saved_tensors
        B, n = x.shape
        ybar = ybar.contiguous().to(x.dtype)
        abar = torch.empty_like(alpha)
        xbar = torch.empty_like(x)
        _call("qg_displace_apply_bwd", _code(x.dtype), _ptr(evecs), _ptr(evals), _ptr(alpha), _ptr(x), _ptr(ybar),
              _ptr(abar), _ptr(xbar), B, n, _stream())
        return abar, xbar, None, None


class Displace:
    r"""qgrad_qutip.py:216-261 -- :math:`D(\alpha)=\exp(\alpha a^\dagger-\alpha^* a)` through the
    eigendecomposition of a real symmetric tridiagonal.  ``Displace(n)(alpha)`` -> (n, n);
    a batch of alphas (B,) -> (B, n, n).  dtype: complex128 as the reference requests (:229)."""

    def __init__(self, n, dtype=C128):
        self.n = n
        self.dtype = dtype
        self.evecs, self.evals = _displace_constants(n, dtype)

    def __call__(self, alpha):
        alpha = _asarray(alpha, self.dtype)
        a1 = alpha.reshape(-1)
        D = _DisplaceFn.apply(a1, self.evecs, self.evals)
        return D[0] if alpha.dim() == 0 else D.reshape(*alpha.shape, self.n, self.n)

    def apply(self, alpha, x):
        """D(alpha) @ x without forming D; x (n,1) / (B,n,1) kets."""
        return displace_apply(self, alpha, x)


def displace_apply(displace, alpha, x):
    """Fused D(alpha)|x> (three O(n^2) products instead of an n^3 build + matvec): the D.x steps
    of ``apply_blocks`` (examples/SNAP_gates.py:150-182) and ``coherent``."""
    x = _asarray(x, displace.dtype)
    alpha = _asarray(alpha, displace.dtype)
    xb, batched = _batchify(x, 2)
    B = max(xb.shape[0], alpha.numel())
    y = _DisplaceApplyFn.apply(alpha.reshape(-1).expand(B), xb.squeeze(-1).expand(B, -1), displace.evecs,
                               displace.evals).unsqueeze(-1)
    return y if (batched or alpha.dim() > 0) else y[0]


def coherent(N, alpha, dtype=C128):
    """qgrad_qutip.py:302-316 -- D(alpha)|0>, (N, 1)."""
    return displace_apply(Displace(N, dtype), alpha, basis(N, 0, dtype))


def snap(hilbert_size, thetas, dtype=C128):
    """examples/SNAP_gates.py:103-119 -- diag(e^{i theta}), zero beyond len(thetas) (the
    reference sums e^{i theta_i}|i><i| over the thetas it is given)."""
    thetas = _asarray(thetas, _real_of(dtype))
    d = torch.zeros(thetas.shape[:-1] + (hilbert_size,), dtype=dtype, device=thetas.device)
    d[..., : thetas.shape[-1]] = torch.exp(1j * thetas.to(dtype))
    return torch.diag_embed(d)


def apply_blocks(alphas, thetas, initial_state, dtype=C128):
    """examples/SNAP_gates.py:150-182 -- T blocks D(alpha_t) SNAP(theta_t) D(-alpha_t) on a ket, for one parameter
    set (alphas (T,), thetas (T, k), state (n, 1) -> (n, 1)) or a batch of them (alphas (B, T), thetas (B, T, k),
    state (n, 1) or (B, n, 1) -> (B, n, 1)).  Neither D nor the SNAP matrix is formed: D.x is the fused three-product
    kernel (qg_displace_apply_*) and the SNAP gate, being diagonal, is an elementwise factor -- ``e^{i theta_i}`` for the
    k thetas given and ZERO beyond them, exactly as ``snap`` sums ``e^{i theta_i}|i><i|`` over the thetas it is handed
    (:103-119; the example pads its theta rows to the cut-off with ``pad_thetas`` first)."""
    alphas = _asarray(alphas, dtype)
    thetas = _asarray(thetas, _real_of(dtype))
    x = _asarray(initial_state, dtype)
    if alphas.shape[-1] != thetas.shape[-2]:
        raise ValueError("The number of alphas and theta vectors should be same")
    batched = alphas.dim() > 1
    n = x.shape[-2]
    ab = alphas.reshape(-1, alphas.shape[-1])
    tb = thetas.reshape(-1, thetas.shape[-2], thetas.shape[-1])
    phase = torch.exp(1j * tb.to(dtype))
    if tb.shape[-1] < n:
        phase = torch.nn.functional.pad(phase, (0, n - tb.shape[-1]))
    disp = Displace(n, dtype)
    xb = x.reshape(-1, n).expand(ab.shape[0], n)
    for t in range(ab.shape[1]):
        xb = _DisplaceApplyFn.apply(ab[:, t], xb, disp.evecs, disp.evals)
        xb = phase[:, t] * xb
        xb = _DisplaceApplyFn.apply(-ab[:, t], xb, disp.evecs, disp.evals)
    out = xb.unsqueeze(-1)
    return out if batched else out[0]


class _SnapCostFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alphas, thetas, init, target, evecs, evals):     # (B, T) complex, (B, T, n) real, (n,), (n,)
        _require_cuda("snap_cost")
        alphas, thetas, init, target = alphas.contiguous(), thetas.contiguous(), init.contiguous(), target.contiguous()
        B, T = alphas.shape
        n = evals.shape[0]
        cost = torch.empty(B, dtype=thetas.dtype, device=alphas.device)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        ab = torch.empty_like(alphas) if need else None
        tb = torch.empty_like(thetas) if need else None
        _call("qg_snap_cost_fwd_grad", _code(alphas.dtype), _ptr(evecs), _ptr(evals), _ptr(alphas), _ptr(thetas), _ptr(init),
              _ptr(target), T, _ptr(cost), _ptr(ab), _ptr(tb), B, n, _stream())
        ctx.save_for_backward(ab, tb)
        return cost

    @staticmethod
    def backward(ctx, g):
        ab, tb = ctx.saved_tensors
        return g.unsqueeze(-1) * ab, g.reshape(-1, 1, 1) * tb, None, None, None, None


def snap_cost(alphas, thetas, initial_state, target_state, dtype=C128):
    """examples/SNAP_gates.py:231-248 -- ``1 - fidelity(target, apply_blocks(alphas, thetas, initial))[0][0]``; a 0-d real, or
    (B,) for a batch of parameter sets sharing the two states.  With full-length theta rows (the example's ``pad_thetas``)
    value and gradient come out of ONE kernel (``qg_snap_cost_fwd_grad``: the state never leaves shared memory, the reverse
    sweep un-applies the unitary factors); shorter rows -- where ``snap`` is a projector -- take the composed path."""
    alphas = _asarray(alphas, dtype)
    thetas = _asarray(thetas, _real_of(dtype))
    init, target = _asarray(initial_state, dtype), _asarray(target_state, dtype)
    n = init.shape[-2]
    fused = thetas.shape[-1] == n and init.dim() == 2 and target.dim() == 2 and init.shape[-1] == 1 and target.shape[-1] == 1
    if not fused:
        evo = apply_blocks(alphas, thetas, init, dtype)
        return 1 - fidelity(target, evo)[..., 0, 0]
    if alphas.shape[-1] != thetas.shape[-2]:
        raise ValueError("The number of alphas and theta vectors should be same")
    batched = alphas.dim() > 1
    disp = Displace(n, dtype)
    out = _SnapCostFn.apply(alphas.reshape(-1, alphas.shape[-1]), thetas.reshape(-1, thetas.shape[-2], n), init.reshape(n),
                            target.reshape(n), disp.evecs, disp.evals)
    return out if batched else out[0]


# ----------------------------------------------------------------------------- rotations / Unitary
class _MakeRotFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, N, i, j, cdtype):    # params (B, 2) real
        _require_cuda("_make_rot")
        params = params.contiguous()
        B = params.shape[0]
        R = torch.empty((B, N, N), dtype=cdtype, device=params.device)
        _call("qg_make_rot_fwd", _code(cdtype), _ptr(params), _ptr(R), B, N, i, j, _stream())
        ctx.save_for_backward(params)
        ctx.meta = (N, i, j, cdtype)
        return R

    @staticmethod
    def backward(ctx, Rbar):
        (params,) = ctx.saved_tensors
        N, i, j, cdtype = ctx.meta
        Rbar = Rbar.contiguous().to(cdtype)
        pbar = torch.empty_like(params)
        _call("qg_make_rot_bwd", _code(cdtype), _ptr(params), _ptr(Rbar), _ptr(pbar), params.shape[0], N, i, j,
              _stream())
        return pbar, None, None, None, None


def _make_rot(N, params, idx, dtype=None):
    """qgrad_qutip.py:424-459 -- N x N identity with four entries replaced."""
    dt = dtype or _DEFAULT["dtype"]
    i, j = int(idx[0]), int(idx[1])
    if isinstance(params, (list, tuple)):
        params = torch.stack([_asarray(p, _real_of(dt)).reshape(()) for p in params])
    params = _asarray(params, _real_of(dt))
    pb, batched = _batchify(params, 1)
    R = _MakeRotFn.apply(pb, int(N), i % N, j % N, dt)
    return R if batched else R[0]


make_rot = _make_rot


class _GivensFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, thetas, phis, omegas, cdtype):     # (B, K), (B, K), (B, N) real
        _require_cuda("Unitary")
        thetas, phis, omegas = thetas.contiguous(), phis.contiguous(), omegas.contiguous()
        B, N = omegas.shape
        U = torch.empty((B, N, N), dtype=cdtype, device=omegas.device)
        _call("qg_givens_unitary_fwd", _code(cdtype), _ptr(thetas), _ptr(phis), _ptr(omegas), _ptr(U), B, N, _stream())
        ctx.save_for_backward(thetas, phis, omegas)
        ctx.cdtype = cdtype
        return U

    @staticmethod
    def backward(ctx, Ubar):
        thetas, phis, omegas = ctx.saved_tensors
        B, N = omegas.shape
        Ubar = Ubar.contiguous().to(ctx.cdtype)
        tb, pb, ob = torch.empty_like(thetas), torch.empty_like(phis), torch.empty_like(omegas)
        _call("qg_givens_unitary_bwd", _code(ctx.cdtype), _ptr(thetas), _ptr(phis), _ptr(omegas), _ptr(Ubar), _ptr(tb),
              _ptr(pb), _ptr(ob), B, N, _stream())
        return tb, pb, ob, None


class Unitary:
    r"""qgrad_qutip.py:462-555 -- ``Unitary(N)(thetas, phis, omegas)``: diag(e^{i omega}) times
    the product of N(N-1)/2 Givens rotations, (N, N) complex64.  Batched parameter sets
    (B, K), (B, K), (B, N) -> (B, N, N).

    North-star extension: ``Unitary(H)`` with an array ``H`` (..., n, n) is the callable
    ``t -> expm(-i H t)`` (dispatch on int vs array keeps ``Unitary(N)`` intact)."""

    def __init__(self, N, dtype=None):
        if isinstance(N, (int, np.integer)) and not isinstance(N, bool):
            self.N = int(N)
            self.H = None
            self.dtype = dtype or _DEFAULT["dtype"]
        else:
            self.H = _asarray(N, dtype, complex_=True)
            self.N = self.H.shape[-1]
            self.dtype = self.H.dtype

    def __call__(self, *args):
        if self.H is not None:
            (t,) = args
            t = _asarray(t, _real_of(self.dtype)).to(self.dtype)
            if t.dim() > 0:
                t = t.reshape(t.shape + (1, 1))
            return expm(-1j * self.H * t)
        thetas, phis, omegas = args
        rdt = _real_of(self.dtype)
        thetas, phis, omegas = _asarray(thetas, rdt), _asarray(phis, rdt), _asarray(omegas, rdt)
        N = self.N
        if omegas.shape[-1] != N:
            raise ValueError("The dimension of omegas should be the same as the unitary")
        if phis.shape[-1] != thetas.shape[-1]:
            raise ValueError("Number of phi and theta rotation parameters should be the same")
        if phis.shape[-1] != N * (N - 1) / 2 or thetas.shape[-1] != N * (N - 1) / 2:
            raise ValueError(
                "Size of each of the rotation parameters should be N * (N - 1) / 2, where N is the size of the"
                " unitary matrix")
        tb, batched = _batchify(thetas, 1)
        pb, _ = _batchify(phis, 1)
        ob, _ = _batchify(omegas, 1)
        U = _GivensFn.apply(tb, pb, ob, self.dtype)
        return U if batched else U[0]


class _RotZyzFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, angles, cdtype):             # (B, 3) real
        _require_cuda("rot")
        angles = angles.contiguous()
        B = angles.shape[0]
        R = torch.empty((B, 2, 2), dtype=cdtype, device=angles.device)
        _call("qg_rot_zyz_fwd", _code(cdtype), _ptr(angles), _ptr(R), B, _stream())
        ctx.save_for_backward(angles)
        ctx.cdtype = cdtype
        return R

    @staticmethod
    def backward(ctx, Rbar):
        (angles,) = ctx.saved_tensors
        Rbar = Rbar.contiguous().to(ctx.cdtype)
        ab = torch.empty_like(angles)
        _call("qg_rot_zyz_bwd", _code(ctx.cdtype), _ptr(angles), _ptr(Rbar), _ptr(ab), angles.shape[0], _stream())
        return ab, None


def rot(phi, theta, omega, dtype=None):
    """examples/Qubit_Rotation.py:40-69 -- RZ(omega) RY(theta) RZ(phi); (2,2) or (B,2,2)."""
    dt = dtype or _DEFAULT["dtype"]
    rdt = _real_of(dt)
    phi, theta, omega = _asarray(phi, rdt), _asarray(theta, rdt), _asarray(omega, rdt)
    batched = phi.dim() > 0 or theta.dim() > 0 or omega.dim() > 0
    ang = torch.stack(torch.broadcast_tensors(phi.reshape(-1), theta.reshape(-1), omega.reshape(-1)), dim=-1)
    R = _RotZyzFn.apply(ang, dt)
    return R if batched else R[0]


class _RotFidelityFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, angles, ket, cdtype):        # angles (B,3) real, ket (B|1, 2) complex
        _require_cuda("rot_fidelity")
        angles, ket = angles.contiguous(), ket.contiguous()
        B = angles.shape[0]
        rdt = _real_of(cdtype)
        fid = torch.empty(B, dtype=rdt, device=angles.device)
        grad_ = torch.empty((B, 3), dtype=rdt, device=angles.device)
        _call("qg_rot_fidelity_fwd_grad", _code(cdtype), _ptr(angles), _ptr(ket), 1 if ket.shape[0] == B and B > 1 else 0,
              _ptr(fid), _ptr(grad_), B, _stream())
        ctx.save_for_backward(grad_)
        return fid

    @staticmethod
    def backward(ctx, g):
        (grad_,) = ctx.saved_tensors
        return g.unsqueeze(-1) * grad_, None, None


def rot_fidelity(phi, theta, omega, ket, dtype=None):
    """Fused cost of examples/Qubit_Rotation.py:72-92: fidelity(rot(phi,theta,omega) @ ket, |0>)
    = (1 + <sigma_z>)/2 for batches of parameter sets, with the gradient produced in the same
    kernel pass.  Returns (B,) reals (0-d for scalar angles)."""
    dt = dtype or _DEFAULT["dtype"]
    rdt = _real_of(dt)
    phi, theta, omega = _asarray(phi, rdt), _asarray(theta, rdt), _asarray(omega, rdt)
    batched = phi.dim() > 0 or theta.dim() > 0 or omega.dim() > 0
    ang = torch.stack(torch.broadcast_tensors(phi.reshape(-1), theta.reshape(-1), omega.reshape(-1)), dim=-1)
    ket = _asarray(ket, dt)
    ket = ket.reshape(-1, 2) if ket.dim() >= 2 else ket.reshape(1, 2)
    f = _RotFidelityFn.apply(ang, ket, dt)
    return f if batched else f[0]


class _UnitaryCostFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, thetas, phis, omegas, kin, kout, cdtype):     # (B,K), (B,K), (B,N) real; kets (M,N) complex
        _require_cuda("unitary_learning_cost")
        thetas, phis, omegas = thetas.contiguous(), phis.contiguous(), omegas.contiguous()
        kin, kout = kin.contiguous(), kout.contiguous()
        B, N = omegas.shape
        rdt = _real_of(cdtype)
        loss = torch.empty(B, dtype=rdt, device=thetas.device)
        tb, pb, ob = torch.empty_like(thetas), torch.empty_like(phis), torch.empty_like(omegas)
        _call("qg_unitary_cost_fwd_grad", _code(cdtype), _ptr(thetas), _ptr(phis), _ptr(omegas), _ptr(kin), _ptr(kout),
              kin.shape[0], _ptr(loss), _ptr(tb), _ptr(pb), _ptr(ob), B, N, _stream())
        ctx.save_for_backward(tb, pb, ob)
        return loss

    @staticmethod
    def backward(ctx, g):
        tb, pb, ob = ctx.saved_tensors
        g = g.unsqueeze(-1)
        return g * tb, g * pb, g * ob, None, None, None


def unitary_learning_cost(thetas, phis, omegas, inputs, outputs, dtype=None):
    """Fused cost of examples/Unitary-Learning-qgrad.py:123-151,
    ``1 - mean_k |Re <out_k| Unitary(N)(thetas, phis, omegas) |in_k>|``, for one parameter set or a batch of them
    (leading dimension) sharing the training pairs ``inputs`` / ``outputs`` (lists of (N,1) kets or (M,N) arrays);
    value and gradient come out of one kernel pass (N <= 16)."""
    dt = dtype or _DEFAULT["dtype"]
    rdt = _real_of(dt)
    thetas, phis, omegas = _asarray(thetas, rdt), _asarray(phis, rdt), _asarray(omegas, rdt)
    batched = omegas.dim() > 1
    N = omegas.shape[-1]

    def kets(x):
        if isinstance(x, (list, tuple)):
            x = torch.stack([_asarray(v, dt).reshape(-1) for v in x])
        return _asarray(x, dt).reshape(-1, N)

    out = _UnitaryCostFn.apply(thetas.reshape(-1, thetas.shape[-1]), phis.reshape(-1, phis.shape[-1]), omegas.reshape(-1, N),
                               kets(inputs), kets(outputs), dt)
    return out if batched else out[0]


# ----------------------------------------------------------------------------- generators (on the device)
def _seed(seed):
    """the reference draws ``onp.random.randint(1000)`` when no seed is given (qgrad_qutip.py:568, 586, 610)"""
    return int(np.random.randint(1000)) if seed is None else int(seed)


def rand_kets(N, batch, seed=0, first=0, dtype=None, kind=_lib.QG_RAND_HAAR):
    """``batch`` Haar-random kets of dimension N, (batch, N, 1), generated on the device (Philox4x32-10, ``qg_rand_ket``):
    kets ``first .. first + batch - 1`` of the sequence ``seed`` -- a rank that owns kets [lo, hi) of a sharded data set
    passes ``first=lo`` and gets exactly the slice a single process would have generated.  Replaces the host-side QuTiP
    ``rand_ket`` loops of examples/Unitary-Learning-qgrad.py:63,88."""
    _require_cuda("rand_kets")
    dt = dtype or _DEFAULT["dtype"]
    out = torch.empty((batch, N), dtype=dt, device=_device())
    _call("qg_rand_ket", _code(dt), _ptr(out), batch, N, int(seed), int(first), int(kind), _stream())
    return out.unsqueeze(-1)


def rand_herm(N, batch=None, seed=0, first=0, scale=None, dtype=None):
    """GUE-distributed Hermitian matrices ``scale (X + X^H) / 2`` on the device (``qg_rand_hermitian``); ``scale=None``
    is the named H / sqrt(N) of the benchmarks, ``scale=1`` tenpy's ``GUE((d, d))`` up to its
    normalisation (examples/Unitary-Learning-no-qgrad.py:106-107).  (N, N), or (batch, N, N)."""
    _require_cuda("rand_herm")
    dt = dtype or _DEFAULT["dtype"]
    b = 1 if batch is None else int(batch)
    out = torch.empty((b, N, N), dtype=dt, device=_device())
    _call("qg_rand_hermitian", _code(dt), _ptr(out), b, N, int(seed), int(first), float(1.0 / math.sqrt(N) if scale is None else scale),
          _stream())
    return out[0] if batch is None else out


def rand_uniform(count, seed=0, lo=0.0, hi=1.0, dtype=None):
    """``count`` reals uniform in [lo, hi) on the device (``qg_rand_uniform``), of the real type matching ``dtype``."""
    _require_cuda("rand_uniform")
    dt = dtype or _DEFAULT["dtype"]
    out = torch.empty(int(count), dtype=_real_of(dt), device=_device())
    _call("qg_rand_uniform", _code(dt), _ptr(out), int(count), int(seed), float(lo), float(hi), _stream())
    return out


def rand_ket(N, seed=None, dtype=None):
    """qgrad_qutip.py:558-573 -- ``uniform(key) + 1j * uniform(key)`` with ONE key (so re == im), normalised, (N, 1).
    Generated on the device; the stream is Philox, not JAX's threefry (the reference's tests pin the norm only)."""
    return rand_kets(N, 1, _seed(seed), 0, dtype, _lib.QG_RAND_REFERENCE)[0]


def rand_dm(N, seed=None, dtype=None):
    """qgrad_qutip.py:576-591 -- ``to_dm(rand_ket(N, seed))`` (rank one, as the reference)."""
    return to_dm(rand_ket(N, seed, dtype))


def rand_unitary(N, seed=None, dtype=None):
    """qgrad_qutip.py:594-620 -- ``Unitary(N)`` on N^2 angles uniform in [0, 2 pi), split N(N-1)/2, N(N-1)/2, N; the
    angles are drawn on the device and never visit the host."""
    params = rand_uniform(N * N, _seed(seed), 0.0, 2 * math.pi, dtype)
    h = N * (N - 1) // 2
    return Unitary(N, dtype)(params[:h], params[h:2 * h], params[2 * h:])


# ----------------------------------------------------------------------------- jax.grad analogue
def grad(fun, argnums=0):
    """``jax.grad`` analogue.  ``fun`` must return a real scalar tensor.  For complex arguments
    the result follows JAX's convention dL/dx - i dL/dy (the conjugate of torch's ``.grad``;
    SURVEY.md 7.2 item 4), so SNAP's complex alphas can be fed to Adam exactly as in the
    reference.  Arguments may be tensors / arrays / scalars or (nested) lists of them."""
    single = isinstance(argnums, int)
    nums = (argnums,) if single else tuple(argnums)

    def leafify(x):
        if isinstance(x, (list, tuple)):
            return type(x)(leafify(v) for v in x)
        x = _asarray(x)
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
