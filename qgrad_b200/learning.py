"""Host-side sequencing of one Hamiltonian-learning step (BASELINE config #5).

Workload definition: the cost of examples/Unitary-Learning-qgrad.py:123-151 /
examples/Unitary-Learning-no-qgrad.py:165-191,

    L(theta) = 1 - (1/M) sum_{s,k} | Re < out_{s,k} | U_s(theta) | in_{s,k} > | ,
    U_s = exp(-i t H_s(theta)),   H_s(theta) = sum_p ctrl[s, p] theta_p P_p   (Pauli strings),

its gradient by the Frechet pull-back (replacing the finite differences of
Unitary-Learning-no-qgrad.py:213-241), and the Adam update
(jax.experimental.optimizers.adam, Unitary-Learning-qgrad.py:218).

One step, all on the device:  A_s = -i t H_s(theta)  ->  U_s = exp(A_s) (powers, inverse and squaring chain kept)
->  Phi = U Psi_in  ->  overlaps, loss, cotangent G of Phi  ->  Z = G Phi^H (= Ubar U^H, the cotangent in the body
frame)  ->  skew-Hermitian part of Abar (all a real theta can see of it: QG_ADJ_SKEW_BODY)  ->  thetabar by
projection on the Pauli terms  ->  all-reduce  ->  Adam.

Sharding: the ``n_set`` control settings -- each with its own training kets -- are
split across ranks; theta is replicated.  Every rank runs expm forward/backward for its own
settings only, so the expm work scales 1/G; one all-reduce of (thetabar, loss) = P + 1 reals per
step combines the shared-parameter gradient.  ``torch.distributed`` (NCCL) is plumbing; every
kernel on the path is ours (C ABI in include/qgrad_b200.h).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from ._lib import QG_ADJ_CONJ, QG_ADJ_SKEW_BODY, QG_EXPM_FWD_VJP
from ._lib import QgradB200Error
from .sharding import GradientReducer
from .qgrad_qutip import C64, C128, _call, _code, _ptr, _real_of, _require_cuda, _stream, expm_plan


def pauli_hamiltonian(theta, ctrl, xmask, zmask, n_qubits, t, dtype=C128, out=None):
    """A[s] = -i t sum_p ctrl[s,p] theta_p P_p, (n_set, n, n)."""
    _require_cuda("pauli_hamiltonian")
    n_set, P = ctrl.shape
    n = 1 << n_qubits
    A = out if out is not None else torch.empty((n_set, n, n), dtype=dtype, device=theta.device)
    _call("qg_pauli_hamiltonian", _code(dtype), _ptr(theta), _ptr(ctrl), _ptr(xmask), _ptr(zmask), P, n_set, n_qubits,
          float(t), _ptr(A), _stream())
    return A


def pauli_project(Abar, ctrl, xmask, zmask, n_qubits, t, out=None, accumulate=False):
    """thetabar[p] = sum_s ctrl[s,p] Re <dA_s/dtheta_p, Abar_s>."""
    n_set, P = ctrl.shape
    tb = out if out is not None else torch.empty(P, dtype=ctrl.dtype, device=ctrl.device)
    _call("qg_pauli_project", _code(Abar.dtype), _ptr(Abar), _ptr(ctrl), _ptr(xmask), _ptr(zmask), P, n_set, n_qubits,
          float(t), _ptr(tb), 1 if accumulate else 0, _stream())
    return tb


def gemm(A, B, out=None):
    """Batched C = A @ B on the tiled complex GEMM kernel; A (b, m, k), B (b, k, n) contiguous."""
    _require_cuda("gemm")
    b, m, k = A.shape
    n = B.shape[-1]
    Cm = out if out is not None else torch.empty((b, m, n), dtype=A.dtype, device=A.device)
    _call("qg_gemm", _code(A.dtype), _ptr(A), _ptr(B), _ptr(Cm), b, m, n, k, k, n, n, m * k, k * n, m * n, _stream())
    return Cm


def overlap_loss(phi, out_kets, inv_m_total, G=None, loss=None):
    """(G, loss_partial): G[:, :, k] = -sign(Re o_k) inv_m out_k, loss_partial = sum_k |Re o_k|."""
    b, n, m_local = phi.shape
    G = G if G is not None else torch.empty_like(phi)
    if loss is None:
        loss = torch.zeros((), dtype=_real_of(phi.dtype), device=phi.device)
    _call("qg_overlap_loss", _code(phi.dtype), _ptr(phi), _ptr(out_kets), _ptr(G), _ptr(loss), b, n, m_local,
          float(inv_m_total), _stream())
    return G, loss


def conj_transpose(src, out=None):
    """out[b] = src[b]^H for src (batch, rows, cols) (``qg_conj_transpose``)."""
    B, r, c = src.shape
    out = torch.empty((B, c, r), dtype=src.dtype, device=src.device) if out is None else out
    _call("qg_conj_transpose", _code(src.dtype), _ptr(src), _ptr(out), B, r, c, _stream())
    return out


def adam_step_device(x, m, v, grad, lr, step, b1=0.9, b2=0.999, eps=1e-8, loss_partial=None, inv_m_total=0.0, loss_out=None):
    """Adam with the step index in device memory (``step``: int64 tensor of one element, advanced by the kernel) --
    the graph-capturable form; optionally writes ``loss_out = 1 - loss_partial * inv_m_total`` in the same launch."""
    code = _code(C128 if x.dtype == torch.float64 else C64)
    _call("qg_adam_step_device", code, _ptr(x), _ptr(m), _ptr(v), _ptr(grad), x.numel(), float(lr), float(b1), float(b2),
          float(eps), _ptr(step), _ptr(loss_partial), float(inv_m_total), _ptr(loss_out), _stream())
    return x


def adam_step(x, m, v, grad, lr, i, b1=0.9, b2=0.999, eps=1e-8):
    """In-place Adam update (jax.experimental.optimizers.adam defaults)."""
    code = _code(C128 if x.dtype == torch.float64 else C64)
    _call("qg_adam_step", code, _ptr(x), _ptr(m), _ptr(v), _ptr(grad), x.numel(), float(lr), float(b1), float(b2),
          float(eps), int(i), _stream())
    return x


class HamiltonianLearner:
    """One rank's share of the learning problem.

    ctrl (n_set_local, P): control coefficients of this rank's settings; psi_in / psi_out
    (n_set_local, n, m_local): training kets as COLUMNS; m_total: kets over ALL ranks (the loss is
    the mean over all of them).  ``group`` is a torch.distributed process group (None = the default group if one is
    initialised; False = never reduce: score this learner's settings alone even inside a multi-rank job)."""

    def __init__(self, n_qubits, xmask, zmask, ctrl, t, psi_in, psi_out, m_total, dtype=C128, lr=1e-2,
                 max_squarings=None, theta_bound=None, group=None, skew_pullback=True):
        _require_cuda("HamiltonianLearner")
        self.nq, self.n, self.t, self.dtype = n_qubits, 1 << n_qubits, float(t), dtype
        if self.n < 64:
            raise ValueError("HamiltonianLearner needs n_qubits >= 6 (the ket products run on the 64x64-tiled GEMM)")
        dev = torch.device("cuda", torch.cuda.current_device())
        rdt = _real_of(dtype)
        self.xmask = torch.as_tensor(np.asarray(xmask, dtype=np.int32)).to(dev)
        self.zmask = torch.as_tensor(np.asarray(zmask, dtype=np.int32)).to(dev)
        self.ctrl = torch.as_tensor(np.asarray(ctrl)).to(rdt).to(dev).contiguous()
        self.n_set, self.P = self.ctrl.shape
        to = lambda x: (x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(dtype).to(dev).contiguous()
        self.psi_in, self.psi_out = to(psi_in), to(psi_out)
        # Pull-back of the expm: A = -i t H is skew-Hermitian and theta is real, so only the skew-Hermitian part of
        # Abar reaches thetabar.  skew_pullback=True computes exactly that part from the body-frame cotangent
        # Z = Ubar U^H = G Phi^H (QG_ADJ_SKEW_BODY: Hermitian / skew-Hermitian products on half the tiles);
        # False runs the general pull-back from Ubar = G Psi_in^H.  Same thetabar either way (to round-off).
        self.skew = bool(skew_pullback)
        self.psi_in_h = None if self.skew else torch.conj(self.psi_in.transpose(-1, -2)).resolve_conj().contiguous()   # static
        self.m_local = self.psi_in.shape[-1]
        if self.m_local % 64 or self.psi_in.shape != (self.n_set, self.n, self.m_local) or self.psi_out.shape != self.psi_in.shape:
            raise ValueError(f"psi_in / psi_out must be (n_set, n, m_local) = ({self.n_set}, {self.n}, multiple of 64) with kets as "
                             f"columns (the ket products run on the 64x64-tiled GEMM); got {tuple(self.psi_in.shape)}, "
                             f"{tuple(self.psi_out.shape)}")
        self.m_total = int(m_total)
        self.lr, self.group = lr, group
        # a-priori bound on ||A||_1: every Pauli string has 1-norm 1
        if max_squarings is None:
            bound = float(theta_bound if theta_bound is not None else 1.0)
            norm = self.t * float(self.ctrl.abs().sum(dim=1).max()) * bound
            top = 1.3 if dtype == C64 else 0.783
            max_squarings = max(1, int(math.ceil(math.log2(max(norm / top, 1.0)))))
        self.ms = int(max_squarings)
        lib = _lib.load()
        code = _code(dtype)
        nbytes = lib.qg_expm_workspace_bytes(code, self.n, self.n_set, QG_EXPM_FWD_VJP, self.ms)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        shape = (self.n_set, self.n, self.n)
        self.A = torch.empty(shape, dtype=dtype, device=dev)
        self.U = torch.empty(shape, dtype=dtype, device=dev)
        self.Ubar = torch.empty(shape, dtype=dtype, device=dev)
        self.Abar = torch.empty(shape, dtype=dtype, device=dev)
        self.phi = torch.empty_like(self.psi_in)
        self.phi_h = torch.empty((self.n_set, self.m_local, self.n), dtype=dtype, device=dev) if self.skew else None
        self.G = torch.empty_like(self.psi_in)
        self.reducer = GradientReducer(self.P, self.m_total, rdt, dev, group)
        self.red = self.reducer.buf                                   # [thetabar | loss_partial]
        self.m = torch.zeros(self.P, dtype=rdt, device=dev)
        self.v = torch.zeros(self.P, dtype=rdt, device=dev)
        self.step_index = 0
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)     # the same index, where a captured Adam launch reads it
        self.loss_out = torch.zeros(1, dtype=rdt, device=dev)             # loss of the last step()
        self._ring, self._loss_ring, self._loss_events = 8, None, None    # pipelined read-back (step_async)
        self._back = None            # streaming input: back buffers, copy stream, events (enable_streaming)
        self._graphs = None          # CUDA graphs of the local part of a step, keyed by the ket buffers in use
        self.graph_kernel_launches = 0   # kernels launched through graph replays (the C-side counter only sees eager launches)

    # -- streaming input: the next step's kets travel host -> device on a copy stream while this step computes
    def enable_streaming(self):
        """Allocate back buffers for (psi_in, psi_out) -- and psi_in^H for the general pull-back -- and a copy
        stream.  Afterwards ``feed(host...)`` starts an asynchronous host->device copy of the NEXT step's kets
        (pinned host tensors) and ``swap()`` makes them current, waiting on the copy only if it has not finished."""
        if self._back is None:
            self._back = [torch.empty_like(self.psi_in), None if self.skew else torch.empty_like(self.psi_in_h),
                          torch.empty_like(self.psi_out)]
            self._copy_stream = torch.cuda.Stream()
            self._copy_done = torch.cuda.Event()
            self._compute_done = torch.cuda.Event()
            self._compute_done.record()
        return self

    def feed(self, host_in, host_in_h, host_out):
        """``host_in_h`` (psi_in^H) is only read by the general pull-back; pass None with skew_pullback."""
        self._copy_stream.wait_event(self._compute_done)          # the back buffers were the step before last's inputs
        with torch.cuda.stream(self._copy_stream):
            self._back[0].copy_(host_in, non_blocking=True)
            if not self.skew:
                self._back[1].copy_(host_in_h, non_blocking=True)
            self._back[2].copy_(host_out, non_blocking=True)
            self._copy_done.record()

    def swap(self):
        torch.cuda.current_stream().wait_event(self._copy_done)
        front = [self.psi_in, self.psi_in_h, self.psi_out]
        self.psi_in, self.psi_in_h, self.psi_out = self._back
        self._back = front

    # -- the local part of one step: kernels only, no host sync
    def _local(self, theta):
        code, st = _code(self.dtype), _stream
        pauli_hamiltonian(theta, self.ctrl, self.xmask, self.zmask, self.nq, self.t, self.dtype, out=self.A)
        _call("qg_expm_fwd_save", code, _ptr(self.A), _ptr(self.U), self.n_set, self.n, _ptr(self.ws), self.ws.numel(),
              self.ms, st())
        gemm(self.U, self.psi_in, out=self.phi)
        self.red.zero_()
        overlap_loss(self.phi, self.psi_out, 1.0 / self.m_total, G=self.G, loss=self.red[self.P:])
        if self.skew:
            conj_transpose(self.phi, out=self.phi_h)
            gemm(self.G, self.phi_h, out=self.Ubar)               # Z = G Phi^H = Ubar U^H
        else:
            gemm(self.G, self.psi_in_h, out=self.Ubar)
        _call("qg_expm_bwd_saved", code, _ptr(self.A), _ptr(self.Ubar), _ptr(self.Abar), self.n_set, self.n,
              QG_ADJ_SKEW_BODY if self.skew else QG_ADJ_CONJ, _ptr(self.ws), self.ws.numel(), self.ms, st())
        pauli_project(self.Abar, self.ctrl, self.xmask, self.zmask, self.nq, self.t, out=self.red[: self.P])

    # -- CUDA-graph replay (whole-step capture)
    def enable_graphs(self, theta, whole_step=True):
        """Capture a step into a CUDA graph and replay it from then on.  ``whole_step=True`` (default) captures
        EVERYTHING ``step`` does -- Hamiltonian build, expm forward, ket products, loss, Frechet pull-back, projection,
        the NCCL all-reduce of (thetabar, loss), Adam (step index in device memory) and the loss write -- so a training
        step is one graph launch with no host work in between; ``loss_and_grad`` replays the local part only (up to the
        projection), as does ``step`` with ``whole_step=False``.  ``theta`` must be the tensor later passed to
        ``step`` / ``loss_and_grad`` (its address is baked in).  With streaming input one graph is captured per
        ket-buffer set, lazily."""
        self._graphs = {}
        self._graph_theta = theta
        self._graph_whole = bool(whole_step)
        self._graph_stream = torch.cuda.Stream()
        return self

    def reset(self, theta, theta0):
        """Back to the start of training: theta <- theta0 (in place: captured graphs keep pointing at it), Adam moments and
        step counters zeroed."""
        theta.copy_(theta0)
        self.m.zero_(); self.v.zero_(); self.step_dev.zero_()
        self.step_index = 0

    def release_graphs(self):
        """Drop every captured graph (back to eager launches).  Call before ``destroy_process_group()``: a live CUDA
        graph that holds captured NCCL kernels keeps the communicator busy and its destruction blocks."""
        if self._graphs:
            torch.cuda.synchronize()
            self._graphs.clear()
        self._graphs = None

    def _graph_for(self, theta, whole):
        """(graph, kernel nodes of ours in it) for the current ket buffers, captured on first use."""
        key = (self.psi_in.data_ptr(), whole)
        g = self._graphs.get(key)
        if g is None:
            s = self._graph_stream
            s.wait_stream(torch.cuda.current_stream())
            if whole:
                # the warm-up run below must not move the training state: snapshot, run, restore
                keep = [t.clone() for t in (theta, self.m, self.v, self.step_dev)]
            with torch.cuda.stream(s):
                # warm-up on the capture stream: first-use allocations, kernel attributes, the NCCL communicator
                self._local(theta)
                if whole:
                    self._tail(theta)
            torch.cuda.current_stream().wait_stream(s)
            if whole:
                torch.cuda.synchronize()
                for t, k in zip((theta, self.m, self.v, self.step_dev), keep):
                    t.copy_(k)
            graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(graph, stream=s):
                self._local(theta)
                if whole:
                    self._tail(theta)
            g = (graph, _lib.launch_count() - before)
            self._graphs[key] = g
        return g

    def _run_local(self, theta):
        if self._graphs is None or theta.data_ptr() != self._graph_theta.data_ptr():
            return self._local(theta)
        g = self._graph_for(theta, False)
        g[0].replay()
        self.graph_kernel_launches += g[1]

    def _reduce(self):
        self.reducer.all_reduce()

    def _tail(self, theta):
        """all-reduce of (thetabar, loss partial) -> Adam -> loss: everything of a step after the local part"""
        self._reduce()
        adam_step_device(theta, self.m, self.v, self.red[: self.P], self.lr, self.step_dev,
                         loss_partial=self.red[self.P:], inv_m_total=1.0 / self.m_total, loss_out=self.loss_out)

    def loss_and_grad(self, theta):
        """(loss, thetabar) as device tensors, reduced over all ranks."""
        self._run_local(theta)
        self._reduce()
        if self._back is not None:
            self._compute_done.record()
        return self.reducer.loss(), self.reducer.grad

    def step(self, theta):
        """One full training step in place on ``theta``; returns the loss as a device scalar (a view of the learner's
        loss slot: read it -- or ``.clone()`` it -- before the next step overwrites it)."""
        if self._graphs is not None and self._graph_whole and theta.data_ptr() == self._graph_theta.data_ptr():
            g = self._graph_for(theta, True)
            g[0].replay()
            self.graph_kernel_launches += g[1]
        else:
            self._run_local(theta)
            self._tail(theta)
        self.step_index += 1
        if self._back is not None:
            self._compute_done.record()
        return self.loss_out[0]

    # -- pipelined read-back: the loss of step i travels device -> pinned host memory behind the step; the host picks it
    #    up one or more steps later, so reading every step's result never drains the launch queue
    def step_async(self, theta):
        """``step`` + an asynchronous device->host copy of its loss; returns a ticket for :meth:`loss_result`."""
        if self._loss_ring is None:
            self._loss_ring = torch.empty(self._ring, dtype=self.loss_out.dtype).pin_memory()
            self._loss_events = [torch.cuda.Event() for _ in range(self._ring)]
        self.step(theta)
        slot = (self.step_index - 1) % self._ring
        self._loss_ring[slot: slot + 1].copy_(self.loss_out, non_blocking=True)
        self._loss_events[slot].record()
        return self.step_index - 1

    def loss_result(self, ticket):
        """Host value of the loss of the step ``ticket`` refers to (waits for that step only).  Raises if the step went
        non-finite -- a theta that outgrew ``theta_bound`` makes the expm kernel flag the matrix (NaN) instead of
        launching more squarings than the workspace was sized for."""
        if self.step_index - 1 - ticket >= self._ring:
            raise QgradB200Error("loss_result: ticket is older than the read-back ring")
        slot = ticket % self._ring
        self._loss_events[slot].synchronize()
        val = float(self._loss_ring[slot])
        if not math.isfinite(val):
            self.check_plan()
            raise QgradB200Error(f"learning step {ticket}: loss is not finite")
        return val

    def check_plan(self):
        """Raise if the last forward pass flagged a matrix as needing more squarings than ``max_squarings`` (theta
        outgrew ``theta_bound``): re-create the learner with a larger bound.  One device->host read."""
        s = expm_plan(self.ws, self.n_set).cpu().numpy()
        off = (-self.ws.data_ptr()) % 256
        flags = self.ws[off + 4 * self.n_set: off + 8 * self.n_set].view(torch.int32).cpu().numpy()
        if flags.any():
            raise QgradB200Error(f"expm needs more than max_squarings = {self.ms} squarings for settings "
                                 f"{np.nonzero(flags)[0].tolist()} (planned s = {s.tolist()}): theta outgrew theta_bound; "
                                 f"construct the learner with a larger theta_bound / max_squarings")
        return s

    # algorithmic flops of the local part of one step (SURVEY.md 8d conventions: 8 flops / complex MAC)
    def gemm_equivalents(self, s):
        """n^3-complex-MAC products one setting's expm value + pull-back executes (tile-exact: Hermitian / skew-
        Hermitian results are computed on the upper triangle of 64x64 tiles, t = (T+1)/2T of a full product)."""
        T = ((self.n + 63) // 64)
        t = (T + 1) / (2.0 * T)
        fwd = 4 * t + 1 + 1 + s                       # A2, A4, A6, U (structured); inverse; R0; squarings
        bwd = (4 * 2 * t + 1 + t + s * (1 + t)) if self.skew else (10 + 2 * s)
        return fwd + bwd

    def flops_per_step(self, s):
        n, np_ = self.n, (self.n + 63) // 64 * 64
        expm = 8.0 * np_ ** 3 * self.gemm_equivalents(s)
        kets = 2 * 8.0 * n * n * self.m_local
        return self.n_set * (expm + kets)
