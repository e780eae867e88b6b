"""world_size-2 gloo test of the N>1 path's host logic: settings are sharded across ranks with
``shard_range``, every rank computes the partial (thetabar, sum |Re o|) of its own settings (here
with the CPU oracle standing in for the kernels), one all-reduce of P+1 reals combines them, and
the result equals the single-process gradient of the full problem."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qgrad_b200.sharding import shard_range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _problem(seed=0, nq=3, n_set=4, P=5, m_local=6, t=0.8):
    rng = np.random.default_rng(seed)
    n = 1 << nq
    xm = rng.integers(0, n, P); zm = rng.integers(0, n, P)
    dense = np.zeros((P, n, n), dtype=np.complex128)
    for p in range(P):
        for c in range(n):
            dense[p, c ^ xm[p], c] = (1j) ** bin(xm[p] & zm[p]).count("1") * (-1) ** bin(c & zm[p]).count("1")
    ctrl = rng.uniform(0.5, 1.5, (n_set, P)); theta = rng.standard_normal(P) * 0.5
    psi = rng.standard_normal((n_set, n, m_local)) + 1j * rng.standard_normal((n_set, n, m_local))
    out = rng.standard_normal((n_set, n, m_local)) + 1j * rng.standard_normal((n_set, n, m_local))
    return dense, ctrl, theta, psi, out, t


def _partial(dense, ctrl, theta, psi, out, t, lo, hi, m_total):
    """[thetabar | sum|Re o|] of settings lo..hi, via torch-CPU autograd (oracle-side arithmetic)"""
    th = torch.tensor(theta, requires_grad=True)
    acc = torch.zeros((), dtype=torch.float64)
    for s in range(lo, hi):
        H = torch.tensordot((torch.tensor(ctrl[s]) * th).to(torch.complex128), torch.tensor(dense), 1)
        U = torch.linalg.matrix_exp(-1j * t * H)
        o = torch.sum(torch.conj(torch.tensor(out[s])) * (U @ torch.tensor(psi[s])), dim=0)
        acc = acc + torch.sum(torch.abs(torch.real(o)))
    red = torch.zeros(len(theta) + 1, dtype=torch.float64)
    if hi > lo:
        (-acc / m_total).backward()
        red[:-1] = th.grad
        red[-1] = acc.detach()
    return red


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dense, ctrl, theta, psi, out, t = _problem()
    n_set, m_local = psi.shape[0], psi.shape[2]
    lo, hi = shard_range(n_set, world, rank)
    red = _partial(dense, ctrl, theta, psi, out, t, lo, hi, n_set * m_local)
    dist.all_reduce(red)
    q.put((rank, red.numpy()))
    dist.destroy_process_group()


def test_sharded_gradient_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    results = dict(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    dense, ctrl, theta, psi, out, t = _problem()
    full = _partial(dense, ctrl, theta, psi, out, t, 0, psi.shape[0], psi.shape[0] * psi.shape[2]).numpy()
    for r in range(world):
        np.testing.assert_allclose(results[r], full, rtol=1e-12, atol=1e-14)
