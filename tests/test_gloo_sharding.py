"""world_size-2 (and 3: ragged shards) gloo tests of the N>1 path's host logic: settings are sharded across ranks with
``shard_range``; every rank writes the partial (thetabar, sum |Re o|) of its own settings into the learner's
``GradientReducer`` buffer (here with the CPU oracle standing in for the kernels -- the reducer, its layout, the
all-reduce and the loss formula are the product's own objects, the ones ``HamiltonianLearner`` drives on the GPU); one
all-reduce of P+1 reals combines them; Adam follows on every rank.  The sharded trajectory (loss of every step, theta
after k steps) equals the single-process trajectory of the full problem."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.examples_ref import Adam
from qgrad_b200.sharding import GradientReducer, shard_range, shard_sizes


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


def _trajectory(world, rank, steps, group):
    """``steps`` training steps of this rank's shard through the product's GradientReducer; returns (losses, theta)"""
    dense, ctrl, theta, psi, out, t = _problem()
    n_set, m_local = psi.shape[0], psi.shape[2]
    lo, hi = shard_range(n_set, world, rank)
    red = GradientReducer(len(theta), n_set * m_local, torch.float64, "cpu", group)
    opt = Adam([theta.copy()], 1e-2)
    losses = []
    for i in range(steps):
        red.buf.zero_()
        red.buf.copy_(_partial(dense, ctrl, opt.x[0], psi, out, t, lo, hi, n_set * m_local))
        red.all_reduce()
        losses.append(float(red.loss()))
        opt.update(i, [red.grad.numpy().copy()])
    return losses, opt.x[0]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert GradientReducer(3, 10, torch.float64, "cpu").reduces()
    assert not GradientReducer(3, 10, torch.float64, "cpu", group=False).reduces()
    losses, theta = _trajectory(world, rank, 3, None)
    q.put((rank, losses, theta))
    dist.destroy_process_group()


def _run(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    results = {r: (l, th) for r, l, th in (q.get(timeout=180) for _ in range(world))}
    [p.join(timeout=60) for p in procs]
    return results


def test_sharded_trajectory_equals_single_process():
    ref_losses, ref_theta = _trajectory(1, 0, 3, False)
    for world in (2, 3):                                  # 4 settings over 3 ranks: shards of 2, 1, 1
        results = _run(world)
        for r in range(world):
            np.testing.assert_allclose(results[r][0], ref_losses, rtol=0, atol=1e-13)
            np.testing.assert_allclose(results[r][1], ref_theta, rtol=1e-13, atol=1e-15)
        for r in range(1, world):                         # every rank holds the same theta, bit for bit
            assert np.array_equal(results[r][1], results[0][1])


def test_shard_ranges_cover_and_balance():
    for n_units in (0, 1, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            rs = [shard_range(n_units, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n_units
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n_units, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n_units


def test_reducer_single_process_is_a_no_op():
    red = GradientReducer(4, 8, torch.float64, "cpu")
    red.buf.copy_(torch.tensor([1.0, 2.0, 3.0, 4.0, 6.0]))
    assert not red.reduces()
    red.all_reduce()
    assert float(red.loss()) == 1.0 - 6.0 / 8 and red.grad.tolist() == [1.0, 2.0, 3.0, 4.0]
