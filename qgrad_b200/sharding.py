"""Batch sharding across ranks: contiguous split of independent units
(Hamiltonian settings, parameter sets, kets) with no data-path collective; the learning step adds
one all-reduce of the shared-parameter gradient."""
from __future__ import annotations


def shard_range(n_units: int, world: int, rank: int):
    """Contiguous [lo, hi) of ``n_units`` owned by ``rank``; remainders go to the low ranks."""
    if world < 1 or not (0 <= rank < world) or n_units < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_units: int, world: int):
    return [shard_range(n_units, world, r)[1] - shard_range(n_units, world, r)[0] for r in range(world)]


class GradientReducer:
    """The one exchange step of the sharded learning loop: every rank accumulates the partial
    shared-parameter gradient of ITS settings and its partial loss sum into one buffer ``[thetabar_0 .. thetabar_{P-1} |
    sum |Re o|]`` of P + 1 reals; one sum all-reduce combines the ranks; the loss is ``1 - sum / m_total`` with
    ``m_total`` the number of kets over ALL ranks (the cotangents were already scaled by 1 / m_total, so the reduced
    gradient is the gradient of the global mean).  Pure host logic over ``torch.distributed``: the CUDA learner
    (learning.HamiltonianLearner) and the CPU gloo test drive the same object.

    ``group``: a process group; None = the default group when one is initialised with more than one rank; False =
    never reduce (score these settings alone, even inside a multi-rank job)."""

    def __init__(self, n_params, m_total, dtype, device, group=None):
        import torch
        self.P, self.m_total, self.group = int(n_params), int(m_total), group
        self.buf = torch.zeros(self.P + 1, dtype=dtype, device=device)

    @property
    def grad(self):
        return self.buf[: self.P]

    @property
    def loss_partial(self):
        return self.buf[self.P:]

    def reduces(self):
        import torch.distributed as dist
        if self.group is False:
            return False
        return self.group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)

    def all_reduce(self):
        if self.reduces():
            import torch.distributed as dist
            dist.all_reduce(self.buf, group=self.group or None)
        return self.buf

    def loss(self):
        return 1.0 - self.buf[self.P] / self.m_total
