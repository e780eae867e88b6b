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
