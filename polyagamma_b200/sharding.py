"""Sharding of the output index range over the GPUs of one box.

Every output element is a pure function of ``(seed, offset, global index, h, z)``, so the path
shards with NO collective: rank r of G takes the contiguous block ``[r*n//G, (r+1)*n//G)`` of
the flat output range, its slice of h and z, and passes ``first_index = lo`` -- the Philox
counter ranges of different ranks are then disjoint and the concatenation of the shards is bit
for bit what one GPU would have produced.  NCCL / NVLink are not used by the product path.
"""
from __future__ import annotations

import os


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``range(n)`` owned by ``rank`` of ``world``."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n < 0:
        raise ValueError("n must be non-negative")
    return (rank * n) // world, ((rank + 1) * n) // world


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process per GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def random_polyagamma_shard(h, z, *, n_total: int, rank: int, world: int, random_state, method=None,
                            out=None, disable_checks=True, ref_compat=False, device=None):
    """Draw this rank's block of a length-``n_total`` fill2.

    ``h`` and ``z`` are this rank's slices (length ``hi - lo``) or scalars.  ``random_state``
    must be the same ``(seed, offset)`` on every rank."""
    from ._api import random_polyagamma

    lo, hi = shard_range(n_total, rank, world)
    kw = dict(method=method, disable_checks=disable_checks, random_state=random_state,
              ref_compat=ref_compat, first_index=lo, device=device)
    if out is not None:
        return random_polyagamma(h, z, out=out, **kw)
    if isinstance(h, (int, float)) and isinstance(z, (int, float)):
        return random_polyagamma(h, z, size=hi - lo, **kw)
    return random_polyagamma(h, z, **kw)
