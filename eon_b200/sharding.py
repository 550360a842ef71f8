"""Walker sharding over the GPUs of one box and the host-side harvest.

Walkers are independent units -- the reference runs them as separate OS processes
(eon/communicator.py:525-559) and only ever gathers their result files (`harvest_job_files`,
eon/communicator.py:135-167).  Here rank r of G (one process, one stream, one handle per GPU) owns a
contiguous block of walker ids, so the images of a NEB band stay on one GPU
(client/NudgedElasticBand.cpp:581-592 couples neighbouring images); there is no collective on the
force path.  `harvest` is the path's only exchange: every rank contributes its `Harvest` arrays and
rank `dst` receives them concatenated in walker order.
"""
from __future__ import annotations

from dataclasses import fields
from typing import Optional, Tuple

import numpy as np

from .walkers import Harvest


def shard_range(n_walkers: int, rank: int, world: int, band: int = 1) -> Tuple[int, int]:
    """[first, last) walker ids of `rank`; blocks are multiples of `band` walkers (NEB images per band)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    if band < 1 or n_walkers % band:
        raise ValueError("n_walkers must be a multiple of the band size")
    units = n_walkers // band
    lo = (units * rank) // world
    hi = (units * (rank + 1)) // world
    return lo * band, hi * band


def owner_of(walker: int, n_walkers: int, world: int, band: int = 1) -> int:
    """Rank that owns walker id `walker` under `shard_range`."""
    for r in range(world):
        lo, hi = shard_range(n_walkers, r, world, band)
        if lo <= walker < hi:
            return r
    raise ValueError("walker id out of range")


def concat_harvests(parts) -> Harvest:
    """Concatenate per-rank harvests in rank (= walker id) order."""
    parts = [p for p in parts if p is not None]
    out = {}
    for f in fields(Harvest):
        vals = [getattr(p, f.name) for p in parts]
        out[f.name] = None if any(v is None for v in vals) else np.concatenate(vals, axis=0)
    return Harvest(**out)


def harvest(local: Harvest, dst: int = 0, group=None) -> Optional[Harvest]:
    """Gather every rank's results on rank `dst` (None elsewhere).  Without an initialised process group
    (single GPU) it returns `local` unchanged."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    return concat_harvests(gathered)


def gather_bands(local: dict, dst: int = 0, group=None) -> Optional[dict]:
    """The same end-of-run exchange for NEB bands: every rank holds the result dictionary of its own bands
    (`eon_b200.neb.band_forces_resident`: arrays with the band index first); rank `dst` gets them concatenated in band
    order, the others None.  A band never spans ranks (its images are coupled, client/NudgedElasticBand.cpp:581-592), so
    this is the only communication."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = {}
    for key in gathered[0]:
        vals = [g[key] for g in gathered]
        out[key] = None if any(v is None for v in vals) else np.concatenate(vals, axis=0)
    return out
