"""Epicentres and displacements for batches of walkers.

Coordination numbers come from the device neighbour list (`B200Potential.neighbors`, the vesin-exact cell-list
kernel) instead of the reference's O(N^2) loop (client/EpiCenters.cpp:188-203: pairs with minimum-image distance
below `neighbor_cutoff`, 3.3 A by default, include/eon/Parameters.h:166); for cells at least two cutoffs wide
the two counts are identical.  Epicentre rules follow client/EpiCenters.cpp:52-112 (least-coordinated free
atom, random among ties; last atom; random free atom), the displacement itself eon/displace.py:180-263
(`eon_b200.workloads.displace`).  Random picks use numpy generators seeded by the caller -- the reference's
generator is its own, so picks are reproducible here, not bit-identical to eonclient's.
"""
from __future__ import annotations

import numpy as np

from .workloads import displace  # noqa: F401  (re-exported: Gaussian displacement around an epicentre)


def coordination(pot, positions, box, cutoff: float = 3.3) -> np.ndarray:
    """Number of atoms within `cutoff` of every atom (client/EpiCenters.cpp:188-203), from the device list."""
    R = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
    pairs, _ = pot.neighbors(R, np.asarray(box, dtype=np.float64).reshape(9), cutoff, full=True)
    return np.bincount(pairs[:, 0], minlength=R.shape[0]).astype(np.int64)


def _free(fixed, n):
    f = np.asarray(fixed).reshape(n, -1)
    return f.sum(axis=1) == 0  # Matter::getFixed(i): any fixed axis counts as fixed (include/eon/Matter.h:217-227)


def min_coordinated_atoms(pot, positions, box, fixed, cutoff: float = 3.3) -> np.ndarray:
    """Free atoms whose coordination equals the smallest coordination among free atoms
    (minCoordination + coordinationLessOrEqual, client/EpiCenters.cpp:205-218,244-258)."""
    c = coordination(pot, positions, box, cutoff)
    free = _free(fixed, len(c))
    if not free.any():
        raise ValueError("no free atom")
    cmin = c[free].min()
    return np.where(free & (c <= cmin))[0]


def min_coordinated_epicentre(pot, positions, box, fixed, rng: np.random.Generator, cutoff: float = 3.3) -> int:
    """client/EpiCenters.cpp:52-84: a random member of the least-coordinated free atoms."""
    cand = min_coordinated_atoms(pot, positions, box, fixed, cutoff)
    return int(cand[int(rng.integers(len(cand)))])


def last_atom_epicentre(positions) -> int:
    """client/EpiCenters.cpp:86-91."""
    return int(np.asarray(positions).reshape(-1, 3).shape[0] - 1)


def random_free_atom_epicentre(fixed, n_atoms: int, rng: np.random.Generator) -> int:
    """client/EpiCenters.cpp:93-112."""
    free = np.where(_free(fixed, n_atoms))[0]
    return int(free[int(rng.integers(len(free)))])
