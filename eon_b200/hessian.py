"""Finite-difference Hessian through one batched force call.

``Hessian::calculate`` of the reference (client/Hessian.cpp:150-231) displaces one coordinate of one moved atom at
a time and asks ``Matter::getForces`` for each displaced configuration: 3 M (forward) or 6 M (central) independent
force evaluations.  Here they are the systems of a single ``forceBatch`` call.  Same arithmetic as the reference:
H_ij = -(F+_j - F0_j) / dr (forward, the default) or -(F+_j - F-_j) / (2 dr) (central), divided by
sqrt(m_i m_j), then symmetrised.  Forces are ``Matter``'s: fixed axes zeroed, and the mean force removed when no
atom is fixed (client/Matter.cpp:324-333,531-536).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def _matter_forces(F: np.ndarray, fixed: Optional[np.ndarray], remove_net_force: bool) -> np.ndarray:
    F = F.copy()
    if fixed is not None and np.asarray(fixed).any():
        F *= (1 - np.asarray(fixed).reshape(1, -1, 3))
    elif remove_net_force:
        F -= F.mean(axis=1, keepdims=True)
    return F


def displaced_systems(positions, atoms: Sequence[int], dr: float = 0.01, central: bool = False) -> np.ndarray:
    """The configurations ``Hessian::calculate`` evaluates (client/Hessian.cpp:150-215), one per row: the undisplaced
    one, then +dr on each coordinate of each atom of `atoms`, then (central scheme) -dr on each."""
    R0 = np.asarray(positions, dtype=np.float64).reshape(-1)
    size = 3 * len(atoms)
    n_sys = 1 + (2 if central else 1) * size
    R = np.broadcast_to(R0, (n_sys, R0.size)).copy()
    for i in range(size):
        col = 3 * int(atoms[i // 3]) + i % 3
        R[1 + i, col] += dr
        if central:
            R[1 + size + i, col] -= dr
    return R


def assemble_hessian(F, atoms: Sequence[int], masses=None, dr: float = 0.01, central: bool = False, fixed=None,
                     remove_net_force: bool = True) -> np.ndarray:
    """Hessian from the forces of `displaced_systems` (same row order): finite difference, mass weighting,
    symmetrisation (client/Hessian.cpp:196-231)."""
    F = np.asarray(F, dtype=np.float64)
    n_sys = F.shape[0]
    N = F.size // (3 * n_sys)
    atoms = [int(a) for a in atoms]
    size = 3 * len(atoms)
    m = np.ones(N) if masses is None else np.asarray(masses, dtype=np.float64)
    F = _matter_forces(F.reshape(n_sys, N, 3), fixed, remove_net_force)
    sel = np.array([[3 * a, 3 * a + 1, 3 * a + 2] for a in atoms]).reshape(-1)
    Ff = F.reshape(n_sys, 3 * N)[:, sel]
    if central:
        H = -(Ff[1:1 + size] - Ff[1 + size:1 + 2 * size]) / (2.0 * dr)
    else:
        H = -(Ff[1:1 + size] - Ff[0]) / dr
    mm = np.repeat(m[atoms], 3)
    eff = np.sqrt(np.outer(mm, mm))
    H = np.where(eff < 1e-300, 0.0, H / np.where(eff < 1e-300, 1.0, eff))  # safe_div(.., 0.0)
    return 0.5 * (H + H.T)


def fd_hessian(pot, positions, atomic_nrs, box, atoms: Sequence[int], masses=None, dr: float = 0.01,
               central: bool = False, fixed=None, remove_net_force: bool = True) -> np.ndarray:
    """Mass-weighted, symmetrised FD Hessian over the coordinates of `atoms` (size 3 len(atoms)); `pot` is any object
    with the ``forceBatch(nSystems, nAtoms, positions, atomicNrs, boxes)`` call of `B200Potential`."""
    atoms = [int(a) for a in atoms]
    R = displaced_systems(positions, atoms, dr, central)
    n_sys, N = R.shape[0], R.shape[1] // 3
    F, _, _ = pot.forceBatch(n_sys, N, R, atomic_nrs, box)
    return assemble_hessian(np.asarray(F).reshape(n_sys, N, 3), atoms, masses, dr, central, fixed, remove_net_force)
