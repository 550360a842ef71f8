"""Harmonic rate prefactors of a batch of processes through ONE batched force call.

``Prefactor::getPrefactors`` of the reference (client/Prefactor.cpp:22-167) picks the atoms that moved in a process,
builds three finite-difference Hessians over them (minimum 1, saddle, minimum 2; client/Hessian.cpp:150-275), takes
their eigenvalues and forms the HTST (or quasi-quantum HTST) prefactor of both directions.  Every Hessian is 3 M
displaced force evaluations (6 M with the central scheme); the reference runs them one ``Matter::getForces`` at a
time.  Here the displaced configurations of all three structures of all processes are the systems of a single
``forceBatch`` call; atom selection, the eigen solve of the small (3 M)^2 matrices and the products stay on the host.

The arithmetic is the reference's: same moved-atom rules (``movedAtoms`` client/Prefactor.cpp:185-227,
``movedAtomsPct`` :229-306), same forward / central formula, mass weighting and symmetrisation, the same
negative-mode checks (:93-131), the overflow-avoiding product order (:137-146) and the 10.18e-15 s time unit.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import warnings

import numpy as np

from .hessian import assemble_hessian, displaced_systems
from .process_search import _mic

TIME_UNIT = 10.18e-15  # s per (amu A^2 / eV)^(1/2), as written in client/Prefactor.cpp:142
STATUS_BAD_PREFACTOR = 7
STATUS_FAILED_PREFACTOR = 10


def _per_atom(fixed, n) -> np.ndarray:
    """``Matter::getFixed(i)``: an atom counts as fixed only when all of its axes are."""
    f = np.asarray(fixed).reshape(n, -1)
    return f.all(axis=1)


def _free_moves(min1, saddle, min2, cell, fixed) -> Tuple[np.ndarray, np.ndarray]:
    """|pbc(saddle - min)| per atom with fixed axes zeroed (client/Prefactor.cpp:193-199)."""
    n = saddle.shape[0]
    free = 1.0 - np.asarray(fixed, dtype=float).reshape(n, -1)
    d1 = _mic(saddle - min1, cell) * free
    d2 = _mic(saddle - min2, cell) * free
    return np.linalg.norm(d1, axis=1), np.linalg.norm(d2, axis=1)


def _distance_rows(saddle, cell, i) -> np.ndarray:
    """``Matter::distance(i, j)`` for all j (client/Matter.cpp:360-364): minimum-image distance."""
    return np.linalg.norm(_mic(saddle - saddle[i], cell), axis=1)


def moved_atoms(min1, saddle, min2, cell, fixed, min_displacement: float = 0.25, within_radius: float = 3.3) -> List[int]:
    """``Prefactor::movedAtoms`` (client/Prefactor.cpp:185-227): atoms displaced by more than `min_displacement`
    between the saddle and either minimum, each followed by its free neighbours within `within_radius` of it at the
    saddle, in the reference's order of discovery."""
    min1, saddle, min2 = (np.asarray(a, dtype=float).reshape(-1, 3) for a in (min1, saddle, min2))
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    n = saddle.shape[0]
    is_fixed = _per_atom(fixed, n)
    n1, n2 = _free_moves(min1, saddle, min2, cell, fixed)
    moved: List[int] = []
    seen = np.zeros(n, dtype=bool)
    for i in range(n):
        if n1[i] > min_displacement or n2[i] > min_displacement:
            if not seen[i]:
                moved.append(i)
                seen[i] = True
            near = (_distance_rows(saddle, cell, i) < within_radius) & ~is_fixed & ~seen
            for j in np.nonzero(near)[0]:
                moved.append(int(j))
                seen[j] = True
    return moved


def moved_atoms_pct(min1, saddle, min2, cell, fixed, filter_fraction: float = 0.9, within_radius: float = 3.3) -> List[int]:
    """``Prefactor::movedAtomsPct`` (client/Prefactor.cpp:229-306): the largest movers that make up `filter_fraction`
    of the summed motion, then the free neighbours of each within `within_radius` at the saddle."""
    min1, saddle, min2 = (np.asarray(a, dtype=float).reshape(-1, 3) for a in (min1, saddle, min2))
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    n = saddle.shape[0]
    is_fixed = _per_atom(fixed, n)
    n_free = int((~is_fixed).sum())
    n1, n2 = _free_moves(min1, saddle, min2, cell, fixed)
    diff = np.maximum(n1, n2)
    total = float(diff.sum())
    mini = 0
    for i in range(n):  # last index of the minimum (:252-258)
        if diff[i] <= diff[mini]:
            mini = i
    moved: List[int] = []
    seen = np.zeros(n, dtype=bool)
    d = 0.0
    while total > 0.0 and d / total <= filter_fraction and len(moved) < n_free:
        maxi = mini
        for i in range(n):  # :268-276: the last not-yet-taken atom that is at least as large as the running pick
            if diff[i] >= diff[maxi] and not seen[i]:
                maxi = i
        moved.append(maxi)
        seen[maxi] = True
        d += diff[maxi]
    n_moved = len(moved)
    for k in range(n_moved):
        near = (_distance_rows(saddle, cell, moved[k]) < within_radius) & ~is_fixed & ~seen
        for j in np.nonzero(near)[0]:
            moved.append(int(j))
            seen[j] = True
    return moved


def remove_zero_freqs(freqs: np.ndarray, n_atoms: int, zero_freq_value: float = 1e-6) -> Optional[np.ndarray]:
    """``Hessian::removeZeroFreqs`` (client/Hessian.cpp:277-305): only when the Hessian spans all 3 N coordinates.  A
    count of trivial modes other than 6 is only logged there; the shortened vector is returned all the same and the
    size check of `rate_prefactors` decides."""
    if freqs.size != 3 * n_atoms:
        return freqs
    keep = np.abs(freqs) > zero_freq_value
    if freqs.size - int(keep.sum()) != 6:
        warnings.warn("removeZeroFreqs: found %d trivial eigenmodes instead of 6" % (freqs.size - int(keep.sum())))
    return freqs[keep]


def rate_prefactors(f_min1: np.ndarray, f_saddle: np.ndarray, f_min2: np.ndarray, rate: str = "htst",
                    temperature: float = 300.0) -> Optional[Tuple[float, float]]:
    """Eigenvalue checks and products of client/Prefactor.cpp:76-165; ``None`` where the reference returns -1."""
    if f_min1.size != f_saddle.size or f_min2.size != f_saddle.size or f_saddle.size == 0:
        return None
    if (f_min1 < 0).sum() != 0 or (f_saddle < 0).sum() != 1 or (f_min2 < 0).sum() != 0:
        return None
    p1 = p2 = 1.0
    if rate == "htst":
        for i in range(f_saddle.size):
            p1 *= f_min1[i]
            p2 *= f_min2[i]
            if f_saddle[i] > 0:
                p1 /= f_saddle[i]
                p2 /= f_saddle[i]
        p1 = math.sqrt(p1) / (2 * math.pi * TIME_UNIT)
        p2 = math.sqrt(p2) / (2 * math.pi * TIME_UNIT)
    elif rate == "qqhtst":
        kB_T = temperature * 8.617332e-5
        h_bar, h = 6.582119e-16, 4.135667e-15
        t = h_bar / (2.0 * kB_T)
        for i in range(f_min1.size):
            p1 *= math.sinh(t * (math.sqrt(f_min1[i]) / TIME_UNIT))
            p2 *= math.sinh(t * (math.sqrt(f_min2[i]) / TIME_UNIT))
            if f_saddle[i] > 0:
                s = math.sinh(t * (math.sqrt(f_saddle[i]) / TIME_UNIT))
                p1 /= s
                p2 /= s
        p1 = 2.0 * kB_T / h * p1
        p2 = 2.0 * kB_T / h * p2
    else:
        raise ValueError("unknown prefactor rate %r" % rate)
    return float(p1), float(p2)


def get_prefactors(pot, min1, saddle, min2, atomic_nrs, cell, fixed, masses, dr: float = 0.01, central: bool = False,
                   filter_scheme: str = "fraction", filter_fraction: float = 0.9, min_displacement: float = 0.25,
                   within_radius: float = 3.3, rate: str = "htst", temperature: float = 300.0,
                   check_rotation: bool = False, zero_freq_value: float = 1e-6, remove_net_force: bool = True,
                   max_systems_per_call: int = 1 << 16):
    """Prefactors of P processes.  `min1`, `saddle`, `min2`: arrays ``[P, 3 N]`` (``[3 N]`` or ``[N, 3]`` for one process) of
    the reactant, saddle and product of each; shared `atomic_nrs`, `cell`, `fixed` (N x 3 mask), `masses`.

    Returns ``(prefactors [P, 2], ok [P], force_calls [P], freqs)``: `ok[p]` is False where the reference's
    ``getPrefactors`` returns -1; `force_calls[p]` counts the displaced evaluations, which is what the reference adds
    to ``force_calls_prefactors`` (its three undisplaced forces are cached by Matter, here they ride along in the
    batch); `freqs[p]` are the three eigenvalue vectors.
    """
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    N = int(np.asarray(atomic_nrs).size)
    n3 = 3 * N
    A = [np.asarray(a, dtype=np.float64).reshape(-1, n3) for a in (min1, saddle, min2)]
    P = A[1].shape[0]
    fixed = np.asarray(fixed).reshape(N, -1)
    any_fixed = bool(fixed.any())
    atom_lists: List[List[int]] = []
    blocks, spans = [], []
    for p in range(P):
        m1, sd, m2 = (a[p].reshape(N, 3) for a in A)
        if filter_scheme == "fraction":
            atoms = moved_atoms_pct(m1, sd, m2, cell, fixed, filter_fraction, within_radius)
        else:
            atoms = moved_atoms(m1, sd, m2, cell, fixed, min_displacement, within_radius)
        atom_lists.append(atoms)
        for a in (m1, sd, m2):
            R = displaced_systems(a, atoms, dr, central) if atoms else np.empty((0, n3))
            spans.append(R.shape[0])
            blocks.append(R)
    R_all = np.concatenate(blocks, axis=0) if blocks else np.empty((0, n3))
    F_all = np.empty((R_all.shape[0], N, 3))
    for s0 in range(0, R_all.shape[0], max_systems_per_call):  # normally one call
        s1 = min(R_all.shape[0], s0 + max_systems_per_call)
        F, _, _ = pot.forceBatch(s1 - s0, N, R_all[s0:s1], atomic_nrs, cell)
        F_all[s0:s1] = np.asarray(F).reshape(s1 - s0, N, 3)
    pref = np.zeros((P, 2))
    ok = np.zeros(P, dtype=bool)
    calls = np.zeros(P, dtype=np.int64)
    freqs = []
    off = 0
    for p in range(P):
        atoms = atom_lists[p]
        f3 = []
        for k in range(3):
            n_sys = spans[3 * p + k]
            if atoms:
                H = assemble_hessian(F_all[off:off + n_sys], atoms, masses, dr, central, fixed if any_fixed else None,
                                     remove_net_force)
                ev = np.linalg.eigvalsh(H) if np.isfinite(H).all() else None
                if ev is not None and check_rotation:
                    ev = remove_zero_freqs(ev, N, zero_freq_value)
                f3.append(ev)
                calls[p] += n_sys - 1
            else:
                f3.append(None)
            off += n_sys
        freqs.append(f3)
        if any(f is None for f in f3):
            continue
        r = rate_prefactors(f3[0], f3[1], f3[2], rate, temperature)
        if r is not None:
            pref[p] = r
            ok[p] = True
    return pref, ok, calls, freqs


def apply_prefactors(results: Sequence[dict], pot, atomic_nrs, cell, fixed, masses, min_value: float = 1e9,
                     max_value: float = 1e21, **kw) -> None:
    """The prefactor step of ``ProcessSearchJob::doProcessSearch`` (client/ProcessSearchJob.cpp:291-315) for the
    result dictionaries of `process_search.process_results`: walkers with status 0 get their prefactors from ONE
    batched force call; a failed calculation ends as status 10, a prefactor outside [min_value, max_value] as 7."""
    from .process_search import STATUS_TEXT
    good = [i for i, r in enumerate(results) if r["termination_reason"] == 0]
    if not good:
        return
    m1 = np.stack([np.asarray(results[i]["reactant"], dtype=float).reshape(-1) for i in good])
    sd = np.stack([np.asarray(results[i]["saddle"], dtype=float).reshape(-1) for i in good])
    m2 = np.stack([np.asarray(results[i]["product"], dtype=float).reshape(-1) for i in good])
    pref, ok, calls, _ = get_prefactors(pot, m1, sd, m2, atomic_nrs, cell, fixed, masses, **kw)
    for k, i in enumerate(good):
        r = results[i]
        if not ok[k]:
            status = STATUS_FAILED_PREFACTOR  # returns before the call count is taken (:297-300)
        else:
            r["force_calls_prefactors"] = int(calls[k])
            r["total_force_calls"] += int(calls[k])
            if not (min_value <= pref[k, 0] <= max_value) or not (min_value <= pref[k, 1] <= max_value):
                status = STATUS_BAD_PREFACTOR
            else:
                status = 0
                r["prefactor_reactant_to_product"] = float(pref[k, 0])
                r["prefactor_product_to_reactant"] = float(pref[k, 1])
        if status:
            r["termination_reason"] = status
            r["termination_reason_text"] = STATUS_TEXT.get(status, "status %d" % status)
