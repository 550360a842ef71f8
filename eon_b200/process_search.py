"""Host-side tail of a batched process search: what ``ProcessSearchJob::doProcessSearch`` does with the
saddle and the two relaxed endpoints (client/ProcessSearchJob.cpp:234-321) and what ``saveData`` writes to
``results.dat`` (client/ProcessSearchJob.cpp:323-363), for every walker of `Walkers.process_search`.

Pure host logic over harvested arrays: connectivity by ``Matter::compare`` (client/Matter.cpp:129-147, the
distinguishable / no-rotation branch eonclient takes for endpoints: per-atom minimum-image distance below
``structure_comparison.distance_difference``, net translation removed first when no atom is fixed,
client/GeometryAnalysis.cpp:240-258), the endpoint swap, barriers and the abort rules with the reference's status
codes (include/eon/MinModeSaddleSearch.h:29-53).  `process_results` takes the ``prefactor.default_value`` path of
the reference; `eon_b200.prefactor.apply_prefactors` then replaces it by the harmonic prefactors of
client/Prefactor.cpp for the walkers that ended well, from one batched force call.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

STATUS_GOOD = 0
STATUS_BAD_NOT_CONNECTED = 6
STATUS_BAD_HIGH_BARRIER = 8
STATUS_BAD_MINIMA = 9
STATUS_NEGATIVE_BARRIER = 14

STATUS_TEXT = {0: "good", 3: "bad high energy", 5: "bad max iterations", 6: "bad not connected", 7: "bad prefactor",
               8: "bad high barrier", 10: "failed prefactor",
               9: "bad minima", 11: "potential failed", 13: "nonlocal abort", 14: "negative barrier",
               16: "no negative mode at saddle", 18: "zero mode abort", 19: "optimizer error"}


def _mic(d: np.ndarray, cell: np.ndarray) -> np.ndarray:
    """``pbc``: minimum image of row vectors (include/eon/Matter.h:42-50: frac -= floor(frac + 0.5))."""
    inv = np.linalg.inv(cell)
    f = d @ inv
    f -= np.floor(f + 0.5)
    return f @ cell


def per_atom_norm(a: np.ndarray, b: np.ndarray, cell: np.ndarray) -> float:
    """``Matter::perAtomNorm`` (client/Matter.cpp:156-168): largest minimum-image distance between matching atoms."""
    return float(np.linalg.norm(_mic(a - b, cell), axis=1).max()) if len(a) else 0.0


def compare(initial: np.ndarray, other: np.ndarray, cell: np.ndarray, any_fixed: bool, distance_difference: float = 0.1,
            remove_translation: bool = True) -> bool:
    """``Matter::compare(matter, indistinguishable=false)`` without rotation check (client/Matter.cpp:142-146)."""
    a = np.asarray(initial, dtype=float).reshape(-1, 3)
    b = np.asarray(other, dtype=float).reshape(-1, 3)
    if a.shape != b.shape:
        return False
    if not any_fixed and remove_translation:  # client/GeometryAnalysis.cpp:240-258
        a = a + _mic(b - a, cell).mean(axis=0)
    return distance_difference > per_atom_norm(a, b, cell)


def process_results(out: Dict[str, object], initial_positions, cell, fixed, max_energy: float,
                    displacement_positions=None, distance_difference: float = 0.1, minimize_first: bool = True,
                    initial_energy: Optional[float] = None, prefactor_default: float = 0.0,
                    potential_type: str = "", random_seed: int = -1) -> List[dict]:
    """One ``results.dat`` dictionary per walker from `Walkers.process_search` output (`out`: harvests "saddle",
    "min1", "min2" with geometry).  Keys and order follow ``ProcessSearchJob::saveData``; ``reactant`` / ``product``
    hold the endpoint positions after the reference's swap rule."""
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    init = np.asarray(initial_positions, dtype=float).reshape(-1, 3)
    # translationRemove applies when Matter::numberOfFixedAtoms() == 0, an atom being fixed when all its axes are
    any_fixed = bool(np.asarray(fixed).reshape(len(init), -1).all(axis=1).any())
    sad, m1, m2 = out["saddle"], out["min1"], out["min2"]
    W = len(sad.status)
    results = []
    for w in range(W):
        fc_saddle = int(sad.force_calls[w])
        status = int(sad.status[w])
        e_s = float(sad.energies[w])
        p1 = m1.positions[w].reshape(-1, 3)
        p2 = m2.positions[w].reshape(-1, 3)
        e1, e2 = float(m1.energies[w]), float(m2.energies[w])
        fc_min = 0
        barriers = [0.0, 0.0]
        if status == STATUS_GOOD:  # ProcessSearchJob.cpp:184-186: a failed climb ends the job there
            fc_min = int(m1.force_calls[w]) + int(m2.force_calls[w])
            if int(m1.status[w]) != 0 or int(m2.status[w]) != 0:
                status = STATUS_BAD_MINIMA
            else:
                c1 = compare(init, p1, cell, any_fixed, distance_difference)
                c2 = compare(init, p2, cell, any_fixed, distance_difference)
                if not c1 and c2:  # :238-242
                    p1, p2, e1, e2, c1, c2 = p2, p1, e2, e1, c2, c1
                if not c1 or c2:  # :244-273: neither endpoint, or both, is the initial state
                    status = STATUS_BAD_NOT_CONNECTED
                else:
                    if not minimize_first and initial_energy is not None:  # :275-277
                        p1, e1 = init, float(initial_energy)
                    barriers = [e_s - e1, e_s - e2]
                    if max_energy < barriers[0] or max_energy < barriers[1]:
                        status = STATUS_BAD_HIGH_BARRIER
                    elif barriers[0] < 0.0 or barriers[1] < 0.0:
                        status = STATUS_NEGATIVE_BARRIER
        disp_dist = 0.0
        if displacement_positions is not None:
            disp_dist = per_atom_norm(np.asarray(displacement_positions[w], dtype=float).reshape(-1, 3),
                                      sad.positions[w].reshape(-1, 3), cell)
        results.append({
            "termination_reason": status,
            "termination_reason_text": STATUS_TEXT.get(status, "status %d" % status),
            "random_seed": random_seed,
            "potential_type": potential_type,
            "total_force_calls": fc_saddle + fc_min,
            "force_calls_minimization": fc_min,
            "force_calls_saddle": fc_saddle,
            "potential_energy_saddle": e_s,
            "potential_energy_reactant": e1,
            "potential_energy_product": e2,
            "barrier_reactant_to_product": barriers[0],
            "barrier_product_to_reactant": barriers[1],
            "displacement_saddle_distance": disp_dist,
            "force_calls_prefactors": 0,
            "prefactor_reactant_to_product": prefactor_default,
            "prefactor_product_to_reactant": prefactor_default,
            "reactant": p1, "product": p2, "saddle": sad.positions[w].reshape(-1, 3),
        })
    return results


_FLOAT_KEYS = {"potential_energy_saddle", "potential_energy_reactant", "potential_energy_product",
               "barrier_reactant_to_product", "barrier_product_to_reactant", "displacement_saddle_distance",
               "prefactor_reactant_to_product", "prefactor_product_to_reactant"}


def format_results_dat(result: dict) -> str:
    """``<value> <key>`` lines in the order and number format of ``ProcessSearchJob::saveData``."""
    lines = []
    for key, val in result.items():
        if key in ("reactant", "product", "saddle"):
            continue
        lines.append(("%.12e %s" % (val, key)) if key in _FLOAT_KEYS else ("%s %s" % (val, key)))
    return "\n".join(lines) + "\n"
