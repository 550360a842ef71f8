"""Job-directory driver for resident walkers: W saddle-search job directories in, W result sets out.

The reference runs one ``eonclient`` per job directory (eon/communicator.py:525-559): the client reads ``config.ini``,
``pos.con``, ``displacement.con`` and ``direction.dat`` (client/SaddleSearchJob.cpp:27-61), searches, and writes
``results.dat``, ``mode.dat`` and ``saddle.con`` (client/SaddleSearchJob.cpp:138-175).  Here the W directories of one
topology are read on the host, searched in ONE device launch (or streamed through the slots of a handle), and written
back in the reference's formats and key order.  Host I/O only; the search itself never leaves the device.
"""
from __future__ import annotations

import configparser
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import confile
from .potential import B200Potential
from .walkers import Walkers, search_params

# MinModeSaddleSearch::statusMessage (include/eon/MinModeSaddleSearch.h:55-83)
STATUS_TEXT = ["Success", "Initialized", "Initial displacement unable to reach convex region", "Barrier too high",
               "Too many iterations in concave region", "Too many iterations", "Saddle is not connected to initial state",
               "Prefactors not within window", "Energy barrier not within window",
               "Minimizations from saddle did not converge", "Hessian calculation failed", "Potential evaluation failed",
               "Nonnegative initial mode, aborting", "Nonlocal abort", "Negative barrier detected",
               "No reaction found during MD trajectory", "Converged to stationary point with zero negative modes",
               "No forward barrier found along minimized band", "Zero mode abort", "Optimizer error", "Dimer lost mode",
               "Dimer restored best"]

_OPT = {"cg": 0, "lbfgs": 1, "fire": 2}
_METRIC = {"norm": 0, "max_atom": 1, "max_component": 2}
_ROT = {"sd": 0, "cg": 1, "lbfgs": 2}
_MINMODE = {"dimer": 0, "lanczos": 1}
# potential strings of [Potential] potential that this backend serves, with the PotType name results.dat prints
_POT = {"lj": ("lj", "LJ"), "lj_cluster": ("lj_cluster", "LJCLUSTER"), "morse_pt": ("morse_pt", "MORSE_PT"),
        "eam_b200": ("eam", "EAM_B200")}


def _ini(path_or_text: str) -> configparser.ConfigParser:
    cp = configparser.ConfigParser(inline_comment_prefixes=("#", ";"), strict=False)
    if os.path.exists(path_or_text):
        cp.read(path_or_text)
    else:
        cp.read_string(path_or_text)
    return cp


def _bool(v: str) -> int:
    return 1 if str(v).strip().lower() in ("1", "true", "yes", "on") else 0


def search_params_from_ini(path_or_text: str, lib=None):
    """``EonB200SearchParams`` from a reference ``config.ini``: the keys of client/ParametersINI.cpp this path reads
    (:343-432 [Optimizer] / [LBFGS] / [CG] / [FIRE], :445-468 [Dimer], :583-590 [Lanczos], :1054-1170 [Saddle Search],
    [Main] finite_difference), with the reference's cross-group defaults (:1075-1077 the saddle's converged_force
    falls back to the optimizer's).  Unknown method names raise, like the reference's enum casts."""
    cp = _ini(path_or_text)
    kw: Dict[str, object] = {}

    def get(sec, key, conv, field, table=None):
        if cp.has_option(sec, key):
            v = cp.get(sec, key).strip()
            if table is not None:
                if v.lower() not in table:
                    raise ValueError("config.ini: [%s] %s = %s is not served by this backend" % (sec, key, v))
                kw[field] = table[v.lower()]
            else:
                kw[field] = conv(v)

    get("Main", "finite_difference", float, "finite_difference")
    get("Optimizer", "opt_method", None, "opt_method", _OPT)
    get("Optimizer", "convergence_metric", None, "convergence_metric", _METRIC)
    get("Optimizer", "converged_force", float, "opt_converged_force")
    get("Optimizer", "max_iterations", int, "opt_max_iterations")
    get("Optimizer", "max_move", float, "max_move")
    get("FIRE", "time_step", float, "time_step")
    get("FIRE", "time_step_max", float, "max_time_step")
    get("LBFGS", "lbfgs_memory", int, "lbfgs_memory")
    get("LBFGS", "lbfgs_inverse_curvature", float, "lbfgs_inverse_curvature")
    get("LBFGS", "lbfgs_auto_scale", _bool, "lbfgs_auto_scale")
    get("LBFGS", "lbfgs_angle_reset", _bool, "lbfgs_angle_reset")
    get("LBFGS", "lbfgs_distance_reset", _bool, "lbfgs_distance_reset")
    get("CG", "cg_no_overshooting", _bool, "cg_no_overshooting")
    get("CG", "cg_knock_out_max_move", _bool, "cg_knock_out_max_move")
    get("CG", "cg_max_iter_before_reset", int, "cg_max_iter_before_reset")
    get("Dimer", "opt_method", None, "dimer_opt_method", _ROT)
    get("Dimer", "converged_angle", float, "dimer_converged_angle")
    get("Dimer", "rotations_max", int, "dimer_rotations_max")
    get("Lanczos", "tolerance", float, "lanczos_tolerance")
    get("Lanczos", "max_iterations", int, "lanczos_max_iterations")
    get("Lanczos", "quit_early", _bool, "lanczos_quit_early")
    get("Saddle Search", "min_mode_method", None, "minmode_method", _MINMODE)
    get("Saddle Search", "max_iterations", int, "saddle_max_iterations")
    get("Saddle Search", "max_energy", float, "saddle_max_energy")
    get("Saddle Search", "perp_force_ratio", float, "saddle_perp_force_ratio")
    get("Saddle Search", "zero_mode_abort_curvature", float, "saddle_zero_mode_abort_curvature")
    get("Saddle Search", "nonlocal_count_abort", int, "saddle_nonlocal_count_abort")
    get("Saddle Search", "nonlocal_distance_abort", float, "saddle_nonlocal_distance_abort")
    if cp.has_option("Saddle Search", "converged_force"):
        kw["saddle_converged_force"] = float(cp.get("Saddle Search", "converged_force"))
    elif "opt_converged_force" in kw:  # ParametersINI.cpp:1075-1077
        kw["saddle_converged_force"] = kw["opt_converged_force"]
    if cp.has_option("Dimer", "improved") and not _bool(cp.get("Dimer", "improved")):
        raise ValueError("config.ini: the classic dimer ([Dimer] improved = false) is not served by this backend")
    # the reference's time steps are divided by the time unit after parsing (ParametersINI.cpp:1335-1340)
    for k in ("time_step", "max_time_step"):
        if k in kw:
            kw[k] = kw[k] / 10.1805055  # Parameters.h:42 timeUnit
    return search_params(lib=lib, **kw)


def job_settings(path_or_text: str) -> dict:
    """What the driver needs beyond the search parameters: job, potential, random seed."""
    cp = _ini(path_or_text)
    pot = cp.get("Potential", "potential", fallback="lj").strip().lower()
    if pot not in _POT:
        raise ValueError("config.ini: [Potential] potential = %s is not served by this backend (served: %s)"
                         % (pot, ", ".join(sorted(_POT))))
    return {"job": cp.get("Main", "job", fallback="saddle_search").strip().lower(), "potential": _POT[pot][0],
            "potential_type": _POT[pot][1], "random_seed": cp.getint("Main", "random_seed", fallback=-1),
            "eam_conservative": _bool(cp.get("Potential", "eam_b200_conservative", fallback="false"))}


def format_saddle_results(status: int, random_seed: int, potential_type: str, total_force_calls: int,
                          force_calls_saddle: int, iterations: int, e_saddle: float, eigenvalue: float,
                          e_reactant: float) -> str:
    """``results.dat`` of ``SaddleSearchJob::saveData`` (client/SaddleSearchJob.cpp:138-163): same keys, same order,
    ``{:f}`` (six decimals) for the reals."""
    text = STATUS_TEXT[status] if 0 <= status < len(STATUS_TEXT) else "Unknown status"
    lines = ["%d termination_reason" % status, "%s termination_reason_text" % text, "saddle_search job_type",
             "%d random_seed" % random_seed, "%s potential_type" % potential_type,
             "%d total_force_calls" % total_force_calls, "%d force_calls_saddle" % force_calls_saddle,
             "%d iterations" % iterations]
    if status != 11:
        lines += ["%f potential_energy_saddle" % e_saddle, "%f final_eigenvalue" % eigenvalue]
    lines.append("%f potential_energy_reactant" % e_reactant)
    return "\n".join(lines) + "\n"


def parse_results_dat(text: str) -> dict:
    """``<value> <key>`` lines -> dict (values kept as strings except where they parse as numbers)."""
    out = {}
    for line in text.splitlines():
        val, _, key = line.strip().partition(" ")
        if not key:
            continue
        if key == "termination_reason_text":
            val, key = line.strip().rsplit(" ", 1)
        try:
            out[key] = int(val)
        except ValueError:
            try:
                out[key] = float(val)
            except ValueError:
                out[key] = val
    return out


def run_saddle_search_jobs(job_dirs: Sequence[str], out_dirs: Optional[Sequence[str]] = None, device: int = 0, lib=None,
                           pot: Optional[B200Potential] = None, slots: int = 0, groups: int = 32) -> List[dict]:
    """``SaddleSearchJob::run`` + ``saveData`` for every directory of `job_dirs` (all of one topology and one
    ``config.ini`` content, as the jobs an AKMC step hands out are): one launch over all of them, or -- with
    ``0 < slots < len(job_dirs)`` -- through that many resident slots: streamed in `groups` groups, or with
    ``groups <= 0`` taken from a device-side queue inside one launch (``eonb200_walkers_queue``).  Writes ``results.dat``, ``mode.dat`` and
    ``saddle.con`` into the matching `out_dirs` (default: the job directories) and returns the parsed results."""
    if not job_dirs:
        return []
    out_dirs = list(out_dirs) if out_dirs is not None else list(job_dirs)
    cfg0 = os.path.join(job_dirs[0], "config.ini")
    settings = job_settings(cfg0)
    if settings["job"] != "saddle_search":
        raise ValueError("job = %s: this driver runs saddle_search jobs" % settings["job"])
    own_pot = pot is None
    if own_pot:
        pot = B200Potential(settings["potential"], device=device, lib=lib, conservative=bool(settings["eam_conservative"]))
    prm = search_params_from_ini(cfg0, lib=pot._lib)
    try:
        reactants, starts, modes = [], [], []
        for d in job_dirs:
            reactants.append(confile.read_con(os.path.join(d, "pos.con")))
            starts.append(confile.read_con(os.path.join(d, "displacement.con")))
            modes.append(confile.read_mode(os.path.join(d, "direction.dat"), starts[-1].n_atoms))
        f0 = reactants[0]
        for fr in reactants + starts:
            if fr.n_atoms != f0.n_atoms or not np.array_equal(fr.atomic_nrs, f0.atomic_nrs) or \
                    not np.array_equal(fr.fixed, f0.fixed) or not np.allclose(fr.cell, f0.cell, atol=0, rtol=0):
                raise ValueError("job directories of different topologies / cells in one batch")
        W, N = len(job_dirs), f0.n_atoms
        fixed = f0.fixed.astype(np.int32)
        R0 = np.stack([fr.positions for fr in reactants]).reshape(W, 3 * N)
        R = np.stack([fr.positions for fr in starts]).reshape(W, 3 * N)
        M = np.stack(modes).reshape(W, 3 * N)
        calls0 = pot.forceCallCounter
        # reactant energies: Matter::getPotentialEnergy of pos.con (SaddleSearchJob.cpp:33,66); identical reactants
        # are evaluated once
        same = all(np.array_equal(R0[0], R0[w]) for w in range(1, W))
        wk0 = Walkers(pot, f0.atomic_nrs, fixed, f0.cell, R0[:1] if same else R0)
        E0, _ = wk0.matter_forces()
        wk0.close()
        E0 = np.broadcast_to(E0, (W,)).copy() if same else E0
        n_slots = W if slots <= 0 or slots >= W else int(slots)
        wk = Walkers(pot, f0.atomic_nrs, fixed, f0.cell, R[:n_slots], M[:n_slots])
        if n_slots == W:
            wk.saddle_search(prm, E0)
            h = wk.harvest()
        elif groups <= 0:  # device-side queue: one launch, every slot takes the next job directory's search when free
            h = wk.queue_search(prm, R, M, E0)
        else:
            h = wk.stream_search(prm, R, M, E0, groups=groups)
        wk.close()
        total_calls = pot.forceCallCounter - calls0
        results = []
        for w, d in enumerate(out_dirs):
            os.makedirs(d, exist_ok=True)
            status = int(h.status[w])
            # the reference's counter is per process: one reactant evaluation + the search's own calls
            per_job_calls = int(h.force_calls[w]) + 1
            text = format_saddle_results(status, settings["random_seed"], settings["potential_type"], per_job_calls,
                                         int(h.force_calls[w]), int(h.iterations[w]), float(h.energies[w]),
                                         float(h.eigenvalues[w]), float(E0[w]))
            with open(os.path.join(d, "results.dat"), "w") as fh:
                fh.write(text)
            confile.write_mode(os.path.join(d, "mode.dat"), h.modes[w].reshape(N, 3), starts[w].free_mask())
            sad = confile.ConFrame(f0.cell, h.positions[w].reshape(N, 3), f0.atomic_nrs, f0.fixed, f0.masses,
                                   f0.atom_index, list(starts[w].header))
            confile.write_con(os.path.join(d, "saddle.con"), sad)
            r = parse_results_dat(text)
            r["batch_force_calls"] = int(total_calls)
            results.append(r)
        return results
    finally:
        if own_pot:
            pot.close()
