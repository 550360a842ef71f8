"""ctypes binding of the CPU oracle (oracle/liboracle.so) and the compiled reference pieces
(oracle/_ref/*.so).  TEST INFRASTRUCTURE: imported only by tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke() -- never by the eon_b200 package."""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

POT_LJ, POT_LJCLUSTER, POT_MORSE_PT, POT_EAM, POT_EAM_NL = 0, 1, 2, 3, 4
OPT_CG, OPT_LBFGS, OPT_FIRE = 0, 1, 2
MINMODE_DIMER, MINMODE_LANCZOS = 0, 1
ROT_SD, ROT_CG, ROT_LBFGS = 0, 1, 2


class EamRow(C.Structure):
    _fields_ = [("Z", C.c_int), ("Dm", C.c_double), ("alphaM", C.c_double), ("Rm", C.c_double),
                ("beta1", C.c_double), ("beta2", C.c_double), ("r_cut", C.c_double),
                ("func_coeff", C.c_double * 9)]


class Potential(C.Structure):
    _fields_ = [("kind", C.c_int), ("cutoff", C.c_double), ("lj_u0", C.c_double), ("lj_psi", C.c_double),
                ("morse_De", C.c_double), ("morse_a", C.c_double), ("morse_re", C.c_double),
                ("n_eam_rows", C.c_int), ("eam_rows", C.POINTER(EamRow)), ("eam_conservative", C.c_int)]


class SearchParams(C.Structure):
    _fields_ = [("finite_difference", C.c_double), ("remove_net_force", C.c_int), ("opt_method", C.c_int),
                ("convergence_metric", C.c_int), ("opt_max_iterations", C.c_long), ("max_move", C.c_double),
                ("opt_converged_force", C.c_double), ("time_step", C.c_double), ("max_time_step", C.c_double),
                ("lbfgs_memory", C.c_long), ("lbfgs_inverse_curvature", C.c_double),
                ("lbfgs_auto_scale", C.c_int), ("lbfgs_angle_reset", C.c_int), ("lbfgs_distance_reset", C.c_int),
                ("cg_no_overshooting", C.c_int), ("cg_knock_out_max_move", C.c_int),
                ("cg_max_iter_before_reset", C.c_long), ("dimer_opt_method", C.c_int),
                ("dimer_converged_angle", C.c_double), ("dimer_rotations_max", C.c_long),
                ("lanczos_tolerance", C.c_double), ("lanczos_max_iterations", C.c_long),
                ("lanczos_quit_early", C.c_int), ("minmode_method", C.c_int),
                ("saddle_max_iterations", C.c_long), ("saddle_converged_force", C.c_double),
                ("saddle_max_energy", C.c_double), ("saddle_perp_force_ratio", C.c_double),
                ("saddle_zero_mode_abort_curvature", C.c_double), ("saddle_nonlocal_count_abort", C.c_long),
                ("saddle_nonlocal_distance_abort", C.c_double)]


class SaddleResult(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_long), ("force_calls", C.c_long),
                ("energy", C.c_double), ("eigenvalue", C.c_double), ("convergence", C.c_double)]


class RelaxResult(C.Structure):
    _fields_ = [("converged", C.c_int), ("iterations", C.c_long), ("force_calls", C.c_long),
                ("energy", C.c_double), ("convergence", C.c_double)]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


_lib = None


def build_oracle():
    """Compile oracle/liboracle.so (and oracle/_ref when the reference checkout exists)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        _lib = C.CDLL(path)
        _lib.orc_force.restype = C.c_int
        _lib.orc_count_ordered_pairs.restype = C.c_long
        _lib.orc_neighbors.restype = C.c_long
    return _lib


def potential(kind, conservative=False, rows=None, **kw) -> Potential:
    p = Potential()
    lib().orc_potential_default(C.byref(p), int(kind))
    p.eam_conservative = 1 if conservative else 0
    for k, v in kw.items():
        setattr(p, k, v)
    if rows is not None:
        arr = (EamRow * len(rows))(*rows)
        p._keepalive = arr
        p.n_eam_rows = len(rows)
        p.eam_rows = C.cast(arr, C.POINTER(EamRow))
    return p


def search_params(**kw) -> SearchParams:
    p = SearchParams()
    lib().orc_search_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def force(pot: Potential, R, Z, box):
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    F = np.zeros_like(R)
    U = C.c_double()
    rc = lib().orc_force(C.byref(pot), C.c_long(len(Z)), _dp(R), _ip(Z), _dp(F), C.byref(U), _dp(box))
    if rc != 0:
        raise RuntimeError(f"orc_force rc={rc}")
    return U.value, F


def count_ordered_pairs(pot, R, Z, box) -> int:
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    return int(lib().orc_count_ordered_pairs(C.byref(pot), C.c_long(len(Z)), _dp(R), _ip(Z), _dp(box)))


def matter_forces(pot, R, Z, fixed, box, remove_net_force=True):
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    fixed = np.ascontiguousarray(fixed, dtype=np.int32).reshape(-1, 3)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    F = np.zeros_like(R)
    U = C.c_double()
    rc = lib().orc_matter_forces(C.byref(pot), int(remove_net_force), C.c_long(len(Z)), _dp(R), _ip(Z),
                                 _ip(fixed), _dp(box), _dp(F), C.byref(U))
    if rc != 0:
        raise RuntimeError(f"orc_matter_forces rc={rc}")
    return U.value, F


def neighbors(R, box, cutoff, full=True, periodic=(1, 1, 1)):
    """Brute-force oracle list -> (pairs[P,2] int64, shifts[P,3] int32)."""
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    per = (C.c_int * 3)(*[int(x) for x in periodic])
    cap = 0
    pairs = np.zeros((1, 2), dtype=np.int64)
    shifts = np.zeros((1, 3), dtype=np.int32)
    for _ in range(2):
        n = lib().orc_neighbors(C.c_long(len(R)), _dp(R), _dp(box), per, C.c_double(cutoff), int(full),
                                C.c_long(cap), pairs.ctypes.data_as(C.POINTER(C.c_int64)),
                                shifts.ctypes.data_as(C.POINTER(C.c_int32)))
        if n <= cap:
            break
        cap = int(n)
        pairs = np.zeros((cap, 2), dtype=np.int64)
        shifts = np.zeros((cap, 3), dtype=np.int32)
    return pairs[:n].copy(), shifts[:n].copy()


def saddle_search(pot, prm, R, Z, fixed, box, mode, reactant_energy):
    R = np.array(R, dtype=np.float64).reshape(-1, 3).copy()
    mode = np.array(mode, dtype=np.float64).reshape(-1, 3).copy()
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    fixed = np.ascontiguousarray(fixed, dtype=np.int32).reshape(-1, 3)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    out = SaddleResult()
    lib().orc_saddle_search(C.byref(pot), C.byref(prm), C.c_long(len(Z)), _dp(R), _ip(Z), _ip(fixed), _dp(box),
                            _dp(mode), C.c_double(reactant_energy), C.byref(out))
    return out, R, mode


def relax(pot, prm, R, Z, fixed, box):
    R = np.array(R, dtype=np.float64).reshape(-1, 3).copy()
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    fixed = np.ascontiguousarray(fixed, dtype=np.int32).reshape(-1, 3)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    out = RelaxResult()
    lib().orc_relax(C.byref(pot), C.byref(prm), C.c_long(len(Z)), _dp(R), _ip(Z), _ip(fixed), _dp(box),
                    C.byref(out))
    return out, R


def symmetric_eig(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    ev = np.zeros(n)
    vec = np.zeros((n, n))
    lib().orc_symmetric_eig(n, _dp(a), _dp(ev), _dp(vec))
    return ev, vec


# ---- compiled reference pieces (oracle/_ref) -------------------------------------------------

def ref_eam_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libeam_ref.so"))


def ref_vesin_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libvesin_ref.so"))


_ref_eam = None


def ref_eam_force(R, Z, box):
    """The reference's own EAM.cpp (compiled from /root/reference by oracle/Makefile)."""
    global _ref_eam
    if _ref_eam is None:
        _ref_eam = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libeam_ref.so"))
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    F = np.zeros_like(R)
    U = C.c_double()
    rc = _ref_eam.ref_eam_force(C.c_long(len(Z)), _dp(R), _ip(Z), _dp(F), C.byref(U), _dp(box))
    if rc != 0:
        raise RuntimeError(f"ref_eam_force rc={rc}")
    return U.value, F


class _VesinOptions(C.Structure):
    # client/thirdparty/vesin/vesin.h: struct VesinOptions
    _fields_ = [("cutoff", C.c_double), ("full", C.c_bool), ("sorted", C.c_bool), ("algorithm", C.c_int),
                ("skin", C.c_double), ("n_threads", C.c_int32), ("return_shifts", C.c_bool),
                ("return_distances", C.c_bool), ("return_vectors", C.c_bool)]


class _VesinDevice(C.Structure):
    _fields_ = [("type", C.c_int), ("device_id", C.c_int)]


class _VesinNeighborList(C.Structure):
    _fields_ = [("length", C.c_size_t), ("device", _VesinDevice), ("pairs", C.POINTER(C.c_size_t)),
                ("shifts", C.POINTER(C.c_int32)), ("distances", C.POINTER(C.c_double)),
                ("vectors", C.POINTER(C.c_double)), ("opaque", C.c_void_p)]


_ref_vesin = None


def ref_vesin_neighbors(R, box, cutoff, full=True, periodic=(1, 1, 1), algorithm=0, vectors=False):
    """The reference's own vendored vesin (compiled from /root/reference by oracle/Makefile):
    -> (pairs[P,2] int64, shifts[P,3] int32[, distances, vectors])."""
    global _ref_vesin
    if _ref_vesin is None:
        _ref_vesin = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libvesin_ref.so"))
        _ref_vesin.vesin_neighbors.restype = C.c_int
        _ref_vesin.vesin_neighbors.argtypes = [C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double),
                                               C.POINTER(C.c_bool), _VesinDevice, _VesinOptions,
                                               C.POINTER(_VesinNeighborList), C.POINTER(C.c_char_p)]
        _ref_vesin.vesin_free.argtypes = [C.POINTER(_VesinNeighborList)]
        _ref_vesin.vesin_free.restype = None
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    per = (C.c_bool * 3)(*[bool(x) for x in periodic])
    opt = _VesinOptions(float(cutoff), bool(full), False, int(algorithm), 0.0, 1, True, bool(vectors), bool(vectors))
    nl = _VesinNeighborList()
    nl.device = _VesinDevice(0, 0)
    err = C.c_char_p()
    rc = _ref_vesin.vesin_neighbors(_dp(R), len(R), _dp(box), per, _VesinDevice(1, 0), opt, C.byref(nl), C.byref(err))
    if rc != 0:
        raise RuntimeError(f"vesin_neighbors failed: {err.value}")
    n = nl.length
    pairs = np.ctypeslib.as_array(nl.pairs, shape=(n, 2)).astype(np.int64) if n else np.zeros((0, 2), np.int64)
    shifts = np.ctypeslib.as_array(nl.shifts, shape=(n, 3)).copy() if n else np.zeros((0, 3), np.int32)
    out = [pairs, shifts]
    if vectors:
        out.append(np.ctypeslib.as_array(nl.distances, shape=(n,)).copy() if n else np.zeros(0))
        out.append(np.ctypeslib.as_array(nl.vectors, shape=(n, 3)).copy() if n else np.zeros((0, 3)))
    _ref_vesin.vesin_free(C.byref(nl))
    return tuple(out)


def load_fixture(name):
    with open(os.path.join(GOLDEN_DIR, name)) as fh:
        d = json.load(fh)
    out = dict(d)
    out["positions"] = np.array(d["positions"], dtype=np.float64)
    out["atomic_nrs"] = np.array(d["atomic_nrs"], dtype=np.int32)
    out["cell"] = np.array(d["cell"], dtype=np.float64)
    out["fixed"] = np.array(d["fixed"], dtype=np.int32)
    for k in ("displacement_positions", "direction", "product_positions"):
        if k in d:
            out[k] = np.array(d[k], dtype=np.float64)
    return out


def goldens():
    with open(os.path.join(GOLDEN_DIR, "goldens.json")) as fh:
        return json.load(fh)
