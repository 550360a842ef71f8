"""Host-side mirror of eOn's Potential plugin interface over the C ABI.

`B200Potential` has the methods and argument meaning of ``eonc::Potential``
(reference: include/eon/Potential.h:23-123): ``force``, ``forceBatch``, the capability queries a
caller checks before batching (client/ImprovedDimer.cpp:160, client/NudgedElasticBand.cpp:475) and
``forceCallCounter``.  Errors surface like the reference's: an unknown EAM element raises
``PotentialError(14324)`` (client/potentials/EAM/EAM.cpp:116-123), any other failure a
``PotentialError`` carrying the library's message (the C-ABI precedent converts rc != 0 to
``std::runtime_error``, client/potentials/Metatomic/MetatomicDynPot.cpp:59-69).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import POT_EAM, POT_LJ, POT_LJCLUSTER, POT_MORSE_PT

# "eam" / "eam_b200": the in-tree EAM class (client/potentials/EAM/EAM.cpp).  The reference's own "eam_al" is rgpot's
# Fortran potential, a DIFFERENT model (golden -5.217864 eV against -5.085058 eV here): the
# spelling is accepted as a deprecated alias of the in-tree class for round-1 scripts, never as that potential.
_KINDS = {"lj": POT_LJ, "lj_cluster": POT_LJCLUSTER, "ljcluster": POT_LJCLUSTER, "morse_pt": POT_MORSE_PT,
          "eam": POT_EAM, "eam_b200": POT_EAM, "eam_al": POT_EAM}


class PotentialError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _row_ptrs(a: np.ndarray, ctype, n: int = 0):
    """`ctype **` over the rows of a C-contiguous 2-D array, without a Python loop over the rows.  A one-row array
    with n > 1 yields n copies of that row's address (a value shared by all systems)."""
    if n > 1 and a.shape[0] == 1:
        addr = np.full(n, a.ctypes.data, dtype=np.uint64)
    else:
        addr = a.ctypes.data + np.arange(a.shape[0], dtype=np.uint64) * np.uint64(a.strides[0])
    addr = np.ascontiguousarray(addr, dtype=np.uint64)
    p = addr.ctypes.data_as(C.POINTER(C.POINTER(ctype)))
    p._keepalive = (addr, a)
    return p


def eam_row(Z, Dm, alphaM, Rm, beta1, beta2, r_cut, func_coeff) -> _capi.EamRow:
    """One element row in the layout of include/eon/potentials/EAM/Parameters.h:15-50."""
    r = _capi.EamRow()
    r.Z, r.Dm, r.alphaM, r.Rm, r.beta1, r.beta2, r.r_cut = int(Z), Dm, alphaM, Rm, beta1, beta2, r_cut
    for k, c in enumerate(func_coeff):
        r.func_coeff[k] = float(c)
    return r


class B200Potential:
    """``eonc::Potential`` for LJ / LJCluster / Morse_Pt / EAM evaluated on a B200.

    Parameters mirror `EonB200Config` (include/eonb200.h); ``kind`` also accepts the reference's
    ``[Potential] potential`` strings (lj, lj_cluster, morse_pt, eam_al).
    """

    def __init__(self, kind, device: int = 0, cutoff: float = 0.0, skin: float = 1.0, conservative: bool = False,
                 eam_rows: Optional[Sequence[_capi.EamRow]] = None, threads_per_walker: int = 0, lib=None, **params):
        if lib is None:
            from . import _lib
            lib = _lib.load()
        self._lib = lib
        if isinstance(kind, str):
            kind = _KINDS[kind.lower()]
        cfg = _capi.Config()
        lib.eonb200_config_default(C.byref(cfg), int(kind))
        cfg.device = int(device)
        if cutoff > 0:
            cfg.cutoff = float(cutoff)
        cfg.skin = float(skin)
        cfg.eam_conservative = 1 if conservative else 0
        cfg.threads_per_walker = int(threads_per_walker)
        for k, v in params.items():
            if not hasattr(cfg, k):
                raise AttributeError(f"unknown potential parameter {k}")
            setattr(cfg, k, v)
        if eam_rows:
            arr = (_capi.EamRow * len(eam_rows))(*eam_rows)
            cfg.n_eam_rows = len(eam_rows)
            cfg.eam_rows = C.cast(arr, C.POINTER(_capi.EamRow))
            self._rows_keepalive = arr
        err = C.create_string_buffer(512)
        self._h = lib.eonb200_create(C.byref(cfg), err, len(err))
        if not self._h:
            raise PotentialError(1, err.value.decode() or "eonb200_create failed")
        self.kind = int(kind)
        self.device = int(device)
        self._pin = bool(cfg.pin_host_buffers)
        self._pin_keepalive = []
        self._ptr_cache = {}
        self._walkers = weakref.WeakSet()  # Walkers created on this handle

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        for wk in list(getattr(self, "_walkers", ())):  # resident walkers go first: they belong to this handle
            wk.close()
        if getattr(self, "_h", None):
            self._lib.eonb200_destroy(self._h)
            self._h = None
            self._pin_keepalive = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise PotentialError(rc, self._lib.eonb200_last_error(self._h).decode())

    # -- capability queries (include/eon/Potential.h:62-103) ---------------------------------
    def supportsBatchEvaluation(self) -> bool:
        return True

    def isThreadSafe(self) -> bool:
        return False

    def isSharedInstanceThreadSafe(self) -> bool:
        return False

    def needsPerImageInstance(self) -> bool:
        return False

    def isSurrogate(self) -> bool:
        return False

    def requiresIsolatedMoleculeLayout(self) -> bool:
        return False

    @property
    def forceCallCounter(self) -> int:
        return int(self._lib.eonb200_force_calls(self._h))

    @property
    def kernelLaunches(self) -> int:
        return int(self._lib.eonb200_kernel_launches(self._h))

    # -- Potential::force (include/eon/Potential.h:55-57) --------------------------------------
    def force(self, nAtoms: int, positions, atomicNrs, box):
        """Energy and forces of one configuration; returns (forces[N,3], energy, variance)."""
        R = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1)
        Z = np.ascontiguousarray(atomicNrs, dtype=np.int32).reshape(-1)
        B = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
        if R.size != 3 * nAtoms or Z.size != nAtoms:
            raise ValueError("positions / atomicNrs do not match nAtoms")
        F = np.empty(3 * nAtoms, dtype=np.float64)
        U = C.c_double()
        var = C.c_double()
        self._check(self._lib.eonb200_force(self._h, nAtoms, _dp(R), _ip(Z), _dp(F), C.byref(U), C.byref(var), _dp(B)))
        return F.reshape(-1, 3), U.value, var.value

    def get_ef(self, pos, atmnrs, box):
        """``Potential::get_ef`` (client/Potential.cpp:98-115): (energy, forces)."""
        pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
        F, U, _ = self.force(pos.shape[0], pos, atmnrs, box)
        return U, F

    # -- Potential::forceBatch (include/eon/Potential.h:108-122) ---------------------------------
    def forceBatch(self, nSystems: int, nAtoms: int, positions, atomicNrs, boxes, out=None):
        """Evaluate nSystems systems of nAtoms atoms in one call.

        positions: [S][3N] (or [S][N][3]); atomicNrs: [S][N] or [N]; boxes: [S][9] or [9].
        Returns (forces[S,N,3], energies[S], variances[S]).  `out` = a C-contiguous float64 array of S * 3N
        elements to receive the forces (the reference's callers own their force buffers too); with
        ``pin_host_buffers=1`` contiguous position / force arrays are page-locked once and copied without staging,
        and this object keeps them alive until it is closed.
        """
        S, N = int(nSystems), int(nAtoms)
        R = np.ascontiguousarray(positions, dtype=np.float64).reshape(S, 3 * N)
        # one topology / one box for all systems: every slot of the pointer array names the same row
        Z = np.ascontiguousarray(atomicNrs, dtype=np.int32).reshape(-1, N)
        B = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 9)
        if Z.shape[0] not in (1, S) or B.shape[0] not in (1, S):
            raise ValueError("atomicNrs / boxes must describe one system or nSystems systems")
        if out is not None:
            if out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"] or out.size != S * 3 * N:
                raise ValueError("out must be a C-contiguous float64 array of nSystems * 3 * nAtoms elements")
            F = out.reshape(S, 3 * N)
        else:
            F = np.empty((S, 3 * N), dtype=np.float64)
        U = np.empty(S, dtype=np.float64)
        V = np.empty(S, dtype=np.float64)
        if self._pin:
            for arr in (R, F):
                if not any(arr is k for k in self._pin_keepalive):
                    self._pin_keepalive.append(arr)
        # arrays of per-system pointers (the reference's `const double *const *`), built from base address + row stride
        pR, pZ, pF, pB = (self._row_ptrs_cached(R, C.c_double), self._row_ptrs_cached(Z, C.c_int, S),
                          self._row_ptrs_cached(F, C.c_double), self._row_ptrs_cached(B, C.c_double, S))
        self._check(self._lib.eonb200_force_batch(self._h, S, N, pR, pZ, pF, _dp(U), _dp(V), pB))
        return F.reshape(S, N, 3), U, V

    def _row_ptrs_cached(self, a: np.ndarray, ctype, n: int = 0):
        """_row_ptrs, remembered per (address, rows, pitch): a caller that hands over the same blocks every step (an
        optimiser stepping its own position / force arrays) does not pay for 4 x nSystems pointers each time."""
        key = (a.__array_interface__["data"][0], a.shape[0], a.strides[0], n, ctype)
        p = self._ptr_cache.get(key)
        if p is None:
            if len(self._ptr_cache) >= 64:
                self._ptr_cache.clear()
            p = _row_ptrs(a, ctype, n)
            p._keepalive = (p._keepalive[0],)  # the pointer table only; the array itself belongs to the caller
            self._ptr_cache[key] = p
        return p

    def forceBatchDevice(self, nSystems: int, nAtoms: int, d_positions: int, d_atomicNrs: int, d_forces: int,
                         d_energies: int, d_boxes: int, sync: bool = True):
        """Same evaluation on arrays already resident in HBM (raw device addresses, e.g. ``tensor.data_ptr()``)."""
        self._check(self._lib.eonb200_force_batch_device(self._h, nSystems, nAtoms, d_positions, d_atomicNrs, d_forces,
                                                         d_energies, d_boxes, 1 if sync else 0))

    # -- neighbour lists (vesin_neighbors, client/thirdparty/vesin/vesin.h:190-199) ---------------
    def neighbors(self, positions, box, cutoff: float, full: bool = True, periodic=(1, 1, 1), return_vectors=False):
        """Device-built neighbour list of one system -> (pairs[P,2] int64, shifts[P,3] int32[, dist, vec])."""
        R = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        out = self.neighborsBatch(R[None], np.asarray(box, dtype=np.float64).reshape(1, 9), cutoff, full, periodic,
                                  return_vectors)
        return tuple(o[0] if isinstance(o, list) else o for o in out)

    def neighborsBatch(self, positions, boxes, cutoff: float, full: bool = True, periodic=(1, 1, 1),
                       return_vectors=False):
        R = np.ascontiguousarray(positions, dtype=np.float64)
        S, N = R.shape[0], R.shape[1]
        R = R.reshape(S, 3 * N)
        B = np.ascontiguousarray(np.broadcast_to(np.asarray(boxes, dtype=np.float64).reshape(-1, 9), (S, 9)))
        per = (C.c_int * 3)(*[int(p) for p in periodic])
        cap = 0
        offs = np.zeros(S + 1, dtype=np.int64)
        pairs = np.zeros((1, 2), dtype=np.int64)
        shifts = np.zeros((1, 3), dtype=np.int32)
        dist = vec = None
        n = C.c_long(0)
        for _ in range(2):
            self._check(self._lib.eonb200_neighbors(
                self._h, S, N, _dp(R), _dp(B), per, float(cutoff), 1 if full else 0, cap,
                offs.ctypes.data_as(C.POINTER(C.c_long)), pairs.ctypes.data_as(C.POINTER(C.c_int64)),
                shifts.ctypes.data_as(C.POINTER(C.c_int32)), _dp(dist) if dist is not None else None,
                _dp(vec) if vec is not None else None, C.byref(n)))
            if n.value <= cap:
                break
            cap = int(n.value)
            pairs = np.zeros((cap, 2), dtype=np.int64)
            shifts = np.zeros((cap, 3), dtype=np.int32)
            if return_vectors:
                dist = np.zeros(cap)
                vec = np.zeros((cap, 3))
        P = [pairs[offs[s]:offs[s + 1]] for s in range(S)]
        Sh = [shifts[offs[s]:offs[s + 1]] for s in range(S)]
        if return_vectors:
            if dist is None:
                dist, vec = np.zeros(0), np.zeros((0, 3))
            return P, Sh, [dist[offs[s]:offs[s + 1]] for s in range(S)], [vec[offs[s]:offs[s + 1]] for s in range(S)]
        return P, Sh

    # -- measurement helpers --------------------------------------------------------------------------
    def nl_stats(self) -> dict:
        out = (C.c_longlong * 5)()
        self._lib.eonb200_nl_stats(self._h, out)
        return {"rebuilt": out[0], "reused": out[1], "row_capacity": out[2], "max_row_fill": out[3],
                "ordered_pairs": out[4]}

    def last_force_ms(self):
        out = (C.c_double * 2)()
        self._lib.eonb200_last_force_ms(self._h, out)
        return out[0], out[1]

    def l2_flush(self, nbytes: int = 512 << 20):
        """Queue an L2-evicting memset on the potential's stream (returns immediately)."""
        self._check(self._lib.eonb200_l2_flush(self._h, int(nbytes)))

    def phase_clocks(self, enable: bool = True):
        """Per-CTA phase clocks of the last fused EAM launch -> int64 array [n_ctas, 16] (see eonb200_phase_clocks)."""
        n = C.c_long(0)
        self._check(self._lib.eonb200_phase_clocks(self._h, 1 if enable else 0, None, 0, C.byref(n)))
        out = np.zeros((max(int(n.value), 0), 16), dtype=np.int64)
        if n.value > 0:
            self._check(self._lib.eonb200_phase_clocks(self._h, 1 if enable else 0,
                                                       out.ctypes.data_as(C.POINTER(C.c_longlong)), n.value, C.byref(n)))
        return out

    def measure_fp64_peak(self, repeats: int = 5) -> float:
        return float(self._lib.eonb200_measure_fp64_peak(self._h, repeats))
