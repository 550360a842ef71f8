"""Device-resident batches of walkers: the batched counterpart of eOn's job drivers.

`Walkers` keeps W copies of one topology (atomic numbers, fixed mask, box) in HBM and runs, for
all of them in one launch with one CTA per walker,

* ``matter_forces``   -- ``Matter::getPotentialEnergy / getForces`` (client/Matter.cpp:324-333,498-538)
* ``saddle_search``   -- ``MinModeSaddleSearch::run`` as driven by ``SaddleSearchJob::run``
                         (client/SaddleSearchJob.cpp:27-107, client/MinModeSaddleSearch.cpp:211-475)
* ``relax``           -- ``helpers::relaxMatter`` (client/HelperFunctions.cpp:314-410)
* ``process_search``  -- the saddle search followed by the two displaced minimisations of
                         ``ProcessSearchJob::run`` (client/ProcessSearchJob.cpp:165-289)

and harvests results to the host (the path's only "collective").
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _capi
from .potential import B200Potential, PotentialError, _dp, _ip

STATUS_GOOD = 0
STATUS_BAD_HIGH_ENERGY = 3
STATUS_BAD_MAX_ITERATIONS = 5
STATUS_POTENTIAL_FAILED = 11
STATUS_NONLOCAL_ABORT = 13
STATUS_BAD_NO_NEGATIVE_MODE_AT_SADDLE = 16
STATUS_ZEROMODE_ABORT = 18
STATUS_OPTIMIZER_ERROR = 19


def search_params(lib=None, **kw) -> _capi.SearchParams:
    """Defaults of the reference's Parameters for this path (include/eon/Parameters.h), overridden by kw."""
    if lib is None:
        from . import _lib
        lib = _lib.load()
    p = _capi.SearchParams()
    lib.eonb200_search_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(f"unknown search parameter {k}")
        setattr(p, k, v)
    return p


@dataclass
class Harvest:
    status: np.ndarray
    iterations: np.ndarray
    force_calls: np.ndarray
    energies: np.ndarray
    eigenvalues: np.ndarray
    convergence: np.ndarray
    positions: Optional[np.ndarray] = None
    modes: Optional[np.ndarray] = None


class Walkers:
    def __init__(self, pot: B200Potential, atomicNrs, fixed, box, positions, modes=None):
        self._pot = pot
        self._lib = pot._lib
        Z = np.ascontiguousarray(atomicNrs, dtype=np.int32).reshape(-1)
        N = Z.size
        R = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3 * N)
        W = R.shape[0]
        fx = np.ascontiguousarray(np.asarray(fixed).astype(np.int32).reshape(N, 3))
        B = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
        M = None
        if modes is not None:
            M = np.ascontiguousarray(modes, dtype=np.float64).reshape(W, 3 * N)
        err = C.create_string_buffer(512)
        self._h = self._lib.eonb200_walkers_create(pot._h, W, N, _ip(Z), _ip(fx), _dp(B), _dp(R),
                                                   _dp(M) if M is not None else None, err, len(err))
        if not self._h:
            raise PotentialError(1, err.value.decode() or "eonb200_walkers_create failed")
        self.W, self.N = W, N
        pot._walkers.add(self)  # the potential closes its walkers before it goes (they hold device memory of its context)

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self._pot, "_h", None):  # (a handle whose potential is gone is not touched)
                self._lib.eonb200_walkers_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        self._pot._check(rc)

    def set(self, positions=None, modes=None):
        R = M = None
        if positions is not None:
            R = np.ascontiguousarray(positions, dtype=np.float64).reshape(self.W, 3 * self.N)
        if modes is not None:
            M = np.ascontiguousarray(modes, dtype=np.float64).reshape(self.W, 3 * self.N)
        self._check(self._lib.eonb200_walkers_set(self._h, _dp(R) if R is not None else None,
                                                  _dp(M) if M is not None else None))

    def matter_forces(self, remove_net_force: bool = True):
        E = np.empty(self.W)
        F = np.empty((self.W, 3 * self.N))
        self._check(self._lib.eonb200_walkers_matter_forces(self._h, 1 if remove_net_force else 0, _dp(E), _dp(F)))
        return E, F.reshape(self.W, self.N, 3)

    def saddle_search(self, prm: _capi.SearchParams, reactant_energies=None):
        re = None
        if reactant_energies is not None:
            re = np.ascontiguousarray(np.broadcast_to(np.asarray(reactant_energies, dtype=np.float64), (self.W,)))
        self._check(self._lib.eonb200_walkers_saddle_search(self._h, C.byref(prm), _dp(re) if re is not None else None))

    def relax(self, prm: _capi.SearchParams):
        self._check(self._lib.eonb200_walkers_relax(self._h, C.byref(prm)))

    def displace_along_mode(self, signed_offset: float):
        self._check(self._lib.eonb200_walkers_displace_along_mode(self._h, float(signed_offset)))

    def harvest(self, geometry: bool = True) -> Harvest:
        W, N = self.W, self.N
        st = np.zeros(W, dtype=np.int32)
        it = np.zeros(W, dtype=np.int64)
        fc = np.zeros(W, dtype=np.int64)
        E, ev, cv = np.zeros(W), np.zeros(W), np.zeros(W)
        R = np.zeros((W, 3 * N)) if geometry else None
        M = np.zeros((W, 3 * N)) if geometry else None
        lp = C.POINTER(C.c_long)
        self._check(self._lib.eonb200_walkers_harvest(
            self._h, _ip(st), it.ctypes.data_as(lp), fc.ctypes.data_as(lp), _dp(E), _dp(ev), _dp(cv),
            _dp(R) if geometry else None, _dp(M) if geometry else None))
        return Harvest(st, it, fc, E, ev, cv, R.reshape(W, N, 3) if geometry else None,
                       M.reshape(W, N, 3) if geometry else None)

    def displace(self, reference_positions, epicentres, radius: float, magnitude: float, seed: int,
                 first_walker_id: int = 0):
        """Displaced start geometries and modes for all W walkers, generated on the device (eon/displace.py:180-263):
        no per-walker Python loop and no upload of W geometries.  `reference_positions`: [N, 3] shared or [W, N, 3];
        `epicentres`: one atom index or [W]."""
        ref = np.ascontiguousarray(reference_positions, dtype=np.float64).reshape(-1, 3 * self.N)
        epi = np.ascontiguousarray(np.broadcast_to(np.asarray(epicentres, dtype=np.int32), (self.W,)))
        self._check(self._lib.eonb200_walkers_displace(self._h, _dp(ref), ref.shape[0], _ip(epi), float(radius),
                                                       float(magnitude), int(seed) & (2 ** 64 - 1), int(first_walker_id)))

    def set_active(self, active=None):
        """Mask for the following launches: walkers with active == 0 are left alone (``None``: all take part)."""
        if active is None:
            self._check(self._lib.eonb200_walkers_set_active(self._h, None))
        else:
            a = np.ascontiguousarray(np.asarray(active).astype(np.int32).reshape(self.W))
            self._check(self._lib.eonb200_walkers_set_active(self._h, _ip(a)))

    # ---- streaming: a queue of searches through the handle's W slots ------------------------------------------
    def queue_search(self, prm: _capi.SearchParams, positions, modes=None, reactant_energies=None, job: int = 1,
                     geometry: bool = True) -> Harvest:
        """Q >= 1 saddle searches (job 1) or relaxations (job 2) through the W resident slots in ONE launch
        (``eonb200_walkers_queue``): every CTA takes the next queue entry from a device counter as soon as its walker is
        done.  Results come back in queue order."""
        R = np.array(positions, dtype=np.float64).reshape(-1, 3 * self.N)  # (copied: the call writes the final geometries)
        Q = R.shape[0]
        M = None if modes is None else np.array(modes, dtype=np.float64).reshape(Q, 3 * self.N)
        E0 = None
        if reactant_energies is not None:
            E0 = np.ascontiguousarray(np.broadcast_to(np.asarray(reactant_energies, dtype=np.float64), (Q,)))
        st = np.zeros(Q, np.int32)
        it, fc = np.zeros(Q, np.int64), np.zeros(Q, np.int64)
        E, ev, cv = np.zeros(Q), np.zeros(Q), np.zeros(Q)
        lp = C.POINTER(C.c_long)
        self._check(self._lib.eonb200_walkers_queue(
            self._h, int(job), C.byref(prm), Q, _dp(R), _dp(M) if M is not None else None,
            _dp(E0) if E0 is not None else None, _ip(st), it.ctypes.data_as(lp), fc.ctypes.data_as(lp), _dp(E), _dp(ev),
            _dp(cv)))
        return Harvest(st, it, fc, E, ev, cv, R.reshape(Q, self.N, 3) if geometry else None,
                       M.reshape(Q, self.N, 3) if (geometry and M is not None) else None)

    def stream_search(self, prm: _capi.SearchParams, positions, modes, reactant_energies=None, groups: int = 32,
                      geometry: bool = True, job: int = 1) -> Harvest:
        """Run Q >= 1 searches (job 1) or relaxations (job 2) through the W slots of this handle, Q arbitrary.

        The slots are split into `groups` slices with a stream each; every slice is refilled from the queue as soon as
        its walkers are done, the way eon/communicator.py:525-559 hands the next job to a free core, so a walker that
        runs into ``max_iterations`` only holds up its own slice.  Results come back in queue order."""
        R = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3 * self.N)
        Q = R.shape[0]
        M = None if modes is None else np.ascontiguousarray(modes, dtype=np.float64).reshape(Q, 3 * self.N)
        E0 = None
        if reactant_energies is not None:
            E0 = np.ascontiguousarray(np.broadcast_to(np.asarray(reactant_energies, dtype=np.float64), (Q,)))
        G = max(1, min(int(groups), self.W, 64))
        self._check(self._lib.eonb200_walkers_groups(self._h, G))
        bounds = [self.W * g // G for g in range(G + 1)]
        out = Harvest(np.zeros(Q, np.int32), np.zeros(Q, np.int64), np.zeros(Q, np.int64), np.zeros(Q), np.zeros(Q),
                      np.zeros(Q), np.zeros((Q, 3 * self.N)) if geometry else None,
                      np.zeros((Q, 3 * self.N)) if geometry else None)
        lp = C.POINTER(C.c_long)
        nxt = 0
        inflight = {}  # group -> (queue start, count)

        def submit(g):
            nonlocal nxt
            cap = bounds[g + 1] - bounds[g]
            n = min(cap, Q - nxt)
            if n <= 0:
                return
            q0 = nxt
            nxt += n
            self._check(self._lib.eonb200_walkers_submit(
                self._h, g, job, C.byref(prm), bounds[g], n, _dp(R[q0:q0 + n]),
                _dp(M[q0:q0 + n]) if M is not None else None, _dp(E0[q0:q0 + n]) if E0 is not None else None))
            inflight[g] = (q0, n)

        def harvest(g):
            q0, n = inflight.pop(g)
            it = np.zeros(n, np.int64)
            fc = np.zeros(n, np.int64)
            st = np.zeros(n, np.int32)
            E, ev, cv = np.zeros(n), np.zeros(n), np.zeros(n)
            P = np.zeros((n, 3 * self.N)) if geometry else None
            Mo = np.zeros((n, 3 * self.N)) if geometry else None
            self._check(self._lib.eonb200_walkers_harvest_group(
                self._h, g, _ip(st), it.ctypes.data_as(lp), fc.ctypes.data_as(lp), _dp(E), _dp(ev), _dp(cv),
                _dp(P) if geometry else None, _dp(Mo) if geometry else None))
            sl = slice(q0, q0 + n)
            out.status[sl], out.iterations[sl], out.force_calls[sl] = st, it, fc
            out.energies[sl], out.eigenvalues[sl], out.convergence[sl] = E, ev, cv
            if geometry:
                out.positions[sl], out.modes[sl] = P, Mo

        for g in range(G):
            submit(g)
        while inflight:
            progressed = False
            for g in list(inflight):
                state = self._lib.eonb200_walkers_poll(self._h, g, 0)
                if state < 0:
                    raise PotentialError(100, "walkers_poll failed")
                if state == 1:
                    harvest(g)
                    submit(g)
                    progressed = True
            if not progressed:  # nothing finished: block on the oldest group instead of spinning
                g = next(iter(inflight))
                if self._lib.eonb200_walkers_poll(self._h, g, 1) < 0:
                    raise PotentialError(100, "walkers_poll failed")
        if geometry:
            out.positions = out.positions.reshape(Q, self.N, 3)
            out.modes = out.modes.reshape(Q, self.N, 3)
        return out

    @property
    def last_ms(self) -> float:
        return float(self._lib.eonb200_walkers_last_ms(self._h))

    # ProcessSearchJob::run without file I/O (client/ProcessSearchJob.cpp:165-289): saddle search, then
    # min1 / min2 = saddle -/+ minimization_offset * mode relaxed on the free coordinates.
    def process_search(self, prm: _capi.SearchParams, reactant_energies=None, minimization_offset=None):
        offset = prm.max_move if minimization_offset is None else minimization_offset  # ParametersINI.cpp:1374-1376
        self.saddle_search(prm, reactant_energies)
        saddle = self.harvest()
        out = {"saddle": saddle}
        # ProcessSearchJob.cpp:184-186 returns right after a failed saddle search: those walkers sit out the two
        # minimisations (no force calls, and no relaxation from a failed geometry)
        good = saddle.status == STATUS_GOOD
        self.set_active(good)
        try:
            for name, sign in (("min1", -1.0), ("min2", +1.0)):  # ProcessSearchJob.cpp:190-202: min1 = saddle - offset * mode
                self.set(positions=saddle.positions, modes=saddle.modes)
                if good.any():
                    self.displace_along_mode(sign * offset)
                    self.relax(prm)
                h = self.harvest()
                bad = ~good
                h.status[bad] = 1
                h.iterations[bad] = 0
                h.force_calls[bad] = 0
                h.energies[bad] = np.nan
                out[name] = h
        finally:
            self.set_active(None)
        self.set(positions=saddle.positions, modes=saddle.modes)
        return out
