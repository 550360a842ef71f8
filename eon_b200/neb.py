"""Batched nudged-elastic-band forces.

``NudgedElasticBand::updateForces`` (client/NudgedElasticBand.cpp:469-646) evaluates the movable images of ONE band
with ``Potential::forceBatch`` and then loops over them on the host for tangent, springs and projection.  Here many
bands go through the same two steps at once: `band_forces` evaluates every movable image of every band in one
``forceBatch`` call (one fused kernel for EAM) and hands positions, masked forces and energies to the
`k_neb_band` kernel (``eonb200_neb_forces``, one CTA per image) for the band algebra.  `convergence_force` is
``NudgedElasticBand::convergenceForce`` (:433-466) per band.  A band's images stay together on one GPU; bands shard
across GPUs like walkers (`eon_b200.sharding`).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _capi
from ._capi import NEB_PROJECTION_DNEB, NEB_PROJECTION_ELASTIC_BAND, NEB_PROJECTION_NEB  # noqa: F401


def neb_params(lib, **kw) -> _capi.NebParams:
    """``EonB200NebParams`` with the reference's defaults (include/eon/Parameters.h:478-489), fields overridden by kw."""
    p = _capi.NebParams()
    lib.eonb200_neb_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown NEB parameter %r" % k)
        setattr(p, k, v)
    return p


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def neb_forces(pot, positions, forces, energies, box, n_free_atoms: Optional[int] = None, prm=None, **kw):
    """Band algebra on the device.  positions / forces ``[B, M + 2, 3 N]`` (endpoints included, forces as
    ``Matter::getForces`` returns them), energies ``[B, M + 2]``.  Returns ``(projected [B, M, 3 N],
    tangents [B, M, 3 N], climbing [B])``."""
    R = np.ascontiguousarray(positions, dtype=np.float64)
    if R.ndim == 2:
        R = R[None]
    B, M2, D = R.shape
    F = np.ascontiguousarray(np.asarray(forces, dtype=np.float64).reshape(B, M2, D))
    E = np.ascontiguousarray(np.asarray(energies, dtype=np.float64).reshape(B, M2))
    M, N = M2 - 2, D // 3
    if M < 1:
        raise ValueError("a band needs at least one movable image")
    if prm is None:
        prm = neb_params(pot._lib, **kw)
    cell = None if box is None else np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    P = np.empty((B, M, D))
    T = np.empty((B, M, D))
    ci = np.zeros(B, dtype=np.int32)
    nfree = N if n_free_atoms is None else int(n_free_atoms)
    pot._check(pot._lib.eonb200_neb_forces(pot._h, B, M, N, nfree, _dp(R), _dp(F), _dp(E),
                                           None if cell is None else _dp(cell), C.byref(prm), _dp(P), _dp(T),
                                           ci.ctypes.data_as(C.POINTER(C.c_int))))
    return P, T, ci


def band_forces(pot, positions, atomic_nrs, box, fixed=None, endpoint_energies=None, remove_net_force: bool = True,
                prm=None, **kw):
    """One NEB force evaluation of B bands: all B x M movable images in one ``forceBatch`` call, then `neb_forces`.
    `endpoint_energies` ``[B, 2]`` are evaluated (one more batch of 2 B systems) when not given.  Returns a dict with
    projected forces, tangents, climbing image, image energies ``[B, M + 2]`` and true (masked) forces."""
    R = np.ascontiguousarray(positions, dtype=np.float64)
    if R.ndim == 2:
        R = R[None]
    B, M2, D = R.shape
    M, N = M2 - 2, D // 3
    E = np.empty((B, M2))
    F = np.zeros((B, M2, D))
    if endpoint_energies is None:
        ends = np.ascontiguousarray(R[:, [0, M + 1]].reshape(2 * B, D))
        _, Ue, _ = pot.forceBatch(2 * B, N, ends, atomic_nrs, box)
        endpoint_energies = np.asarray(Ue).reshape(B, 2)
    E[:, [0, M + 1]] = np.asarray(endpoint_energies, dtype=float).reshape(B, 2)
    mov = np.ascontiguousarray(R[:, 1:M + 1].reshape(B * M, D))
    Fm, Um, _ = pot.forceBatch(B * M, N, mov, atomic_nrs, box)
    Fm = np.asarray(Fm, dtype=np.float64).reshape(B * M, N, 3).copy()
    any_fixed = fixed is not None and bool(np.asarray(fixed).any())
    if any_fixed:  # Matter::getForces (client/Matter.cpp:324-333,531-536)
        Fm *= (1 - np.asarray(fixed).reshape(1, N, 3))
    elif remove_net_force:
        Fm -= Fm.mean(axis=1, keepdims=True)
    F[:, 1:M + 1] = Fm.reshape(B, M, D)
    E[:, 1:M + 1] = np.asarray(Um).reshape(B, M)
    # Matter::numberOfFreeAtoms counts an atom as fixed only when all three axes are (client/Matter.cpp getFixed)
    n_free = N - int(np.asarray(fixed).reshape(N, 3).all(axis=1).sum()) if fixed is not None else N
    P, T, ci = neb_forces(pot, R, F, E, box, n_free, prm, **kw)
    return {"projected": P, "tangents": T, "climbing": ci, "energies": E, "forces": F[:, 1:M + 1]}


def band_forces_resident(pot, positions, atomic_nrs, box, fixed=None, remove_net_force: bool = True, prm=None,
                         want_tangents: bool = False, **kw):
    """`band_forces` as ONE library call (``eonb200_neb_band_forces``): positions up once, all images of all bands in
    one batched potential launch (endpoints included, so their energies need no separate call), masking and band
    algebra on the device, only projected forces, energies and climbing images back."""
    R = np.ascontiguousarray(positions, dtype=np.float64)
    if R.ndim == 2:
        R = R[None]
    B, M2, D = R.shape
    M, N = M2 - 2, D // 3
    if M < 1:
        raise ValueError("a band needs at least one movable image")
    if prm is None:
        prm = neb_params(pot._lib, **kw)
    Z = np.ascontiguousarray(atomic_nrs, dtype=np.int32).reshape(N)
    cell = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    fx = None if fixed is None else np.ascontiguousarray(fixed, dtype=np.int32).reshape(D)
    P = np.empty((B, M, D))
    E = np.empty((B, M2))
    T = np.empty((B, M, D)) if want_tangents else None
    ci = np.zeros(B, dtype=np.int32)
    ip = C.POINTER(C.c_int)
    pot._check(pot._lib.eonb200_neb_band_forces(pot._h, B, M, N, _dp(R), Z.ctypes.data_as(ip),
                                                None if fx is None else fx.ctypes.data_as(ip), _dp(cell),
                                                1 if remove_net_force else 0, C.byref(prm), _dp(P), _dp(E),
                                                None if T is None else _dp(T), ci.ctypes.data_as(ip)))
    return {"projected": P, "tangents": T, "climbing": ci, "energies": E}


def convergence_force(projected, climbing=None, metric: str = "norm", fixed=None, ci_converged_only: bool = False):
    """``NudgedElasticBand::convergenceForce`` per band (client/NudgedElasticBand.cpp:433-466)."""
    P = np.asarray(projected, dtype=float)
    B, M, D = P.shape
    N = D // 3
    if metric == "norm":
        per = np.linalg.norm(P, axis=2)
    elif metric == "max_atom":
        a = np.linalg.norm(P.reshape(B, M, N, 3), axis=3)
        if fixed is not None:
            a = a[:, :, ~np.asarray(fixed).reshape(N, 3).all(axis=1)]
        per = a.max(axis=2)
    elif metric == "max_component":
        per = P.max(axis=2)
    else:
        raise ValueError("unknown convergence_metric: %s" % metric)
    out = np.maximum(per.max(axis=1), 0.0)
    if ci_converged_only and climbing is not None:
        for b in range(B):
            if climbing[b] != 0:
                out[b] = max(0.0, per[b, climbing[b] - 1])
    return out


# ---- the band optimiser ------------------------------------------------------------------------------------------
def _max_atom_motion(v: np.ndarray) -> float:
    """``helpers::maxAtomMotionV`` (client/HelperFunctions.cpp): largest per-atom displacement of a 3 n vector."""
    return float(np.sqrt((v.reshape(-1, 3) ** 2).sum(axis=1)).max()) if v.size else 0.0


def _max_motion_applied(v: np.ndarray, max_move: float) -> np.ndarray:
    """``helpers::maxAtomMotionAppliedV``: scale the whole vector so that no atom moves further than max_move."""
    m = _max_atom_motion(v)
    return v * (max_move / m) if m > max_move else v


def _normalized(v: np.ndarray) -> np.ndarray:
    n = float(np.linalg.norm(v))
    return v / n if n > 1e-300 else np.zeros_like(v)


class BandLBFGS:
    """``LBFGS`` (client/LBFGS.cpp:19-186) for the concatenated images of one band (NEBObjectiveFunction.cpp:16-97):
    history update with the degenerate-curvature reset, H0 from the curvature along the last step (auto_scale) or a
    finite-difference probe on the first step, two-loop recursion, distance / angle resets, max_move scaling."""
    LBFGS_EPS = 1e-30

    def __init__(self, memory=20, inverse_curvature=0.01, auto_scale=True, angle_reset=True, distance_reset=True,
                 finite_difference=0.01):
        self.m, self.h0, self.auto, self.areset, self.dreset, self.fd = (memory, inverse_curvature, auto_scale,
                                                                        angle_reset, distance_reset, finite_difference)
        self.s, self.y, self.rho = [], [], []
        self.r_prev = self.f_prev = None
        self.iteration = 0

    def reset(self):
        self.s, self.y, self.rho = [], [], []

    def wants_probe(self) -> bool:
        return self.iteration == 0 and self.auto

    def probe_positions(self, r, f):
        return r + self.fd * _normalized(f)

    def step(self, r, f, diff, max_move, f_probe=None):
        """One ``LBFGS::step``: returns the displacement to add to r.  `diff(a, b)` is the objective's pbc difference;
        `f_probe` the force at `probe_positions` when `wants_probe()`."""
        if self.iteration > 0:  # LBFGS::update
            s0 = diff(r, self.r_prev)
            y0 = self.f_prev - f
            sy = float(s0 @ y0)
            if abs(sy) < self.LBFGS_EPS:
                self.reset()
            else:
                self.rho.append(1.0 / sy)
                self.s.append(s0)
                self.y.append(y0)
                if len(self.s) > self.m:
                    self.s.pop(0), self.y.pop(0), self.rho.pop(0)
        dr = self._get_step(r, f, diff, max_move, f_probe)
        self.r_prev, self.f_prev = r.copy(), f.copy()
        self.iteration += 1
        return dr

    def _get_step(self, r, f, diff, max_move, f_probe):
        H0 = self.h0
        if self.iteration > 0:
            dr = diff(r, self.r_prev)
            df = self.f_prev - f
            den = float(dr @ df)
            Cc = float(df @ df) / den if abs(den) >= 1e-300 else -1.0
            if Cc < 0:
                self.reset()
                return _max_motion_applied(1000.0 * f, max_move)
            if self.auto:
                H0 = 1.0 / Cc if abs(Cc) >= 1e-300 else -1.0
        if self.iteration == 0 and self.auto:
            fh = _normalized(f)
            dg = -f_probe + f  # getGradient(true) + f, the gradient being minus the projected force
            Cc = float(dg @ fh) / self.fd
            H0 = 1.0 / Cc if abs(Cc) >= 1e-300 else -1.0
            if H0 < 0:
                self.reset()
                return _max_motion_applied(1000.0 * f, max_move)
        q = -f.copy()
        a = [0.0] * len(self.s)
        for i in range(len(self.s) - 1, -1, -1):
            a[i] = self.rho[i] * float(self.s[i] @ q)
            q -= a[i] * self.y[i]
        z = H0 * q
        for i in range(len(self.s)):
            b = self.rho[i] * float(self.y[i] @ z)
            z += self.s[i] * (a[i] - b)
        d = -z
        if _max_atom_motion(d) >= max_move and self.dreset:
            self.reset()
            return _max_motion_applied(H0 * f, max_move)
        vd = min(1.0, max(-1.0, float(_normalized(d) @ _normalized(f))))
        if np.degrees(np.arccos(vd)) > 90.0 and self.areset:
            self.reset()
            return _max_motion_applied(H0 * f, max_move)
        return _max_motion_applied(d, max_move)


def spline_extrema(energies, parallel_forces, distances):
    """``findSplineExtrema`` (client/NEBSplineExtrema.cpp:73-165): cubic through (U_i, F_i) per interval,
    ``parallel_forces[i]`` the true force along the tangent at image i (endpoints: along the first / last segment),
    ``distances[i]`` the length of interval i.  Returns (positions, energies, curvatures)."""
    E = np.asarray(energies, dtype=float)
    Fp = np.asarray(parallel_forces, dtype=float)
    D = np.asarray(distances, dtype=float)
    pos, en, cu = [], [], []
    for i in range(len(E) - 1):
        F1, F2, U1, U2 = Fp[i] * D[i], Fp[i + 1] * D[i], E[i], E[i + 1]
        a, b, c, d = U1, -F1, 3.0 * (U2 - U1) + 2.0 * F1 + F2, -2.0 * (U2 - U1) - (F1 + F2)
        disc = c * c - 3.0 * b * d
        if disc < 0:
            continue
        roots = []
        f = -1.0
        if d == 0 and c != 0:
            f = -b / (2.0 * c)
        elif d != 0:
            f = -(c + np.sqrt(disc)) / (3.0 * d)
        roots.append(f)
        if d != 0:
            f = -(c - np.sqrt(disc)) / (3.0 * d)
            roots.append(f)  # (the reference re-tests the first root when d == 0; it cannot be in [0, 1] twice)
        for f in roots:
            if 0.0 <= f <= 1.0:
                pos.append(i + f)
                en.append(((d * f + c) * f + b) * f + a)
                cu.append(6.0 * d * f + 2.0 * c)
    return np.array(pos), np.array(en), np.array(cu)


def run_bands(pot, reactants, products, atomic_nrs, box, fixed=None, images: int = 5, spring: float = 5.0,
              max_iterations: int = 1000, converged_force: float = 0.01, max_move: float = 0.2,
              climbing_image: bool = True, ci_converged_only: bool = True, metric: str = "norm", lbfgs=None,
              finite_difference: float = 0.01, **kw):
    """``NudgedElasticBand::compute`` (client/NudgedElasticBand.cpp:204-430) for B bands at once: linear initial paths
    (pbc interpolation), LBFGS over the concatenated images of each band, every force evaluation of ALL bands one
    ``eonb200_neb_band_forces`` call (potential + masking + band algebra on the device).  The climbing image is on
    from the first iteration (trigger_force = infinity, Parameters.h:503) and, with ``ci_converged_only``, convergence
    looks at the climbing image alone (:433-466).  Returns a dict: status [B] (0 converged, 1 max iterations),
    iterations, force_calls (per band, movable images only), positions [B, M + 2, N, 3], energies [B, M + 2],
    projected, climbing, extrema (per band: positions, energies, curvatures)."""
    R0 = np.asarray(reactants, dtype=float)
    R1 = np.asarray(products, dtype=float)
    Z = np.asarray(atomic_nrs)
    N = Z.size
    R0 = R0.reshape(-1, N, 3)
    R1 = R1.reshape(-1, N, 3)
    B, M = R0.shape[0], int(images)
    cell = np.asarray(box, dtype=float).reshape(3, 3)
    inv = np.linalg.inv(cell)

    def pbc(d):  # include/eon/Matter.h:42-50
        fr = d @ inv
        fr -= np.floor(fr + 0.5)
        return fr @ cell

    lam = np.linspace(0.0, 1.0, M + 2).reshape(1, -1, 1, 1)
    path = R0[:, None] + lam * pbc(R1 - R0)[:, None]  # [B, M + 2, N, 3]
    path = path.reshape(B, M + 2, 3 * N)
    opts = dict(memory=20, inverse_curvature=0.01, auto_scale=True, angle_reset=True, distance_reset=True,
                finite_difference=finite_difference)
    opts.update(lbfgs or {})
    opt = [BandLBFGS(**opts) for _ in range(B)]
    seg = 3 * N

    def evaluate(p, ci):
        return band_forces_resident(pot, p, Z, cell.reshape(9), fixed, spring_constant=spring,
                                    climbing_image=1 if ci else 0, want_tangents=True, **kw)

    def diff(a, b):  # NEBObjectiveFunction::difference: pbc per image
        return pbc((a - b).reshape(-1, 3)).reshape(-1)

    def conv(out, b):
        return float(convergence_force(out["projected"][b:b + 1], out["climbing"][b:b + 1], metric, fixed,
                                       ci_converged_only and climbing_image)[0])

    out = evaluate(path, climbing_image)
    status = np.full(B, -1, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int64)
    calls = np.full(B, M, dtype=np.int64)
    active = np.ones(B, dtype=bool)
    for b in range(B):
        if conv(out, b) < converged_force:
            status[b], active[b] = 0, False
    iteration = 1
    while active.any():
        if iteration >= max_iterations:
            status[active] = 1
            break
        x = path[:, 1:M + 1].reshape(B, M * seg).copy()
        f = out["projected"].reshape(B, M * seg)
        probe = None
        if any(opt[b].wants_probe() for b in range(B) if active[b]):  # first step: finite-difference curvature
            pp = path.copy()
            for b in range(B):
                if active[b] and opt[b].wants_probe():
                    pp[b, 1:M + 1] = opt[b].probe_positions(x[b], f[b]).reshape(M, seg)
            probe = evaluate(pp, climbing_image)["projected"].reshape(B, M * seg)
            calls[active] += M
        for b in range(B):
            if not active[b]:
                continue
            dr = opt[b].step(x[b], f[b], diff, max_move, None if probe is None else probe[b])
            path[b, 1:M + 1] = (x[b] + dr).reshape(M, seg)
        out = evaluate(path, climbing_image)
        calls[active] += M
        iters[active] = iteration
        iteration += 1
        for b in range(B):
            if active[b] and conv(out, b) < converged_force:
                status[b], active[b] = 0, False
    # extrema of the energy along each band (NEBSplineExtrema.cpp)
    extrema = []
    for b in range(B):
        P = path[b].reshape(M + 2, N, 3)
        T = out["tangents"][b].reshape(M, N, 3)
        Ftrue = _true_forces(pot, path[b], Z, cell, fixed)
        dist = np.array([np.linalg.norm(pbc(P[i + 1] - P[i])) for i in range(M + 1)])
        tan = [pbc(P[1] - P[0])] + [T[i] for i in range(M)] + [pbc(P[M + 1] - P[M])]
        fpar = np.array([float((Ftrue[i] * (tan[i] / max(np.linalg.norm(tan[i]), 1e-300))).sum()) for i in range(M + 2)])
        extrema.append(spline_extrema(out["energies"][b], fpar, dist))
    return {"status": status, "iterations": iters, "force_calls": calls, "positions": path.reshape(B, M + 2, N, 3),
            "energies": out["energies"], "projected": out["projected"], "climbing": out["climbing"],
            "extrema": extrema}


def _true_forces(pot, band, atomic_nrs, cell, fixed):
    """``Matter::getForces`` of every image of one band ([M + 2, N, 3]): potential, mean force removed when nothing is
    fixed, fixed axes zeroed."""
    M2 = band.shape[0]
    N = len(atomic_nrs)
    F, _, _ = pot.forceBatch(M2, N, np.ascontiguousarray(band.reshape(M2, 3 * N)), atomic_nrs, cell.reshape(9))
    F = np.asarray(F, dtype=float).reshape(M2, N, 3).copy()
    if fixed is not None and np.asarray(fixed).any():
        F *= (1 - np.asarray(fixed).reshape(1, N, 3))
    else:
        F -= F.mean(axis=1, keepdims=True)
    return F
