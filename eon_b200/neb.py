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
