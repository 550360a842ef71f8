"""CPU oracle of the NEB band forces -- TEST INFRASTRUCTURE ONLY (imported by tests/, never by eon_b200).

numpy restatement, image by image, of what ``NudgedElasticBand::updateForces`` does after the potential call
(client/NudgedElasticBand.cpp:551-646), with the helpers it dispatches to: ``ImprovedTangent`` / ``SimpleTangent``
(client/NEBTangent.cpp:21-69), ``UniformSpring`` / ``WeightedSpring`` and the weighted spring constants
(client/NEBSpringForce.cpp:17-42,113-136), the projections (client/NEBProjection.cpp:17-45), ``computeDNEB``,
``climbingImageForce`` and ``zeroTranslation`` (client/NEBForceProjection.cpp:65-110).

Parity pinning: the reference's NEB cannot be built here (needs eonclient).  Every helper below
is pinned by the hand-computed vectors of the reference's own unit tests, client/unit_tests/NEBSubsystemTest.cpp:47-353
(tangent cases, forcePerp (0, 4, 0), climbingImageForce (-3, 4), computeDNEB switching 2/pi atan(1/4), zeroTranslation,
uniform / weighted spring 8 t) and :376-477 (cubic extrema), reproduced in
tests/test_neb.py::test_reference_neb_subsystem_vectors; `band_forces` is composed of exactly these helpers.  The band
OPTIMISER's known answer (data/reference/neb_lj.dat) is not reproduced: see DESIGN.md section 8-2.
"""
import numpy as np


def pbc(d, cell):
    """include/eon/Matter.h:42-50"""
    if cell is None:
        return d
    f = d @ np.linalg.inv(cell)
    f -= np.floor(f + 0.5)
    return f @ cell


def _normalize(t, d_next):
    n = np.linalg.norm(t)
    if n > 1e-10:
        return t / n
    t = d_next.copy()
    n = np.linalg.norm(t)
    return t / n if n > 1e-10 else t


def tangent(d_next, d_prev, e, e_prev, e_next, old_tangent=False):
    if old_tangent or (e_next > e and e > e_prev):
        t = d_next
    elif e > e_next and e_prev > e:
        t = d_prev
    else:
        dp, dn = e_prev - e, e_next - e
        lo, hi = min(abs(dp), abs(dn)), max(abs(dp), abs(dn))
        t = d_next * lo + d_prev * hi if dp > dn else d_next * hi + d_prev * lo
    return _normalize(t, d_next)


def spring_constants(E, e_max, e_ref, k_min, k_max):
    M = len(E) - 2
    k = np.full(M + 2, k_min)
    rng = e_max - e_ref
    if not rng < 1e-10:
        for idx in range(1, M + 2):
            ei = max(E[idx], E[idx - 1])
            if ei > e_ref:
                al = max(0.0, min(1.0, (e_max - ei) / rng))
                k[idx - 1] = (1.0 - al) * k_max + al * k_min
    return k


def dneb(f_spring, t, f_perp):
    sp = f_spring - np.sum(f_spring * t) * t
    a, b = np.linalg.norm(sp), np.linalg.norm(f_perp)
    if a > 1e-10 and b > 1e-10:
        fn = f_perp / b
        return (sp - np.sum(sp * fn) * fn) * (2.0 / np.pi * np.arctan(b * b / (a * a)))
    return np.zeros_like(t)


def force_perp(f, t):
    """client/NEBForceProjection.cpp forcePerp"""
    return f - np.sum(f * t) * t


def climbing_image_force(f, t, f_dneb=0.0):
    """client/NEBForceProjection.cpp climbingImageForce: F - 2 (F.t) t + DNEB part"""
    return f - 2.0 * np.sum(f * t) * t + f_dneb


def zero_translation(f, n_free, n_atoms):
    """client/NEBForceProjection.cpp zeroTranslation: only when every atom is free"""
    return f - f.sum(axis=0) / n_atoms if n_free == n_atoms else f


def uniform_spring_par(ksp, t, dist_next, dist_prev):
    """client/NEBSpringForce.cpp:17-30"""
    return ksp * (dist_next - dist_prev) * t


def weighted_spring_par(k, i, t, dist_next, dist_prev):
    """client/NEBSpringForce.cpp:32-42: (k[i] d_next - k[i - 1] d_prev) t, full spring force zero"""
    return (k[i] * dist_next - k[i - 1] * dist_prev) * t


def cubic(U1, U2, F1, F2):
    """client/NEBSplineExtrema.cpp:105-108"""
    return U1, -F1, 3.0 * (U2 - U1) + 2.0 * F1 + F2, -2.0 * (U2 - U1) - (F1 + F2)


def band_forces(R, F, E, cell, n_free, ksp=5.0, climbing_image=False, old_tangent=False, projection=0,
                energy_weighted=False, k_min=0.97, k_max=9.7):
    """One band: R, F [M + 2, N, 3]; E [M + 2].  Returns (projected [M, N, 3], tangents [M, N, 3], climbing image)."""
    M, N = R.shape[0] - 2, R.shape[1]
    imax = 1
    for j in range(2, M + 1):
        if E[imax] < E[j]:
            imax = j
    e_max = E[imax]
    if energy_weighted:
        projection = 0
        kk = spring_constants(E, e_max, min(E[0], E[M + 1]), k_min, k_max)
    P = np.zeros((M, N, 3))
    T = np.zeros((M, N, 3))
    for i in range(1, M + 1):
        dn = pbc(R[i + 1] - R[i], cell)
        dp = pbc(R[i] - R[i - 1], cell)
        ln, lp = np.linalg.norm(dn), np.linalg.norm(dp)
        t = tangent(dn, dp, E[i], E[i - 1], E[i + 1], old_tangent)
        if energy_weighted:
            f_par = weighted_spring_par(kk, i, t, ln, lp)
            f_spring = np.zeros_like(t)
        else:
            f_par = uniform_spring_par(ksp, t, ln, lp)
            f_spring = ksp * pbc(dn - dp, cell)
        f = F[i]
        f_perp = force_perp(f, t)
        if climbing_image and i == imax:
            extra = dneb(f_spring, t, f_perp) if projection == 1 else 0.0
            out = climbing_image_force(f, t, extra)
        elif projection == 2:
            out = f_spring + f
        elif projection == 1:
            out = f_par + f_perp + dneb(f_spring, t, f_perp)
        else:
            out = f_par + f_perp
        out = zero_translation(out, n_free, N)
        P[i - 1], T[i - 1] = out, t
    return P, T, (imax if climbing_image else 0)
