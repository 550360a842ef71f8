"""NEB band forces (client/NudgedElasticBand.cpp:551-646 and the helpers it dispatches to) on the device against the
numpy oracle in oracle/neb_band.py, plus the analytic properties the reference's helper tests rely on."""
import os
import sys

import numpy as np
import pytest

from eon_b200 import B200Potential, Walkers, neb as NEB, search_params

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import neb_band as ONEB  # noqa: E402


def _random_bands(seed, B, M, N, cell, fixed_atoms=0):
    rng = np.random.default_rng(seed)
    L = np.diag(cell)
    a = rng.uniform(0.0, 1.0, (B, 1, N, 3)) * L
    b = a + rng.normal(0.0, 0.4, (B, 1, N, 3))
    lam = np.linspace(0.0, 1.0, M + 2).reshape(1, M + 2, 1, 1)
    R = a + lam * (b - a) + rng.normal(0.0, 0.03, (B, M + 2, N, 3))
    R = R - np.floor(R / L) * L          # wrapped: neighbouring images differ by whole cells for some atoms
    F = rng.normal(0.0, 0.5, (B, M + 2, N, 3))
    F[:, :, :fixed_atoms] = 0.0
    E = rng.normal(0.0, 1.0, (B, M + 2))
    return R, F, E


def _energy_shapes(E):
    E = E.copy()
    E[0] = np.array([0.0, 0.2, 0.5, 0.9, 0.6, 0.3, 0.1])[:E.shape[1]]        # single maximum
    if E.shape[0] > 1:
        E[1] = np.linspace(0.0, 1.0, E.shape[1])                              # monotonic: highest is the last movable
    if E.shape[0] > 2:
        E[2] = np.array([0.0, 0.4, 0.4, 0.1, 0.4, 0.4, 0.0])[:E.shape[1]]    # ties: first maximum wins, flat neighbours
    return E


CASES = [dict(), dict(climbing_image=1), dict(old_tangent=1), dict(projection=1), dict(projection=1, climbing_image=1),
         dict(projection=2), dict(energy_weighted=1, climbing_image=1), dict(energy_weighted=1, projection=2),
         dict(spring_constant=0.3, climbing_image=1)]


@pytest.mark.parametrize("n_fixed", [0, 5])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s%s" % kv for kv in c.items()) or "default")
def test_band_forces_match_oracle(backend, case, n_fixed):
    B, M, N = 5, 5, 23
    cell = np.diag([9.0, 10.0, 11.0])
    R, F, E = _random_bands(3 + n_fixed, B, M, N, cell, n_fixed)
    E = _energy_shapes(E)
    with B200Potential("lj", lib=backend) as pot:
        P, T, ci = NEB.neb_forces(pot, R.reshape(B, M + 2, -1), F.reshape(B, M + 2, -1), E, cell, N - n_fixed, **case)
    okw = dict(ksp=case.get("spring_constant", 5.0), climbing_image=bool(case.get("climbing_image")),
               old_tangent=bool(case.get("old_tangent")), projection=case.get("projection", 0),
               energy_weighted=bool(case.get("energy_weighted")))
    for b in range(B):
        P0, T0, ci0 = ONEB.band_forces(R[b], F[b], E[b], cell, N - n_fixed, **okw)
        assert ci[b] == ci0
        assert np.abs(T[b].reshape(M, N, 3) - T0).max() <= 1e-13
        assert np.abs(P[b].reshape(M, N, 3) - P0).max() <= 1e-11 * max(1.0, np.abs(P0).max())
    assert ci[0] == (3 if case.get("climbing_image") else 0)
    if case.get("climbing_image"):
        assert ci[1] == M and ci[2] == 1


@pytest.mark.parametrize("n_atoms", [256, 301, 601, 1000, 2500])
@pytest.mark.parametrize("case", [dict(climbing_image=1), dict(projection=1, climbing_image=1)], ids=["neb", "dneb"])
def test_band_forces_sizes(backend, case, n_atoms):
    """Both kernels on the GPU: rows staged in shared memory by bulk copies (odd and even row lengths, so aligned and
    shifted copies and the end-of-array fallback all occur) and, at 2500 atoms, the any-size kernel; all atoms free."""
    B, M = 2, 3
    cell = np.diag([28.638, 28.638, 40.125])
    R, F, E = _random_bands(n_atoms, B, M, n_atoms, cell)
    with B200Potential("lj", lib=backend) as pot:
        P, T, ci = NEB.neb_forces(pot, R.reshape(B, M + 2, -1), F.reshape(B, M + 2, -1), E, cell, n_atoms, **case)
    for b in range(B):
        P0, T0, ci0 = ONEB.band_forces(R[b], F[b], E[b], cell, n_atoms, climbing_image=True,
                                       projection=case.get("projection", 0))
        assert ci[b] == ci0
        assert np.abs(T[b].reshape(M, n_atoms, 3) - T0).max() <= 1e-13
        assert np.abs(P[b].reshape(M, n_atoms, 3) - P0).max() <= 1e-11 * max(1.0, np.abs(P0).max())


def test_band_forces_triclinic_cell(backend):
    """General (sheared) cell: the full fractional-coordinate minimum image instead of the orthorhombic shortcut."""
    B, M, N = 2, 4, 19
    cell = np.array([[9.0, 0.0, 0.0], [2.5, 8.0, 0.0], [-1.0, 1.5, 10.0]])
    rng = np.random.default_rng(21)
    frac = rng.uniform(0.0, 1.0, (B, 1, N, 3)) + np.linspace(0.0, 1.0, M + 2).reshape(1, -1, 1, 1) * rng.normal(0, 0.05, (B, 1, N, 3))
    frac -= np.floor(frac)                      # wrapped per image: neighbouring images may sit a whole cell apart
    R = frac @ cell
    F = rng.normal(0.0, 0.5, (B, M + 2, N, 3))
    E = rng.normal(0.0, 1.0, (B, M + 2))
    for case in (dict(climbing_image=1), dict(projection=1), dict(projection=2)):
        with B200Potential("lj", lib=backend) as pot:
            P, T, ci = NEB.neb_forces(pot, R.reshape(B, M + 2, -1), F.reshape(B, M + 2, -1), E, cell, N - 1, **case)
        for b in range(B):
            P0, T0, ci0 = ONEB.band_forces(R[b], F[b], E[b], cell, N - 1, climbing_image=bool(case.get("climbing_image")),
                                           projection=case.get("projection", 0))
            assert ci[b] == ci0
            assert np.abs(T[b].reshape(M, N, 3) - T0).max() <= 1e-12
            assert np.abs(P[b].reshape(M, N, 3) - P0).max() <= 1e-10 * max(1.0, np.abs(P0).max())


def test_band_force_properties(backend):
    """Unit tangents, nudging (only the spring acts along the tangent), climbing-image inversion, zero translation."""
    B, M, N = 3, 5, 17
    cell = np.diag([8.0, 8.0, 8.0])
    R, F, E = _random_bands(11, B, M, N, cell)
    E = _energy_shapes(E)
    D = 3 * N
    with B200Potential("lj", lib=backend) as pot:
        P, T, ci = NEB.neb_forces(pot, R.reshape(B, M + 2, D), F.reshape(B, M + 2, D), E, cell, N - 1, climbing_image=1)
        Pz, _, _ = NEB.neb_forces(pot, R.reshape(B, M + 2, D), F.reshape(B, M + 2, D), E, cell, N, climbing_image=1)
        # no cell: plain differences
        Pn, Tn, _ = NEB.neb_forces(pot, R.reshape(B, M + 2, D), F.reshape(B, M + 2, D), E, None, N - 1)
    assert np.allclose(np.linalg.norm(T, axis=2), 1.0, atol=1e-13)
    Fm = F[:, 1:M + 1].reshape(B, M, D)
    for b in range(B):
        for i in range(M):
            dn = ONEB.pbc(R[b, i + 2] - R[b, i + 1], cell)
            dp = ONEB.pbc(R[b, i + 1] - R[b, i], cell)
            along = P[b, i] @ T[b, i]
            if ci[b] == i + 1:
                assert along == pytest.approx(-(Fm[b, i] @ T[b, i]), abs=1e-11)
                perp = P[b, i] - along * T[b, i]
                assert np.allclose(perp, Fm[b, i] - (Fm[b, i] @ T[b, i]) * T[b, i], atol=1e-11)
            else:
                assert along == pytest.approx(5.0 * (np.linalg.norm(dn) - np.linalg.norm(dp)), abs=1e-11)
    assert np.abs(Pz.reshape(B, M, N, 3).sum(axis=2)).max() < 1e-11          # zeroTranslation when every atom is free
    assert np.abs(P.reshape(B, M, N, 3).sum(axis=2)).max() > 1e-3
    for b in range(B):
        P0, T0, _ = ONEB.band_forces(R[b], F[b], E[b], None, N - 1)
        assert np.abs(Pn[b].reshape(M, N, 3) - P0).max() <= 1e-11 and np.abs(Tn[b].reshape(M, N, 3) - T0).max() <= 1e-13
    assert NEB.convergence_force(P, ci).shape == (B,)
    assert np.allclose(NEB.convergence_force(P), np.linalg.norm(P, axis=2).max(axis=1))
    only = NEB.convergence_force(P, ci, ci_converged_only=True)
    assert all(only[b] == pytest.approx(np.linalg.norm(P[b, ci[b] - 1])) for b in range(B))


def test_degenerate_band_and_errors(backend):
    N, M = 4, 3
    R = np.tile(np.arange(3.0 * N).reshape(1, 1, -1), (1, M + 2, 1))   # every image identical: no direction at all
    F = np.ones((1, M + 2, 3 * N))
    E = np.zeros((1, M + 2))
    with B200Potential("lj", lib=backend) as pot:
        P, T, ci = NEB.neb_forces(pot, R, F, E, np.diag([50.0, 50.0, 50.0]), N - 1, climbing_image=1)
        assert np.isfinite(P).all() and np.all(T == 0.0)               # NEBTangent.cpp:21-33: left unnormalised
        assert np.array_equal(P, F[:, 1:M + 1]) and ci[0] == 1
        with pytest.raises(ValueError):
            NEB.neb_forces(pot, R[:, :2], F[:, :2], E[:, :2], None)
        with pytest.raises(RuntimeError, match="projection"):
            NEB.neb_forces(pot, R, F, E, None, projection=7)
        with pytest.raises(TypeError):
            NEB.neb_forces(pot, R, F, E, None, springiness=1.0)


def test_lj13_band_end_to_end(backend, oracle):
    """`band_forces`: one forceBatch over the movable images of two bands (reactant -> product of the reference's
    neb_morse fixture, and the reverse), masked / net-force-free forces, then the band kernel; against the C oracle's
    forces fed through the numpy band oracle."""
    a = oracle.load_fixture("lj13.json")
    b = oracle.load_fixture("lj13_product.json")
    Z, cell = a["atomic_nrs"], a["cell"]
    N, M = len(Z), 5
    lam = np.linspace(0.0, 1.0, M + 2).reshape(-1, 1, 1)
    band = a["positions"][None] + lam * ONEB.pbc(b["positions"] - a["positions"], cell)[None]
    R = np.stack([band, band[::-1]])
    opot = oracle.potential(oracle.POT_LJ)
    with B200Potential("lj", lib=backend) as pot:
        out = NEB.band_forces(pot, R.reshape(2, M + 2, -1), Z, cell, climbing_image=1)
        assert pot.forceCallCounter == 2 * 2 + 2 * M
    for k in range(2):
        E0 = np.empty(M + 2)
        F0 = np.zeros((M + 2, N, 3))
        for i in range(M + 2):
            E0[i], f = oracle.force(opot, R[k, i], Z, cell)
            F0[i] = f - f.mean(axis=0)
        assert np.abs(out["energies"][k] - E0).max() <= 1e-8
        P0, T0, ci0 = ONEB.band_forces(R[k], F0, E0, cell, N, climbing_image=True)
        assert out["climbing"][k] == ci0 and 1 <= ci0 <= M
        assert np.abs(out["tangents"][k].reshape(M, N, 3) - T0).max() <= 1e-12
        assert np.abs(out["projected"][k].reshape(M, N, 3) - P0).max() <= 1e-9 * max(1.0, np.abs(P0).max())
    # the reversed band is the same path: same energies mirrored, same climbing image mirrored
    assert np.allclose(out["energies"][0], out["energies"][1][::-1], atol=1e-10)
    assert out["climbing"][0] == M + 1 - out["climbing"][1] or np.isclose(*[out["energies"][k][out["climbing"][k]] for k in (0, 1)])


def test_reference_neb_lj_golden_endpoint_energies(backend, oracle):
    """client/unit_tests/data/reference/neb_lj.dat (the job of JobIntegrationTest.cpp:463-515: LJ13 reactant / product,
    3 images): `energy_reference` -39.965351 and `image4_energy` 631.805887 (product relative to reactant) are what one
    band evaluation can be held to.  The golden's 69 optimiser iterations are not: it comes from the SVN-era client
    (no climbing image -- its highest image has |F| = 30.9 but a projected force of 0.003 --, endpoints not minimised),
    the product lies 632 eV above the reactant, and the band between them is chaotic; the reference's own test only asks
    for a valid status, <= 209 force calls and energy_reference (JobIntegrationTest.cpp:492-513)."""
    a = oracle.load_fixture("lj13.json")
    b = oracle.load_fixture("lj13_product.json")
    Z, cell = a["atomic_nrs"], a["cell"]
    M = 3
    lam = np.linspace(0.0, 1.0, M + 2).reshape(-1, 1, 1)
    band = a["positions"][None] + lam * ONEB.pbc(b["positions"] - a["positions"], cell)[None]
    with B200Potential("lj", lib=backend) as pot:
        out = NEB.band_forces(pot, band.reshape(1, M + 2, -1), Z, cell, climbing_image=1)
    E = out["energies"][0]
    assert E[0] == pytest.approx(-39.965351, abs=1e-6)          # energy_reference
    assert E[M + 1] - E[0] == pytest.approx(631.805887, abs=1e-5)  # image4_energy
    assert np.all(np.isfinite(E))


@pytest.mark.parametrize("system", ["lj13", "morse_pt"])
def test_resident_band_evaluation_matches_host_path(backend, oracle, system):
    """``eonb200_neb_band_forces`` (one call: all images evaluated and projected on the device) against
    `band_forces` (forceBatch + host masking + band kernel): all-free LJ13 (mean force removed, zeroTranslation) and the
    Morse slab with fixed atoms (axes masked in the kernel)."""
    M, B = 4, 3
    if system == "lj13":
        a, b = oracle.load_fixture("lj13.json"), oracle.load_fixture("lj13_product.json")
        R0, R1, fixed, kind = a["positions"], b["positions"], None, "lj"
    else:
        a = oracle.load_fixture("morse_pt_saddle.json")
        R0, fixed, kind = a["positions"], a["fixed"], "morse_pt"
        R1 = R0 + (a["displacement_positions"] - R0) * 2.0
    Z, cell = a["atomic_nrs"], a["cell"]
    N = len(Z)
    lam = np.linspace(0.0, 1.0, M + 2).reshape(1, -1, 1, 1)
    rng = np.random.default_rng(2)
    free = 1.0 if fixed is None else (1 - fixed).reshape(1, 1, N, 3)
    R = R0[None, None] + lam * ONEB.pbc(R1 - R0, cell)[None, None] + rng.normal(0.0, 0.01, (B, M + 2, N, 3)) * free
    R = R.reshape(B, M + 2, -1)
    for case in (dict(climbing_image=1), dict(projection=1)):
        with B200Potential(kind, lib=backend) as pot:
            ref = NEB.band_forces(pot, R, Z, cell, fixed, **case)
        with B200Potential(kind, lib=backend) as pot:
            out = NEB.band_forces_resident(pot, R, Z, cell, fixed, want_tangents=True, **case)
            assert pot.forceCallCounter == B * (M + 2)
            again = NEB.band_forces_resident(pot, R, Z, cell, fixed, **case)  # cached neighbour lists, no tangents
        assert np.array_equal(out["climbing"], ref["climbing"])
        assert np.abs(out["energies"] - ref["energies"]).max() <= 1e-9
        assert np.abs(out["tangents"] - ref["tangents"]).max() <= 1e-12
        scale = max(1.0, np.abs(ref["projected"]).max())
        assert np.abs(out["projected"] - ref["projected"]).max() <= 1e-10 * scale
        assert np.abs(again["projected"] - out["projected"]).max() <= 1e-10 * scale and again["tangents"] is None


@pytest.mark.parametrize("kind,fixture", [("eam", "neb_al.json"), ("morse_pt", "morse_pt_saddle.json")])
def test_cell_change_on_one_handle_invalidates_cached_lists(backend, oracle, kind, fixture):
    """A handle evaluated in one cell and then in another (same positions, so no atom has moved) must not reuse the
    lists and image shifts of the first cell: the build cell is recorded with every cached list and compared in the
    Verlet check.  Checked against a fresh handle, through the resident band call and forceBatch."""
    a = oracle.load_fixture(fixture)
    Z, N = a["atomic_nrs"], len(a["atomic_nrs"])
    cell_a = np.array(a["cell"], dtype=np.float64)
    cell_b = cell_a.copy()
    cell_b[0, 0] *= 0.97  # atoms near the +x face now see their periodic neighbours 3 % closer
    M, B = 3, 2
    rng = np.random.default_rng(5)
    R = a["positions"][None, None] + rng.normal(0.0, 0.01, (B, M + 2, N, 3))
    R = R.reshape(B, M + 2, -1)
    with B200Potential(kind, lib=backend) as pot:
        NEB.band_forces_resident(pot, R, Z, cell_a, None, climbing_image=1)
        out = NEB.band_forces_resident(pot, R, Z, cell_b, None, climbing_image=1)
        F1, U1, _ = pot.forceBatch(B * (M + 2), N, R.reshape(B * (M + 2), -1), Z, cell_a.reshape(9))
    with B200Potential(kind, lib=backend) as pot:
        ref = NEB.band_forces_resident(pot, R, Z, cell_b, None, climbing_image=1)
    with B200Potential(kind, lib=backend) as pot:
        F0, U0, _ = pot.forceBatch(B * (M + 2), N, R.reshape(B * (M + 2), -1), Z, cell_a.reshape(9))
    assert np.abs(out["energies"] - ref["energies"]).max() <= 1e-9
    assert np.abs(out["projected"] - ref["projected"]).max() <= 1e-10 * max(1.0, np.abs(ref["projected"]).max())
    assert np.abs(U1 - U0).max() <= 1e-9 and np.abs(F1 - F0).max() <= 1e-10 * max(1.0, np.abs(F0).max())


def _m3(*vals):
    return np.array(vals, dtype=float).reshape(3, 3)


def _device_image(backend, d_next, d_prev, e, e_prev, e_next, force, n_free=2, **kw):
    """One movable image between two endpoints on the device: tangent and projected force of the middle image."""
    cur = np.full((3, 3), 50.0) + np.arange(9).reshape(3, 3)  # far from the faces of a 100 A cell: no wrap
    R = np.stack([cur - d_prev, cur, cur + d_next]).reshape(1, 3, 9)
    F = np.zeros((1, 3, 9))
    F[0, 1] = np.asarray(force, dtype=float).reshape(9)
    E = np.array([[e_prev, e, e_next]], dtype=float)
    with B200Potential("lj", lib=backend) as pot:
        P, T, ci = NEB.neb_forces(pot, R, F, E, np.eye(3) * 100.0, n_free, **kw)
    return P[0, 0].reshape(3, 3), T[0, 0].reshape(3, 3), int(ci[0])


def test_reference_neb_subsystem_vectors(backend):
    """Every hand-computed expectation of client/unit_tests/NEBSubsystemTest.cpp:47-353 (tangents, forcePerp,
    climbingImageForce, computeDNEB, zeroTranslation, springs) and :376-477 (cubic extrema), against the numpy oracle's
    helpers, the product's `spline_extrema`, and -- wherever a three-image band can express the case -- the device
    kernel."""
    nx = _m3(1, 0, 0, 1, 0, 0, 1, 0, 0)
    # SimpleTangent (:47-74): normalised, along diffNext, blind to energies and diffPrev
    t = ONEB.tangent(nx, -nx, 0.0, 0.0, 0.0, old_tangent=True)
    assert np.linalg.norm(t) == pytest.approx(1.0, rel=1e-12) and (t * nx).sum() > 0
    d2, d3 = _m3(0, 2, 0, 0, 2, 0, 0, 2, 0), _m3(-3, 0, 0, -3, 0, 0, -3, 0, 0)
    assert np.allclose(ONEB.tangent(d2, d3, 1.0, 0.5, 2.0, True), ONEB.tangent(d2, d3, 5.0, 10.0, -3.0, True), atol=1e-12)
    _, T, _ = _device_image(backend, d2, d3, 1.0, 0.5, 2.0, np.zeros(9), old_tangent=1)
    assert np.allclose(T, d2 / np.linalg.norm(d2), atol=1e-12)
    # ImprovedTangent (:78-151)
    dn, dp = _m3(1, 0, 0, 0, 0, 0, 0, 0, 0), _m3(0, 1, 0, 0, 0, 0, 0, 0, 0)
    for (e, ep, en), want in (((1.0, 0.0, 2.0), dn), ((1.0, 2.0, 0.0), dp), ((0.0, 3.0, 1.0), dn * 1.0 + dp * 3.0),
                              ((5.0, 2.0, 1.0), dn * 3.0 + dp * 4.0)):
        want = want / np.linalg.norm(want)
        got = ONEB.tangent(dn, dp, e, ep, en)
        assert np.linalg.norm(got) == pytest.approx(1.0, rel=1e-12) and np.allclose(got, want, atol=1e-10)
        _, T, _ = _device_image(backend, dn, dp, e, ep, en, np.zeros(9))
        assert np.allclose(T, want, atol=1e-10)
    # forcePerp (:155-197)
    tx = dn / np.linalg.norm(dn)
    f34 = _m3(3, 4, 0, 0, 0, 0, 0, 0, 0)
    fp = ONEB.force_perp(f34, tx)
    assert abs((fp * tx).sum()) < 1e-12 and abs(fp[0, 0]) < 1e-12 and fp[0, 1] == pytest.approx(4.0)
    assert np.linalg.norm(ONEB.force_perp(_m3(5, 0, 0, 0, 0, 0, 0, 0, 0), tx)) < 1e-12
    f7 = _m3(0, 7, 0, 0, 0, 0, 0, 0, 0)
    assert np.allclose(ONEB.force_perp(f7, tx), f7, atol=1e-12)
    # device: equal spacing along x, rising energies -> tangent x, no parallel spring force: projected = forcePerp
    P, T, _ = _device_image(backend, dn, dn, 1.0, 0.0, 2.0, f34)
    assert np.allclose(T, tx, atol=1e-12) and np.allclose(P, _m3(0, 4, 0, 0, 0, 0, 0, 0, 0), atol=1e-12)
    # climbingImageForce (:201-229): (3, 4) -> (-3, 4); with a DNEB part (0, 1): (3, 0) -> (-3, 1)
    assert np.allclose(ONEB.climbing_image_force(f34, tx)[0], [-3.0, 4.0, 0.0])
    assert np.allclose(ONEB.climbing_image_force(_m3(3, 0, 0, 0, 0, 0, 0, 0, 0), tx, _m3(0, 1, 0, 0, 0, 0, 0, 0, 0))[0],
                       [-3.0, 1.0, 0.0])
    P, _, ci = _device_image(backend, dn, dn, 1.0, 0.0, 2.0, f34, climbing_image=1)  # the only image climbs
    assert ci == 1 and np.allclose(P[0], [-3.0, 4.0, 0.0], atol=1e-12)
    # computeDNEB (:233-281)
    assert np.linalg.norm(ONEB.dneb(_m3(0, 2, 0, 0, 0, 0, 0, 0, 0), tx, np.zeros((3, 3)))) < 1e-12
    assert np.linalg.norm(ONEB.dneb(_m3(5, 0, 0, 0, 0, 0, 0, 0, 0), tx, _m3(0, 3, 0, 0, 0, 0, 0, 0, 0))) < 1e-12
    dn3 = ONEB.dneb(_m3(0, 0, 2, 0, 0, 0, 0, 0, 0), tx, _m3(0, 1, 0, 0, 0, 0, 0, 0, 0))
    assert dn3[0, 2] == pytest.approx(2.0 * (2.0 / np.pi * np.arctan(0.25)), rel=1e-10)
    assert abs(dn3[0, 0]) < 1e-12 and abs(dn3[0, 1]) < 1e-12
    # zeroTranslation (:285-309)
    f369 = _m3(3, 6, 9, 1, 2, 3, 2, 4, 6)
    assert np.allclose(ONEB.zero_translation(f369, 3, 3).sum(axis=0), 0.0, atol=1e-12)
    assert np.array_equal(ONEB.zero_translation(f369, 2, 3), f369)
    P3, _, _ = _device_image(backend, dn, dn, 1.0, 0.0, 2.0, f369, n_free=3)
    P2, _, _ = _device_image(backend, dn, dn, 1.0, 0.0, 2.0, f369, n_free=2)
    assert np.allclose(P3.sum(axis=0), 0.0, atol=1e-12)
    assert np.allclose(P2, ONEB.force_perp(f369, tx), atol=1e-12)  # untouched beyond the projection
    # UniformSpring (:313-345): ksp (dNext - dPrev) t
    assert np.linalg.norm(ONEB.uniform_spring_par(5.0, tx, 1.0, 1.0)) < 1e-12
    assert np.allclose(ONEB.uniform_spring_par(5.0, tx, 2.0, 1.0), 5.0 * tx, atol=1e-12)
    P, _, _ = _device_image(backend, 2.0 * dn, dn, 1.0, 0.0, 2.0, np.zeros(9), spring_constant=5.0)
    assert np.allclose(P, 5.0 * tx, atol=1e-12)
    # WeightedSpring (:349-372): k = {2, 4, 6}, image 1, distances 3 and 2 -> 8 t
    ty = dp / np.linalg.norm(dp)
    assert np.allclose(ONEB.weighted_spring_par([2.0, 4.0, 6.0], 1, ty, 3.0, 2.0), 8.0 * ty, atol=1e-12)
    # cubic extrema (:376-477)
    a, b, c, d = ONEB.cubic(0.0, 0.0, -1.0, 1.0)
    assert (a, b, c, d) == (0.0, 1.0, -1.0, 0.0)
    pos, en, cu = NEB.spline_extrema([0.0, 0.0], [-1.0, 1.0], [1.0])
    assert pos.tolist() == [0.5] and en[0] == pytest.approx(0.25) and cu[0] == pytest.approx(-2.0)
    a, b, c, d = ONEB.cubic(0.0, 0.5, -2.0, 0.5)
    assert (b, c, d) == (2.0, -2.0, 0.5) and c * c - 3.0 * b * d > 0
    pos, en, cu = NEB.spline_extrema([0.0, 0.5], [-2.0, 0.5], [1.0])
    assert len(pos) >= 1
    for f in pos:
        assert abs((3 * d * f + 2 * c) * f + b) < 1e-10
    a, b, c, d = ONEB.cubic(0.0, 10.0, -20.0, -20.0)
    assert c * c - 3.0 * b * d < 0 and len(NEB.spline_extrema([0.0, 10.0], [-20.0, -20.0], [1.0])[0]) == 0


@pytest.mark.parametrize("images", [4, 5])
def test_band_optimiser_converges_to_the_dimer_saddle(backend, oracle, images):
    """`run_bands` (NudgedElasticBand::compute with LBFGS over the concatenated images, every evaluation one
    eonb200_neb_band_forces call): a climbing-image band between the two minima of the reference's Pt adatom hop
    (process_search_morse_pt.dat) converges, and its climbing image / spline maximum sit on the saddle the dimer finds
    (saddle energy of the golden, 1e-3 eV).  Two identical bands in one batch stay identical."""
    g = oracle.goldens()["process_search_lbfgs"]
    d = oracle.load_fixture("morse_pt_saddle.json")
    Z, box, fixed = d["atomic_nrs"], d["cell"], d["fixed"]
    prm = search_params(lib=backend, opt_method=1, opt_converged_force=1e-3, saddle_max_energy=10.0)
    with B200Potential("morse_pt", lib=backend) as pot:
        wk = Walkers(pot, Z, fixed, box, d["positions"].reshape(1, -1))
        E0, _ = wk.matter_forces()
        wk.set(positions=d["displacement_positions"].reshape(1, -1), modes=d["direction"].reshape(1, -1))
        out = wk.process_search(prm, E0)
        wk.close()
        m1, m2 = out["min1"].positions[0], out["min2"].positions[0]
        r = NEB.run_bands(pot, np.stack([m1, m1]), np.stack([m2, m2]), Z, box, fixed, images=images, spring=5.0,
                          max_iterations=300, converged_force=0.01, max_move=0.2)
    assert np.all(r["status"] == 0) and np.all(r["iterations"] < 100)
    assert np.array_equal(r["force_calls"], (r["iterations"] + 2) * images)  # first evaluation + curvature probe + steps
    E = r["energies"]
    assert np.abs(E[0] - E[1]).max() < 1e-12
    assert E[0].max() == pytest.approx(g["energy_saddle"], abs=1e-3)
    assert E[0, 0] == pytest.approx(out["min1"].energies[0], abs=1e-8) and E[0, -1] == pytest.approx(out["min2"].energies[0], abs=1e-8)
    pos, en, cu = r["extrema"][0]
    k = int(np.argmax(en))
    assert en[k] == pytest.approx(g["energy_saddle"], abs=1e-3) and cu[k] < 0
    assert abs(pos[k] - r["climbing"][0]) < 0.25
