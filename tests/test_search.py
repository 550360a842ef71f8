"""Device-resident walkers against the CPU oracle and the reference's known-answer job results.

Tolerances are the north star's: each walker converges to the same saddle within 1e-5 eV and 1e-4 A.
Where the reference pins force-call counts (client/unit_tests/data/reference/*.dat) they are asserted too.
"""
import numpy as np
import pytest

from eon_b200 import B200Potential, Walkers, search_params, workloads as WL

E_TOL = 1e-5
R_TOL = 1e-4


def mic_max(a, b, cell):
    d = (a - b).reshape(-1, 3)
    L = np.diag(cell)
    d = d - np.round(d / L) * L
    return np.abs(d).max()


def same_params(oracle, backend, **kw):
    return oracle.search_params(**kw), search_params(lib=backend, **kw)


def test_matter_forces_mask_and_net_force(backend, oracle):
    """Matter::getForces: fixed-axis mask; mean force removed only when nothing is fixed (Matter.cpp:324-333,531-536)."""
    for kind, okind, fixture in (("morse_pt", "POT_MORSE_PT", "akmc_pt.json"), ("eam", "POT_EAM", "akmc_al.json"),
                                 ("lj", "POT_LJ", "lj13.json")):
        d = oracle.load_fixture(fixture)
        rng = np.random.default_rng(1)
        W = 3
        R = d["positions"][None] + rng.uniform(-0.05, 0.05, (W,) + d["positions"].shape)
        opot = oracle.potential(getattr(oracle, okind))
        for fixed in (d["fixed"], np.zeros_like(d["fixed"])):
            with B200Potential(kind, lib=backend) as pot:
                wk = Walkers(pot, d["atomic_nrs"], fixed, d["cell"], R)
                for rnf in (True, False):
                    E, F = wk.matter_forces(rnf)
                    for w in range(W):
                        U0, F0 = oracle.matter_forces(opot, R[w], d["atomic_nrs"], fixed, d["cell"], rnf)
                        assert abs(E[w] - U0) < 1e-8
                        assert np.abs(F[w] - F0).max() <= 1e-10 * np.abs(F0).max()
                        assert np.all(F[w][fixed.astype(bool)] == 0.0)
                wk.close()


@pytest.mark.parametrize("label,kw,golden", [
    ("dimer-cg", {}, "saddle_dimer_cg"),
    ("dimer-lbfgs", {"opt_method": 1}, "process_search_lbfgs"),
    ("lanczos-cg", {"minmode_method": 1}, "saddle_lanczos_cg"),
    ("dimer-fire", {"opt_method": 2}, None),
    ("dimer-sd-rotation", {"dimer_opt_method": 0}, None),
    ("dimer-lbfgs-rotation", {"dimer_opt_method": 2}, None),
])
def test_morse_pt_saddle_search_known_answer(backend, oracle, label, kw, golden):
    """systems/saddle_search (336 Pt, one free atom): saddle_search_morse_pt.dat, saddle_search_lanczos_morse_pt.dat,
    process_search_morse_pt.dat."""
    g = oracle.goldens()
    d = oracle.load_fixture("morse_pt_saddle.json")
    Z, box, fixed = d["atomic_nrs"], d["cell"], d["fixed"]
    oprm, prm = same_params(oracle, backend, **kw)
    opot = oracle.potential(oracle.POT_MORSE_PT)
    E0, _ = oracle.force(opot, d["positions"], Z, box)
    ref, Rref, Mref = oracle.saddle_search(opot, oprm, d["displacement_positions"], Z, fixed, box, d["direction"], E0)
    with B200Potential("morse_pt", lib=backend) as pot:
        wk = Walkers(pot, Z, fixed, box, d["displacement_positions"].reshape(1, -1), d["direction"].reshape(1, -1))
        wk.saddle_search(prm, E0)
        h = wk.harvest()
        wk.close()
    assert h.status[0] == ref.status == 0
    assert abs(h.energies[0] - ref.energy) < E_TOL
    assert mic_max(h.positions[0], Rref, box) < R_TOL
    assert abs(h.eigenvalues[0] - ref.eigenvalue) < 1e-3
    assert abs(abs(np.vdot(h.modes[0], Mref)) - 1.0) < 1e-4
    assert h.iterations[0] == ref.iterations and h.force_calls[0] == ref.force_calls
    assert np.all(h.modes[0][fixed.astype(bool)] == 0.0)  # DimerTest.cpp:573-639
    if golden:
        k = g[golden]
        assert h.energies[0] == pytest.approx(k["energy_saddle"], abs=2e-5)
        if "force_calls_saddle" in k:
            assert h.force_calls[0] <= k["force_calls_saddle"]
        if golden == "saddle_dimer_cg":
            assert h.iterations[0] == k["iterations"] and h.force_calls[0] == k["force_calls_saddle"]
            assert h.eigenvalues[0] == pytest.approx(k["eigenvalue"], abs=5e-6)


@pytest.mark.parametrize("method,key", [(1, "lj_minimization"), (0, "lj_min_cg"), (2, "lj_min_fire")])
def test_lj13_relaxation_known_answer(backend, oracle, method, key):
    """minimization_lj.dat / min_lj_cg.dat / min_lj_fire.dat."""
    g = oracle.goldens()[key]
    d = oracle.load_fixture("lj13.json")
    oprm, prm = same_params(oracle, backend, opt_method=method, opt_converged_force=1e-3)
    ref, Rref = oracle.relax(oracle.potential(oracle.POT_LJ), oprm, d["positions"], d["atomic_nrs"], d["fixed"], d["cell"])
    with B200Potential("lj", lib=backend) as pot:
        wk = Walkers(pot, d["atomic_nrs"], d["fixed"], d["cell"], d["positions"].reshape(1, -1))
        wk.relax(prm)
        h = wk.harvest()
        wk.close()
    assert h.status[0] == 0 and ref.converged == 1
    assert h.energies[0] == pytest.approx(g["energy"], abs=2e-6)
    assert abs(h.energies[0] - ref.energy) < 1e-8
    assert h.force_calls[0] == ref.force_calls <= g["force_calls"]
    assert np.abs(h.positions[0] - Rref).max() < 1e-6


def test_process_search_known_answer(backend, oracle):
    """process_search_morse_pt.dat: saddle, then min1/min2 = saddle -+ max_move * mode relaxed with LBFGS;
    product energy and both barriers (ProcessSearchJob.cpp:188-289)."""
    g = oracle.goldens()["process_search_lbfgs"]
    d = oracle.load_fixture("morse_pt_saddle.json")
    Z, box, fixed = d["atomic_nrs"], d["cell"], d["fixed"]
    # JobIntegrationTest.cpp:1020-1040: lbfgs, converged_force 0.001, max_energy 10
    prm = search_params(lib=backend, opt_method=1, opt_converged_force=1e-3, saddle_max_energy=10.0)
    with B200Potential("morse_pt", lib=backend) as pot:
        wk = Walkers(pot, Z, fixed, box, d["positions"].reshape(1, -1))
        E0, _ = wk.matter_forces()
        wk.set(positions=d["displacement_positions"].reshape(1, -1), modes=d["direction"].reshape(1, -1))
        out = wk.process_search(prm, E0)
        wk.close()
    Es = out["saddle"].energies[0]
    mins = sorted([out["min1"].energies[0], out["min2"].energies[0]])
    assert out["saddle"].status[0] == 0 and out["min1"].status[0] == 0 and out["min2"].status[0] == 0
    assert Es == pytest.approx(g["energy_saddle"], abs=2e-6)
    assert mins[0] == pytest.approx(g["energy_product"], abs=2e-6)
    assert mins[1] == pytest.approx(g["energy_reactant"], abs=2e-6)
    assert Es - mins[1] == pytest.approx(g["barrier_reactant_to_product"], abs=5e-6)
    assert Es - mins[0] == pytest.approx(g["barrier_product_to_reactant"], abs=5e-6)
    calls = out["saddle"].force_calls[0] + out["min1"].force_calls[0] + out["min2"].force_calls[0]
    assert out["saddle"].force_calls[0] == g["force_calls_saddle"]
    assert calls <= g["force_calls_saddle"] + g["force_calls_minimization"] + 2


def test_eam_walkers_match_oracle(backend, oracle):
    """C2-style walkers (akmc-al slab, displaced adatom neighbourhood, conservative EAM): several different walkers in
    one launch land on the oracle's saddles."""
    cfg = WL.AKMC_AL
    s = WL.load_structure(cfg["fixture"])
    kw = dict(opt_method=1, opt_converged_force=1e-3, minmode_method=0, saddle_max_energy=10.0, saddle_max_iterations=300)
    oprm, prm = same_params(oracle, backend, **kw)
    opot = oracle.potential(oracle.POT_EAM_NL, conservative=True)
    relp_o, relp = same_params(oracle, backend, opt_method=1, opt_converged_force=1e-3)
    res, Rmin = oracle.relax(opot, relp_o, s.positions, s.atomic_nrs, s.fixed, s.cell)
    assert res.converged == 1
    W = 3
    _, R, M = WL.walkers_for(cfg, Rmin, W)
    with B200Potential("eam", lib=backend, conservative=True) as pot:
        wk = Walkers(pot, s.atomic_nrs, s.fixed, s.cell, s.positions.reshape(1, -1).repeat(W, 0))
        wk.relax(relp)
        hm = wk.harvest()
        assert np.all(hm.status == 0) and np.abs(hm.energies - res.energy).max() < E_TOL
        wk.set(positions=R, modes=M)
        wk.saddle_search(prm, res.energy)
        h = wk.harvest()
        wk.close()
    for w in range(W):
        ref, Rref, _ = oracle.saddle_search(opot, oprm, R[w], s.atomic_nrs, s.fixed, s.cell, M[w], res.energy)
        assert h.status[w] == ref.status
        if ref.status == 0:
            assert abs(h.energies[w] - ref.energy) < E_TOL
            assert mic_max(h.positions[w], Rref, s.cell) < R_TOL


def test_status_codes(backend, oracle):
    """MinModeSaddleSearch status words: 5 max iterations, 3 high energy (MinModeSaddleSearch.h:29-53)."""
    d = oracle.load_fixture("morse_pt_saddle.json")
    Z, box, fixed = d["atomic_nrs"], d["cell"], d["fixed"]
    opot = oracle.potential(oracle.POT_MORSE_PT)
    E0, _ = oracle.force(opot, d["positions"], Z, box)
    for kw, want in (({"saddle_max_iterations": 3}, 5), ({"saddle_max_energy": 0.05}, 3)):
        oprm, prm = same_params(oracle, backend, **kw)
        ref, _, _ = oracle.saddle_search(opot, oprm, d["displacement_positions"], Z, fixed, box, d["direction"], E0)
        with B200Potential("morse_pt", lib=backend) as pot:
            wk = Walkers(pot, Z, fixed, box, d["displacement_positions"].reshape(1, -1), d["direction"].reshape(1, -1))
            wk.saddle_search(prm, E0)
            h = wk.harvest(geometry=False)
            wk.close()
        assert h.status[0] == ref.status == want
        assert h.iterations[0] == ref.iterations and h.force_calls[0] == ref.force_calls


def test_force_call_counter_counts_walker_evaluations(backend, oracle):
    """Potential::forceCallCounter: one increment per walker evaluation (Potential.h:119-120)."""
    d = oracle.load_fixture("lj13.json")
    with B200Potential("lj", lib=backend) as pot:
        wk = Walkers(pot, d["atomic_nrs"], d["fixed"], d["cell"], np.stack([d["positions"].reshape(-1)] * 4))
        wk.matter_forces()
        assert pot.forceCallCounter == 4
        wk.relax(search_params(lib=backend, opt_method=1, opt_converged_force=1e-3))
        h = wk.harvest(geometry=False)
        assert pot.forceCallCounter == 4 + int(h.force_calls.sum())
        wk.close()
