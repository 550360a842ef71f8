"""Harmonic prefactors (client/Prefactor.cpp:22-306) through one batched force call: moved-atom selection, the
eigenvalue rules, and the known answer of data/reference/process_search_morse_pt.dat."""
import numpy as np
import pytest

from eon_b200 import B200Potential, Walkers, prefactor as PF, process_search as PS, search_params


def test_moved_atoms_rules():
    cell = np.diag([20.0, 20.0, 20.0])
    sad = np.array([[1.0, 1.0, 1.0], [3.0, 1.0, 1.0], [19.5, 1.0, 1.0], [10.0, 10.0, 10.0], [1.0, 3.5, 1.0]])
    fixed = np.zeros((5, 3), dtype=int)
    fixed[1] = 1
    m1 = sad.copy()
    m2 = sad.copy()
    m1[0, 0] -= 0.4   # atom 0 moves 0.4 towards min1
    m2[4, 1] += 0.1   # atom 4 moves a little
    m1[1, 2] += 5.0   # fixed atom: its motion is masked out
    # movedAtoms: atom 0 passes the threshold; its free neighbours within 3.3 at the saddle follow in index order
    # (2 is across the periodic face at 1.5, 4 at 2.5; 1 is fixed; 3 is far)
    assert PF.moved_atoms(m1, sad, m2, cell, fixed, min_displacement=0.25, within_radius=3.3) == [0, 2, 4]
    assert PF.moved_atoms(m1, sad, m2, cell, fixed, min_displacement=0.05, within_radius=1.0) == [0, 4]
    assert PF.moved_atoms(m1, sad, m2, cell, fixed, min_displacement=0.5) == []
    # movedAtomsPct: 0.4 of 0.5 total is 80 % <= 90 % -> the next mover (atom 4) is taken too
    assert PF.moved_atoms_pct(m1, sad, m2, cell, fixed, filter_fraction=0.9, within_radius=1.0) == [0, 4]
    assert PF.moved_atoms_pct(m1, sad, m2, cell, fixed, filter_fraction=0.7, within_radius=1.0) == [0]
    assert PF.moved_atoms_pct(m1, sad, m2, cell, fixed, filter_fraction=0.7, within_radius=3.3) == [0, 2, 4]
    assert PF.moved_atoms_pct(sad, sad, sad, cell, fixed) == []


def test_rate_rules():
    f1 = np.array([1.0, 2.0, 3.0])
    fs = np.array([-0.5, 1.5, 2.5])
    f2 = np.array([0.5, 2.5, 3.5])
    p1, p2 = PF.rate_prefactors(f1, fs, f2)
    assert p1 == pytest.approx(np.sqrt(1.0 * 2.0 * 3.0 / (1.5 * 2.5)) / (2 * np.pi * 10.18e-15), rel=1e-14)
    assert p2 == pytest.approx(np.sqrt(0.5 * 2.5 * 3.5 / (1.5 * 2.5)) / (2 * np.pi * 10.18e-15), rel=1e-14)
    assert PF.rate_prefactors(f1, np.abs(fs), f2) is None          # no negative mode at the saddle
    assert PF.rate_prefactors(-f1, fs, f2) is None                 # negative mode at a minimum
    assert PF.rate_prefactors(f1, fs, f2[:2]) is None              # sizes differ
    q1, q2 = PF.rate_prefactors(f1, fs, f2, rate="qqhtst", temperature=300.0)
    assert q1 > 0 and q2 > 0 and q1 != p1
    # high-temperature limit of the quasi-quantum rate is the classical one
    h1, _ = PF.rate_prefactors(f1, fs, f2, rate="qqhtst", temperature=3.0e5)
    assert h1 == pytest.approx(p1, rel=1e-3)
    z = np.concatenate([np.zeros(6), np.arange(1.0, 4.0)])
    assert PF.remove_zero_freqs(z, 3).tolist() == [1.0, 2.0, 3.0]
    assert PF.remove_zero_freqs(z[1:], 3) is z[1:] or PF.remove_zero_freqs(z[1:], 3).size == 8  # not 3N: untouched
    # five trivial modes instead of six: Hessian.cpp:296-301 only logs it and returns the shortened vector
    with pytest.warns(UserWarning):
        assert PF.remove_zero_freqs(np.concatenate([np.zeros(5), np.arange(1.0, 5.0)]), 3).tolist() == [1.0, 2.0, 3.0, 4.0]


def test_prefactors_known_answer(backend, oracle):
    """prefactor_* and force_calls_prefactors of data/reference/process_search_morse_pt.dat
    (JobIntegrationTest.cpp:1020-1069) from the batched process search followed by one batched Hessian call."""
    g = oracle.goldens()["process_search_lbfgs"]
    d = oracle.load_fixture("morse_pt_saddle.json")
    Z, box, fixed, masses = d["atomic_nrs"], d["cell"], d["fixed"], d["masses"]
    prm = search_params(lib=backend, opt_method=1, opt_converged_force=1e-3, saddle_max_energy=10.0)
    with B200Potential("morse_pt", lib=backend) as pot:
        wk = Walkers(pot, Z, fixed, box, d["positions"].reshape(1, -1))
        E0, _ = wk.matter_forces()
        wk.set(positions=d["displacement_positions"].reshape(1, -1), modes=d["direction"].reshape(1, -1))
        out = wk.process_search(prm, E0)
        wk.close()
        res = PS.process_results(out, d["positions"], box, fixed, max_energy=10.0, potential_type="MORSE_PT")
        before = pot.forceCallCounter
        PF.apply_prefactors(res, pot, Z, box, fixed, masses)
        assert pot.forceCallCounter - before == 3 * (1 + 3)  # one batch: three structures x (undisplaced + 3 displaced)
    r = res[0]
    assert r["termination_reason"] == 0
    assert r["force_calls_prefactors"] == g["force_calls_prefactors"]
    assert r["total_force_calls"] == r["force_calls_saddle"] + r["force_calls_minimization"] + 9
    # measured agreement with the golden's 13 digits: 3e-13 relative
    assert r["prefactor_reactant_to_product"] == pytest.approx(g["prefactor_reactant_to_product"], rel=1e-9)
    assert r["prefactor_product_to_reactant"] == pytest.approx(g["prefactor_product_to_reactant"], rel=1e-9)
    assert "%.12e prefactor_reactant_to_product" % r["prefactor_reactant_to_product"] in PS.format_results_dat(r)
    # window rule (ProcessSearchJob.cpp:303-312)
    res2 = PS.process_results(out, d["positions"], box, fixed, max_energy=10.0)
    with B200Potential("morse_pt", lib=backend) as pot:
        PF.apply_prefactors(res2, pot, Z, box, fixed, masses, min_value=1e13)
    assert res2[0]["termination_reason"] == PF.STATUS_BAD_PREFACTOR
