"""Host-side tail of the batched process search: Matter::compare connectivity, endpoint
swap, barrier / abort rules and the results.dat keys of client/ProcessSearchJob.cpp:234-363."""
import numpy as np
import pytest

from eon_b200 import B200Potential, Walkers, process_search as PS, search_params
from eon_b200.walkers import Harvest


def _h(status, energies, positions, force_calls):
    n = len(status)
    return Harvest(status=np.array(status, dtype=np.int32), iterations=np.zeros(n, dtype=np.int64),
                   force_calls=np.array(force_calls, dtype=np.int64), energies=np.array(energies, dtype=float),
                   eigenvalues=np.zeros(n), convergence=np.zeros(n), positions=np.array(positions, dtype=float),
                   modes=None)


def test_compare_and_per_atom_norm():
    cell = np.diag([10.0, 10.0, 10.0])
    a = np.array([[0.1, 0.1, 0.1], [5.0, 5.0, 5.0]])
    b = a.copy()
    b[0, 0] = 9.95  # across the face: minimum image distance 0.15
    assert PS.per_atom_norm(a, b, cell) == pytest.approx(0.15)
    assert not PS.compare(a, b, cell, any_fixed=True, distance_difference=0.1)
    assert PS.compare(a, b, cell, any_fixed=True, distance_difference=0.2)
    # a rigid translation is removed only when no atom is fixed (client/Matter.cpp:142-146)
    c = a + 0.3
    assert PS.compare(a, c, cell, any_fixed=False)
    assert not PS.compare(a, c, cell, any_fixed=True)
    assert not PS.compare(a, a[:1], cell, any_fixed=True)


def test_status_rules_and_swap():
    cell = np.diag([20.0, 20.0, 20.0])
    init = np.array([[1.0, 1.0, 1.0], [4.0, 4.0, 4.0]])
    moved = init + np.array([[0.0, 0.0, 0.0], [1.5, 0.0, 0.0]])
    fixed = np.array([[1, 1, 1], [0, 0, 0]])
    flat = lambda p: p.reshape(1, -1)
    W = 6
    sad = _h([0, 0, 0, 0, 3, 0], [-9.0] * W, [flat(moved * 0.5 + init * 0.5)[0]] * W, [30] * W)
    #            good (min2 = initial -> swapped)   both moved     both initial   relax failed   climb failed   high barrier
    min1 = _h([0, 0, 0, 5, 0, 0], [-10.2, -10.2, -10.0, -10.0, -10.0, -40.0],
              [flat(moved)[0], flat(moved)[0], flat(init)[0], flat(init)[0], flat(init)[0], flat(init)[0]], [10] * W)
    min2 = _h([0, 0, 0, 0, 0, 0], [-10.0, -10.1, -10.0, -10.2, -10.2, -10.2],
              [flat(init)[0], flat(moved + 0.5)[0], flat(init + 0.01)[0], flat(moved)[0], flat(moved)[0], flat(moved)[0]],
              [12] * W)
    res = PS.process_results({"saddle": sad, "min1": min1, "min2": min2}, init, cell, fixed, max_energy=20.0)
    assert [r["termination_reason"] for r in res] == [0, PS.STATUS_BAD_NOT_CONNECTED, PS.STATUS_BAD_NOT_CONNECTED,
                                                        PS.STATUS_BAD_MINIMA, 3, PS.STATUS_BAD_HIGH_BARRIER]
    r = res[0]  # endpoints swapped: the reactant is the one that matches the initial state
    assert r["potential_energy_reactant"] == -10.0 and r["potential_energy_product"] == -10.2
    assert r["barrier_reactant_to_product"] == pytest.approx(1.0) and r["barrier_product_to_reactant"] == pytest.approx(1.2)
    assert r["total_force_calls"] == 52 and r["force_calls_saddle"] == 30 and r["force_calls_minimization"] == 22
    assert np.allclose(r["reactant"], init) and np.allclose(r["product"], moved)
    assert res[4]["force_calls_minimization"] == 0  # a failed climb ends the job before the minimisations
    text = PS.format_results_dat(r)
    lines = text.strip().split("\n")
    assert lines[0] == "0 termination_reason" and lines[1] == "good termination_reason_text"
    assert "-9.000000000000e+00 potential_energy_saddle" in lines
    assert [ln.split()[-1] for ln in lines][:9] == ["termination_reason", "termination_reason_text", "random_seed",
                                                    "potential_type", "total_force_calls", "force_calls_minimization",
                                                    "force_calls_saddle", "potential_energy_saddle",
                                                    "potential_energy_reactant"]
    # negative barrier
    sad2 = _h([0], [-10.1], [flat(init)[0]], [5])
    res2 = PS.process_results({"saddle": sad2, "min1": _h([0], [-10.0], [flat(init)[0]], [1]),
                               "min2": _h([0], [-10.2], [flat(moved)[0]], [1])}, init, cell, fixed, max_energy=20.0)
    assert res2[0]["termination_reason"] == PS.STATUS_NEGATIVE_BARRIER


def test_process_search_results_known_answer(backend, oracle):
    """data/reference/process_search_morse_pt.dat through the whole batched path: status 0, reactant / product /
    barriers of the golden, force-call bookkeeping (JobIntegrationTest.cpp:1020-1069)."""
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
    res = PS.process_results(out, d["positions"], box, fixed, max_energy=10.0,
                             displacement_positions=d["displacement_positions"].reshape(1, -1, 3),
                             potential_type="MORSE_PT")[0]
    assert res["termination_reason"] == 0
    assert res["potential_energy_saddle"] == pytest.approx(g["energy_saddle"], abs=2e-6)
    assert res["potential_energy_reactant"] == pytest.approx(g["energy_reactant"], abs=2e-6)
    assert res["potential_energy_product"] == pytest.approx(g["energy_product"], abs=2e-6)
    assert res["barrier_reactant_to_product"] == pytest.approx(g["barrier_reactant_to_product"], abs=5e-6)
    assert res["barrier_product_to_reactant"] == pytest.approx(g["barrier_product_to_reactant"], abs=5e-6)
    assert res["force_calls_saddle"] == g["force_calls_saddle"]
    assert res["total_force_calls"] <= g["force_calls_saddle"] + g["force_calls_minimization"] + 2
    assert PS.compare(d["positions"], res["reactant"], box, True) and not PS.compare(d["positions"], res["product"], box, True)
    assert 0.0 < res["displacement_saddle_distance"] < 2.0
    assert PS.format_results_dat(res).count("\n") == 16
