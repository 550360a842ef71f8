"""Host-side tail of the batched process search: Matter::compare connectivity, endpoint
swap, barrier / abort rules and the results.dat keys of client/ProcessSearchJob.cpp:234-363."""
import os

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


def test_job_directory_driver_on_the_reference_system(backend, oracle, tmp_path):
    """client/unit_tests/data/systems/saddle_search (config.ini, pos.con, displacement.con, direction.dat -- the files
    the fixture tests/golden/morse_pt_saddle.json was cut from) through `jobs.run_saddle_search_jobs`: three copies of
    the job directory in one launch, `results.dat` with the keys and order of SaddleSearchJob::saveData
    (client/SaddleSearchJob.cpp:138-163) and the known answers of data/reference/saddle_search_morse_pt.dat, plus
    `mode.dat` / `saddle.con` that read back to the harvested saddle."""
    from eon_b200 import confile, jobs
    g = oracle.goldens()["saddle_dimer_cg"]
    d = oracle.load_fixture("morse_pt_saddle.json")
    N = len(d["atomic_nrs"])
    cfg = """[Main]
job = saddle_search
temperature = 300
random_seed = 706253457

[Potential]
potential = morse_pt

[Optimizer]
converged_force = 0.001
max_iterations = 1000

[Saddle Search]
displace_least_coordinated_weight = 1.0
displace_radius = 3.3
displace_magnitude = 0.01 # Actually from the displacement.con DO NOT CHANGE
min_mode_method = dimer
max_energy = 10.0
"""
    prm = jobs.search_params_from_ini(cfg, lib=backend)
    assert prm.opt_converged_force == 0.001 and prm.saddle_converged_force == 0.001  # ParametersINI.cpp:1075-1077
    assert prm.opt_max_iterations == 1000 and prm.minmode_method == 0 and prm.saddle_max_energy == 10.0
    with pytest.raises(ValueError):
        jobs.job_settings(cfg.replace("morse_pt", "emt"))
    dirs = []
    masses = np.full(N, 195.08)
    for k in range(3):
        jd = tmp_path / ("job%d" % k)
        jd.mkdir()
        (jd / "config.ini").write_text(cfg)
        for name, pos in (("pos.con", d["positions"]), ("displacement.con", d["displacement_positions"])):
            confile.write_con(str(jd / name), confile.ConFrame(d["cell"], pos, d["atomic_nrs"], d["fixed"].astype(bool),
                                                               masses, np.arange(N)))
        confile.write_mode(str(jd / "direction.dat"), d["direction"])
        dirs.append(str(jd))
    with B200Potential("morse_pt", lib=backend) as pot:
        res = jobs.run_saddle_search_jobs(dirs, pot=pot)
    for k, r in enumerate(res):
        text = open(os.path.join(dirs[k], "results.dat")).read()
        keys = [ln.split()[-1] for ln in text.splitlines()]
        assert keys == ["termination_reason", "termination_reason_text", "job_type", "random_seed", "potential_type",
                        "total_force_calls", "force_calls_saddle", "iterations", "potential_energy_saddle",
                        "final_eigenvalue", "potential_energy_reactant"]
        assert r["termination_reason"] == 0 and r["termination_reason_text"] == "Success" and r["job_type"] == "saddle_search"
        assert r["potential_type"] == "MORSE_PT" and r["random_seed"] == 706253457
        assert r["potential_energy_saddle"] == pytest.approx(g["energy_saddle"], abs=2e-5)
        assert r["final_eigenvalue"] == pytest.approx(g["eigenvalue"], abs=5e-6)
        assert r["iterations"] == g["iterations"] and r["force_calls_saddle"] == g["force_calls_saddle"]
        assert r["total_force_calls"] == g["force_calls_saddle"] + 1
        sad = confile.read_con(os.path.join(dirs[k], "saddle.con"))
        mode = confile.read_mode(os.path.join(dirs[k], "mode.dat"), N)
        assert sad.n_atoms == N and np.array_equal(sad.fixed, d["fixed"].astype(bool))
        assert np.all(mode[d["fixed"].astype(bool)] == 0.0) and abs(np.linalg.norm(mode) - 1.0) < 1e-9
    assert np.abs(confile.read_con(os.path.join(dirs[0], "saddle.con")).positions -
                  confile.read_con(os.path.join(dirs[2], "saddle.con")).positions).max() == 0.0
    # the same three directories through two resident slots and the device-side queue: same results files
    outs = [str(tmp_path / ("queued%d" % k)) for k in range(3)]
    with B200Potential("morse_pt", lib=backend) as pot:
        res_q = jobs.run_saddle_search_jobs(dirs, out_dirs=outs, pot=pot, slots=2, groups=0)
    for k in range(3):
        assert res_q[k]["iterations"] == res[k]["iterations"] and res_q[k]["force_calls_saddle"] == res[k]["force_calls_saddle"]
        assert res_q[k]["potential_energy_saddle"] == pytest.approx(res[k]["potential_energy_saddle"], abs=1e-9)
        assert np.abs(confile.read_con(os.path.join(outs[k], "saddle.con")).positions -
                      confile.read_con(os.path.join(dirs[k], "saddle.con")).positions).max() < 1e-9
