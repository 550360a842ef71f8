#!/usr/bin/env python
"""Regenerate tests/golden/*.json from the reference checkout.

Run HERE (the container that has /root/reference); the GPU box only ever sees
the committed JSON.  Structures are the reference's own test/example inputs
(.con data files, not source code); golden numbers are the known-answer values
of the reference's Catch2 / job-integration tests, each with its file:line.

    python tests/golden/make_fixtures.py [/root/reference]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from eon_b200.confile import read_con, read_mode  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SYS = os.path.join(REF, "client/unit_tests/data/systems")


def frame_dict(path, src):
    f = read_con(path)
    return {
        "source": src,
        "cell": f.cell.tolist(),
        "positions": f.positions.tolist(),
        "atomic_nrs": f.atomic_nrs.tolist(),
        "fixed": f.fixed.astype(int).tolist(),
        "masses": f.masses.tolist(),
    }


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as fh:
        json.dump(obj, fh)
        fh.write("\n")
    print("wrote", name)


def main():
    dump("lj13.json", frame_dict(os.path.join(SYS, "neb_morse/reactant.con"),
                                 "client/unit_tests/data/systems/neb_morse/reactant.con"))
    dump("lj13_product.json", frame_dict(os.path.join(SYS, "neb_morse/product.con"),
                                         "client/unit_tests/data/systems/neb_morse/product.con"))
    ss = frame_dict(os.path.join(SYS, "saddle_search/pos.con"),
                    "client/unit_tests/data/systems/saddle_search/pos.con")
    disp = read_con(os.path.join(SYS, "saddle_search/displacement.con"))
    ss["displacement_positions"] = disp.positions.tolist()
    ss["direction"] = read_mode(os.path.join(SYS, "saddle_search/direction.dat"), disp.n_atoms).tolist()
    ss["displacement_source"] = "client/unit_tests/data/systems/saddle_search/{displacement.con,direction.dat}"
    dump("morse_pt_saddle.json", ss)
    dump("al_fcc.json", frame_dict(os.path.join(SYS, "al_fcc/pos.con"),
                                   "client/unit_tests/data/systems/al_fcc/pos.con"))
    dump("akmc_al.json", frame_dict(os.path.join(REF, "examples/akmc-al/pos.con"), "examples/akmc-al/pos.con"))
    dump("akmc_pt.json", frame_dict(os.path.join(REF, "examples/akmc-pt/pos.con"), "examples/akmc-pt/pos.con"))
    neb = frame_dict(os.path.join(REF, "examples/neb-al/reactant.con"), "examples/neb-al/reactant.con")
    neb["product_positions"] = read_con(os.path.join(REF, "examples/neb-al/product.con")).positions.tolist()
    dump("neb_al.json", neb)

    goldens = {
        "lj_point": {"energy": -39.965351, "max_force": 0.004704, "fixture": "lj13.json",
                     "cite": "client/unit_tests/PotTest.cpp:38-43; client/unit_tests/data/reference/point_lj.dat:1-2",
                     "params": {"u0": 1.0, "psi": 1.0, "cutoff": 15.0, "shifted": True}},
        "ljcluster_point": {"energy": -39.965379, "fixture": "lj13.json",
                            "cite": "client/unit_tests/PotTest.cpp:130-132"},
        "lj_minimization": {"energy": -39.965352, "force_calls": 11,
                            "cite": "client/unit_tests/data/reference/minimization_lj.dat:4-5"},
        "lj_min_cg": {"energy": -39.965352, "force_calls": 15,
                      "cite": "client/unit_tests/data/reference/min_lj_cg.dat:4-5"},
        "lj_min_fire": {"energy": -39.965352, "force_calls": 21,
                        "cite": "client/unit_tests/data/reference/min_lj_fire.dat:4-5"},
        "morse_pt_point": {"energy": -1462.166782, "max_force": 0.003638, "fixture": "morse_pt_saddle.json",
                           "cite": "client/unit_tests/data/reference/point_morse_pt.dat:1-2",
                           "params": {"De": 0.7102, "a": 1.6047, "re": 2.8970, "cutoff": 9.5, "shifted": True}},
        "saddle_dimer_cg": {"status": 0, "energy_saddle": -1462.008706, "eigenvalue": -1.014995,
                            "iterations": 10, "total_force_calls": 39, "force_calls_saddle": 38,
                            "energy_reactant": -1462.166782,
                            "cite": "client/unit_tests/data/reference/saddle_search_morse_pt.dat:1-10; "
                                    "client/unit_tests/JobIntegrationTest.cpp:365-416"},
        "saddle_lanczos_cg": {"status": 0, "energy_saddle": -1462.008706, "eigenvalue": -1.011347,
                              "eigenvalue_tol": 1e-3, "iterations": 10, "total_force_calls": 44,
                              "cite": "client/unit_tests/data/reference/saddle_search_lanczos_morse_pt.dat:1-10; "
                                      "client/unit_tests/JobIntegrationTest.cpp:418-460"},
        "process_search_lbfgs": {"energy_saddle": -1.462008705949e+03, "energy_reactant": -1.462166782734e+03,
                                 "energy_product": -1.462179028847e+03,
                                 "barrier_reactant_to_product": 1.580767852306e-01,
                                 "barrier_product_to_reactant": 1.703228982960e-01,
                                 "force_calls_saddle": 31, "force_calls_minimization": 27,
                                 "total_force_calls": 67, "force_calls_prefactors": 9,
                                 "prefactor_reactant_to_product": 1.166188480043e+12,
                                 "prefactor_product_to_reactant": 1.130999754565e+12,
                                 "cite": "client/unit_tests/data/reference/process_search_morse_pt.dat:1-15; "
                                         "client/unit_tests/JobIntegrationTest.cpp:1020-1069"},
        "eam_al_rgpot_point": {"energy": -5.217864, "max_force": 0.968647, "fixture": "al_fcc.json",
                               "cite": "client/unit_tests/EAMAlTest.cpp:29-34 (rgpot Fortran eam_al: a DIFFERENT model "
                                       "from the in-tree EAM class; not reproducible here, listed for the record)"},
    }
    dump("goldens.json", goldens)


if __name__ == "__main__":
    main()
