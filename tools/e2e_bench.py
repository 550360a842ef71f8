"""End-to-end time of one plugin call (host buffers in, host buffers out) through the C++ shim
(eon_b200/host/B200Potential.hpp) driven by tests/host/harness: C2 shape, 256 walkers x 385 atoms, one heap block
per system.  Development aid; bench.py reports the same measurement in its `e2e` object.

    python tools/e2e_bench.py [steps=K] [chunks=1,2,4,8] [device=0]
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eon_b200 import _lib, workloads as WL  # noqa: E402

opt = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
steps = int(opt.get("steps", 50))
chunks = [int(c) for c in opt.get("chunks", "1,2,4,8").split(",")]
W = int(opt.get("w", 256))


def harness_input(struct, positions):
    lines = []
    for w in range(positions.shape[0]):
        lines.append(" ".join(repr(float(x)) for x in positions[w].reshape(-1)))
        lines.append(" ".join(str(int(z)) for z in struct.atomic_nrs))
        lines.append(" ".join(repr(float(x)) for x in struct.cell.reshape(-1)))
    return "\n".join(lines) + "\n"


s = WL.load_structure("akmc_al.json")
R = WL.jitter(s.positions, W, 0.05, seed=0)
text = harness_input(s, R)
exe = os.path.join(ROOT, "tests", "host", "harness")
lib = _lib.library_path() if hasattr(_lib, "library_path") else os.path.join(ROOT, "eon_b200", "libeonb200.so")
for pin in ("", "pin"):
    for c in chunks:
        env = dict(os.environ, EONB200_PIPE_CHUNKS=str(c))
        args = [exe, lib, "3", str(W), str(s.n_atoms), "bench", str(steps)] + ([pin] if pin else []) + ["device=" + opt.get("device", "0")]
        out = subprocess.run(args, input=text, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        print(f"chunks {c} {pin or 'staged'}: rc {out.returncode} {out.stdout.strip()}")
        if out.stderr.strip():
            print("\n".join(out.stderr.strip().splitlines()[-int(opt.get("errlines", 8)):]))
