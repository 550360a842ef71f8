"""End-to-end time of B200Potential.forceBatch with caller-owned contiguous numpy blocks (page-locked by the handle):
the measurement bench.py reports as `e2e`.  Development aid.   python tools/e2e_py.py [steps=K]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from eon_b200 import B200Potential  # noqa: E402

opt = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
steps = int(opt.get("steps", 100))
struct, sets = bench.force_workload(0)
W, N = sets.shape[1], struct.n_atoms
for pin in (1, 0):
    pot = B200Potential("eam", device=0, pin_host_buffers=pin)
    h_R = [np.ascontiguousarray(sets[k].reshape(W, 3 * N)) for k in range(len(sets))]
    h_F = np.empty((W, 3 * N))
    for k in range(6):
        pot.forceBatch(W, N, h_R[k % len(h_R)], struct.atomic_nrs, struct.cell.reshape(9), out=h_F)
    ts = []
    for k in range(steps):
        t0 = time.perf_counter()
        pot.forceBatch(W, N, h_R[k % len(h_R)], struct.atomic_nrs, struct.cell.reshape(9), out=h_F)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    print(f"chunks {os.environ.get('EONB200_PIPE_CHUNKS', 'auto')} pin {pin}: e2e ms/step mean {ts.mean():.4f} min {ts.min():.4f}")
    pot.close()
