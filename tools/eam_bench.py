"""Quick timing of the batched EAM kernels on the bench workloads (C2 256 x 385 atoms, C3 W x 2047 atoms):
ms per step with a cold L2, FP64 roofline fraction, per-CTA phase clocks.  Development aid; bench.py is the contract.

    python tools/eam_bench.py [c3] [mode=0|1] [threads=N] [steps=K] [clocks] [hot] [w3=W]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
import bench  # noqa: E402
from eon_b200 import B200Potential, workloads as WL  # noqa: E402

opt = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
flags = [a for a in sys.argv[1:] if "=" not in a]
mode = int(opt.get("mode", 0))
threads = int(opt.get("threads", 0))
steps = int(opt.get("steps", 20))
skin = float(opt.get("skin", 1.0))
if "c3" in flags:
    struct = WL.fcc_vacancy(8, 3.615, 13)
    W = int(opt.get("w3", 592))
    sets = np.stack([WL.jitter(struct.positions, W, 0.05, seed=1000 * k) for k in range(2)])
else:
    struct, sets = bench.force_workload(0)
    W = 256
nsets = len(sets)
N = struct.n_atoms
dev = torch.device("cuda", 0)
libpath = opt.get("lib")  # another build of the library (timing ablations)
lib = None
if libpath:
    import ctypes
    from eon_b200 import _capi
    lib = _capi.bind(ctypes.CDLL(os.path.abspath(libpath)))
rot = int(opt.get("rot", 1))  # handles cycled (each with its own walkers and lists): inputs larger than L2, no flush
pots = [B200Potential("eam", device=0, skin=skin, eam_pair_mode=mode, threads_per_walker=threads, lib=lib) for _ in range(rot)]
pot = pots[0]
d_R = torch.from_numpy(sets.reshape(nsets, W, 3 * N)).to(dev)
d_Z = torch.from_numpy(np.broadcast_to(struct.atomic_nrs, (W, N)).copy()).to(dev)
d_box = torch.from_numpy(np.broadcast_to(struct.cell.reshape(9), (W, 9)).copy()).to(dev)
d_F = torch.empty((W, 3 * N), dtype=torch.float64, device=dev)
d_U = torch.empty((W,), dtype=torch.float64, device=dev)


def step(k):
    p = pots[k % rot]
    p.forceBatchDevice(W, N, d_R[(k // rot) % nsets].data_ptr(), d_Z.data_ptr(), d_F.data_ptr(), d_U.data_ptr(),
                       d_box.data_ptr(), sync=True)
    return sum(p.last_force_ms())


for k in range(4 * rot):
    step(k)
pairs = float(pot.nl_stats()["ordered_pairs"])
peak = pot.measure_fp64_peak(5)
ts = []
for k in range(steps):
    if "hot" not in flags and rot == 1:
        pot.l2_flush(512 << 20)
    ts.append(step(k))
ts = np.array(ts)
flops = bench.eam_flops(pairs, W * N)
ms = float(np.mean(ts))
print(f"rot {rot} mode {mode} threads {threads} W {W} N {N}: ms/step mean {ms:.4f} min {ts.min():.4f} max {ts.max():.4f}; "
      f"pairs {pairs:.0f}; {flops / (ms * 1e-3) / 1e12:.2f} TFLOP/s of {peak / 1e12:.2f} -> frac {flops / (ms * 1e-3) / peak:.3f}; "
      f"{W * N / (ms * 1e-3):.3e} atom-force-evals/s; rebuilt {pot.nl_stats()['rebuilt']}")
if "clocks" in flags:
    pot.phase_clocks(True)
    if "hot" not in flags:
        pot.l2_flush(512 << 20)
    ms1 = step(1)
    c = pot.phase_clocks(False).astype(np.float64)
    t0 = c[:, 0].min()
    start_us = (c[:, 0] - t0) / 1e3
    print("clocked launch ms", ms1, "ctas", len(c), "entry spread us: med %.1f max %.1f" % (np.median(start_us), start_us.max()))
    if mode == 0:
        seq = [(2, "entry"), (3, "positions landed"), (4, "converted"), (12, "list decision"),
               (6, "density done (t0)"), (7, "density barrier"), (8, "embedding done"), (9, "own force + barrier"), (10, "recv force, exit")]
        prev = None
        for slot, nm in seq:
            if prev is not None:
                d = (c[:, slot] - c[:, prev]) / 1.965e3
                print("  %-22s med %.2f p90 %.2f max %.2f us" % (nm, np.median(d), np.percentile(d, 90), d.max()))
            prev = slot
        for a_, b_, nm in [(12, 5, "last warp leaves density"), (8, 11, "last warp leaves own force"), (9, 15, "thread 0 leaves recv force"),
                           (9, 13, "last warp leaves recv force"), (13, 14, "energy reduction"), (14, 10, "exit")]:
            d = (c[:, b_] - c[:, a_]) / 1.965e3
            print("  [%-28s] med %.2f p90 %.2f max %.2f us" % (nm, np.median(d), np.percentile(d, 90), d.max()))
        tot = (c[:, 10] - c[:, 2]) / 1.965e3
        print("  cta total med %.1f max %.1f; end (start + total) max %.1f us" % (np.median(tot), tot.max(), (start_us + tot).max()))
    smid = c[:, 1].astype(int)
    print("  ctas per SM: min %d max %d" % (np.bincount(smid)[np.bincount(smid) > 0].min(), np.bincount(smid).max()))
