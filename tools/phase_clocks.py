import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from eon_b200 import B200Potential
tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if "c3" in sys.argv:  # BASELINE.json configs[2] geometry, two rounds of CTAs
    from eon_b200 import workloads as WL
    struct = WL.fcc_vacancy(8, 3.615, 13)
    W = 148
    sets = np.stack([WL.jitter(struct.positions, W, 0.05, seed=1000 * k) for k in range(4)])
else:
    struct, sets = bench.force_workload(0)
    W = 256
N = struct.n_atoms
dev = torch.device("cuda", 0)
pot = B200Potential("eam", device=0, skin=1.0, tiles_override=tiles)
d_R = torch.from_numpy(sets.reshape(4, W, 3 * N)).to(dev)
d_Z = torch.from_numpy(np.broadcast_to(struct.atomic_nrs, (W, N)).copy()).to(dev)
d_box = torch.from_numpy(np.broadcast_to(struct.cell.reshape(9), (W, 9)).copy()).to(dev)
d_F = torch.empty((W, 3 * N), dtype=torch.float64, device=dev); d_U = torch.empty((W,), dtype=torch.float64, device=dev)
def step(k):
    pot.forceBatchDevice(W, N, d_R[k % 4].data_ptr(), d_Z.data_ptr(), d_F.data_ptr(), d_U.data_ptr(), d_box.data_ptr(), sync=True)
    return sum(pot.last_force_ms())
for k in range(4): step(k)
pot.phase_clocks(True)
if "hot" not in sys.argv:
    pot.l2_flush(512 << 20)  # queued on the potential's stream like in bench.py; "hot" keeps the L2 contents
ms = step(5)
c = pot.phase_clocks(False).astype(np.float64)
print("tiles", tiles, "ctas", len(c), "ms", ms)
t0 = c[:, 0].min()
start_us = (c[:, 0] - t0) / 1e3
names = ["landed", "converted", "filter(t0)", "density(t0)", "cta density", "cluster bar", "force", "exit"]
cyc = c[:, 2:11]
d = np.diff(cyc, axis=1) / 1.965e3  # us at 1965 MHz
print("entry spread us: min %.1f med %.1f max %.1f" % (start_us.min(), np.median(start_us), start_us.max()))
for k, nm in enumerate(names):
    print("%-12s med %.2f  p90 %.2f  max %.2f us" % (nm, np.median(d[:, k]), np.percentile(d[:, k], 90), d[:, k].max()))
for a, b, nm in [(2, 11, "entry -> loads issued"), (11, 3, "-> positions landed"), (4, 12, "check done -> list decision"),
                 (12, 13, "-> element row / margins"), (13, 5, "-> filter done")]:
    dd = (c[:, b] - c[:, a]) / 1.965e3
    print("  detail %-28s med %.2f max %.2f us   (first CTA of each cluster: med %.2f, others: med %.2f)" % (
        nm, np.median(dd), dd.max(), np.median(dd[0::max(1, len(c) // W)]), np.median(np.delete(dd, np.s_[0::max(1, len(c) // W)])) if len(c) > W else float("nan")))
tot = (cyc[:, -1] - cyc[:, 0]) / 1.965e3
print("cta total med %.1f max %.1f; end time (start+total) max %.1f us" % (np.median(tot), tot.max(), (start_us + tot).max()))
sm = c[:, 1].astype(int)
print("ctas per SM: min %d max %d" % (np.bincount(sm).min(), np.bincount(sm).max()))

# resident search: distribution of force calls per walker (critical path of the launch)
if "search" in sys.argv:
    from eon_b200 import Walkers, search_params, workloads as WL
    s2, scale = bench.search_workload()
    spot = B200Potential("eam", device=0, skin=1.0, conservative=True)
    wk = Walkers(spot, s2.atomic_nrs, s2.fixed, s2.cell, np.broadcast_to(s2.positions.reshape(1, -1), (W, 3 * N)).copy())
    wk.relax(search_params(opt_method=1, opt_converged_force=1e-4))
    hmin = wk.harvest()
    Rmin, E0 = hmin.positions[0], float(hmin.energies[0])
    cfg = dict(WL.AKMC_AL)
    R = np.empty((W, N, 3)); M = np.empty((W, N, 3))
    for k in range(W):
        R[k], M[k] = WL.displace(s2, Rmin, N - 1, cfg["radius"] * scale, cfg["magnitude"], cfg["seed"] + k)
    wk.set(positions=R, modes=M)
    wk.saddle_search(search_params(**bench.SEARCH_PARAMS), E0)
    h = wk.harvest(geometry=False)
    fc = np.sort(h.force_calls)
    print("search ms %.1f; force calls per walker: min %d med %d p90 %d max %d sum %d; iterations max %d" % (
        wk.last_ms, fc[0], fc[len(fc) // 2], fc[int(0.9 * len(fc))], fc[-1], fc.sum(), h.iterations.max()))
    print("us per force call on the critical path: %.1f" % (1e3 * wk.last_ms / fc[-1]))
