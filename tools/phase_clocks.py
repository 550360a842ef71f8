import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from eon_b200 import B200Potential
tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 0
struct, sets = bench.force_workload(0)
N, W = struct.n_atoms, 256
dev = torch.device("cuda", 0)
pot = B200Potential("eam", device=0, skin=1.0, tiles_override=tiles)
d_R = torch.from_numpy(sets.reshape(4, W, 3 * N)).to(dev)
d_Z = torch.from_numpy(np.broadcast_to(struct.atomic_nrs, (W, N)).copy()).to(dev)
d_box = torch.from_numpy(np.broadcast_to(struct.cell.reshape(9), (W, 9)).copy()).to(dev)
d_F = torch.empty((W, 3 * N), dtype=torch.float64, device=dev); d_U = torch.empty((W,), dtype=torch.float64, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def step(k):
    pot.forceBatchDevice(W, N, d_R[k % 4].data_ptr(), d_Z.data_ptr(), d_F.data_ptr(), d_U.data_ptr(), d_box.data_ptr(), sync=True)
    return sum(pot.last_force_ms())
for k in range(4): step(k)
pot.phase_clocks(True)
flush.zero_(); torch.cuda.synchronize()
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
tot = (cyc[:, -1] - cyc[:, 0]) / 1.965e3
print("cta total med %.1f max %.1f; end time (start+total) max %.1f us" % (np.median(tot), tot.max(), (start_us + tot).max()))
sm = c[:, 1].astype(int)
print("ctas per SM: min %d max %d" % (np.bincount(sm).min(), np.bincount(sm).max()))
