// nlist.cuh -- per-walker periodic cell-list neighbour build, one CTA per walker.
//
// Replaces VesinNeighbors::compute -> vesin_neighbors -> cell_list_neighbors
// (client/VesinNeighbors.cpp:42-89; client/thirdparty/vesin/vesin-single-build.cpp:1125-1274,
// 1345-1519) for the batched path.  Two final predicates share one cell walk:
//
//   NL_VESIN  the set {(i, j, S) : |(r_j - r_i) + S.H|^2 < rc^2, (i != j or S != 0)} with the
//             un-fused operation order of vesin-single-build.cpp:1199-1202 -- bit-identical to
//             vesin's full list, whatever the cell decomposition (candidate generation is free,
//             only the predicate is pinned).  Used by the LJ / Morse kernels and exported through
//             eonb200_neighbors.
//   NL_MIC    EAM.cpp's single minimum image (client/potentials/EAM/EAM.cpp:56-90,157-172):
//             coordinates shifted non-negative and wrapped, d = r_i - r_j folded once by +-L,
//             r^2 <= rc^2 (inclusive), at most one entry per (i, j).
//
// The list is cached with a Verlet skin (rc_list = rc + skin) exactly like vesin's own Verlet
// mode (vesin-single-build.cpp:1868-1975): candidates are frozen at build time, every force
// evaluation re-derives the vector from current positions and re-applies the exact predicate,
// and the list is rebuilt as soon as any atom has moved more than skin/2 from its build position.
#pragma once
#include "types.cuh"

namespace eonb200 {

struct NlParams {
  int mode;       // NL_VESIN / NL_MIC
  double rc;      // list cutoff (potential cutoff + skin)
  double rc2;     // rc * rc, rounded once on the host exactly like `cutoff * cutoff`
  double rc_core2;  // the potential's own cutoff^2: rows list these pairs first, the skin shell after them
  double rc_core;   // the potential's own cutoff
  double dstar;     // NL_MIC rows come in three classes of certainty for displacements up to dstar (= skin / 4) from
                    // the build positions: pairs that stay inside the cutoff, pairs that may cross it, pairs that
                    // stay outside (a pair distance changes by at most twice the largest displacement)
  double skin;    // 0 -> rebuild on every evaluation
  int per[3];
};

__device__ inline bool invert3(const double *m, double *inv) {
  const double det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                     m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (!(fabs(det) >= 1e-30)) return false;
  inv[0] = (m[4] * m[8] - m[7] * m[5]) / det;
  inv[1] = (m[2] * m[7] - m[1] * m[8]) / det;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) / det;
  inv[3] = (m[5] * m[6] - m[3] * m[8]) / det;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) / det;
  inv[5] = (m[3] * m[2] - m[0] * m[5]) / det;
  inv[6] = (m[3] * m[7] - m[6] * m[4]) / det;
  inv[7] = (m[6] * m[1] - m[0] * m[7]) / det;
  inv[8] = (m[0] * m[4] - m[3] * m[1]) / det;
  return true;
}

__device__ inline int floor_div(int a, int b) {
  int q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
  return q;
}

// Bring the raw positions of the walker into shared memory (sm.stage, [3N] interleaved).  A single
// bulk-copy instruction when the source is 16-byte aligned and a multiple of 16 bytes, plain
// coalesced loads otherwise.  Ends with a barrier.
__device__ inline void stage_raw_positions(const SysView &sys, Smem &sm, uint32_t &phase) {
  const int D = 3 * sys.N;
  // an odd 3N is rounded up to the next even count when the walker's row has the padding for it
  const int nd = ((D & 1) && sys.Rcap > D) ? D + 1 : D;
  const uint32_t bytes = static_cast<uint32_t>(nd) * 8u;
  const bool aligned = ((reinterpret_cast<uintptr_t>(sys.R) & 15u) == 0) && (bytes % 16u == 0);
  if (aligned) {
    if (threadIdx.x == 0) bulk_g2s(sm.stage, sys.R, bytes, sm.mbar);
    mbar_wait(sm.mbar, phase);
    phase ^= 1u;
  } else {
    for (int k = threadIdx.x; k < D; k += blockDim.x) sm.stage[k] = sys.R[k];
  }
  __syncthreads();
}

// NL_MIC image offsets: id = (fx + 1) + 3 (fy + 1) + 9 (fz + 1) -> -(fx Lx, fy Ly, fz Lz)  (EAM.cpp:157-168)
__device__ inline void fill_sh27(const Smem &sm, double Lx, double Ly, double Lz) {
  for (int k = threadIdx.x; k < 27; k += blockDim.x) {
    const double fx = static_cast<double>(k % 3 - 1), fy = static_cast<double>((k / 3) % 3 - 1), fz = static_cast<double>(k / 9 - 1);
    const double4 v = make_double4(-mul_rn(fx, Lx), -mul_rn(fy, Ly), -mul_rn(fz, Lz), 0.0);
    sm.sh27[k] = v;
    if (sm.sh27f) sm.sh27f[k] = make_float4(static_cast<float>(v.x), static_cast<float>(v.y), static_cast<float>(v.z), 0.0f);
  }
}

// Geometry of the box + positions in the layout the predicate wants (sm.pos):
//   NL_VESIN: raw coordinates.   NL_MIC: EAM.cpp:56-90 shifted / wrapped coordinates.
// Also fills the shift tables and the exp table, and returns (to every thread) the largest squared
// displacement of the prepared positions from the ones the cached list was built with.
// tables_ready: box geometry, exp table and image tables in shared memory are still valid from an earlier call with
// the same box (resident walkers evaluate hundreds of configurations of one cell): only the coordinate shift is redone.
__device__ inline double prepare_positions(const NlParams &np, const SysView &sys, Smem &sm, const double *exp_tab_g,
                                           bool tables_ready = false) {
  const int N = sys.N;
  const bool have_list = *sys.state != 0;
  double v[3] = {1e300, 1e300, 1e300};
  if (np.mode == NL_MIC) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      v[0] = fmin(v[0], sm.stage[3 * i]);
      v[1] = fmin(v[1], sm.stage[3 * i + 1]);
      v[2] = fmin(v[2], sm.stage[3 * i + 2]);
    }
    const int op[3] = {RED_MIN, RED_MIN, RED_MIN};
    block_reduce<3>(v, op, sm.red);
  }
  if (threadIdx.x == 0 && tables_ready) {
    Geom &g = *sm.geom;
    for (int k = 0; k < 3; k++) {
      const double lo = v[k] > 0.0 ? 0.0 : v[k];
      g.lo[k] = np.mode == NL_MIC ? fabs(lo) : 0.0;
    }
  }
  if (threadIdx.x == 0 && !tables_ready) {
    Geom &g = *sm.geom;
    for (int k = 0; k < 9; k++) g.H[k] = sys.box[k];
    g.ok = invert3(g.H, g.Hinv) ? 1 : 0;
    for (int k = 0; k < 3; k++) {
      const double lo = v[k] > 0.0 ? 0.0 : v[k];  // EAM.cpp:56-76: only negative minima shift
      g.lo[k] = np.mode == NL_MIC ? fabs(lo) : 0.0;
      g.per[k] = g.ok ? np.per[k] : 0;
    }
    if (np.mode == NL_MIC) {
      // stored folds stay the reference's minimum image while every box length >= 2 rc_list + skin
      const double need = 2.0 * np.rc + np.skin;
      g.tab_ok = (g.H[0] >= need && g.H[4] >= need && g.H[8] >= need) ? 1 : 0;
    } else {
      g.tab_ok = (g.H[1] == 0.0 && g.H[2] == 0.0 && g.H[3] == 0.0 && g.H[5] == 0.0 && g.H[6] == 0.0 && g.H[7] == 0.0) ? 1 : 0;
    }
  }
  if (!tables_ready)
    for (int k = threadIdx.x; k < 128; k += blockDim.x) sm.exp_tab[k] = exp_tab_g[k];
  __syncthreads();
  const Geom &g = *sm.geom;
  if (!tables_ready) {
    for (int k = threadIdx.x; k < 48; k += blockDim.x) {
      const int axis = k >> 4;
      const double sft = static_cast<double>((k & 15) - 8);
      const double len = g.H[4 * axis];
      // NL_MIC: d = r_i - r_j, fold +1 means "dx > L/2 -> dx -= L" (EAM.cpp:157-168); NL_VESIN: + s * H_kk
      const double val = np.mode == NL_MIC ? -mul_rn(sft, len) : mul_rn(sft, len);
      (axis == 0 ? sm.shx : (axis == 1 ? sm.shy : sm.shz))[k & 15] = val;
    }
    fill_sh27(sm, g.H[0], g.H[4], g.H[8]);
  }
  double disp2 = 0.0;
  const double L[3] = {g.H[0], g.H[4], g.H[8]};  // EAM.cpp:31 reads the diagonal only
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double p[3];
    for (int k = 0; k < 3; k++) {
      p[k] = sm.stage[3 * i + k];
      if (np.mode == NL_MIC) {
        p[k] += g.lo[k];  // EAM.cpp:78-83
        if (L[k] > 0.0) {
          int guard = 0;
          while (p[k] > L[k] && guard++ < 1000000) p[k] -= L[k];  // EAM.cpp:86-90
        }
      }
    }
    sm.pos[i] = make_double4(p[0], p[1], p[2], 0.0);
    if (have_list) {
      const double dx = p[0] - sys.rbuild[3 * i], dy = p[1] - sys.rbuild[3 * i + 1], dz = p[2] - sys.rbuild[3 * i + 2];
      disp2 = fmax(disp2, dx * dx + dy * dy + dz * dz);
    }
  }
  return block_max1(disp2, sm.red);
}

// Choose the cell grid (thread 0), given the list cutoff.  Cells are at least rc_eff wide so that
// +-ns cells cover the cutoff; rc_eff carries a 1e-9 relative margin against mis-binning by
// rounding in the fractional coordinates.
__device__ inline void choose_cells(const NlParams &np, const SysView &sys, Smem &sm, const double *fmin,
                                    const double *fmax) {
  Geom &g = *sm.geom;
  const double rc_eff = np.rc * (1.0 + 1e-9) + 1e-12;
  double width[3];
  if (np.mode == NL_MIC) {
    width[0] = fabs(g.H[0]);
    width[1] = fabs(g.H[4]);
    width[2] = fabs(g.H[8]);
  } else {
    // face distances, vesin-single-build.cpp:99-128
    const double *a = g.H, *b = g.H + 3, *c = g.H + 6;
    const double bc[3] = {b[1] * c[2] - b[2] * c[1], b[2] * c[0] - b[0] * c[2], b[0] * c[1] - b[1] * c[0]};
    const double ca[3] = {c[1] * a[2] - c[2] * a[1], c[2] * a[0] - c[0] * a[2], c[0] * a[1] - c[1] * a[0]};
    const double ab[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const double nbc = sqrt(bc[0] * bc[0] + bc[1] * bc[1] + bc[2] * bc[2]);
    const double nca = sqrt(ca[0] * ca[0] + ca[1] * ca[1] + ca[2] * ca[2]);
    const double nab = sqrt(ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2]);
    width[0] = nbc > 0 ? fabs(bc[0] * a[0] + bc[1] * a[1] + bc[2] * a[2]) / nbc : 0.0;
    width[1] = nca > 0 ? fabs(ca[0] * b[0] + ca[1] * b[1] + ca[2] * b[2]) / nca : 0.0;
    width[2] = nab > 0 ? fabs(ab[0] * c[0] + ab[1] * c[1] + ab[2] * c[2]) / nab : 0.0;
  }
  double extent[3];
  for (int k = 0; k < 3; k++) {
    if (np.mode == NL_MIC || g.per[k]) {
      g.fmin[k] = 0.0;
      g.fspan[k] = 1.0;
      extent[k] = width[k];
    } else {
      // open axis: bin the occupied fractional range only
      g.fmin[k] = fmin[k];
      g.fspan[k] = fmax[k] - fmin[k];
      if (!(g.fspan[k] > 0.0)) g.fspan[k] = 1.0;
      extent[k] = g.fspan[k] * width[k];
    }
    double q = extent[k] / rc_eff;
    int n = (q >= 1.0 && q < 1e6) ? static_cast<int>(q) : (q >= 1e6 ? 1000000 : 1);
    if (n < 1) n = 1;
    g.nc[k] = n;
  }
  // cap the grid: never more cells than shared memory holds, nor many more than atoms
  int cap = sm.maxCells;
  for (;;) {
    const long long tot = static_cast<long long>(g.nc[0]) * g.nc[1] * g.nc[2];
    if (tot <= cap) break;
    int kmax = 0;
    for (int k = 1; k < 3; k++)
      if (g.nc[k] > g.nc[kmax]) kmax = k;
    const double shrink = cbrt(static_cast<double>(cap) / static_cast<double>(tot));
    int n = static_cast<int>(g.nc[kmax] * shrink);
    if (n >= g.nc[kmax]) n = g.nc[kmax] - 1;
    if (n < 1) n = 1;
    g.nc[kmax] = n;
  }
  g.ncells = g.nc[0] * g.nc[1] * g.nc[2];
  for (int k = 0; k < 3; k++) {
    const double cellw = extent[k] / g.nc[k];
    int ns = cellw > 0.0 ? static_cast<int>(ceil(rc_eff / cellw)) : 1;
    if (ns < 1) ns = 1;
    if (!(np.mode == NL_MIC || g.per[k]) && ns > g.nc[k] - 1) ns = g.nc[k] > 1 ? g.nc[k] - 1 : 0;
    g.ns[k] = ns;
  }
}

// Build this walker's list from sm.pos / sm.stage.  Returns (to all threads) the largest row fill.
__device__ inline int nl_build(const NlParams &np, const SysView &sys, Smem &sm) {
  const int N = sys.N;
  const int tid = threadIdx.x, nt = blockDim.x;
  Geom &g = *sm.geom;

  // fractional range for open axes (VESIN mode)
  double fr[6] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300};
  const bool any_open = np.mode == NL_VESIN && !(np.per[0] && np.per[1] && np.per[2]);
  if (any_open && g.ok) {
    for (int i = tid; i < N; i += nt) {
      const double4 p = sm.pos[i];
      for (int k = 0; k < 3; k++) {
        const double f = p.x * g.Hinv[k] + p.y * g.Hinv[3 + k] + p.z * g.Hinv[6 + k];
        fr[k] = fmin(fr[k], f);
        fr[3 + k] = fmax(fr[3 + k], f);
      }
    }
    const int op[6] = {RED_MIN, RED_MIN, RED_MIN, RED_MAX, RED_MAX, RED_MAX};
    block_reduce<6>(fr, op, sm.red);
  }
  if (tid == 0) choose_cells(np, sys, sm, fr, fr + 3);
  __syncthreads();
  const int ncells = g.ncells;
  for (int c = tid; c <= ncells; c += nt) {
    sm.cell_start[c] = 0;
    sm.cell_fill[c] = 0;
  }
  __syncthreads();

  // 1. bin atoms
  for (int i = tid; i < N; i += nt) {
    const double4 p = sm.pos[i];
    int c[3], w[3] = {0, 0, 0};
    if (np.mode == NL_MIC) {
      const double L[3] = {g.H[0], g.H[4], g.H[8]};
      const double q[3] = {p.x, p.y, p.z};
      for (int k = 0; k < 3; k++) {
        int ck = L[k] > 0.0 ? static_cast<int>(q[k] / L[k] * g.nc[k]) : 0;
        c[k] = min(max(ck, 0), g.nc[k] - 1);
      }
    } else {
      for (int k = 0; k < 3; k++) {
        double f = g.ok ? p.x * g.Hinv[k] + p.y * g.Hinv[3 + k] + p.z * g.Hinv[6 + k] : 0.0;
        if (g.per[k]) {
          const double fl = floor(f);
          w[k] = static_cast<int>(fl);
          f -= fl;
        } else {
          f = (f - g.fmin[k]) / g.fspan[k];
        }
        int ck = static_cast<int>(f * g.nc[k]);
        c[k] = min(max(ck, 0), g.nc[k] - 1);
      }
    }
    const int cell = (c[0] * g.nc[1] + c[1]) * g.nc[2] + c[2];
    sm.cell_of[i] = cell;
    // wrap shifts packed as 3 x 10 bits, biased by 512
    sm.wrapc[i] = ((w[0] + 512) & 1023) | (((w[1] + 512) & 1023) << 10) | (((w[2] + 512) & 1023) << 20);
    atomicAdd(&sm.cell_fill[cell], 1);
  }
  __syncthreads();

  // 2. exclusive scan of the cell counts: up to 32 lanes scan contiguous chunks, thread 0 joins them
  {
    const int nl = nt < 32 ? nt : 32;
    const int chunk = (ncells + nl - 1) / nl;
    int *lane_tot = reinterpret_cast<int *>(sm.red);
    if (tid < nl) {
      int s = 0;
      const int c0 = tid * chunk, c1 = min(ncells, c0 + chunk);
      for (int c = c0; c < c1; c++) {
        const int n = sm.cell_fill[c];
        sm.cell_start[c] = s;
        s += n;
      }
      lane_tot[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
      int s = 0;
      for (int l = 0; l < nl; l++) {
        const int n = lane_tot[l];
        lane_tot[l] = s;
        s += n;
      }
    }
    __syncthreads();
    if (tid < nl) {
      const int c0 = tid * chunk, c1 = min(ncells, c0 + chunk);
      const int base = lane_tot[tid];
      for (int c = c0; c < c1; c++) sm.cell_start[c] += base;
    }
    if (tid == 0) sm.cell_start[ncells] = N;
    __syncthreads();
    for (int c = tid; c < ncells; c += nt) sm.cell_fill[c] = 0;
    __syncthreads();
  }

  // 3. scatter atoms into their cells, then order each cell by atom index (deterministic rows)
  for (int i = tid; i < N; i += nt) {
    const int cell = sm.cell_of[i];
    const int slot = atomicAdd(&sm.cell_fill[cell], 1);
    sm.order[sm.cell_start[cell] + slot] = i;
  }
  __syncthreads();
  for (int c = tid; c < ncells; c += nt) {
    const int b = sm.cell_start[c], e = sm.cell_start[c + 1];
    for (int a = b + 1; a < e; a++) {
      const int key = sm.order[a];
      int q = a - 1;
      while (q >= b && sm.order[q] > key) {
        sm.order[q + 1] = sm.order[q];
        q--;
      }
      sm.order[q + 1] = key;
    }
  }
  __syncthreads();

  // 4. walk the neighbouring cells of every atom and apply the exact predicate
  int my_max = 0;
  bool bad_shift = false;
  const int K = sys.K, Np = sys.Np;
  for (int i = tid; i < N; i += nt) {
    const double4 pi = sm.pos[i];
    const int ci = sm.cell_of[i];
    const int cz = ci % g.nc[2], cy = (ci / g.nc[2]) % g.nc[1], cx = ci / (g.nc[2] * g.nc[1]);
    const int wi = sm.wrapc[i];
    const int wix = (wi & 1023) - 512, wiy = ((wi >> 10) & 1023) - 512, wiz = ((wi >> 20) & 1023) - 512;
    // core pairs (inside the potential's cutoff) fill the row from the front, the skin shell from the back;
    // the shell is moved down behind the core afterwards, so that the lanes of a warp accept / reject together
    // NL_MIC: three classes (see NlParams::dstar) -- sure-in from the front, sure-out from the back, the uncertain
    // ones parked in this atom's column of nbr2 and moved in between afterwards
    const double r_in = np.rc_core - 2.0 * np.dstar, r_out = np.rc_core + 2.0 * np.dstar;
    const double r_in2 = r_in > 0.0 ? r_in * r_in * (1.0 - 1e-12) : -1.0, r_out2 = r_out * r_out * (1.0 + 1e-12);
    int n0 = 0, n1 = 0, n2 = 0;
    auto emit3 = [&](int code, double r2) {
      const int cls = r2 < r_in2 ? 0 : (r2 > r_out2 ? 2 : 1);
      if (n0 + n1 + n2 < K) {
        if (cls == 0)
          sys.nbr[static_cast<size_t>(n0) * Np + i] = code;
        else if (cls == 2)
          sys.nbr[static_cast<size_t>(K - 1 - n2) * Np + i] = code;
        else
          sys.nbr2[static_cast<size_t>(n1) * Np + i] = code;
      }
      if (cls == 0)
        n0++;
      else if (cls == 2)
        n2++;
      else
        n1++;
    };
    int ncore = 0, nshell = 0;
    auto emit = [&](int code, bool core) {
      if (ncore + nshell < K) {
        if (core)
          sys.nbr[static_cast<size_t>(ncore) * Np + i] = code;
        else
          sys.nbr[static_cast<size_t>(K - 1 - nshell) * Np + i] = code;
      }
      if (core)
        ncore++;
      else
        nshell++;
    };
    for (int ox = -g.ns[0]; ox <= g.ns[0]; ox++) {
      int qx = cx + ox, sx = 0;
      if (np.mode == NL_MIC || g.per[0]) {
        sx = floor_div(qx, g.nc[0]);
        qx -= sx * g.nc[0];
      } else if (qx < 0 || qx >= g.nc[0]) {
        continue;
      }
      for (int oy = -g.ns[1]; oy <= g.ns[1]; oy++) {
        int qy = cy + oy, sy = 0;
        if (np.mode == NL_MIC || g.per[1]) {
          sy = floor_div(qy, g.nc[1]);
          qy -= sy * g.nc[1];
        } else if (qy < 0 || qy >= g.nc[1]) {
          continue;
        }
        for (int oz = -g.ns[2]; oz <= g.ns[2]; oz++) {
          int qz = cz + oz, sz = 0;
          if (np.mode == NL_MIC || g.per[2]) {
            sz = floor_div(qz, g.nc[2]);
            qz -= sz * g.nc[2];
          } else if (qz < 0 || qz >= g.nc[2]) {
            continue;
          }
          const int cj = (qx * g.nc[1] + qy) * g.nc[2] + qz;
          const int b = sm.cell_start[cj], e = sm.cell_start[cj + 1];
          for (int a = b; a < e; a++) {
            const int j = sm.order[a];
            const double4 pj = sm.pos[j];
            if (np.mode == NL_MIC) {
              if (j == i) continue;
              // EAM.cpp:152-172: d = r_i - r_j, one fold by +-L; this cell image is the minimum
              // image iff its shift equals the fold the reference would apply.
              const double hx = g.H[0] * 0.5, hy = g.H[4] * 0.5, hz = g.H[8] * 0.5;
              double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
              const int fx = dx > hx ? 1 : (dx < -hx ? -1 : 0);
              const int fy = dy > hy ? 1 : (dy < -hy ? -1 : 0);
              const int fz = dz > hz ? 1 : (dz < -hz ? -1 : 0);
              if (fx != sx || fy != sy || fz != sz) continue;
              if (fx > 0) dx -= g.H[0]; else if (fx < 0) dx += g.H[0];
              if (fy > 0) dy -= g.H[4]; else if (fy < 0) dy += g.H[4];
              if (fz > 0) dz -= g.H[8]; else if (fz < 0) dz += g.H[8];
              const double r2 = norm2_rn(dx, dy, dz);
              if (!(r2 <= np.rc2)) continue;
              emit3(pack_mic(j, fx, fy, fz), r2);
            } else {
              const int wj = sm.wrapc[j];
              const int Sx = sx + wix - ((wj & 1023) - 512);
              const int Sy = sy + wiy - (((wj >> 10) & 1023) - 512);
              const int Sz = sz + wiz - (((wj >> 20) & 1023) - 512);
              if (j == i && Sx == 0 && Sy == 0 && Sz == 0) continue;
              // vesin-single-build.cpp:1199-1202: vector = (r_j - r_i) + S.H, un-fused
              const double fx = static_cast<double>(Sx), fy = static_cast<double>(Sy), fz = static_cast<double>(Sz);
              const double tx = add_rn(add_rn(mul_rn(fx, g.H[0]), mul_rn(fy, g.H[3])), mul_rn(fz, g.H[6]));
              const double ty = add_rn(add_rn(mul_rn(fx, g.H[1]), mul_rn(fy, g.H[4])), mul_rn(fz, g.H[7]));
              const double tz = add_rn(add_rn(mul_rn(fx, g.H[2]), mul_rn(fy, g.H[5])), mul_rn(fz, g.H[8]));
              const double vx = add_rn(sub_rn(pj.x, pi.x), tx);
              const double vy = add_rn(sub_rn(pj.y, pi.y), ty);
              const double vz = add_rn(sub_rn(pj.z, pi.z), tz);
              const double d2 = norm2_rn(vx, vy, vz);
              if (!(d2 < np.rc2)) continue;
              if (Sx < -8 || Sx > 7 || Sy < -8 || Sy > 7 || Sz < -8 || Sz > 7) {
                bad_shift = true;
                continue;
              }
              if (sys.nbr2 && sys.cls)
                emit3(pack_nbr(j, Sx, Sy, Sz), d2);
              else
                emit(pack_nbr(j, Sx, Sy, Sz), d2 < np.rc_core2);
            }
          }
        }
      }
    }
    if (np.mode == NL_MIC || (sys.nbr2 && sys.cls)) {
      if (n0 + n1 + n2 <= K)
        for (int t = 0; t < n1; t++)
          sys.nbr[static_cast<size_t>(n0 + t) * Np + i] = sys.nbr2[static_cast<size_t>(t) * Np + i];
      sys.cls[i] = n0 + n1 + n2 <= K ? (n0 | (n1 << 16)) : 0;  // a truncated row reads as empty (the host regrows K)
      ncore = n0 + n1;
      nshell = n2;
    }
    const int count = ncore + nshell;
    if (count <= K) {
      // fill the hole between core and shell from the top of the shell
      const int holes = K - nshell - ncore;
      const int moves = nshell < holes ? nshell : holes;
      for (int t = 0; t < moves; t++)
        sys.nbr[static_cast<size_t>(ncore + t) * Np + i] = sys.nbr[static_cast<size_t>(K - 1 - t) * Np + i];
    }
    if (np.mode == NL_MIC && count <= K) {
      // pad to a multiple of 8 with codes of the far-away ghost atom N (fused EAM filter: no bounds tests)
      for (int k = count; (k & 7) && k < K; k++) sys.nbr[static_cast<size_t>(k) * Np + i] = pack_mic(N, 0, 0, 0);
    }
    sys.cnt[i] = count <= K ? count : 0;
    my_max = max(my_max, count);
  }
  if (bad_shift) my_max = 1 << 30;  // unrepresentable image shift: reported as an overflow that never fits
  double m = block_max1(static_cast<double>(my_max), sm.red);
  const int row_max = static_cast<int>(m);
  if (tid == 0) atomicMax(sys.overflow, row_max);
  return row_max;
}

// A cached list is only good for the cell it was built in (image shifts and folds are stored per pair).
__device__ __forceinline__ bool nl_have_list(const SysView &sys) {
  bool have = ld_cg(sys.state) != 0;
  if (have && sys.box_build) {
    for (int k = 0; k < 9; k++) have = have && ld_cg(sys.box_build + k) == sys.box[k];
  }
  return have;
}

// Stage positions (sm.stage raw, sm.pos in the predicate's layout) and decide whether the cached candidate
// list still covers them (Verlet check).  CTA-uniform result; every thread has read *sys.state on return.
__device__ inline bool nl_stage_check(const NlParams &np, const SysView &sys, Smem &sm, uint32_t &phase,
                                      bool force_rebuild, const double *exp_tab_g, bool tables_ready = false) {
  stage_raw_positions(sys, sm, phase);
  const double disp2 = prepare_positions(np, sys, sm, exp_tab_g, tables_ready);
  if (threadIdx.x == 0) sm.geom->disp2 = disp2;
  const bool have = nl_have_list(sys);
  const double lim = 0.5 * np.skin;
  const bool rebuild = force_rebuild || !have || np.skin <= 0.0 || !(disp2 <= lim * lim);
  __syncthreads();  // every thread has read *sys.state before it can change
  return rebuild;
}

// Rebuild this walker's list from sm.pos and record the build positions.  Returns the largest row fill.
__device__ EONB200_NOINLINE int nl_rebuild(const NlParams &np, const SysView &sys, Smem &sm) {
  const int row_max = nl_build(np, sys, sm);
  for (int i = threadIdx.x; i < sys.N; i += blockDim.x) {
    const double4 p = sm.pos[i];
    sys.rbuild[3 * i] = p.x;
    sys.rbuild[3 * i + 1] = p.y;
    sys.rbuild[3 * i + 2] = p.z;
  }
  if (threadIdx.x == 0) {
    if (sys.box_build)
      for (int k = 0; k < 9; k++) sys.box_build[k] = sys.box[k];
    *sys.state = 1;
    atomicAdd(&sys.stats[0], 1ULL);
  }
  __syncthreads();
  return row_max;
}

// Stage positions and make sure a valid candidate list exists for them (Verlet check + rebuild).
// On return sm.pos holds the positions in the predicate's layout and sm.stage the raw ones.
// Returns (to all threads) the largest row fill of a rebuild (> sys.K: the list is truncated), 0 if reused.
__device__ inline int nl_ensure(const NlParams &np, const SysView &sys, Smem &sm, uint32_t &phase,
                                bool force_rebuild, const double *exp_tab_g, bool tables_ready = false) {
  const bool rebuild = nl_stage_check(np, sys, sm, phase, force_rebuild, exp_tab_g, tables_ready);
  int row_max = 0;
  if (rebuild) {
    if (threadIdx.x == 0) sm.geom->disp2 = 0.0;  // the fresh list is built from these very positions
    row_max = nl_rebuild(np, sys, sm);
  } else if (threadIdx.x == 0) {
    atomicAdd(&sys.stats[1], 1ULL);
  }
  return row_max;
}

}  // namespace eonb200
