// search.cuh -- device-resident min-mode saddle search and relaxation, one CTA per walker.
//
// Every function here is executed by all threads of the walker's CTA with CTA-uniform control
// flow; vectors live in the walker's HBM workspace (L2 resident), scalars come out of fixed-order
// block reductions.  No host round trip happens inside a search.
//
// Reference semantics followed (paths relative to the eOn checkout):
//   Matter force cache / wrap / mask / net force   client/Matter.cpp:25-72,273-357,498-538; include/eon/Matter.h:42-63
//   MinModeObjectiveFunction                        client/MinModeSaddleSearch.cpp:51-164
//   MinModeSaddleSearch::run                        client/MinModeSaddleSearch.cpp:211-475
//   ImprovedDimer::compute                          client/ImprovedDimer.cpp:63-434
//   Lanczos::compute                                client/Lanczos.cpp:46-173; client/MobileAtoms.cpp:47-135
//   ConjugateGradients                              client/ConjugateGradients.cpp:17-43,114-190
//   LBFGS                                           client/LBFGS.cpp:19-165
//   FIRE                                            client/FIRE.cpp:19-68; include/eon/FIRE.h:22-35
//   relaxMatter / MatterObjectiveFunction           client/HelperFunctions.cpp:273-410
//   maxAtomMotion*, SafeMath                        client/GeometryAnalysis.cpp:269-327; include/eon/SafeMath.h:19-80
#pragma once
#include "../../include/eonb200.h"
#include "potentials.cuh"

namespace eonb200 {

#define VLOOP(k, n) for (int k = threadIdx.x; k < (n); k += blockDim.x)

constexpr int kMaxLbfgsMemory = 64;
constexpr int kMaxLanczosIters = 24;
constexpr int kMaxRotHistory = 32;

// Workspace vector slots (each Dp doubles per walker).
enum Slot {
  S_X1_POS = 0, S_X1_FRC, S_X1P_POS, S_X1P_FRC,            // dimer images / Lanczos probe
  S_THETA, S_THETA_OLD, S_FR, S_FR_OLD, S_G1, S_TAUP, S_TAU_OLD,  // ImprovedDimer
  S_GRAD,                                                  // objective gradient
  S_A, S_B, S_C, S_D, S_E, S_F,                            // optimizer vectors
  S_LZ_R, S_LZ_U, S_LZ_EV,                                 // Lanczos
  S_COUNT
};

struct WalkerArrays {
  int W, N, D, Dp, nfree, any_fixed;
  double *pos;   // [W][Dp] positions of the walker's Matter (wrapped)
  double *frc;   // [W][Dp] its masked forces (valid when dirty == 0)
  double *mode;  // [W][Dp] eigenvector in / out
  double *energy;  // [W]
  int *dirty;      // [W]
  double *ws;      // [W][ws_stride]
  size_t ws_stride;
  size_t off_lbfgs_s, off_lbfgs_y, off_lanczos_q, off_rot_s, off_rot_y;  // history blocks inside ws
  const double *mask;     // [Dp] 1.0 free / 0.0 fixed (Matter::getFree)
  const int *free_atoms;  // [nfree] atoms with at least one free axis
  const int *Z;           // [N]
  const double *box;      // [9]
  const double *reactant_energy;  // [W] or null
  int *status;
  long long *iterations;
  long long *force_calls;
  double *eigenvalue;
  double *convergence;
};

struct MatterS {
  double *pos;
  double *frc;
  double energy;
  int dirty;
};

struct Ctx {
  int N, D, Dp, nfree, Dfree, any_fixed, cell_ok, lpa;
  const double *mask;
  const int *free_atoms;
  const double *cell;  // shared memory copy
  const double *inv;
  const PotParams *P;
  NlParams np;
  SysView sys;
  Smem sm;
  uint32_t phase;
  const EonB200SearchParams *prm;  // shared memory copy
  long long force_calls;
  int tables_ready;  // shared-memory box geometry / exp / image tables are valid (set by the first evaluation)
  int row_flags;     // eam_assign_rows result of this walker's (constant) atomic numbers, -1 = not computed yet
  int failed;  // potential failure (NaN energy / truncated neighbour list)
  int remove_net_force;
  double *ws;
  const WalkerArrays *wa;
  __device__ double *slot(int s) const { return ws + static_cast<size_t>(s) * Dp; }
};

// ---- Matter -----------------------------------------------------------------------------------------
// Legacy wrap of Matter::setPositions (client/Matter.cpp:273-280, include/eon/Matter.h:54-63):
// frac = r.inv; frac = fmod(frac + 1, 1); r = frac.cell
__device__ __forceinline__ void wrap_atom(const Ctx &c, double &x, double &y, double &z) {
  if (!c.cell_ok) return;
  double f[3];
#pragma unroll
  for (int k = 0; k < 3; k++) f[k] = x * c.inv[k] + y * c.inv[3 + k] + z * c.inv[6 + k];
#pragma unroll
  for (int k = 0; k < 3; k++) f[k] = fmod(f[k] + 1.0, 1.0);
  x = f[0] * c.cell[0] + f[1] * c.cell[3] + f[2] * c.cell[6];
  y = f[0] * c.cell[1] + f[1] * c.cell[4] + f[2] * c.cell[7];
  z = f[0] * c.cell[2] + f[1] * c.cell[5] + f[2] * c.cell[8];
}

// m.pos = wrap(src + alpha * dir) over all 3N coordinates; dir may be null.  src may alias m.pos.
__device__ inline void matter_set_positions(const Ctx &c, MatterS &m, const double *src, const double *dir,
                                            double alpha) {
  VLOOP(i, c.N) {
    double x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
    if (dir) {
      x += alpha * dir[3 * i];
      y += alpha * dir[3 * i + 1];
      z += alpha * dir[3 * i + 2];
    }
    wrap_atom(c, x, y, z);
    m.pos[3 * i] = x;
    m.pos[3 * i + 1] = y;
    m.pos[3 * i + 2] = z;
  }
  m.dirty = 1;
  __syncthreads();
}

// Matter::setPositionsFreeV: packed free-atom coordinates xf (+ alpha * df) scattered, then wrap all.
__device__ inline void matter_set_positions_free(const Ctx &c, MatterS &m, const double *xf, const double *df,
                                                 double alpha) {
  VLOOP(a, c.nfree) {
    const int i = c.free_atoms[a];
    for (int k = 0; k < 3; k++) m.pos[3 * i + k] = xf[3 * a + k] + (df ? alpha * df[3 * a + k] : 0.0);
  }
  __syncthreads();
  VLOOP(i, c.N) {
    double x = m.pos[3 * i], y = m.pos[3 * i + 1], z = m.pos[3 * i + 2];
    wrap_atom(c, x, y, z);
    m.pos[3 * i] = x;
    m.pos[3 * i + 1] = y;
    m.pos[3 * i + 2] = z;
  }
  m.dirty = 1;
  __syncthreads();
}

// Matter::computePotential (client/Matter.cpp:498-538) + the mask of getForces (:324-333).
__device__ EONB200_NOINLINE void matter_compute(Ctx &c, MatterS &m) {
  if (!m.dirty) return;
  c.sys.R = m.pos;
  // (box geometry / tables and the element rows of the walker are computed on its first evaluation only)
  if (nl_ensure(c.np, c.sys, c.sm, c.phase, false, c.P->exp_tab, c.tables_ready != 0) > c.sys.K) c.failed = 1;  // truncated list
  c.tables_ready = 1;
  double U = walker_forces(*c.P, c.sys, c.sm, c.lpa, &c.row_flags);
  if (U != U) c.failed = 1;
  c.force_calls++;
  m.energy = U;
  m.dirty = 0;
  double mean[3] = {0.0, 0.0, 0.0};
  if (!c.any_fixed && c.remove_net_force) {  // Matter.cpp:531-536
    VLOOP(i, c.N) {
      mean[0] += c.sm.stage[3 * i];
      mean[1] += c.sm.stage[3 * i + 1];
      mean[2] += c.sm.stage[3 * i + 2];
    }
    block_sum<3>(mean, c.sm.red);
    for (int k = 0; k < 3; k++) mean[k] /= static_cast<double>(c.N);
  }
  VLOOP(i, c.N) {
#pragma unroll
    for (int k = 0; k < 3; k++) m.frc[3 * i + k] = (c.sm.stage[3 * i + k] - mean[k]) * c.mask[3 * i + k];
  }
  __syncthreads();
}

__device__ inline void matter_copy(const Ctx &c, MatterS &dst, const MatterS &src) {
  VLOOP(k, c.D) {
    dst.pos[k] = src.pos[k];
    dst.frc[k] = src.frc[k];
  }
  dst.energy = src.energy;
  dst.dirty = src.dirty;
  __syncthreads();
}

// ---- small vector helpers -----------------------------------------------------------------------------
__device__ inline double vdot(const Ctx &c, const double *a, const double *b, int n) {
  double s = 0.0;
  VLOOP(k, n) s += a[k] * b[k];
  return block_sum1(s, c.sm.red);
}
__device__ inline void vcopy(double *dst, const double *src, int n) {
  VLOOP(k, n) dst[k] = src[k];
  __syncthreads();
}
__device__ inline void vzero(double *dst, int n) {
  VLOOP(k, n) dst[k] = 0.0;
  __syncthreads();
}
// GeometryAnalysis.cpp:273-296
__device__ inline double max_atom_motion(const Ctx &c, const double *v, int n) {
  double mx = 0.0;
  VLOOP(i, n / 3) {
    const double nn = sqrt(v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
    mx = fmax(mx, nn);
  }
  return block_max1(mx, c.sm.red);
}
// GeometryAnalysis.cpp:321-327, in place: v *= max_motion / mx when mx > max_motion
__device__ inline void max_atom_motion_apply(const Ctx &c, double *v, int n, double max_motion) {
  const double mx = max_atom_motion(c, v, n);
  if (mx > max_motion) {
    const double s = max_motion / mx;
    VLOOP(k, n) v[k] *= s;
    __syncthreads();
  }
}
// pbc::apply on a packed difference (include/eon/Matter.h:42-50): out = mic(a - b)
__device__ inline void pbc_difference(const Ctx &c, const double *a, const double *b, double *out, int n) {
  VLOOP(i, n / 3) {
    double d[3] = {a[3 * i] - b[3 * i], a[3 * i + 1] - b[3 * i + 1], a[3 * i + 2] - b[3 * i + 2]};
    if (c.cell_ok) {
      double f[3];
      for (int k = 0; k < 3; k++) f[k] = d[0] * c.inv[k] + d[1] * c.inv[3 + k] + d[2] * c.inv[6 + k];
      for (int k = 0; k < 3; k++) f[k] -= floor(f[k] + 0.5);
      for (int k = 0; k < 3; k++) d[k] = f[0] * c.cell[k] + f[1] * c.cell[3 + k] + f[2] * c.cell[6 + k];
    }
    out[3 * i] = d[0];
    out[3 * i + 1] = d[1];
    out[3 * i + 2] = d[2];
  }
  __syncthreads();
}

// Convergence metric of the true (masked) force (MinModeSaddleSearch.cpp:150-164,
// HelperFunctions.cpp:296-308).  Fully fixed atoms carry zero force, so sums over all atoms equal
// the reference's sums over free atoms.
__device__ inline double convergence_of(Ctx &c, MatterS &m) {
  matter_compute(c, m);
  const int metric = c.prm->convergence_metric;
  if (metric == EONB200_METRIC_MAX_ATOM) return max_atom_motion(c, m.frc, c.D);
  if (metric == EONB200_METRIC_MAX_COMPONENT) {
    double mx = -1e300;
    VLOOP(k, c.D) mx = fmax(mx, m.frc[k]);  // AtomMatrix::maxCoeff (signed), as in the reference
    return block_max1(mx, c.sm.red);
  }
  return sqrt(vdot(c, m.frc, m.frc, c.D));
}

// ---- lowest eigenmode: ImprovedDimer -----------------------------------------------------------------
struct MinModeState {
  // ImprovedDimer
  int init_cg, init_lbfgs, rot_hist;
  double C_tau;
  // Lanczos
  double lowestEw;
  long long rotations_total;
  int stats_rotations;
};

__device__ inline double minmode_eigenvalue(const Ctx &c, const MinModeState &s) {
  return c.prm->minmode_method == EONB200_MINMODE_LANCZOS ? s.lowestEw : s.C_tau;
}

// tau lives in the walker's mode array (it is the eigenvector handed back to the objective).
__device__ EONB200_NOINLINE void dimer_compute(Ctx &c, MinModeState &st, MatterS &x0, double *tau) {
  const EonB200SearchParams &prm = *c.prm;
  const int D = c.D;
  double *theta = c.slot(S_THETA), *thetaOld = c.slot(S_THETA_OLD), *FR = c.slot(S_FR), *FRold = c.slot(S_FR_OLD);
  double *g1 = c.slot(S_G1), *taup = c.slot(S_TAUP), *tauOld = c.slot(S_TAU_OLD);
  MatterS x1{c.slot(S_X1_POS), c.slot(S_X1_FRC), 0.0, 1};
  MatterS x1p{c.slot(S_X1P_POS), c.slot(S_X1P_FRC), 0.0, 1};
  const double delta = prm.finite_difference;

  // tau = direction .* free; normalise (ImprovedDimer.cpp:68-86)
  {
    double s = 0.0;
    VLOOP(k, D) {
      const double t = tau[k] * c.mask[k];
      tau[k] = t;
      s += t * t;
    }
    double nrm = sqrt(block_sum1(s, c.sm.red));
    if (!(nrm > 1e-10)) {
      // the reference draws a random tangent; same deterministic stand-in as the oracle
      s = 0.0;
      VLOOP(k, D) {
        const unsigned long long h = static_cast<unsigned long long>(k) * 2654435761ULL % 1000ULL;
        const double t = c.mask[k] * (static_cast<double>(h) / 1000.0 - 0.5);
        tau[k] = t;
        s += t * t;
      }
      nrm = sqrt(block_sum1(s, c.sm.red));
    }
    if (nrm < kSafeEps) {
      VLOOP(k, D) tau[k] = 0.0;
    } else {
      VLOOP(k, D) tau[k] /= nrm;
    }
    __syncthreads();
  }

  // x0 is the caller's Matter (never modified here); x1 = x0 + delta tau (ImprovedDimer.cpp:89-103)
  matter_compute(c, x0);
  matter_set_positions(c, x1, x0.pos, tau, delta);
  matter_compute(c, x1);
  if (x1.energy - x0.energy > 10.0 * delta) {
    VLOOP(k, D) tau[k] = -tau[k];
    __syncthreads();
    matter_set_positions(c, x1, x0.pos, tau, delta);
  }
  if (prm.dimer_opt_method == EONB200_ROT_LBFGS) {
    st.rot_hist = 0;
    st.init_lbfgs = 1;
  }
  const double phi_tol = kPi * (prm.dimer_converged_angle / 180.0);
  double phi_prime = 0.0, phi_min = 0.0;
  st.stats_rotations = 0;
  matter_compute(c, x1);
  const double *F0 = x0.frc;  // g0 = -F0
  VLOOP(k, D) g1[k] = -x1.frc[k];
  __syncthreads();

  do {
    // dg = g1 - g0 = g1 + F0;  F_R = -2 dg + 2 (dg.tau) tau   (ImprovedDimer.cpp:224)
    double dgtau;
    {
      double s = 0.0;
      VLOOP(k, D) s += (g1[k] + F0[k]) * tau[k];
      dgtau = block_sum1(s, c.sm.red);
    }
    double r3[3] = {0.0, 0.0, 0.0};  // F_R.F_R_Old, F_R_Old.F_R_Old, F_R.(F_R - F_R_Old)
    VLOOP(k, D) {
      const double dg = g1[k] + F0[k];
      const double fr = -2.0 * dg + 2.0 * dgtau * tau[k];
      FR[k] = fr;
      const double fo = FRold[k];
      r3[0] += fr * fo;
      r3[1] += fo * fo;
      r3[2] += fr * (fr - fo);
    }
    block_sum<3>(r3, c.sm.red);

    if (prm.dimer_opt_method == EONB200_ROT_SD) {
      const double n2 = vdot(c, FR, FR, D);
      const double nrm = sqrt(n2);
      const double sc = nrm < kSafeEps ? 0.0 : 1.0 / nrm;
      VLOOP(k, D) theta[k] = FR[k] * sc;
      __syncthreads();
    } else if (prm.dimer_opt_method == EONB200_ROT_CG) {
      double gamma = 0.0;
      if (st.init_cg) {
        st.init_cg = 0;
      } else {
        const double a = fabs(r3[0]), b = r3[1];
        gamma = (a < 0.5 * b) ? safe_div(r3[2], b, 0.0) : 0.0;
      }
      double s = 0.0;
      VLOOP(k, D) {
        const double t = (gamma == 0.0) ? FR[k] : FR[k] + gamma * thetaOld[k];
        theta[k] = t;
        s += t * tau[k];
      }
      const double tt = block_sum1(s, c.sm.red);
      s = 0.0;
      VLOOP(k, D) {
        const double t = theta[k] - tt * tau[k];
        theta[k] = t;
        thetaOld[k] = t;
        s += t * t;
      }
      double nrm = sqrt(block_sum1(s, c.sm.red));
      if (nrm < kSafeEps) {
        const double ft = vdot(c, FR, tau, D);
        s = 0.0;
        VLOOP(k, D) {
          const double t = FR[k] - ft * tau[k];
          theta[k] = t;
          s += t * t;
        }
        nrm = sqrt(block_sum1(s, c.sm.red));
      }
      if (nrm < kSafeEps) {
        VLOOP(k, D) theta[k] = 0.0;
      } else {
        VLOOP(k, D) theta[k] /= nrm;
      }
      VLOOP(k, D) FRold[k] = FR[k];
      __syncthreads();
    } else {
      // L-BFGS rotation plane (ImprovedDimer.cpp:252-299)
      double *ls = c.ws + c.wa->off_rot_s, *ly = c.ws + c.wa->off_rot_y;
      __shared__ double lrho[kMaxRotHistory], lalpha[kMaxRotHistory];
      if (!st.init_lbfgs) {
        if (st.rot_hist < kMaxRotHistory) {
          double *s0 = ls + static_cast<size_t>(st.rot_hist) * c.Dp, *y0 = ly + static_cast<size_t>(st.rot_hist) * c.Dp;
          const double invd = safe_div(1.0, delta, 0.0);
          double s = 0.0;
          VLOOP(k, D) {
            const double sv = tau[k] - tauOld[k];
            const double yv = (FRold[k] - FR[k]) * invd;
            s0[k] = sv;
            y0[k] = yv;
            s += sv * yv;
          }
          const double sy = block_sum1(s, c.sm.red);
          if (threadIdx.x == 0) lrho[st.rot_hist] = safe_recip(sy, 0.0);
          st.rot_hist++;
          __syncthreads();
        }
      } else {
        st.init_lbfgs = 0;
      }
      const double H0 = 1.0 / 60.0;
      const int loopmax = st.rot_hist;
      double *q = theta;  // q / z accumulate in theta's storage
      VLOOP(k, D) q[k] = -FR[k];
      __syncthreads();
      for (int i = loopmax - 1; i >= 0; i--) {
        const double *si = ls + static_cast<size_t>(i) * c.Dp, *yi = ly + static_cast<size_t>(i) * c.Dp;
        const double a = lrho[i] * vdot(c, si, q, D);
        if (threadIdx.x == 0) lalpha[i] = a;
        VLOOP(k, D) q[k] -= a * yi[k];
        __syncthreads();
      }
      VLOOP(k, D) q[k] *= H0;
      __syncthreads();
      for (int i = 0; i < loopmax; i++) {
        const double *si = ls + static_cast<size_t>(i) * c.Dp, *yi = ly + static_cast<size_t>(i) * c.Dp;
        const double bv = lrho[i] * vdot(c, yi, q, D);
        const double a = lalpha[i];
        VLOOP(k, D) q[k] += (a - bv) * si[k];
        __syncthreads();
      }
      double r[3] = {0.0, 0.0, 0.0};  // z.z, FR.FR, z.FR
      VLOOP(k, D) {
        r[0] += q[k] * q[k];
        r[1] += FR[k] * FR[k];
        r[2] += q[k] * FR[k];
      }
      block_sum<3>(r, c.sm.red);
      const double nz = sqrt(r[0]), nf = sqrt(r[1]);
      const double cz = nz < kSafeEps ? 0.0 : 1.0 / nz, cf = nf < kSafeEps ? 0.0 : 1.0 / nf;
      const double vd = fmin(1.0, fmax(-1.0, -(r[2] * cz * cf)));
      const double angle = safe_acos(vd) * (180.0 / kPi);
      double zn = cz;
      if (angle > 87.0) {
        st.rot_hist = 0;
        VLOOP(k, D) q[k] = -FR[k];
        __syncthreads();
        zn = cf;
      }
      // theta = -normalized(z); theta -= (theta.tau) tau; normalise
      double s = 0.0;
      VLOOP(k, D) {
        const double t = -(q[k] * zn);
        theta[k] = t;
        s += t * tau[k];
      }
      const double tt = block_sum1(s, c.sm.red);
      s = 0.0;
      VLOOP(k, D) {
        const double t = theta[k] - tt * tau[k];
        theta[k] = t;
        s += t * t;
      }
      const double nrm = sqrt(block_sum1(s, c.sm.red));
      VLOOP(k, D) {
        const double t = nrm < kSafeEps ? 0.0 : theta[k] / nrm;
        theta[k] = t;
        thetaOld[k] = t;
        FRold[k] = FR[k];
        tauOld[k] = tau[k];
      }
      __syncthreads();
    }

    // curvature and its derivative (ImprovedDimer.cpp:302-318)
    st.C_tau = safe_div(dgtau, delta, 0.0);
    double dgtheta;
    {
      double s = 0.0;
      VLOOP(k, D) s += (g1[k] + F0[k]) * theta[k];
      dgtheta = block_sum1(s, c.sm.red);
    }
    const double dC = 2.0 * safe_div(dgtheta, delta, 0.0);
    phi_prime = -0.5 * safe_atan_ratio(dC, 2.0 * fabs(st.C_tau), 0.0);

    if (fabs(phi_prime) > phi_tol) {
      const double b1 = 0.5 * dC;
      // trial rotation (ImprovedDimer.cpp:323-340)
      {
        const double cp = cos(phi_prime), sp = sin(phi_prime);
        double s = 0.0;
        VLOOP(k, D) {
          const double t = tau[k] * cp + theta[k] * sp;
          taup[k] = t;
          s += t * t;
        }
        const double nrm = sqrt(block_sum1(s, c.sm.red));
        const double sc = nrm < kSafeEps ? 0.0 : 1.0 / nrm;
        VLOOP(k, D) taup[k] *= sc;
        __syncthreads();
      }
      matter_set_positions(c, x1p, x0.pos, taup, delta);
      matter_compute(c, x1p);
      double ctp;
      {
        double s = 0.0;
        VLOOP(k, D) s += (-x1p.frc[k] + F0[k]) * taup[k];
        ctp = block_sum1(s, c.sm.red);
      }
      const double C_tau_prime = safe_div(ctp, delta, 0.0);
      // Fourier coefficients and the optimal angle (ImprovedDimer.cpp:342-363)
      const double a1 =
          safe_div(st.C_tau - C_tau_prime + b1 * sin(2.0 * phi_prime), 1.0 - cos(2.0 * phi_prime), 0.0);
      const double a0 = 2.0 * (st.C_tau - a1);
      phi_min = 0.5 * safe_atan_ratio(b1, a1, 0.0);
      double C_tau_min = 0.5 * a0 + a1 * cos(2.0 * phi_min) + b1 * sin(2.0 * phi_min);
      if (C_tau_min > st.C_tau) {
        phi_min += kPi * 0.5;
        C_tau_min = 0.5 * a0 + a1 * cos(2.0 * phi_min) + b1 * sin(2.0 * phi_min);
      }
      if (phi_min > kPi * 0.5) phi_min -= kPi;
      {
        const double cm = cos(phi_min), sn = sin(phi_min);
        double s = 0.0;
        VLOOP(k, D) {
          const double t = tau[k] * cm + theta[k] * sn;
          tau[k] = t;
          s += t * t;
        }
        const double nrm = sqrt(block_sum1(s, c.sm.red));
        const double sc = nrm < kSafeEps ? 0.0 : 1.0 / nrm;
        VLOOP(k, D) tau[k] *= sc;
        __syncthreads();
      }
      // (x1 is re-placed at x0 + delta tau in the reference; its cache is never read again)
      st.C_tau = C_tau_min;
      // g1 interpolation (ImprovedDimer.cpp:385-390)
      const double sin_pp = sin(phi_prime);
      const double c1 = safe_div(sin(phi_prime - phi_min), sin_pp, 0.0);
      const double c2 = safe_div(sin(phi_min), sin_pp, 0.0);
      const double c3 = 1.0 - cos(phi_min) - sin(phi_min) * tan(phi_prime * 0.5);
      VLOOP(k, D) g1[k] = g1[k] * c1 + (-x1p.frc[k]) * c2 + (-F0[k]) * c3;
      __syncthreads();
      st.stats_rotations += 1;
    }
  } while (fabs(phi_prime) > fabs(phi_tol) && fabs(phi_min) > fabs(phi_tol) &&
           st.stats_rotations < prm.dimer_rotations_max && !c.failed);
  st.rotations_total += st.stats_rotations;
}

// ---- lowest eigenmode: Lanczos --------------------------------------------------------------------------
// Lowest eigenpair of the symmetric TRIDIAGONAL Lanczos matrix (stands in for Eigen's SelfAdjointEigenSolver,
// Lanczos.cpp:117-119), by the first warp of the CTA: the eigenvalue by 32-way multisection on the Sturm count (a
// dozen rounds reach the last bit of the Gershgorin interval), the eigenvector by three steps of inverse iteration with
// a pivoted tridiagonal solve.  The cyclic Jacobi this replaces cost one thread ~10^5 dependent shared-memory
// operations per Lanczos iteration -- thirty times the force evaluation next to it.
// d[i] = T[i][i], e[i] = T[i][i + 1] read from T (row stride ld); work: 5 * n doubles of shared memory.
__device__ inline void tridiag_lowest(int n, const double *T, int ld, double *work, double *ew, double *evec) {
#ifdef EONB200_EMU
  const int lane = 0, nl = 1;
#else
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x, nl = 32;
#endif
  double lo = 1e300, hi = -1e300;
  for (int i = 0; i < n; i++) {
    const double dl = i > 0 ? fabs(T[(i - 1) * ld + i]) : 0.0, dr = i + 1 < n ? fabs(T[i * ld + i + 1]) : 0.0;
    lo = fmin(lo, T[i * ld + i] - dl - dr);
    hi = fmax(hi, T[i * ld + i] + dl + dr);
  }
  const double span0 = fmax(hi - lo, 1e-300);
  lo -= 1e-12 * span0;
  hi += 1e-12 * span0;
  for (int round = 0; round < 64; round++) {
    if (!(hi - lo > 4e-16 * fmax(fabs(lo), fabs(hi)) + 1e-300)) break;
    const double x = lo + (hi - lo) * static_cast<double>(lane + 1) / static_cast<double>(nl + 1);
    // Sturm count: eigenvalues below x
    int cnt = 0;
    double q = T[0] - x;
    cnt += q < 0.0 ? 1 : 0;
    for (int i = 1; i < n; i++) {
      if (fabs(q) < 1e-300) q = q < 0.0 ? -1e-300 : 1e-300;
      const double e = T[(i - 1) * ld + i];
      q = T[i * ld + i] - x - e * e / q;
      cnt += q < 0.0 ? 1 : 0;
    }
#ifdef EONB200_EMU
    if (cnt >= 1) hi = x; else lo = x;
#else
    const unsigned below = __ballot_sync(0xffffffffu, cnt >= 1);  // lanes whose point lies above the lowest eigenvalue
    const int first = below ? __ffs(below) - 1 : 32;
    const double xs_hi = lo + (hi - lo) * static_cast<double>(first + 1) / static_cast<double>(nl + 1);
    const double xs_lo = lo + (hi - lo) * static_cast<double>(first) / static_cast<double>(nl + 1);
    hi = first < 32 ? xs_hi : hi;
    lo = first > 0 ? xs_lo : lo;
#endif
  }
  const double lam = 0.5 * (lo + hi);
  if (lane == 0) {
    *ew = lam;
    // inverse iteration on (T - lam) with Gaussian elimination and partial pivoting (band grows to two superdiagonals)
    double *dd = work, *du = work + n, *du2 = work + 2 * n, *dl = work + 3 * n, *x = work + 4 * n;
    const double shift = lam - 1e-14 * fmax(fabs(lam), span0 * 1e-3);  // keep the shifted matrix non-singular
    for (int i = 0; i < n; i++) x[i] = 1.0 / sqrt(static_cast<double>(n)) * ((i & 1) ? 0.9 : 1.0);
    for (int it = 0; it < 3; it++) {
      for (int i = 0; i < n; i++) {
        dd[i] = T[i * ld + i] - shift;
        du[i] = i + 1 < n ? T[i * ld + i + 1] : 0.0;
        dl[i] = du[i];
        du2[i] = 0.0;
      }
      for (int i = 0; i + 1 < n; i++) {  // eliminate dl[i]
        if (fabs(dd[i]) >= fabs(dl[i])) {
          const double piv = fabs(dd[i]) < 1e-300 ? 1e-300 : dd[i];
          const double f = dl[i] / piv;
          dd[i + 1] -= f * du[i];
          x[i + 1] -= f * x[i];
          dd[i] = piv;
        } else {  // swap rows i and i + 1
          const double f = dd[i] / dl[i];
          const double t0 = dd[i + 1], t1 = i + 2 < n ? du[i + 1] : 0.0, tx = x[i + 1];
          dd[i] = dl[i];
          dd[i + 1] = du[i] - f * t0;
          du2[i] = t1;
          if (i + 2 < n) du[i + 1] = -f * t1;
          du[i] = t0;
          x[i + 1] = x[i] - f * tx;
          x[i] = tx;
        }
      }
      for (int i = n - 1; i >= 0; i--) {  // back substitution
        double t = x[i];
        if (i + 1 < n) t -= du[i] * x[i + 1];
        if (i + 2 < n) t -= du2[i] * x[i + 2];
        const double piv = fabs(dd[i]) < 1e-300 ? 1e-300 : dd[i];
        x[i] = t / piv;
      }
      double nn = 0.0;
      for (int i = 0; i < n; i++) nn += x[i] * x[i];
      nn = 1.0 / sqrt(nn);
      for (int i = 0; i < n; i++) x[i] *= nn;
    }
    for (int i = 0; i < n; i++) evec[i] = x[i];
  }
}

// Lanczos::compute with phva_atoms = All (mobile set = atoms that are not fully fixed).  `mode`
// holds the initial direction on entry and the lowest eigenvector (zero on fixed atoms) on exit.
__device__ EONB200_NOINLINE void lanczos_compute(Ctx &c, MinModeState &st, MatterS &matter, double *mode) {
  const EonB200SearchParams &prm = *c.prm;
  const int size = c.Dfree;
  st.stats_rotations = 0;
  if (size == 0) {
    st.lowestEw = 0.0;
    vzero(mode, c.D);
    return;
  }
  int maxIters = static_cast<int>(prm.lanczos_max_iterations);
  if (maxIters > kMaxLanczosIters) maxIters = kMaxLanczosIters;
  if (maxIters < 1) maxIters = 1;
  __shared__ double T[kMaxLanczosIters * kMaxLanczosIters];
  __shared__ double twork[5 * kMaxLanczosIters];
  __shared__ double evec[kMaxLanczosIters];
  __shared__ double ew_sh;
  VLOOP(k, maxIters * maxIters) T[k] = 0.0;
  double *r = c.slot(S_LZ_R), *u = c.slot(S_LZ_U), *evEst = c.slot(S_LZ_EV);
  double *Q = c.ws + c.wa->off_lanczos_q;
  const size_t qs = static_cast<size_t>(c.Dp);
  MatterS tmp{c.slot(S_X1P_POS), c.slot(S_X1P_FRC), 0.0, 1};
  // r = pack(direction)
  double beta;
  {
    double s = 0.0;
    VLOOP(a, c.nfree) {
      const int i = c.free_atoms[a];
      for (int k = 0; k < 3; k++) {
        const double t = mode[3 * i + k];
        r[3 * a + k] = t;
        s += t * t;
      }
    }
    beta = sqrt(block_sum1(s, c.sm.red));
  }
  if (beta < kSafeEps) {
    st.lowestEw = 0.0;
    vzero(mode, c.D);
    return;
  }
  double ew = 0.0, ewOld = 0.0, alpha = 0.0;
  const double dr = prm.finite_difference;
  bool have_ev = false;
  matter_compute(c, matter);  // force0 = cached centre forces
  const double *pos0 = matter.pos, *F0 = matter.frc;

  for (int i = 0; i < size; i++) {
    st.stats_rotations = i;
    double *Qi = Q + qs * i;
    {
      const double sc = 1.0 / beta;
      VLOOP(k, size) Qi[k] = r[k] * sc;
      __syncthreads();
    }
    // u = H q ~ -(F(x0 + dr q) - F(x0)) / dr on the mobile coordinates (Lanczos.cpp:79-84)
    VLOOP(k, c.D) tmp.pos[k] = pos0[k];
    __syncthreads();
    VLOOP(a, c.nfree) {
      const int at = c.free_atoms[a];
      for (int k = 0; k < 3; k++) tmp.pos[3 * at + k] = pos0[3 * at + k] + dr * Qi[3 * a + k];
    }
    __syncthreads();
    matter_set_positions(c, tmp, tmp.pos, nullptr, 0.0);
    matter_compute(c, tmp);
    VLOOP(a, c.nfree) {
      const int at = c.free_atoms[a];
      for (int k = 0; k < 3; k++) u[3 * a + k] = -(tmp.frc[3 * at + k] - F0[3 * at + k]) / dr;
    }
    __syncthreads();
    {
      const double *Qm = i > 0 ? Q + qs * (i - 1) : nullptr;
      double s = 0.0;
      VLOOP(k, size) {
        const double t = (i == 0) ? u[k] : u[k] - beta * Qm[k];
        r[k] = t;
        s += Qi[k] * t;
      }
      alpha = block_sum1(s, c.sm.red);
      s = 0.0;
      VLOOP(k, size) {
        const double t = r[k] - alpha * Qi[k];
        r[k] = t;
        s += t * t;
      }
      if (threadIdx.x == 0) {
        T[i * maxIters + i] = alpha;
        if (i > 0) {
          T[(i - 1) * maxIters + i] = beta;
          T[i * maxIters + (i - 1)] = beta;
        }
      }
      beta = sqrt(block_sum1(s, c.sm.red));
    }
    if (beta <= 1e-10 * fabs(alpha)) {
      if (i == 0) {
        ew = alpha;
        vcopy(evEst, Q, size);
        have_ev = true;
      }
      break;
    }
    if (i >= 1) {
      const int m = i + 1;
      __syncthreads();  // T[i][i], T[i - 1][i] written by thread 0
      tridiag_lowest(m, T, maxIters, twork, &ew_sh, evec);
      __syncthreads();
      ew = ew_sh;
      const double err = safe_div(fabs(ew - ewOld), fabs(ewOld), 1.0);
      ewOld = ew;
      double s = 0.0;
      VLOOP(k, size) {
        double t = 0.0;
        for (int a = 0; a < m; a++) t += Q[qs * a + k] * evec[a];
        evEst[k] = t;
        s += t * t;
      }
      const double nn = sqrt(block_sum1(s, c.sm.red));
      VLOOP(k, size) evEst[k] /= nn;
      __syncthreads();
      have_ev = true;
      if (err < prm.lanczos_tolerance) break;
    } else {
      ew = alpha;
      ewOld = ew;
      vcopy(evEst, Q, size);
      have_ev = true;
      if (st.lowestEw != 0.0 && prm.lanczos_quit_early) {
        const double Cprev = st.lowestEw;
        const double Cnew = vdot(c, u, Qi, size);
        const double err = safe_div(fabs(Cnew - Cprev), fabs(Cprev), 1.0);
        if (err <= prm.lanczos_tolerance) break;
      }
    }
    if (i >= maxIters - 1) break;
    if (c.failed) break;
  }
  st.rotations_total += st.stats_rotations;
  st.lowestEw = ew;
  VLOOP(k, c.D) mode[k] = 0.0;
  __syncthreads();
  if (have_ev) {
    VLOOP(a, c.nfree) {
      const int at = c.free_atoms[a];
      for (int k = 0; k < 3; k++) mode[3 * at + k] = evEst[3 * a + k];
    }
    __syncthreads();
  }
}

// ---- objective functions -----------------------------------------------------------------------------------
enum ObjKind { OBJ_MINMODE = 0, OBJ_MATTER = 1 };

struct Objective {
  int kind;
  int dof;          // 3N (min-mode) or 3 * nfree (relaxation)
  MatterS *matter;
  double *mode;     // eigenvector (min-mode objective)
  MinModeState *mm;
  int iteration;    // MinModeObjectiveFunction::iteration
};

__device__ inline void obj_get_positions(const Ctx &c, const Objective &o, double *out) {
  if (o.kind == OBJ_MINMODE) {
    VLOOP(k, c.D) out[k] = o.matter->pos[k];
  } else {
    VLOOP(a, c.nfree) {
      const int i = c.free_atoms[a];
      for (int k = 0; k < 3; k++) out[3 * a + k] = o.matter->pos[3 * i + k];
    }
  }
  __syncthreads();
}
// positions = x + alpha * dir (objective space)
__device__ inline void obj_set_positions(const Ctx &c, Objective &o, const double *x, const double *dir, double alpha) {
  if (o.kind == OBJ_MINMODE)
    matter_set_positions(c, *o.matter, x, dir, alpha);
  else
    matter_set_positions_free(c, *o.matter, x, dir, alpha);
}
__device__ inline double obj_energy(Ctx &c, Objective &o) {
  matter_compute(c, *o.matter);
  return o.matter->energy;
}
__device__ inline double obj_convergence(Ctx &c, Objective &o) { return convergence_of(c, *o.matter); }
__device__ inline bool obj_converged(Ctx &c, Objective &o) {
  const EonB200SearchParams &p = *c.prm;
  double tol = p.opt_converged_force;
  if (o.kind == OBJ_MINMODE && p.saddle_converged_force > 0.0) tol = p.saddle_converged_force;
  return obj_convergence(c, o) < tol;
}

// out = gradient (minus the effective force).  MinModeSaddleSearch.cpp:51-140 / HelperFunctions.cpp:281-289.
__device__ EONB200_NOINLINE void obj_gradient(Ctx &c, Objective &o, double *out, bool fdstep) {
  MatterS &m = *o.matter;
  matter_compute(c, m);
  if (o.kind == OBJ_MATTER) {
    VLOOP(a, c.nfree) {
      const int i = c.free_atoms[a];
      for (int k = 0; k < 3; k++) out[3 * a + k] = -m.frc[3 * i + k];
    }
    __syncthreads();
    return;
  }
  if (!fdstep || o.iteration == 0) {
    if (c.prm->minmode_method == EONB200_MINMODE_LANCZOS)
      lanczos_compute(c, *o.mm, m, o.mode);
    else
      dimer_compute(c, *o.mm, m, o.mode);
    o.iteration++;
  }
  const double eigenvalue = minmode_eigenvalue(c, *o.mm);
  double r[2] = {0.0, 0.0};
  VLOOP(k, c.D) {
    r[0] += m.frc[k] * o.mode[k];
    r[1] += o.mode[k] * o.mode[k];
  }
  block_sum<2>(r, c.sm.red);
  const double nrm = sqrt(r[1]);
  const double sc = nrm < kSafeEps ? 0.0 : 1.0 / nrm;
  const double fdot = r[0];
  const double d = c.prm->saddle_perp_force_ratio;
  VLOOP(k, c.D) {
    const double proj = (o.mode[k] * sc) * fdot;
    double f;
    if (eigenvalue > 0.0) {
      f = d > 0.0 ? d * m.frc[k] - (1.0 + d) * proj : -proj;
    } else {
      f = m.frc[k] - 2.0 * proj;
    }
    out[k] = -f;
  }
  __syncthreads();
}

// ---- optimizers ------------------------------------------------------------------------------------------------
struct OptState {
  long long iteration;
  // CG
  long long cg_i;
  // LBFGS
  int hist, hist_begin;  // ring buffer fill / oldest entry
  // FIRE
  double dt, alpha;
  long long fire_N;
};

__device__ inline void opt_init(const Ctx &c, OptState &s, const Objective &o) {
  s.iteration = 0;
  s.cg_i = 0;
  s.hist = 0;
  s.hist_begin = 0;
  s.dt = c.prm->time_step;
  s.alpha = 0.1;
  s.fire_N = 0;
  // CG: forceOld / directionOld start at zero; FIRE: velocity zero
  double *fo = c.slot(S_B), *dirOld = c.slot(S_C);
  VLOOP(k, o.dof) {
    fo[k] = 0.0;
    dirOld[k] = 0.0;
  }
  __syncthreads();
}

// ConjugateGradients::single_step + getStep.  Returns 1 when converged afterwards.
__device__ EONB200_NOINLINE int cg_step(Ctx &c, OptState &s, Objective &o, double maxMove) {
  const EonB200SearchParams &prm = *c.prm;
  const int n = o.dof;
  double *force = c.slot(S_A), *forceOld = c.slot(S_B), *dirOld = c.slot(S_C), *dirN = c.slot(S_D);
  double *pos = c.slot(S_E), *grad = c.slot(S_GRAD);
  obj_gradient(c, o, grad, false);
  obj_get_positions(c, o, pos);
  // getStep (ConjugateGradients.cpp:17-43)
  double r[3] = {0.0, 0.0, 0.0};  // f.fOld, fOld.fOld, f.(f - fOld)
  VLOOP(k, n) {
    const double f = -grad[k];
    force[k] = f;
    const double fo = forceOld[k];
    r[0] += f * fo;
    r[1] += fo * fo;
    r[2] += f * (f - fo);
  }
  block_sum<3>(r, c.sm.red);
  double gamma = 0.0;
  if (fabs(r[0]) < 0.5 * r[1]) gamma = safe_div(r[2], r[1], 0.0);
  double s2 = 0.0;
  VLOOP(k, n) {
    const double dnew = force[k] + gamma * dirOld[k];
    dirOld[k] = dnew;  // m_directionOld = m_direction
    forceOld[k] = force[k];
    s2 += dnew * dnew;
  }
  const double nrm = sqrt(block_sum1(s2, c.sm.red));
  if (nrm < kSafeEps) {
    VLOOP(k, n) dirN[k] = 0.0;
  } else {
    VLOOP(k, n) dirN[k] = dirOld[k] / nrm;
  }
  __syncthreads();
  if (prm.cg_max_iter_before_reset > 0 && prm.cg_max_iter_before_reset <= s.cg_i) {
    s.cg_i = 0;
    VLOOP(k, n) {
      forceOld[k] = 0.0;
      dirOld[k] = 0.0;
    }
    __syncthreads();
  }
  s.cg_i += 1;
  // finite-difference curvature along the direction (ConjugateGradients.cpp:114-155)
  obj_set_positions(c, o, pos, dirN, prm.finite_difference);
  obj_gradient(c, o, grad, true);
  double p2[2] = {0.0, 0.0};
  VLOOP(k, n) {
    p2[0] += force[k] * dirN[k];
    p2[1] += -grad[k] * dirN[k];
  }
  block_sum<2>(p2, c.sm.red);
  const double pf1 = p2[0];
  double pf2 = p2[1];
  const double curvature = safe_div(pf1 - pf2, prm.finite_difference, 0.0);
  double stepSize = maxMove;
  if (curvature > 0.0) stepSize = safe_div(pf1, curvature, maxMove);
  double *stepv = c.slot(S_F);
  if (!prm.cg_no_overshooting) {
    VLOOP(k, n) stepv[k] = dirN[k] * stepSize;
    __syncthreads();
    max_atom_motion_apply(c, stepv, n, maxMove);
    VLOOP(k, n) pos[k] += stepv[k];
    __syncthreads();
    obj_set_positions(c, o, pos, nullptr, 0.0);
  } else {
    double passedMinimum = -1.0;
    while (passedMinimum < 0.0 && 0.1 * fabs(pf1) < fabs(pf2) && !c.failed) {
      VLOOP(k, n) stepv[k] = dirN[k] * stepSize;
      __syncthreads();
      max_atom_motion_apply(c, stepv, n, maxMove);
      obj_set_positions(c, o, pos, stepv, 1.0);
      obj_gradient(c, o, grad, true);
      double sdot = 0.0;
      VLOOP(k, n) sdot += -grad[k] * dirN[k];
      pf2 = block_sum1(sdot, c.sm.red);
      passedMinimum = pf1 * pf2;
      if (passedMinimum < 0.0 && 0.1 * fabs(pf1) < fabs(pf2)) {
        const double forceChange = pf1 - pf2;
        stepSize = safe_div(pf1, forceChange, 0.0) * stepSize;
      }
    }
  }
  if (prm.cg_knock_out_max_move && stepSize >= maxMove) {
    VLOOP(k, n) {
      dirOld[k] = 0.0;
      forceOld[k] = 0.0;
    }
    __syncthreads();
  }
  s.iteration++;
  return obj_converged(c, o) ? 1 : 0;
}

// LBFGS::step / getStep / update / reset (client/LBFGS.cpp:19-165).  The (s, y, rho) history is a
// ring buffer in the walker's workspace.
__device__ EONB200_NOINLINE int lbfgs_step(Ctx &c, OptState &s, Objective &o, double maxMove) {
  const EonB200SearchParams &prm = *c.prm;
  const int n = o.dof;
  int memory = static_cast<int>(prm.lbfgs_memory < o.dof ? prm.lbfgs_memory : o.dof);
  if (memory > kMaxLbfgsMemory) memory = kMaxLbfgsMemory;
  if (memory < 1) memory = 1;
  double *r = c.slot(S_A), *f = c.slot(S_B), *rPrev = c.slot(S_C), *fPrev = c.slot(S_D);
  double *q = c.slot(S_E), *tmp = c.slot(S_F), *grad = c.slot(S_GRAD);
  double *hs = c.ws + c.wa->off_lbfgs_s, *hy = c.ws + c.wa->off_lbfgs_y;
  const size_t hstride = static_cast<size_t>(c.Dp);
  __shared__ double rho[kMaxLbfgsMemory], alph[kMaxLbfgsMemory];

  obj_get_positions(c, o, r);
  obj_gradient(c, o, grad, false);
  VLOOP(k, n) f[k] = -grad[k];
  __syncthreads();

  bool have_step = false;  // the step lands in q
  double H0 = prm.lbfgs_inverse_curvature;
  if (s.iteration > 0) {
    // update(r, rPrev, f, fPrev): s0 = mic(r - rPrev), y0 = fPrev - f
    pbc_difference(c, r, rPrev, tmp, n);
    double sy = 0.0;
    VLOOP(k, n) sy += tmp[k] * (fPrev[k] - f[k]);
    sy = block_sum1(sy, c.sm.red);
    if (fabs(sy) < 1e-30) {  // LBFGS_EPS
      s.hist = 0;
      s.hist_begin = 0;
    } else {
      int slot;
      if (s.hist < memory) {
        slot = (s.hist_begin + s.hist) % memory;
        s.hist++;
      } else {
        slot = s.hist_begin;
        s.hist_begin = (s.hist_begin + 1) % memory;
      }
      double *s0 = hs + hstride * slot, *y0 = hy + hstride * slot;
      VLOOP(k, n) {
        s0[k] = tmp[k];
        y0[k] = fPrev[k] - f[k];
      }
      if (threadIdx.x == 0) rho[slot] = safe_recip(sy, 0.0);
      __syncthreads();
    }
    // getStep: curvature along the last move (LBFGS.cpp:23-40); r is the current position
    double cc[2] = {0.0, 0.0};
    VLOOP(k, n) {
      const double df = fPrev[k] - f[k];
      cc[0] += df * df;
      cc[1] += tmp[k] * df;
    }
    block_sum<2>(cc, c.sm.red);
    const double C = safe_div(cc[0], cc[1], -1.0);
    if (C < 0) {
      s.hist = 0;
      s.hist_begin = 0;
      VLOOP(k, n) q[k] = f[k] * 1000.0;
      __syncthreads();
      max_atom_motion_apply(c, q, n, maxMove);
      have_step = true;
    } else if (prm.lbfgs_auto_scale) {
      H0 = safe_recip(C, -1.0);
    }
  }
  if (!have_step && s.iteration == 0 && prm.lbfgs_auto_scale) {
    // finite-difference curvature along the force (LBFGS.cpp:42-61)
    const double fn = sqrt(vdot(c, f, f, n));
    const double sc = fn < kSafeEps ? 0.0 : 1.0 / fn;
    VLOOP(k, n) tmp[k] = f[k] * sc;
    __syncthreads();
    obj_set_positions(c, o, r, tmp, prm.finite_difference);
    obj_gradient(c, o, grad, true);
    double sd = 0.0;
    VLOOP(k, n) sd += (grad[k] + f[k]) * tmp[k];
    const double C = block_sum1(sd, c.sm.red) / prm.finite_difference;
    H0 = safe_recip(C, -1.0);
    obj_set_positions(c, o, r, nullptr, 0.0);
    if (H0 < 0) {
      s.hist = 0;
      s.hist_begin = 0;
      VLOOP(k, n) q[k] = f[k] * 1000.0;
      __syncthreads();
      max_atom_motion_apply(c, q, n, maxMove);
      have_step = true;
    }
  }
  constexpr int kQr = 4;  // elements of q a thread keeps in registers during the two-loop recursion
  if (!have_step) {
    if (n <= kQr * static_cast<int>(blockDim.x)) {  // (bodies below keep their original indentation)
    // Two-loop recursion with q in registers: element k = tid + u * blockDim of every vector belongs to one thread,
    // so each history pair costs one pass over (s_i, y_i) and one block reduction instead of three passes over q in
    // the workspace and three barriers.  Same products, same summation order as the general version below.
    const int loopmax = s.hist;
    const int tid = threadIdx.x, nt = blockDim.x;
    double qr[kQr];
#pragma unroll
    for (int u = 0; u < kQr; u++) {
      const int k = tid + u * nt;
      qr[u] = k < n ? -f[k] : 0.0;
    }
    for (int i = loopmax - 1; i >= 0; i--) {
      const int slot = (s.hist_begin + i) % memory;
      const double *si = hs + hstride * slot, *yi = hy + hstride * slot;
      double sv[kQr], yv[kQr];
#pragma unroll
      for (int u = 0; u < kQr; u++) {
        const int k = tid + u * nt;
        sv[u] = k < n ? si[k] : 0.0;
        yv[u] = k < n ? yi[k] : 0.0;
      }
      double part = 0.0;
#pragma unroll
      for (int u = 0; u < kQr; u++)
        if (tid + u * nt < n) part += sv[u] * qr[u];
      const double a = rho[slot] * block_sum1(part, c.sm.red);
      if (tid == 0) alph[slot] = a;
#pragma unroll
      for (int u = 0; u < kQr; u++) qr[u] -= a * yv[u];
    }
#pragma unroll
    for (int u = 0; u < kQr; u++) qr[u] *= H0;
    for (int i = 0; i < loopmax; i++) {
      const int slot = (s.hist_begin + i) % memory;
      const double *si = hs + hstride * slot, *yi = hy + hstride * slot;
      double sv[kQr], yv[kQr];
#pragma unroll
      for (int u = 0; u < kQr; u++) {
        const int k = tid + u * nt;
        sv[u] = k < n ? si[k] : 0.0;
        yv[u] = k < n ? yi[k] : 0.0;
      }
      double part = 0.0;
#pragma unroll
      for (int u = 0; u < kQr; u++)
        if (tid + u * nt < n) part += yv[u] * qr[u];
      const double b = rho[slot] * block_sum1(part, c.sm.red);
      const double a = alph[slot];
#pragma unroll
      for (int u = 0; u < kQr; u++) qr[u] += (a - b) * sv[u];
    }
#pragma unroll
    for (int u = 0; u < kQr; u++) {
      const int k = tid + u * nt;
      if (k < n) q[k] = qr[u];
    }
    __syncthreads();
    } else {
    const int loopmax = s.hist;
    VLOOP(k, n) q[k] = -f[k];
    __syncthreads();
    for (int i = loopmax - 1; i >= 0; i--) {
      const int slot = (s.hist_begin + i) % memory;
      const double *si = hs + hstride * slot, *yi = hy + hstride * slot;
      const double a = rho[slot] * vdot(c, si, q, n);
      if (threadIdx.x == 0) alph[slot] = a;
      VLOOP(k, n) q[k] -= a * yi[k];
      __syncthreads();
    }
    VLOOP(k, n) q[k] *= H0;
    __syncthreads();
    for (int i = 0; i < loopmax; i++) {
      const int slot = (s.hist_begin + i) % memory;
      const double *si = hs + hstride * slot, *yi = hy + hstride * slot;
      const double b = rho[slot] * vdot(c, yi, q, n);
      const double a = alph[slot];
      VLOOP(k, n) q[k] += (a - b) * si[k];
      __syncthreads();
    }
    }
    // d = -z
    double rr[3] = {0.0, 0.0, 0.0};  // d.d, f.f, d.f
    VLOOP(k, n) {
      const double d = -q[k];
      q[k] = d;
      rr[0] += d * d;
      rr[1] += f[k] * f[k];
      rr[2] += d * f[k];
    }
    block_sum<3>(rr, c.sm.red);
    const double distance = max_atom_motion(c, q, n);
    bool reset_sd = false;
    if (distance >= maxMove && prm.lbfgs_distance_reset) {
      reset_sd = true;
    } else {
      const double nd = sqrt(rr[0]), nf = sqrt(rr[1]);
      const double cd = nd < kSafeEps ? 0.0 : 1.0 / nd, cf = nf < kSafeEps ? 0.0 : 1.0 / nf;
      const double vd = fmin(1.0, fmax(-1.0, rr[2] * cd * cf));
      const double angle = safe_acos(vd) * (180.0 / kPi);
      if (angle > 90.0 && prm.lbfgs_angle_reset) reset_sd = true;
    }
    if (reset_sd) {
      s.hist = 0;
      s.hist_begin = 0;
      VLOOP(k, n) q[k] = f[k] * H0;
      __syncthreads();
    }
    max_atom_motion_apply(c, q, n, maxMove);
  }
  // objf->setPositions(r + dr); rPrev = r; fPrev = f
  obj_set_positions(c, o, r, q, 1.0);
  VLOOP(k, n) {
    rPrev[k] = r[k];
    fPrev[k] = f[k];
  }
  __syncthreads();
  s.iteration++;
  return obj_converged(c, o) ? 1 : 0;
}

// FIRE::step (client/FIRE.cpp:19-68).  Returns -1 when the time step collapses (the reference throws).
__device__ EONB200_NOINLINE int fire_step(Ctx &c, OptState &s, Objective &o, double /*maxMove*/) {
  const EonB200SearchParams &prm = *c.prm;
  const int n = o.dof;
  const double f_inc = 1.1, f_dec = 0.5, f_a = 0.99, alpha_start = 0.1;
  const long long N_min = 5;
  double *vel = c.slot(S_C), *x = c.slot(S_A), *dx = c.slot(S_E), *grad = c.slot(S_GRAD);
  if (obj_converged(c, o)) return 1;
  obj_gradient(c, o, grad, false);
  obj_get_positions(c, o, x);
  VLOOP(k, n) {
    const double v = vel[k] + (-grad[k]) * s.dt;
    vel[k] = v;
    dx[k] = v * s.dt;
  }
  __syncthreads();
  max_atom_motion_apply(c, dx, n, prm.max_move);
  obj_set_positions(c, o, x, dx, 1.0);
  obj_gradient(c, o, grad, false);
  double r[3] = {0.0, 0.0, 0.0};  // f.f, f.v, v.v
  VLOOP(k, n) {
    const double f = -grad[k];
    r[0] += f * f;
    r[1] += f * vel[k];
    r[2] += vel[k] * vel[k];
  }
  block_sum<3>(r, c.sm.red);
  const double invf = 1.0 / sqrt(r[0]);
  const double P = r[1], vn = sqrt(r[2]);
  VLOOP(k, n) vel[k] = (1 - s.alpha) * vel[k] + s.alpha * ((-grad[k]) * invf) * vn;
  __syncthreads();
  if (P >= 0) {
    s.fire_N++;
    if (s.fire_N > N_min) {
      s.dt = fmin(s.dt * f_inc, prm.max_time_step);
      s.alpha = s.alpha * f_a;
    }
  } else {
    s.dt = s.dt * f_dec;
    VLOOP(k, n) vel[k] = 0.0;
    __syncthreads();
    s.alpha = alpha_start;
    s.fire_N = 0;
  }
  s.iteration++;
  if (s.dt < 1e-6) return -1;
  return obj_converged(c, o) ? 1 : 0;
}

__device__ inline int opt_step(Ctx &c, OptState &s, Objective &o, double maxMove) {
  switch (c.prm->opt_method) {
    case EONB200_OPT_LBFGS:
      return lbfgs_step(c, s, o, maxMove);
    case EONB200_OPT_FIRE:
      return fire_step(c, s, o, maxMove);
    default:
      return cg_step(c, s, o, maxMove);
  }
}

// ---- drivers ---------------------------------------------------------------------------------------------------------
// MinModeSaddleSearch::run (client/MinModeSaddleSearch.cpp:211-475).
__device__ inline void saddle_search_walker(Ctx &c, MatterS &matter, double *mode, double reactant_energy, int &status_out,
                                            long long &iterations_out, double &eigenvalue_out, double &conv_out) {
  const EonB200SearchParams &prm = *c.prm;
  MinModeState mm;
  mm.init_cg = prm.dimer_opt_method == EONB200_ROT_CG ? 1 : 0;
  mm.init_lbfgs = 0;
  mm.rot_hist = 0;
  mm.C_tau = 0.0;
  mm.lowestEw = 0.0;
  mm.rotations_total = 0;
  mm.stats_rotations = 0;
  Objective objf{OBJ_MINMODE, c.D, &matter, mode, &mm, 0};
  OptState opt;
  opt_init(c, opt, objf);
  vzero(c.slot(S_THETA_OLD), c.D);
  vzero(c.slot(S_FR_OLD), c.D);
  int status = 0;
  long long iteration = 0;
  bool first = true;
  double *initial = nullptr;
  if (prm.saddle_nonlocal_count_abort != 0) {
    // a slot the active min-mode method does not use
    const bool dimer_lbfgs = prm.minmode_method == EONB200_MINMODE_DIMER && prm.dimer_opt_method == EONB200_ROT_LBFGS;
    initial = dimer_lbfgs ? c.slot(S_LZ_U) : c.slot(S_TAU_OLD);
    vcopy(initial, matter.pos, c.D);
  }
  while ((!obj_converged(c, objf) || iteration == 0) && !c.failed) {
    if (!first) {
      if (prm.saddle_nonlocal_count_abort != 0) {
        double nm = 0.0;
        VLOOP(i, c.N) {
          const double dx = initial[3 * i] - matter.pos[3 * i], dy = initial[3 * i + 1] - matter.pos[3 * i + 1],
                       dz = initial[3 * i + 2] - matter.pos[3 * i + 2];
          if (sqrt(dx * dx + dy * dy + dz * dz) >= prm.saddle_nonlocal_distance_abort) nm += 1.0;
        }
        nm = block_sum1(nm, c.sm.red);
        if (nm >= static_cast<double>(prm.saddle_nonlocal_count_abort)) {
          status = 13;  // STATUS_NONLOCAL_ABORT
          break;
        }
      }
      if (fabs(minmode_eigenvalue(c, mm)) < prm.saddle_zero_mode_abort_curvature) {
        status = 18;  // STATUS_ZEROMODE_ABORT
        break;
      }
    }
    first = false;
    if (iteration >= prm.saddle_max_iterations) {
      status = 5;  // STATUS_BAD_MAX_ITERATIONS
      break;
    }
    const int optStatus = opt_step(c, opt, objf, prm.max_move);
    if (optStatus < 0) {
      status = 19;  // STATUS_OPTIMIZER_ERROR
      break;
    }
    const double de = obj_energy(c, objf) - reactant_energy;
    iteration++;
    if (de > prm.saddle_max_energy) {
      status = 3;  // STATUS_BAD_HIGH_ENERGY
      break;
    }
  }
  if (!c.failed) {
    if (iteration == 0) {
      if (prm.minmode_method == EONB200_MINMODE_LANCZOS)
        lanczos_compute(c, mm, matter, mode);
      else
        dimer_compute(c, mm, matter, mode);
    }
    const bool conv = obj_converged(c, objf);
    if (status == 0 && !conv) status = 5;  // finalizeClimbStatus, MinModeSaddleSearch.h:90-96
    if (minmode_eigenvalue(c, mm) > 0.0 && status == 0) status = 16;
    conv_out = obj_convergence(c, objf);
  }
  if (c.failed) {
    status = 11;  // STATUS_POTENTIAL_FAILED (SaddleSearchJob.cpp:114-123)
    conv_out = nan("");
  }
  status_out = status;
  iterations_out = iteration;
  eigenvalue_out = minmode_eigenvalue(c, mm);
}

// helpers::relaxMatter (client/HelperFunctions.cpp:314-410).
__device__ inline void relax_walker(Ctx &c, MatterS &matter, int &status_out, long long &iterations_out,
                                    double &conv_out) {
  const EonB200SearchParams &prm = *c.prm;
  Objective objf{OBJ_MATTER, c.Dfree, &matter, nullptr, nullptr, 0};
  OptState opt;
  opt_init(c, opt, objf);
  long long iteration = 0;
  while (!obj_converged(c, objf) && iteration < prm.opt_max_iterations && !c.failed) {
    if (opt_step(c, opt, objf, prm.max_move) < 0) break;
    iteration++;
  }
  if (c.failed) {
    status_out = 11;
    conv_out = nan("");
  } else {
    status_out = obj_converged(c, objf) ? 0 : 1;
    conv_out = obj_convergence(c, objf);
  }
  iterations_out = iteration;
}

}  // namespace eonb200
