// potentials.cuh -- FP64 energy/force evaluation of one walker by its CTA.
//
// Thread mapping: `lpa` adjacent lanes share one atom and split its neighbour row; the atom's
// force and density are combined with an xor butterfly inside the lane group (fixed order), and
// per-walker scalars (energy, pair count) with one block reduction.  Neighbour rows come from the
// cached candidate list (nlist.cuh); every candidate is re-tested with the potential's own exact
// predicate, so the result does not depend on the skin.
//
//   LJ / LJCluster / Morse_Pt: arithmetic of the un-vendored rgpot kernels the reference reaches
//     through RgpotAdapter (include/eon/potentials/RgpotAdapter/RgpotAdapter.h:41-50,68-94,
//     factory arms client/Potential.cpp:140-153); forms pinned by the reference's goldens
//     (client/unit_tests/PotTest.cpp:38-43,130-132; data/reference/point_morse_pt.dat).
//   EAM: client/potentials/EAM/EAM.cpp:125-226,349-365 as a density pass, the embedding
//     polynomial, and a force pass that gathers F'_i g'_i(r) + F'_j g'_j(r) + pair'(r).
#pragma once
#include "nlist.cuh"

namespace eonb200 {

__device__ __forceinline__ double group_sum_rt(double v, int lpa) {
#ifndef EONB200_EMU
  for (int o = lpa >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}

// ---- LJ / Morse --------------------------------------------------------------------------------
// Forces land in sm.stage ([3N], raw, unmasked).  Returns the energy to every thread.
template <int KIND>
__device__ EONB200_NOINLINE double pair_forces(const PotParams &P, const SysView &sys, Smem &sm, int lpa) {
  const int N = sys.N;
  const int group = threadIdx.x / lpa, gl = threadIdx.x % lpa, ngroups = blockDim.x / lpa;
  const Geom &g = *sm.geom;
  double H[9];
#pragma unroll
  for (int k = 0; k < 9; k++) H[k] = g.H[k];
  const double rc2 = P.cutoff2, eshift = P.eshift;
  const double De = P.morse_De, ma = P.morse_a, re = P.morse_re;
  const double u0 = P.lj_u0, psi2 = P.lj_psi * P.lj_psi;
  double e_acc = 0.0, n_acc = 0.0;
  for (int i0 = 0; i0 < N; i0 += ngroups) {
    const int i = i0 + group;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (i < N) {
      const double4 pi = sm.pos[i];
      const int n = sys.cnt[i];
      for (int k = gl; k < n; k += lpa) {
        const int code = sys.nbr[static_cast<size_t>(k) * sys.Np + i];
        int j, sx, sy, sz;
        unpack_nbr(code, j, sx, sy, sz);
        const double4 pj = sm.pos[j];
        double vx = sub_rn(pj.x, pi.x), vy = sub_rn(pj.y, pi.y), vz = sub_rn(pj.z, pi.z);
        if ((sx | sy | sz) != 0) {
          const double ax = static_cast<double>(sx), ay = static_cast<double>(sy), az = static_cast<double>(sz);
          vx = add_rn(vx, add_rn(add_rn(mul_rn(ax, H[0]), mul_rn(ay, H[3])), mul_rn(az, H[6])));
          vy = add_rn(vy, add_rn(add_rn(mul_rn(ax, H[1]), mul_rn(ay, H[4])), mul_rn(az, H[7])));
          vz = add_rn(vz, add_rn(add_rn(mul_rn(ax, H[2]), mul_rn(ay, H[5])), mul_rn(az, H[8])));
        }
        const double d2 = norm2_rn(vx, vy, vz);
        if (!(d2 < rc2)) continue;
        double e, s;  // pair energy and (dU/dr)/r
        if (KIND == POT_MORSE_PT) {
          const double r = sqrt(d2);
          const double ex = exp(-ma * (r - re));
          const double d = 1.0 - ex;
          e = De * d * d - De;
          s = (2.0 * De * ma * d * ex) / r;
        } else {
          const double sr2 = psi2 / d2;
          const double sr6 = sr2 * sr2 * sr2;
          e = 4.0 * u0 * (sr6 * sr6 - sr6);
          s = -24.0 * u0 * (2.0 * sr6 * sr6 - sr6) / d2;
        }
        // F_i = +(dU/dr) v / r with v = r_j - r_i; each unordered pair is seen from both ends
        fx += s * vx;
        fy += s * vy;
        fz += s * vz;
        e_acc += 0.5 * (e - eshift);
        n_acc += 1.0;
      }
    }
    fx = group_sum_rt(fx, lpa);
    fy = group_sum_rt(fy, lpa);
    fz = group_sum_rt(fz, lpa);
    if (i < N && gl == 0) {
      sm.stage[3 * i] = fx;
      sm.stage[3 * i + 1] = fy;
      sm.stage[3 * i + 2] = fz;
    }
  }
  double v[2] = {e_acc, n_acc};
  block_sum<2>(v, sm.red);
  if (threadIdx.x == 0) atomicAdd(&sys.stats[2], static_cast<unsigned long long>(v[1]));
  return v[0];
}

// ---- EAM ------------------------------------------------------------------------------------------
__device__ __forceinline__ double eam_embed(const double *c, double rho) {  // EAM.cpp:349-356
  double acc = c[8];
#pragma unroll
  for (int k = 7; k >= 0; k--) acc = acc * rho + c[k];
  return acc;
}
__device__ __forceinline__ double eam_dembed(const double *dc, double rho) {  // +dF/drho, EAM.cpp:358-365
  double acc = dc[8];
#pragma unroll
  for (int k = 7; k >= 1; k--) acc = acc * rho + dc[k];
  return acc;
}
__device__ __forceinline__ void eam_fold(double &d, double L, double h) {  // EAM.cpp:157-168
  if (d > h)
    d -= L;
  else if (d < -h)
    d += L;
}

// Element rows of the atoms (sm.rowid); returns false to every thread if an element is unknown.
__device__ inline bool eam_assign_rows(const PotParams &P, const SysView &sys, Smem &sm, bool &multi) {
  double bad = 0.0, mixed = 0.0;
  const int z0 = sys.Z[0];
  for (int i = threadIdx.x; i < sys.N; i += blockDim.x) {
    const int z = sys.Z[i];
    int r = 255;
    for (int q = 0; q < P.n_rows; q++)
      if (P.rows[q].Z == z) {
        r = q;
        break;
      }
    sm.rowid[i] = static_cast<unsigned char>(r);
    if (r == 255) bad = 1.0;
    if (z != z0) mixed = 1.0;
  }
  double v[2] = {bad, mixed};
  const int op[2] = {RED_MAX, RED_MAX};
  block_reduce<2>(v, op, sm.red);
  multi = v[1] != 0.0;
  return v[0] == 0.0;
}

// g(r) = r^6 (e^{-b1 r} + 512 e^{-2 b2 r}) and g'(r), EAM.cpp:174-190
__device__ __forceinline__ void eam_density_terms(const EamRowDev &p, double r, double r2, double &gval, double &gder) {
  const double r6 = r2 * r2 * r2;
  const double e1 = exp(-p.beta1 * r);
  const double e2 = exp(-2.0 * p.beta2 * r);
  const double rho_pair = e1 + 512.0 * e2;
  gval = r6 * rho_pair;
  const double r5 = r2 * r2 * r;
  gder = 6.0 * r5 * rho_pair + r6 * (-p.beta1 * e1 - 1024.0 * p.beta2 * e2);
}

template <bool MULTI>
__device__ EONB200_NOINLINE double eam_forces(const PotParams &P, const SysView &sys, Smem &sm, int lpa) {
  const int N = sys.N;
  const int group = threadIdx.x / lpa, gl = threadIdx.x % lpa, ngroups = blockDim.x / lpa;
  const Geom &g = *sm.geom;
  const double Lx = g.H[0], Ly = g.H[4], Lz = g.H[8];
  const double hx = Lx * 0.5, hy = Ly * 0.5, hz = Lz * 0.5;
  const EamRowDev row0 = P.rows[sm.rowid[0]];
  double e_acc = 0.0, n_acc = 0.0;

  // pass 1: densities with the row of atom i (EAM.cpp:140,172-183) -> embedding energy, dF/drho
  for (int i0 = 0; i0 < N; i0 += ngroups) {
    const int i = i0 + group;
    double dens = 0.0;
    if (i < N) {
      const double4 pi = sm.pos[i];
      const int n = sys.cnt[i];
      const EamRowDev *pr = MULTI ? &P.rows[sm.rowid[i]] : &row0;
      const double b1 = pr->beta1, b2x2 = 2.0 * pr->beta2, rc2 = pr->rc2;
      for (int k = gl; k < n; k += lpa) {
        const int j = sys.nbr[static_cast<size_t>(k) * sys.Np + i] & kNbrIndexMask;
        const double4 pj = sm.pos[j];
        double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        eam_fold(dx, Lx, hx);
        eam_fold(dy, Ly, hy);
        eam_fold(dz, Lz, hz);
        const double r2 = norm2_rn(dx, dy, dz);
        if (r2 > rc2) continue;  // EAM.cpp:172 (inclusive cutoff)
        const double r = sqrt(r2);
        const double r6 = r2 * r2 * r2;
        dens += r6 * (exp(-b1 * r) + 512.0 * exp(-b2x2 * r));
        n_acc += 1.0;
      }
    }
    dens = group_sum_rt(dens, lpa);
    if (i < N && gl == 0) {
      const EamRowDev *pr = MULTI ? &P.rows[sm.rowid[i]] : &row0;
      e_acc += eam_embed(pr->c, dens);
      sm.dfd[i] = eam_dembed(pr->dc, dens);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) sm.pos[i].w = sm.dfd[i];
  __syncthreads();

  // pass 2: F_i = -sum_j [F'_i g'_i(r) + F'_j g'_j(r) + s phi'_m(r)] d / r, d = r_i - r_j,
  // m = min(i, j) (EAM.cpp:200: the pair part uses the row of the lower index), s = +1 for
  // F = -grad U and -1 for the reference's own sign (EAM.cpp:203-215).
  const double pair_sign = P.eam_conservative ? 1.0 : -1.0;
  for (int i0 = 0; i0 < N; i0 += ngroups) {
    const int i = i0 + group;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (i < N) {
      const double4 pi = sm.pos[i];
      const int n = sys.cnt[i];
      const EamRowDev *pri = MULTI ? &P.rows[sm.rowid[i]] : &row0;
      for (int k = gl; k < n; k += lpa) {
        const int j = sys.nbr[static_cast<size_t>(k) * sys.Np + i] & kNbrIndexMask;
        const double4 pj = sm.pos[j];
        double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        eam_fold(dx, Lx, hx);
        eam_fold(dy, Ly, hy);
        eam_fold(dz, Lz, hz);
        const double r2 = norm2_rn(dx, dy, dz);
        double coef = 0.0;
        double r, invr;
        if (!MULTI) {
          if (r2 > pri->rc2) continue;
          r = sqrt(r2);
          invr = 1.0 / r;
          double gv, gd;
          eam_density_terms(*pri, r, r2, gv, gd);
          coef = (pi.w + pj.w) * gd;
          const double ex = exp(-pri->alphaM * (r - pri->Rm));
          const double d = 1.0 - ex;
          coef += pair_sign * 2.0 * pri->alphaM * pri->Dm * d * ex;
          e_acc += 0.5 * (pri->Dm * d * d - pri->Dm);
        } else {
          const EamRowDev *prj = &P.rows[sm.rowid[j]];
          const EamRowDev *prm = (i < j) ? pri : prj;
          const bool in_i = !(r2 > pri->rc2), in_j = !(r2 > prj->rc2), in_m = !(r2 > prm->rc2);
          if (!(in_i || in_j)) continue;
          r = sqrt(r2);
          invr = 1.0 / r;
          double gv, gd;
          if (in_i) {
            eam_density_terms(*pri, r, r2, gv, gd);
            coef += pi.w * gd;
          }
          if (in_j) {
            eam_density_terms(*prj, r, r2, gv, gd);
            coef += pj.w * gd;
          }
          if (in_m) {
            const double ex = exp(-prm->alphaM * (r - prm->Rm));
            const double d = 1.0 - ex;
            coef += pair_sign * 2.0 * prm->alphaM * prm->Dm * d * ex;
            e_acc += 0.5 * (prm->Dm * d * d - prm->Dm);
          }
        }
        const double s3 = -coef * invr;
        fx += s3 * dx;
        fy += s3 * dy;
        fz += s3 * dz;
      }
    }
    fx = group_sum_rt(fx, lpa);
    fy = group_sum_rt(fy, lpa);
    fz = group_sum_rt(fz, lpa);
    if (i < N && gl == 0) {
      sm.stage[3 * i] = fx;
      sm.stage[3 * i + 1] = fy;
      sm.stage[3 * i + 2] = fz;
    }
  }
  double v[2] = {e_acc, n_acc};
  block_sum<2>(v, sm.red);
  if (threadIdx.x == 0) atomicAdd(&sys.stats[2], static_cast<unsigned long long>(v[1]));
  return v[0];
}

// Energy + raw forces (into sm.stage) of the walker whose positions are staged in sm.pos.
// Returns NaN if an element has no EAM row (the reference throws 14324, EAM.cpp:116-123).
__device__ inline double walker_forces(const PotParams &P, const SysView &sys, Smem &sm, int lpa) {
  switch (P.kind) {
    case POT_MORSE_PT:
      return pair_forces<POT_MORSE_PT>(P, sys, sm, lpa);
    case POT_EAM: {
      bool multi = false;
      if (!eam_assign_rows(P, sys, sm, multi)) {
        for (int k = threadIdx.x; k < 3 * sys.N; k += blockDim.x) sm.stage[k] = 0.0;
        __syncthreads();
        return nan("");
      }
      return multi ? eam_forces<true>(P, sys, sm, lpa) : eam_forces<false>(P, sys, sm, lpa);
    }
    default:
      return pair_forces<POT_LJ>(P, sys, sm, lpa);
  }
}

}  // namespace eonb200
