// potentials.cuh -- FP64 energy/force evaluation of (a tile of) one walker by a CTA.
//
// Thread mapping: `lpa` adjacent lanes share one atom and split its neighbour row; the atom's
// force / density are combined with an xor butterfly inside the lane group (fixed order) and the
// per-walker scalars (energy, pair count) with a block reduction.  Rows come from the cached
// candidate list (nlist.cuh): pairs inside the potential's cutoff first, the Verlet skin shell
// after them, so the lanes of a warp accept and reject together.  Every candidate is re-tested
// with the potential's own exact (un-fused) predicate, so results do not depend on the skin.
//
//   LJ / LJCluster / Morse_Pt: arithmetic of the un-vendored rgpot kernels the reference reaches
//     through RgpotAdapter (include/eon/potentials/RgpotAdapter/RgpotAdapter.h:41-50,68-94,
//     factory arms client/Potential.cpp:140-153); forms pinned by the reference's goldens
//     (client/unit_tests/PotTest.cpp:38-43,130-132; data/reference/point_morse_pt.dat).
//   EAM: client/potentials/EAM/EAM.cpp:125-226,349-365 as a pipeline:
//     pass 1 (density)  per ordered pair: r, the three exponentials, rho_i += g(r); the pair's
//                       g'(r)/r and +-phi'(r)/r are kept in HBM/L2 scratch (pairA / pairB);
//     embedding         F(rho_i), F'(rho_i) by Horner (EAM.cpp:349-365);
//     pass 2 (force)    F_i = -sum_j [(F'_i + F'_j) g'(r)/r +- phi'(r)/r] d, a pure gather with no
//                       transcendental, sqrt or division.
//   Systems with several elements take a slower variant of pass 2 that re-derives the per-row terms.
#pragma once
#include "nlist.cuh"

namespace eonb200 {

__device__ __forceinline__ double group_sum_rt(double v, int lpa) {
#ifndef EONB200_EMU
  __syncwarp();  // (callers come out of pair loops of lane-dependent length: reconverge before the shuffles)
  for (int o = lpa >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}

// ---- exp ---------------------------------------------------------------------------------------------
// Table-driven exp in the style of glibc's: x = (128 k + j) ln2/128 + r, |r| <= ln2/256,
// e^x = 2^k * T[j] * (1 + r + ... + r^5/120).  T = 2^(j/128) lives in shared memory.  About a third of
// the FP64 instructions of the library exp; error below 1 ulp over the range the potentials use
// (arguments in [-700, 700]; smaller arguments return 0).
__device__ __forceinline__ double exp_tab(double x, const double *T) {
  const double kInv = 184.6649652337873;            // 128 / ln 2
  const double kShift = 6755399441055744.0;         // 1.5 * 2^52
  const double kLn2Hi = 6.93147180369123816490e-01 / 128.0;  // high part of ln2 (trailing zeros), exact / 128
  const double kLn2Lo = 1.90821492927058770002e-10 / 128.0;
  double kd = fma(x, kInv, kShift);
#ifdef EONB200_EMU
  long long ki;
  memcpy(&ki, &kd, 8);
#else
  const long long ki = __double_as_longlong(kd);
#endif
  kd -= kShift;
  double r = fma(kd, -kLn2Hi, x);
  r = fma(kd, -kLn2Lo, r);
  const int k = static_cast<int>(ki);
  const double t = T[k & 127];
  double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double res = fma(t, q * r, t);
  const long long e = static_cast<long long>(k >> 7) << 52;
#ifdef EONB200_EMU
  long long bits;
  memcpy(&bits, &res, 8);
  bits += e;
  double out;
  memcpy(&out, &bits, 8);
#else
  const double out = __longlong_as_double(__double_as_longlong(res) + e);
#endif
  return x < -700.0 ? 0.0 : out;
}

__device__ __forceinline__ double rsqrt_f64(double x) {
#ifdef EONB200_EMU
  return 1.0 / sqrt(x);
#else
  return rsqrt(x);
#endif
}

// Exp arguments stay far above the underflow guard of exp_tab for any physical row (the host checks
// PotParams::eam_fast_ok), so the guard is dropped here.
__device__ __forceinline__ double exp_tab_fast(double x, const double *T) {
  const double kInv = 184.6649652337873;
  const double kShift = 6755399441055744.0;
  const double kLn2Hi = 6.93147180369123816490e-01 / 128.0;
  const double kLn2Lo = 1.90821492927058770002e-10 / 128.0;
  double kd = fma(x, kInv, kShift);
#ifdef EONB200_EMU
  long long ki;
  memcpy(&ki, &kd, 8);
#else
  const long long ki = __double_as_longlong(kd);
#endif
  kd -= kShift;
  double r = fma(kd, -kLn2Hi, x);
  r = fma(kd, -kLn2Lo, r);
  const int k = static_cast<int>(ki);
  const double t = T[k & 127];
  double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double res = fma(t, q * r, t);
#ifdef EONB200_EMU
  const long long e = static_cast<long long>(k >> 7) << 52;
  long long bits;
  memcpy(&bits, &res, 8);
  bits += e;
  double out;
  memcpy(&out, &bits, 8);
  return out;
#else
  // times 2^(k >> 7): one 32-bit add on the exponent word
  return __hiloint2double(__double2hiint(res) + ((k >> 7) << 20), __double2loint(res));
#endif
}

// 1/sqrt(x) for normal positive x (pair distances): hardware seed + one third-order step, without the library's
// special-case branch.
__device__ __forceinline__ double rsqrt_fast(double x) {
#ifdef EONB200_EMU
  return 1.0 / sqrt(x);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(x * y, -y, 1.0);
  const double c = fma(e, 0.375, 0.5);
  return fma(c * e, y, y);
#endif
}

// d = r_i - r_j folded by the image recorded at build time (image table valid: Geom::tab_ok).
__device__ __forceinline__ void mic_vec_tab(const Smem &sm, int code, const double4 &pi, const double4 &pj, double &dx,
                                            double &dy, double &dz) {
  const double4 sh = sm.sh27[static_cast<unsigned>(code) >> kNbrIndexBits];
  dx = (pi.x - pj.x) + sh.x;
  dy = (pi.y - pj.y) + sh.y;
  dz = (pi.z - pj.z) + sh.z;
}

// ---- pair vectors --------------------------------------------------------------------------------------
// NL_VESIN: v = (r_j - r_i) + S.H, un-fused (vesin-single-build.cpp:1199-1202).
__device__ __forceinline__ void vesin_vec(const Smem &sm, int code, const double4 &pi, const double4 &pj, double &vx,
                                          double &vy, double &vz) {
  const Geom &g = *sm.geom;
  vx = sub_rn(pj.x, pi.x);
  vy = sub_rn(pj.y, pi.y);
  vz = sub_rn(pj.z, pi.z);
  const unsigned s = static_cast<unsigned>(code) >> kNbrIndexBits;
  if (s == static_cast<unsigned>(kShiftNeutral)) return;
  if (g.tab_ok) {  // orthorhombic cell: S.H is one product per axis
    vx = add_rn(vx, sm.shx[s & 15]);
    vy = add_rn(vy, sm.shy[(s >> 4) & 15]);
    vz = add_rn(vz, sm.shz[(s >> 8) & 15]);
  } else {
    const double ax = static_cast<double>(static_cast<int>(s & 15) - 8), ay = static_cast<double>(static_cast<int>((s >> 4) & 15) - 8),
                 az = static_cast<double>(static_cast<int>((s >> 8) & 15) - 8);
    vx = add_rn(vx, add_rn(add_rn(mul_rn(ax, g.H[0]), mul_rn(ay, g.H[3])), mul_rn(az, g.H[6])));
    vy = add_rn(vy, add_rn(add_rn(mul_rn(ax, g.H[1]), mul_rn(ay, g.H[4])), mul_rn(az, g.H[7])));
    vz = add_rn(vz, add_rn(add_rn(mul_rn(ax, g.H[2]), mul_rn(ay, g.H[5])), mul_rn(az, g.H[8])));
  }
}

// NL_MIC: d = r_i - r_j folded once by +-L (EAM.cpp:152-168).  The fold recorded at build time is the
// reference's fold while the box is at least 2 rc_list + skin wide; otherwise it is re-derived.
__device__ __forceinline__ void mic_vec(const Smem &sm, int code, const double4 &pi, const double4 &pj, double &dx, double &dy,
                                        double &dz) {
  const Geom &g = *sm.geom;
  dx = pi.x - pj.x;
  dy = pi.y - pj.y;
  dz = pi.z - pj.z;
  if (g.tab_ok) {
    const double4 sh = sm.sh27[static_cast<unsigned>(code) >> kNbrIndexBits];
    dx += sh.x;
    dy += sh.y;
    dz += sh.z;
  } else {
    const double Lx = g.H[0], Ly = g.H[4], Lz = g.H[8];
    if (dx > 0.5 * Lx) dx -= Lx; else if (dx < -0.5 * Lx) dx += Lx;
    if (dy > 0.5 * Ly) dy -= Ly; else if (dy < -0.5 * Ly) dy += Ly;
    if (dz > 0.5 * Lz) dz -= Lz; else if (dz < -0.5 * Lz) dz += Lz;
  }
}

// ---- LJ / Morse ------------------------------------------------------------------------------------------
// Atoms [i0, i1) of the walker staged in sm.pos; forces to F[3 (i - fbase)...] (shared or global).
// acc[0] += energy share, acc[1] += ordered pairs (per-lane partials, reduced by the caller).
template <int KIND, bool FASTEXP>
// near_build: every atom within NlParams::dstar of the list's build positions -- the leading sure-in rows skip the
// cutoff test, the trailing sure-out rows are not visited.
__device__ __forceinline__ void pair_pass_body(const PotParams &P, const SysView &sys, const Smem &sm, int lpa, int i0, int i1,
                                               double *F, double (&acc)[2], bool near_build) {
  const int group = threadIdx.x / lpa, gl = threadIdx.x % lpa, ngroups = blockDim.x / lpa;
  const double rc2 = P.cutoff2, eshift = P.eshift;
  const double De = P.morse_De, ma = P.morse_a, re = P.morse_re;
  const double u0 = P.lj_u0, psi2 = P.lj_psi * P.lj_psi;
  double e_acc = 0.0, n_acc = 0.0;
  for (int ib = i0; ib < i1; ib += ngroups) {
    const int i = ib + group;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (i < i1) {
      const double4 pi = sm.pos[i];
      int n = sys.cnt[i], n_sure = 0;
      if (near_build && sys.cls) {
        const int c = sys.cls[i];
        n_sure = c & 0xffff;
        n = n_sure + (c >> 16);
      }
      // eight candidate codes of this lane in flight at a time (the loads are a dependent L2 round trip otherwise)
      constexpr int kPairBatch = 8;
      const size_t Np = static_cast<size_t>(sys.Np);
      for (int k0 = gl; k0 < n; k0 += kPairBatch * lpa) {
        int cb[kPairBatch];
#pragma unroll
        for (int u = 0; u < kPairBatch; u++) {
          const int kk = k0 + u * lpa;
          cb[u] = kk < n ? sys.nbr[static_cast<size_t>(kk) * Np + i] : -1;
        }
#pragma unroll
        for (int u = 0; u < kPairBatch; u++) {
          const int code = cb[u];
          if (code == -1) continue;
          const int k = k0 + u * lpa;
          const double4 pj = sm.pos[code & kNbrIndexMask];
          double vx, vy, vz;
          vesin_vec(sm, code, pi, pj, vx, vy, vz);
          const double d2 = norm2_rn(vx, vy, vz);
          if (k >= n_sure && !(d2 < rc2)) continue;
          double e, s;  // pair energy and (dU/dr)/r
          if (KIND == POT_MORSE_PT) {
            const double invr = rsqrt_fast(d2);
            const double r = d2 * invr;
            const double ex = FASTEXP ? exp_tab_fast(-ma * (r - re), sm.exp_tab) : exp_tab(-ma * (r - re), sm.exp_tab);
            const double d = 1.0 - ex;
            e = De * d * d - De;
            s = (2.0 * De * ma * d * ex) * invr;
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
    }
    fx = group_sum_rt(fx, lpa);
    fy = group_sum_rt(fy, lpa);
    fz = group_sum_rt(fz, lpa);
    if (i < i1 && gl == 0) {
      F[3 * i] = fx;
      F[3 * i + 1] = fy;
      F[3 * i + 2] = fz;
    }
  }
  acc[0] += e_acc;
  acc[1] += n_acc;
}

template <int KIND>
__device__ __forceinline__ void pair_pass_inl(const PotParams &P, const SysView &sys, const Smem &sm, int lpa, int i0, int i1,
                                              double *F, double (&acc)[2], bool near_build) {
  if (KIND == POT_MORSE_PT && P.eam_fast_ok)  // unguarded exp: decided once, outside the pair loop
    pair_pass_body<KIND, true>(P, sys, sm, lpa, i0, i1, F, acc, near_build);
  else
    pair_pass_body<KIND, false>(P, sys, sm, lpa, i0, i1, F, acc, near_build);
}

// Out-of-line instance for the resident walker kernel (its Smem view lives in the walker context anyway).
template <int KIND>
__device__ EONB200_NOINLINE void pair_pass(const PotParams &P, const SysView &sys, const Smem &sm, int lpa, int i0, int i1,
                                          double *F, double (&acc)[2]) {
  pair_pass_inl<KIND>(P, sys, sm, lpa, i0, i1, F, acc, sm.geom->disp2 <= 0.0625 * P.skin * P.skin);
}

// ---- EAM ------------------------------------------------------------------------------------------------------
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

// Element rows of the atoms -> rowid[] (shared or global); flags: bit 0 mixed elements, bit 1 unknown element.
__device__ inline int eam_assign_rows(const PotParams &P, const SysView &sys, unsigned char *rowid, double *red) {
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
    rowid[i] = static_cast<unsigned char>(r);
    if (r == 255) bad = 1.0;
    if (z != z0) mixed = 1.0;
  }
  double v[2] = {bad, mixed};
  const int op[2] = {RED_MAX, RED_MAX};
  block_reduce<2>(v, op, red);
  return (v[1] != 0.0 ? 1 : 0) | (v[0] != 0.0 ? 2 : 0);
}

// g'(r) of one row: d/dr [r^6 (e^{-b1 r} + 512 e^{-2 b2 r})], EAM.cpp:174-190
__device__ __forceinline__ double eam_gder(const EamRowDev &p, double r, double r2, const double *T) {
  const double r6 = r2 * r2 * r2;
  const double e1 = exp_tab(-p.beta1 * r, T);
  const double e2 = exp_tab(-2.0 * p.beta2 * r, T);
  const double rho_pair = e1 + 512.0 * e2;
  const double r5 = r2 * r2 * r;
  return 6.0 * r5 * rho_pair + r6 * (-p.beta1 * e1 - 1024.0 * p.beta2 * e2);
}

// Pass 1 over atoms [i0, i1): densities with the row of atom i (EAM.cpp:140,172-183) -> dF/drho into
// dfd[i], embedding energy; single-element systems also do the whole pair arithmetic here.
template <bool MULTI>
__device__ EONB200_NOINLINE void eam_pass1(const PotParams &P, const SysView &sys, const Smem &sm, int lpa, int i0, int i1,
                                          double *dfd, double (&acc)[2]) {
  const int group = threadIdx.x / lpa, gl = threadIdx.x % lpa, ngroups = blockDim.x / lpa;
  const double *T = sm.exp_tab;
  const EamRowDev *row0 = &P.rows[sm.rowid[0]];
  const double pair_sign = P.eam_conservative ? 1.0 : -1.0;
  double e_acc = 0.0, n_acc = 0.0;
  for (int ib = i0; ib < i1; ib += ngroups) {
    const int i = ib + group;
    double dens = 0.0;
    if (i < i1) {
      const double4 pi = sm.pos[i];
      const int n = sys.cnt[i];
      const EamRowDev *pr = MULTI ? &P.rows[sm.rowid[i]] : row0;
      const double b1 = pr->beta1, b2 = pr->beta2, rc2 = pr->rc2;
      const double aM = pr->alphaM, Rm = pr->Rm, Dm = pr->Dm;
      for (int k = gl; k < n; k += lpa) {
        const size_t slot = static_cast<size_t>(k) * sys.Np + i;
        const int code = sys.nbr[slot];
        const double4 pj = sm.pos[code & kNbrIndexMask];
        double dx, dy, dz;
        mic_vec(sm, code, pi, pj, dx, dy, dz);
        const double r2 = norm2_rn(dx, dy, dz);
        if (r2 > rc2) continue;  // EAM.cpp:172 (inclusive cutoff)
        const double invr = rsqrt_f64(r2);
        const double r = r2 * invr;
        const double r6 = r2 * r2 * r2;
        const double e1 = exp_tab(-b1 * r, T);
        const double e2 = exp_tab(-2.0 * b2 * r, T);
        const double rho_pair = e1 + 512.0 * e2;
        dens += r6 * rho_pair;
        n_acc += 1.0;
        if (!MULTI) {
          const double r5 = r2 * r2 * r;
          const double gder = 6.0 * r5 * rho_pair + r6 * (-b1 * e1 - 1024.0 * b2 * e2);
          sys.pairA[slot] = gder * invr;
          const double ex = exp_tab(-aM * (r - Rm), T);
          const double d = 1.0 - ex;
          sys.pairB[slot] = pair_sign * (2.0 * aM * Dm * d * ex) * invr;
          e_acc += 0.5 * (Dm * d * d - Dm);  // the pair energy is counted from both ends
        }
      }
    }
    dens = group_sum_rt(dens, lpa);
    if (i < i1 && gl == 0) {
      const EamRowDev *pr = MULTI ? &P.rows[sm.rowid[i]] : row0;
      e_acc += eam_embed(pr->c, dens);
      dfd[i] = eam_dembed(pr->dc, dens);
    }
  }
  acc[0] += e_acc;
  acc[1] += n_acc;
}

// Pass 2 over atoms [i0, i1): sm.pos[j].w must hold dF/drho of atom j.
//   F_i = -sum_j [F'_i g'_i(r) + F'_j g'_j(r) + s phi'_m(r)] d / r,  d = r_i - r_j, m = min(i, j)
// (EAM.cpp:200: the pair part uses the row of the lower index), s = +1 for F = -grad U and -1 for the
// reference's own sign (EAM.cpp:203-215).
template <bool MULTI>
__device__ EONB200_NOINLINE void eam_pass2(const PotParams &P, const SysView &sys, const Smem &sm, int lpa, int i0, int i1,
                                          double *F, double (&acc)[2]) {
  const int group = threadIdx.x / lpa, gl = threadIdx.x % lpa, ngroups = blockDim.x / lpa;
  const double *T = sm.exp_tab;
  const EamRowDev *row0 = &P.rows[sm.rowid[0]];
  const double pair_sign = P.eam_conservative ? 1.0 : -1.0;
  double e_acc = 0.0;
  for (int ib = i0; ib < i1; ib += ngroups) {
    const int i = ib + group;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (i < i1) {
      const double4 pi = sm.pos[i];
      const int n = sys.cnt[i];
      const EamRowDev *pri = MULTI ? &P.rows[sm.rowid[i]] : row0;
      const double rc2 = pri->rc2;
      for (int k = gl; k < n; k += lpa) {
        const size_t slot = static_cast<size_t>(k) * sys.Np + i;
        const int code = sys.nbr[slot];
        const int j = code & kNbrIndexMask;
        const double4 pj = sm.pos[j];
        double dx, dy, dz;
        mic_vec(sm, code, pi, pj, dx, dy, dz);
        const double r2 = norm2_rn(dx, dy, dz);
        double s3;
        if (!MULTI) {
          if (r2 > rc2) continue;
          s3 = -((pi.w + pj.w) * sys.pairA[slot] + sys.pairB[slot]);
        } else {
          const EamRowDev *prj = &P.rows[sm.rowid[j]];
          const EamRowDev *prm = (i < j) ? pri : prj;
          const bool in_i = !(r2 > pri->rc2), in_j = !(r2 > prj->rc2), in_m = !(r2 > prm->rc2);
          if (!(in_i || in_j || in_m)) continue;
          const double invr = rsqrt_f64(r2);
          const double r = r2 * invr;
          double coef = 0.0;
          if (in_i) coef += pi.w * eam_gder(*pri, r, r2, T);
          if (in_j) coef += pj.w * eam_gder(*prj, r, r2, T);
          if (in_m) {
            const double ex = exp_tab(-prm->alphaM * (r - prm->Rm), T);
            const double d = 1.0 - ex;
            coef += pair_sign * 2.0 * prm->alphaM * prm->Dm * d * ex;
            e_acc += 0.5 * (prm->Dm * d * d - prm->Dm);
          }
          s3 = -coef * invr;
        }
        fx += s3 * dx;
        fy += s3 * dy;
        fz += s3 * dz;
      }
    }
    fx = group_sum_rt(fx, lpa);
    fy = group_sum_rt(fy, lpa);
    fz = group_sum_rt(fz, lpa);
    if (i < i1 && gl == 0) {
      F[3 * i] = fx;
      F[3 * i + 1] = fy;
      F[3 * i + 2] = fz;
    }
  }
  acc[0] += e_acc;
}

// ---- EAM, single element, one lane per atom: filter -> density -> (barrier) -> force ----------------------------
// The fast path of the fused kernel.  Per atom i (one thread):
//   filter   every cached candidate is screened in single precision against rc^2 +- a rounding band; inside the
//            band the reference's own predicate decides (EAM.cpp:152-172, double, un-fused, inclusive cutoff).
//            Survivors are compacted to the front of the atom's column of sys.nbr2, so the lanes of a warp run the
//            expensive body with (nearly) all lanes busy instead of idling through the skin shell;
//   density  per surviving ordered pair: r, the three exponentials (EAM.cpp:174-215); rho_i += g(r); the pair's
//            g'(r)/r goes to sys.pairA (L2 resident, 8 B per pair); the Morse pair force is applied right away;
//            then F(rho_i), F'(rho_i) by Horner (EAM.cpp:349-365);
//   force    F_i -= sum_j (F'_i + F'_j) g'(r)/r d  -- no transcendental, square root or division.
struct EamLane {   // what one atom's thread carries from the density pass into the force pass
  double fx, fy, fz;  // Morse pair part of the force, accumulated in the density pass
  double dfd;         // F'(rho_i)
  int n0;             // leading rows of the cached list taken as they are (pairs certain to be inside the cutoff)
  int nin;            // pairs the filter let through (compacted in nbr2, padded to a multiple of kForceBatch)
};

constexpr int kFilterBatch = 8;  // candidate codes a lane keeps in flight; NL_MIC rows are padded to a multiple of it
constexpr int kForceBatch = 4;   // same for the force pass over the compacted pairs

// Single-precision screening margin above rc^2: beyond it the float distance rejects, everything else goes on to
// the density pass, which applies the reference's own predicate (EAM.cpp:152-172, double, un-fused, inclusive
// cutoff) to the exact distance it computes anyway.  |p| <= L after the wrap, so every float coordinate is off by
// at most 2^-24 L and the float r^2 by well under 1e-6 (L + rc) rc; the margin keeps an 8x reserve over that.
__device__ __forceinline__ float eam_filter_band(double Lmax, double rc) {
  return static_cast<float>(8.0e-6 * (Lmax + rc) * rc);
}

// Which rows of atom i's cached list need the filter: with the atoms no further than NlParams::dstar from the
// list's build positions the leading sure-in rows are taken unfiltered and the trailing sure-out rows are skipped;
// otherwise every row is screened.
__device__ __forceinline__ void eam_filter_range(const SysView &sys, int i, bool near_build, int &n0, int &k_end) {
  if (near_build) {
    const int c = ld_cg(sys.cls + i);
    k_end = (c & 0xffff) + (c >> 16);
    n0 = (c & 0xffff) & ~(kForceBatch - 1);  // whole force-pass batches; the few rows left over just pass the filter
  } else {
    n0 = 0;
    k_end = ld_cg(sys.cnt + i);
  }
}

// Filter of rows [k_begin, k_end) of atom i: cur[] holds rows (k_begin & ~7) .. + 7 (already loaded).  Rows past
// k_end, up to the next multiple of 8, are valid codes too (far pairs or the ghost atom N of the row padding), so
// there is no upper bounds test.
__device__ __forceinline__ int eam_fast_filter(const SysView &sys, const Smem &sm, int i, int k_begin, int k_end,
                                               int (&cur)[kFilterBatch], float rc2_hi) {
  const size_t step = static_cast<size_t>(sys.Np);
  const int k_first = k_begin & ~(kFilterBatch - 1);  // batches stay aligned to the row padding
  const int *row = sys.nbr + i + static_cast<size_t>(k_first) * step;
  int *row2 = sys.nbr2 + i;
  const float4 pif = sm.posf[i];
  int nin = 0;
  for (int k0 = k_first; k0 < k_end; k0 += kFilterBatch) {
    int nxt[kFilterBatch];
    row += kFilterBatch * step;
    if (k0 + kFilterBatch < k_end) {
#pragma unroll
      for (int u = 0; u < kFilterBatch; u++) nxt[u] = ld_cg(row + u * step);
    }
#pragma unroll
    for (int u = 0; u < kFilterBatch; u++) {
      const int c = cur[u];
      const float4 shf = sm.sh27f[static_cast<unsigned>(c) >> kNbrIndexBits];
      const float4 pjf = sm.posf[c & kNbrIndexMask];
      const float dxf = (pif.x - pjf.x) + shf.x;
      const float dyf = (pif.y - pjf.y) + shf.y;
      const float dzf = (pif.z - pjf.z) + shf.z;
      const float r2f = fmaf(dzf, dzf, fmaf(dyf, dyf, dxf * dxf));
      if (r2f <= rc2_hi && k0 + u >= k_begin) {
        *row2 = c;
        row2 += step;
        nin++;
      }
    }
#pragma unroll
    for (int u = 0; u < kFilterBatch; u++) cur[u] = nxt[u];
  }
  return nin;
}

__device__ __forceinline__ void eam_fast_density(const PotParams &P, const EamRowDev *pr, const SysView &sys, const Smem &sm,
                                                 int i, int n0, int nin, EamLane &L, double (&acc)[2]) {
  const double *T = sm.exp_tab;
  const size_t step = static_cast<size_t>(sys.Np);
  const double4 pi = sm.pos[i];
  const int *row = sys.nbr + i;
  int *row2 = sys.nbr2 + i;
  double *rowA = sys.pairA + i;
  const double rc2 = pr->rc2;
  const double nb1 = -pr->beta1, nb2x2 = -2.0 * pr->beta2, nb2x1024 = -1024.0 * pr->beta2;
  const double naM = -pr->alphaM, aMRm = pr->alphaM * pr->Rm, hDm = 0.5 * pr->Dm;
  // sign of the reference's own pair force (EAM.cpp:203-215) or of -grad U
  const double nK2 = -(P.eam_conservative ? 1.0 : -1.0) * 2.0 * pr->alphaM * pr->Dm;
  double dens = 0.0, e_acc = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
  int npairs = 0;
  const int ntot = n0 + nin;
  const int *pc = n0 > 0 ? row : row2;  // walks the cached list's leading rows, then the compacted column
  int cnext = ntot > 0 ? ld_cg(pc) : 0;
  for (int kk = 0; kk < ntot; kk++) {
    const int code = cnext;
    pc += step;
    if (kk + 1 == n0) pc = row2;
    if (kk + 1 < ntot) cnext = ld_cg(pc);
    const double4 pj = sm.pos[code & kNbrIndexMask];
    const double4 sh = sm.sh27[static_cast<unsigned>(code) >> kNbrIndexBits];
    const double dx = (pi.x - pj.x) + sh.x;
    const double dy = (pi.y - pj.y) + sh.y;
    const double dz = (pi.z - pj.z) + sh.z;
    const double r2 = norm2_rn(dx, dy, dz);
    if (EONB200_UNLIKELY(r2 > rc2)) {  // EAM.cpp:172 -- only pairs inside the filter's single-precision margin get here
      *rowA = 0.0;
      rowA += step;
      continue;
    }
    npairs++;
    const double invr = rsqrt_fast(r2);
    const double r = r2 * invr;
    const double r4 = r2 * r2;
    const double e1 = exp_tab_fast(nb1 * r, T);
    const double e2 = exp_tab_fast(nb2x2 * r, T);
    const double ex = exp_tab_fast(fma(naM, r, aMRm), T);
    const double rho_pair = fma(512.0, e2, e1);
    const double r6 = r4 * r2;
    dens = fma(r6, rho_pair, dens);
    const double u = fma(nb2x1024, e2, nb1 * e1);
    // g'(r)/r = [6 r^5 rho_pair + r^6 u] / r = r^4 (6 rho_pair + r u)
    *rowA = r4 * fma(r, u, 6.0 * rho_pair);
    rowA += step;
    const double dm = 1.0 - ex;
    const double nB = (nK2 * dm) * (ex * invr);  // -(+-phi'(r)/r)
    fx = fma(nB, dx, fx);
    fy = fma(nB, dy, fy);
    fz = fma(nB, dz, fz);
    e_acc = fma(hDm, fma(dm, dm, -1.0), e_acc);  // half of Dm (d^2 - 1): the pair is seen from both ends
  }
  // pad the pair sequence to a multiple of kForceBatch with null pairs (weight 0, partner = the atom itself)
  int nin_padded = nin;
  for (int kk = ntot; kk & (kForceBatch - 1); kk++) {
    row2[static_cast<size_t>(nin_padded) * step] = i | (13 << kNbrIndexBits);
    *rowA = 0.0;
    rowA += step;
    nin_padded++;
  }
  e_acc += eam_embed(pr->c, dens);
  L.dfd = eam_dembed(pr->dc, dens);
  L.fx = fx;
  L.fy = fy;
  L.fz = fz;
  L.n0 = n0;
  L.nin = nin_padded;
  acc[0] += e_acc;
  acc[1] += static_cast<double>(npairs);
}

// sm.pos[j].w must hold F'(rho_j) of every atom of the walker.  L.n0 + L.nin is a multiple of kForceBatch.
__device__ __forceinline__ void eam_fast_force(const SysView &sys, const Smem &sm, int i, const EamLane &L, double *F) {
  constexpr int B = kForceBatch;
  const size_t step = static_cast<size_t>(sys.Np);
  const double4 pi = sm.pos[i];
  const int *row = sys.nbr + i;
  const int *row2 = sys.nbr2 + i;
  const double *rowA = sys.pairA + i;
  double fx = L.fx, fy = L.fy, fz = L.fz;
  const double dfi = L.dfd;
  const int n0 = L.n0, ntot = L.n0 + L.nin;
  int c[B];
  double A[B];
  const int *pc = n0 > 0 ? row : row2;  // n0 is a multiple of B: a batch never straddles the two sources
  if (ntot > 0) {
#pragma unroll
    for (int u = 0; u < B; u++) {
      c[u] = ld_cg(pc + u * step);
      A[u] = ld_cg(rowA + u * step);
    }
  }
  for (int k0 = 0; k0 < ntot; k0 += B) {
    int cn[B];
    double An[B];
    rowA += B * step;
    pc += B * step;
    if (k0 + B == n0) pc = row2;
    if (k0 + B < ntot) {
#pragma unroll
      for (int u = 0; u < B; u++) {
        cn[u] = ld_cg(pc + u * step);
        An[u] = ld_cg(rowA + u * step);
      }
    }
#pragma unroll
    for (int u = 0; u < B; u++) {
      const double4 pj = sm.pos[c[u] & kNbrIndexMask];
      const double4 sh = sm.sh27[static_cast<unsigned>(c[u]) >> kNbrIndexBits];
      const double dx = (pi.x - pj.x) + sh.x;
      const double dy = (pi.y - pj.y) + sh.y;
      const double dz = (pi.z - pj.z) + sh.z;
      const double ns = -(dfi + pj.w) * A[u];
      fx = fma(ns, dx, fx);
      fy = fma(ns, dy, fy);
      fz = fma(ns, dz, fz);
    }
#pragma unroll
    for (int u = 0; u < B; u++) {
      c[u] = cn[u];
      A[u] = An[u];
    }
  }
  F[3 * i] = fx;
  F[3 * i + 1] = fy;
  F[3 * i + 2] = fz;
}

// Whole walker in one CTA (resident kernels): the same filter / density / force lane functions as the fused kernel,
// thread-strided over the atoms; forces to sm.stage.  sm.pos holds the prepared positions.
__device__ EONB200_NOINLINE double walker_forces_eam_fast(const PotParams &P, const EamRowDev *row, const SysView &sys, Smem sm,
                                                         double dstar) {
  const int N = sys.N;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < N; i += nt) {
    const double4 p = sm.pos[i];
    sm.posf[i] = make_float4(static_cast<float>(p.x), static_cast<float>(p.y), static_cast<float>(p.z), 0.0f);
  }
  if (tid == 0) {  // the ghost atom the row padding points at
    sm.pos[N] = make_double4(1e150, 1e150, 1e150, 0.0);
    sm.posf[N] = make_float4(1e18f, 1e18f, 1e18f, 0.0f);
  }
  __syncthreads();
  const Geom &g = *sm.geom;
  const float rc2_hi = static_cast<float>(row->rc2) + eam_filter_band(fmax(fmax(g.H[0], g.H[4]), g.H[8]), row->r_cut);
  const bool single = N <= nt;
  const bool near_build = g.disp2 <= dstar * dstar;
  double acc[2] = {0.0, 0.0};
  EamLane L = {0.0, 0.0, 0.0, 0.0, 0, 0};
  for (int i = tid; i < N; i += nt) {
    int n0, k_end;
    eam_filter_range(sys, i, near_build, n0, k_end);
    int cur[kFilterBatch];
#pragma unroll
    for (int u = 0; u < kFilterBatch; u++)
      cur[u] = n0 < k_end ? sys.nbr[static_cast<size_t>((n0 & ~(kFilterBatch - 1)) + u) * sys.Np + i] : 0;
    const int nin = n0 < k_end ? eam_fast_filter(sys, sm, i, n0, k_end, cur, rc2_hi) : 0;
    eam_fast_density(P, row, sys, sm, i, n0, nin, L, acc);
    sm.pos[i].w = L.dfd;
    if (!single) {
      sm.stage[3 * i] = L.fx;
      sm.stage[3 * i + 1] = L.fy;
      sm.stage[3 * i + 2] = L.fz;
      sys.dfd[i] = L.dfd;
      sys.cnt2[i] = L.n0 | (L.nin << 16);
    }
  }
  block_sum<2>(acc, sm.red);  // (barriers: every F'(rho) is in place)
  for (int i = tid; i < N; i += nt) {
    if (!single) {
      L.fx = sm.stage[3 * i];
      L.fy = sm.stage[3 * i + 1];
      L.fz = sm.stage[3 * i + 2];
      L.dfd = sys.dfd[i];
      L.n0 = sys.cnt2[i] & 0xffff;
      L.nin = sys.cnt2[i] >> 16;
    }
    eam_fast_force(sys, sm, i, L, sm.stage);
  }
  __syncthreads();
  if (tid == 0) atomicAdd(&sys.stats[2], static_cast<unsigned long long>(acc[1]));
  return acc[0];
}

// Energy + raw forces (into sm.stage) of the whole walker staged in sm.pos -- the resident kernels' path.
// Returns NaN if an element has no EAM row (the reference throws 14324, EAM.cpp:116-123).
// row_flags_cache (optional): the walker's atomic numbers never change, so its element rows are assigned once.
__device__ inline double walker_forces(const PotParams &P, const SysView &sys, Smem &sm, int lpa, int *row_flags_cache = nullptr) {
  double acc[2] = {0.0, 0.0};
  const int N = sys.N;
  if (P.kind == POT_EAM) {
    int flags;
    if (row_flags_cache && *row_flags_cache >= 0) {
      flags = *row_flags_cache;
    } else {
      flags = eam_assign_rows(P, sys, sm.rowid, sm.red);
      if (row_flags_cache) *row_flags_cache = flags;
    }
    if (flags & 2) {
      for (int k = threadIdx.x; k < 3 * N; k += blockDim.x) sm.stage[k] = 0.0;
      __syncthreads();
      return nan("");
    }
    const bool multi = (flags & 1) != 0;
    if (!multi && sm.geom->tab_ok && P.eam_fast_ok && sm.posf != nullptr)
      return walker_forces_eam_fast(P, &P.rows[sm.rowid[0]], sys, sm, 0.25 * P.skin);
    if (multi)
      eam_pass1<true>(P, sys, sm, lpa, 0, N, sm.dfd, acc);
    else
      eam_pass1<false>(P, sys, sm, lpa, 0, N, sm.dfd, acc);
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) sm.pos[i].w = sm.dfd[i];
    __syncthreads();
    if (multi)
      eam_pass2<true>(P, sys, sm, lpa, 0, N, sm.stage, acc);
    else
      eam_pass2<false>(P, sys, sm, lpa, 0, N, sm.stage, acc);
  } else if (P.kind == POT_MORSE_PT) {
    pair_pass<POT_MORSE_PT>(P, sys, sm, lpa, 0, N, sm.stage, acc);
  } else {
    pair_pass<POT_LJ>(P, sys, sm, lpa, 0, N, sm.stage, acc);
  }
  block_sum<2>(acc, sm.red);
  if (threadIdx.x == 0) atomicAdd(&sys.stats[2], static_cast<unsigned long long>(acc[1]));
  return acc[0];
}

}  // namespace eonb200
