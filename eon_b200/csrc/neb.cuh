// neb.cuh -- nudged-elastic-band force projection for many bands at once.
//
// What NudgedElasticBand::updateForces does per movable image after the potential call
// (client/NudgedElasticBand.cpp:551-646): minimum-image differences to both neighbours, the tangent
// (client/NEBTangent.cpp:36-69), the spring force (client/NEBSpringForce.cpp:17-42,113-136), the projection
// (client/NEBProjection.cpp:17-31, client/NEBForceProjection.cpp:65-112), the climbing image and
// zeroTranslation.  One CTA per (band, image); a handful of passes over 3 N doubles, so the kernel is HBM bound:
// per image its position and force rows are read and tangent and projected force written (4 x 3 N x 8 bytes; the
// neighbours' position rows are shared with their own CTAs through L2).  k_neb_band_tma stages the rows in shared
// memory with bulk copies and is the one normally launched; k_neb_band is the any-size fallback working out of HBM/L2.
#pragma once
#include "common.cuh"

namespace eonb200 {

enum NebProjection { NEB_PROJ_NEB = 0, NEB_PROJ_DNEB = 1, NEB_PROJ_ELASTIC_BAND = 2 };

struct NebArgs {
  const double *pos;     // [bands][images + 2][3 N]
  const double *frc;     // same shape: Matter forces (fixed axes zeroed); endpoint rows are not read
  const double *energy;  // [bands][images + 2]
  double *projected;     // [bands][images][3 N]
  double *tangent;       // [bands][images][3 N]
  int *climbing;         // [bands]: index of the climbing image (0 = none), may be null
  const double *mask;    // [3 N] 1 = free axis, 0 = fixed (Matter::getForces, client/Matter.cpp:324-333); null = frc is
                         // already masked.  Only set by eonb200_neb_band_forces, which owns the force rows.
  int remove_net;        // subtract the mean force of each image first (client/Matter.cpp:531-536; no fixed atoms)
  double cell[9], inv[9];
  double ksp, k_min, k_max;
  int images, atoms;
  int cell_ok, all_free, ci_active, old_tangent, projection, weighted;
};

// pbc of one atom row (include/eon/Matter.h:42-50)
__device__ __forceinline__ void neb_mic(const NebArgs &a, double (&d)[3]) {
  if (!a.cell_ok) return;
  double f[3];
  for (int k = 0; k < 3; k++) f[k] = d[0] * a.inv[k] + d[1] * a.inv[3 + k] + d[2] * a.inv[6 + k];
  for (int k = 0; k < 3; k++) f[k] -= floor(f[k] + 0.5);
  for (int k = 0; k < 3; k++) d[k] = f[0] * a.cell[k] + f[1] * a.cell[3 + k] + f[2] * a.cell[6 + k];
}

__device__ __forceinline__ void neb_diffs(const NebArgs &a, const double *rp, const double *rc, const double *rn, int at,
                                          double (&dn)[3], double (&dp)[3]) {
  for (int k = 0; k < 3; k++) {
    dn[k] = rn[3 * at + k] - rc[3 * at + k];
    dp[k] = rc[3 * at + k] - rp[3 * at + k];
  }
  neb_mic(a, dn);
  neb_mic(a, dp);
}

// forceSpring of UniformSpring::compute (client/NEBSpringForce.cpp:24): ksp * pbc(dNext - dPrev)
__device__ __forceinline__ void neb_spring(const NebArgs &a, const double (&dn)[3], const double (&dp)[3], double (&s)[3]) {
  for (int k = 0; k < 3; k++) s[k] = dn[k] - dp[k];
  neb_mic(a, s);
  for (int k = 0; k < 3; k++) s[k] *= a.ksp;
}

struct NebImage {
  int imax, rule;  // highest movable image; tangent rule 0 = next, 1 = previous, 2 = energy-weighted mix
  double e_max, c_next, c_prev;
  bool climbs;
};

__device__ __forceinline__ NebImage neb_image(const NebArgs &a, const double *E, int img) {
  NebImage s;
  // highest movable image: std::max_element keeps the first of equal maxima (NudgedElasticBand.cpp:551-560)
  s.imax = 1;
  for (int j = 2; j <= a.images; j++)
    if (E[s.imax] < E[j]) s.imax = j;
  s.e_max = E[s.imax];
  s.climbs = a.ci_active && img == s.imax;
  // tangent rule (NEBTangent.cpp:44-66)
  const double e = E[img], e_prev = E[img - 1], e_next = E[img + 1];
  s.rule = 2;
  s.c_next = s.c_prev = 0.0;
  if (a.old_tangent || (e_next > e && e > e_prev)) {
    s.rule = 0;
  } else if (e > e_next && e_prev > e) {
    s.rule = 1;
  } else {
    const double dprev = e_prev - e, dnext = e_next - e;
    const double lo = fmin(fabs(dprev), fabs(dnext)), hi = fmax(fabs(dprev), fabs(dnext));
    s.c_next = dprev > dnext ? lo : hi;
    s.c_prev = dprev > dnext ? hi : lo;
  }
  return s;
}

// coefficient of the tangent in forceSpringPar (NEBSpringForce.cpp:23,35-36; spring constants :113-131)
__device__ __forceinline__ double neb_parallel(const NebArgs &a, const double *E, int img, double e_max, double dist_next,
                                               double dist_prev) {
  if (!a.weighted) return a.ksp * (dist_next - dist_prev);
  const int M = a.images;
  const double e_ref = fmin(E[0], E[M + 1]);
  const double range = e_max - e_ref;
  double kk[2];
  for (int s = 0; s < 2; s++) {
    const int idx = img + s;  // springConstants[idx - 1] sits between images idx - 1 and idx
    double kv = a.k_min;
    if (!(range < 1e-10)) {
      const double ei = fmax(E[idx], E[idx - 1]);
      if (ei > e_ref) {
        double al = (e_max - ei) / range;
        al = fmax(0.0, fmin(1.0, al));
        kv = (1.0 - al) * a.k_max + al * a.k_min;
      }
    }
    kk[s] = kv;  // kk[0] = springConstants[img - 1], kk[1] = springConstants[img]
  }
  return kk[1] * dist_next - kk[0] * dist_prev;
}

// Matter::getForces on a raw force row in place: fixed axes zeroed, or the mean force removed.  Every thread touches
// only the atoms it owns in the later passes (atom = thread + k * blockDim).
__device__ __forceinline__ void neb_matter_forces(const NebArgs &a, double *f, double *red) {
  if (!a.mask && !a.remove_net) return;
  const int N = a.atoms;
  double c[3] = {0.0, 0.0, 0.0};
  for (int at = threadIdx.x; at < N; at += blockDim.x)
    for (int k = 0; k < 3; k++) {
      double v = f[3 * at + k];
      if (a.mask) v *= a.mask[3 * at + k];
      f[3 * at + k] = v;
      c[k] += v;
    }
  if (a.remove_net) {
    block_sum<3>(c, red);
    for (int at = threadIdx.x; at < N; at += blockDim.x)
      for (int k = 0; k < 3; k++) f[3 * at + k] -= c[k] / static_cast<double>(N);
  }
}

__global__ void __launch_bounds__(256) k_neb_band(NebArgs a) {
  __shared__ double red[4 * kMaxWarps];
  const int M = a.images, N = a.atoms, D = 3 * N;
  const int band = blockIdx.x / M, img = blockIdx.x % M + 1;
  const double *E = a.energy + static_cast<size_t>(band) * (M + 2);
  const size_t row = static_cast<size_t>(band) * (M + 2) + img;
  const double *rc = a.pos + row * D, *rp = rc - D, *rn = rc + D;
  const double *fc = a.frc + row * D;
  double *T = a.tangent + (static_cast<size_t>(band) * M + img - 1) * D;
  double *P = a.projected + (static_cast<size_t>(band) * M + img - 1) * D;

  neb_matter_forces(a, const_cast<double *>(fc), red);
  const NebImage im = neb_image(a, E, img);
  const int rule = im.rule;
  const double c_next = im.c_next, c_prev = im.c_prev;
  const bool climbs = im.climbs;
  if (a.climbing && img == 1 && threadIdx.x == 0) a.climbing[band] = a.ci_active ? im.imax : 0;

  // pass 1: raw tangent, |dNext|^2, |dPrev|^2, |tangent|^2
  double acc[3] = {0.0, 0.0, 0.0};
  for (int at = threadIdx.x; at < N; at += blockDim.x) {
    double dn[3], dp[3];
    neb_diffs(a, rp, rc, rn, at, dn, dp);
    for (int k = 0; k < 3; k++) {
      const double t = rule == 0 ? dn[k] : (rule == 1 ? dp[k] : __dadd_rn(__dmul_rn(dn[k], c_next), __dmul_rn(dp[k], c_prev)));
      T[3 * at + k] = t;
      acc[0] += dn[k] * dn[k];
      acc[1] += dp[k] * dp[k];
      acc[2] += t * t;
    }
  }
  block_sum<3>(acc, red);
  const double dist_next = sqrt(acc[0]), dist_prev = sqrt(acc[1]);
  double t_norm = sqrt(acc[2]);
  const bool fallback = !(t_norm > 1e-10);  // NEBTangent.cpp:21-33
  if (fallback) t_norm = dist_next;
  const bool scale = t_norm > 1e-10;
  const bool want_spring = a.projection != NEB_PROJ_NEB;

  // pass 2: normalise; F.t and spring.t
  double dot[2] = {0.0, 0.0};
  for (int at = threadIdx.x; at < N; at += blockDim.x) {
    double dn[3], dp[3], s[3] = {0.0, 0.0, 0.0};
    if (fallback || want_spring) neb_diffs(a, rp, rc, rn, at, dn, dp);
    if (want_spring) neb_spring(a, dn, dp, s);
    for (int k = 0; k < 3; k++) {
      double t = fallback ? dn[k] : T[3 * at + k];
      if (scale) t /= t_norm;
      T[3 * at + k] = t;
      dot[0] += fc[3 * at + k] * t;
      dot[1] += s[k] * t;
    }
  }
  block_sum<2>(dot, red);
  const double f_t = dot[0], s_t = dot[1];

  // doubly nudged term (NEBForceProjection.cpp:73-97)
  double dneb_switch = 0.0, sp_fn = 0.0, f_perp_norm = 1.0;
  if (a.projection == NEB_PROJ_DNEB) {
    double q[3] = {0.0, 0.0, 0.0};
    for (int at = threadIdx.x; at < N; at += blockDim.x) {
      double dn[3], dp[3], s[3];
      neb_diffs(a, rp, rc, rn, at, dn, dp);
      neb_spring(a, dn, dp, s);
      for (int k = 0; k < 3; k++) {
        const double t = T[3 * at + k];
        const double sp = s[k] - s_t * t, fp = fc[3 * at + k] - f_t * t;
        q[0] += sp * sp;
        q[1] += fp * fp;
        q[2] += sp * fp;
      }
    }
    block_sum<3>(q, red);
    const double sp_norm = sqrt(q[0]);
    f_perp_norm = sqrt(q[1]);
    if (sp_norm > 1e-10 && f_perp_norm > 1e-10) {
      sp_fn = q[2] / f_perp_norm;
      dneb_switch = 2.0 / 3.14159265358979323846 * atan(f_perp_norm * f_perp_norm / (sp_norm * sp_norm));
    } else {
      f_perp_norm = 1.0;
    }
  }

  const double par = neb_parallel(a, E, img, im.e_max, dist_next, dist_prev);

  // pass 3: the projected force
  double col[3] = {0.0, 0.0, 0.0};
  for (int at = threadIdx.x; at < N; at += blockDim.x) {
    double dn[3], dp[3], s[3] = {0.0, 0.0, 0.0};
    if (want_spring) {
      neb_diffs(a, rp, rc, rn, at, dn, dp);
      neb_spring(a, dn, dp, s);
    }
    for (int k = 0; k < 3; k++) {
      const double t = T[3 * at + k], f = fc[3 * at + k];
      double dneb = 0.0;
      if (a.projection == NEB_PROJ_DNEB && dneb_switch != 0.0) {
        const double sp = s[k] - s_t * t, fpn = (f - f_t * t) / f_perp_norm;
        dneb = (sp - sp_fn * fpn) * dneb_switch;
      }
      double out;
      if (climbs) out = f - 2.0 * f_t * t + dneb;                       // NEBForceProjection.cpp:69-71
      else if (a.projection == NEB_PROJ_ELASTIC_BAND) out = s[k] + f;   // NEBProjection.cpp:17-19
      else out = par * t + (f - f_t * t) + dneb;                        // :21-31
      P[3 * at + k] = out;
      col[k] += out;
    }
  }
  if (a.all_free) {  // zeroTranslation (NEBForceProjection.cpp:101-110)
    block_sum<3>(col, red);
    for (int at = threadIdx.x; at < N; at += blockDim.x)
      for (int k = 0; k < 3; k++) P[3 * at + k] -= col[k] / static_cast<double>(N);
  }
}

// TMA-staged variant: one elected thread brings the three position rows (contiguous in HBM) and the force row into
// shared memory with two bulk copies (cp.async.bulk completing on an mbarrier), so a CTA has its whole 4 x 3 N x 8
// bytes in flight at once without holding them in registers; the differences, the tangent and the projected force then
// overwrite the staged rows in place and leave with one store each.  Rows start on 8-byte boundaries, so each copy
// starts at the 16-byte boundary below the row; a copy that would run past the end of the caller's array falls back
// to plain loads.  `total_doubles` is the length of the position / force arrays.
__host__ __device__ inline size_t neb_stage_bytes(int n_atoms) {
  return (static_cast<size_t>(4) * 3 * n_atoms + 10) * sizeof(double);
}

__global__ void __launch_bounds__(256) k_neb_band_tma(NebArgs a, size_t total_doubles) {
  EONB200_DYN_SMEM(smem_raw);
  __shared__ double red[4 * kMaxWarps];
  __shared__ __align__(8) uint64_t mbar[1];
  const int M = a.images, N = a.atoms, D = 3 * N;
  const int band = blockIdx.x / M, img = blockIdx.x % M + 1;
  const double *E = a.energy + static_cast<size_t>(band) * (M + 2);
  const size_t row = static_cast<size_t>(band) * (M + 2) + img;
  double *T = a.tangent + (static_cast<size_t>(band) * M + img - 1) * D;
  double *P = a.projected + (static_cast<size_t>(band) * M + img - 1) * D;

  double *spos = reinterpret_cast<double *>(smem_raw);  // 3 D + 2 doubles used
  double *sfrc = spos + ((3 * D + 5) & ~1);              // D + 2 doubles used; even offset keeps it 16-byte aligned
  const size_t p0 = (row - 1) * D, f0 = row * D;        // first double of the position block / the force row
  const int ppad = static_cast<int>(p0 & 1), fpad = static_cast<int>(f0 & 1);
  const size_t plen = (static_cast<size_t>(3) * D + ppad + 1) & ~static_cast<size_t>(1);
  const size_t flen = (static_cast<size_t>(D) + fpad + 1) & ~static_cast<size_t>(1);
  const bool bulk = ((reinterpret_cast<uintptr_t>(a.pos) | reinterpret_cast<uintptr_t>(a.frc)) & 15u) == 0 &&
                    p0 - ppad + plen <= total_doubles && f0 - fpad + flen <= total_doubles;
  if (threadIdx.x == 0) {
    mbar_init(mbar, 1);
    if (bulk) {
      mbar_expect(mbar, static_cast<uint32_t>((plen + flen) * 8));
      bulk_g2s_issue(spos, a.pos + (p0 - ppad), static_cast<uint32_t>(plen * 8), mbar);
      bulk_g2s_issue(sfrc, a.frc + (f0 - fpad), static_cast<uint32_t>(flen * 8), mbar);
    }
  }
  const NebImage im = neb_image(a, E, img);
  if (a.climbing && img == 1 && threadIdx.x == 0) a.climbing[band] = a.ci_active ? im.imax : 0;
  __syncthreads();  // mbarrier initialised
  if (bulk) {
    mbar_wait(mbar, 0);
  } else {
    for (int k = threadIdx.x; k < 3 * D; k += blockDim.x) spos[ppad + k] = a.pos[p0 + k];
    for (int k = threadIdx.x; k < D; k += blockDim.x) sfrc[fpad + k] = a.frc[f0 + k];
    __syncthreads();
  }
  double *rp = spos + ppad, *rc = rp + D, *rn = rc + D, *fc = sfrc + fpad;
  neb_matter_forces(a, fc, red);

  // pass 1: differences in place (rn <- dNext, rp <- dPrev), raw tangent into rc
  double acc[3] = {0.0, 0.0, 0.0};
  for (int at = threadIdx.x; at < N; at += blockDim.x) {
    double dn[3], dp[3];
    neb_diffs(a, rp, rc, rn, at, dn, dp);
    for (int k = 0; k < 3; k++) {
      const double t = im.rule == 0 ? dn[k] : (im.rule == 1 ? dp[k] : __dadd_rn(__dmul_rn(dn[k], im.c_next), __dmul_rn(dp[k], im.c_prev)));
      rn[3 * at + k] = dn[k];
      rp[3 * at + k] = dp[k];
      rc[3 * at + k] = t;
      acc[0] += dn[k] * dn[k];
      acc[1] += dp[k] * dp[k];
      acc[2] += t * t;
    }
  }
  block_sum<3>(acc, red);
  const double dist_next = sqrt(acc[0]), dist_prev = sqrt(acc[1]);
  double t_norm = sqrt(acc[2]);
  const bool fallback = !(t_norm > 1e-10);
  if (fallback) t_norm = dist_next;
  const bool scale = t_norm > 1e-10;
  const bool want_spring = a.projection != NEB_PROJ_NEB;

  // pass 2: normalised tangent (kept in rc, stored to HBM), F.t and spring.t
  double dot[2] = {0.0, 0.0};
  for (int at = threadIdx.x; at < N; at += blockDim.x) {
    double dn[3], dp[3], s[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < 3; k++) {
      dn[k] = rn[3 * at + k];
      dp[k] = rp[3 * at + k];
    }
    if (want_spring) neb_spring(a, dn, dp, s);
    for (int k = 0; k < 3; k++) {
      double t = fallback ? dn[k] : rc[3 * at + k];
      if (scale) t /= t_norm;
      rc[3 * at + k] = t;
      T[3 * at + k] = t;
      dot[0] += fc[3 * at + k] * t;
      dot[1] += s[k] * t;
    }
  }
  block_sum<2>(dot, red);
  const double f_t = dot[0], s_t = dot[1];

  double dneb_switch = 0.0, sp_fn = 0.0, f_perp_norm = 1.0;
  if (a.projection == NEB_PROJ_DNEB) {
    double q[3] = {0.0, 0.0, 0.0};
    for (int at = threadIdx.x; at < N; at += blockDim.x) {
      double dn[3], dp[3], s[3];
      for (int k = 0; k < 3; k++) {
        dn[k] = rn[3 * at + k];
        dp[k] = rp[3 * at + k];
      }
      neb_spring(a, dn, dp, s);
      for (int k = 0; k < 3; k++) {
        const double t = rc[3 * at + k];
        const double sp = s[k] - s_t * t, fp = fc[3 * at + k] - f_t * t;
        q[0] += sp * sp;
        q[1] += fp * fp;
        q[2] += sp * fp;
      }
    }
    block_sum<3>(q, red);
    const double sp_norm = sqrt(q[0]);
    f_perp_norm = sqrt(q[1]);
    if (sp_norm > 1e-10 && f_perp_norm > 1e-10) {
      sp_fn = q[2] / f_perp_norm;
      dneb_switch = 2.0 / 3.14159265358979323846 * atan(f_perp_norm * f_perp_norm / (sp_norm * sp_norm));
    } else {
      f_perp_norm = 1.0;
    }
  }
  const double par = neb_parallel(a, E, img, im.e_max, dist_next, dist_prev);

  // pass 3: the projected force; with zeroTranslation it waits in the force row for the column sums
  double col[3] = {0.0, 0.0, 0.0};
  for (int at = threadIdx.x; at < N; at += blockDim.x) {
    double dn[3], dp[3], s[3] = {0.0, 0.0, 0.0};
    if (want_spring) {
      for (int k = 0; k < 3; k++) {
        dn[k] = rn[3 * at + k];
        dp[k] = rp[3 * at + k];
      }
      neb_spring(a, dn, dp, s);
    }
    for (int k = 0; k < 3; k++) {
      const double t = rc[3 * at + k], f = fc[3 * at + k];
      double dneb = 0.0;
      if (a.projection == NEB_PROJ_DNEB && dneb_switch != 0.0) {
        const double sp = s[k] - s_t * t, fpn = (f - f_t * t) / f_perp_norm;
        dneb = (sp - sp_fn * fpn) * dneb_switch;
      }
      double out;
      if (im.climbs) out = f - 2.0 * f_t * t + dneb;
      else if (a.projection == NEB_PROJ_ELASTIC_BAND) out = s[k] + f;
      else out = par * t + (f - f_t * t) + dneb;
      if (a.all_free) fc[3 * at + k] = out;
      else P[3 * at + k] = out;
      col[k] += out;
    }
  }
  if (a.all_free) {
    block_sum<3>(col, red);
    for (int at = threadIdx.x; at < N; at += blockDim.x)
      for (int k = 0; k < 3; k++) P[3 * at + k] = fc[3 * at + k] - col[k] / static_cast<double>(N);
  }
}

}  // namespace eonb200
