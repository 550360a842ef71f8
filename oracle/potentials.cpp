// oracle/potentials.cpp -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
//
// Restated potentials and the brute-force neighbour predicate.  Compiled WITHOUT FMA
// contraction (-ffp-contract=off, no -march) so the arithmetic matches an x86-64 build of the
// reference (client/meson.build:237-249: -O3, native_arch / fast_math off).
#include "oracle.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

// Element rows of include/eon/potentials/EAM/Parameters.h:15-50 (Z, Dm, alphaM, Rm, beta1,
// beta2, r_cut, embedding polynomial c0..c8).  Physical constants, not code.
const OrcEamRow kBuiltinEam[3] = {
    {13, 3.63262, 1.49303, 2.12396, 3.45762, 3.54208, 5.555,
     {-0.0847305, 17.0887, 60.7909, -75.8412, 39.9948, -6.21563, -3.07993, 1.51365, -0.196034}},
    {46, 1.6054, 1.55083, 2.35604, 3.35148, 3.34715, 5.4120,
     {0.0992627, 4.8356, -0.368647, 17.2739, -18.450, 9.5018, -2.69876, 0.405341, -0.0251768}},
    {79, 0.669134, 1.89707, 2.57021, 3.69785, 3.69674, 5.5155,
     {-0.198791, -27.7663, 224.068, -594.101, 1316.31, -2070.62, 1973.14, -1012.02, 214.373}}};

struct Vec3 {
  double x, y, z;
};

// vesin Matrix::determinant / cross products, vesin-single-build.cpp:50-57,99-128,222-236.
Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
Vec3 normalize(Vec3 a) {
  double s = 1.0 / std::sqrt(dot(a, a));
  return {s * a.x, s * a.y, s * a.z};
}

void face_distances(const double *box, double w[3]) {
  Vec3 a{box[0], box[1], box[2]}, b{box[3], box[4], box[5]}, c{box[6], box[7], box[8]};
  w[0] = std::fabs(dot(normalize(cross(b, c)), a));
  w[1] = std::fabs(dot(normalize(cross(c, a)), b));
  w[2] = std::fabs(dot(normalize(cross(a, b)), c));
}

bool invert3(const double *m, double *inv) {
  double det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
               m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (std::fabs(det) < 1e-30) return false;
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

// Visit every (i, j, S) of vesin's FULL list: S integer on periodic axes, (i != j or S != 0),
// |(r_j - r_i) + S.H|^2 < cutoff^2 with the exact operation order of
// vesin-single-build.cpp:41-43,62-76,138-144,372-383,1199-1202.  Candidate enumeration is a
// plain triple loop over shifts around the minimum-image shift (vesin's cell walk only ever
// changes which candidates are looked at, never the predicate).  After the minimum-image shift
// the fractional separation is within [-0.5, 0.5], so an image k cells further is at least
// (k - 0.5) face distances away: k <= ceil(cutoff / w + 0.5) covers every image inside the cutoff.
template <typename F>
void visit_full(long n, const double *R, const double *box, const int periodic[3], double cutoff, F &&emit) {
  double inv[9];
  bool ok = invert3(box, inv);
  double w[3] = {1, 1, 1};
  if (ok) face_distances(box, w);
  int span[3];
  for (int k = 0; k < 3; k++) span[k] = (periodic[k] && ok) ? static_cast<int>(std::ceil(cutoff / w[k] + 0.5)) : 0;
  const double cutoff2 = cutoff * cutoff;
  std::vector<double> frac(3 * n, 0.0);
  if (ok) {
    for (long i = 0; i < n; i++)
      for (int k = 0; k < 3; k++)
        frac[3 * i + k] = R[3 * i] * inv[0 + k] + R[3 * i + 1] * inv[3 + k] + R[3 * i + 2] * inv[6 + k];
  }
  for (long i = 0; i < n; i++) {
    for (long j = 0; j < n; j++) {
      int base[3];
      for (int k = 0; k < 3; k++)
        base[k] = periodic[k] ? -static_cast<int>(std::floor(frac[3 * j + k] - frac[3 * i + k] + 0.5)) : 0;
      for (int sx = base[0] - span[0]; sx <= base[0] + span[0]; sx++)
        for (int sy = base[1] - span[1]; sy <= base[1] + span[1]; sy++)
          for (int sz = base[2] - span[2]; sz <= base[2] + span[2]; sz++) {
            if (i == j && sx == 0 && sy == 0 && sz == 0) continue;
            const double fx = static_cast<double>(sx), fy = static_cast<double>(sy), fz = static_cast<double>(sz);
            // Vector * Matrix, row-major box: v[k] = s0*H[0][k] + s1*H[1][k] + s2*H[2][k]
            const double cx = fx * box[0] + fy * box[3] + fz * box[6];
            const double cy = fx * box[1] + fy * box[4] + fz * box[7];
            const double cz = fx * box[2] + fy * box[5] + fz * box[8];
            const double vx = (R[3 * j] - R[3 * i]) + cx;
            const double vy = (R[3 * j + 1] - R[3 * i + 1]) + cy;
            const double vz = (R[3 * j + 2] - R[3 * i + 2]) + cz;
            const double d2 = vx * vx + vy * vy + vz * vz;
            if (d2 < cutoff2) emit(i, j, sx, sy, sz, vx, vy, vz, d2);
          }
    }
  }
}

// Half-list filter of vesin-single-build.cpp:1165-1196.
bool keep_in_half_list(long first, long second, int sx, int sy, int sz) {
  if (first > second) return false;
  if (first == second) {
    const int sum = sx + sy + sz;
    if (sum < 0) return false;
    if (sum == 0 && (sz < 0 || (sz == 0 && sy < 0))) return false;
  }
  return true;
}

const int kAllPeriodic[3] = {1, 1, 1};

// ---- pair potentials (rgpot forms, pinned by goldens; see oracle.h) ------------------------
// LJ: U = sum_{pairs, r<rc} 4 u0 [(psi/r)^12 - (psi/r)^6] - U(rc) (LJCluster: no shift).
// Morse: U = sum De (1 - exp(-a (r - re)))^2 - De - U(rc).
// Force on i from pair (i,j), d = r_i - r_j: F_i = -(dU/dr) d / r.
void pair_force(const OrcPotential *pot, long n, const double *R, double *F, double *U, const double *box) {
  for (long k = 0; k < 3 * n; k++) F[k] = 0.0;
  double energy = 0.0;
  const double rc = pot->cutoff;
  double shift = 0.0;
  if (pot->kind == ORC_POT_LJ) {
    const double s6 = std::pow(pot->lj_psi / rc, 6);
    shift = 4.0 * pot->lj_u0 * (s6 * s6 - s6);
  } else if (pot->kind == ORC_POT_MORSE_PT) {
    const double d = 1.0 - std::exp(-pot->morse_a * (rc - pot->morse_re));
    shift = pot->morse_De * d * d - pot->morse_De;
  }
  visit_full(n, R, box, kAllPeriodic, rc,
             [&](long i, long j, int sx, int sy, int sz, double vx, double vy, double vz, double d2) {
               if (!keep_in_half_list(i, j, sx, sy, sz)) return;
               const double r = std::sqrt(d2);
               double e, dudr;
               if (pot->kind == ORC_POT_MORSE_PT) {
                 const double ex = std::exp(-pot->morse_a * (r - pot->morse_re));
                 const double d = 1.0 - ex;
                 e = pot->morse_De * d * d - pot->morse_De;
                 dudr = 2.0 * pot->morse_De * pot->morse_a * d * ex;
               } else {
                 const double sr2 = (pot->lj_psi * pot->lj_psi) / d2;
                 const double sr6 = sr2 * sr2 * sr2;
                 e = 4.0 * pot->lj_u0 * (sr6 * sr6 - sr6);
                 dudr = -24.0 * pot->lj_u0 * (2.0 * sr6 * sr6 - sr6) / r;
               }
               energy += e - shift;
               // vector = r_j - r_i (+ image); F_i = +dU/dr * vector / r, F_j = -F_i
               const double s = dudr / r;
               F[3 * i] += s * vx;
               F[3 * i + 1] += s * vy;
               F[3 * i + 2] += s * vz;
               if (i != j) {
                 F[3 * j] -= s * vx;
                 F[3 * j + 1] -= s * vy;
                 F[3 * j + 2] -= s * vz;
               }
             });
  *U = energy;
}

// ---- EAM (client/potentials/EAM/EAM.cpp) ---------------------------------------------------
const OrcEamRow *find_row(const OrcPotential *pot, int z) {
  const OrcEamRow *rows = pot->n_eam_rows > 0 ? pot->eam_rows : kBuiltinEam;
  const int nrows = pot->n_eam_rows > 0 ? pot->n_eam_rows : 3;
  for (int r = 0; r < nrows; r++)
    if (rows[r].Z == z) return &rows[r];
  return nullptr;  // EAM.cpp:116-123 throws 14324
}

// EAM.cpp:349-356 / :358-365 (note the sign: embedding_force returns -dF/drho).
double embed(const double *c, double rho) {
  double acc = c[8];
  for (int k = 7; k >= 0; k--) acc = acc * rho + c[k];
  return acc;
}
double embed_force(const double *c, double rho) {
  double acc = c[8] * 8;
  for (int k = 7; k >= 1; k--) acc = acc * rho + k * c[k];
  return -acc;
}

// EAM.cpp:31,56-90: the cell diagonal only; shift so no coordinate is negative; wrap > box.
void eam_prepare(long n, const double *R, const double *fullbox, double box[3], std::vector<double> &Rw) {
  box[0] = fullbox[0];
  box[1] = fullbox[4];
  box[2] = fullbox[8];
  double lo[3] = {std::numeric_limits<double>::max(), std::numeric_limits<double>::max(),
                  std::numeric_limits<double>::max()};
  Rw.assign(R, R + 3 * n);
  for (long i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) lo[k] = std::min(lo[k], R[3 * i + k]);
  for (int k = 0; k < 3; k++)
    if (lo[k] > 0) lo[k] = 0;
  for (long i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) Rw[3 * i + k] += std::fabs(lo[k]);
  for (long q = 0; q < 3 * n; q++)
    while (Rw[q] > box[q % 3]) Rw[q] -= box[q % 3];
}

inline void eam_mic(double &d, double box, double half) {
  if (d > half)
    d -= box;
  else if (d < -half)
    d += box;
}

// EAM.cpp:125-226, same loop nest and accumulation order (so the result is bit-identical to
// the compiled reference): for every i, every j != i inside rc_i contributes density and
// d(rho_i)/dR to a 3N scratch; the Morse pair part only for j > i with the row of atom i; after
// the j loop the embedding energy is added and F -= F'(rho_i) * drho over ALL 3N entries.
int eam_force_n2(const OrcPotential *pot, long n, const double *Rin, const int *Z, double *F, double *U,
                 const double *fullbox) {
  double box[3];
  std::vector<double> R;
  eam_prepare(n, Rin, fullbox, box, R);
  std::vector<double> drho(3 * n);
  const double pair_sign = pot->eam_conservative ? -1.0 : 1.0;
  for (long k = 0; k < 3 * n; k++) F[k] = 0.0;
  *U = 0;
  const double hx = box[0] * 0.5, hy = box[1] * 0.5, hz = box[2] * 0.5;
  for (long i = 0; i < n; i++) {
    std::fill(drho.begin(), drho.end(), 0.0);
    const OrcEamRow *p = find_row(pot, Z[i]);
    if (!p) return 14324;
    const double cutoff2 = p->r_cut * p->r_cut;
    double dens = 0.0;
    for (long j = 0; j < n; j++) {
      if (i == j) continue;
      double dx = R[3 * i] - R[3 * j], dy = R[3 * i + 1] - R[3 * j + 1], dz = R[3 * i + 2] - R[3 * j + 2];
      eam_mic(dx, box[0], hx);
      eam_mic(dy, box[1], hy);
      eam_mic(dz, box[2], hz);
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 > cutoff2) continue;
      const double r = std::sqrt(r2);
      const double r6 = r2 * r2 * r2;
      const double e1 = std::exp(-p->beta1 * r);
      const double e2 = std::exp(-2.0 * p->beta2 * r);
      const double rho_pair = e1 + 512.0 * e2;
      dens += r6 * rho_pair;
      const double r5 = r2 * r2 * r;
      const double dden = 6.0 * r5 * rho_pair + r6 * (-p->beta1 * e1 - 1024.0 * p->beta2 * e2);
      const double invr = 1.0 / r;
      const double fs = -dden * invr;
      drho[3 * i] += fs * dx;
      drho[3 * i + 1] += fs * dy;
      drho[3 * i + 2] += fs * dz;
      drho[3 * j] -= fs * dx;
      drho[3 * j + 1] -= fs * dy;
      drho[3 * j + 2] -= fs * dz;
      if (j > i) {
        const double ex = std::exp(-p->alphaM * (r - p->Rm));
        const double d = 1.0 - ex;
        const double phi = p->Dm * d * d - p->Dm;
        const double mag = 2.0 * p->alphaM * p->Dm * d * (d - 1.0);
        // EAM.cpp:203-215 applies F_i -= (mag/r) d with mag = -dphi/dr, i.e. the pair force enters
        // with the opposite sign of -grad(U) (finite differences of the reference's own energy
        // confirm it).  pair_sign = +1 reproduces the reference bit for bit; eam_conservative flips
        // it so that F = -grad(U), which is what a saddle search needs to converge.
        const double fc = pair_sign * mag * invr;
        F[3 * i] -= fc * dx;
        F[3 * i + 1] -= fc * dy;
        F[3 * i + 2] -= fc * dz;
        F[3 * j] += fc * dx;
        F[3 * j + 1] += fc * dy;
        F[3 * j + 2] += fc * dz;
        *U += phi;
      }
    }
    *U += embed(p->func_coeff, dens);
    const double dF = embed_force(p->func_coeff, dens);
    for (long k = 0; k < 3 * n; k++) F[k] -= dF * drho[k];
  }
  return 0;
}

// "oracle-NL": the same EAM arithmetic per pair, but candidates from a uniform cell grid and a
// per-atom gather (density pass, then force pass) instead of the reference's O(N^2) loop and
// O(N^2) axpy.  This is the CPU baseline that does not flatter the GPU.
// Valid when every box length is >= 2*max r_cut (single minimum image), otherwise falls back.
int eam_force_nl(const OrcPotential *pot, long n, const double *Rin, const int *Z, double *F, double *U,
                 const double *fullbox) {
  double box[3];
  std::vector<double> R;
  eam_prepare(n, Rin, fullbox, box, R);
  double rcmax = 0.0;
  std::vector<const OrcEamRow *> row(n);
  for (long i = 0; i < n; i++) {
    row[i] = find_row(pot, Z[i]);
    if (!row[i]) return 14324;
    rcmax = std::max(rcmax, row[i]->r_cut);
  }
  int nc[3];
  for (int k = 0; k < 3; k++) {
    if (!(box[k] >= 2.0 * rcmax)) return eam_force_n2(pot, n, Rin, Z, F, U, fullbox);
    nc[k] = std::max(1, static_cast<int>(box[k] / rcmax));
    if (nc[k] < 3) nc[k] = 1;  // fewer than 3 cells: neighbour cells alias; use one cell on that axis
  }
  const long ncell = static_cast<long>(nc[0]) * nc[1] * nc[2];
  std::vector<long> head(ncell + 1, 0), cell_of(n), order(n);
  for (long i = 0; i < n; i++) {
    int c[3];
    for (int k = 0; k < 3; k++) {
      c[k] = static_cast<int>(R[3 * i + k] / box[k] * nc[k]);
      c[k] = std::min(std::max(c[k], 0), nc[k] - 1);
    }
    cell_of[i] = (static_cast<long>(c[0]) * nc[1] + c[1]) * nc[2] + c[2];
    head[cell_of[i] + 1]++;
  }
  for (long c = 0; c < ncell; c++) head[c + 1] += head[c];
  {
    std::vector<long> fill(head.begin(), head.end() - 1);
    for (long i = 0; i < n; i++) order[fill[cell_of[i]]++] = i;
  }
  // neighbour rows (CSR), j ascending inside a row so the summation order is well defined
  std::vector<long> nstart(n + 1, 0);
  std::vector<long> nbr;
  nbr.reserve(64 * n);
  const double hx = box[0] * 0.5, hy = box[1] * 0.5, hz = box[2] * 0.5;
  std::vector<long> tmp;
  for (long i = 0; i < n; i++) {
    tmp.clear();
    long ci = cell_of[i];
    int c[3] = {static_cast<int>(ci / (nc[1] * nc[2])), static_cast<int>((ci / nc[2]) % nc[1]),
                static_cast<int>(ci % nc[2])};
    int lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
      lo[k] = nc[k] == 1 ? 0 : -1;
      hi[k] = nc[k] == 1 ? 0 : 1;
    }
    for (int ax = lo[0]; ax <= hi[0]; ax++)
      for (int ay = lo[1]; ay <= hi[1]; ay++)
        for (int az = lo[2]; az <= hi[2]; az++) {
          const int qx = (c[0] + ax + nc[0]) % nc[0], qy = (c[1] + ay + nc[1]) % nc[1],
                    qz = (c[2] + az + nc[2]) % nc[2];
          const long cj = (static_cast<long>(qx) * nc[1] + qy) * nc[2] + qz;
          for (long s = head[cj]; s < head[cj + 1]; s++) {
            const long j = order[s];
            if (j == i) continue;
            double dx = R[3 * i] - R[3 * j], dy = R[3 * i + 1] - R[3 * j + 1], dz = R[3 * i + 2] - R[3 * j + 2];
            eam_mic(dx, box[0], hx);
            eam_mic(dy, box[1], hy);
            eam_mic(dz, box[2], hz);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 <= rcmax * rcmax) tmp.push_back(j);
          }
        }
    std::sort(tmp.begin(), tmp.end());
    nbr.insert(nbr.end(), tmp.begin(), tmp.end());
    nstart[i + 1] = static_cast<long>(nbr.size());
  }
  // pass 1: densities with the row of atom i (EAM.cpp:140,172-183)
  std::vector<double> dFd(n);
  double energy = 0.0;
  for (long i = 0; i < n; i++) {
    const OrcEamRow *p = row[i];
    const double cutoff2 = p->r_cut * p->r_cut;
    double dens = 0.0;
    for (long s = nstart[i]; s < nstart[i + 1]; s++) {
      const long j = nbr[s];
      double dx = R[3 * i] - R[3 * j], dy = R[3 * i + 1] - R[3 * j + 1], dz = R[3 * i + 2] - R[3 * j + 2];
      eam_mic(dx, box[0], hx);
      eam_mic(dy, box[1], hy);
      eam_mic(dz, box[2], hz);
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 > cutoff2) continue;
      const double r = std::sqrt(r2);
      dens += (r2 * r2 * r2) * (std::exp(-p->beta1 * r) + 512.0 * std::exp(-2.0 * p->beta2 * r));
    }
    energy += embed(p->func_coeff, dens);
    dFd[i] = -embed_force(p->func_coeff, dens);  // +dF/drho
  }
  // pass 2: F_k = -sum_j [F'_k g'_k(r) + F'_j g'_j(r) (r <= rc_j)] d/r - pair'(r; row of min(k,j)) d/r
  for (long k = 0; k < n; k++) {
    double fx = 0, fy = 0, fz = 0;
    for (long s = nstart[k]; s < nstart[k + 1]; s++) {
      const long j = nbr[s];
      double dx = R[3 * k] - R[3 * j], dy = R[3 * k + 1] - R[3 * j + 1], dz = R[3 * k + 2] - R[3 * j + 2];
      eam_mic(dx, box[0], hx);
      eam_mic(dy, box[1], hy);
      eam_mic(dz, box[2], hz);
      const double r2 = dx * dx + dy * dy + dz * dz;
      const double r = std::sqrt(r2);
      const double invr = 1.0 / r;
      double coef = 0.0;  // dE/dr along the pair
      const OrcEamRow *pk = row[k], *pj = row[j];
      const double r5 = r2 * r2 * r, r6 = r2 * r2 * r2;
      if (r2 <= pk->r_cut * pk->r_cut) {
        const double e1 = std::exp(-pk->beta1 * r), e2 = std::exp(-2.0 * pk->beta2 * r);
        coef += dFd[k] * (6.0 * r5 * (e1 + 512.0 * e2) + r6 * (-pk->beta1 * e1 - 1024.0 * pk->beta2 * e2));
      }
      if (r2 <= pj->r_cut * pj->r_cut) {
        const double e1 = std::exp(-pj->beta1 * r), e2 = std::exp(-2.0 * pj->beta2 * r);
        coef += dFd[j] * (6.0 * r5 * (e1 + 512.0 * e2) + r6 * (-pj->beta1 * e1 - 1024.0 * pj->beta2 * e2));
      }
      const OrcEamRow *pm = (k < j) ? pk : pj;  // Morse part uses the row of the lower index (EAM.cpp:200)
      if (r2 <= pm->r_cut * pm->r_cut) {
        const double ex = std::exp(-pm->alphaM * (r - pm->Rm));
        const double d = 1.0 - ex;
        // reference sign quirk (see eam_force_n2): the pair term enters with the wrong sign
        // unless eam_conservative is set
        coef += (pot->eam_conservative ? 1.0 : -1.0) * 2.0 * pm->alphaM * pm->Dm * d * ex;
        if (k < j) energy += pm->Dm * d * d - pm->Dm;
      }
      const double s3 = -coef * invr;
      fx += s3 * dx;
      fy += s3 * dy;
      fz += s3 * dz;
    }
    F[3 * k] = fx;
    F[3 * k + 1] = fy;
    F[3 * k + 2] = fz;
  }
  *U = energy;
  return 0;
}

}  // namespace

extern "C" {

void orc_potential_default(OrcPotential *pot, int kind) {
  std::memset(pot, 0, sizeof(*pot));
  pot->kind = kind;
  pot->lj_u0 = 1.0;
  pot->lj_psi = 1.0;
  pot->morse_De = 0.7102;
  pot->morse_a = 1.6047;
  pot->morse_re = 2.8970;
  if (kind == ORC_POT_LJ || kind == ORC_POT_LJCLUSTER) pot->cutoff = 15.0;
  if (kind == ORC_POT_MORSE_PT) pot->cutoff = 9.5;
}

int orc_force(const OrcPotential *pot, long n, const double *R, const int *Z, double *F, double *U,
              const double *box) {
  switch (pot->kind) {
    case ORC_POT_LJ:
    case ORC_POT_LJCLUSTER:
    case ORC_POT_MORSE_PT:
      pair_force(pot, n, R, F, U, box);
      return 0;
    case ORC_POT_EAM:
      return eam_force_n2(pot, n, R, Z, F, U, box);
    case ORC_POT_EAM_NL:
      return eam_force_nl(pot, n, R, Z, F, U, box);
  }
  return -1;
}

long orc_count_ordered_pairs(const OrcPotential *pot, long n, const double *R, const int *Z, const double *box) {
  long count = 0;
  if (pot->kind == ORC_POT_EAM || pot->kind == ORC_POT_EAM_NL) {
    double b[3];
    std::vector<double> Rw;
    eam_prepare(n, R, box, b, Rw);
    for (long i = 0; i < n; i++) {
      const OrcEamRow *p = find_row(pot, Z[i]);
      if (!p) return -1;
      for (long j = 0; j < n; j++) {
        if (i == j) continue;
        double dx = Rw[3 * i] - Rw[3 * j], dy = Rw[3 * i + 1] - Rw[3 * j + 1], dz = Rw[3 * i + 2] - Rw[3 * j + 2];
        eam_mic(dx, b[0], b[0] * 0.5);
        eam_mic(dy, b[1], b[1] * 0.5);
        eam_mic(dz, b[2], b[2] * 0.5);
        if (dx * dx + dy * dy + dz * dz <= p->r_cut * p->r_cut) count++;
      }
    }
    return count;
  }
  visit_full(n, R, box, kAllPeriodic, pot->cutoff,
             [&](long, long, int, int, int, double, double, double, double) { count++; });
  return count;
}

long orc_neighbors(long n, const double *points, const double *box, const int periodic[3], double cutoff,
                   int full, long cap, int64_t *pairs, int32_t *shifts) {
  long count = 0;
  visit_full(n, points, box, periodic, cutoff,
             [&](long i, long j, int sx, int sy, int sz, double, double, double, double) {
               if (!full && !keep_in_half_list(i, j, sx, sy, sz)) return;
               if (count < cap) {
                 pairs[2 * count] = i;
                 pairs[2 * count + 1] = j;
                 shifts[3 * count] = sx;
                 shifts[3 * count + 1] = sy;
                 shifts[3 * count + 2] = sz;
               }
               count++;
             });
  return count;
}

}  // extern "C"
