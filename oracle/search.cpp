// oracle/search.cpp -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
//
// Restatement (no Eigen) of the reference's min-mode saddle search stack:
//   Matter force cache / PBC / fixed mask   client/Matter.cpp:25-72,273-357,498-538; include/eon/Matter.h:42-63
//   MinModeObjectiveFunction                client/MinModeSaddleSearch.cpp:51-164
//   MinModeSaddleSearch::run                client/MinModeSaddleSearch.cpp:211-475
//   ImprovedDimer::compute                  client/ImprovedDimer.cpp:63-434
//   Lanczos::compute                        client/Lanczos.cpp:46-173; client/MobileAtoms.cpp:47-135
//   ConjugateGradients                      client/ConjugateGradients.cpp:17-43,114-199
//   LBFGS                                   client/LBFGS.cpp:19-175
//   FIRE                                    client/FIRE.cpp:19-75; include/eon/FIRE.h:22-35
//   relaxMatter + MatterObjectiveFunction   client/HelperFunctions.cpp:273-410
//   maxAtomMotion*, safe math               client/GeometryAnalysis.cpp:269-327; include/eon/SafeMath.h:19-80
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <vector>

namespace {

using Vec = std::vector<double>;
constexpr double kPi = 3.14159265358979323846;
constexpr double kEps = 1e-300;  // SafeMath.h:19

double safe_div(double num, double den, double fb = 0.0) { return std::fabs(den) < kEps ? fb : num / den; }
double safe_recip(double x, double fb = 0.0) { return safe_div(1.0, x, fb); }
double safe_acos(double x) { return std::acos(std::min(1.0, std::max(-1.0, x))); }
double safe_atan_ratio(double num, double den, double fb = 0.0) {
  return std::fabs(den) < kEps ? fb : std::atan(num / den);
}

double vdot(const Vec &a, const Vec &b) {
  double s = 0.0;
  for (size_t k = 0; k < a.size(); k++) s += a[k] * b[k];
  return s;
}
double vnorm(const Vec &a) { return std::sqrt(vdot(a, a)); }
Vec vscale(const Vec &a, double s) {
  Vec r(a);
  for (double &x : r) x *= s;
  return r;
}
Vec vadd(const Vec &a, const Vec &b, double sb = 1.0) {
  Vec r(a);
  for (size_t k = 0; k < r.size(); k++) r[k] += sb * b[k];
  return r;
}
Vec vsub(const Vec &a, const Vec &b) { return vadd(a, b, -1.0); }
Vec safe_normalized(const Vec &a) {
  const double n = vnorm(a);
  if (n < kEps) return Vec(a.size(), 0.0);
  return vscale(a, 1.0 / n);  // Eigen: v / n
}
void safe_normalize_inplace(Vec &a) {
  const double n = vnorm(a);
  if (n < kEps) {
    std::fill(a.begin(), a.end(), 0.0);
  } else {
    for (double &x : a) x /= n;
  }
}

// GeometryAnalysis.cpp:273-296,321-327
double max_atom_motion_v(const Vec &v) {
  const long n = static_cast<long>(v.size());
  if (n < 3) return vnorm(v);
  double mx = 0.0;
  for (long i = 0; i + 3 <= n; i += 3) {
    const double nn = std::sqrt(v[i] * v[i] + v[i + 1] * v[i + 1] + v[i + 2] * v[i + 2]);
    if (mx < nn) mx = nn;
  }
  return mx;
}
Vec max_atom_motion_applied_v(const Vec &v, double max_motion) {
  const double mx = max_atom_motion_v(v);
  if (mx > max_motion) return vscale(v, max_motion / mx);
  return v;
}

struct ForceCounter {
  long calls = 0;
};

using PotFn = std::function<int(long, const double *, const int *, double *, double *, const double *)>;

// ---------------------------------------------------------------------------------------------
// Matter: positions + lazily evaluated force cache (Matter.cpp:498-538).  Copy-assignment copies
// the cache AND the dirty flag (Matter.cpp:25-72).
struct Topology {
  long n = 0;
  const int *Z = nullptr;
  Vec free_mask;             // 3N, 1.0 free / 0.0 fixed (Matter::getFree)
  std::vector<long> free_atoms;  // atoms with at least one free axis (!getFixed(i))
  bool any_fixed_axis = false;
  double cell[9];
  double inv[9];
  bool cell_ok = false;
  bool remove_net_force = true;
  PotFn pot;
  ForceCounter *counter = nullptr;
};

bool invert3(const double *m, double *inv) {
  const double det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
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

class Matter {
public:
  explicit Matter(std::shared_ptr<const Topology> topo) : t(std::move(topo)) {
    pos.assign(3 * t->n, 0.0);
    frc.assign(3 * t->n, 0.0);
  }

  // Matter::setPositions + applyPeriodicBoundary, Legacy convention (Matter.h:54-63)
  void setPositionsV(const Vec &x) {
    pos = x;
    wrap();
    dirty = true;
  }
  void setPositionsFreeV(const Vec &x) {
    for (size_t a = 0; a < t->free_atoms.size(); a++)
      for (int k = 0; k < 3; k++) pos[3 * t->free_atoms[a] + k] = x[3 * a + k];
    wrap();
    dirty = true;
  }
  const Vec &getPositionsV() const { return pos; }
  Vec getPositionsFreeV() const { return pack_free(pos); }

  double getPotentialEnergy() {
    compute();
    return energy;
  }
  // Matter::getForces: raw forces times the per-axis free mask (Matter.cpp:324-333)
  Vec getForcesV() {
    compute();
    Vec f(frc);
    for (size_t k = 0; k < f.size(); k++) f[k] *= t->free_mask[k];
    return f;
  }
  Vec getForcesFreeV() { return pack_free(getForcesV()); }
  double maxForce() {
    const Vec f = getForcesV();
    double mx = 0.0;
    for (long i = 0; i < t->n; i++)
      mx = std::max(mx, std::sqrt(f[3 * i] * f[3 * i] + f[3 * i + 1] * f[3 * i + 1] + f[3 * i + 2] * f[3 * i + 2]));
    return mx;
  }
  // pbc::apply on a difference (Matter.h:42-50)
  Vec pbcV(const Vec &d) const {
    if (!t->cell_ok) return d;
    Vec r(d.size());
    for (size_t i = 0; i + 3 <= d.size(); i += 3) {
      double f[3];
      for (int k = 0; k < 3; k++) f[k] = d[i] * t->inv[k] + d[i + 1] * t->inv[3 + k] + d[i + 2] * t->inv[6 + k];
      for (int k = 0; k < 3; k++) f[k] -= std::floor(f[k] + 0.5);
      for (int k = 0; k < 3; k++) r[i + k] = f[0] * t->cell[k] + f[1] * t->cell[3 + k] + f[2] * t->cell[6 + k];
    }
    return r;
  }
  Vec pack_free(const Vec &full) const {
    Vec r(3 * t->free_atoms.size());
    for (size_t a = 0; a < t->free_atoms.size(); a++)
      for (int k = 0; k < 3; k++) r[3 * a + k] = full[3 * t->free_atoms[a] + k];
    return r;
  }
  long degreesFree() const { return 3 * static_cast<long>(t->free_atoms.size()); }
  long nAtoms() const { return t->n; }
  std::shared_ptr<const Topology> topo() const { return t; }

  bool failed = false;

private:
  void wrap() {
    if (!t->cell_ok) return;
    for (long i = 0; i < t->n; i++) {
      double f[3];
      for (int k = 0; k < 3; k++)
        f[k] = pos[3 * i] * t->inv[k] + pos[3 * i + 1] * t->inv[3 + k] + pos[3 * i + 2] * t->inv[6 + k];
      for (int k = 0; k < 3; k++) f[k] = std::fmod(f[k] + 1.0, 1.0);
      for (int k = 0; k < 3; k++) pos[3 * i + k] = f[0] * t->cell[k] + f[1] * t->cell[3 + k] + f[2] * t->cell[6 + k];
    }
  }
  void compute() {
    if (!dirty) return;
    const int rc = t->pot(t->n, pos.data(), t->Z, frc.data(), &energy, t->cell);
    if (rc != 0) throw std::runtime_error("potential failed");
    if (t->counter) t->counter->calls++;
    dirty = false;
    // Matter.cpp:531-536: only when NO axis of ANY atom is fixed
    if (!t->any_fixed_axis && t->remove_net_force) {
      double m[3] = {0, 0, 0};
      for (long i = 0; i < t->n; i++)
        for (int k = 0; k < 3; k++) m[k] += frc[3 * i + k];
      for (int k = 0; k < 3; k++) m[k] /= static_cast<double>(t->n);
      for (long i = 0; i < t->n; i++)
        for (int k = 0; k < 3; k++) frc[3 * i + k] -= m[k];
    }
  }

  std::shared_ptr<const Topology> t;
  Vec pos, frc;
  double energy = 0.0;
  bool dirty = true;
};

// ---------------------------------------------------------------------------------------------
struct ObjectiveFunction {
  virtual ~ObjectiveFunction() = default;
  virtual double getEnergy() = 0;
  virtual Vec getGradient(bool fdstep = false) = 0;
  virtual void setPositions(const Vec &x) = 0;
  virtual Vec getPositions() = 0;
  virtual int degreesOfFreedom() = 0;
  virtual bool isConverged() = 0;
  virtual double getConvergence() = 0;
  virtual Vec difference(const Vec &a, const Vec &b) = 0;
};

struct Optimizer {
  virtual ~Optimizer() = default;
  virtual int step(double maxMove) = 0;
};

// ConjugateGradients.cpp:17-43,114-190 (single_step; line_search is off by default and not restated)
class CG : public Optimizer {
public:
  CG(ObjectiveFunction *o, const OrcSearchParams &p) : objf(o), prm(p) {
    const size_t d = objf->getPositions().size();
    directionOld.assign(d, 0.0);
    forceOld.assign(d, 0.0);
  }
  int step(double maxMove) override {
    force = vscale(objf->getGradient(), -1.0);
    Vec pos = objf->getPositions();
    getStep();
    Vec posStep = vadd(pos, directionNorm, prm.finite_difference);
    objf->setPositions(posStep);
    Vec forceAfter = vscale(objf->getGradient(true), -1.0);
    const double pf1 = vdot(force, directionNorm);
    double pf2 = vdot(forceAfter, directionNorm);
    const double curvature = safe_div(pf1 - pf2, prm.finite_difference, 0.0);
    double stepSize = maxMove;
    if (curvature > 0.0) stepSize = safe_div(pf1, curvature, maxMove);
    if (!prm.cg_no_overshooting) {
      pos = vadd(pos, max_atom_motion_applied_v(vscale(directionNorm, stepSize), maxMove));
      objf->setPositions(pos);
    } else {
      double passedMinimum = -1.;
      while (passedMinimum < 0.0 && 0.1 * std::fabs(pf1) < std::fabs(pf2)) {
        posStep = vadd(pos, max_atom_motion_applied_v(vscale(directionNorm, stepSize), maxMove));
        objf->setPositions(posStep);
        forceAfter = vscale(objf->getGradient(true), -1.0);
        pf2 = vdot(forceAfter, directionNorm);
        passedMinimum = pf1 * pf2;
        if (passedMinimum < 0.0 && 0.1 * std::fabs(pf1) < std::fabs(pf2)) {
          const double forceChange = pf1 - pf2;
          stepSize = safe_div(pf1, forceChange, 0.0) * stepSize;
        }
      }
    }
    if (prm.cg_knock_out_max_move && stepSize >= maxMove) {
      std::fill(directionOld.begin(), directionOld.end(), 0.0);
      std::fill(forceOld.begin(), forceOld.end(), 0.0);
    }
    return objf->isConverged() ? 1 : 0;
  }

private:
  void getStep() {
    const double a = std::fabs(vdot(force, forceOld));
    const double b = vdot(forceOld, forceOld);
    double gamma = 0.0;
    if (a < 0.5 * b) gamma = safe_div(vdot(force, vsub(force, forceOld)), b, 0.0);
    direction = vadd(force, directionOld, gamma);
    directionNorm = direction;
    safe_normalize_inplace(directionNorm);
    directionOld = direction;
    forceOld = force;
    if (prm.cg_max_iter_before_reset > 0 && prm.cg_max_iter_before_reset <= cg_i) {
      cg_i = 0;
      std::fill(forceOld.begin(), forceOld.end(), 0.0);
      std::fill(directionOld.begin(), directionOld.end(), 0.0);
    }
    cg_i += 1;
  }
  ObjectiveFunction *objf;
  OrcSearchParams prm;
  Vec force, forceOld, direction, directionOld, directionNorm;
  long cg_i = 0;
};

// LBFGS.cpp:19-165
class LBFGS : public Optimizer {
public:
  LBFGS(ObjectiveFunction *o, const OrcSearchParams &p)
      : objf(o), prm(p), memory(std::min<long>(o->degreesOfFreedom(), p.lbfgs_memory)) {}
  int step(double maxMove) override {
    Vec r = objf->getPositions();
    Vec f = vscale(objf->getGradient(), -1.0);
    if (iteration > 0) update(r, rPrev, f, fPrev);
    Vec dr = getStep(maxMove, f);
    objf->setPositions(vadd(r, dr));
    rPrev = r;
    fPrev = f;
    iteration++;
    return objf->isConverged() ? 1 : 0;
  }

private:
  void reset() {
    s.clear();
    y.clear();
    rho.clear();
  }
  void update(const Vec &r1, const Vec &r0, const Vec &f1, const Vec &f0) {
    Vec s0 = objf->difference(r1, r0);
    Vec y0 = vsub(f0, f1);
    const double sy = vdot(s0, y0);
    if (std::fabs(sy) < 1e-30) {  // LBFGS_EPS, LBFGS.h:24
      reset();
      return;
    }
    rho.push_back(safe_recip(sy, 0.0));
    s.push_back(std::move(s0));
    y.push_back(std::move(y0));
    if (static_cast<long>(s.size()) > memory) {
      s.pop_front();
      y.pop_front();
      rho.pop_front();
    }
  }
  Vec getStep(double maxMove, const Vec &f) {
    double H0 = prm.lbfgs_inverse_curvature;
    Vec r = objf->getPositions();
    if (iteration > 0) {
      Vec dr = objf->difference(r, rPrev);
      Vec df = vsub(fPrev, f);
      const double C = safe_div(vdot(df, df), vdot(dr, df), -1.0);
      if (C < 0) {
        reset();
        return max_atom_motion_applied_v(vscale(f, 1000.0), maxMove);
      }
      if (prm.lbfgs_auto_scale) H0 = safe_recip(C, -1.0);
    }
    if (iteration == 0 && prm.lbfgs_auto_scale) {
      const Vec fhat = safe_normalized(f);
      objf->setPositions(vadd(r, fhat, prm.finite_difference));
      Vec dg = vadd(objf->getGradient(true), f);
      const double C = vdot(dg, fhat) / prm.finite_difference;
      H0 = safe_recip(C, -1.0);
      objf->setPositions(r);
      if (H0 < 0) {
        reset();
        return max_atom_motion_applied_v(vscale(f, 1000.0), maxMove);
      }
    }
    const int loopmax = static_cast<int>(s.size());
    std::vector<double> a(loopmax);
    Vec q = vscale(f, -1.0);
    for (int i = loopmax - 1; i >= 0; i--) {
      a[i] = rho[i] * vdot(s[i], q);
      q = vadd(q, y[i], -a[i]);
    }
    Vec z = vscale(q, H0);
    for (int i = 0; i < loopmax; i++) {
      const double b = rho[i] * vdot(y[i], z);
      z = vadd(z, s[i], a[i] - b);
    }
    Vec d = vscale(z, -1.0);
    const double distance = max_atom_motion_v(d);
    if (distance >= maxMove && prm.lbfgs_distance_reset) {
      reset();
      return max_atom_motion_applied_v(vscale(f, H0), maxMove);
    }
    double vd = vdot(safe_normalized(d), safe_normalized(f));
    vd = std::min(1.0, std::max(-1.0, vd));
    const double angle = safe_acos(vd) * (180.0 / kPi);
    if (angle > 90.0 && prm.lbfgs_angle_reset) {
      reset();
      return max_atom_motion_applied_v(vscale(f, H0), maxMove);
    }
    return max_atom_motion_applied_v(d, maxMove);
  }
  ObjectiveFunction *objf;
  OrcSearchParams prm;
  long memory;
  long iteration = 0;
  std::deque<Vec> s, y;
  std::deque<double> rho;
  Vec rPrev, fPrev;
};

// FIRE.cpp:19-68, FIRE.h:22-35
class FIRE : public Optimizer {
public:
  FIRE(ObjectiveFunction *o, const OrcSearchParams &p)
      : objf(o), dt(p.time_step), dt_max(p.max_time_step), max_move(p.max_move), vel(o->degreesOfFreedom(), 0.0) {}
  int step(double /*maxMove*/) override {
    if (objf->isConverged()) return 1;
    Vec f = vscale(objf->getGradient(), -1.0);
    Vec x = objf->getPositions();
    vel = vadd(vel, f, dt);
    Vec dx = vscale(vel, dt);
    dx = max_atom_motion_applied_v(dx, max_move);
    objf->setPositions(vadd(x, dx));
    f = vscale(objf->getGradient(), -1.0);
    const Vec f_unit = vscale(f, 1.0 / vnorm(f));
    const double P = vdot(f, vel);
    const double vn = vnorm(vel);
    for (size_t k = 0; k < vel.size(); k++) vel[k] = (1 - alpha) * vel[k] + alpha * f_unit[k] * vn;
    if (P >= 0) {
      N++;
      if (N > N_min) {
        dt = std::min(dt * f_inc, dt_max);
        alpha = alpha * f_a;
      }
    } else {
      dt = dt * f_dec;
      std::fill(vel.begin(), vel.end(), 0.0);
      alpha = alpha_start;
      N = 0;
    }
    if (dt < 1e-6) throw std::runtime_error("[FIRE] m_dt is too small");
    return objf->isConverged() ? 1 : 0;
  }

private:
  ObjectiveFunction *objf;
  double dt, dt_max, max_move;
  size_t N_min = 5, N = 0;
  Vec vel;
  double alpha_start = 0.1, alpha = 0.1, f_inc = 1.1, f_dec = 0.5, f_a = 0.99;
};

std::unique_ptr<Optimizer> make_optimizer(ObjectiveFunction *o, const OrcSearchParams &p) {
  switch (p.opt_method) {
    case ORC_OPT_LBFGS:
      return std::make_unique<LBFGS>(o, p);
    case ORC_OPT_FIRE:
      return std::make_unique<FIRE>(o, p);
    default:
      return std::make_unique<CG>(o, p);
  }
}

// ---------------------------------------------------------------------------------------------
struct LowestEigenmode {
  virtual ~LowestEigenmode() = default;
  virtual void compute(Matter &matter, const Vec &direction) = 0;
  virtual double getEigenvalue() = 0;
  virtual Vec getEigenvector() = 0;
  long statsRotations = 0;
};

// ImprovedDimer.cpp:63-434 (classical rotation backend, remove_rotation = false, OCINEB mode-loss
// check off: angle_tol/use_mmf are NEB-only options and default to off).
class ImprovedDimer : public LowestEigenmode {
public:
  ImprovedDimer(const Matter &m, const OrcSearchParams &p) : prm(p), x0(m), x1(m) {
    tau.assign(3 * m.nAtoms(), 0.0);
    if (prm.dimer_opt_method == ORC_ROT_CG) init_cg = true;
  }
  void compute(Matter &matter, const Vec &initialDirection) override {
    const Vec &mask = matter.topo()->free_mask;
    tau = initialDirection;
    for (size_t k = 0; k < tau.size(); k++) tau[k] *= mask[k];
    if (vnorm(tau) > 1e-10) {
      safe_normalize_inplace(tau);
    } else {
      // reference draws a random tangent here; deterministic stand-in for the oracle
      for (size_t k = 0; k < tau.size(); k++) tau[k] = mask[k] * ((k * 2654435761u % 1000) / 1000.0 - 0.5);
      safe_normalize_inplace(tau);
    }
    x0 = matter;
    x1 = matter;
    Vec x0_r = x0.getPositionsV();
    const double delta = prm.finite_difference;
    x1.setPositionsV(vadd(x0_r, tau, delta));
    if (x1.getPotentialEnergy() - x0.getPotentialEnergy() > 10.0 * delta) {
      tau = vscale(tau, -1.0);
      x1.setPositionsV(vadd(x0_r, tau, delta));
    }
    if (prm.dimer_opt_method == ORC_ROT_LBFGS) {
      ls.clear();
      ly.clear();
      lrho.clear();
      init_lbfgs = true;
    }
    const double phi_tol = kPi * (prm.dimer_converged_angle / 180.0);
    double phi_prime = 0.0, phi_min = 0.0;
    statsRotations = 0;
    Matter x1p(x1);
    Vec g0 = vscale(x0.getForcesV(), -1.0);
    Vec g1 = vscale(x1.getForcesV(), -1.0);
    Vec tau_Old;
    do {
      const Vec dg = vsub(g1, g0);
      // F_R = -2 (g1-g0) + 2 ((g1-g0).tau) tau
      F_R = vadd(vscale(dg, -2.0), tau, 2.0 * vdot(dg, tau));
      if (prm.dimer_opt_method == ORC_ROT_SD) {
        theta = safe_normalized(F_R);
      } else if (prm.dimer_opt_method == ORC_ROT_CG) {
        double gamma;
        if (init_cg) {
          init_cg = false;
          gamma = 0.0;
        } else {
          const double a = std::fabs(vdot(F_R, F_R_Old));
          const double b = vdot(F_R_Old, F_R_Old);
          gamma = (a < 0.5 * b) ? safe_div(vdot(F_R, vsub(F_R, F_R_Old)), b, 0.0) : 0.0;
        }
        theta = (gamma == 0.0) ? F_R : vadd(F_R, thetaOld, gamma);
        theta = vadd(theta, tau, -vdot(theta, tau));
        thetaOld = theta;
        if (vnorm(theta) < kEps) theta = vadd(F_R, tau, -vdot(F_R, tau));
        safe_normalize_inplace(theta);
        F_R_Old = F_R;
      } else {
        if (!init_lbfgs) {
          Vec s0 = vsub(tau, tau_Old);
          Vec y0 = vscale(vsub(F_R_Old, F_R), safe_div(1.0, delta, 0.0));
          lrho.push_back(safe_recip(vdot(s0, y0), 0.0));
          ls.push_back(std::move(s0));
          ly.push_back(std::move(y0));
        } else {
          init_lbfgs = false;
        }
        const double H0 = 1.0 / 60.0;
        const long loopmax = static_cast<long>(ls.size());
        std::vector<double> alpha(loopmax);
        Vec q = vscale(F_R, -1.0);
        for (long i = loopmax - 1; i >= 0; i--) {
          alpha[i] = lrho[i] * vdot(ls[i], q);
          q = vadd(q, ly[i], -alpha[i]);
        }
        Vec z = vscale(q, H0);
        for (long i = 0; i < loopmax; i++) {
          const double bv = lrho[i] * vdot(ly[i], z);
          z = vadd(z, ls[i], alpha[i] - bv);
        }
        const double vd = std::min(1.0, std::max(-1.0, -vdot(safe_normalized(z), safe_normalized(F_R))));
        const double angle = safe_acos(vd) * (180.0 / kPi);
        if (angle > 87.0) {
          ls.clear();
          ly.clear();
          lrho.clear();
          z = vscale(F_R, -1.0);
        }
        theta = vscale(safe_normalized(z), -1.0);
        theta = vadd(theta, tau, -vdot(theta, tau));
        safe_normalize_inplace(theta);
        thetaOld = theta;
        F_R_Old = F_R;
        tau_Old = tau;
      }
      C_tau = safe_div(vdot(dg, tau), delta, 0.0);
      const double dC = 2.0 * safe_div(vdot(dg, theta), delta, 0.0);
      phi_prime = -0.5 * safe_atan_ratio(dC, 2.0 * std::fabs(C_tau), 0.0);
      if (std::fabs(phi_prime) > phi_tol) {
        const double b1 = 0.5 * dC;
        x0_r = x0.getPositionsV();
        Vec tau_prime = vadd(vscale(tau, std::cos(phi_prime)), theta, std::sin(phi_prime));
        tau_prime = safe_normalized(tau_prime);
        const Vec x1_rp = vadd(x0_r, tau_prime, delta);
        x1p = x1;
        x1p.setPositionsV(x1_rp);
        const Vec g1_prime = vscale(x1p.getForcesV(), -1.0);
        const double C_tau_prime = safe_div(vdot(vsub(g1_prime, g0), tau_prime), delta, 0.0);
        const double a1 =
            safe_div(C_tau - C_tau_prime + b1 * std::sin(2.0 * phi_prime), 1.0 - std::cos(2.0 * phi_prime), 0.0);
        const double a0 = 2.0 * (C_tau - a1);
        phi_min = 0.5 * safe_atan_ratio(b1, a1, 0.0);
        double C_tau_min = 0.5 * a0 + a1 * std::cos(2.0 * phi_min) + b1 * std::sin(2.0 * phi_min);
        if (C_tau_min > C_tau) {
          phi_min += kPi * 0.5;
          C_tau_min = 0.5 * a0 + a1 * std::cos(2.0 * phi_min) + b1 * std::sin(2.0 * phi_min);
        }
        if (phi_min > kPi * 0.5) phi_min -= kPi;
        tau = vadd(vscale(tau, std::cos(phi_min)), theta, std::sin(phi_min));
        tau = safe_normalized(tau);
        x1.setPositionsV(vadd(x0_r, tau, delta));
        C_tau = C_tau_min;
        const double sin_pp = std::sin(phi_prime);
        const double c1 = safe_div(std::sin(phi_prime - phi_min), sin_pp, 0.0);
        const double c2 = safe_div(std::sin(phi_min), sin_pp, 0.0);
        const double c3 = 1.0 - std::cos(phi_min) - std::sin(phi_min) * std::tan(phi_prime * 0.5);
        for (size_t k = 0; k < g1.size(); k++) g1[k] = g1[k] * c1 + g1_prime[k] * c2 + g0[k] * c3;
        statsRotations += 1;
      }
    } while (std::fabs(phi_prime) > std::fabs(phi_tol) && std::fabs(phi_min) > std::fabs(phi_tol) &&
             statsRotations < prm.dimer_rotations_max);
  }
  double getEigenvalue() override { return C_tau; }
  Vec getEigenvector() override { return tau; }

private:
  OrcSearchParams prm;
  Matter x0, x1;
  Vec tau, theta, thetaOld, F_R, F_R_Old;
  double C_tau = 0.0;
  bool init_cg = false, init_lbfgs = false;
  std::vector<Vec> ls, ly;
  std::vector<double> lrho;
};

}  // namespace

// Cyclic Jacobi for a small symmetric matrix (stands in for Eigen::SelfAdjointEigenSolver,
// Lanczos.cpp:117-119).  Eigenvalues ascending; eigenvector k in vecs[k*n .. k*n+n).
extern "C" void orc_symmetric_eig(int n, const double *a_in, double *evals, double *vecs) {
  std::vector<double> a(a_in, a_in + n * n), v(n * n, 0.0);
  for (int i = 0; i < n; i++) v[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += a[p * n + q] * a[p * n + q];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        if (std::fabs(a[p * n + q]) < 1e-300) continue;
        const double th = (a[q * n + q] - a[p * n + p]) / (2.0 * a[p * n + q]);
        const double tt = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
        const double c = 1.0 / std::sqrt(tt * tt + 1.0), s = tt * c;
        for (int k = 0; k < n; k++) {
          const double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = v[k * n + p], vkq = v[k * n + q];
          v[k * n + p] = c * vkp - s * vkq;
          v[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return a[x * n + x] < a[y * n + y]; });
  for (int k = 0; k < n; k++) {
    evals[k] = a[order[k] * n + order[k]];
    for (int i = 0; i < n; i++) vecs[k * n + i] = v[i * n + order[k]];
  }
}

namespace {

// Lanczos.cpp:46-173 with phva_atoms = All (mobile set = every atom that is not fully fixed).
class Lanczos : public LowestEigenmode {
public:
  Lanczos(const Matter &m, const OrcSearchParams &p) : prm(p) { lowestEv.assign(3 * m.nAtoms(), 0.0); }
  void compute(Matter &matter, const Vec &direction) override {
    statsRotations = 0;
    std::fill(lowestEv.begin(), lowestEv.end(), 0.0);
    const int size = static_cast<int>(matter.degreesFree());
    if (size == 0) {
      lowestEw = 0.0;
      return;
    }
    const long maxIters = prm.lanczos_max_iterations;
    std::vector<double> T(maxIters * maxIters, 0.0);
    std::vector<Vec> Q;
    Vec r = matter.pack_free(direction), u;
    double alpha, beta = vnorm(r);
    if (beta < kEps) {
      lowestEw = 0.0;
      return;
    }
    double ew = 0, ewOld = 0, ewAbsRelErr;
    const double dr = prm.finite_difference;
    Vec evEst, evOldEst;
    Matter tmp(matter);
    const Vec pos0free = matter.getPositionsFreeV();
    const Vec pos0 = matter.getPositionsV();
    const Vec force0 = tmp.getForcesFreeV();
    auto applyH = [&](const Vec &v) {
      // AtomMatrix pos = pos0; unpack(pos0_mobile + dr v); setPositions(pos)
      Vec pos = pos0;
      const auto &fa = matter.topo()->free_atoms;
      for (size_t a = 0; a < fa.size(); a++)
        for (int k = 0; k < 3; k++) pos[3 * fa[a] + k] = pos0free[3 * a + k] + dr * v[3 * a + k];
      tmp.setPositionsV(pos);
      Vec f = tmp.getForcesFreeV();
      for (size_t k = 0; k < f.size(); k++) f[k] = -(f[k] - force0[k]) / dr;
      return f;
    };
    for (int i = 0; i < size; i++) {
      statsRotations = i;
      // NOTE the reference allocates Q(size, maxIters); i never reaches maxIters because of the
      // max-iterations break below.
      Q.push_back(vscale(r, 1.0 / beta));
      u = applyH(Q[i]);
      if (i == 0)
        r = u;
      else
        r = vadd(u, Q[i - 1], -beta);
      alpha = vdot(Q[i], r);
      r = vadd(r, Q[i], -alpha);
      T[i * maxIters + i] = alpha;
      if (i > 0) {
        T[(i - 1) * maxIters + i] = beta;
        T[i * maxIters + (i - 1)] = beta;
      }
      beta = vnorm(r);
      if (beta <= 1e-10 * std::fabs(alpha)) {
        if (i == 0) {
          ew = alpha;
          evEst = Q[0];
        }
        break;
      }
      if (i >= 1) {
        const int m = i + 1;
        std::vector<double> blk(m * m), ev(m), evec(m * m);
        for (int a = 0; a < m; a++)
          for (int b = 0; b < m; b++) blk[a * m + b] = T[a * maxIters + b];
        orc_symmetric_eig(m, blk.data(), ev.data(), evec.data());
        ew = ev[0];
        ewAbsRelErr = safe_div(std::fabs(ew - ewOld), std::fabs(ewOld), 1.0);
        ewOld = ew;
        evEst.assign(size, 0.0);
        for (int a = 0; a < m; a++)
          for (int k = 0; k < size; k++) evEst[k] += Q[a][k] * evec[a];
        const double nn = vnorm(evEst);
        for (double &x : evEst) x /= nn;
        evOldEst = evEst;
        if (ewAbsRelErr < prm.lanczos_tolerance) break;
      } else {
        ew = alpha;
        ewOld = ew;
        evEst = Q[0];
        evOldEst = Q[0];
        if (lowestEw != 0.0 && prm.lanczos_quit_early) {
          const double Cprev = lowestEw;
          const double Cnew = vdot(u, Q[i]);
          ewAbsRelErr = safe_div(std::fabs(Cnew - Cprev), std::fabs(Cprev), 1.0);
          if (ewAbsRelErr <= prm.lanczos_tolerance) break;
        }
      }
      if (i >= prm.lanczos_max_iterations - 1) break;
    }
    lowestEw = ew;
    std::fill(lowestEv.begin(), lowestEv.end(), 0.0);
    if (static_cast<int>(evEst.size()) == size) {
      const auto &fa = matter.topo()->free_atoms;
      for (size_t a = 0; a < fa.size(); a++)
        for (int k = 0; k < 3; k++) lowestEv[3 * fa[a] + k] = evEst[3 * a + k];
    }
  }
  double getEigenvalue() override { return lowestEw; }
  Vec getEigenvector() override { return lowestEv; }

private:
  OrcSearchParams prm;
  Vec lowestEv;
  double lowestEw = 0.0;
};

double convergence_of(Matter &m, int metric) {
  if (metric == ORC_METRIC_MAX_ATOM) return m.maxForce();
  if (metric == ORC_METRIC_MAX_COMPONENT) {
    const Vec f = m.getForcesV();
    return *std::max_element(f.begin(), f.end());  // AtomMatrix::maxCoeff (signed), as in the reference
  }
  return vnorm(m.getForcesFreeV());
}

// MinModeSaddleSearch.cpp:32-164
class MinModeObjective : public ObjectiveFunction {
public:
  MinModeObjective(Matter *m, LowestEigenmode *mm, const Vec &mode, const OrcSearchParams &p)
      : matter(m), minMode(mm), eigenvector(mode), prm(p) {}
  Vec getGradient(bool fdstep = false) override {
    Vec force = matter->getForcesV();
    if (!fdstep || iteration == 0) {
      minMode->compute(*matter, eigenvector);
      iteration++;
    }
    eigenvector = minMode->getEigenvector();
    const double eigenvalue = minMode->getEigenvalue();
    const Vec proj = vscale(safe_normalized(eigenvector), vdot(force, eigenvector));
    if (eigenvalue > 0.0) {
      if (prm.saddle_perp_force_ratio > 0.0) {
        const double d = prm.saddle_perp_force_ratio;
        for (size_t k = 0; k < force.size(); k++) force[k] = d * force[k] - (1.0 + d) * proj[k];
      } else {
        force = vscale(proj, -1.0);
      }
    } else {
      force = vadd(force, proj, -2.0);
    }
    return vscale(force, -1.0);
  }
  double getEnergy() override { return matter->getPotentialEnergy(); }
  void setPositions(const Vec &x) override { matter->setPositionsV(x); }
  Vec getPositions() override { return matter->getPositionsV(); }
  int degreesOfFreedom() override { return 3 * static_cast<int>(matter->nAtoms()); }
  bool isConverged() override {
    const double tol = prm.saddle_converged_force > 0.0 ? prm.saddle_converged_force : prm.opt_converged_force;
    return getConvergence() < tol;
  }
  double getConvergence() override { return convergence_of(*matter, prm.convergence_metric); }
  Vec difference(const Vec &a, const Vec &b) override { return matter->pbcV(vsub(a, b)); }

private:
  Matter *matter;
  LowestEigenmode *minMode;
  Vec eigenvector;
  OrcSearchParams prm;
  int iteration = 0;
};

// HelperFunctions.cpp:273-311
class MatterObjective : public ObjectiveFunction {
public:
  MatterObjective(Matter *m, const OrcSearchParams &p) : matter(m), prm(p) {}
  double getEnergy() override { return matter->getPotentialEnergy(); }
  Vec getGradient(bool = false) override { return vscale(matter->getForcesFreeV(), -1.0); }
  void setPositions(const Vec &x) override { matter->setPositionsFreeV(x); }
  Vec getPositions() override { return matter->getPositionsFreeV(); }
  int degreesOfFreedom() override { return static_cast<int>(matter->degreesFree()); }
  bool isConverged() override { return getConvergence() < prm.opt_converged_force; }
  double getConvergence() override { return convergence_of(*matter, prm.convergence_metric); }
  Vec difference(const Vec &a, const Vec &b) override { return matter->pbcV(vsub(a, b)); }

private:
  Matter *matter;
  OrcSearchParams prm;
};

std::shared_ptr<Topology> make_topology(const OrcPotential *pot, int remove_net_force, long n, const int *Z,
                                        const int *fixed, const double *box, ForceCounter *counter) {
  auto t = std::make_shared<Topology>();
  t->n = n;
  t->Z = Z;
  t->free_mask.resize(3 * n);
  for (long i = 0; i < n; i++) {
    bool all_fixed = true;
    for (int k = 0; k < 3; k++) {
      const bool fx = fixed && fixed[3 * i + k] != 0;
      t->free_mask[3 * i + k] = fx ? 0.0 : 1.0;
      if (fx) t->any_fixed_axis = true;
      all_fixed = all_fixed && fx;
    }
    if (!all_fixed) t->free_atoms.push_back(i);
  }
  std::memcpy(t->cell, box, 9 * sizeof(double));
  t->cell_ok = invert3(box, t->inv);
  t->remove_net_force = remove_net_force != 0;
  const OrcPotential potcopy = *pot;
  t->pot = [potcopy](long nn, const double *R, const int *ZZ, double *F, double *U, const double *bx) {
    return orc_force(&potcopy, nn, R, ZZ, F, U, bx);
  };
  t->counter = counter;
  return t;
}

}  // namespace

extern "C" {

void orc_search_params_default(OrcSearchParams *p) {
  std::memset(p, 0, sizeof(*p));
  p->finite_difference = 0.01;
  p->remove_net_force = 1;
  p->opt_method = ORC_OPT_CG;
  p->convergence_metric = ORC_METRIC_NORM;
  p->opt_max_iterations = 1000;
  p->max_move = 0.2;
  p->opt_converged_force = 0.01;
  p->time_step = 1.0 / 10.1805055;
  p->max_time_step = 2.5 / 10.1805055;
  p->lbfgs_memory = 20;
  p->lbfgs_inverse_curvature = 0.01;
  p->lbfgs_auto_scale = 1;
  p->lbfgs_angle_reset = 1;
  p->lbfgs_distance_reset = 1;
  p->cg_no_overshooting = 0;
  p->cg_knock_out_max_move = 0;
  p->cg_max_iter_before_reset = 0;
  p->dimer_opt_method = ORC_ROT_CG;
  p->dimer_converged_angle = 5.0;
  p->dimer_rotations_max = 10;
  p->lanczos_tolerance = 0.01;
  p->lanczos_max_iterations = 20;
  p->lanczos_quit_early = 1;
  p->minmode_method = ORC_MINMODE_DIMER;
  p->saddle_max_iterations = 1000;
  p->saddle_converged_force = 0.0;
  p->saddle_max_energy = 20.0;
  p->saddle_perp_force_ratio = 0.0;
  p->saddle_zero_mode_abort_curvature = 0.0;
  p->saddle_nonlocal_count_abort = 0;
  p->saddle_nonlocal_distance_abort = 0.0;
}

int orc_matter_forces(const OrcPotential *pot, int remove_net_force, long n, const double *positions, const int *Z,
                      const int *fixed, const double *box, double *forces, double *energy) {
  try {
    ForceCounter counter;
    auto topo = make_topology(pot, remove_net_force, n, Z, fixed, box, &counter);
    Matter m(topo);
    m.setPositionsV(Vec(positions, positions + 3 * n));
    *energy = m.getPotentialEnergy();
    const Vec f = m.getForcesV();
    std::memcpy(forces, f.data(), 3 * n * sizeof(double));
    return 0;
  } catch (...) {
    return 11;
  }
}

int orc_saddle_search(const OrcPotential *pot, const OrcSearchParams *prm_in, long n, double *positions,
                      const int *Z, const int *fixed, const double *box, double *mode_io, double reactant_energy,
                      OrcSaddleResult *out) {
  const OrcSearchParams prm = *prm_in;
  ForceCounter counter;
  auto topo = make_topology(pot, prm.remove_net_force, n, Z, fixed, box, &counter);
  Matter matter(topo);
  matter.setPositionsV(Vec(positions, positions + 3 * n));
  Vec mode(mode_io, mode_io + 3 * n);
  int status = 0;  // STATUS_GOOD
  long iteration = 0;
  std::unique_ptr<LowestEigenmode> minMode;
  try {
    if (prm.minmode_method == ORC_MINMODE_LANCZOS)
      minMode = std::make_unique<Lanczos>(matter, prm);
    else
      minMode = std::make_unique<ImprovedDimer>(matter, prm);
    MinModeObjective objf(&matter, minMode.get(), mode, prm);
    const Vec initialPosition = matter.getPositionsV();
    auto optim = make_optimizer(&objf, prm);
    bool firstIteration = true;
    while (!objf.isConverged() || iteration == 0) {
      if (!firstIteration) {
        if (prm.saddle_nonlocal_count_abort != 0) {
          const Vec d = vsub(initialPosition, matter.getPositionsV());  // no MIC, as in the reference
          long nm = 0;
          for (long i = 0; i < n; i++)
            if (std::sqrt(d[3 * i] * d[3 * i] + d[3 * i + 1] * d[3 * i + 1] + d[3 * i + 2] * d[3 * i + 2]) >=
                prm.saddle_nonlocal_distance_abort)
              nm++;
          if (nm >= prm.saddle_nonlocal_count_abort) {
            status = 13;
            break;
          }
        }
        if (std::fabs(minMode->getEigenvalue()) < prm.saddle_zero_mode_abort_curvature) {
          status = 18;
          break;
        }
      }
      firstIteration = false;
      if (iteration >= prm.saddle_max_iterations) {
        status = 5;
        break;
      }
      const int optStatus = optim->step(prm.max_move);
      if (optStatus < 0) {
        status = 19;
        break;
      }
      const double de = objf.getEnergy() - reactant_energy;
      iteration++;
      if (de > prm.saddle_max_energy) {
        status = 3;
        break;
      }
    }
    if (iteration == 0) minMode->compute(matter, mode);
    const bool climbConverged = objf.isConverged();
    if (status == 0 && !climbConverged) status = 5;  // finalizeClimbStatus, MinModeSaddleSearch.h:90-96
    if (minMode->getEigenvalue() > 0.0 && status == 0) status = 16;
    out->convergence = objf.getConvergence();
  } catch (const std::runtime_error &) {
    status = 11;  // STATUS_POTENTIAL_FAILED (SaddleSearchJob.cpp:114-123)
    out->convergence = std::numeric_limits<double>::quiet_NaN();
  }
  out->status = status;
  out->iterations = iteration;
  out->energy = status == 11 ? std::numeric_limits<double>::quiet_NaN() : matter.getPotentialEnergy();
  out->force_calls = counter.calls;
  out->eigenvalue = minMode ? minMode->getEigenvalue() : 0.0;
  const Vec &p = matter.getPositionsV();
  std::memcpy(positions, p.data(), 3 * n * sizeof(double));
  if (minMode) {
    const Vec ev = minMode->getEigenvector();
    std::memcpy(mode_io, ev.data(), 3 * n * sizeof(double));
  }
  return 0;
}

int orc_relax(const OrcPotential *pot, const OrcSearchParams *prm_in, long n, double *positions, const int *Z,
              const int *fixed, const double *box, OrcRelaxResult *out) {
  const OrcSearchParams prm = *prm_in;
  ForceCounter counter;
  auto topo = make_topology(pot, prm.remove_net_force, n, Z, fixed, box, &counter);
  Matter matter(topo);
  matter.setPositionsV(Vec(positions, positions + 3 * n));
  long iteration = 0;
  try {
    MatterObjective objf(&matter, prm);
    auto optim = make_optimizer(&objf, prm);
    while (!objf.isConverged() && iteration < prm.opt_max_iterations) {
      optim->step(prm.max_move);
      iteration++;
    }
    out->converged = objf.isConverged() ? 1 : 0;
    out->convergence = objf.getConvergence();
    out->energy = matter.getPotentialEnergy();
  } catch (const std::runtime_error &) {
    out->converged = 0;
    out->convergence = std::numeric_limits<double>::quiet_NaN();
    out->energy = std::numeric_limits<double>::quiet_NaN();
  }
  out->iterations = iteration;
  out->force_calls = counter.calls;
  const Vec &p = matter.getPositionsV();
  std::memcpy(positions, p.data(), 3 * n * sizeof(double));
  return 0;
}

}  // extern "C"
