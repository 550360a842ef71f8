/*
 * oracle/oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain C++17 restatement (C ABI, no Eigen) of the reference algorithms on the
 * batched saddle-search hot path of TheochemUI/eOn.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or the timed CPU baseline -- the
 * product (libeonb200.so) never links, loads or calls anything in oracle/.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - LJ / LJCluster / Morse_Pt: arithmetic lives in the un-vendored third-party
 *     rgpot (subprojects/rgpot.wrap:1-15, branch feat/force-batch-impl, no tag);
 *     restated from the published functional forms and PINNED against the
 *     reference's own goldens (client/unit_tests/PotTest.cpp:38-43,130-132;
 *     client/unit_tests/data/reference/point_morse_pt.dat).
 *   - EAM: restates client/potentials/EAM/EAM.cpp:28-114,125-226,349-365 and is
 *     PINNED bit-for-bit against the reference's own EAM.cpp compiled from where
 *     it lies (oracle/_ref/libeam_ref.so, recipe in oracle/Makefile).  The
 *     reference holds no golden for that class, so beyond "equals the reference
 *     binary" its parity is "unpinned by reference tests".
 *   - Neighbour sets: brute-force restatement of vesin's final predicate
 *     (client/thirdparty/vesin/vesin-single-build.cpp:1165-1202), PINNED against
 *     vesin itself compiled into oracle/_ref/libvesin_ref.so.
 *   - Dimer / Lanczos / CG / LBFGS / FIRE / MinModeSaddleSearch / relax: restated
 *     from the cited lines; PINNED against the known-answer job results in
 *     client/unit_tests/data/reference/{saddle_search_morse_pt,
 *     saddle_search_lanczos_morse_pt,process_search_morse_pt,minimization_lj,
 *     min_lj_cg,min_lj_fire}.dat.
 */
#ifndef EONB200_ORACLE_H
#define EONB200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_POT_LJ = 0, ORC_POT_LJCLUSTER = 1, ORC_POT_MORSE_PT = 2, ORC_POT_EAM = 3, ORC_POT_EAM_NL = 4 };

typedef struct OrcEamRow {
  int Z;
  double Dm, alphaM, Rm, beta1, beta2, r_cut;
  double func_coeff[9];
} OrcEamRow;

/* Potential description (same fields as the product's EonB200Config). */
typedef struct OrcPotential {
  int kind;
  double cutoff;          /* LJ / Morse cutoff */
  double lj_u0, lj_psi;
  double morse_De, morse_a, morse_re;
  int n_eam_rows;         /* 0 -> built-in Al/Pd/Au rows of EAM/Parameters.h:15-50 */
  const OrcEamRow *eam_rows;
  int eam_conservative;   /* 0: pair-force sign exactly as EAM.cpp:203-215 (NOT -grad U);
                             1: F = -grad U (needed for searches to converge) */
} OrcPotential;

void orc_potential_default(OrcPotential *pot, int kind);

/* One energy+force evaluation with the reference's Potential::force signature
 * (include/eon/Potential.h:55-57).  Returns 0, or 14324 for an unknown element
 * (EAM.cpp:116-123). */
int orc_force(const OrcPotential *pot, long nAtoms, const double *positions, const int *atomicNrs,
              double *forces, double *energy, const double *box);

/* Count of ordered in-cutoff pairs the potential visits (for the roofline's algorithmic work). */
long orc_count_ordered_pairs(const OrcPotential *pot, long nAtoms, const double *positions,
                             const int *atomicNrs, const double *box);

/* Brute-force neighbour list with vesin's predicate and half-list rules.
 * pairs: [cap][2] int64, shifts: [cap][3] int32.  Returns the number of pairs (may exceed
 * cap, in which case only cap were written). */
long orc_neighbors(long n, const double *points, const double *box, const int periodic[3], double cutoff,
                   int full, long cap, int64_t *pairs, int32_t *shifts);

/* ---- search algorithms ------------------------------------------------------------------ */

enum { ORC_OPT_CG = 0, ORC_OPT_LBFGS = 1, ORC_OPT_FIRE = 2 };
enum { ORC_MINMODE_DIMER = 0, ORC_MINMODE_LANCZOS = 1 };
enum { ORC_ROT_SD = 0, ORC_ROT_CG = 1, ORC_ROT_LBFGS = 2 };
enum { ORC_METRIC_NORM = 0, ORC_METRIC_MAX_ATOM = 1, ORC_METRIC_MAX_COMPONENT = 2 };

/* Subset of eonc::Parameters the path reads (include/eon/Parameters.h:55,58,176-197,
 * 220-277,418-423; links client/ParametersINI.cpp:1075-1077,1335-1377). */
typedef struct OrcSearchParams {
  double finite_difference;      /* [Main] finite_difference, 0.01 */
  int remove_net_force;          /* [Main] remove_net_force, true */
  int opt_method;                /* [Optimizer] opt_method, cg */
  int convergence_metric;        /* norm */
  long opt_max_iterations;       /* 1000 */
  double max_move;               /* 0.2 */
  double opt_converged_force;    /* 0.01 */
  double time_step;              /* FIRE, internal units: 1.0 fs / 10.1805055 */
  double max_time_step;          /* 2.5 fs / 10.1805055 */
  long lbfgs_memory;             /* 20 */
  double lbfgs_inverse_curvature;/* 0.01 */
  int lbfgs_auto_scale, lbfgs_angle_reset, lbfgs_distance_reset; /* true */
  int cg_no_overshooting, cg_knock_out_max_move; /* false */
  long cg_max_iter_before_reset; /* 0 */
  int dimer_opt_method;          /* [Dimer] opt_method, cg */
  double dimer_converged_angle;  /* 5.0 degrees */
  long dimer_rotations_max;      /* 10 */
  double lanczos_tolerance;      /* 0.01 */
  long lanczos_max_iterations;   /* 20 */
  int lanczos_quit_early;        /* true */
  int minmode_method;            /* dimer */
  long saddle_max_iterations;    /* 1000 */
  double saddle_converged_force; /* <=0 -> opt_converged_force (ParametersINI.cpp:1075-1077) */
  double saddle_max_energy;      /* 20.0 */
  double saddle_perp_force_ratio;/* 0.0 */
  double saddle_zero_mode_abort_curvature; /* 0.0 */
  long saddle_nonlocal_count_abort;        /* 0 */
  double saddle_nonlocal_distance_abort;   /* 0.0 */
} OrcSearchParams;

void orc_search_params_default(OrcSearchParams *p);

typedef struct OrcSaddleResult {
  int status;              /* MinModeSaddleSearch::Status, include/eon/MinModeSaddleSearch.h:29-53 */
  long iterations;
  long force_calls;        /* force calls made by the climb (excludes the reactant energy call) */
  double energy;           /* potential energy at the final (saddle) geometry */
  double eigenvalue;
  double convergence;      /* final value of the convergence metric */
} OrcSaddleResult;

/* MinModeSaddleSearch::run on one walker (client/MinModeSaddleSearch.cpp:211-475).
 * positions (in/out): displaced start -> final geometry; mode (in/out): initial direction -> final
 * eigenvector; fixed: [N][3] 0/1 mask; reactant_energy used for the max_energy check. */
int orc_saddle_search(const OrcPotential *pot, const OrcSearchParams *prm, long nAtoms, double *positions,
                      const int *atomicNrs, const int *fixed, const double *box, double *mode,
                      double reactant_energy, OrcSaddleResult *out);

typedef struct OrcRelaxResult {
  int converged;
  long iterations;
  long force_calls;
  double energy;
  double convergence;
} OrcRelaxResult;

/* helpers::relaxMatter (client/HelperFunctions.cpp:314-410) on the free DOF. */
int orc_relax(const OrcPotential *pot, const OrcSearchParams *prm, long nAtoms, double *positions,
              const int *atomicNrs, const int *fixed, const double *box, OrcRelaxResult *out);

/* Matter-level single evaluation: potential + fixed mask + net-force removal
 * (client/Matter.cpp:324-333,498-538).  forces are the masked forces Matter::getForces returns. */
int orc_matter_forces(const OrcPotential *pot, int remove_net_force, long nAtoms, const double *positions,
                      const int *atomicNrs, const int *fixed, const double *box, double *forces, double *energy);

/* Eigenvalues (ascending) and eigenvectors (column k at vecs[k*n..]) of a small symmetric matrix,
 * cyclic Jacobi.  Exposed so the device tridiagonal solver can be checked against it. */
void orc_symmetric_eig(int n, const double *a, double *evals, double *vecs);

#ifdef __cplusplus
}
#endif
#endif
