/*
 * eonb200.h -- C ABI of libeonb200.so, the B200 (sm_100a) backend for eOn's batched
 * saddle-search hot path.
 *
 * Plain C: pointers and sizes only, no C++/torch types.  Every entry point names the
 * reference interface (paths relative to the eOn checkout) it stands in for.  The pattern
 * (opaque handle, create/destroy/force/abi_version, int return code + char* errbuf) is the
 * reference's own dlopen precedent:
 *     include/eon/potentials/Metatomic/metatomic_c_abi.h:30-79
 *     client/potentials/Metatomic/MetatomicDynPot.cpp:39-70
 *
 * All floating point is FP64.  Positions are row-major [x0,y0,z0,x1,...] in Angstrom, boxes
 * are row-major 3x3 with the cell vectors as rows (include/eon/Eigen.h:28-37,
 * client/Matter.cpp:523-524).  The library needs a CUDA device; there is no CPU fallback --
 * eonb200_create fails with a message when no sm_100 device is usable.
 */
#ifndef EONB200_H
#define EONB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EONB200_ABI_VERSION 6

#if defined(_WIN32)
#define EONB200_API __declspec(dllexport)
#else
#define EONB200_API __attribute__((visibility("default")))
#endif

/* ---- potential description ------------------------------------------------------------- */

/* PotType arms this backend re-points (include/eon/BaseStructures.h:36-71,
 * client/Potential.cpp:140-153): LJ, LJCLUSTER, MORSE_PT (rgpot kernels) and the in-tree EAM
 * class (client/potentials/EAM/EAM.cpp). */
enum {
  EONB200_POT_LJ = 0,        /* 4 u0 [(psi/r)^12 - (psi/r)^6] - U(rc), r < rc        */
  EONB200_POT_LJCLUSTER = 1, /* same, no energy shift                                */
  EONB200_POT_MORSE_PT = 2,  /* De (1 - exp(-a (r - re)))^2 - De - U(rc), r < rc     */
  EONB200_POT_EAM = 3        /* EAM.cpp:125-226: r^6 (e^-b1 r + 512 e^-2 b2 r) density,
                                degree-8 embedding polynomial, Morse pair part        */
};

/* One element row of include/eon/potentials/EAM/Parameters.h:15-50. */
typedef struct EonB200EamRow {
  int Z;
  double Dm, alphaM, Rm, beta1, beta2, r_cut;
  double func_coeff[9];
} EonB200EamRow;

typedef struct EonB200Config {
  int kind;               /* EONB200_POT_*                                                     */
  int device;             /* CUDA device ordinal                                               */
  double cutoff;          /* LJ / Morse cutoff; <= 0 -> 15.0 (LJ) / 9.5 (Morse_Pt)              */
  double lj_u0, lj_psi;   /* <= 0 -> 1.0                                                       */
  double morse_De, morse_a, morse_re; /* all 0 -> Pt values 0.7102 / 1.6047 / 2.8970           */
  int n_eam_rows;         /* 0 -> the reference's built-in Al / Pd / Au rows                   */
  const EonB200EamRow *eam_rows; /* caller-supplied table (e.g. a Cu row), copied at create    */
  int eam_conservative;   /* 0: pair-force sign exactly as EAM.cpp:203-215 (bit-faithful to the
                             reference class, which is NOT -grad U); 1: F = -grad U            */
  double skin;            /* Verlet skin of the cached neighbour lists, Angstrom; < 0 -> 1.0;
                             0 -> exact list rebuilt on every evaluation                       */
  int threads_per_walker; /* CTA size override of the resident walker kernel, 0 -> from nAtoms  */
  int tiles_override;     /* CTAs per system: cluster size of the fused EAM kernel (1..8), tiles of the LJ / Morse
                             force pass (1..16); 0 -> chosen from nSystems and nAtoms                    */
  int lpa_override;       /* lanes sharing one atom's neighbour row (1, 2, 4, 8), 0 -> automatic */
  int pin_host_buffers;   /* 1: when the systems of a force_batch call lie back to back in host memory (positions[s]
                             = positions[0] + s * 3N, same for forces), page-lock those ranges once
                             (cudaHostRegister, cached per handle, released at destroy) and copy straight from / to
                             them instead of staging through the handle's own pinned buffers.  The caller must keep
                             such buffers allocated until eonb200_destroy.  0 (default): always stage.          */
  int eam_pair_mode;      /* EAM force_batch kernel.  0 (default): every unordered pair evaluated once (Newton's third
                             law; both atoms accumulate in shared memory, so results agree run to run to rounding,
                             ~1e-16 relative).  1: every ordered pair evaluated from its own atom's row, fixed summation
                             order (bit-reproducible; about twice the arithmetic).  Systems with several elements or a
                             box narrower than 2 (r_cut + skin) + skin always take the ordered-pair general code.        */
} EonB200Config;

typedef struct EonB200Pot EonB200Pot; /* opaque */

/* Fill cfg with the defaults of the given potential kind. */
EONB200_API void eonb200_config_default(EonB200Config *cfg, int kind);

/* replaces helpers::makePotential for these PotTypes (client/Potential.cpp:124-303).
 * NULL on failure with a message in errbuf. */
EONB200_API EonB200Pot *eonb200_create(const EonB200Config *cfg, char *errbuf, size_t errlen);
EONB200_API void eonb200_destroy(EonB200Pot *pot);
EONB200_API int eonb200_abi_version(void);
/* message of the last failing call on this handle ("" if none) */
EONB200_API const char *eonb200_last_error(const EonB200Pot *pot);

/* Return codes: 0 ok; 14324 unknown element (EAM.cpp:116-123 `throw 14324`); 100 potential
 * failed (the reference's legacy `throw 100`, SaddleSearchJob.cpp:116-123); other non-zero =
 * CUDA / argument error (see eonb200_last_error). */

/* replaces Potential::force (include/eon/Potential.h:55-57).  Host pointers; forces is fully
 * overwritten; variance may be NULL (written 0 otherwise).  Ignores the fixed mask. */
EONB200_API int eonb200_force(EonB200Pot *pot, long nAtoms, const double *positions, const int *atomicNrs,
                              double *forces, double *energy, double *variance, const double *box);

/* replaces Potential::forceBatch (include/eon/Potential.h:98-122) as consumed by
 * client/ImprovedDimer.cpp:160-193, client/Dimer.cpp:170-210, client/NudgedElasticBand.cpp:475-521.
 * Arrays of nSystems host pointers; every system has nAtoms atoms.  One H2D copy, one batched
 * launch sequence, one D2H copy. */
EONB200_API int eonb200_force_batch(EonB200Pot *pot, long nSystems, long nAtoms, const double *const *positions,
                                    const int *const *atomicNrs, double *const *forces, double *energies,
                                    double *variances, const double *const *boxes);

/* Same evaluation with everything already resident in HBM (contiguous device arrays):
 * d_positions [S][3N], d_atomicNrs [S][N], d_boxes [S][9] in; d_forces [S][3N], d_energies [S] out.
 * Asynchronous on the handle's stream unless sync != 0.  atomicNrs are NOT validated here. */
EONB200_API int eonb200_force_batch_device(EonB200Pot *pot, long nSystems, long nAtoms, const double *d_positions,
                                           const int *d_atomicNrs, double *d_forces, double *d_energies,
                                           const double *d_boxes, int sync);

/* Force-call counter of the handle (Potential::forceCallCounter, include/eon/Potential.h:32):
 * one increment per system evaluated through force / force_batch / resident walkers. */
EONB200_API unsigned long long eonb200_force_calls(const EonB200Pot *pot);

/* ---- neighbour lists -------------------------------------------------------------------- */

/* replaces vesin_neighbors (client/thirdparty/vesin/vesin.h:190-199) as wrapped by
 * VesinNeighbors::compute (client/VesinNeighbors.cpp:42-89): periodic cell-list build on the
 * device for nSystems independent systems of nAtoms points.  The set of (i, j, shift) per system
 * equals vesin's bit for bit (same final predicate |(r_j - r_i) + S.H|^2 < cutoff^2 with
 * un-fused arithmetic, same half-list rule); order is by i, unspecified within i.
 *   full != 0: both (i,j,S) and (j,i,-S).
 * Output (host, caller-allocated): pair_offsets [nSystems+1]; pairs [cap][2] (int64), shifts
 * [cap][3] (int32); distances [cap] and vectors [cap][3] may be NULL.  Returns the total number
 * of pairs through *n_pairs; if it exceeds cap nothing is written beyond cap and the call
 * returns 0 with *n_pairs > cap (call again with a larger buffer). */
EONB200_API int eonb200_neighbors(EonB200Pot *pot, long nSystems, long nAtoms, const double *positions,
                                  const double *boxes, const int periodic[3], double cutoff, int full,
                                  long cap, long *pair_offsets, int64_t *pairs, int32_t *shifts,
                                  double *distances, double *vectors, long *n_pairs);

/* ---- resident walkers ------------------------------------------------------------------- */

enum { EONB200_OPT_CG = 0, EONB200_OPT_LBFGS = 1, EONB200_OPT_FIRE = 2 };
enum { EONB200_MINMODE_DIMER = 0, EONB200_MINMODE_LANCZOS = 1 };
enum { EONB200_ROT_SD = 0, EONB200_ROT_CG = 1, EONB200_ROT_LBFGS = 2 };
enum { EONB200_METRIC_NORM = 0, EONB200_METRIC_MAX_ATOM = 1, EONB200_METRIC_MAX_COMPONENT = 2 };

/* The subset of eonc::Parameters this path reads (include/eon/Parameters.h:55,58,176-197,
 * 220-277,418-423; links client/ParametersINI.cpp:1075-1077,1335-1377).  Field-for-field the
 * same as the oracle's OrcSearchParams. */
typedef struct EonB200SearchParams {
  double finite_difference;
  int remove_net_force;
  int opt_method;
  int convergence_metric;
  long opt_max_iterations;
  double max_move;
  double opt_converged_force;
  double time_step;
  double max_time_step;
  long lbfgs_memory;
  double lbfgs_inverse_curvature;
  int lbfgs_auto_scale, lbfgs_angle_reset, lbfgs_distance_reset;
  int cg_no_overshooting, cg_knock_out_max_move;
  long cg_max_iter_before_reset;
  int dimer_opt_method;
  double dimer_converged_angle;
  long dimer_rotations_max;
  double lanczos_tolerance;
  long lanczos_max_iterations;
  int lanczos_quit_early;
  int minmode_method;
  long saddle_max_iterations;
  double saddle_converged_force;
  double saddle_max_energy;
  double saddle_perp_force_ratio;
  double saddle_zero_mode_abort_curvature;
  long saddle_nonlocal_count_abort;
  double saddle_nonlocal_distance_abort;
} EonB200SearchParams;

EONB200_API void eonb200_search_params_default(EonB200SearchParams *p);

typedef struct EonB200Walkers EonB200Walkers; /* opaque: W walkers of one topology, resident in HBM */

/* Upload W walkers that share nAtoms, atomicNrs, the per-axis fixed mask [N][3] (0 free / 1 fixed,
 * include/eon/Matter.h:217-227) and the box.  positions [W][3N]: start geometries (the displaced
 * structures, SaddleSearchJob.cpp:27-61); modes [W][3N]: initial directions (direction.dat), may be
 * NULL for relaxation-only use.  Everything stays on the device until harvest. */
EONB200_API EonB200Walkers *eonb200_walkers_create(EonB200Pot *pot, long nWalkers, long nAtoms, const int *atomicNrs,
                                                   const int *fixed, const double *box, const double *positions,
                                                   const double *modes, char *errbuf, size_t errlen);
EONB200_API void eonb200_walkers_destroy(EonB200Walkers *w);

/* Replace the geometries / modes of all walkers (host arrays [W][3N]; either may be NULL). */
EONB200_API int eonb200_walkers_set(EonB200Walkers *w, const double *positions, const double *modes);

/* Matter::getPotentialEnergy + getForces on every walker (client/Matter.cpp:324-333,498-538):
 * potential, net-force removal when nothing is fixed, fixed-axis mask.  Host outputs, may be NULL. */
EONB200_API int eonb200_walkers_matter_forces(EonB200Walkers *w, int remove_net_force, double *energies,
                                              double *forces);

/* MinModeSaddleSearch::run for every walker, device resident, one CTA per walker, no host
 * round trips inside a search (client/MinModeSaddleSearch.cpp:211-475 with
 * client/ImprovedDimer.cpp:63-434 or client/Lanczos.cpp:46-173 and client/{ConjugateGradients,
 * LBFGS,FIRE}.cpp).  reactant_energies [W] (or NULL -> 0) feed the max_energy abort. */
EONB200_API int eonb200_walkers_saddle_search(EonB200Walkers *w, const EonB200SearchParams *prm,
                                              const double *reactant_energies);

/* helpers::relaxMatter on every walker (client/HelperFunctions.cpp:273-410): minimisation over
 * the free degrees of freedom with the optimizer named in prm. */
EONB200_API int eonb200_walkers_relax(EonB200Walkers *w, const EonB200SearchParams *prm);

/* ProcessSearchJob's displacement of the two minima from the saddle
 * (client/ProcessSearchJob.cpp:188-202): positions += sign * offset * mode on the free axes. */
EONB200_API int eonb200_walkers_displace_along_mode(EonB200Walkers *w, double signed_offset);

/* Host-side harvest (the only "collective" of the path): any output may be NULL.
 *   status [W]       MinModeSaddleSearch::Status (include/eon/MinModeSaddleSearch.h:29-53) or, after
 *                    relax, 0 converged / 1 not converged
 *   iterations [W], force_calls [W]
 *   energies [W], eigenvalues [W], convergence [W]
 *   positions [W][3N], modes [W][3N] */
EONB200_API int eonb200_walkers_harvest(EonB200Walkers *w, int *status, long *iterations, long *force_calls,
                                        double *energies, double *eigenvalues, double *convergence,
                                        double *positions, double *modes);

/* Start geometries and initial modes made on the device (eon/displace.py:180-263, the displacement the server makes
 * for every job it hands out): walker w is displaced around atom epicentres[w] -- that atom and every atom within
 * `radius` of it (minimum image) get a normal deviate of standard deviation `sigma` on each free axis -- starting from
 * reference_positions ([3N] shared when nReferences == 1, else [W][3N]); its mode is the normalised displacement.
 * Random numbers: Philox4x32-10, counter (atom, global walker id = first_walker_id + w), key `seed`; the reference
 * draws from numpy's generator, so displacements are reproducible from the seed here, not bit-identical to the server's. */
EONB200_API int eonb200_walkers_displace(EonB200Walkers *w, const double *reference_positions, long nReferences,
                                         const int *epicentres, double radius, double sigma, unsigned long long seed,
                                         long first_walker_id);

/* Which walkers take part in the following saddle_search / relax / displace launches: active [W], 0 = leave the walker
 * alone (its slot keeps its last results).  NULL: all.  ProcessSearchJob returns right after a failed saddle search
 * (client/ProcessSearchJob.cpp:184-186): the batched tail masks those walkers out of the two minimisations. */
EONB200_API int eonb200_walkers_set_active(EonB200Walkers *w, const int *active);

/* ---- streaming: keep every slot busy -------------------------------------------------------------------------------
 * The reference keeps its cores busy by handing out the next search as soon as one returns
 * (eon/communicator.py:525-559).  Here the W slots of a handle are worked in groups: every group has its own stream,
 * a job on a slice of slots is submitted without waiting, and a finished group is harvested and refilled while the
 * others keep running -- a walker that runs into max_iterations only holds up its own group.
 *   groups(n)        create n group streams (1..64) and size the neighbour rows (asynchronous jobs cannot regrow them)
 *   submit(g, job, prm, first, count, positions, modes, reactant_energies)
 *                    upload the inputs of slots [first, first + count) ([count][3N], [count]; positions / modes may be
 *                    NULL to keep what the slots hold) and start job 1 (saddle search) or 2 (relax) on them; every job
 *                    in flight must use the same prm
 *   poll(g, wait)    1 = done or idle, 0 = still running, -1 = error; wait != 0 blocks until done
 *   harvest_group(g, ...)  outputs of the group's slice, sized [count] / [count][3N] like harvest; frees the group */
EONB200_API int eonb200_walkers_groups(EonB200Walkers *w, int nGroups);
EONB200_API int eonb200_walkers_submit(EonB200Walkers *w, int group, int job, const EonB200SearchParams *prm, long first,
                                       long count, const double *positions, const double *modes,
                                       const double *reactant_energies);
EONB200_API int eonb200_walkers_poll(EonB200Walkers *w, int group, int wait);
EONB200_API int eonb200_walkers_harvest_group(EonB200Walkers *w, int group, int *status, long *iterations,
                                              long *force_calls, double *energies, double *eigenvalues,
                                              double *convergence, double *positions, double *modes);

/* A queue of nQueue >= 1 saddle searches (job 1) or relaxations (job 2) through the handle's W resident slots in ONE
 * launch: every CTA takes the next entry of the queue from a device counter the moment its walker is done, so no slot
 * waits for a straggler -- the device-side counterpart of the server handing the next job to a free core
 * (eon/communicator.py:525-559).  positions / modes: [nQueue][3N] start geometries and modes in, final ones out (modes
 * may be NULL for job 2); reactant_energies: [nQueue] or NULL; outputs [nQueue], any may be NULL.  Which slot runs an
 * entry does not change its result: an entry starts from a freshly loaded Matter, as SaddleSearchJob does. */
EONB200_API int eonb200_walkers_queue(EonB200Walkers *w, int job, const EonB200SearchParams *prm, long nQueue,
                                      double *positions, double *modes, const double *reactant_energies, int *status,
                                      long *iterations, long *force_calls, double *energies, double *eigenvalues,
                                      double *convergence);

/* Device time in milliseconds of the last walkers_* launch sequence (CUDA events on the
 * handle's stream), for the measurement harness. */
EONB200_API double eonb200_walkers_last_ms(const EonB200Walkers *w);

/* ---- measurement helpers ---------------------------------------------------------------- */

/* Register-resident DFMA microbenchmark: returns the measured non-tensor FP64 rate in FLOP/s
 * (2 flop per DFMA) on the handle's device; the roofline denominator for the pair kernels. */
EONB200_API double eonb200_measure_fp64_peak(EonB200Pot *pot, int repeats);

/* Kernel launch counter of the handle (every kernel this library launches bumps it). */
EONB200_API unsigned long long eonb200_kernel_launches(const EonB200Pot *pot);

/* Neighbour-list statistics of the last force_batch / walkers call on the handle:
 * out[0] = lists rebuilt, out[1] = lists reused, out[2] = row capacity K, out[3] = max row fill,
 * out[4] = ordered in-cutoff pairs visited by the last force evaluation (sum over systems). */
EONB200_API int eonb200_nl_stats(EonB200Pot *pot, long long out[5]);

/* ---- NEB band forces ----------------------------------------------------------------------
 * Replaces the per-image loop of NudgedElasticBand::updateForces after the potential call
 * (client/NudgedElasticBand.cpp:551-646) for many bands at once: minimum-image neighbour differences, tangent
 * (client/NEBTangent.cpp:36-69), spring force (client/NEBSpringForce.cpp:17-42,113-136), projection
 * (client/NEBProjection.cpp:17-45, client/NEBForceProjection.cpp:65-97), climbing image and zeroTranslation (:101-110).
 * The Onsager-Machlup spring (NEBSpringForce.cpp:46-111) is not covered. */
enum { EONB200_NEB_PROJECTION_NEB = 0, EONB200_NEB_PROJECTION_DNEB = 1, EONB200_NEB_PROJECTION_ELASTIC_BAND = 2 };

typedef struct EonB200NebParams {
  double spring_constant; /* neb_options.spring.constant (5.0) */
  double k_min, k_max;    /* neb_options.spring.weighting.k_min / k_max (0.97 / 9.7), used when energy_weighted */
  int climbing_image;     /* ci_active of the reference: the highest movable image climbs */
  int old_tangent;        /* neb_options.climbing_image.use_old_tangent: tangent = direction to the next image */
  int projection;         /* EONB200_NEB_PROJECTION_*: spring.doubly_nudged / spring.use_elastic_band */
  int energy_weighted;    /* neb_options.spring.weighting.enabled (forces the plain NEB projection, as the reference) */
} EonB200NebParams;

EONB200_API void eonb200_neb_params_default(EonB200NebParams *p);

/* Host buffers, all contiguous: positions / forces [nBands][nImages + 2][3 nAtoms] (both endpoints included; forces are
 * Matter::getForces values, i.e. fixed axes already zeroed, endpoint rows unused), energies [nBands][nImages + 2], box[9]
 * row-major cell shared by all bands (null = no periodic wrap).  Writes projected and tangents
 * [nBands][nImages][3 nAtoms] (tangents may be null) and climbing[nBands] (may be null): index of the climbing image,
 * 0 when climbing_image is off.  nFreeAtoms == nAtoms switches zeroTranslation on.  Returns 0 or an error code. */
EONB200_API int eonb200_neb_forces(EonB200Pot *pot, long nBands, long nImages, long nAtoms, long nFreeAtoms,
                                   const double *positions, const double *forces, const double *energies,
                                   const double *box, const EonB200NebParams *prm, double *projected, double *tangents,
                                   int *climbing);

/* One NEB force evaluation of many bands in a single call: positions [nBands][nImages + 2][3 nAtoms] go to the device
 * once, every image (the two endpoints ride along, which also yields their energies) is evaluated by the handle's
 * potential in ONE batched launch, Matter::getForces is applied in the band kernel (fixed[3 nAtoms] axes zeroed, or --
 * when nothing is fixed and remove_net_force is set -- the mean force removed; client/Matter.cpp:324-333,531-536), and
 * only the results come back: projected [nBands][nImages][3 nAtoms], energies [nBands][nImages + 2] (nullable),
 * tangents (nullable), climbing[nBands] (nullable).  atomicNrs[nAtoms] and box[9] are shared by all images; fixed may
 * be null.  Replaces the forceBatch call site AND the per-image loop of NudgedElasticBand::updateForces
 * (client/NudgedElasticBand.cpp:475-521,584-643).  The force-call counter advances by nBands * (nImages + 2). */
EONB200_API int eonb200_neb_band_forces(EonB200Pot *pot, long nBands, long nImages, long nAtoms, const double *positions,
                                        const int *atomicNrs, const int *fixed, const double *box, int remove_net_force,
                                        const EonB200NebParams *prm, double *projected, double *energies,
                                        double *tangents, int *climbing);

/* Same with DEVICE pointers for everything but box and prm (bands kept resident between optimiser steps); asynchronous
 * on the handle's stream. */
EONB200_API int eonb200_neb_forces_device(EonB200Pot *pot, long nBands, long nImages, long nAtoms, long nFreeAtoms,
                                          const double *d_positions, const double *d_forces, const double *d_energies,
                                          const double *box, const EonB200NebParams *prm, double *d_projected,
                                          double *d_tangents, int *d_climbing);

/* CUDA-event milliseconds of the last (synchronous) force_batch / force_batch_device call, split as
 * out[0] = k_prepare (staging, Verlet check, cell-list rebuilds), out[1] = the force pass kernel(s). */
EONB200_API int eonb200_last_force_ms(EonB200Pot *pot, double out[2]);

/* Measurement aid: overwrite `bytes` of scratch HBM on the handle's stream (asynchronously), which evicts the L2
 * cache (126 MB on B200) ahead of the next evaluation without letting the stream run dry. */
EONB200_API int eonb200_l2_flush(EonB200Pot *pot, size_t bytes);

/* Profiling aid of the fused EAM kernel: with `enable` != 0 the next launches record, per CTA, 16 words
 * (0: %globaltimer ns at entry, 1: SM id, 2..10: SM cycle counter at entry / positions landed / conversion and
 * Verlet check done / filter done / density done / CTA density done / cluster barrier passed / force done / exit;
 * thread 0 of the CTA).  The half-list kernel (k_eam_n3) uses the words as: 2 entry, 3 coordinates landed, 4 converted,
 * 12 list decision, 6 / 7 density done / barrier passed, 8 embedding done, 9 owner pass done, 15 thread 0 leaves the
 * receiver pass, 14 energy reduction done, 10 exit, and 5 / 11 / 13 = the LAST warp of the CTA leaving the density /
 * owner / receiver loops (atomic max).  Copies the last launch's records (n_ctas = nSystems * CTAs per system) to `out`. */
EONB200_API int eonb200_phase_clocks(EonB200Pot *pot, int enable, long long *out, long cap_ctas, long *n_ctas);

#ifdef __cplusplus
}
#endif
#endif /* EONB200_H */
