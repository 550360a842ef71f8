// capi.cu -- the C ABI of libeonb200.so (include/eonb200.h): handles, HBM residency, launches.
//
// Host logic only; every floating-point result comes from the kernels in kernels.cuh.  There is
// no CPU evaluation path: without a usable CUDA device eonb200_create fails.
#include "kernels.cuh"
#include "neb.cuh"

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

using namespace eonb200;

namespace {

const EonB200EamRow kBuiltinRows[3] = {
    // include/eon/potentials/EAM/Parameters.h:15-50 (Z, Dm, alphaM, Rm, beta1, beta2, r_cut, c0..c8)
    {13, 3.63262, 1.49303, 2.12396, 3.45762, 3.54208, 5.555,
     {-0.0847305, 17.0887, 60.7909, -75.8412, 39.9948, -6.21563, -3.07993, 1.51365, -0.196034}},
    {46, 1.6054, 1.55083, 2.35604, 3.35148, 3.34715, 5.4120,
     {0.0992627, 4.8356, -0.368647, 17.2739, -18.450, 9.5018, -2.69876, 0.405341, -0.0251768}},
    {79, 0.669134, 1.89707, 2.57021, 3.69785, 3.69674, 5.5155,
     {-0.198791, -27.7663, 224.068, -594.101, 1316.31, -2070.62, 1973.14, -1012.02, 214.373}}};

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  // grow-only; contents are NOT preserved
  cudaError_t reserve(size_t count) {
    if (count <= n) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t n = 0;
  ~PinBuf() {
    if (p) cudaFreeHost(p);
  }
  cudaError_t reserve(size_t count) {
    if (count <= n) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&p), std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

struct NlStore {
  DevBuf<int> nbr, cnt, state, flags, done, nbr2, cnt2, cls, cellscratch;
  int scratch_stride = 0;
  int rowmax_seen = 0;  // largest row fill any build of this store reported (rows worth prefetching)
  DevBuf<double> rbuild, pairA, pairB, dfd, upart, box_build;
  DevBuf<double4> pos4;
  DevBuf<unsigned char> rowid;
  // owner / receiver lists of the Newton's-third-law EAM kernel (types.cuh: NlBatch::own_*)
  DevBuf<int> recv_code, roff, hl_hdr, lane_atom, atom_lane;
  DevBuf<int2> own_cs;
  DevBuf<int4> lhdr;
  DevBuf<double2> own_ab;
  DevBuf<double> lo_build;
  int n3 = 0;          // lists allocated
  int slots_seen = 0;  // most receiver slots any walker of this store needed (sizes the launch's shared memory)
  long S = 0, N = 0;
  int K = 0, Np = 0;
};

}  // namespace

constexpr int kPipeMax = 8;  // most chunks a batch is split into

struct EonB200Pot {
  EonB200Config cfg;
  PotParams hp;
  std::vector<EonB200EamRow> rows;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  // chunked host <-> device pipeline of eonb200_force_batch: copy-in stream, second compute stream, copy-out stream
  cudaStream_t s_in = nullptr, s_k2 = nullptr, s_out = nullptr;
  cudaEvent_t ev_m = nullptr, ev_in[kPipeMax] = {}, ev_k[kPipeMax] = {}, ev_out[kPipeMax] = {};
  DevBuf<PotParams> dP;
  DevBuf<double> d_exp_tab;
  DevBuf<unsigned char> d_flush;  // eonb200_l2_flush scratch
  DevBuf<long long> d_clocks;  // per-CTA phase clocks of the last fused EAM launch (profiling aid)
  int clocks_on = 0;
  long clocks_ctas = 0;
  // counters of a launch: stats[0..3], and the row-capacity flag in the word behind them (d_ovf / h_ovf), so that one
  // memset clears and one small copy fetches them all
  DevBuf<unsigned long long> d_stats;
  PinBuf<unsigned long long> h_stats;
  int *d_ovf = nullptr, *h_ovf = nullptr;
  // stateless batch staging
  DevBuf<double> dR, dF, dU, dbox;
  DevBuf<int> dZ;
  PinBuf<double> hR, hF, hU, hbox;
  PinBuf<int> hZ;
  std::vector<double> last_boxes;
  long z_cached_S = 0, z_cached_N = 0;  // shape of the atomic numbers held in hZ / dZ (0: none)
  std::vector<std::pair<const char *, size_t>> pinned;  // caller ranges page-locked by this handle (pin_host_buffers)
  // systems scattered over the caller's heap (one block per Matter): per-system device-visible addresses
  long pin_registrations = 0;
  DevBuf<unsigned long long> d_rowptr;  // [2 S]: coordinates rows, then force rows
  PinBuf<unsigned long long> h_rowptr;
  std::vector<unsigned long long> rowptr_sent;
  NlStore nl;  // cache of the stateless batch slots
  DevBuf<double> neb_R, neb_F, neb_E, neb_P, neb_T;  // eonb200_neb_forces staging
  DevBuf<int> neb_C, neb_Z;
  DevBuf<double> neb_box, neb_mask;
  std::vector<double> neb_hmask;
  int neb_attr = 0;
  unsigned long long force_calls = 0, launches = 0;
  long long last_stats[5] = {0, 0, 0, 0, 0};
  double last_ms[2] = {0.0, 0.0};
  int max_dyn_smem = 0;
  int sm_count = 148;
  std::string err;
};

struct EonB200Walkers {
  EonB200Pot *pot = nullptr;
  long W = 0, N = 0;
  int D = 0, Dp = 0, nfree = 0, any_fixed = 0;
  DevBuf<double> pos, frc, mode, energy, ws, mask, box, react, eig, conv;
  DevBuf<int> dirty, free_atoms, Z, status;
  DevBuf<long long> iters, fcalls;
  DevBuf<EonB200SearchParams> prm;
  NlStore nl;
  size_t ws_stride = 0;
  size_t off_lbfgs_s = 0, off_lbfgs_y = 0, off_lanczos_q = 0, off_rot_s = 0, off_rot_y = 0;
  std::vector<double> h_box;
  double last_ms = 0.0;
  bool have_modes = false;
  // in-kernel queue (eonb200_walkers_queue): per-entry inputs / results
  DevBuf<double> q_pos, q_mode, q_react, q_energy, q_eig, q_conv;
  DevBuf<int> q_status, q_next;
  DevBuf<long long> q_iter, q_fcalls;
  DevBuf<int> active;       // [W] 1 = takes part in the next launches
  bool use_active = false;
  // group-wise asynchronous jobs (eonb200_walkers_submit): one stream and one completion event per group
  std::vector<cudaStream_t> g_stream;
  std::vector<cudaEvent_t> g_done;
  std::vector<std::array<long, 2>> g_range;  // [first, count] of the job in flight on the group, count 0 = idle
};

namespace {

int fail(EonB200Pot *pot, int rc, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (pot) pot->err = buf;
  return rc;
}

#define CU_TRY(pot, call)                                                                      \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) return fail((pot), 700 + static_cast<int>(e_), "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

void fill_params(EonB200Pot *pot) {
  const EonB200Config &c = pot->cfg;
  PotParams &P = pot->hp;
  memset(&P, 0, sizeof(P));
  P.kind = c.kind;
  P.lj_u0 = c.lj_u0;
  P.lj_psi = c.lj_psi;
  P.morse_De = c.morse_De;
  P.morse_a = c.morse_a;
  P.morse_re = c.morse_re;
  P.eam_conservative = c.eam_conservative;
  P.skin = c.skin;
  P.cutoff = c.cutoff;
  P.eshift = 0.0;
  if (c.kind == EONB200_POT_EAM) {
    P.nl_mode = NL_MIC;
    P.n_rows = static_cast<int>(pot->rows.size());
    double rcmax = 0.0;
    for (int r = 0; r < P.n_rows; r++) {
      const EonB200EamRow &s = pot->rows[r];
      EamRowDev &d = P.rows[r];
      d.Z = s.Z;
      d.Dm = s.Dm;
      d.alphaM = s.alphaM;
      d.Rm = s.Rm;
      d.beta1 = s.beta1;
      d.beta2 = s.beta2;
      d.r_cut = s.r_cut;
      d.rc2 = s.r_cut * s.r_cut;
      for (int k = 0; k < 9; k++) {
        d.c[k] = s.func_coeff[k];
        d.dc[k] = k * s.func_coeff[k];
      }
      rcmax = std::max(rcmax, s.r_cut);
    }
    P.cutoff = rcmax;
    // exp_tab_fast has no underflow guard: every exponent argument a row can produce inside its cutoff must stay
    // well inside the double range
    P.eam_fast_ok = 1;
    for (int r = 0; r < P.n_rows; r++) {
      const EonB200EamRow &s = pot->rows[r];
      const double worst = std::max(std::max(std::fabs(s.beta1), 2.0 * std::fabs(s.beta2)), std::fabs(s.alphaM)) *
                           (s.r_cut + std::fabs(s.Rm));
      if (!(worst < 600.0)) P.eam_fast_ok = 0;
    }
  } else {
    P.nl_mode = NL_VESIN;
    const double rc = P.cutoff;
    if (c.kind == EONB200_POT_LJ) {
      const double s6 = std::pow(P.lj_psi / rc, 6);
      P.eshift = 4.0 * P.lj_u0 * (s6 * s6 - s6);
    } else if (c.kind == EONB200_POT_MORSE_PT) {
      const double d = 1.0 - std::exp(-P.morse_a * (rc - P.morse_re));
      P.eshift = P.morse_De * d * d - P.morse_De;
      // the unguarded exp of the pair loop is fine while the exponent stays far inside the double range
      P.eam_fast_ok = std::fabs(P.morse_a) * (rc + std::fabs(P.morse_re)) < 600.0 ? 1 : 0;
    }
  }
  P.cutoff2 = P.cutoff * P.cutoff;
  P.rc_list = P.cutoff + P.skin;
  P.rc_list2 = P.rc_list * P.rc_list;
}

int pick_lpa(const EonB200Pot *pot, long N, int threads_cap) {
  (void)pot;
#ifdef EONB200_EMU
  (void)N;
  (void)threads_cap;
  return 1;
#else
  int lpa = 1;
  while (lpa < 8 && N * (lpa * 2) <= threads_cap) lpa *= 2;
  return lpa;
#endif
}

int pick_threads(const EonB200Pot *pot, long N, int lpa) {
#ifdef EONB200_EMU
  (void)pot;
  (void)N;
  (void)lpa;
  return 1;
#else
  if (pot->cfg.threads_per_walker > 0) {
    int t = (pot->cfg.threads_per_walker + 31) / 32 * 32;
    return std::min(std::max(t, 32), 512);
  }
  long t = (N * lpa + 31) / 32 * 32;
  return static_cast<int>(std::min<long>(std::max<long>(t, 32), 512));
#endif
}

// Tiling of the stateless force passes: enough CTAs to give every SM several, at most kMaxTiles per system,
// tiles of at least 32 atoms; lanes per atom so that a tile fills a CTA of a few hundred threads.
void pick_tiling(const EonB200Pot *pot, long S, long N, int *tiles, int *lpa) {
#ifdef EONB200_EMU
  (void)pot;
  *tiles = S > 1 ? 2 : 3;  // exercise the multi-tile path
  if (*tiles > N) *tiles = static_cast<int>(N);
  *lpa = 1;
#else
  const long want_ctas = 14L * pot->sm_count;  // (Morse_Pt, 256 x 343 atoms on B200: 3 tiles x 2 lanes per atom 205 us, 8 x 4 185 us)
  long t = (want_ctas + S - 1) / S;
  t = std::max<long>(1, std::min<long>(t, kMaxTiles));
  t = std::min<long>(t, std::max<long>(1, N / 32));
  if (pot->cfg.tiles_override > 0) t = std::min<long>(pot->cfg.tiles_override, std::min<long>(kMaxTiles, N));
  *tiles = static_cast<int>(t);
  const long tile_atoms = (N + t - 1) / t;
  int l = 1;
  while (l < 8 && tile_atoms * (l * 2) <= 256) l *= 2;
  if (pot->cfg.lpa_override > 0) l = pot->cfg.lpa_override;
  *lpa = l;
#endif
}

int tile_threads_for(int tile_atoms, int lpa) {
#ifdef EONB200_EMU
  (void)tile_atoms;
  (void)lpa;
  return 1;
#else
  long t = (static_cast<long>(tile_atoms) * lpa + 31) / 32 * 32;
  return static_cast<int>(std::min<long>(std::max<long>(t, 32), kMaxThreadsPerWalker));
#endif
}

// Fused EAM kernel: CTAs per walker (= cluster size, at most 8) and threads per CTA.  Atoms are dealt out in whole
// warps; CTAs of three to four warps keep the load per SM even when a few hundred walkers meet 148 SMs, a
// handful of systems (the dimer's two images) is spread as wide as the cluster allows.
void pick_fused(const EonB200Pot *pot, long S, long N, int *T_out, int *threads_out) {
#ifdef EONB200_EMU
  (void)pot;
  (void)S;
  (void)N;
  *T_out = 1;
  *threads_out = 1;
#else
  const long Wn = (N + 31) / 32;
  const long Tmin = (Wn + 15) / 16;  // at most 512 threads per CTA
  long T;
  if (smem_bytes_fused(static_cast<int>(N)) > 100 * 1024) {
    T = Tmin;  // one large CTA per SM: the walker's positions are replicated in every CTA of its cluster
  } else if (S * Wn < 4L * pot->sm_count) {
    T = Wn;
  } else {
    // the fewest CTAs per walker that stay within 448 threads: under the 72-register budget two such CTAs (or four
    // of 224 threads) share an SM, so a few hundred walkers are all resident in one round, and the prologue (whole-
    // walker staging, Verlet check) is replicated as little as possible.  Measured on 256 x 385 atoms: one CTA per
    // walker 79 us, two 81 us, four 108 us (clusters of four no longer pack into the GPCs in one round).
    T = (Wn + 13) / 14;
  }
  if (pot->cfg.tiles_override > 0) T = pot->cfg.tiles_override;
  T = std::max<long>(std::max<long>(T, Tmin), 1);
  T = std::min<long>(std::min<long>(T, 8), Wn);
  if (pot->cfg.tiles_override <= 0) {  // cluster sizes 1, 2, 4, 8
    long p2 = 1;
    while (p2 < T) p2 *= 2;
    T = std::min<long>(p2, 8);
    while (T > 1 && T > Wn) T /= 2;
  }
  long wpc = (Wn + T - 1) / T;
  // the largest slice of the warp-granular split w0 = rank * Wn / T
  for (long r = 0; r < T; r++) wpc = std::max(wpc, (r + 1) * Wn / T - r * Wn / T);
  long threads = 32 * wpc;
  if (pot->cfg.threads_per_walker > 0) threads = std::max<long>(32, (pot->cfg.threads_per_walker + 31) / 32 * 32);
  *T_out = static_cast<int>(T);
  *threads_out = static_cast<int>(std::min<long>(threads, kMaxThreadsPerWalker));
#endif
}

// Newton's-third-law EAM kernel: one lane per atom up to 1024 atoms, more atoms per lane beyond; two lanes per atom
// (half the row per lane) on request.
int pick_n3_lpa(const EonB200Pot *pot, long N) {
#ifdef EONB200_EMU
  (void)pot;
  (void)N;
  return 1;  // (one thread per CTA: no partner lane)
#else
  static const int forced = getenv("EONB200_N3_LPA") ? atoi(getenv("EONB200_N3_LPA")) : 0;  // development aid
  if (pot->cfg.threads_per_walker > 0) return pot->cfg.threads_per_walker >= 2 * N && forced != 1 ? 2 : 1;
  if (forced == 1 || forced == 2) return 2 * N <= 1024 ? forced : 1;
  return 1;  // measured on B200 (C2, 385 atoms): two lanes per atom 85 us per step against 73 us with one
#endif
}
int pick_n3_threads(const EonB200Pot *pot, long N, int lpa) {
#ifdef EONB200_EMU
  (void)pot;
  (void)N;
  (void)lpa;
  return 1;
#else
  long t;
  if (pot->cfg.threads_per_walker > 0)
    t = (pot->cfg.threads_per_walker + 31) / 32 * 32;
  else if (lpa * N <= 1024)
    t = std::max<long>(64, (lpa * N + 31) / 32 * 32);
  else
    t = std::min<long>(1024, (N + 1) / 2 + 31) / 32 * 32;
  return static_cast<int>(std::min<long>(std::max<long>(t, 32), 1024));
#endif
}

bool use_n3(const EonB200Pot *pot, long N) {
  return pot->cfg.kind == EONB200_POT_EAM && pot->cfg.eam_pair_mode != 1 &&
         N <= kN3MaxAtoms && static_cast<long>(smem_bytes_n3(static_cast<int>(N), 0, 4)) + 8 * 8 * N <= pot->max_dyn_smem &&
         pot->cfg.skin > 0.0;
}

// Receiver slots a launch wants shared memory for: what the lists built so far needed plus a margin, or -- before the
// first build -- an estimate from the row capacity.
long n3_slots_wanted(const NlStore &nl, long N) {
  const long want = nl.slots_seen > 0 ? nl.slots_seen + nl.slots_seen / 16 + 64 : N * nl.K / 4;
  return std::max<long>(want, 3 * N + 4);  // the raw coordinates land in the slot buffer
}

// Copies of every atom in the kernel's shared position arrays, as n3_shift = 4 + log2(copies): eight (conflict-free
// gathers, eam_n3.cuh) when they fit beside the receiver slots, fewer for larger systems.  Lists carry the shift they
// were built for (N3H_VALID); slots_seen only grows, so a store changes its shift at most three times.
int pick_n3_shift(const EonB200Pot *pot, const NlStore &nl, long N) {
#ifdef EONB200_EMU
  (void)pot;
  (void)nl;
  (void)N;
  return 5;  // (two copies: the emulation exercises the copy arithmetic)
#else
  static const int forced = getenv("EONB200_N3_REP") ? atoi(getenv("EONB200_N3_REP")) : 0;  // development aid: 1, 2, 4, 8
  const long want = n3_slots_wanted(nl, N);
  for (int shift = 7; shift > 4; shift--) {
    if (forced > 0 && (16 << (shift - 4)) != 16 * forced) continue;
    if (static_cast<long>(smem_bytes_n3(static_cast<int>(N), static_cast<int>(want), shift)) <= pot->max_dyn_smem) return shift;
  }
  return 4;
#endif
}

// Receiver slots the next launch reserves shared memory for; never more than the SM has.
int pick_slot_cap(const EonB200Pot *pot, const NlStore &nl, long N, int shift) {
  long want = n3_slots_wanted(nl, N);
  const long room = (pot->max_dyn_smem - static_cast<long>(smem_bytes_n3(static_cast<int>(N), 0, shift))) / 8 - 16;
  want = std::min(want, room);
  // two CTAs per SM when that is within reach
  const long two = ((227 * 1024) / 2 - 1024 - static_cast<long>(smem_bytes_n3(static_cast<int>(N), 0, shift))) / 8 - 16;
  if (nl.slots_seen > 0 && nl.slots_seen <= two) want = std::min(want, two);
  want = std::max<long>(want, 3 * N + 4);  // the raw coordinates land in the slot buffer
  return static_cast<int>((std::max<long>(want, 0) + 15) / 16 * 16);
}

// Argument checks shared by the entry points that take host atomic numbers: the packed neighbour codes hold
// kNbrIndexMask atoms, and an EAM element without a parameter row is the reference's 14324 (EAM.cpp:116-123).
int check_system(EonB200Pot *pot, long N, const int *Z, const char *who) {
  if (N > kNbrIndexMask) return fail(pot, 2, "%s: more than %d atoms per system", who, kNbrIndexMask);
  if (pot->cfg.kind == EONB200_POT_EAM && Z) {
    for (long i = 0; i < N; i++) {
      bool found = false;
      for (const auto &r : pot->rows) found = found || r.Z == Z[i];
      if (!found) return fail(pot, 14324, "EAM: no parameter row for atomic number %d", Z[i]);
    }
  }
  return 0;
}

int pick_max_cells(long N) { return static_cast<int>(std::min<long>(std::max<long>(64, N), 4096)); }

// Row capacity guess from the mean density; truncated builds are detected and regrown.
int guess_K(double rc, long N, const double *box) {
  const double *m = box;
  const double det = std::fabs(m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                               m[2] * (m[3] * m[7] - m[4] * m[6]));
  double expect = det > 1e-30 ? static_cast<double>(N) / det * (4.0 / 3.0) * kPi * rc * rc * rc : static_cast<double>(N);
  long K = static_cast<long>(std::ceil(1.5 * expect)) + 16;
  K = std::min<long>(K, std::max<long>(N - 1, 16) + 32);
  K = (K + 7) / 8 * 8;
  return static_cast<int>(std::max<long>(K, 16));  // (the half-list EAM kernel requests its first 16 rows unconditionally)
}

int nl_alloc(EonB200Pot *pot, NlStore &nl, long S, long N, int K, bool n3_lists = false) {
  const int Np = static_cast<int>((N + 31) / 32 * 32);
  CU_TRY(pot, nl.nbr.reserve(static_cast<size_t>(S) * K * Np));
  CU_TRY(pot, nl.cnt.reserve(static_cast<size_t>(S) * Np));
  CU_TRY(pot, nl.rbuild.reserve(static_cast<size_t>(S) * 3 * N));
  CU_TRY(pot, nl.state.reserve(static_cast<size_t>(S)));
  CU_TRY(pot, nl.flags.reserve(static_cast<size_t>(S)));
  CU_TRY(pot, nl.done.reserve(static_cast<size_t>(S)));
  CU_TRY(pot, nl.pos4.reserve(static_cast<size_t>(S) * Np));
  CU_TRY(pot, nl.dfd.reserve(static_cast<size_t>(S) * Np));
  CU_TRY(pot, nl.rowid.reserve(static_cast<size_t>(S) * Np));
  CU_TRY(pot, nl.upart.reserve(static_cast<size_t>(S) * kMaxTiles * 4));
  // certainty classes of the rows (and the scratch column the build parks the uncertain codes in): every potential
  CU_TRY(pot, nl.nbr2.reserve(static_cast<size_t>(S) * K * Np));
  CU_TRY(pot, nl.cls.reserve(static_cast<size_t>(S) * Np));
  if (pot->cfg.kind == EONB200_POT_EAM) {
    CU_TRY(pot, nl.pairA.reserve(static_cast<size_t>(S) * K * Np));
    CU_TRY(pot, nl.pairB.reserve(static_cast<size_t>(S) * K * Np));
    CU_TRY(pot, nl.cnt2.reserve(static_cast<size_t>(S) * Np));
    nl.scratch_stride = static_cast<int>(fused_scratch_ints(static_cast<int>(N), pick_max_cells(N)));
    CU_TRY(pot, nl.cellscratch.reserve(static_cast<size_t>(S) * nl.scratch_stride));
    nl.n3 = 0;
    nl.slots_seen = 0;
    if (n3_lists && use_n3(pot, N)) {
      nl.n3 = 1;
      const size_t cells = static_cast<size_t>(S) * K * Np;
      const size_t slack = static_cast<size_t>(kN3Slack) * Np;  // rows the pair loops request ahead of a row's end
      CU_TRY(pot, nl.own_cs.reserve(cells + slack));
      CU_TRY(pot, nl.recv_code.reserve(cells + slack));
      CU_TRY(pot, nl.own_ab.reserve(cells + slack));
      CU_TRY(pot, cudaMemsetAsync(nl.own_cs.p, 0, sizeof(int2) * (cells + slack), pot->stream));
      CU_TRY(pot, cudaMemsetAsync(nl.recv_code.p, 0, sizeof(int) * (cells + slack), pot->stream));
      CU_TRY(pot, cudaMemsetAsync(nl.own_ab.p, 0, sizeof(double2) * (cells + slack), pot->stream));
      CU_TRY(pot, nl.lhdr.reserve(static_cast<size_t>(S) * Np * 2));
      CU_TRY(pot, nl.roff.reserve(static_cast<size_t>(S) * (2 * Np + 1)));
      CU_TRY(pot, nl.lane_atom.reserve(static_cast<size_t>(S) * Np));
      CU_TRY(pot, nl.atom_lane.reserve(static_cast<size_t>(S) * Np));
      CU_TRY(pot, cudaMemsetAsync(nl.lane_atom.p, 0, sizeof(int) * S * Np, pot->stream));
      CU_TRY(pot, nl.lo_build.reserve(static_cast<size_t>(S) * 4));
      CU_TRY(pot, cudaMemsetAsync(nl.lo_build.p, 0, sizeof(double) * 4 * S, pot->stream));
      CU_TRY(pot, nl.hl_hdr.reserve(static_cast<size_t>(S) * 8));
      CU_TRY(pot, cudaMemsetAsync(nl.hl_hdr.p, 0, sizeof(int) * 8 * S, pot->stream));
      CU_TRY(pot, cudaMemsetAsync(nl.lhdr.p, 0, sizeof(int4) * S * Np * 2, pot->stream));
      CU_TRY(pot, cudaMemsetAsync(nl.roff.p, 0, sizeof(int) * S * (2 * Np + 1), pot->stream));
    }
  }
  CU_TRY(pot, nl.box_build.reserve(static_cast<size_t>(S) * 9));
  CU_TRY(pot, cudaMemsetAsync(nl.state.p, 0, sizeof(int) * S, pot->stream));
  CU_TRY(pot, cudaMemsetAsync(nl.done.p, 0, sizeof(int) * S, pot->stream));
  CU_TRY(pot, cudaMemsetAsync(nl.flags.p, 0, sizeof(int) * S, pot->stream));
  nl.S = S;
  nl.N = N;
  nl.K = K;
  nl.Np = Np;
  nl.rowmax_seen = 0;
  return 0;
}

NlBatch nl_view(EonB200Pot *pot, NlStore &nl) {
  NlBatch b;
  b.nbr = nl.nbr.p;
  b.cnt = nl.cnt.p;
  b.rbuild = nl.rbuild.p;
  b.state = nl.state.p;
  b.K = nl.K;
  b.Np = nl.Np;
  b.overflow = pot->d_ovf;
  b.stats = pot->d_stats.p;
  b.pairA = nl.pairA.p;
  b.pairB = nl.pairB.p;
  b.nbr2 = nl.nbr2.p;
  b.cnt2 = nl.cnt2.p;
  b.cls = nl.cls.p;
  b.cellscratch = nl.cellscratch.p;
  b.scratch_stride = nl.scratch_stride;
  b.pos4 = nl.pos4.p;
  b.dfd = nl.dfd.p;
  b.rowid = nl.rowid.p;
  b.flags = nl.flags.p;
  b.upart = nl.upart.p;
  b.done = nl.done.p;
  b.box_build = nl.box_build.p;
  b.own_cs = nl.n3 ? nl.own_cs.p : nullptr;
  b.recv_code = nl.n3 ? nl.recv_code.p : nullptr;
  b.own_ab = nl.n3 ? nl.own_ab.p : nullptr;
  b.lhdr = nl.n3 ? nl.lhdr.p : nullptr;
  b.roff = nl.n3 ? nl.roff.p : nullptr;
  b.lane_atom = nl.n3 ? nl.lane_atom.p : nullptr;
  b.atom_lane = nl.n3 ? nl.atom_lane.p : nullptr;
  b.hl_hdr = nl.n3 ? nl.hl_hdr.p : nullptr;
  b.lo_build = nl.n3 ? nl.lo_build.p : nullptr;
  b.slot_cap = 0;
  b.n3_lpa = 1;
  b.n3_shift = 4;
  return b;
}

int check_smem(EonB200Pot *pot, long N, int maxCells, size_t *bytes) {
  *bytes = smem_bytes(static_cast<int>(N), maxCells);
  if (static_cast<long>(*bytes) > pot->max_dyn_smem)
    return fail(pot, 3, "system of %ld atoms needs %zu bytes of shared memory per walker CTA (limit %d)", N, *bytes,
                pot->max_dyn_smem);
  return 0;
}

// pin_host_buffers: are the S per-system arrays one contiguous block?  Then page-lock it (once) for direct DMA.
template <typename T>
bool rows_contiguous(T *const *rows, long S, size_t elems) {
  for (long s = 1; s < S; s++)
    if (rows[s] != rows[0] + static_cast<size_t>(s) * elems) return false;
  return true;
}

bool pin_range(EonB200Pot *pot, const void *p, size_t bytes) {
#ifdef EONB200_EMU
  (void)pot;
  (void)p;
  (void)bytes;
  return false;
#else
  if (!pot->cfg.pin_host_buffers || bytes == 0) return false;
  const char *b = static_cast<const char *>(p);
  const char *e = b + bytes;
  for (auto &r : pot->pinned)
    if (b >= r.first && e <= r.first + r.second) return true;
  // ranges of ours that overlap the new one are given up and re-registered as one (registrations must not overlap)
  for (size_t k = 0; k < pot->pinned.size();) {
    const char *rb = pot->pinned[k].first, *re = rb + pot->pinned[k].second;
    if (rb < e && b < re) {
      cudaHostUnregister(const_cast<char *>(rb));
      b = std::min(b, rb);
      e = std::max(e, re);
      pot->pinned.erase(pot->pinned.begin() + k);
    } else {
      k++;
    }
  }
  if (pot->pinned.size() >= 32) {  // bounded cache: drop the oldest registration
    cudaHostUnregister(const_cast<char *>(pot->pinned.front().first));
    pot->pinned.erase(pot->pinned.begin());
  }
  // A caller that hands over fresh memory on every call (new Matter objects each time) would pay a registration --
  // hundreds of microseconds per megabyte -- per call: after 256 of them this handle stops and stages instead.
  if (++pot->pin_registrations > 256) return false;
  if (cudaHostRegister(const_cast<char *>(b), static_cast<size_t>(e - b), cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();  // e.g. overlaps a range somebody else registered: fall back to staging
    return false;
  }
  pot->pinned.emplace_back(b, static_cast<size_t>(e - b));
  return true;
#endif
}

// Page-lock the span of the caller's heap that holds the S rows of `elems` values each (and the rows of a second
// array), when it is not much larger than the rows themselves: one block per system, allocated one after the other,
// is what eonclient's Matter objects look like from here.
template <typename T, typename U>
bool pin_rows_span(EonB200Pot *pot, T *const *rows, U *const *rows2, long S, size_t elems) {
#ifdef EONB200_EMU
  (void)pot; (void)rows; (void)rows2; (void)S; (void)elems;
  return false;
#else
  if (!pot->cfg.pin_host_buffers || S <= 0) return false;
  uintptr_t lo = ~static_cast<uintptr_t>(0), hi = 0;
  for (long s = 0; s < S; s++) {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(rows[s]), a1 = reinterpret_cast<uintptr_t>(rows2[s]);
    lo = std::min(lo, std::min(a0, a1));
    hi = std::max(hi, std::max(a0, a1) + elems * sizeof(double));
  }
  const size_t payload = 2 * static_cast<size_t>(S) * elems * sizeof(double);
  lo &= ~static_cast<uintptr_t>(4095);
  hi = (hi + 4095) & ~static_cast<uintptr_t>(4095);
  if (hi - lo > 8 * payload + (4u << 20)) return false;
  return pin_range(pot, reinterpret_cast<const void *>(lo), hi - lo);
#endif
}

// Launch the batched force evaluation on device-resident arrays; handles truncated lists by
// regrowing the row capacity and re-running.  Synchronises when it has to look at the flags.
// A batch split into chunks of systems whose copy-in, kernel and copy-out overlap on three streams (chunk c + 1 comes
// up while chunk c computes and chunk c - 1 goes back).  in / out enqueue the copies of systems [s0, s1) on the stream
// they are given.
struct BatchPipe {
  int chunks = 1;
  bool no_out = false;  // the kernel writes the results into the caller's memory itself: nothing to copy back
  std::function<int(long s0, long s1, cudaStream_t st)> in, out;
};

int run_force_batch(EonB200Pot *pot, NlStore &nl, ForceBatchArgs a, const double *h_box0, bool sync_and_check,
                    const std::function<int()> &enqueue_readback = nullptr, const BatchPipe *pipe = nullptr) {
  const long S = a.S, N = a.N;
  if (nl.S < S || nl.N != N || nl.K == 0) {
    double unit[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const int K = std::max(nl.N == N ? nl.K : 0, guess_K(pot->hp.rc_list, N, h_box0 ? h_box0 : unit));
    int rc = nl_alloc(pot, nl, S, N, K, true);
    if (rc) return rc;
  }
  a.maxCells = pick_max_cells(N);
  const bool eam = pot->cfg.kind == EONB200_POT_EAM;
  const int prep_threads = pick_threads(pot, N, 1);
  size_t smem = 0;
  int rc = 0;
  // LJ / Morse force pass: T tiles per system so that S * T CTAs cover the SMs a few times over
  int tile_threads = 0;
  size_t tile_smem = 0;
  // EAM: one cluster of T CTAs per walker
  int fused_T = 1, fused_threads = 1;
  size_t fused_smem = 0;
  const bool n3 = eam && use_n3(pot, N) && nl.n3;
  int slot_cap = 0;
  if (n3) {
    fused_T = 1;
    a.lpa = pick_n3_lpa(pot, N);
    fused_threads = pick_n3_threads(pot, N, a.lpa);
    a.tiles = 1;
    a.kmax = nl.K;
    a.clocks = nullptr;
    if (pot->clocks_on) {
      pot->clocks_ctas = S;
      CU_TRY(pot, pot->d_clocks.reserve(static_cast<size_t>(pot->clocks_ctas) * 16));
      CU_TRY(pot, cudaMemsetAsync(pot->d_clocks.p, 0, sizeof(long long) * pot->clocks_ctas * 16, pot->stream));
      a.clocks = pot->d_clocks.p;
    }
  } else if (eam) {
    pick_fused(pot, S, N, &fused_T, &fused_threads);
    fused_smem = smem_bytes_fused(static_cast<int>(N));
    if (static_cast<long>(fused_smem) > pot->max_dyn_smem)
      return fail(pot, 3, "system of %ld atoms needs %zu bytes of shared memory per CTA (limit %d)", N, fused_smem,
                  pot->max_dyn_smem);
    a.tiles = fused_T;
    a.lpa = 1;
    a.kmax = nl.rowmax_seen > 0 ? std::min(nl.rowmax_seen, nl.K) : nl.K;
    a.clocks = nullptr;
    if (pot->clocks_on) {
      pot->clocks_ctas = S * fused_T;
      CU_TRY(pot, pot->d_clocks.reserve(static_cast<size_t>(pot->clocks_ctas) * 16));
      CU_TRY(pot, cudaMemsetAsync(pot->d_clocks.p, 0, sizeof(long long) * pot->clocks_ctas * 16, pot->stream));
      a.clocks = pot->d_clocks.p;
    }
  } else {
    rc = check_smem(pot, N, a.maxCells, &smem);
    if (rc) return rc;
    pick_tiling(pot, S, N, &a.tiles, &a.lpa);
    const int tile_atoms = static_cast<int>((N + a.tiles - 1) / a.tiles);
    tile_threads = tile_threads_for(tile_atoms, a.lpa);
    tile_smem = smem_bytes_tile(static_cast<int>(N));
    if (static_cast<long>(tile_smem) > pot->max_dyn_smem) return fail(pot, 3, "system too large for the tiled force kernels");
  }
  for (int attempt = 0; attempt < 6; attempt++) {
    CU_TRY(pot, cudaMemsetAsync(pot->d_stats.p, 0, 5 * sizeof(unsigned long long), pot->stream));  // (stats and d_ovf)
    NlBatch view = nl_view(pot, nl);
    if (n3) {
      const int shift = pick_n3_shift(pot, nl, N);
      slot_cap = pick_slot_cap(pot, nl, N, shift);
      fused_smem = smem_bytes_n3(static_cast<int>(N), slot_cap, shift);
      view.slot_cap = slot_cap;
      view.n3_lpa = a.lpa;
      view.n3_shift = shift;
    }
    // one launch (or launch pair) over Sc systems of the view vw / the arrays of ac, on stream st
    auto launch = [&](const NlBatch &vw, const ForceBatchArgs &ac, long Sc, cudaStream_t st) -> int {
    if (n3) {
        const dim3 grid(1u, static_cast<unsigned>(Sc));
#ifdef EONB200_EMU
        EONB200_LAUNCH(k_eam_n3, grid, fused_threads, fused_smem, st, pot->hp, pot->dP.p, vw, ac);
#else
        static const int n3_minb = getenv("EONB200_N3_MINB") ? atoi(getenv("EONB200_N3_MINB")) : 0;  // development aid
        const bool two = n3_minb != 1 && 2 * (fused_smem + 1024) <= 227 * 1024;
        if (n3_minb == 2 && fused_threads <= 224 && two)
          EONB200_LAUNCH((k_eam_n3_t<224, 2>), grid, fused_threads, fused_smem, st, pot->hp, pot->dP.p, vw, ac);
        else if (n3_minb != 1 && fused_threads <= 224 && 4 * (fused_smem + 1024) <= 227 * 1024)
          EONB200_LAUNCH((k_eam_n3_t<224, 4>), grid, fused_threads, fused_smem, st, pot->hp, pot->dP.p, vw, ac);
        else if (fused_threads <= 448 && two)
          EONB200_LAUNCH((k_eam_n3_t<448, 2>), grid, fused_threads, fused_smem, st, pot->hp, pot->dP.p, vw, ac);
        else if (fused_threads <= 512 && two)
          EONB200_LAUNCH((k_eam_n3_t<512, 2>), grid, fused_threads, fused_smem, st, pot->hp, pot->dP.p, vw, ac);
        else if (fused_threads <= 416)  // (one CTA per SM: up to 152 registers per thread)
          EONB200_LAUNCH((k_eam_n3_t<416, 1>), grid, fused_threads, fused_smem, st, pot->hp, pot->dP.p, vw, ac);
        else if (fused_threads <= 512)
          EONB200_LAUNCH((k_eam_n3_t<512, 1>), grid, fused_threads, fused_smem, st, pot->hp, pot->dP.p, vw, ac);
        else if (fused_threads <= 800)
          EONB200_LAUNCH((k_eam_n3_t<800, 1>), grid, fused_threads, fused_smem, st, pot->hp, pot->dP.p, vw, ac);
        else
          EONB200_LAUNCH((k_eam_n3_t<1024, 1>), grid, fused_threads, fused_smem, st, pot->hp, pot->dP.p, vw, ac);
#endif
        pot->launches += 1;
      } else if (eam) {
        const dim3 grid(static_cast<unsigned>(fused_T), static_cast<unsigned>(Sc));
#ifdef EONB200_EMU
        EONB200_LAUNCH_CLUSTER(k_eam_fused, grid, fused_threads, fused_smem, st, fused_T, pot->dP.p, vw, ac);
#else
        // register budgets: every walker of a few-hundred-walker batch must be resident at once
        if (fused_threads <= 128)
          EONB200_LAUNCH_CLUSTER((k_eam_fused_t<128, 7>), grid, fused_threads, fused_smem, st, fused_T, pot->dP.p, vw, ac);
        else if (fused_threads <= 224)
          EONB200_LAUNCH_CLUSTER((k_eam_fused_t<224, 4>), grid, fused_threads, fused_smem, st, fused_T, pot->dP.p, vw, ac);
        else if (fused_threads <= 448)
          EONB200_LAUNCH_CLUSTER((k_eam_fused_t<448, 2>), grid, fused_threads, fused_smem, st, fused_T, pot->dP.p, vw, ac);
        else if (2 * (fused_smem + 1024) <= 227 * 1024)
          EONB200_LAUNCH_CLUSTER((k_eam_fused_t<512, 2>), grid, fused_threads, fused_smem, st, fused_T, pot->dP.p, vw, ac);
        else
          EONB200_LAUNCH_CLUSTER((k_eam_fused_t<512, 1>), grid, fused_threads, fused_smem, st, fused_T, pot->dP.p, vw, ac);
#endif
        pot->launches += 1;
      } else {
        EONB200_LAUNCH(k_prepare, dim3(static_cast<unsigned>(Sc)), prep_threads, smem, st, pot->dP.p, vw, ac);
        CU_TRY(pot, cudaEventRecord(pot->ev[1], st));
        const dim3 grid(static_cast<unsigned>(a.tiles), static_cast<unsigned>(Sc));
        EONB200_LAUNCH(k_pair, grid, tile_threads, tile_smem, st, pot->dP.p, vw, ac);
        pot->launches += 2;
      }
      return 0;
    };
    CU_TRY(pot, cudaEventRecord(pot->ev[0], pot->stream));
#ifndef EONB200_EMU
    if (pipe && pipe->chunks > 1 && sync_and_check) {
      const int nc = std::min(pipe->chunks, kPipeMax);
      // Without copies back the read-back stream is not touched (measured: the second KERNEL stream has to stay -- with
      // every chunk on one stream the call takes 0.207 ms instead of 0.171, the chunks' tails no longer overlap).
      const bool lean = pipe->no_out;
      CU_TRY(pot, cudaEventRecord(pot->ev_m, pot->stream));  // the counters are cleared
      CU_TRY(pot, cudaStreamWaitEvent(pot->s_k2, pot->ev_m, 0));
      for (int c = 0; c < nc; c++) {
        const long s0 = S * c / nc, s1 = S * (c + 1) / nc;
        if (s1 <= s0) continue;
        int rcp = pipe->in(s0, s1, pot->s_in);
        if (rcp) return rcp;
        CU_TRY(pot, cudaEventRecord(pot->ev_in[c], pot->s_in));
        cudaStream_t ks = (c & 1) ? pot->s_k2 : pot->stream;
        CU_TRY(pot, cudaStreamWaitEvent(ks, pot->ev_in[c], 0));
        NlBatch vw = view;
        offset_systems(vw, static_cast<int>(N), s0);
        ForceBatchArgs ac = a;
        ac.S = static_cast<int>(s1 - s0);
        ac.R += static_cast<size_t>(s0) * a.strideR;
        ac.Z += static_cast<size_t>(s0) * a.strideZ;
        ac.box += static_cast<size_t>(s0) * a.strideBox;
        ac.F += static_cast<size_t>(s0) * a.strideR;
        ac.U += s0;
        if (ac.clocks) ac.clocks += static_cast<size_t>(s0) * 16 * std::max(1, a.tiles);
        rcp = launch(vw, ac, s1 - s0, ks);
        if (rcp) return rcp;
        CU_TRY(pot, cudaEventRecord(pot->ev_k[c], ks));
        if (lean) {
          if (c & 1) CU_TRY(pot, cudaStreamWaitEvent(pot->stream, pot->ev_k[c], 0));
          continue;
        }
        CU_TRY(pot, cudaStreamWaitEvent(pot->s_out, pot->ev_k[c], 0));
        rcp = pipe->out(s0, s1, pot->s_out);
        if (rcp) return rcp;
        CU_TRY(pot, cudaEventRecord(pot->ev_out[c], pot->s_out));
        if (c & 1) CU_TRY(pot, cudaStreamWaitEvent(pot->stream, pot->ev_k[c], 0));  // the counters are read on pot->stream
      }
    } else
#endif
    {
      const int rcl = launch(view, a, S, pot->stream);
      if (rcl) return rcl;
    }
    CU_TRY(pot, cudaEventRecord(pot->ev[2], pot->stream));
    CU_TRY(pot, cudaGetLastError());
    pot->force_calls += static_cast<unsigned long long>(S);
    if (!sync_and_check) return 0;
    if (enqueue_readback && !(pipe && pipe->chunks > 1)) {  // results follow the kernel without a host round trip; a truncated list re-runs everything
      const int rb = enqueue_readback();
      if (rb) return rb;
    }
    CU_TRY(pot, cudaMemcpyAsync(pot->h_stats.p, pot->d_stats.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                pot->stream));  // (stats and d_ovf)
    CU_TRY(pot, cudaStreamSynchronize(pot->stream));
    float ms0 = 0, ms1 = 0;
    if (eam) {  // one kernel: no split point
      cudaEventElapsedTime(&ms1, pot->ev[0], pot->ev[2]);
    } else {
      cudaEventElapsedTime(&ms0, pot->ev[0], pot->ev[1]);
      cudaEventElapsedTime(&ms1, pot->ev[1], pot->ev[2]);
    }
    pot->last_ms[0] = ms0;
    pot->last_ms[1] = ms1;
    const int need = pot->h_ovf[0];
    pot->last_stats[0] = static_cast<long long>(pot->h_stats.p[0]);
    pot->last_stats[1] = eam ? S - static_cast<long long>(pot->h_stats.p[0]) : static_cast<long long>(pot->h_stats.p[1]);
    pot->last_stats[2] = nl.K;
    pot->last_stats[3] = need;
    pot->last_stats[4] = static_cast<long long>(pot->h_stats.p[2]);
    if (n3) nl.slots_seen = std::max(nl.slots_seen, static_cast<int>(pot->h_stats.p[3]));
    if (need <= nl.K) {
      nl.rowmax_seen = std::max(nl.rowmax_seen, need);
      return 0;
    }
    if (need >= (1 << 30))
      return fail(pot, 100, "neighbour image shift outside the representable range (box much smaller than the cutoff)");
    pot->force_calls -= static_cast<unsigned long long>(S);
    const int K = (static_cast<int>(need * 1.2) + 8 + 7) / 8 * 8;
    nl.K = 0;
    rc = nl_alloc(pot, nl, S, N, K, true);
    if (rc) return rc;
    a.force_rebuild = 1;
  }
  return fail(pot, 100, "neighbour list capacity did not converge");
}

}  // namespace

extern "C" {

void eonb200_config_default(EonB200Config *cfg, int kind) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->kind = kind;
  cfg->lj_u0 = 1.0;
  cfg->lj_psi = 1.0;
  cfg->morse_De = 0.7102;
  cfg->morse_a = 1.6047;
  cfg->morse_re = 2.8970;
  if (kind == EONB200_POT_LJ || kind == EONB200_POT_LJCLUSTER) cfg->cutoff = 15.0;
  if (kind == EONB200_POT_MORSE_PT) cfg->cutoff = 9.5;
  cfg->skin = 1.0;
}

int eonb200_abi_version(void) { return EONB200_ABI_VERSION; }

EonB200Pot *eonb200_create(const EonB200Config *cfg_in, char *errbuf, size_t errlen) {
  auto bail = [&](const std::string &msg) -> EonB200Pot * {
    if (errbuf && errlen > 0) {
      strncpy(errbuf, msg.c_str(), errlen - 1);
      errbuf[errlen - 1] = 0;
    }
    return nullptr;
  };
  if (!cfg_in) return bail("eonb200_create: null config");
  EonB200Config cfg = *cfg_in;
  if (cfg.kind < EONB200_POT_LJ || cfg.kind > EONB200_POT_EAM) return bail("eonb200_create: unknown potential kind");
  EonB200Config def;
  eonb200_config_default(&def, cfg.kind);
  if (cfg.cutoff <= 0) cfg.cutoff = def.cutoff;
  if (cfg.lj_u0 <= 0) cfg.lj_u0 = def.lj_u0;
  if (cfg.lj_psi <= 0) cfg.lj_psi = def.lj_psi;
  if (cfg.morse_De == 0 && cfg.morse_a == 0 && cfg.morse_re == 0) {
    cfg.morse_De = def.morse_De;
    cfg.morse_a = def.morse_a;
    cfg.morse_re = def.morse_re;
  }
  if (cfg.skin < 0) cfg.skin = 1.0;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return bail(std::string("eonb200_create: no CUDA device (") + cudaGetErrorString(e) +
                "); this backend has no CPU fallback");
  if (cfg.device < 0 || cfg.device >= ndev) return bail("eonb200_create: device ordinal out of range");
  if ((e = cudaSetDevice(cfg.device)) != cudaSuccess) return bail(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, cfg.device)) != cudaSuccess)
    return bail(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major < 10) return bail("eonb200_create: device is not sm_100 class; libeonb200 is built for sm_100a only");

  EonB200Pot *pot = new EonB200Pot();
  pot->cfg = cfg;
  pot->device = cfg.device;
  if (cfg.kind == EONB200_POT_EAM) {
    if (cfg.n_eam_rows > 0 && cfg.eam_rows) {
      if (cfg.n_eam_rows > kMaxEamRows) {
        delete pot;
        return bail("eonb200_create: too many EAM rows");
      }
      pot->rows.assign(cfg.eam_rows, cfg.eam_rows + cfg.n_eam_rows);
    } else {
      pot->rows.assign(kBuiltinRows, kBuiltinRows + 3);
    }
  }
  pot->cfg.eam_rows = nullptr;
  pot->sm_count = prop.multiProcessorCount;
  fill_params(pot);
  {
    // 2^(j/128) rounded from long double: the table of exp_tab (potentials.cuh)
    double tab[128];
    for (int j = 0; j < 128; j++) tab[j] = static_cast<double>(exp2l(static_cast<long double>(j) / 128.0L));
    if (pot->d_exp_tab.reserve(128) != cudaSuccess ||
        cudaMemcpy(pot->d_exp_tab.p, tab, sizeof(tab), cudaMemcpyHostToDevice) != cudaSuccess) {
      delete pot;
      return bail("eonb200_create: cannot allocate device memory");
    }
    pot->hp.exp_tab = pot->d_exp_tab.p;
    for (int j = 0; j < 16; j++) pot->hp.exp_t16[j] = static_cast<double>(exp2l(static_cast<long double>(j) / 16.0L));
  }
  bool ok = cudaStreamCreate(&pot->stream) == cudaSuccess;
  for (int k = 0; k < 3 && ok; k++) ok = cudaEventCreate(&pot->ev[k]) == cudaSuccess;
#ifndef EONB200_EMU
  ok = ok && cudaStreamCreateWithFlags(&pot->s_in, cudaStreamNonBlocking) == cudaSuccess &&
       cudaStreamCreateWithFlags(&pot->s_k2, cudaStreamNonBlocking) == cudaSuccess &&
       cudaStreamCreateWithFlags(&pot->s_out, cudaStreamNonBlocking) == cudaSuccess &&
       cudaEventCreateWithFlags(&pot->ev_m, cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < kPipeMax && ok; k++)
    ok = cudaEventCreateWithFlags(&pot->ev_in[k], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&pot->ev_k[k], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&pot->ev_out[k], cudaEventDisableTiming) == cudaSuccess;
#endif
  ok = ok && pot->dP.reserve(1) == cudaSuccess && pot->d_stats.reserve(5) == cudaSuccess &&
       pot->h_stats.reserve(5) == cudaSuccess;
  if (ok) {
    pot->d_ovf = reinterpret_cast<int *>(pot->d_stats.p + 4);
    pot->h_ovf = reinterpret_cast<int *>(pot->h_stats.p + 4);
  }
  ok = ok && cudaMemcpy(pot->dP.p, &pot->hp, sizeof(PotParams), cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok) {
    // opt in to the full shared memory of the SM for the walker kernels
    cudaFuncAttributes fa;
    int limit = static_cast<int>(prop.sharedMemPerBlockOptin);
    int worst_static = 0;
    if (cudaFuncGetAttributes(&fa, reinterpret_cast<const void *>(k_walkers)) == cudaSuccess)
      worst_static = std::max<int>(worst_static, static_cast<int>(fa.sharedSizeBytes));
    if (cudaFuncGetAttributes(&fa, reinterpret_cast<const void *>(k_prepare)) == cudaSuccess)
      worst_static = std::max<int>(worst_static, static_cast<int>(fa.sharedSizeBytes));
    pot->max_dyn_smem = limit - worst_static - 1024;
    ok = cudaFuncSetAttribute(reinterpret_cast<const void *>(k_walkers), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
#ifndef EONB200_EMU
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_walkers_x2), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
#endif
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_prepare), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
#ifndef EONB200_EMU
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_fused_t<128, 7>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_fused_t<224, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_fused_t<448, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_fused_t<512, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_fused_t<512, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_n3_t<224, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_n3_t<224, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_n3_t<448, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_n3_t<512, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_n3_t<512, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_n3_t<416, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_n3_t<800, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_eam_n3_t<1024, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
#endif
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_pair), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess &&
         cudaFuncSetAttribute(reinterpret_cast<const void *>(k_neighbors), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              pot->max_dyn_smem) == cudaSuccess;
  }
  if (!ok) {
    std::string msg = std::string("eonb200_create: CUDA setup failed: ") + cudaGetErrorString(cudaGetLastError());
    eonb200_destroy(pot);
    return bail(msg);
  }
  return pot;
}

void eonb200_destroy(EonB200Pot *pot) {
  if (!pot) return;
  cudaSetDevice(pot->device);
  if (pot->stream) cudaStreamSynchronize(pot->stream);
#ifndef EONB200_EMU
  for (auto &r : pot->pinned) cudaHostUnregister(const_cast<char *>(r.first));
#endif
  for (int k = 0; k < 3; k++)
    if (pot->ev[k]) cudaEventDestroy(pot->ev[k]);
#ifndef EONB200_EMU
  for (int k = 0; k < kPipeMax; k++) {
    if (pot->ev_in[k]) cudaEventDestroy(pot->ev_in[k]);
    if (pot->ev_k[k]) cudaEventDestroy(pot->ev_k[k]);
    if (pot->ev_out[k]) cudaEventDestroy(pot->ev_out[k]);
  }
  if (pot->ev_m) cudaEventDestroy(pot->ev_m);
  if (pot->s_in) cudaStreamDestroy(pot->s_in);
  if (pot->s_k2) cudaStreamDestroy(pot->s_k2);
  if (pot->s_out) cudaStreamDestroy(pot->s_out);
#endif
  if (pot->stream) cudaStreamDestroy(pot->stream);
  delete pot;
}

const char *eonb200_last_error(const EonB200Pot *pot) { return pot ? pot->err.c_str() : "null handle"; }
unsigned long long eonb200_force_calls(const EonB200Pot *pot) { return pot ? pot->force_calls : 0; }
unsigned long long eonb200_kernel_launches(const EonB200Pot *pot) { return pot ? pot->launches : 0; }

int eonb200_nl_stats(EonB200Pot *pot, long long out[5]) {
  if (!pot) return 1;
  for (int k = 0; k < 5; k++) out[k] = pot->last_stats[k];
  return 0;
}

int eonb200_l2_flush(EonB200Pot *pot, size_t bytes) {
  if (!pot) return 1;
  CU_TRY(pot, cudaSetDevice(pot->device));
  CU_TRY(pot, pot->d_flush.reserve(bytes));
  CU_TRY(pot, cudaMemsetAsync(pot->d_flush.p, 0, bytes, pot->stream));
  return 0;
}

int eonb200_phase_clocks(EonB200Pot *pot, int enable, long long *out, long cap_ctas, long *n_ctas) {
  if (!pot) return 1;
  pot->clocks_on = enable ? 1 : 0;
  if (n_ctas) *n_ctas = pot->clocks_ctas;
  if (out && pot->d_clocks.p && pot->clocks_ctas > 0) {
    CU_TRY(pot, cudaSetDevice(pot->device));
    const long n = std::min(cap_ctas, pot->clocks_ctas);
    CU_TRY(pot, cudaMemcpy(out, pot->d_clocks.p, sizeof(long long) * n * 16, cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ---- NEB band forces ---------------------------------------------------------------------
static bool host_invert3(const double *m, double *inv) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  if (!(std::fabs(det) > 1e-300)) return false;
  inv[0] = c00 / det;
  inv[1] = (m[2] * m[7] - m[1] * m[8]) / det;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) / det;
  inv[3] = c01 / det;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) / det;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) / det;
  inv[6] = c02 / det;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) / det;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) / det;
  return true;
}

void eonb200_neb_params_default(EonB200NebParams *p) {
  if (!p) return;
  // include/eon/Parameters.h:478-489: spring.constant 5.0, weighting k_min 0.97 / k_max 9.7 (off), improved tangent,
  // plain NEB projection; the climbing image is switched on by the caller once its trigger is met
  p->spring_constant = 5.0;
  p->k_min = 0.97;
  p->k_max = 9.7;
  p->climbing_image = 0;
  p->old_tangent = 0;
  p->projection = EONB200_NEB_PROJECTION_NEB;
  p->energy_weighted = 0;
}

static int neb_launch(EonB200Pot *pot, long nBands, long nImages, long nAtoms, long nFreeAtoms, const double *d_positions,
                      const double *d_forces, const double *d_energies, const double *box, const EonB200NebParams *prm,
                      double *d_projected, double *d_tangents, int *d_climbing, const double *d_mask, int remove_net) {
  if (!pot) return 1;
  if (nBands <= 0 || nImages <= 0 || nAtoms <= 0) return fail(pot, 2, "neb_forces: bands, images and atoms must be positive");
  if (!d_positions || !d_forces || !d_energies || !prm || !d_projected || !d_tangents)
    return fail(pot, 2, "neb_forces: null argument");
  if (prm->projection < 0 || prm->projection > 2) return fail(pot, 2, "neb_forces: unknown projection %d", prm->projection);
  if (nBands * nImages > 2147483647L / 2) return fail(pot, 2, "neb_forces: too many images");
  CU_TRY(pot, cudaSetDevice(pot->device));
  NebArgs a;
  a.pos = d_positions;
  a.frc = d_forces;
  a.energy = d_energies;
  a.projected = d_projected;
  a.tangent = d_tangents;
  a.climbing = d_climbing;
  a.mask = d_mask;
  a.remove_net = remove_net;
  a.cell_ok = 0;
  for (int k = 0; k < 9; k++) a.cell[k] = a.inv[k] = 0.0;
  if (box) {
    for (int k = 0; k < 9; k++) a.cell[k] = box[k];
    a.cell_ok = host_invert3(a.cell, a.inv) ? 1 : 0;
  }
  a.ksp = prm->spring_constant;
  a.k_min = prm->k_min;
  a.k_max = prm->k_max;
  a.images = static_cast<int>(nImages);
  a.atoms = static_cast<int>(nAtoms);
  a.all_free = nFreeAtoms == nAtoms ? 1 : 0;
  a.ci_active = prm->climbing_image ? 1 : 0;
  a.old_tangent = prm->old_tangent ? 1 : 0;
  a.weighted = prm->energy_weighted ? 1 : 0;
  // elastic band / doubly nudged only without energy weighting (client/NEBProjection.cpp:33-45)
  a.projection = a.weighted ? static_cast<int>(NEB_PROJ_NEB) : prm->projection;
  const unsigned grid = static_cast<unsigned>(nBands * nImages);
  bool staged = false;
#ifndef EONB200_EMU  // (the emulation runs the any-size kernel)
  const size_t stage = neb_stage_bytes(static_cast<int>(nAtoms));
  if (stage <= static_cast<size_t>(pot->max_dyn_smem)) {
    if (!pot->neb_attr) {
      CU_TRY(pot, cudaFuncSetAttribute(reinterpret_cast<const void *>(k_neb_band_tma),
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, pot->max_dyn_smem));
      pot->neb_attr = 1;
    }
    const size_t total = static_cast<size_t>(nBands) * (nImages + 2) * 3 * nAtoms;
    EONB200_LAUNCH(k_neb_band_tma, grid, 256, stage, pot->stream, a, total);
    staged = true;
  }
#endif
  if (!staged) EONB200_LAUNCH(k_neb_band, grid, 256, 0, pot->stream, a);
  CU_TRY(pot, cudaGetLastError());
  pot->launches += 1;
  return 0;
}

int eonb200_neb_forces_device(EonB200Pot *pot, long nBands, long nImages, long nAtoms, long nFreeAtoms,
                              const double *d_positions, const double *d_forces, const double *d_energies,
                              const double *box, const EonB200NebParams *prm, double *d_projected, double *d_tangents,
                              int *d_climbing) {
  return neb_launch(pot, nBands, nImages, nAtoms, nFreeAtoms, d_positions, d_forces, d_energies, box, prm, d_projected,
                    d_tangents, d_climbing, nullptr, 0);
}

int eonb200_neb_band_forces(EonB200Pot *pot, long nBands, long nImages, long nAtoms, const double *positions,
                            const int *atomicNrs, const int *fixed, const double *box, int remove_net_force,
                            const EonB200NebParams *prm, double *projected, double *energies, double *tangents,
                            int *climbing) {
  if (!pot) return 1;
  if (nBands <= 0 || nImages <= 0 || nAtoms <= 0) return fail(pot, 2, "neb_band_forces: bands, images and atoms must be positive");
  if (!positions || !atomicNrs || !box || !prm || !projected) return fail(pot, 2, "neb_band_forces: null argument");
  {
    const int rc = check_system(pot, nAtoms, atomicNrs, "neb_band_forces");
    if (rc) return rc;
  }
  CU_TRY(pot, cudaSetDevice(pot->device));
  const long N = nAtoms, S = nBands * (nImages + 2);
  const size_t D = 3 * static_cast<size_t>(N);
  const size_t rows = static_cast<size_t>(S), mov = static_cast<size_t>(nBands) * nImages;
  if (S > 2147483647L / 2) return fail(pot, 2, "neb_band_forces: too many images");
  CU_TRY(pot, pot->neb_R.reserve(rows * D));
  CU_TRY(pot, pot->neb_F.reserve(rows * D));
  CU_TRY(pot, pot->neb_E.reserve(rows));
  CU_TRY(pot, pot->neb_P.reserve(mov * D));
  CU_TRY(pot, pot->neb_T.reserve(mov * D));
  CU_TRY(pot, pot->neb_C.reserve(static_cast<size_t>(nBands)));
  CU_TRY(pot, pot->neb_Z.reserve(static_cast<size_t>(N)));
  CU_TRY(pot, pot->neb_box.reserve(9));
  CU_TRY(pot, pot->neb_mask.reserve(D));
  long n_fixed_atoms = 0;
  bool any_fixed = false;
  if (fixed) {
    pot->neb_hmask.resize(D);
    for (long i = 0; i < N; i++) {  // Matter::getFixed(i): all three axes; the mask itself is per axis
      bool all_fx = true;
      for (int k = 0; k < 3; k++) {
        pot->neb_hmask[3 * i + k] = fixed[3 * i + k] ? 0.0 : 1.0;
        all_fx = all_fx && fixed[3 * i + k] != 0;
        any_fixed = any_fixed || fixed[3 * i + k] != 0;
      }
      n_fixed_atoms += all_fx ? 1 : 0;
    }
  }
  pin_range(pot, positions, rows * D * sizeof(double));
  pin_range(pot, projected, mov * D * sizeof(double));
  CU_TRY(pot, cudaMemcpyAsync(pot->neb_R.p, positions, rows * D * sizeof(double), cudaMemcpyHostToDevice, pot->stream));
  CU_TRY(pot, cudaMemcpyAsync(pot->neb_Z.p, atomicNrs, static_cast<size_t>(N) * sizeof(int), cudaMemcpyHostToDevice, pot->stream));
  CU_TRY(pot, cudaMemcpyAsync(pot->neb_box.p, box, 9 * sizeof(double), cudaMemcpyHostToDevice, pot->stream));
  if (any_fixed)
    CU_TRY(pot, cudaMemcpyAsync(pot->neb_mask.p, pot->neb_hmask.data(), D * sizeof(double), cudaMemcpyHostToDevice, pot->stream));
  ForceBatchArgs a;
  memset(&a, 0, sizeof(a));
  a.S = static_cast<int>(S);
  a.N = static_cast<int>(N);
  a.R = pot->neb_R.p;
  a.strideR = static_cast<int>(D);
  a.Z = pot->neb_Z.p;
  a.strideZ = 0;  // one atomicNrs row and one cell shared by every image
  a.box = pot->neb_box.p;
  a.strideBox = 0;
  a.F = pot->neb_F.p;
  a.U = pot->neb_E.p;
  a.force_rebuild = 0;
  // endpoints ride along in the same launch (2 of nImages + 2 rows): positions, forces and energies keep the
  // [band][image] layout the band kernel reads, and the endpoint energies come for free
  auto tail = [&]() -> int {
    const int rc = neb_launch(pot, nBands, nImages, N, N - n_fixed_atoms, pot->neb_R.p, pot->neb_F.p, pot->neb_E.p, box, prm,
                              pot->neb_P.p, pot->neb_T.p, pot->neb_C.p, any_fixed ? pot->neb_mask.p : nullptr,
                              (!any_fixed && remove_net_force) ? 1 : 0);
    if (rc) return rc;
    CU_TRY(pot, cudaMemcpyAsync(projected, pot->neb_P.p, mov * D * sizeof(double), cudaMemcpyDeviceToHost, pot->stream));
    if (energies) CU_TRY(pot, cudaMemcpyAsync(energies, pot->neb_E.p, rows * sizeof(double), cudaMemcpyDeviceToHost, pot->stream));
    if (tangents)
      CU_TRY(pot, cudaMemcpyAsync(tangents, pot->neb_T.p, mov * D * sizeof(double), cudaMemcpyDeviceToHost, pot->stream));
    if (climbing)
      CU_TRY(pot, cudaMemcpyAsync(climbing, pot->neb_C.p, static_cast<size_t>(nBands) * sizeof(int), cudaMemcpyDeviceToHost,
                                  pot->stream));
    return 0;
  };
  return run_force_batch(pot, pot->nl, a, box, true, tail);
}

int eonb200_neb_forces(EonB200Pot *pot, long nBands, long nImages, long nAtoms, long nFreeAtoms, const double *positions,
                       const double *forces, const double *energies, const double *box, const EonB200NebParams *prm,
                       double *projected, double *tangents, int *climbing) {
  if (!pot) return 1;
  if (nBands <= 0 || nImages <= 0 || nAtoms <= 0) return fail(pot, 2, "neb_forces: bands, images and atoms must be positive");
  if (!positions || !forces || !energies || !prm || !projected) return fail(pot, 2, "neb_forces: null argument");
  CU_TRY(pot, cudaSetDevice(pot->device));
  const size_t D = 3 * static_cast<size_t>(nAtoms);
  const size_t rows = static_cast<size_t>(nBands) * (nImages + 2), mov = static_cast<size_t>(nBands) * nImages;
  CU_TRY(pot, pot->neb_R.reserve(rows * D));
  CU_TRY(pot, pot->neb_F.reserve(rows * D));
  CU_TRY(pot, pot->neb_E.reserve(rows));
  CU_TRY(pot, pot->neb_P.reserve(mov * D));
  CU_TRY(pot, pot->neb_T.reserve(mov * D));
  CU_TRY(pot, pot->neb_C.reserve(static_cast<size_t>(nBands)));
  CU_TRY(pot, cudaMemcpyAsync(pot->neb_R.p, positions, rows * D * sizeof(double), cudaMemcpyHostToDevice, pot->stream));
  CU_TRY(pot, cudaMemcpyAsync(pot->neb_F.p, forces, rows * D * sizeof(double), cudaMemcpyHostToDevice, pot->stream));
  CU_TRY(pot, cudaMemcpyAsync(pot->neb_E.p, energies, rows * sizeof(double), cudaMemcpyHostToDevice, pot->stream));
  const int rc = eonb200_neb_forces_device(pot, nBands, nImages, nAtoms, nFreeAtoms, pot->neb_R.p, pot->neb_F.p,
                                           pot->neb_E.p, box, prm, pot->neb_P.p, pot->neb_T.p, pot->neb_C.p);
  if (rc) return rc;
  CU_TRY(pot, cudaMemcpyAsync(projected, pot->neb_P.p, mov * D * sizeof(double), cudaMemcpyDeviceToHost, pot->stream));
  if (tangents)
    CU_TRY(pot, cudaMemcpyAsync(tangents, pot->neb_T.p, mov * D * sizeof(double), cudaMemcpyDeviceToHost, pot->stream));
  if (climbing)
    CU_TRY(pot, cudaMemcpyAsync(climbing, pot->neb_C.p, static_cast<size_t>(nBands) * sizeof(int), cudaMemcpyDeviceToHost,
                                pot->stream));
  CU_TRY(pot, cudaStreamSynchronize(pot->stream));
  return 0;
}

int eonb200_last_force_ms(EonB200Pot *pot, double out[2]) {
  if (!pot) return 1;
  out[0] = pot->last_ms[0];
  out[1] = pot->last_ms[1];
  return 0;
}

int eonb200_force_batch(EonB200Pot *pot, long S, long N, const double *const *positions, const int *const *atomicNrs,
                        double *const *forces, double *energies, double *variances, const double *const *boxes) {
  if (!pot) return 1;
  static const bool trace = getenv("EONB200_TRACE") != nullptr;  // development aid: host-side phase times on stderr
  const auto t_begin = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (trace)
      fprintf(stderr, "[force_batch] %-12s %8.1f us\n", what,
              std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count());
  };
  if (S <= 0 || N <= 0) return fail(pot, 2, "force_batch: nSystems and nAtoms must be positive");
  CU_TRY(pot, cudaSetDevice(pot->device));
  // one topology / one box shared by all systems (the same pointer in every slot): keep a single copy
  bool same_Z = true, same_box = true;
  for (long s = 1; s < S; s++) {
    same_Z = same_Z && atomicNrs[s] == atomicNrs[0];
    same_box = same_box && boxes[s] == boxes[0];
  }
  const long SZ = same_Z ? 1 : S, SB = same_box ? 1 : S;
  if (N > kNbrIndexMask) return fail(pot, 2, "force_batch: more than %d atoms per system", kNbrIndexMask);
  lap("checked");
  const size_t D = 3 * static_cast<size_t>(N);
  const bool direct_in = rows_contiguous(positions, S, D) && pin_range(pot, positions[0], S * D * sizeof(double));
  const bool direct_out = rows_contiguous(forces, S, D) && pin_range(pot, forces[0], S * D * sizeof(double));
  // device rows: the caller's own pitch when both blocks are copied straight from / to page-locked memory (one linear
  // DMA per chunk instead of one descriptor per 9 KB row; the kernels stage rows of any 8-byte alignment), otherwise
  // an even pitch so that every system starts 16-byte aligned
  const size_t strideR = (direct_in && direct_out) ? D : (D + 1) / 2 * 2;
  // Zero copy: the half-list EAM kernel stages a walker's coordinates with one bulk copy and returns its forces with
  // one bulk store, so it can work on the caller's page-locked blocks directly -- every CTA fetches its own 9 KB over
  // PCIe when it starts and writes its row back when it ends, transfers and arithmetic of different walkers overlap
  // by themselves, and the call is one launch instead of a chunked copy / kernel / copy pipeline driven by ~20 API
  // calls.  (Other kernels write forces atom by atom and keep the staged path.)
  const double *zc_R = nullptr;
  double *zc_F = nullptr;
#ifndef EONB200_EMU
  {
    static const int zc_env = getenv("EONB200_ZERO_COPY") ? atoi(getenv("EONB200_ZERO_COPY")) : 1;  // development aid: 0 off
    if (zc_env && direct_in && direct_out && use_n3(pot, N) && pot->nl.n3 && pot->nl.N == N && pot->nl.S >= S &&
        pick_n3_lpa(pot, N) * N <= 1024) {
      void *pr = nullptr, *pf = nullptr;
      if (cudaHostGetDevicePointer(&pr, const_cast<double *>(positions[0]), 0) == cudaSuccess &&
          cudaHostGetDevicePointer(&pf, forces[0], 0) == cudaSuccess) {
        zc_R = static_cast<const double *>(pr);
        zc_F = static_cast<double *>(pf);
      } else {
        cudaGetLastError();
      }
    }
  }
#endif
  static const int zc_mode = getenv("EONB200_ZERO_COPY") ? atoi(getenv("EONB200_ZERO_COPY")) : 1;  // 2: coordinates too
  // Coordinates fetched by the CTAs over PCIe arrive at about half the copy engine's rate (measured on B200: 256 x
  // 385 atoms 0.174 ms per call against 0.135 ms with copied coordinates), so by default only the forces take the
  // direct way: posted writes, spread over the kernel's life, no read-back call at all.
  // Systems scattered over the caller's heap (one block per system, what eonclient's Matter objects hand over): the
  // span that holds them is page-locked once and the kernel gets one device-visible address per row -- both ways
  // direct, because the alternative is a 2 x 2.4 MB gather / scatter memcpy on the calling thread.
  bool scattered = false;
#ifndef EONB200_EMU
  if (zc_mode >= 1 && !zc_F && !(direct_in && direct_out) && use_n3(pot, N) && pot->nl.n3 && pot->nl.N == N &&
      pot->nl.S >= S && pick_n3_lpa(pot, N) * N <= 1024 && pin_rows_span(pot, positions, forces, S, D)) {
    const char *hb = nullptr;
    size_t hn = 0;
    for (auto &r : pot->pinned)
      if (reinterpret_cast<const char *>(positions[0]) >= r.first &&
          reinterpret_cast<const char *>(positions[0]) < r.first + r.second) {
        hb = r.first;
        hn = r.second;
      }
    void *db = nullptr;
    if (pot->d_rowptr.n < static_cast<size_t>(2 * S)) pot->rowptr_sent.clear();  // (regrown: contents lost)
    if (hb && cudaHostGetDevicePointer(&db, const_cast<char *>(hb), 0) == cudaSuccess &&
        pot->h_rowptr.reserve(2 * S) == cudaSuccess && pot->d_rowptr.reserve(2 * S) == cudaSuccess) {
      bool inside = true;
      for (long s = 0; s < S && inside; s++) {
        const char *pr = reinterpret_cast<const char *>(positions[s]), *pf = reinterpret_cast<const char *>(forces[s]);
        inside = pr >= hb && pr + D * sizeof(double) <= hb + hn && pf >= hb && pf + D * sizeof(double) <= hb + hn;
        pot->h_rowptr.p[s] = reinterpret_cast<unsigned long long>(static_cast<char *>(db) + (pr - hb));
        pot->h_rowptr.p[S + s] = reinterpret_cast<unsigned long long>(static_cast<char *>(db) + (pf - hb));
      }
      if (inside) {
        if (pot->rowptr_sent.size() != static_cast<size_t>(2 * S) ||
            memcmp(pot->rowptr_sent.data(), pot->h_rowptr.p, 2 * S * sizeof(unsigned long long)) != 0) {
          CU_TRY(pot, cudaMemcpyAsync(pot->d_rowptr.p, pot->h_rowptr.p, 2 * S * sizeof(unsigned long long),
                                      cudaMemcpyHostToDevice, pot->stream));
          pot->rowptr_sent.assign(pot->h_rowptr.p, pot->h_rowptr.p + 2 * S);
        }
        scattered = true;
      }
    } else {
      cudaGetLastError();
    }
  }
#endif
  const bool zero_copy = (zc_R != nullptr && zc_mode >= 2) || scattered;
  const bool zero_copy_out = zc_F != nullptr || scattered;
  double *zc_U = nullptr;  // energies likewise (set below, once the handle's page-locked block is in place)
  CU_TRY(pot, pot->dR.reserve(S * strideR));
  CU_TRY(pot, pot->dF.reserve(S * strideR));
  CU_TRY(pot, pot->dU.reserve(S));
  if (pot->dZ.n < static_cast<size_t>(S * N) || pot->hZ.n < static_cast<size_t>(S * N)) pot->z_cached_S = 0;  // regrown: contents lost
  if (pot->dbox.n < static_cast<size_t>(S * 9)) pot->last_boxes.clear();
  CU_TRY(pot, pot->dbox.reserve(S * 9));
  CU_TRY(pot, pot->dZ.reserve(S * N));
  CU_TRY(pot, pot->hR.reserve(S * strideR));
  CU_TRY(pot, pot->hF.reserve(S * strideR));
  CU_TRY(pot, pot->hU.reserve(S));
#ifndef EONB200_EMU
  if (zero_copy_out) {  // one 8-byte posted write per walker
    void *pu = nullptr;
    if (cudaHostGetDevicePointer(&pu, pot->hU.p, 0) == cudaSuccess)
      zc_U = static_cast<double *>(pu);
    else
      cudaGetLastError();
  }
#endif
  CU_TRY(pot, pot->hbox.reserve(S * 9));
  CU_TRY(pot, pot->hZ.reserve(S * N));
  bool box_changed = pot->last_boxes.size() != static_cast<size_t>(SB) * 9;
  // atomic numbers: checked (EAM.cpp:116-123) and sent only when they differ from the previous call's
  bool z_changed = pot->z_cached_S != SZ || pot->z_cached_N != N;
  for (long s = 0; s < SZ && !z_changed; s++)
    z_changed = memcmp(pot->hZ.p + s * N, atomicNrs[s], N * sizeof(int)) != 0;
  if (z_changed) {
    pot->z_cached_S = pot->z_cached_N = 0;
    for (long s = 0; s < SZ; s++) {
      const int rc = check_system(pot, N, atomicNrs[s], "force_batch");
      if (rc) return rc;
      memcpy(pot->hZ.p + s * N, atomicNrs[s], N * sizeof(int));
    }
  }
  for (long s = 0; s < SB; s++) memcpy(pot->hbox.p + s * 9, boxes[s], 9 * sizeof(double));
  if (!box_changed) box_changed = memcmp(pot->last_boxes.data(), pot->hbox.p, SB * 9 * sizeof(double)) != 0;
  if (box_changed) pot->last_boxes.assign(pot->hbox.p, pot->hbox.p + SB * 9);
  if (z_changed) {
    CU_TRY(pot, cudaMemcpyAsync(pot->dZ.p, pot->hZ.p, SZ * N * sizeof(int), cudaMemcpyHostToDevice, pot->stream));
    pot->z_cached_S = SZ;
    pot->z_cached_N = N;
  }
  if (box_changed)
    CU_TRY(pot, cudaMemcpyAsync(pot->dbox.p, pot->hbox.p, SB * 9 * sizeof(double), cudaMemcpyHostToDevice, pot->stream));
  // positions of systems [s0, s1) up / forces and energies down, on the given stream
  auto copy_in = [&](long s0, long s1, cudaStream_t st) -> int {
    const size_t n = static_cast<size_t>(s1 - s0);
    if (direct_in && strideR == D) {  // straight from the caller's page-locked block
      CU_TRY(pot, cudaMemcpyAsync(pot->dR.p + s0 * D, positions[s0], n * D * sizeof(double), cudaMemcpyHostToDevice, st));
    } else if (direct_in) {  // ... re-pitched to the 16-byte aligned device rows
      CU_TRY(pot, cudaMemcpy2DAsync(pot->dR.p + s0 * strideR, strideR * sizeof(double), positions[s0], D * sizeof(double),
                                    D * sizeof(double), n, cudaMemcpyHostToDevice, st));
    } else {
      for (long s = s0; s < s1; s++) {
        memcpy(pot->hR.p + s * strideR, positions[s], D * sizeof(double));
        if (strideR > D) pot->hR.p[s * strideR + D] = 0.0;
      }
      CU_TRY(pot, cudaMemcpyAsync(pot->dR.p + s0 * strideR, pot->hR.p + s0 * strideR, n * strideR * sizeof(double),
                                  cudaMemcpyHostToDevice, st));
    }
    return 0;
  };
  auto copy_out = [&](long s0, long s1, cudaStream_t st) -> int {
    const size_t n = static_cast<size_t>(s1 - s0);
    if (zero_copy_out) {  // the kernel wrote the caller's block itself
      if (zc_U) return 0;
    } else if (direct_out && strideR == D)
      CU_TRY(pot, cudaMemcpyAsync(forces[s0], pot->dF.p + s0 * D, n * D * sizeof(double), cudaMemcpyDeviceToHost, st));
    else if (direct_out)
      CU_TRY(pot, cudaMemcpy2DAsync(forces[s0], D * sizeof(double), pot->dF.p + s0 * strideR, strideR * sizeof(double),
                                    D * sizeof(double), n, cudaMemcpyDeviceToHost, st));
    else
      CU_TRY(pot, cudaMemcpyAsync(pot->hF.p + s0 * strideR, pot->dF.p + s0 * strideR, n * strideR * sizeof(double),
                                  cudaMemcpyDeviceToHost, st));
    CU_TRY(pot, cudaMemcpyAsync(pot->hU.p + s0, pot->dU.p + s0, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    return 0;
  };
  // Large batches go through in chunks: the copy of chunk c + 1 overlaps the kernel of chunk c and the read-back of
  // chunk c - 1 (the serial H2D -> kernel -> D2H of one big launch left the PCIe link idle two thirds of the call).
  BatchPipe pipe;
  pipe.chunks = 1;
#ifndef EONB200_EMU
  {
    static const int forced = getenv("EONB200_PIPE_CHUNKS") ? atoi(getenv("EONB200_PIPE_CHUNKS")) : 0;  // development aid
    const size_t bytes = static_cast<size_t>(S) * D * sizeof(double);
    if (zero_copy)
      pipe.chunks = 1;
    else if (forced > 0)
      pipe.chunks = std::min<long>(std::min(forced, kPipeMax), S);
    else if (S >= 64 && bytes >= (512u << 10))
      pipe.chunks = static_cast<int>(std::min<long>(kPipeMax, std::max<long>(2, S / 128)));  // (C2, 256 x 385 atoms,
    // measured on B200 through the Python plugin call: 1 chunk 0.241 ms, 2 chunks 0.218, 4 chunks 0.233, 8 chunks 0.32)
  }
#endif
  pipe.in = copy_in;
  pipe.out = copy_out;
  pipe.no_out = zero_copy_out && zc_U != nullptr;
  if (pipe.chunks == 1 && !zero_copy) {
    const int rc0 = copy_in(0, S, pot->stream);
    if (rc0) return rc0;
  }
  ForceBatchArgs a;
  memset(&a, 0, sizeof(a));
  a.S = static_cast<int>(S);
  a.N = static_cast<int>(N);
  a.R = (zero_copy && !scattered) ? zc_R : pot->dR.p;
  a.strideR = static_cast<int>(strideR);
  a.Z = pot->dZ.p;
  a.strideZ = same_Z ? 0 : static_cast<int>(N);
  a.box = pot->dbox.p;
  a.strideBox = same_box ? 0 : 9;
  a.F = (zero_copy_out && !scattered) ? zc_F : pot->dF.p;
  if (scattered) {
    a.Rp = reinterpret_cast<const double *const *>(pot->d_rowptr.p);
    a.Fp = reinterpret_cast<double *const *>(pot->d_rowptr.p + S);
    a.strideR = static_cast<int>(D);
  }
  a.U = zc_U ? zc_U : pot->dU.p;
  a.force_rebuild = box_changed ? 1 : 0;
  auto readback = [&]() -> int { return copy_out(0, S, pot->stream); };
  lap("staged");
  int rc = run_force_batch(pot, pot->nl, a, pot->hbox.p, true, readback, &pipe);
  if (rc) return rc;
  lap("ran");
  const int nc = pipe.chunks;
  for (int c = 0; c < nc; c++) {
    const long s0 = S * c / nc, s1 = S * (c + 1) / nc;
    if (s1 <= s0) continue;
#ifndef EONB200_EMU
    if (nc > 1 && !pipe.no_out) CU_TRY(pot, cudaEventSynchronize(pot->ev_out[c]));  // chunk c is back; later chunks still in flight
#endif
    for (long s = s0; s < s1; s++) {
      if (!direct_out && !scattered) memcpy(forces[s], pot->hF.p + s * strideR, D * sizeof(double));
      energies[s] = pot->hU.p[s];
      if (variances) variances[s] = 0.0;
    }
  }
  lap("unpacked");
  return 0;
}

int eonb200_force(EonB200Pot *pot, long N, const double *positions, const int *atomicNrs, double *forces, double *energy,
                  double *variance, const double *box) {
  const double *R[1] = {positions};
  const int *Z[1] = {atomicNrs};
  double *F[1] = {forces};
  const double *B[1] = {box};
  double var = 0.0;
  int rc = eonb200_force_batch(pot, 1, N, R, Z, F, energy, &var, B);
  if (variance) *variance = var;
  return rc;
}

int eonb200_force_batch_device(EonB200Pot *pot, long S, long N, const double *d_positions, const int *d_atomicNrs,
                               double *d_forces, double *d_energies, const double *d_boxes, int sync) {
  if (!pot) return 1;
  if (S <= 0 || N <= 0) return fail(pot, 2, "force_batch_device: nSystems and nAtoms must be positive");
  {
    const int rc = check_system(pot, N, nullptr, "force_batch_device");  // (atomic numbers live on the device: an
    if (rc) return rc;                                                    // unknown element yields NaN energies)
  }
  CU_TRY(pot, cudaSetDevice(pot->device));
  ForceBatchArgs a;
  memset(&a, 0, sizeof(a));
  a.S = static_cast<int>(S);
  a.N = static_cast<int>(N);
  a.R = d_positions;
  a.strideR = static_cast<int>(3 * N);
  a.Z = d_atomicNrs;
  a.strideZ = static_cast<int>(N);
  a.box = d_boxes;
  a.strideBox = 9;
  a.F = d_forces;
  a.U = d_energies;
  a.force_rebuild = 0;
  double hb[9];
  const double *hb0 = nullptr;
  if (pot->nl.S < S || pot->nl.N != N) {
    CU_TRY(pot, cudaMemcpy(hb, d_boxes, sizeof(hb), cudaMemcpyDeviceToHost));
    hb0 = hb;
  }
  return run_force_batch(pot, pot->nl, a, hb0, sync != 0);
}

// ---- neighbour lists ---------------------------------------------------------------------------------------
int eonb200_neighbors(EonB200Pot *pot, long S, long N, const double *positions, const double *boxes,
                      const int periodic[3], double cutoff, int full, long cap, long *pair_offsets, int64_t *pairs,
                      int32_t *shifts, double *distances, double *vectors, long *n_pairs) {
  if (!pot) return 1;
  if (S <= 0 || N <= 0 || !(cutoff > 0)) return fail(pot, 2, "neighbors: bad arguments");
  CU_TRY(pot, cudaSetDevice(pot->device));
  DevBuf<double> dR, dB;
  CU_TRY(pot, dR.reserve(S * 3 * N));
  CU_TRY(pot, dB.reserve(S * 9));
  CU_TRY(pot, cudaMemcpyAsync(dR.p, positions, S * 3 * N * sizeof(double), cudaMemcpyHostToDevice, pot->stream));
  CU_TRY(pot, cudaMemcpyAsync(dB.p, boxes, S * 9 * sizeof(double), cudaMemcpyHostToDevice, pot->stream));
  NlParams np;
  np.mode = NL_VESIN;
  np.rc = cutoff;
  np.rc2 = cutoff * cutoff;
  np.rc_core2 = np.rc2;
  np.rc_core = cutoff;
  np.dstar = 0.0;
  np.skin = 0.0;
  for (int k = 0; k < 3; k++) np.per[k] = periodic ? (periodic[k] != 0) : 1;
  NlStore nl;
  int K = guess_K(cutoff, N, boxes);
  const int maxCells = pick_max_cells(N);
  const int threads = pick_threads(pot, N, 1);
  size_t smem;
  int rc = check_smem(pot, N, maxCells, &smem);
  if (rc) return rc;
  for (int attempt = 0;; attempt++) {
    if (attempt >= 6) return fail(pot, 100, "neighbors: list capacity did not converge");
    rc = nl_alloc(pot, nl, S, N, K);
    if (rc) return rc;
    CU_TRY(pot, cudaMemsetAsync(pot->d_ovf, 0, sizeof(int), pot->stream));
    CU_TRY(pot, cudaMemsetAsync(pot->d_stats.p, 0, 4 * sizeof(unsigned long long), pot->stream));
    NlBatch view = nl_view(pot, nl);
    EONB200_LAUNCH(k_neighbors, static_cast<unsigned>(S), threads, smem, pot->stream, view, np, static_cast<int>(S),
                   static_cast<int>(N), dR.p, dB.p, maxCells, pot->d_exp_tab.p);
    pot->launches += 1;
    CU_TRY(pot, cudaGetLastError());
    CU_TRY(pot, cudaMemcpyAsync(pot->h_ovf, pot->d_ovf, sizeof(int), cudaMemcpyDeviceToHost, pot->stream));
    CU_TRY(pot, cudaStreamSynchronize(pot->stream));
    const int need = pot->h_ovf[0];
    if (need >= (1 << 30)) return fail(pot, 100, "neighbors: image shift outside the representable range");
    if (need <= K) break;
    K = (need + 3) / 4 * 4;
    nl.K = 0;
  }
  std::vector<int> hn(static_cast<size_t>(S) * nl.K * nl.Np), hc(static_cast<size_t>(S) * nl.Np);
  CU_TRY(pot, cudaMemcpy(hn.data(), nl.nbr.p, hn.size() * sizeof(int), cudaMemcpyDeviceToHost));
  CU_TRY(pot, cudaMemcpy(hc.data(), nl.cnt.p, hc.size() * sizeof(int), cudaMemcpyDeviceToHost));
  long total = 0;
  for (long s = 0; s < S; s++) {
    if (pair_offsets) pair_offsets[s] = total;
    const double *R = positions + s * 3 * N, *H = boxes + s * 9;
    for (long i = 0; i < N; i++) {
      const int n = hc[s * nl.Np + i];
      for (int k = 0; k < n; k++) {
        int j, sx, sy, sz;
        unpack_nbr(hn[(static_cast<size_t>(s) * nl.K + k) * nl.Np + i], j, sx, sy, sz);
        if (!full) {
          // half-list rule, vesin-single-build.cpp:1165-1196
          if (i > j) continue;
          if (i == j) {
            const int sum = sx + sy + sz;
            if (sum < 0) continue;
            if (sum == 0 && (sz < 0 || (sz == 0 && sy < 0))) continue;
          }
        }
        if (total < cap) {
          if (pairs) {
            pairs[2 * total] = i;
            pairs[2 * total + 1] = j;
          }
          if (shifts) {
            shifts[3 * total] = sx;
            shifts[3 * total + 1] = sy;
            shifts[3 * total + 2] = sz;
          }
          if (distances || vectors) {
            // same un-fused order as the device predicate (host code is compiled without contraction)
            const volatile double tx = (sx * H[0] + sy * H[3]) + sz * H[6];
            const volatile double ty = (sx * H[1] + sy * H[4]) + sz * H[7];
            const volatile double tz = (sx * H[2] + sy * H[5]) + sz * H[8];
            const volatile double vx = (R[3 * j] - R[3 * i]) + tx;
            const volatile double vy = (R[3 * j + 1] - R[3 * i + 1]) + ty;
            const volatile double vz = (R[3 * j + 2] - R[3 * i + 2]) + tz;
            if (vectors) {
              vectors[3 * total] = vx;
              vectors[3 * total + 1] = vy;
              vectors[3 * total + 2] = vz;
            }
            if (distances) {
              const volatile double xx = vx * vx, yy = vy * vy, zz = vz * vz;
              const volatile double s2 = xx + yy;
              distances[total] = std::sqrt(s2 + zz);
            }
          }
        }
        total++;
      }
    }
  }
  if (pair_offsets) pair_offsets[S] = total;
  if (n_pairs) *n_pairs = total;
  return 0;
}

// ---- resident walkers --------------------------------------------------------------------------------------------
void eonb200_search_params_default(EonB200SearchParams *p) {
  memset(p, 0, sizeof(*p));
  p->finite_difference = 0.01;
  p->remove_net_force = 1;
  p->opt_method = EONB200_OPT_CG;
  p->convergence_metric = EONB200_METRIC_NORM;
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
  p->dimer_opt_method = EONB200_ROT_CG;
  p->dimer_converged_angle = 5.0;
  p->dimer_rotations_max = 10;
  p->lanczos_tolerance = 0.01;
  p->lanczos_max_iterations = 20;
  p->lanczos_quit_early = 1;
  p->minmode_method = EONB200_MINMODE_DIMER;
  p->saddle_max_iterations = 1000;
  p->saddle_max_energy = 20.0;
}

EonB200Walkers *eonb200_walkers_create(EonB200Pot *pot, long W, long N, const int *atomicNrs, const int *fixed,
                                       const double *box, const double *positions, const double *modes, char *errbuf,
                                       size_t errlen) {
  auto bail = [&](const std::string &msg) -> EonB200Walkers * {
    if (errbuf && errlen > 0) {
      strncpy(errbuf, msg.c_str(), errlen - 1);
      errbuf[errlen - 1] = 0;
    }
    return nullptr;
  };
  if (!pot || W <= 0 || N <= 0 || !atomicNrs || !box || !positions) return bail("walkers_create: bad arguments");
  if (N > kNbrIndexMask) return bail("walkers_create: too many atoms per walker");
  if (pot->cfg.kind == EONB200_POT_EAM) {
    for (long i = 0; i < N; i++) {
      bool found = false;
      for (const auto &r : pot->rows) found = found || r.Z == atomicNrs[i];
      if (!found) return bail("walkers_create: EAM has no parameter row for an element (14324)");
    }
  }
  cudaSetDevice(pot->device);
  EonB200Walkers *w = new EonB200Walkers();
  w->pot = pot;
  w->W = W;
  w->N = N;
  w->D = static_cast<int>(3 * N);
  w->Dp = (w->D + 1) / 2 * 2;
  std::vector<double> mask(w->Dp, 0.0);
  std::vector<int> free_atoms;
  for (long i = 0; i < N; i++) {
    bool all_fixed = true;
    for (int k = 0; k < 3; k++) {
      const bool fx = fixed && fixed[3 * i + k] != 0;
      mask[3 * i + k] = fx ? 0.0 : 1.0;
      if (fx) w->any_fixed = 1;
      all_fixed = all_fixed && fx;
    }
    if (!all_fixed) free_atoms.push_back(static_cast<int>(i));
  }
  w->nfree = static_cast<int>(free_atoms.size());
  w->h_box.assign(box, box + 9);
  const size_t WD = static_cast<size_t>(W) * w->Dp;
  bool ok = w->pos.reserve(WD) == cudaSuccess && w->frc.reserve(WD) == cudaSuccess && w->mode.reserve(WD) == cudaSuccess &&
            w->energy.reserve(W) == cudaSuccess && w->dirty.reserve(W) == cudaSuccess &&
            w->mask.reserve(w->Dp) == cudaSuccess && w->free_atoms.reserve(std::max<size_t>(1, free_atoms.size())) == cudaSuccess &&
            w->Z.reserve(N) == cudaSuccess && w->box.reserve(9) == cudaSuccess && w->react.reserve(W) == cudaSuccess &&
            w->eig.reserve(W) == cudaSuccess && w->conv.reserve(W) == cudaSuccess && w->status.reserve(W) == cudaSuccess &&
            w->iters.reserve(W) == cudaSuccess && w->fcalls.reserve(W) == cudaSuccess && w->prm.reserve(1) == cudaSuccess;
  if (!ok) {
    delete w;
    return bail(std::string("walkers_create: out of device memory: ") + cudaGetErrorString(cudaGetLastError()));
  }
  cudaMemcpyAsync(w->mask.p, mask.data(), sizeof(double) * w->Dp, cudaMemcpyHostToDevice, pot->stream);
  if (!free_atoms.empty())
    cudaMemcpyAsync(w->free_atoms.p, free_atoms.data(), sizeof(int) * free_atoms.size(), cudaMemcpyHostToDevice, pot->stream);
  cudaMemcpyAsync(w->Z.p, atomicNrs, sizeof(int) * N, cudaMemcpyHostToDevice, pot->stream);
  cudaMemcpyAsync(w->box.p, box, sizeof(double) * 9, cudaMemcpyHostToDevice, pot->stream);
  cudaMemsetAsync(w->status.p, 0, sizeof(int) * W, pot->stream);
  cudaMemsetAsync(w->iters.p, 0, sizeof(long long) * W, pot->stream);
  cudaMemsetAsync(w->fcalls.p, 0, sizeof(long long) * W, pot->stream);
  cudaMemsetAsync(w->eig.p, 0, sizeof(double) * W, pot->stream);
  cudaMemsetAsync(w->conv.p, 0, sizeof(double) * W, pot->stream);
  cudaMemsetAsync(w->energy.p, 0, sizeof(double) * W, pot->stream);
  cudaMemsetAsync(w->frc.p, 0, sizeof(double) * WD, pot->stream);
  cudaMemsetAsync(w->mode.p, 0, sizeof(double) * WD, pot->stream);
  cudaStreamSynchronize(pot->stream);
  if (eonb200_walkers_set(w, positions, modes) != 0) {
    std::string msg = pot->err;
    delete w;
    return bail(msg);
  }
  return w;
}

void eonb200_walkers_destroy(EonB200Walkers *w) {
  if (!w) return;
  cudaSetDevice(w->pot->device);
  cudaStreamSynchronize(w->pot->stream);
  for (size_t g = 0; g < w->g_stream.size(); g++) {
    cudaStreamSynchronize(w->g_stream[g]);
    cudaEventDestroy(w->g_done[g]);
    cudaStreamDestroy(w->g_stream[g]);
  }
  delete w;
}

int eonb200_walkers_set(EonB200Walkers *w, const double *positions, const double *modes) {
  if (!w) return 1;
  EonB200Pot *pot = w->pot;
  CU_TRY(pot, cudaSetDevice(pot->device));
  const size_t D = w->D, Dp = w->Dp;
  if (positions) {
    CU_TRY(pot, cudaMemcpy2DAsync(w->pos.p, Dp * sizeof(double), positions, D * sizeof(double), D * sizeof(double), w->W,
                                  cudaMemcpyHostToDevice, pot->stream));
    std::vector<int> ones(w->W, 1);
    CU_TRY(pot, cudaMemcpyAsync(w->dirty.p, ones.data(), sizeof(int) * w->W, cudaMemcpyHostToDevice, pot->stream));
    CU_TRY(pot, cudaStreamSynchronize(pot->stream));
  }
  if (modes) {
    CU_TRY(pot, cudaMemcpy2DAsync(w->mode.p, Dp * sizeof(double), modes, D * sizeof(double), D * sizeof(double), w->W,
                                  cudaMemcpyHostToDevice, pot->stream));
    CU_TRY(pot, cudaStreamSynchronize(pot->stream));
    w->have_modes = true;
  }
  return 0;
}

namespace {

// workspace: S_COUNT vector slots + optimizer / min-mode histories, per walker
int walkers_workspace(EonB200Walkers *w, const EonB200SearchParams *prm) {
  EonB200Pot *pot = w->pot;
  const size_t Dp = w->Dp;
  size_t n = static_cast<size_t>(S_COUNT) * Dp;
  w->off_lbfgs_s = w->off_lbfgs_y = w->off_lanczos_q = w->off_rot_s = w->off_rot_y = 0;
  if (prm) {
    if (prm->opt_method == EONB200_OPT_LBFGS) {
      const long m = std::max<long>(1, std::min<long>(prm->lbfgs_memory, kMaxLbfgsMemory));
      w->off_lbfgs_s = n;
      n += m * Dp;
      w->off_lbfgs_y = n;
      n += m * Dp;
    }
    if (prm->minmode_method == EONB200_MINMODE_LANCZOS) {
      const long m = std::max<long>(1, std::min<long>(prm->lanczos_max_iterations, kMaxLanczosIters));
      w->off_lanczos_q = n;
      n += m * Dp;
    } else if (prm->dimer_opt_method == EONB200_ROT_LBFGS) {
      w->off_rot_s = n;
      n += static_cast<size_t>(kMaxRotHistory) * Dp;
      w->off_rot_y = n;
      n += static_cast<size_t>(kMaxRotHistory) * Dp;
    }
  }
  w->ws_stride = n;
  CU_TRY(pot, w->ws.reserve(static_cast<size_t>(w->W) * n));
  return 0;
}

// group < 0: all walkers (or the slice [first, first + count)) on the handle's stream, synchronous, with the row-capacity
// check; group >= 0: the slice on that group's stream, asynchronous (rows must have been sized before).
int walkers_launch(EonB200Walkers *w, int job, const EonB200SearchParams *prm, int remove_net_force, double scalar,
                   long first = 0, long count = -1, int group = -1, const WalkerLaunch *queue = nullptr) {
  EonB200Pot *pot = w->pot;
  CU_TRY(pot, cudaSetDevice(pot->device));
  EonB200SearchParams local;
  if (!prm) {
    eonb200_search_params_default(&local);
    prm = &local;
  }
  const bool search_job = job == 1 || job == 5;  // (5 / 6: searches / relaxations taken from a queue inside the launch)
  if (prm->lanczos_max_iterations > kMaxLanczosIters && prm->minmode_method == EONB200_MINMODE_LANCZOS && search_job)
    return fail(pot, 2, "lanczos_max_iterations above the supported %d", kMaxLanczosIters);
  if (prm->lbfgs_memory > kMaxLbfgsMemory && prm->opt_method == EONB200_OPT_LBFGS && job != 0)
    return fail(pot, 2, "lbfgs_memory above the supported %d", kMaxLbfgsMemory);
  if (prm->dimer_rotations_max >= kMaxRotHistory && prm->dimer_opt_method == EONB200_ROT_LBFGS && search_job)
    return fail(pot, 2, "dimer rotations_max above the supported %d for the lbfgs rotation", kMaxRotHistory - 1);
  int rc = walkers_workspace(w, job == 0 ? nullptr : prm);
  if (rc) return rc;
  cudaStream_t st = group >= 0 ? w->g_stream[group] : pot->stream;
  CU_TRY(pot, cudaMemcpyAsync(w->prm.p, prm, sizeof(*prm), cudaMemcpyHostToDevice, st));
  const long W = w->W, N = w->N;
  if (count < 0) count = W - first;
  if (first < 0 || count <= 0 || first + count > W) return fail(pot, 2, "walkers: slice [%ld, %ld) outside 0..%ld", first, first + count, W);
  if (w->nl.S < W || w->nl.N != N || w->nl.K == 0) {
    rc = nl_alloc(pot, w->nl, W, N, std::max(w->nl.K, guess_K(pot->hp.rc_list, N, w->h_box.data())));
    if (rc) return rc;
  }
  WalkerLaunch wl;
  memset(&wl, 0, sizeof(wl));
  if (queue) wl = *queue;  // (the q_* fields; everything else is set below)
  wl.maxCells = pick_max_cells(N);
  wl.lpa = pick_lpa(pot, N, 512);
  wl.job = job;
  wl.remove_net_force = remove_net_force;
  wl.scalar = scalar;
  wl.first = static_cast<int>(first);
  wl.active = w->use_active ? w->active.p : nullptr;
  const int threads = pick_threads(pot, N, wl.lpa);
  size_t smem;
  rc = check_smem(pot, N, wl.maxCells, &smem);
  if (rc) return rc;
  wl.fast_smem = 0;
  if (pot->cfg.kind == EONB200_POT_EAM) {
    const size_t with_fast = smem_bytes(static_cast<int>(N), wl.maxCells, true);
    if (static_cast<long>(with_fast) <= pot->max_dyn_smem) {
      smem = with_fast;
      wl.fast_smem = 1;
    }
  }
  WalkerArrays wa;
  memset(&wa, 0, sizeof(wa));
  wa.W = static_cast<int>(W);
  wa.N = static_cast<int>(N);
  wa.D = w->D;
  wa.Dp = w->Dp;
  wa.nfree = w->nfree;
  wa.any_fixed = w->any_fixed;
  wa.pos = w->pos.p;
  wa.frc = w->frc.p;
  wa.mode = w->mode.p;
  wa.energy = w->energy.p;
  wa.dirty = w->dirty.p;
  wa.ws = w->ws.p;
  wa.ws_stride = w->ws_stride;
  wa.off_lbfgs_s = w->off_lbfgs_s;
  wa.off_lbfgs_y = w->off_lbfgs_y;
  wa.off_lanczos_q = w->off_lanczos_q;
  wa.off_rot_s = w->off_rot_s;
  wa.off_rot_y = w->off_rot_y;
  wa.mask = w->mask.p;
  wa.free_atoms = w->free_atoms.p;
  wa.Z = w->Z.p;
  wa.box = w->box.p;
  wa.reactant_energy = w->react.p;
  wa.status = w->status.p;
  wa.iterations = w->iters.p;
  wa.force_calls = w->fcalls.p;
  wa.eigenvalue = w->eig.p;
  wa.convergence = w->conv.p;
  // a job that may move atoms needs head-room in the rows: probe the fill with a force-only pass first
  for (int attempt = 0; attempt < 6; attempt++) {
    if (group < 0) {
      CU_TRY(pot, cudaMemsetAsync(pot->d_ovf, 0, sizeof(int), st));
      CU_TRY(pot, cudaMemsetAsync(pot->d_stats.p, 0, 4 * sizeof(unsigned long long), st));
    }
    NlBatch view = nl_view(pot, w->nl);
    if (group < 0) CU_TRY(pot, cudaEventRecord(pot->ev[0], st));
#ifndef EONB200_EMU
    if (threads <= 448 && 2 * (smem + 17 * 1024) <= 220 * 1024 && W > pot->sm_count)
      EONB200_LAUNCH(k_walkers_x2, static_cast<unsigned>(count), threads, smem, st, pot->dP.p, view, wa, wl, w->prm.p);
    else
#endif
      EONB200_LAUNCH(k_walkers, static_cast<unsigned>(count), threads, smem, st, pot->dP.p, view, wa, wl, w->prm.p);
    pot->launches += 1;
    CU_TRY(pot, cudaGetLastError());
    if (group >= 0) {  // asynchronous: the caller polls the group's event and harvests the slice
      CU_TRY(pot, cudaEventRecord(w->g_done[group], st));
      return 0;
    }
    CU_TRY(pot, cudaEventRecord(pot->ev[2], pot->stream));
    CU_TRY(pot, cudaMemcpyAsync(pot->h_ovf, pot->d_ovf, sizeof(int), cudaMemcpyDeviceToHost, pot->stream));
    CU_TRY(pot, cudaMemcpyAsync(pot->h_stats.p, pot->d_stats.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                pot->stream));
    CU_TRY(pot, cudaStreamSynchronize(pot->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, pot->ev[0], pot->ev[2]);
    w->last_ms = ms;
    const int need = pot->h_ovf[0];
    pot->last_stats[0] = static_cast<long long>(pot->h_stats.p[0]);
    pot->last_stats[1] = static_cast<long long>(pot->h_stats.p[1]);
    pot->last_stats[2] = w->nl.K;
    pot->last_stats[3] = need;
    pot->last_stats[4] = static_cast<long long>(pot->h_stats.p[2]);
    if (need >= (1 << 30)) return fail(pot, 100, "neighbour image shift outside the representable range");
    if ((job != 0 && job != 4) || need <= w->nl.K) {
      // count the force calls the walkers made
      if (job != 3 && job != 4 && job < 5) {
        std::vector<long long> fc(count);
        std::vector<int> act(count, 1);
        CU_TRY(pot, cudaMemcpy(fc.data(), w->fcalls.p + first, sizeof(long long) * count, cudaMemcpyDeviceToHost));
        if (w->use_active) CU_TRY(pot, cudaMemcpy(act.data(), w->active.p + first, sizeof(int) * count, cudaMemcpyDeviceToHost));
        for (long k = 0; k < count; k++)
          if (act[k]) pot->force_calls += static_cast<unsigned long long>(fc[k]);
      }
      return 0;
    }
    const int K = (static_cast<int>(need * 1.3) + 8 + 7) / 8 * 8;
    w->nl.K = 0;
    rc = nl_alloc(pot, w->nl, W, N, K);
    if (rc) return rc;
  }
  return fail(pot, 100, "walkers: neighbour list capacity did not converge");
}

// Make sure the rows have head-room before a job that moves atoms: one force-only pass sizes K.
int walkers_size_rows(EonB200Walkers *w) {
  if (w->nl.K != 0 && w->nl.S >= w->W) {
    return 0;
  }
  int rc = walkers_launch(w, 4, nullptr, 1, 0.0);  // probe: build the lists, touch nothing else
  if (rc) return rc;
  EonB200Pot *pot = w->pot;
  const int fill = static_cast<int>(pot->last_stats[3]);
  const int want = (static_cast<int>(fill * 1.3) + 8 + 3) / 4 * 4;
  if (want > w->nl.K) {
    w->nl.K = 0;
    rc = nl_alloc(pot, w->nl, w->W, w->N, want);
  }
  return rc;
}

}  // namespace

int eonb200_walkers_matter_forces(EonB200Walkers *w, int remove_net_force, double *energies, double *forces) {
  if (!w) return 1;
  EonB200Pot *pot = w->pot;
  int rc = walkers_launch(w, 0, nullptr, remove_net_force, 0.0);
  if (rc) return rc;
  if (energies) CU_TRY(pot, cudaMemcpy(energies, w->energy.p, sizeof(double) * w->W, cudaMemcpyDeviceToHost));
  if (forces)
    CU_TRY(pot, cudaMemcpy2D(forces, w->D * sizeof(double), w->frc.p, w->Dp * sizeof(double), w->D * sizeof(double), w->W,
                             cudaMemcpyDeviceToHost));
  return 0;
}

int eonb200_walkers_saddle_search(EonB200Walkers *w, const EonB200SearchParams *prm, const double *reactant_energies) {
  if (!w || !prm) return 1;
  EonB200Pot *pot = w->pot;
  if (!w->have_modes) return fail(pot, 2, "walkers_saddle_search: no initial modes uploaded");
  int rc = walkers_size_rows(w);
  if (rc) return rc;
  if (reactant_energies)
    CU_TRY(pot, cudaMemcpyAsync(w->react.p, reactant_energies, sizeof(double) * w->W, cudaMemcpyHostToDevice, pot->stream));
  else
    CU_TRY(pot, cudaMemsetAsync(w->react.p, 0, sizeof(double) * w->W, pot->stream));
  CU_TRY(pot, cudaStreamSynchronize(pot->stream));
  return walkers_launch(w, 1, prm, prm->remove_net_force, 0.0);
}

int eonb200_walkers_relax(EonB200Walkers *w, const EonB200SearchParams *prm) {
  if (!w || !prm) return 1;
  int rc = walkers_size_rows(w);
  if (rc) return rc;
  return walkers_launch(w, 2, prm, prm->remove_net_force, 0.0);
}

int eonb200_walkers_queue(EonB200Walkers *w, int job, const EonB200SearchParams *prm, long nQueue, double *positions,
                          double *modes, const double *reactant_energies, int *status, long *iterations, long *force_calls,
                          double *energies, double *eigenvalues, double *convergence) {
  if (!w || !prm || !positions || nQueue <= 0) return 1;
  EonB200Pot *pot = w->pot;
  if (job != 1 && job != 2) return fail(pot, 2, "walkers_queue: job must be 1 (saddle search) or 2 (relax)");
  if (job == 1 && !modes) return fail(pot, 2, "walkers_queue: saddle searches need initial modes");
  CU_TRY(pot, cudaSetDevice(pot->device));
  int rc = walkers_size_rows(w);
  if (rc) return rc;
  const size_t Q = static_cast<size_t>(nQueue), D = static_cast<size_t>(w->D), Dp = static_cast<size_t>(w->Dp);
  CU_TRY(pot, w->q_pos.reserve(Q * Dp));
  if (job == 1) CU_TRY(pot, w->q_mode.reserve(Q * Dp));
  CU_TRY(pot, w->q_react.reserve(Q));
  CU_TRY(pot, w->q_energy.reserve(Q));
  CU_TRY(pot, w->q_eig.reserve(Q));
  CU_TRY(pot, w->q_conv.reserve(Q));
  CU_TRY(pot, w->q_status.reserve(Q));
  CU_TRY(pot, w->q_iter.reserve(Q));
  CU_TRY(pot, w->q_fcalls.reserve(Q));
  CU_TRY(pot, w->q_next.reserve(1));
  cudaStream_t st = pot->stream;
  CU_TRY(pot, cudaMemsetAsync(w->q_pos.p, 0, sizeof(double) * Q * Dp, st));
  CU_TRY(pot, cudaMemcpy2DAsync(w->q_pos.p, Dp * sizeof(double), positions, D * sizeof(double), D * sizeof(double), Q,
                                cudaMemcpyHostToDevice, st));
  if (job == 1) {
    CU_TRY(pot, cudaMemsetAsync(w->q_mode.p, 0, sizeof(double) * Q * Dp, st));
    CU_TRY(pot, cudaMemcpy2DAsync(w->q_mode.p, Dp * sizeof(double), modes, D * sizeof(double), D * sizeof(double), Q,
                                  cudaMemcpyHostToDevice, st));
  }
  if (reactant_energies)
    CU_TRY(pot, cudaMemcpyAsync(w->q_react.p, reactant_energies, sizeof(double) * Q, cudaMemcpyHostToDevice, st));
  else
    CU_TRY(pot, cudaMemsetAsync(w->q_react.p, 0, sizeof(double) * Q, st));
  CU_TRY(pot, cudaMemsetAsync(w->q_next.p, 0, sizeof(int), st));
  CU_TRY(pot, cudaMemsetAsync(w->q_status.p, 0, sizeof(int) * Q, st));
  CU_TRY(pot, cudaStreamSynchronize(st));
  WalkerLaunch q;
  memset(&q, 0, sizeof(q));
  q.q_next = w->q_next.p;
  q.q_len = static_cast<int>(nQueue);
  q.q_pos = w->q_pos.p;
  q.q_mode = job == 1 ? w->q_mode.p : nullptr;
  q.q_react = w->q_react.p;
  q.q_energy = w->q_energy.p;
  q.q_eig = w->q_eig.p;
  q.q_conv = w->q_conv.p;
  q.q_status = w->q_status.p;
  q.q_iter = w->q_iter.p;
  q.q_fcalls = w->q_fcalls.p;
  const long ctas = std::min<long>(w->W, nQueue);
  rc = walkers_launch(w, job == 1 ? 5 : 6, prm, prm->remove_net_force, 0.0, 0, ctas, -1, &q);
  if (rc) return rc;
  std::vector<long long> it(Q), fc(Q);
  CU_TRY(pot, cudaMemcpy(it.data(), w->q_iter.p, sizeof(long long) * Q, cudaMemcpyDeviceToHost));
  CU_TRY(pot, cudaMemcpy(fc.data(), w->q_fcalls.p, sizeof(long long) * Q, cudaMemcpyDeviceToHost));
  for (size_t k = 0; k < Q; k++) {
    pot->force_calls += static_cast<unsigned long long>(fc[k]);
    if (iterations) iterations[k] = static_cast<long>(it[k]);
    if (force_calls) force_calls[k] = static_cast<long>(fc[k]);
  }
  if (status) CU_TRY(pot, cudaMemcpy(status, w->q_status.p, sizeof(int) * Q, cudaMemcpyDeviceToHost));
  if (energies) CU_TRY(pot, cudaMemcpy(energies, w->q_energy.p, sizeof(double) * Q, cudaMemcpyDeviceToHost));
  if (eigenvalues) CU_TRY(pot, cudaMemcpy(eigenvalues, w->q_eig.p, sizeof(double) * Q, cudaMemcpyDeviceToHost));
  if (convergence) CU_TRY(pot, cudaMemcpy(convergence, w->q_conv.p, sizeof(double) * Q, cudaMemcpyDeviceToHost));
  CU_TRY(pot, cudaMemcpy2D(positions, D * sizeof(double), w->q_pos.p, Dp * sizeof(double), D * sizeof(double), Q,
                           cudaMemcpyDeviceToHost));
  if (job == 1)
    CU_TRY(pot, cudaMemcpy2D(modes, D * sizeof(double), w->q_mode.p, Dp * sizeof(double), D * sizeof(double), Q,
                             cudaMemcpyDeviceToHost));
  return 0;
}

int eonb200_walkers_displace_along_mode(EonB200Walkers *w, double signed_offset) {
  if (!w) return 1;
  return walkers_launch(w, 3, nullptr, 1, signed_offset);
}

int eonb200_walkers_harvest(EonB200Walkers *w, int *status, long *iterations, long *force_calls, double *energies,
                            double *eigenvalues, double *convergence, double *positions, double *modes) {
  if (!w) return 1;
  EonB200Pot *pot = w->pot;
  CU_TRY(pot, cudaSetDevice(pot->device));
  CU_TRY(pot, cudaStreamSynchronize(pot->stream));
  const long W = w->W;
  if (status) CU_TRY(pot, cudaMemcpy(status, w->status.p, sizeof(int) * W, cudaMemcpyDeviceToHost));
  if (iterations) {
    std::vector<long long> t(W);
    CU_TRY(pot, cudaMemcpy(t.data(), w->iters.p, sizeof(long long) * W, cudaMemcpyDeviceToHost));
    for (long k = 0; k < W; k++) iterations[k] = static_cast<long>(t[k]);
  }
  if (force_calls) {
    std::vector<long long> t(W);
    CU_TRY(pot, cudaMemcpy(t.data(), w->fcalls.p, sizeof(long long) * W, cudaMemcpyDeviceToHost));
    for (long k = 0; k < W; k++) force_calls[k] = static_cast<long>(t[k]);
  }
  if (energies) CU_TRY(pot, cudaMemcpy(energies, w->energy.p, sizeof(double) * W, cudaMemcpyDeviceToHost));
  if (eigenvalues) CU_TRY(pot, cudaMemcpy(eigenvalues, w->eig.p, sizeof(double) * W, cudaMemcpyDeviceToHost));
  if (convergence) CU_TRY(pot, cudaMemcpy(convergence, w->conv.p, sizeof(double) * W, cudaMemcpyDeviceToHost));
  if (positions)
    CU_TRY(pot, cudaMemcpy2D(positions, w->D * sizeof(double), w->pos.p, w->Dp * sizeof(double), w->D * sizeof(double), W,
                             cudaMemcpyDeviceToHost));
  if (modes)
    CU_TRY(pot, cudaMemcpy2D(modes, w->D * sizeof(double), w->mode.p, w->Dp * sizeof(double), w->D * sizeof(double), W,
                             cudaMemcpyDeviceToHost));
  return 0;
}

int eonb200_walkers_displace(EonB200Walkers *w, const double *reference_positions, long nReferences, const int *epicentres,
                             double radius, double sigma, unsigned long long seed, long first_walker_id) {
  if (!w || !reference_positions || !epicentres) return 1;
  EonB200Pot *pot = w->pot;
  if (nReferences != 1 && nReferences != w->W) return fail(pot, 2, "walkers_displace: one reference or one per walker");
  if (!(radius >= 0.0) || !(sigma >= 0.0)) return fail(pot, 2, "walkers_displace: radius and magnitude must not be negative");
  for (long k = 0; k < w->W; k++)
    if (epicentres[k] < 0 || epicentres[k] >= w->N) return fail(pot, 2, "walkers_displace: epicentre %d of walker %ld", epicentres[k], k);
  CU_TRY(pot, cudaSetDevice(pot->device));
  const size_t D = w->D, Dp = w->Dp;
  DevBuf<double> ref;
  DevBuf<int> epi;
  CU_TRY(pot, ref.reserve(static_cast<size_t>(nReferences) * Dp));
  CU_TRY(pot, epi.reserve(static_cast<size_t>(w->W)));
  CU_TRY(pot, cudaMemcpy2DAsync(ref.p, Dp * sizeof(double), reference_positions, D * sizeof(double), D * sizeof(double),
                                nReferences, cudaMemcpyHostToDevice, pot->stream));
  CU_TRY(pot, cudaMemcpyAsync(epi.p, epicentres, sizeof(int) * w->W, cudaMemcpyHostToDevice, pot->stream));
  WalkerArrays wa;
  memset(&wa, 0, sizeof(wa));
  wa.W = static_cast<int>(w->W);
  wa.N = static_cast<int>(w->N);
  wa.D = w->D;
  wa.Dp = w->Dp;
  wa.pos = w->pos.p;
  wa.mode = w->mode.p;
  wa.dirty = w->dirty.p;
  wa.mask = w->mask.p;
  wa.box = w->box.p;
  DisplaceArgs da;
  da.ref = ref.p;
  da.ref_stride = nReferences == 1 ? 0 : Dp;
  da.epicentre = epi.p;
  da.radius = radius;
  da.sigma = sigma;
  da.seed = seed;
  da.first_id = first_walker_id;
  const int threads = std::max(32, std::min<int>(kMaxThreadsPerWalker, static_cast<int>((w->N + 31) / 32 * 32)));
#ifdef EONB200_EMU
  EONB200_LAUNCH(k_displace, static_cast<unsigned>(w->W), 1, 0, pot->stream, wa, da);
  (void)threads;
#else
  EONB200_LAUNCH(k_displace, static_cast<unsigned>(w->W), threads, 0, pot->stream, wa, da);
#endif
  pot->launches += 1;
  CU_TRY(pot, cudaGetLastError());
  CU_TRY(pot, cudaStreamSynchronize(pot->stream));
  w->have_modes = true;
  return 0;
}

int eonb200_walkers_set_active(EonB200Walkers *w, const int *active) {
  if (!w) return 1;
  EonB200Pot *pot = w->pot;
  CU_TRY(pot, cudaSetDevice(pot->device));
  w->use_active = active != nullptr;
  if (active) {
    CU_TRY(pot, w->active.reserve(static_cast<size_t>(w->W)));
    CU_TRY(pot, cudaMemcpy(w->active.p, active, sizeof(int) * w->W, cudaMemcpyHostToDevice));
  }
  return 0;
}

int eonb200_walkers_groups(EonB200Walkers *w, int nGroups) {
  if (!w) return 1;
  EonB200Pot *pot = w->pot;
  if (nGroups < 1 || nGroups > 64) return fail(pot, 2, "walkers_groups: 1..64 groups");
  CU_TRY(pot, cudaSetDevice(pot->device));
  int rc = walkers_size_rows(w);  // asynchronous jobs cannot regrow the neighbour rows: size them now
  if (rc) return rc;
  while (static_cast<int>(w->g_stream.size()) < nGroups) {
    cudaStream_t st = nullptr;
    cudaEvent_t ev = nullptr;
#ifdef EONB200_EMU
    CU_TRY(pot, cudaStreamCreate(&st));
    CU_TRY(pot, cudaEventCreate(&ev));
#else
    CU_TRY(pot, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CU_TRY(pot, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
#endif
    w->g_stream.push_back(st);
    w->g_done.push_back(ev);
    w->g_range.push_back({0, 0});
  }
  return 0;
}

int eonb200_walkers_submit(EonB200Walkers *w, int group, int job, const EonB200SearchParams *prm, long first, long count,
                           const double *positions, const double *modes, const double *reactant_energies) {
  if (!w || !prm) return 1;
  EonB200Pot *pot = w->pot;
  if (group < 0 || group >= static_cast<int>(w->g_stream.size())) return fail(pot, 2, "walkers_submit: no such group");
  if (job != 1 && job != 2) return fail(pot, 2, "walkers_submit: job must be 1 (saddle search) or 2 (relax)");
  if (first < 0 || count <= 0 || first + count > w->W) return fail(pot, 2, "walkers_submit: slice outside the walkers");
  if (w->g_range[group][1] != 0) return fail(pot, 2, "walkers_submit: group %d still has a job in flight", group);
  CU_TRY(pot, cudaSetDevice(pot->device));
  cudaStream_t st = w->g_stream[group];
  const size_t D = w->D, Dp = w->Dp;
  if (positions) {
    CU_TRY(pot, cudaMemcpy2DAsync(w->pos.p + first * Dp, Dp * sizeof(double), positions, D * sizeof(double),
                                  D * sizeof(double), count, cudaMemcpyHostToDevice, st));
    std::vector<int> ones(count, 1);
    CU_TRY(pot, cudaMemcpyAsync(w->dirty.p + first, ones.data(), sizeof(int) * count, cudaMemcpyHostToDevice, st));
    CU_TRY(pot, cudaStreamSynchronize(st));  // (ones is a stack object; the copies are a few KB)
  }
  if (modes) {
    CU_TRY(pot, cudaMemcpy2DAsync(w->mode.p + first * Dp, Dp * sizeof(double), modes, D * sizeof(double), D * sizeof(double),
                                  count, cudaMemcpyHostToDevice, st));
    w->have_modes = true;
  }
  if (job == 1 && !w->have_modes) return fail(pot, 2, "walkers_submit: no initial modes uploaded");
  if (reactant_energies)
    CU_TRY(pot, cudaMemcpyAsync(w->react.p + first, reactant_energies, sizeof(double) * count, cudaMemcpyHostToDevice, st));
  else if (job == 1)
    CU_TRY(pot, cudaMemsetAsync(w->react.p + first, 0, sizeof(double) * count, st));
  const int rc = walkers_launch(w, job, prm, prm->remove_net_force, 0.0, first, count, group);
  if (rc) return rc;
  w->g_range[group] = {first, count};
  return 0;
}

int eonb200_walkers_poll(EonB200Walkers *w, int group, int wait) {
  if (!w) return -1;
  EonB200Pot *pot = w->pot;
  if (group < 0 || group >= static_cast<int>(w->g_stream.size())) return -1;
  if (w->g_range[group][1] == 0) return 1;  // idle
#ifdef EONB200_EMU
  (void)wait;
  return 1;  // the emulation runs jobs to completion inside submit
#else
  cudaSetDevice(pot->device);
  if (wait) return cudaEventSynchronize(w->g_done[group]) == cudaSuccess ? 1 : -1;
  const cudaError_t e = cudaEventQuery(w->g_done[group]);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) return 0;
  cudaGetLastError();
  return -1;
#endif
}

int eonb200_walkers_harvest_group(EonB200Walkers *w, int group, int *status, long *iterations, long *force_calls,
                                  double *energies, double *eigenvalues, double *convergence, double *positions,
                                  double *modes) {
  if (!w) return 1;
  EonB200Pot *pot = w->pot;
  if (group < 0 || group >= static_cast<int>(w->g_stream.size())) return fail(pot, 2, "walkers_harvest_group: no such group");
  const long first = w->g_range[group][0], n = w->g_range[group][1];
  if (n == 0) return fail(pot, 2, "walkers_harvest_group: group %d is idle", group);
  CU_TRY(pot, cudaSetDevice(pot->device));
  cudaStream_t st = w->g_stream[group];
  CU_TRY(pot, cudaStreamSynchronize(st));
  const size_t D = w->D, Dp = w->Dp;
  std::vector<long long> it(n), fc(n);
  if (status) CU_TRY(pot, cudaMemcpyAsync(status, w->status.p + first, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaMemcpyAsync(it.data(), w->iters.p + first, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaMemcpyAsync(fc.data(), w->fcalls.p + first, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
  if (energies) CU_TRY(pot, cudaMemcpyAsync(energies, w->energy.p + first, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (eigenvalues) CU_TRY(pot, cudaMemcpyAsync(eigenvalues, w->eig.p + first, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (convergence) CU_TRY(pot, cudaMemcpyAsync(convergence, w->conv.p + first, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (positions)
    CU_TRY(pot, cudaMemcpy2DAsync(positions, D * sizeof(double), w->pos.p + first * Dp, Dp * sizeof(double),
                                  D * sizeof(double), n, cudaMemcpyDeviceToHost, st));
  if (modes)
    CU_TRY(pot, cudaMemcpy2DAsync(modes, D * sizeof(double), w->mode.p + first * Dp, Dp * sizeof(double), D * sizeof(double),
                                  n, cudaMemcpyDeviceToHost, st));
  CU_TRY(pot, cudaStreamSynchronize(st));
  for (long k = 0; k < n; k++) {
    if (iterations) iterations[k] = static_cast<long>(it[k]);
    if (force_calls) force_calls[k] = static_cast<long>(fc[k]);
    pot->force_calls += static_cast<unsigned long long>(fc[k]);
  }
  w->g_range[group] = {0, 0};
  return 0;
}

double eonb200_walkers_last_ms(const EonB200Walkers *w) { return w ? w->last_ms : 0.0; }

double eonb200_measure_fp64_peak(EonB200Pot *pot, int repeats) {
  if (!pot) return 0.0;
  cudaSetDevice(pot->device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, pot->device) != cudaSuccess) return 0.0;
  DevBuf<double> out;
  if (out.reserve(1) != cudaSuccess) return 0.0;
  const int iters = 1 << 14, threads = 256, blocks = prop.multiProcessorCount * 8;
  double best = 0.0;
  for (int r = 0; r < std::max(1, repeats) + 1; r++) {
    cudaEventRecord(pot->ev[0], pot->stream);
    EONB200_LAUNCH(k_dfma_peak, blocks, threads, 0, pot->stream, out.p, iters);
    cudaEventRecord(pot->ev[2], pot->stream);
    cudaStreamSynchronize(pot->stream);
    pot->launches += 1;
    float ms = 0;
    cudaEventElapsedTime(&ms, pot->ev[0], pot->ev[2]);
    if (r == 0 || ms <= 0) continue;  // warm-up
    const double flops = 2.0 * 8.0 * static_cast<double>(iters) * threads * blocks;
    best = std::max(best, flops / (ms * 1e-3));
  }
  return best;
}

}  // extern "C"
