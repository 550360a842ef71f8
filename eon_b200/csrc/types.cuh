// types.cuh -- device-side parameter blocks and the per-CTA view of one walker.
#pragma once
#include "common.cuh"

namespace eonb200 {

enum PotKind { POT_LJ = 0, POT_LJCLUSTER = 1, POT_MORSE_PT = 2, POT_EAM = 3 };
enum NlMode { NL_VESIN = 0, NL_MIC = 1 };

constexpr int kMaxEamRows = 8;
constexpr int kMaxTiles = 16;

// One element row of include/eon/potentials/EAM/Parameters.h:15-50 plus derived constants.
struct EamRowDev {
  int Z;
  double Dm, alphaM, Rm, beta1, beta2, r_cut, rc2;
  double c[9];   // embedding polynomial c0..c8 (EAM.cpp:349-356)
  double dc[9];  // k * c[k], derivative polynomial (EAM.cpp:358-365)
};

struct PotParams {
  int kind;
  int nl_mode;            // NL_VESIN for the rgpot pair potentials, NL_MIC for EAM.cpp's minimum image
  double cutoff, cutoff2; // potential cutoff (LJ / Morse); for EAM the largest row cutoff
  double rc_list;         // cutoff + skin: what the cached candidate list is built with
  double rc_list2;
  double skin;
  double eshift;          // U(rc) subtracted per pair (LJ, Morse_Pt)
  double lj_u0, lj_psi;
  double morse_De, morse_a, morse_re;
  int n_rows;
  int eam_conservative;
  int eam_fast_ok;        // every EAM row (or the Morse parameters) keeps its exp arguments inside exp_tab_fast's
                          // range (set by the host)
  const double *exp_tab;  // [128] 2^(j/128), device memory (table of the exp used by the pair loops)
  double exp_t16[16];     // 2^(j/16): the one-line table of exp_t16 (Newton's-third-law EAM kernel)
  EamRowDev rows[kMaxEamRows];
};

// Shift code packed above the neighbour index: j in bits 0..19, (s+8) per axis in 4 bits each.
constexpr int kNbrIndexBits = 20;
constexpr int kNbrIndexMask = (1 << kNbrIndexBits) - 1;
constexpr int kShiftNeutral = (8 << 0) | (8 << 4) | (8 << 8);
__device__ __host__ __forceinline__ int pack_nbr(int j, int sx, int sy, int sz) {
  return j | ((sx + 8) << kNbrIndexBits) | ((sy + 8) << (kNbrIndexBits + 4)) | ((sz + 8) << (kNbrIndexBits + 8));
}
__device__ __host__ __forceinline__ void unpack_nbr(int code, int &j, int &sx, int &sy, int &sz) {
  j = code & kNbrIndexMask;
  const int s = static_cast<unsigned>(code) >> kNbrIndexBits;
  sx = (s & 15) - 8;
  sy = ((s >> 4) & 15) - 8;
  sz = ((s >> 8) & 15) - 8;
}

// NL_MIC lists (EAM.cpp's single minimum image) only ever fold by -1, 0 or +1 box lengths per axis: their codes carry
// one image id in [0, 27) above the index, which selects a ready-made Cartesian offset (Smem::sh27).
__device__ __host__ __forceinline__ int pack_mic(int j, int fx, int fy, int fz) {
  return j | (((fx + 1) + 3 * (fy + 1) + 9 * (fz + 1)) << kNbrIndexBits);
}

// Cached neighbour lists of a batch of systems (one row block per system).
struct NlBatch {
  int *nbr;        // [S][K][Np]  packed (j, shift); slot-major so that lanes = atoms read coalesced
  int *cnt;        // [S][Np]     row fill
  double *rbuild;  // [S][3N]     positions the list was built from (Verlet displacement check)
  int *state;      // [S]         0 = no list yet, 1 = list valid for rbuild
  double *box_build;  // [S][9]   cell the list was built for (a different cell invalidates it)
  int K, Np;
  int *overflow;             // [1] largest row fill any build needed (> K means the list is truncated)
  unsigned long long *stats; // [0] rebuilt, [1] reused, [2] ordered pairs of the last evaluation
  // scratch of the tiled / two-pass force pipeline
  double *pairA;    // [S][K][Np] EAM: g'(r)/r of the ordered pair in slot (k, i), written by the density pass
  double *pairB;    // [S][K][Np] EAM: +-phi'(r)/r
  int *nbr2;        // [S][K][Np] fused EAM kernel: this evaluation's in-cutoff pairs, compacted per atom
  int *cnt2;        // [S][Np]    their count (only used when a thread owns more than one atom)
  int *cls;         // [S][Np]    NL_MIC rows: sure-in count | uncertain count << 16 (NlParams::dstar)
  int *cellscratch; // [S][scratch_stride] cell-list scratch of an in-kernel rebuild (fused kernel)
  int scratch_stride;
  double4 *pos4;    // [S][Np]    prepared positions (stateless tiled kernels)
  double *dfd;      // [S][Np]    EAM dF/drho
  unsigned char *rowid;  // [S][Np]
  int *flags;       // [S] bit 0: more than one element, bit 1: element without a row
  double *upart;    // [S][kMaxTiles][4] energy / pair-count partials per tile and pass
  int *done;        // [S] tiles finished (last one sums the partials in tile order)
  // owner / receiver lists of the Newton's-third-law EAM kernel (eam_n3.cuh).  Every unordered pair of the sure-in
  // and uncertain classes is OWNED by one of its atoms (parity rule): the owner evaluates it and hands the partner's
  // share over through a slot of shared memory.  Slot-major like nbr (lanes = atoms read consecutive words).
  int2 *own_cs;     // [S][K][Np] pairs atom i owns, sure-in first, then uncertain: x = code (32 j | fold bits << 16,
                    //            eam_n3.cuh), y = byte offset of the partner's receiver slot in the CTA's slot buffer
  int *recv_code;   // [S][K][Np] codes of the pairs atom i receives, in slot order
  int4 *lhdr;       // [S][2 Np]  list column t (n3_lpa columns per atom): x = owned sure-in | owned << 16, y = received, z = first slot,
                    //            w = atom | (fold is -L along x, y, z) << 24
  int *lane_atom;   // [S][Np]    atom of lane position t (atoms ranked by owned pairs); the lists above are stored by lane
  int *atom_lane;   // [S][Np]    its inverse
  int *roff;        // [S][2 Np + 1] list-build scratch: exclusive scan of the received counts
  double2 *own_ab;  // [S][K][Np] (g'(r)/r, -(+-phi'(r)/r)) of the owned pairs, written by the density pass
  int *hl_hdr;      // [S][8]     [0] slots the fast path needs, [1] 1 = lists valid, [2] list rows in use,
                    //            [3] element row | Z << 8, [4] fold-sign conflict
  double *lo_build; // [S][4]     coordinate shift (EAM.cpp:56-83) the lists were built under
  int slot_cap;     // receiver slots the launch has shared memory for
  int n3_lpa;       // lanes (list columns) per atom of this launch, 1 or 2
  int n3_shift;     // 4 + log2(copies of every atom in the kernel's shared position arrays): 7, 6, 5 or 4
};

// One system as seen by its CTA.
struct SysView {
  int N;
  int Rcap;          // doubles readable at R (>= 3N; the row stride of the walker arrays)
  const double *R;   // [3N] positions, global
  const int *Z;      // [N]
  const double *box; // [9]
  int *nbr;          // this system's [K][Np] block
  int *cnt;          // [Np]
  double *rbuild;    // [3N]
  int *state;        // [1]
  double *box_build; // [9]
  int K, Np;
  int *overflow;
  unsigned long long *stats;
  double *pairA, *pairB;  // this system's [K][Np] blocks (null for the pair potentials)
  int *nbr2;
  int *cnt2;
  int *cls;
  int *cellscratch;
  double4 *pos4;
  double *dfd;
  unsigned char *rowid;
  int *flags;
  double *upart;
  int *done;
  int2 *own_cs;
  int *recv_code;
  int4 *lhdr;
  int *lane_atom, *atom_lane;
  int *roff;
  double2 *own_ab;
  int *hl_hdr;
  double *lo_build;
  int slot_cap;
  int n3_lpa;
  int n3_shift;
};

__device__ __forceinline__ SysView make_sys(const NlBatch &nl, int s, int N, const double *R, const int *Z,
                                            const double *box) {
  SysView v;
  v.N = N;
  v.Rcap = 3 * N;
  v.R = R;
  v.Z = Z;
  v.box = box;
  v.nbr = nl.nbr + static_cast<size_t>(s) * nl.K * nl.Np;
  v.cnt = nl.cnt + static_cast<size_t>(s) * nl.Np;
  v.rbuild = nl.rbuild + static_cast<size_t>(s) * 3 * N;
  v.state = nl.state + s;
  v.box_build = nl.box_build ? nl.box_build + static_cast<size_t>(s) * 9 : nullptr;
  v.K = nl.K;
  v.Np = nl.Np;
  v.overflow = nl.overflow;
  v.stats = nl.stats;
  v.pairA = nl.pairA ? nl.pairA + static_cast<size_t>(s) * nl.K * nl.Np : nullptr;
  v.pairB = nl.pairB ? nl.pairB + static_cast<size_t>(s) * nl.K * nl.Np : nullptr;
  v.nbr2 = nl.nbr2 ? nl.nbr2 + static_cast<size_t>(s) * nl.K * nl.Np : nullptr;
  v.cnt2 = nl.cnt2 ? nl.cnt2 + static_cast<size_t>(s) * nl.Np : nullptr;
  v.cls = nl.cls ? nl.cls + static_cast<size_t>(s) * nl.Np : nullptr;
  v.cellscratch = nl.cellscratch ? nl.cellscratch + static_cast<size_t>(s) * nl.scratch_stride : nullptr;
  v.pos4 = nl.pos4 ? nl.pos4 + static_cast<size_t>(s) * nl.Np : nullptr;
  v.dfd = nl.dfd ? nl.dfd + static_cast<size_t>(s) * nl.Np : nullptr;
  v.rowid = nl.rowid ? nl.rowid + static_cast<size_t>(s) * nl.Np : nullptr;
  v.flags = nl.flags ? nl.flags + s : nullptr;
  v.upart = nl.upart ? nl.upart + static_cast<size_t>(s) * kMaxTiles * 4 : nullptr;
  v.done = nl.done ? nl.done + s : nullptr;
  const size_t kn = static_cast<size_t>(s) * nl.K * nl.Np;
  v.own_cs = nl.own_cs ? nl.own_cs + kn : nullptr;
  v.recv_code = nl.recv_code ? nl.recv_code + kn : nullptr;
  v.own_ab = nl.own_ab ? nl.own_ab + kn : nullptr;
  v.lhdr = nl.lhdr ? nl.lhdr + static_cast<size_t>(s) * nl.Np * 2 : nullptr;
  v.roff = nl.roff ? nl.roff + static_cast<size_t>(s) * (2 * nl.Np + 1) : nullptr;
  v.lane_atom = nl.lane_atom ? nl.lane_atom + static_cast<size_t>(s) * nl.Np : nullptr;
  v.atom_lane = nl.atom_lane ? nl.atom_lane + static_cast<size_t>(s) * nl.Np : nullptr;
  v.hl_hdr = nl.hl_hdr ? nl.hl_hdr + static_cast<size_t>(s) * 8 : nullptr;
  v.lo_build = nl.lo_build ? nl.lo_build + static_cast<size_t>(s) * 4 : nullptr;
  v.slot_cap = nl.slot_cap;
  v.n3_lpa = nl.n3_lpa;
  v.n3_shift = nl.n3_shift;
  return v;
}

// The same store seen from system s0 on (a launch over a chunk of the batch): every per-system array moves by s0
// strides -- the strides of make_sys.
__host__ __device__ inline void offset_systems(NlBatch &nl, int N, long s0) {
  const size_t s = static_cast<size_t>(s0);
  const size_t kn = s * nl.K * nl.Np;
#define EONB200_OFF(field, stride) \
  if (nl.field) nl.field += s * (stride)
  nl.nbr += kn;
  nl.cnt += s * nl.Np;
  nl.rbuild += s * 3 * N;
  nl.state += s;
  EONB200_OFF(box_build, 9);
  if (nl.pairA) nl.pairA += kn;
  if (nl.pairB) nl.pairB += kn;
  if (nl.nbr2) nl.nbr2 += kn;
  EONB200_OFF(cnt2, nl.Np);
  EONB200_OFF(cls, nl.Np);
  EONB200_OFF(cellscratch, nl.scratch_stride);
  EONB200_OFF(pos4, nl.Np);
  EONB200_OFF(dfd, nl.Np);
  EONB200_OFF(rowid, nl.Np);
  EONB200_OFF(flags, 1);
  EONB200_OFF(upart, kMaxTiles * 4);
  EONB200_OFF(done, 1);
  if (nl.own_cs) nl.own_cs += kn;
  if (nl.recv_code) nl.recv_code += kn;
  if (nl.own_ab) nl.own_ab += kn;
  EONB200_OFF(lhdr, nl.Np * 2);
  EONB200_OFF(roff, 2 * nl.Np + 1);
  EONB200_OFF(lane_atom, nl.Np);
  EONB200_OFF(atom_lane, nl.Np);
  EONB200_OFF(hl_hdr, 8);
  EONB200_OFF(lo_build, 4);
#undef EONB200_OFF
}

// Cell geometry of one system, computed by thread 0 of the CTA from the box.
struct Geom {
  double H[9];
  double Hinv[9];
  double lo[3];     // MIC mode: per-axis coordinate shift of EAM.cpp:56-83 (|min(0, lowest coordinate)|)
  double fmin[3];   // VESIN mode, non-periodic axes: lowest fractional coordinate
  double fspan[3];  // VESIN mode, non-periodic axes: fractional extent
  int nc[3];        // cells per axis
  int ns[3];        // search range in cells per axis
  int per[3];       // periodic flags
  int ncells;
  int ok;           // box invertible
  int tab_ok;       // the per-axis shift tables (Smem::shx/shy/shz) are valid for this box
  double disp2;     // largest squared displacement of the prepared positions from the list's build positions
};

// Carve-up of the dynamic shared memory of a walker CTA.
struct Smem {
  double4 *pos;      // [N]   staged positions; .w carries dF/drho in the EAM force pass
  double *red;       // [8 * kMaxWarps] reduction scratch
  double *dfd;       // [N]   EAM: dF/drho per atom
  double *stage;     // [3N + 2] landing zone of the bulk copy / outgoing forces
  int *cell_of;      // [N]
  int *wrapc;        // [N]   packed wrap shift of each atom (VESIN mode)
  int *order;        // [N]   atoms sorted by cell
  int *cell_start;   // [maxCells + 1]
  int *cell_fill;    // [maxCells + 1]
  unsigned char *rowid;  // [N] EAM row of each atom
  uint64_t *mbar;
  Geom *geom;
  double *exp_tab;   // [128] 2^(j/128)
  double *shx, *shy, *shz;  // [16] each: image shift s in [-8, 7] -> Cartesian offset along the axis (NL_VESIN)
  double4 *sh27;     // [27] NL_MIC: image id -> offset added to r_i - r_j (fold +1 means "d > L/2 -> d -= L")
  float4 *sh27f;     // [27] single-precision copy (fused kernel pre-filter)
  float4 *posf;      // [N]   fused kernel: single-precision copy of the positions (pre-filter)
  double *clu;       // [2 * 8] fused kernel: energy / pair-count partials of the CTAs of a cluster (rank 0 sums them)
  double *rho;       // (unused)
  double *facc;      // (unused)
  // Newton's-third-law EAM kernel (eam_n3.cuh)
  double2 *pxy;      // [(N + 1) kN3Rep] (x, y) of the prepared positions, kN3Rep copies per atom (one per 16-byte bank group)
  double2 *pzw;      // [(N + 1) kN3Rep] (z, F'(rho))
  double *t16;       // [16] 2^(j/16), one 128-byte line: any lookup of a warp is conflict free
  double *slots;     // [slot_cap] receiver slots: the partner's share of a pair (g, then the force scalar); landing
                     //            zone of the raw coordinates before that (slot_cap >= 3N + 4)
  int maxCells;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// `fast`: also room for the single-precision copies the fast EAM lane functions filter with -- only the resident
// walker kernel asks for it, and only when it fits.
__host__ __device__ inline size_t smem_bytes(int N, int maxCells, bool fast = false) {
  size_t b = 0;
  b += align_up(sizeof(double4) * (N + 1), 128);
  if (fast) b += align_up(sizeof(float4) * (N + 1), 128) + sizeof(float4) * 28;
  b += align_up(sizeof(double) * 8 * kMaxWarps, 128);
  b += align_up(sizeof(double) * N, 128);
  b += align_up(sizeof(double) * (3 * static_cast<size_t>(N) + 2), 128);
  b += 3 * align_up(sizeof(int) * N, 128);
  b += 2 * align_up(sizeof(int) * (maxCells + 1), 128);
  b += align_up(static_cast<size_t>(N), 128);
  b += 128;  // mbarrier
  b += align_up(sizeof(Geom), 128);
  b += sizeof(double) * (128 + 48) + sizeof(double4) * 28;
  return b;
}

// Tiled force kernels need only the positions, reduction scratch, row ids and tables.
__host__ __device__ inline size_t smem_bytes_tile(int N) {
  size_t b = 0;
  b += align_up(sizeof(double4) * N, 128);
  b += align_up(sizeof(double) * 8 * kMaxWarps, 128);
  b += align_up(static_cast<size_t>(N), 128);
  b += 128;
  b += align_up(sizeof(Geom), 128);
  b += sizeof(double) * (128 + 48) + sizeof(double4) * 28;
  return b;
}

__device__ inline Smem carve_smem_tile(unsigned char *base, int N) {
  Smem s;
  size_t o = 0;
  s.pos = reinterpret_cast<double4 *>(base + o);
  o += align_up(sizeof(double4) * N, 128);
  s.red = reinterpret_cast<double *>(base + o);
  o += align_up(sizeof(double) * 8 * kMaxWarps, 128);
  s.rowid = base + o;
  o += align_up(static_cast<size_t>(N), 128);
  s.mbar = reinterpret_cast<uint64_t *>(base + o);
  o += 128;
  s.geom = reinterpret_cast<Geom *>(base + o);
  o += align_up(sizeof(Geom), 128);
  s.exp_tab = reinterpret_cast<double *>(base + o);
  s.shx = s.exp_tab + 128;
  s.shy = s.shx + 16;
  s.shz = s.shy + 16;
  s.sh27 = reinterpret_cast<double4 *>(s.shz + 16);
  s.sh27f = nullptr;
  s.dfd = nullptr;
  s.stage = nullptr;
  s.clu = nullptr;
  s.posf = nullptr;
  s.rho = nullptr;
  s.facc = nullptr;
  s.pxy = s.pzw = nullptr;
  s.t16 = nullptr;
  s.slots = nullptr;
  s.cell_of = s.wrapc = s.order = s.cell_start = s.cell_fill = nullptr;
  s.maxCells = 0;
  return s;
}

// Fused EAM kernel: positions, raw staging, reductions, rows, tables.  The cell-list scratch of the (rare)
// in-kernel rebuild lives in HBM (SysView::cellscratch) so that it does not cost shared memory on every launch.
__host__ __device__ inline size_t smem_bytes_fused(int N) {
  size_t b = 0;
  b += align_up(sizeof(double4) * (N + 1) + 32, 128);  // positions; the raw coordinates land in the tail and are converted in place
  b += align_up(sizeof(float4) * (N + 1), 128);
  b += align_up(sizeof(double) * 8 * kMaxWarps, 128);
  b += align_up(static_cast<size_t>(N), 128);
  b += 128;  // mbarrier
  b += 128;  // cluster partials
  b += align_up(sizeof(Geom), 128);
  b += sizeof(double) * (128 + 48) + sizeof(double4) * 28 + sizeof(float4) * 28;
  return b;
}

__host__ __device__ inline size_t fused_scratch_ints(int N, int maxCells) {
  const size_t Np = (static_cast<size_t>(N) + 31) / 32 * 32;
  return 3 * Np + 2 * ((static_cast<size_t>(maxCells) + 1 + 31) / 32 * 32);
}

__device__ inline Smem carve_smem_fused(unsigned char *base, int N, int maxCells, int *scratch) {
  Smem s;
  s.rho = nullptr;
  s.facc = nullptr;
  s.pxy = s.pzw = nullptr;
  s.t16 = nullptr;
  s.slots = nullptr;
  size_t o = 0;
  s.pos = reinterpret_cast<double4 *>(base + o);
  o += align_up(sizeof(double4) * (N + 1) + 32, 128);
  s.posf = reinterpret_cast<float4 *>(base + o);
  o += align_up(sizeof(float4) * (N + 1), 128);
  s.red = reinterpret_cast<double *>(base + o);
  o += align_up(sizeof(double) * 8 * kMaxWarps, 128);
  s.rowid = base + o;
  o += align_up(static_cast<size_t>(N), 128);
  s.mbar = reinterpret_cast<uint64_t *>(base + o);
  o += 128;
  s.clu = reinterpret_cast<double *>(base + o);
  o += 128;
  s.geom = reinterpret_cast<Geom *>(base + o);
  o += align_up(sizeof(Geom), 128);
  s.exp_tab = reinterpret_cast<double *>(base + o);
  s.shx = s.exp_tab + 128;
  s.shy = s.shx + 16;
  s.shz = s.shy + 16;
  s.sh27 = reinterpret_cast<double4 *>(s.shz + 16);
  s.sh27f = reinterpret_cast<float4 *>(s.sh27 + 28);
  s.dfd = nullptr;
  s.stage = nullptr;
  const size_t Np = (static_cast<size_t>(N) + 31) / 32 * 32;
  const size_t Cp = (static_cast<size_t>(maxCells) + 1 + 31) / 32 * 32;
  s.cell_of = scratch;
  s.wrapc = scratch ? scratch + Np : nullptr;
  s.order = scratch ? scratch + 2 * Np : nullptr;
  s.cell_start = scratch ? scratch + 3 * Np : nullptr;
  s.cell_fill = scratch ? scratch + 3 * Np + Cp : nullptr;
  s.maxCells = maxCells;
  return s;
}

// Newton's-third-law EAM kernel: the fused kernel's layout (its cold paths and its ordered-pair lane functions run
// unchanged) followed by the split position arrays, the exp table line and the receiver slots.
__host__ __device__ inline size_t smem_bytes_n3(int N, int slot_cap, int shift) {  // shift: 4 + log2(copies per atom)
  return align_up(smem_bytes_fused(N), 128) + 2 * align_up((static_cast<size_t>(N) + 1) << shift, 128) + 128 +
         align_up(sizeof(double) * static_cast<size_t>(slot_cap), 128);
}

__device__ inline Smem carve_smem_n3(unsigned char *base, int N, int maxCells, int *scratch, int shift) {
  Smem s = carve_smem_fused(base, N, maxCells, scratch);
  size_t o = align_up(smem_bytes_fused(N), 128);
  s.pxy = reinterpret_cast<double2 *>(base + o);
  o += align_up((static_cast<size_t>(N) + 1) << shift, 128);
  s.pzw = reinterpret_cast<double2 *>(base + o);
  o += align_up((static_cast<size_t>(N) + 1) << shift, 128);
  s.t16 = reinterpret_cast<double *>(base + o);
  o += 128;
  s.slots = reinterpret_cast<double *>(base + o);
  return s;
}

__device__ inline Smem carve_smem(unsigned char *base, int N, int maxCells, bool fast = false) {
  Smem s;
  size_t o = 0;
  s.pos = reinterpret_cast<double4 *>(base + o);
  o += align_up(sizeof(double4) * (N + 1), 128);
  s.posf = nullptr;
  s.sh27f = nullptr;
  if (fast) {
    s.posf = reinterpret_cast<float4 *>(base + o);
    o += align_up(sizeof(float4) * (N + 1), 128);
    s.sh27f = reinterpret_cast<float4 *>(base + o);
    o += sizeof(float4) * 28;
  }
  s.red = reinterpret_cast<double *>(base + o);
  o += align_up(sizeof(double) * 8 * kMaxWarps, 128);
  s.dfd = reinterpret_cast<double *>(base + o);
  o += align_up(sizeof(double) * N, 128);
  s.stage = reinterpret_cast<double *>(base + o);
  o += align_up(sizeof(double) * (3 * static_cast<size_t>(N) + 2), 128);
  s.cell_of = reinterpret_cast<int *>(base + o);
  o += align_up(sizeof(int) * N, 128);
  s.wrapc = reinterpret_cast<int *>(base + o);
  o += align_up(sizeof(int) * N, 128);
  s.order = reinterpret_cast<int *>(base + o);
  o += align_up(sizeof(int) * N, 128);
  s.cell_start = reinterpret_cast<int *>(base + o);
  o += align_up(sizeof(int) * (maxCells + 1), 128);
  s.cell_fill = reinterpret_cast<int *>(base + o);
  o += align_up(sizeof(int) * (maxCells + 1), 128);
  s.rowid = base + o;
  o += align_up(static_cast<size_t>(N), 128);
  s.mbar = reinterpret_cast<uint64_t *>(base + o);
  o += 128;
  s.geom = reinterpret_cast<Geom *>(base + o);
  o += align_up(sizeof(Geom), 128);
  s.exp_tab = reinterpret_cast<double *>(base + o);
  s.shx = s.exp_tab + 128;
  s.shy = s.shx + 16;
  s.shz = s.shy + 16;
  s.sh27 = reinterpret_cast<double4 *>(s.shz + 16);
  s.clu = nullptr;
  s.rho = nullptr;
  s.facc = nullptr;
  s.pxy = s.pzw = nullptr;
  s.t16 = nullptr;
  s.slots = nullptr;
  s.maxCells = maxCells;
  return s;
}

}  // namespace eonb200
