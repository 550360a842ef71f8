// eam_n3.cuh -- batched EAM energy + forces with every unordered pair evaluated ONCE (Newton's third law), without
// atomics and with a fixed summation order.
//
// Reference arithmetic: EAM::calc_force (client/potentials/EAM/EAM.cpp:125-226) with the embedding polynomial of
// EAM.cpp:349-365.  The reference evaluates the Morse pair part for j > i only (EAM.cpp:200) but the density part for
// every ordered pair; for a single-element system both are symmetric in (i, j) -- the predicate r^2 <= rc^2 is
// evaluated on d = r_i - r_j folded by +-L, whose square is bit-identical from either end (EAM.cpp:152-172) -- so one
// evaluation of the three exponentials serves both atoms.
//
// One CTA per walker, one lane per atom.  Every unordered pair of the cached list's sure-in and uncertain classes is
// OWNED by one of its two atoms (parity rule, so every atom owns about half of its row); the partner RECEIVES it
// through a slot of shared memory that only this pair writes (plain, non-blocking stores -- sm_100a's shared FP64
// atomics are compare-and-swap loops, measured 3-4x slower here and order-dependent).  Four passes:
//   density   owner: r, the three exponentials; g(r) into its own rho and into the partner's slot; g'(r)/r and the
//             Morse -phi'(r)/r go to the owner's scratch column (coalesced, L2 resident)
//   [barrier] rho_i = own part + its slots in slot order; F(rho_i), F'(rho_i) by Horner; F' into pos[i].w   [barrier]
//   force 1   owner: s = -(F'_i + F'_j) g'/r - phi'/r;  F_i += s d;  s into the partner's slot
//   [barrier]
//   force 2   receiver: F_i += s d' with its own d' = r_i - r_j (= -d), slots in slot order.
// Summation orders are fixed, so results are bit-reproducible.  The pair loops carry no branches and no bounds tests:
// rows are padded with ghost entries that fail the cutoff (owned rows) or have zero length (received rows), and pairs
// outside the cutoff are computed and discarded by selects.  Anything but the common case -- lists present and valid,
// every atom within skin/4 of its build position, one element, non-negative coordinates, slots that fit -- leaves the
// hot instruction stream through ONE out-of-line call (eam_n3_cold), which rebuilds the lists or finishes the walker
// with the ordered-pair / general functions of potentials.cuh.
#pragma once

#ifndef EONB200_N3_UD
#define EONB200_N3_UD 2
#endif
#ifndef EONB200_N3_UF
#define EONB200_N3_UF 4
#endif

namespace eonb200 {

constexpr int kN3UD = EONB200_N3_UD;   // pairs per iteration of the density loop
constexpr int kN3UF = EONB200_N3_UF;   // pairs per iteration of the force loops; owned / received rows are padded to this
constexpr int kN3Slack = 8;  // list rows a loop may request beyond a row's end (the allocations carry this many extra)

// List words of this kernel: bits n3_shift.. = j << n3_shift, the byte offset of atom j's first copy in the two position
// arrays (x, y) and (z, F') (j <= 2047; the ghost atom is j = N); bits 0, 1, 2 = "add this atom's box fold along x, y, z"
// -- an atom in the upper half of the box only ever meets images at -L, one in the lower half at +L (EAM.cpp:152-172
// with wrapped coordinates), so the sign is a property of the lane (N3Lane::sL*), not of the pair.
//
// The position arrays hold `rep` = 8 (4, 2, 1 when shared memory is short: SysView::n3_shift = 4 + log2 rep) copies of
// every atom, 16 bytes each, one per 16-byte bank group of shared memory (128 bytes per atom and array), and lane l
// always reads copy l & (rep - 1): the eight lanes a 128-bit shared load serves per wavefront then hit eight different
// bank groups WHATEVER atoms they ask for.  A gather over random atoms costs the minimum of four wavefronts; with one
// copy it cost 9 - 12 (ncu, profiles/r2_n3_*: 68 % of the kernel's shared-memory wavefronts were bank conflicts of
// these gathers, and the L1 data stage was its busiest unit).
constexpr int kN3FoldX = 1, kN3FoldY = 2, kN3FoldZ = 4;
constexpr int kN3MaxAtoms = 2047;
__device__ __forceinline__ int n3_code_atom(int code, int shift) { return static_cast<unsigned>(code) >> shift; }
// what N3H_VALID holds for lists built for this launch shape
__device__ __forceinline__ int n3_valid_tag(const SysView &sys) { return sys.n3_lpa | (sys.n3_shift << 8); }
// byte offset of this lane's copy inside an atom's 16 rep bytes
__device__ __forceinline__ int n3_lane_off(int shift) {
#ifdef EONB200_EMU
  (void)shift;
  return 0;
#else
  return (static_cast<int>(threadIdx.x) << 4) & ((1 << shift) - 1);
#endif
}
// all copies of atom i (copy order rotated by the lane, so that the lanes of a warp store into different bank groups)
__device__ __forceinline__ void n3_store_xy(double2 *pxy, int shift, int i, double x, double y) {
  if (EONB200_LIKELY(shift == 7)) {  // eight copies: unrolled
#pragma unroll
    for (int k = 0; k < 8; k++) pxy[(i << 3) + ((k + static_cast<int>(threadIdx.x)) & 7)] = make_double2(x, y);
    return;
  }
  const int rep = 1 << (shift - 4);
  for (int k = 0; k < rep; k++) pxy[(i << (shift - 4)) + ((k + static_cast<int>(threadIdx.x)) & (rep - 1))] = make_double2(x, y);
}
__device__ __forceinline__ void n3_store_zw(double2 *pzw, int shift, int i, double z, double w) {
  if (EONB200_LIKELY(shift == 7)) {
#pragma unroll
    for (int k = 0; k < 8; k++) pzw[(i << 3) + ((k + static_cast<int>(threadIdx.x)) & 7)] = make_double2(z, w);
    return;
  }
  const int rep = 1 << (shift - 4);
  for (int k = 0; k < rep; k++) pzw[(i << (shift - 4)) + ((k + static_cast<int>(threadIdx.x)) & (rep - 1))] = make_double2(z, w);
}
__device__ __forceinline__ void n3_store_w(double2 *pzw, int shift, int i, double w) {
  if (EONB200_LIKELY(shift == 7)) {
#pragma unroll
    for (int k = 0; k < 8; k++) pzw[(i << 3) + ((k + static_cast<int>(threadIdx.x)) & 7)].y = w;
    return;
  }
  const int rep = 1 << (shift - 4);
  for (int k = 0; k < rep; k++) pzw[(i << (shift - 4)) + ((k + static_cast<int>(threadIdx.x)) & (rep - 1))].y = w;
}

// hl_hdr words of a walker
enum { N3H_SLOTS = 0, N3H_VALID = 1, N3H_ROWS = 2, N3H_ROWZ = 3, N3H_CONFLICT = 4 };  // VALID: 0, or the lists' lanes per atom

// The atom that evaluates the unordered pair {i, j}: the one with the shorter row (surface atoms own all their pairs
// with bulk atoms, so that every atom owns about the same number); between atoms of equal row length the lower index
// when i + j is even, the higher when odd.
#ifndef EONB200_N3_OWNER
#define EONB200_N3_OWNER 1
#endif
__device__ __forceinline__ bool n3_owner(int i, int j, int deg_i, int deg_j) {
#if EONB200_N3_OWNER == 0
  if (deg_i != deg_j) return deg_i < deg_j;
#elif EONB200_N3_OWNER == 2
  if (deg_i + 8 < deg_j) return true;
  if (deg_j + 8 < deg_i) return false;
#endif
  return (i < j) == (((i + j) & 1) == 0);
}
__device__ __forceinline__ int n3_deg(const SysView &sys, int i) {
  const int c = sys.cls[i];
  return (c & 0xffff) + (c >> 16);
}

__device__ __forceinline__ int n3_round_up(int n, int m) { return (n + m - 1) / m * m; }

// Owner / receiver lists from the freshly built full rows (sys.nbr / cls).  CTA-wide; runs after a list rebuild only.
//
// Lanes do not take the atoms in index order: atoms are ranked by the number of pairs they own (descending, ties by
// index), and the atom of rank r is worked on by the `lpa` adjacent lanes (columns) r lpa .. r lpa + lpa - 1: its owned
// pairs and its received pairs are dealt out to them in turn, every column has its own run of slots, and the columns
// of an atom add their partial density / force with a warp shuffle.  Every list below is stored by COLUMN, so that
// the lanes of a warp read consecutive words; codes and slots refer to atom indices / slot numbers.
__device__ inline void n3_build(const SysView &sys, const Smem &sm, const PotParams &P, int lpa) {
  const int N = sys.N;
  const size_t step = static_cast<size_t>(sys.Np) * lpa;  // columns per list row
  const size_t Np = static_cast<size_t>(sys.Np);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int krows = sys.K / lpa;  // list rows a column has
  int *key = sys.cnt2;  // [Np] scratch: pairs owned by atom i
  if (tid == 0) {
    sys.hl_hdr[N3H_VALID] = 0;
    sys.hl_hdr[N3H_ROWS] = 0;
    sys.hl_hdr[N3H_CONFLICT] = 0;
  }
  for (int i = tid; i < N; i += nt) {
    const int n01 = n3_deg(sys, i);
    int no = 0;
    for (int k = 0; k < n01; k++) {
      const int j = sys.nbr[k * Np + i] & kNbrIndexMask;
      if (j != i && n3_owner(i, j, n01, n3_deg(sys, j))) no++;
    }
    key[i] = no;
  }
  __syncthreads();
  for (int i = tid; i < N; i += nt) {  // rank = atoms that own more pairs, or as many with a lower index
    const int ki = key[i];
    int rank = 0;
    for (int j = 0; j < N; j++) {
      const int kj = key[j];
      rank += (kj > ki || (kj == ki && j < i)) ? 1 : 0;
    }
    sys.lane_atom[rank] = i;
    sys.atom_lane[i] = rank;
  }
  __syncthreads();
  const int shift = sys.n3_shift;
  const int ghost = N << shift;
  int rows = 0;  // rows of the owner / receiver arrays in use
  for (int i = tid; i < N; i += nt) {
    const size_t col0 = static_cast<size_t>(sys.atom_lane[i]) * lpa;
    const int c = sys.cls[i];
    const int n0 = c & 0xffff, n01 = n0 + (c >> 16);
    int no_s = 0, no_u = 0, nr = 0, minus = 0, plus = 0;
    for (int k = 0; k < n01; k++) {
      const int code = sys.nbr[k * Np + i];
      const int j = code & kNbrIndexMask;
      if (j == i) continue;  // a self image only exists in boxes the fast path never sees
      const int id = static_cast<unsigned>(code) >> kNbrIndexBits;  // (fx + 1) + 3 (fy + 1) + 9 (fz + 1), pack_mic
      const int f[3] = {id % 3 - 1, (id / 3) % 3 - 1, id / 9 - 1};   // sh27: offset added to r_i - r_j is -f L
      int w = j << shift;
      for (int a = 0; a < 3; a++) {
        if (f[a] != 0) w |= kN3FoldX << a;
        if (f[a] > 0) minus |= 1 << a;
        if (f[a] < 0) plus |= 1 << a;
      }
      if (n3_owner(i, j, n01, n3_deg(sys, j))) {
        const int e = no_s + no_u;
        if (e / lpa < krows) sys.own_cs[(e / lpa) * step + col0 + e % lpa] = make_int2(w, 0);
        if (k < n0) no_s++; else no_u++;
      } else {
        if (nr / lpa < krows) sys.recv_code[(nr / lpa) * step + col0 + nr % lpa] = w;
        nr++;
      }
    }
    if (minus & plus) atomicOr(sys.hl_hdr + N3H_CONFLICT, 1);  // (cannot happen while the box holds two cutoffs)
    const int nown = no_s + no_u;
    for (int h = 0; h < lpa; h++) {
      const int nown_h = (nown - h + lpa - 1) / lpa, nso_h = (no_s - h + lpa - 1) / lpa, nr_h = (nr - h + lpa - 1) / lpa;
      const int nown_pad = n3_round_up(nown_h, kN3UF), nr_pad = n3_round_up(nr_h, kN3UF);
      for (int k = nown_h; k < nown_pad && k < krows; k++) {  // ghost pairs: beyond every cutoff, zero weights
        sys.own_cs[k * step + col0 + h] = make_int2(ghost, -1);
        sys.own_ab[k * step + col0 + h] = make_double2(0.0, 0.0);
      }
      for (int k = nr_h; k < nr_pad && k < krows; k++) sys.recv_code[k * step + col0 + h] = i << shift;  // zero-length pairs
      sys.lhdr[col0 + h] = make_int4(nso_h | (nown_h << 16), nr_h, 0, i | (minus << 24));
      sys.roff[col0 + h] = nr_h;
      rows = max(rows, max(nown_pad, nr_pad));
    }
  }
  __syncthreads();
  atomicMax(sys.hl_hdr + N3H_ROWS, rows);
  // exclusive scan of the received counts (column order) -> first slot of every column
  const int C = N * lpa;
#ifdef EONB200_EMU
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < C; t++) {
      const int v = sys.roff[t];
      sys.roff[t] = run;
      run += v;
    }
    sys.roff[C] = run;
  }
#else
  if (tid < 32) {
    int carry = 0;
    for (int base = 0; base < C; base += 32) {
      const int t = base + tid;
      const int v = t < C ? sys.roff[t] : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (tid >= o) x += y;
      }
      if (t < C) sys.roff[t] = carry + x - v;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (tid == 0) sys.roff[C] = carry;
  }
#endif
  __syncthreads();
  // every owned pair learns the slot its partner reads it from (as a byte offset into the slot buffer); ghost pairs
  // write the spare slot behind the last one
  const int total = sys.roff[C];
  for (int col = tid; col < C; col += nt) {
    const int i = sys.lane_atom[col / lpa];
    int4 h = sys.lhdr[col];
    h.z = sys.roff[col];
    sys.lhdr[col] = h;
    const int nown_h = h.x >> 16;
    const int nown_pad = n3_round_up(nown_h, kN3UF);
    for (int k = 0; k < nown_pad && k < krows; k++) {
      const int code = sys.own_cs[k * step + col].x;
      int slot = total;
      if (k < nown_h) {
        const size_t cj = static_cast<size_t>(sys.atom_lane[n3_code_atom(code, shift)]) * lpa;
        int nrj = 0;
        for (int q = 0; q < lpa; q++) nrj += sys.lhdr[cj + q].y;
        int r = 0;
        while (r < nrj && n3_code_atom(sys.recv_code[(r / lpa) * step + cj + r % lpa], shift) != i) r++;
        if (r >= nrj) r = 0;
        slot = sys.roff[cj + r % lpa] + r / lpa;
      }
      sys.own_cs[k * step + col] = make_int2(code, slot * 8);
    }
  }
  if (tid == 0) {
    int r0 = 255;
    const int z0 = sys.Z[0];
    for (int q = P.n_rows - 1; q >= 0; q--)
      if (P.rows[q].Z == z0) r0 = q;
    for (int k = 0; k < 3; k++) sys.lo_build[k] = sm.geom->lo[k];  // (prepare_positions of the rebuild just before)
    sys.hl_hdr[N3H_SLOTS] = total + kN3UF;  // slots the fast path needs (spare slots for the ghost pairs / padded reads)
    sys.hl_hdr[N3H_ROWZ] = r0 | (z0 << 8);
    atomicMax(&sys.stats[3], static_cast<unsigned long long>(total + kN3UF));
  }
  __syncthreads();
  if (tid == 0 && ld_cg(sys.hl_hdr + N3H_ROWS) <= krows) sys.hl_hdr[N3H_VALID] = n3_valid_tag(sys);
  __syncthreads();
}

struct N3Lane {  // what one lane carries through the passes for one of its atoms
  double rho_own, dfd, fx, fy, fz;
  double sLx, sLy, sLz;  // this atom's box fold per axis (+L or -L)
  int atom;              // the atom this lane works for (lanes take the atoms by rank, see n3_build)
  int nso, nown, nrecv;  // owned sure-in pairs, owned pairs, received pairs
  int rbase;             // its first slot
  int off, mask;         // n3_lane_off(): the copy of the position arrays this lane reads; mask of a code's offset bits
};

struct N3Box {
  double Lx, Ly, Lz;
};

__device__ __forceinline__ void n3_unpack(const int4 &h, const N3Box &bx, int shift, N3Lane &L) {
  L.nso = h.x & 0xffff;
  L.nown = static_cast<unsigned>(h.x) >> 16;
  L.nrecv = h.y;
  L.rbase = h.z;
  L.off = n3_lane_off(shift);
  L.mask = ~((1 << shift) - 1);
  L.atom = h.w & 0xffffff;
  L.sLx = (h.w & (1 << 24)) ? -bx.Lx : bx.Lx;
  L.sLy = (h.w & (1 << 25)) ? -bx.Ly : bx.Ly;
  L.sLz = (h.w & (1 << 26)) ? -bx.Lz : bx.Lz;
}

// exp with a 16-entry table that fills exactly one 128-byte line of shared memory (a warp's lookups never conflict):
// x = (16 k + j) ln2/16 + r, |r| <= ln2/32, e^x = 2^k T[j] (1 + r + ... + r^6/720).  Truncation r^7/5040 < 4.5e-16 and
// the one-step reduction (|k| <= ~1200 times the rounding of ln2/16) < 5e-15 relative: five orders below the parity
// bar of 1e-10.  No underflow guard (PotParams::eam_fast_ok).
__device__ __forceinline__ double exp_t16(double x, const double *T) {
#ifdef EONB200_ABL_EXP  // timing ablation (never defined in the product build): no exponential
  return fma(x, 1e-3, T[0]);
#endif
  const double kInv = 23.083120654223414;    // 16 / ln 2
  const double kShift = 6755399441055744.0;  // 1.5 * 2^52
  const double kLn2_16 = 6.93147180559945309417e-01 / 16.0;
  double kd = fma(x, kInv, kShift);
#ifdef EONB200_EMU
  long long ki;
  memcpy(&ki, &kd, 8);
  const int k = static_cast<int>(ki);
#else
  const int k = __double2loint(kd);
#endif
  kd -= kShift;
  const double r = fma(kd, -kLn2_16, x);
  const double t = T[k & 15];
  double q = fma(r, 1.0 / 720.0, 1.0 / 120.0);
  q = fma(q, r, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double res = fma(t, q * r, t);
#ifdef EONB200_EMU
  const long long e = static_cast<long long>(k >> 4) << 52;
  long long bits;
  memcpy(&bits, &res, 8);
  bits += e;
  double out;
  memcpy(&out, &bits, 8);
  return out;
#else
  return __hiloint2double(__double2hiint(res) + ((k << 16) & 0xfff00000), __double2loint(res));
#endif
}

// d = r_i - r_j folded by the image recorded at build time.  pxyb / pzwb: the position arrays as bytes.
__device__ __forceinline__ void n3_vec(const unsigned char *pxyb, const unsigned char *pzwb, int c, const N3Lane &L,
                                       double xi, double yi, double zi, double &dx, double &dy, double &dz, double &wj) {
#ifdef EONB200_ABL_GATHER
  const int o = ((threadIdx.x & 255) << 4) + 0 * c;
#else
  const int o = (c & L.mask) | L.off;
#endif
  const double2 jxy = *reinterpret_cast<const double2 *>(pxyb + o);
  const double2 jzw = *reinterpret_cast<const double2 *>(pzwb + o);
  dx = xi - jxy.x;
  dy = yi - jxy.y;
  dz = zi - jzw.x;
  if (c & kN3FoldX) dx += L.sLx;
  if (c & kN3FoldY) dy += L.sLy;
  if (c & kN3FoldZ) dz += L.sLz;
  wj = jzw.y;
}
__device__ __forceinline__ void n3_vec3(const unsigned char *pxyb, const unsigned char *pzwb, int c, const N3Lane &L,
                                        double xi, double yi, double zi, double &dx, double &dy, double &dz) {
#ifdef EONB200_ABL_GATHER  // timing ablation: conflict-free, lane-consecutive "gather"
  const int o = ((threadIdx.x & 255) << 4) + 0 * c;
#else
  const int o = (c & L.mask) | L.off;
#endif
  const double2 jxy = *reinterpret_cast<const double2 *>(pxyb + o);
  const double jz = *reinterpret_cast<const double *>(pzwb + o);
  dx = xi - jxy.x;
  dy = yi - jxy.y;
  dz = zi - jz;
  if (c & kN3FoldX) dx += L.sLx;
  if (c & kN3FoldY) dy += L.sLy;
  if (c & kN3FoldZ) dz += L.sLz;
}

struct N3K {  // pair constants of the walker's element row
  double rc2, nb1, nb2x2, nb2x1024, naM, aMRm, Dm, nK2;
};

// Owner density pass of lane position t.
__device__ __forceinline__ void n3_density(const N3K &K, const SysView &sys, const Smem &sm, int t, N3Lane &L,
                                           int2 (&cur)[kN3UD], double (&acc)[2]) {
  const double *T = sm.t16;
  const unsigned char *pxyb = reinterpret_cast<const unsigned char *>(sm.pxy);
  const unsigned char *pzwb = reinterpret_cast<const unsigned char *>(sm.pzw);
  unsigned char *slotb = reinterpret_cast<unsigned char *>(sm.slots);
  const size_t step = static_cast<size_t>(sys.Np) * sys.n3_lpa;
  const int2 *oc = sys.own_cs + t + kN3UD * step;
  double2 *oab = sys.own_ab + t;
  const int own = (L.atom << sys.n3_shift) | L.off;
  const double2 pixy = *reinterpret_cast<const double2 *>(pxyb + own);
  const double piz = *reinterpret_cast<const double *>(pzwb + own);
  double dens = 0.0, e_acc = 0.0;
  int npairs = 0;
  const int nown = L.nown, nso = L.nso;
  for (int k = 0; k < nown; k += kN3UD) {
    int2 cs[kN3UD];
#pragma unroll
    for (int u = 0; u < kN3UD; u++) {
      cs[u] = cur[u];
      cur[u] = ld_cg(oc + u * step);  // (rows beyond the end are valid memory: kN3Slack)
    }
    oc += kN3UD * step;
#pragma unroll
    for (int u = 0; u < kN3UD; u++) {
      const int c = cs[u].x;
      double dx, dy, dz;
      n3_vec3(pxyb, pzwb, c, L, pixy.x, pixy.y, piz, dx, dy, dz);
      const double r2 = norm2_rn(dx, dy, dz);
      // EAM.cpp:172 decides for the uncertain class; sure-in pairs are inside; ghost pairs are far outside
      const bool in = (k + u < nso) || !(r2 > K.rc2);
      const double invr = rsqrt_fast(r2);
      const double r = r2 * invr;
      const double r4 = r2 * r2;
      const double e1 = exp_t16(K.nb1 * r, T);
      const double e2 = exp_t16(K.nb2x2 * r, T);
      const double ex = exp_t16(fma(K.naM, r, K.aMRm), T);
      const double rho_pair = fma(512.0, e2, e1);
      double g = (r4 * r2) * rho_pair;  // EAM.cpp:178-183
      const double uu = fma(K.nb2x1024, e2, K.nb1 * e1);
      // g'(r)/r = [6 r^5 rho_pair + r^6 u] / r = r^4 (6 rho_pair + r u)   (EAM.cpp:186-198)
      double A = r4 * fma(r, uu, 6.0 * rho_pair);
      const double dm = 1.0 - ex;
      double nB = (K.nK2 * dm) * (ex * invr);  // -(+-phi'(r)/r), EAM.cpp:200-217
      double em = K.Dm * fma(dm, dm, -1.0);     // Dm (d^2 - 1): once per unordered pair
      g = in ? g : 0.0;
      A = in ? A : 0.0;
      nB = in ? nB : 0.0;
      em = in ? em : 0.0;
      npairs += in ? 1 : 0;
      dens += g;
      e_acc += em;
      *reinterpret_cast<double *>(slotb + cs[u].y) = g;
      oab[u * step] = make_double2(A, nB);
    }
    oab += kN3UD * step;
  }
  L.rho_own = dens;
  acc[0] += e_acc;
  acc[1] += static_cast<double>(npairs);
}

// This column's part of rho_i: own part + its received slots (fixed order: four interleaved partial sums).
__device__ __forceinline__ double n3_rho_part(const Smem &sm, const N3Lane &L) {
  const double *sl = sm.slots + L.rbase;
  const int nr = L.nrecv;
  double s0 = L.rho_own, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int r = 0;
  for (; r + 4 <= nr; r += 4) {
    s0 += sl[r];
    s1 += sl[r + 1];
    s2 += sl[r + 2];
    s3 += sl[r + 3];
  }
  for (; r < nr; r++) s0 += sl[r];
  return (s0 + s1) + (s2 + s3);
}
// The value a pair of adjacent lanes (the two columns of an atom) add up to; both lanes get it.
__device__ __forceinline__ double n3_pair_sum(double v) {
#ifndef EONB200_EMU
  return v + __shfl_xor_sync(0xffffffffu, v, 1);
#else
  return v;
#endif
}

// Owner force pass.  Pairs outside the cutoff and ghost pairs carry zero weights (written by the density pass / the
// list build).
__device__ __forceinline__ void n3_own_prefetch(const SysView &sys, int t, int2 (&cq)[kN3UF], double2 (&wq)[kN3UF]) {
  const size_t step = static_cast<size_t>(sys.Np) * sys.n3_lpa;
#pragma unroll
  for (int u = 0; u < kN3UF; u++) {
    cq[u] = ld_cg(sys.own_cs + t + u * step);
    wq[u] = ld_cg(sys.own_ab + t + u * step);
  }
}
__device__ __forceinline__ void n3_force_own(const SysView &sys, const Smem &sm, int t, N3Lane &L, int2 (&cq)[kN3UF],
                                             double2 (&wq)[kN3UF]) {
  const unsigned char *pxyb = reinterpret_cast<const unsigned char *>(sm.pxy);
  const unsigned char *pzwb = reinterpret_cast<const unsigned char *>(sm.pzw);
  unsigned char *slotb = reinterpret_cast<unsigned char *>(sm.slots);
  const size_t step = static_cast<size_t>(sys.Np) * sys.n3_lpa;
  const int2 *oc = sys.own_cs + t + kN3UF * step;
  const double2 *oab = sys.own_ab + t + kN3UF * step;
  const int own = (L.atom << sys.n3_shift) | L.off;
  const double2 pixy = *reinterpret_cast<const double2 *>(pxyb + own);
  const double piz = *reinterpret_cast<const double *>(pzwb + own);
  const double ndfi = -L.dfd;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  const int nown = L.nown;
  for (int k0 = 0; k0 < nown; k0 += kN3UF) {
    int2 cs[kN3UF];
    double2 w[kN3UF];
#pragma unroll
    for (int u = 0; u < kN3UF; u++) {
      cs[u] = cq[u];
      w[u] = wq[u];
      cq[u] = ld_cg(oc + u * step);
      wq[u] = ld_cg(oab + u * step);
    }
    oc += kN3UF * step;
    oab += kN3UF * step;
#pragma unroll
    for (int u = 0; u < kN3UF; u++) {
      double dx, dy, dz, dfj;
      n3_vec(pxyb, pzwb, cs[u].x, L, pixy.x, pixy.y, piz, dx, dy, dz, dfj);
      const double sc = fma(ndfi - dfj, w[u].x, w[u].y);
      fx = fma(sc, dx, fx);
      fy = fma(sc, dy, fy);
      fz = fma(sc, dz, fz);
      *reinterpret_cast<double *>(slotb + cs[u].y) = sc;
    }
  }
  L.fx = fx;
  L.fy = fy;
  L.fz = fz;
}

// Receiver force pass: F_i += s d' over the atom's slots in slot order.  Padded entries have d' = 0.
__device__ __forceinline__ void n3_recv_prefetch(const SysView &sys, int t, int (&rq)[kN3UF]) {
  const size_t step = static_cast<size_t>(sys.Np) * sys.n3_lpa;
#pragma unroll
  for (int u = 0; u < kN3UF; u++) rq[u] = ld_cg(sys.recv_code + t + u * step);
}
__device__ __forceinline__ void n3_force_recv(const SysView &sys, const Smem &sm, int t, N3Lane &L, int (&rq)[kN3UF]) {
  const unsigned char *pxyb = reinterpret_cast<const unsigned char *>(sm.pxy);
  const unsigned char *pzwb = reinterpret_cast<const unsigned char *>(sm.pzw);
  const size_t step = static_cast<size_t>(sys.Np) * sys.n3_lpa;
  const int *rc = sys.recv_code + t + kN3UF * step;
  const double *sl = sm.slots + L.rbase;
  const int own = (L.atom << sys.n3_shift) | L.off;
  const double2 pixy = *reinterpret_cast<const double2 *>(pxyb + own);
  const double piz = *reinterpret_cast<const double *>(pzwb + own);
  double fx = L.fx, fy = L.fy, fz = L.fz;
  const int nr = L.nrecv;
  for (int k0 = 0; k0 < nr; k0 += kN3UF) {
    int c[kN3UF];
#pragma unroll
    for (int u = 0; u < kN3UF; u++) {
      c[u] = rq[u];
      rq[u] = ld_cg(rc + u * step);
    }
    rc += kN3UF * step;
#pragma unroll
    for (int u = 0; u < kN3UF; u++) {
      double dx, dy, dz;
      n3_vec3(pxyb, pzwb, c[u], L, pixy.x, pixy.y, piz, dx, dy, dz);
      const double sc = sl[k0 + u];
      fx = fma(sc, dx, fx);
      fy = fma(sc, dy, fy);
      fz = fma(sc, dz, fz);
    }
  }
  L.fx = fx;
  L.fy = fy;
  L.fz = fz;
}

// Sum of two per-thread values over the CTA; thread 0 gets the result.  One barrier.
__device__ __forceinline__ void cta_sum2_to0(double (&v)[2], double *scratch, long long *dbg = nullptr) {
#ifndef EONB200_EMU
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  // The lanes of a warp leave the pair loop before this call at different iterations; without a reconvergence point
  // the shuffles below take their divergent-warp path (WARPSYNC.COLLECTIVE per shuffle: 5 of 13 warps did, ncu: 2.6 us
  // per walker).  A __syncwarp() here is not enough -- ptxas sinks it below the full-mask shuffles in some builds -- so
  // the CTA meets at a barrier first (the warps would wait at the one below anyway).
  __syncthreads();
  v[0] = warp_sum(v[0]);
  v[1] = warp_sum(v[1]);
  if (lane == 0) {
    scratch[warp] = v[0];
    scratch[kMaxWarps + warp] = v[1];
  }
#ifdef EONB200_DBG_RED
  phase_clock_last(dbg, 6);
#endif
  __syncthreads();
#ifdef EONB200_DBG_RED
  phase_clock(dbg, 7);
#endif
  if (warp == 0) {
    v[0] = warp_sum(lane < nwarps ? scratch[lane] : 0.0);
    v[1] = warp_sum(lane < nwarps ? scratch[kMaxWarps + lane] : 0.0);
  }
#else
  (void)scratch;
  (void)dbg;
#endif
}

// One list column per thread (SINGLE: the lane state stays in registers between the passes, the first words of the
// force passes are requested while the CTA meets at the barriers; one or two columns per atom) or several (one column
// per atom; state through HBM scratch: sys.dfd holds the own density part, the output rows the own force part).
template <bool SINGLE>
__device__ __forceinline__ void eam_n3_fast(const N3K &K, const double *c, const double *dc, const SysView &sys,
                                            const Smem &sm, const N3Box &bx, double *F, double *U, const int4 &h0,
                                            int2 (&cd)[kN3UD], long long *clk) {
  const int lpa = SINGLE ? sys.n3_lpa : 1;
  const int C = sys.N * lpa;  // list columns
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t step = static_cast<size_t>(sys.Np) * sys.n3_lpa;
  const bool first = lpa == 1 || (tid & 1) == 0;  // the column that speaks for the atom
  double acc[2] = {0.0, 0.0};
  N3Lane L;
  int2 cq[kN3UF];
  double2 wq[kN3UF];
  int rq[kN3UF];
  n3_unpack(h0, bx, sys.n3_shift, L);
  L.rho_own = 0.0;
  // staging row of the outgoing forces (the cold paths' position block, idle here): element `fhead` is 16-byte aligned
  // in shared memory exactly when it is in the caller's row, as a bulk copy wants
  const int fhead = (reinterpret_cast<uintptr_t>(F) & 15u) ? 1 : 0;
  double *fstage = reinterpret_cast<double *>(sm.pos) + fhead;
  if (SINGLE) {
    if (tid < C) n3_density(K, sys, sm, tid, L, cd, acc);
  } else {
    for (int t = tid; t < C; t += nt) {
      if (t != tid) {
        n3_unpack(ld_cg(sys.lhdr + t), bx, sys.n3_shift, L);
#pragma unroll
        for (int u = 0; u < kN3UD; u++) cd[u] = ld_cg(sys.own_cs + t + u * step);
      }
      n3_density(K, sys, sm, t, L, cd, acc);
      sys.dfd[t] = L.rho_own;
    }
  }
  phase_clock_last(clk, 5);
  if (SINGLE && tid < C) {  // first words of the two force passes
    n3_own_prefetch(sys, tid, cq, wq);
    n3_recv_prefetch(sys, tid, rq);
  }
#ifndef EONB200_DBG_RED
  phase_clock(clk, 6);
#endif
  __syncthreads();
#ifndef EONB200_DBG_RED
  phase_clock(clk, 7);
#endif
  if (SINGLE) {
    double part = tid < C ? n3_rho_part(sm, L) : 0.0;
    __syncwarp();  // (rows of different lengths: meet again before the two polynomials)
    if (lpa == 2) part = n3_pair_sum(part);
    if (tid < C) {
      L.dfd = eam_dembed(dc, part);
      if (first) {
        acc[0] += eam_embed(c, part);
        n3_store_w(sm.pzw, sys.n3_shift, L.atom, L.dfd);
      }
    }
  } else {
    for (int t = tid; t < C; t += nt) {
      n3_unpack(ld_cg(sys.lhdr + t), bx, sys.n3_shift, L);
      L.rho_own = sys.dfd[t];
      const double dens = n3_rho_part(sm, L);
      acc[0] += eam_embed(c, dens);
      n3_store_w(sm.pzw, sys.n3_shift, L.atom, eam_dembed(dc, dens));
    }
  }
  __syncthreads();
  phase_clock(clk, 8);
  for (int t = tid; t < C; t += nt) {
    if (!SINGLE) {
      n3_unpack(ld_cg(sys.lhdr + t), bx, sys.n3_shift, L);
      L.dfd = sm.pzw[L.atom << (sys.n3_shift - 4)].y;
      n3_own_prefetch(sys, t, cq, wq);
    }
    n3_force_own(sys, sm, t, L, cq, wq);
    if (!SINGLE) {
      F[3 * L.atom] = L.fx;
      F[3 * L.atom + 1] = L.fy;
      F[3 * L.atom + 2] = L.fz;
    }
  }
  phase_clock_last(clk, 11);
  __syncthreads();
  phase_clock(clk, 9);
  if (SINGLE) {
    if (tid < C)
      n3_force_recv(sys, sm, tid, L, rq);
    else
      L.fx = L.fy = L.fz = 0.0;
    if (lpa == 2) {
      L.fx = n3_pair_sum(L.fx);
      L.fy = n3_pair_sum(L.fy);
      L.fz = n3_pair_sum(L.fz);
    }
    // Forces leave through shared memory: lanes hold atoms in rank order, so direct stores would be 3 N scattered
    // 8-byte writes; one bulk copy moves the walker's row instead (issued below, after the energy reduction's barrier).
    if (tid < C && first) {
      fstage[3 * L.atom] = L.fx;
      fstage[3 * L.atom + 1] = L.fy;
      fstage[3 * L.atom + 2] = L.fz;
    }
  } else {
    for (int t = tid; t < C; t += nt) {
      n3_unpack(ld_cg(sys.lhdr + t), bx, sys.n3_shift, L);
      L.fx = F[3 * L.atom];
      L.fy = F[3 * L.atom + 1];
      L.fz = F[3 * L.atom + 2];
      n3_recv_prefetch(sys, t, rq);
      n3_force_recv(sys, sm, t, L, rq);
      F[3 * L.atom] = L.fx;
      F[3 * L.atom + 1] = L.fy;
      F[3 * L.atom + 2] = L.fz;
    }
  }
  phase_clock(clk, 15);
  phase_clock_last(clk, 13);
  cta_sum2_to0(acc, sm.red, clk);  // (its barrier also publishes fstage)
  phase_clock(clk, 14);
  if (SINGLE && tid == 32 % nt) {
    const int D = 3 * sys.N;
    const int body = (D - fhead) & ~1;
    if (body > 0) bulk_s2g(F + fhead, fstage + fhead, static_cast<uint32_t>(body) * 8u);
    if (fhead) F[0] = fstage[0];
    if (fhead + body < D) F[D - 1] = fstage[D - 1];
  }
  if (tid == 0) {
    *U = acc[0];
    atomicAdd(&sys.stats[2], 2ULL * static_cast<unsigned long long>(acc[1]));  // ordered in-cutoff pairs
  }
}

__device__ __forceinline__ void n3_cold_views(const FusedCold &fc, SysView &sys, Smem &sm) {
  EONB200_DYN_SMEM(smem_raw);
  sys = batch_sys(fc.nl, fc.a, blockIdx.y);
  sm = carve_smem_n3(smem_raw, fc.a.N, fc.a.maxCells, sys.cellscratch, sys.n3_shift);
}

// The ordered-pair lane functions (filter -> density -> force over the walker's full rows, potentials.cuh) inside this
// kernel: walkers beyond skin/4 of their build positions, or whose slots do not fit.
__device__ EONB200_NOINLINE void eam_n3_ordered(const FusedCold *fc, int r0, bool near_build) {
  SysView sys;
  Smem sm;
  fused_cold_views(*fc, sys, sm);
  const PotParams &P = *fc->Pp;
  const int N = sys.N, tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < N; i += nt) {
    const double4 p = sm.pos[i];
    sm.pos[i].w = 0.0;
    sm.posf[i] = make_float4(static_cast<float>(p.x), static_cast<float>(p.y), static_cast<float>(p.z), 0.0f);
  }
  if (tid == 0) {  // the ghost atom the row padding points at: far outside every cutoff
    sm.pos[N] = make_double4(1e150, 1e150, 1e150, 0.0);
    sm.posf[N] = make_float4(1e18f, 1e18f, 1e18f, 0.0f);
  }
  __syncthreads();
  FusedFast ff;
  ff.row = &P.rows[r0];
  ff.rc2 = ff.row->rc2;
  ff.rc2_hi = static_cast<float>(ff.rc2) + eam_filter_band(fmax(fmax(sys.box[0], sys.box[4]), sys.box[8]), ff.row->r_cut);
  ff.near_build = near_build;
  int n0_first = 0, kend_first = 0;
  int cur[kFilterBatch];
  if (tid < N) {
    eam_filter_range(sys, tid, near_build, n0_first, kend_first);
    if (n0_first < kend_first) {
#pragma unroll
      for (int u = 0; u < kFilterBatch; u++)
        cur[u] = ld_cg(sys.nbr + static_cast<size_t>((n0_first & ~(kFilterBatch - 1)) + u) * sys.Np + tid);
    }
  }
  double *F = batch_F(fc->a, blockIdx.y);
  double *U = fc->a.U + blockIdx.y;
  if (N <= nt)
    eam_fused_fast<true>(P, ff, sys, sm, 0, N, 1u, 0u, F, U, n0_first, kend_first, cur, nullptr);
  else
    eam_fused_fast<false>(P, ff, sys, sm, 0, N, 1u, 0u, F, U, n0_first, kend_first, cur, nullptr);
}

// Everything but the common case.  disp2_thread: this thread's largest squared displacement from the build positions
// (meaningless without a list).  Returns 1 when the walker is (now) fit for the fast path -- lists valid for the
// positions in sm.pos -- and 0 when the walker has been finished here.
__device__ EONB200_NOINLINE int eam_n3_cold(FusedCold *fc, const NlBatch *nl, const ForceBatchArgs *a, const PotParams *Pp,
                                            double disp2_thread) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    fc->nl = *nl;
    fc->a = *a;
    fc->Pp = Pp;
  }
  __syncthreads();
  SysView sys;
  Smem sm;
  n3_cold_views(*fc, sys, sm);
  const PotParams &P = *Pp;
  const int N = sys.N;
  for (int i = tid; i < N; i += nt) {
    const int i0 = i << (sys.n3_shift - 4);
    sm.pos[i] = make_double4(sm.pxy[i0].x, sm.pxy[i0].y, sm.pzw[i0].x, 0.0);
  }
  for (int k = tid; k < 128; k += nt) sm.exp_tab[k] = P.exp_tab[k];
  const double Lbox[3] = {sys.box[0], sys.box[4], sys.box[8]};
  fill_sh27(sm, Lbox[0], Lbox[1], Lbox[2]);
  const bool have = nl_have_list(sys);
  const int z0 = sys.Z[0];
  double mixed_t = 0.0;
  for (int i = tid; i < N; i += nt)
    if (sys.Z[i] != z0) mixed_t = 1.0;
  double red[2] = {disp2_thread, mixed_t};
  {
    const int op[2] = {RED_MAX, RED_MAX};
    block_reduce<2>(red, op, sm.red);
  }
  const bool mixed = red[1] != 0.0;
  double disp2 = red[0];
  const double need = 2.0 * P.rc_list + P.skin;  // stored folds stay the reference's minimum image
  const bool tab_ok = Lbox[0] >= need && Lbox[1] >= need && Lbox[2] >= need;
  // small (or open) boxes: the hot path's nearest-image positions mean nothing there; EAM.cpp:56-90 as written
  if (!tab_ok) disp2 = eam_fused_reconvert(fc, have);
  const double lim = 0.5 * P.skin, dstar = 0.25 * P.skin;
  const int h_valid = ld_cg(sys.hl_hdr + N3H_VALID);
  const int h_rowz = ld_cg(sys.hl_hdr + N3H_ROWZ);
  const bool rebuild = fc->a.force_rebuild != 0 || !have || P.skin <= 0.0 || !(disp2 <= lim * lim) ||
                       h_valid != n3_valid_tag(sys) || (h_rowz >> 8) != z0;
  const bool near_build = rebuild || disp2 <= dstar * dstar;  // a fresh list is built from these very positions
  if (rebuild) {
    __syncthreads();
    if (tid == 0) sys.hl_hdr[N3H_VALID] = 0;
    eam_fused_rebuild(fc, Pp);
    n3_build(sys, sm, P, sys.n3_lpa);
  }
  if (mixed || !tab_ok || !P.eam_fast_ok) {  // general rows: full element table, geometry for mic_vec
    if (tid == 0) sm.pos[N] = make_double4(1e150, 1e150, 1e150, 0.0);
    __syncthreads();
    eam_fused_general(fc, tab_ok, 0, N);
    return 0;
  }
  int r0 = -1;
  for (int q = P.n_rows - 1; q >= 0; q--)
    if (P.rows[q].Z == z0) r0 = q;
  double *F = batch_F(fc->a, blockIdx.y);
  if (r0 < 0) {  // element without a row: the reference throws 14324 (EAM.cpp:116-123)
    for (int k = tid; k < 3 * N; k += nt) F[k] = 0.0;
    if (tid == 0) fc->a.U[blockIdx.y] = nan("");
    return 0;
  }
  const bool fits = ld_cg(sys.hl_hdr + N3H_VALID) == n3_valid_tag(sys) && ld_cg(sys.hl_hdr + N3H_SLOTS) <= sys.slot_cap &&
                    ld_cg(sys.hl_hdr + N3H_CONFLICT) == 0 && N <= kN3MaxAtoms;
  if (!near_build || !fits) {
    eam_n3_ordered(fc, r0, near_build);
    return 0;
  }
  return 1;
}

// grid (1, S): one CTA per walker.  Pg: the parameter block as a kernel argument (constant bank: no load latency on
// the hot path); Pp: the same block in HBM for the out-of-line cold paths.
__device__ __forceinline__ void eam_n3_body(const PotParams &P, const PotParams *__restrict__ Pp, const NlBatch &nl,
                                            const ForceBatchArgs &a) {
  EONB200_DYN_SMEM(smem_raw);
  const int s = blockIdx.y;
  __shared__ FusedCold cold;  // filled only when the cold path is about to run
  __shared__ double coef[18];  // embedding polynomial and its derivative (EamRowDev::c, dc) of the walker's element
  __shared__ double invL[3];   // 1 / box length per axis
  const SysView sys = batch_sys(nl, a, s);
  const Smem sm = carve_smem_n3(smem_raw, a.N, a.maxCells, sys.cellscratch, sys.n3_shift);
  const int N = a.N, tid = threadIdx.x, nt = blockDim.x;
  const size_t step = static_cast<size_t>(sys.Np) * sys.n3_lpa;
  long long *clk = a.clocks ? a.clocks + static_cast<size_t>(s) * 16 : nullptr;
#ifndef EONB200_EMU
  if (clk && tid == 0) {
    unsigned long long gt;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    clk[0] = static_cast<long long>(gt);
    clk[1] = smid;
    clk[2] = clock64();
  }
#endif
  // ---- every HBM access of the prologue in flight at once ------------------------------------------------------------
  // The walker's raw coordinates land in the slot buffer (free until the density pass).  A bulk copy needs 16-byte
  // aligned ends on both sides: rows of an odd 3N start on odd multiples of 8 bytes for every second walker, so the copy
  // moves the aligned interior and up to one leading and one trailing coordinate are fetched by plain loads.
  const int D = 3 * N;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(sys.R);
  const bool aligned = (addr & 7u) == 0;
  const int head = (addr & 15u) ? 1 : 0;
  const int body = (D - head) & ~1;
  double *raw = sm.slots + head;  // &raw[head] is 16-byte aligned
  if (tid == 0) {
    mbar_init(sm.mbar, 1);
    if (aligned) {
      mbar_expect(sm.mbar, static_cast<uint32_t>(body) * 8u);
      bulk_g2s_issue(raw + head, sys.R + head, static_cast<uint32_t>(body) * 8u, sm.mbar);
    }
  }
  if (aligned && tid == 32 % nt && head) raw[0] = sys.R[0];
  if (aligned && tid == 33 % nt && head + body < D) raw[D - 1] = sys.R[D - 1];
  // list header, this lane's header and the first batch of its owned pairs (speculative: lists that turn out stale are
  // read again after the rebuild).  Nothing is USED before the positions have landed.
  const int4 hdr = ld_cg(reinterpret_cast<const int4 *>(sys.hl_hdr));
  int4 lh = make_int4(0, 0, 0, 0);
  int2 cd[kN3UD];
#pragma unroll
  for (int u = 0; u < kN3UD; u++) cd[u] = make_int2(0, 0);
  const int C = N * sys.n3_lpa;  // list columns
  if (tid < C) {
    lh = ld_cg(sys.lhdr + tid);
#pragma unroll
    for (int u = 0; u < kN3UD; u++) cd[u] = ld_cg(sys.own_cs + tid + u * step);
  }
#ifndef EONB200_EMU
  // The pair loops read one list row per iteration and request it only one iteration ahead: a row that has to come
  // from HBM costs every iteration a full miss (~0.8 us measured: the three loops of a walker took 12 + 6 + 5
  // iterations x that).  One thread asks the memory system to bring the rows in use into L2 now, while the CTA waits
  // for its coordinates; the loops then see L2 latency.
  // (the first rows now, the rest once the coordinates are converted: 140 KB per CTA requested at once delayed the
  // prologue's own loads by 2 us)
  constexpr int kRowsFirst = 6;
  if (tid == nt - 1 && hdr.y == n3_valid_tag(sys)) {
    const int rows = min(min(hdr.z, sys.K / sys.n3_lpa), kRowsFirst);
    l2_prefetch_bulk(sys.own_cs, static_cast<uint32_t>(static_cast<size_t>(rows) * step * sizeof(int2)));
  }
#endif
  N3Box bx;
  bx.Lx = sys.box[0];  // EAM.cpp:31 reads the diagonal only
  bx.Ly = sys.box[4];
  bx.Lz = sys.box[8];
  for (int k = tid; k < 16; k += nt) sm.t16[k] = P.exp_t16[k];
  for (int k = tid; k < 3; k += nt) invL[k] = 1.0 / sys.box[4 * k];  // (one division per axis, not one per thread)
  const double lo[3] = {ld_cg(sys.lo_build), ld_cg(sys.lo_build + 1), ld_cg(sys.lo_build + 2)};
  const double Lb[3] = {ld_cg(sys.box_build), ld_cg(sys.box_build + 4), ld_cg(sys.box_build + 8)};  // (the fast path
  // is for orthorhombic use: EAM.cpp:31 reads the diagonal only; the cold path compares the whole cell)
  // build positions and atomic numbers of the atoms this thread converts in the first round
  double rb[3] = {0.0, 0.0, 0.0};
  int zz = 0;
  if (tid < N) {
    zz = sys.Z[tid];
    rb[0] = ld_cg(sys.rbuild + 3 * tid);
    rb[1] = ld_cg(sys.rbuild + 3 * tid + 1);
    rb[2] = ld_cg(sys.rbuild + 3 * tid + 2);
  }
  __syncthreads();  // mbarrier initialised
  if (EONB200_LIKELY(aligned)) {
    mbar_wait(sm.mbar, 0);
  } else {
    for (int k = tid; k < D; k += nt) raw[k] = sys.R[k];
    __syncthreads();
  }
  phase_clock(clk, 3);
  // ---- raw coordinates -> positions continuous with the list's build positions; Verlet displacement; element check ----
  // EAM.cpp:56-90 moves the walker up by its lowest negative coordinate and wraps it into the box.  Any rigid shift
  // and any choice of periodic images gives the same pair vectors as long as it is the one the stored folds were
  // recorded under, so the fast path applies the BUILD's shift and takes the image of every atom nearest to its build
  // position: no reduction over the walker, no re-wrap jumps while the list is valid.
  const int z0 = hdr.w >> 8;
  const double dstar = 0.25 * P.skin;
  const double Lk[3] = {bx.Lx, bx.Ly, bx.Lz};
  const double iL[3] = {invL[0], invL[1], invL[2]};
  double disp2_max = 0.0;
  int odd = 0;
  for (int i = tid; i < N; i += nt) {
    if (i != tid) {
      zz = sys.Z[i];
      rb[0] = ld_cg(sys.rbuild + 3 * i);
      rb[1] = ld_cg(sys.rbuild + 3 * i + 1);
      rb[2] = ld_cg(sys.rbuild + 3 * i + 2);
    }
    double q[3], d2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      q[k] = raw[3 * i + k] + lo[k];
      const double n = rint((q[k] - rb[k]) * iL[k]);
      q[k] = fma(-n, Lk[k], q[k]);
      const double d = q[k] - rb[k];
      d2 = fma(d, d, d2);
    }
    n3_store_xy(sm.pxy, sys.n3_shift, i, q[0], q[1]);
    n3_store_zw(sm.pzw, sys.n3_shift, i, q[2], 0.0);
    disp2_max = fmax(disp2_max, d2);
    if (zz != z0) odd |= 1;
  }
  if (tid == 0) {  // the ghost atom of the row padding
    n3_store_xy(sm.pxy, sys.n3_shift, N, 1e150, 1e150);
    n3_store_zw(sm.pzw, sys.n3_shift, N, 1e150, 0.0);
  }
  if (!(disp2_max <= dstar * dstar)) odd |= 1;
  const double need = 2.0 * P.rc_list + P.skin;  // stored folds stay the reference's minimum image
  const bool tab_ok = bx.Lx >= need && bx.Ly >= need && bx.Lz >= need;
  const int r0h = hdr.w & 0xff;
  // The common case is decided by ONE barrier; anything else takes the out-of-line path.
  const bool uniform_ok = a.force_rebuild == 0 && P.skin > 0.0 && tab_ok && P.eam_fast_ok != 0 &&
                          hdr.y == n3_valid_tag(sys) && hdr.x <= sys.slot_cap && r0h < P.n_rows && N <= kN3MaxAtoms &&
                          Lb[0] == bx.Lx && Lb[1] == bx.Ly && Lb[2] == bx.Lz;
  {
    const int rq = min(r0h, P.n_rows - 1);
    for (int k = tid; k < 18; k += nt) coef[k] = k < 9 ? P.rows[rq].c[k] : P.rows[rq].dc[k - 9];
  }
  phase_clock(clk, 4);
  int r0 = r0h, nslots = hdr.x;
  if (EONB200_UNLIKELY(cta_any(odd) != 0 || !uniform_ok)) {
#ifdef EONB200_DEBUG_COLD
    if (tid == 0 && s == 0) printf("cold: odd %d fr %d skin %g tab %d fast %d hdr %d %d %d %d cap %d nrows %d N %d\n", odd, a.force_rebuild, P.skin, (int)tab_ok, P.eam_fast_ok, hdr.x, hdr.y, hdr.z, hdr.w, sys.slot_cap, P.n_rows, N);
#endif
    if (eam_n3_cold(&cold, &nl, &a, Pp, disp2_max) == 0) return;
    __syncthreads();
    r0 = ld_cg(sys.hl_hdr + N3H_ROWZ) & 0xff;
    nslots = ld_cg(sys.hl_hdr + N3H_SLOTS);
    for (int k = tid; k < 18; k += nt) coef[k] = k < 9 ? P.rows[r0].c[k] : P.rows[r0].dc[k - 9];
    if (tid < C) {
      lh = ld_cg(sys.lhdr + tid);
#pragma unroll
      for (int u = 0; u < kN3UD; u++) cd[u] = ld_cg(sys.own_cs + tid + u * step);
    }
    for (int i = tid; i < N; i += nt) {  // (a rebuild prepared the positions anew)
      const double4 p = sm.pos[i];
      n3_store_xy(sm.pxy, sys.n3_shift, i, p.x, p.y);
      n3_store_zw(sm.pzw, sys.n3_shift, i, p.z, 0.0);
    }
    __syncthreads();
  }
#ifndef EONB200_EMU
  if (tid == nt - 1 && hdr.y == n3_valid_tag(sys)) {
    const int rows = min(hdr.z, sys.K / sys.n3_lpa);
    const char *pc = reinterpret_cast<const char *>(sys.own_cs);
    const char *pr = reinterpret_cast<const char *>(sys.recv_code);
    const size_t nc = static_cast<size_t>(rows) * step * sizeof(int2), nr = static_cast<size_t>(rows) * step * sizeof(int);
    constexpr size_t kChunk = 32768;
    for (size_t o = static_cast<size_t>(kRowsFirst) * step * sizeof(int2); o < nc; o += kChunk)
      l2_prefetch_bulk(pc + o, static_cast<uint32_t>(nc - o < kChunk ? nc - o : kChunk));
    for (size_t o = 0; o < nr; o += kChunk) l2_prefetch_bulk(pr + o, static_cast<uint32_t>(nr - o < kChunk ? nr - o : kChunk));
  }
#endif
  phase_clock(clk, 12);
  // spare slots behind the last one: written by ghost pairs, read (times a zero vector) by padded receiver rows
  for (int k = tid; k < kN3UF; k += nt) sm.slots[nslots - kN3UF + k] = 0.0;
  double *F = batch_F(a, s);
  const EamRowDev &row = P.rows[r0];
  N3K K;
  K.rc2 = row.rc2;
  K.nb1 = -row.beta1;
  K.nb2x2 = -2.0 * row.beta2;
  K.nb2x1024 = -1024.0 * row.beta2;
  K.naM = -row.alphaM;
  K.aMRm = row.alphaM * row.Rm;
  K.Dm = row.Dm;
  // sign of the reference's own pair force (EAM.cpp:203-215) or of -grad U
  K.nK2 = -(P.eam_conservative ? 1.0 : -1.0) * 2.0 * row.alphaM * row.Dm;
  if (EONB200_LIKELY(C <= nt))
    eam_n3_fast<true>(K, coef, coef + 9, sys, sm, bx, F, a.U + s, lh, cd, clk);
  else
    eam_n3_fast<false>(K, coef, coef + 9, sys, sm, bx, F, a.U + s, lh, cd, clk);
  phase_clock(clk, 10);
}

#ifdef EONB200_EMU
inline void k_eam_n3(PotParams Pg, const PotParams *Pp, NlBatch nl, ForceBatchArgs a) { eam_n3_body(Pg, Pp, nl, a); }
#else
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_eam_n3_t(const __grid_constant__ PotParams Pg, const PotParams *__restrict__ Pp,
                                                         const __grid_constant__ NlBatch nl, const __grid_constant__ ForceBatchArgs a) {
  eam_n3_body(Pg, Pp, nl, a);
}
#endif

}  // namespace eonb200
