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
// Summation orders are fixed, so results are bit-reproducible.  Walkers whose atoms have drifted beyond skin/4 from
// the list's build positions (every candidate must be screened again) and walkers whose slots do not fit the launch's
// shared memory take the ordered-pair lane functions of potentials.cuh inside the same kernel; mixed-element systems
// and small boxes the general row functions.
#pragma once

namespace eonb200 {

constexpr int kN3Batch = 4;  // list words / scratch words a lane keeps in flight
constexpr int kN3Pre = 16;   // receiver codes a lane requests ahead of the receiver pass

// Lists of this kernel carry the image of a pair as 2 bits per axis above the index (0: none, 1: + L, 2: - L is added
// to r_i - r_j), so that the fold is a pair of selects on the box length instead of a table lookup in shared memory.
__device__ __forceinline__ int n3_code(int mic_code) {
  const int id = static_cast<unsigned>(mic_code) >> kNbrIndexBits;  // (fx + 1) + 3 (fy + 1) + 9 (fz + 1), pack_mic
  const int f[3] = {id % 3 - 1, (id / 3) % 3 - 1, id / 9 - 1};
  int bits = 0;
  for (int k = 0; k < 3; k++) bits |= (f[k] == 0 ? 0 : (f[k] < 0 ? 1 : 2)) << (2 * k);  // sh27: offset = -f L
  return (mic_code & kNbrIndexMask) | (bits << kNbrIndexBits);
}
__device__ __forceinline__ double n3_fold(int bits, double L) { return (bits & 1) ? L : ((bits & 2) ? -L : 0.0); }

// The atom that evaluates the unordered pair {i, j}: the lower index when i + j is even, the higher when odd.
__device__ __forceinline__ bool n3_owner(int i, int j) { return (i < j) == (((i + j) & 1) == 0); }

// Owner / receiver lists from the freshly built full rows (sys.nbr / cls).  CTA-wide; runs after a list rebuild only.
//
// Lanes do not take the atoms in index order: atoms are ranked by the number of pairs they own (descending, ties by
// index) and lane t takes the atom of rank t, so the lanes of a warp run the same number of iterations (surface
// atoms own half as many pairs as bulk atoms).  Every per-atom list below is stored by LANE (column t), so that the
// lanes of a warp still read consecutive words; codes and slots refer to atom indices / slot numbers.
// Rows are padded with null entries (count them out, compute them anyway: no branches in the pair loops): owned rows
// to a multiple of kN3Batch, received rows to a multiple of kN3Batch and at least kN3Pre.
__device__ inline void n3_build(const SysView &sys) {
  const int N = sys.N;
  const size_t Np = static_cast<size_t>(sys.Np);
  const int tid = threadIdx.x, nt = blockDim.x;
  int *key = sys.cnt2;  // [Np] scratch: pairs owned by atom i
  if (tid == 0) sys.hl_hdr[2] = 0;
  for (int i = tid; i < N; i += nt) {
    const int c = sys.cls[i];
    const int n01 = (c & 0xffff) + (c >> 16);
    int no = 0;
    for (int k = 0; k < n01; k++) {
      const int j = sys.nbr[k * Np + i] & kNbrIndexMask;
      if (j != i && n3_owner(i, j)) no++;
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
  int rows = 0;  // rows of the owner / receiver arrays in use
  for (int i = tid; i < N; i += nt) {
    const size_t t = static_cast<size_t>(sys.atom_lane[i]);
    const int c = sys.cls[i];
    const int n0 = c & 0xffff, n01 = n0 + (c >> 16);
    int no_s = 0, no_u = 0, nr = 0;
    for (int k = 0; k < n01; k++) {
      const int code = sys.nbr[k * Np + i];
      const int j = code & kNbrIndexMask;
      if (j == i) continue;  // a self image only exists in boxes the fast path never sees
      if (n3_owner(i, j)) {
        sys.own_cs[(no_s + no_u) * Np + t] = make_int2(n3_code(code), 0);
        if (k < n0) no_s++; else no_u++;
      } else {
        sys.recv_code[nr * Np + t] = n3_code(code);
        nr++;
      }
    }
    const int self = i;  // partner = the atom itself, no image: a harmless null entry
    const int nown = no_s + no_u;
    const int nown_pad = max(kN3Batch, (nown + kN3Batch - 1) / kN3Batch * kN3Batch);
    for (int k = nown; k < nown_pad && k < sys.K; k++) sys.own_cs[k * Np + t] = make_int2(nown > 0 ? sys.own_cs[t].x : self, -1);
    const int nr_pad = max(kN3Pre, (nr + kN3Batch - 1) / kN3Batch * kN3Batch);
    for (int k = nr; k < nr_pad && k < sys.K; k++) sys.recv_code[k * Np + t] = self;
    sys.cls2[t] = make_int2(no_s | (no_u << 16), nr);
    sys.roff[t] = nr;
    rows = max(rows, max(nown_pad, nr_pad));
  }
  __syncthreads();
  atomicMax(sys.hl_hdr + 2, min(rows, sys.K));
  // exclusive scan of the received counts (lane order) -> first slot of every lane's atom
#ifdef EONB200_EMU
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < N; t++) {
      const int v = sys.roff[t];
      sys.roff[t] = run;
      run += v;
    }
    sys.roff[N] = run;
  }
#else
  if (tid < 32) {
    int carry = 0;
    for (int base = 0; base < N; base += 32) {
      const int t = base + tid;
      const int v = t < N ? sys.roff[t] : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (tid >= o) x += y;
      }
      if (t < N) sys.roff[t] = carry + x - v;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (tid == 0) sys.roff[N] = carry;
  }
#endif
  __syncthreads();
  // every owned pair learns the slot its partner reads it from; null entries point at the spare slot behind the last one
  const int total = sys.roff[N];
  for (int i = tid; i < N; i += nt) {
    const size_t t = static_cast<size_t>(sys.atom_lane[i]);
    const int cx = sys.cls2[t].x;
    const int nown = (cx & 0xffff) + (cx >> 16);
    const int nown_pad = max(kN3Batch, (nown + kN3Batch - 1) / kN3Batch * kN3Batch);
    for (int k = 0; k < nown_pad && k < sys.K; k++) {
      const int code = sys.own_cs[k * Np + t].x;
      int slot = total;
      if (k < nown) {
        const size_t tj = static_cast<size_t>(sys.atom_lane[code & kNbrIndexMask]);
        const int nrj = sys.cls2[tj].y;
        int r = 0;
        while (r < nrj && (sys.recv_code[r * Np + tj] & kNbrIndexMask) != i) r++;
        slot = sys.roff[tj] + (r < nrj ? r : 0);
      }
      sys.own_cs[k * Np + t] = make_int2(code, slot);
    }
  }
  if (tid == 0) {
    sys.hl_hdr[0] = total + 1;  // slots the fast path needs (one spare for the null entries)
    sys.hl_hdr[1] = 1;
    atomicMax(&sys.stats[3], static_cast<unsigned long long>(total + 1));
  }
  __syncthreads();
}

__device__ EONB200_NOINLINE void eam_n3_build(const FusedCold *fc) {
  SysView sys;
  Smem sm;
  fused_cold_views(*fc, sys, sm);
  n3_build(sys);
}


struct N3Lane {  // what one lane carries through the passes for one of its atoms
  double rho_own, dfd, fx, fy, fz;
  int atom;              // the atom this lane works for (lanes take the atoms by rank, see n3_build)
  int nso, nown, nrecv;  // owned sure-in pairs, owned pairs, received pairs
  int rbase;             // its first slot
};

// Raw list words of lane position t (requested early, unpacked when needed: a warp stalls at the first USE of a load).
struct N3Raw {
  int2 c2;
  int rbase, atom;
};
__device__ __forceinline__ void n3_raw_load(const SysView &sys, int t, N3Raw &r) {
  r.c2 = ld_cg(sys.cls2 + t);
  r.rbase = ld_cg(sys.roff + t);
  r.atom = ld_cg(sys.lane_atom + t);
}
__device__ __forceinline__ void n3_unpack(const N3Raw &r, N3Lane &L) {
  L.nso = r.c2.x & 0xffff;
  L.nown = L.nso + (r.c2.x >> 16);
  L.nrecv = r.c2.y;
  L.rbase = r.rbase;
  L.atom = r.atom;
}
__device__ __forceinline__ int n3_pad(int n) { return max(kN3Batch, (n + kN3Batch - 1) & ~(kN3Batch - 1)); }

// exp with a 16-entry table that fills exactly one 128-byte line of shared memory (a warp's lookups never conflict):
// x = (16 k + j) ln2/16 + r, |r| <= ln2/32, e^x = 2^k T[j] (1 + r + ... + r^7/5040); error < 1 ulp (r^8/8! = 1e-18).
// No underflow guard (PotParams::eam_fast_ok).
__device__ __forceinline__ double exp_t16(double x, const double *T) {
  const double kInv = 23.083120654223414;     // 16 / ln 2
  const double kShift = 6755399441055744.0;   // 1.5 * 2^52
  const double kLn2Hi = 6.93147180369123816490e-01 / 16.0;  // high part of ln2 (trailing zeros), exact / 16
  const double kLn2Lo = 1.90821492927058770002e-10 / 16.0;
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
  const double t = T[k & 15];
  double q = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
  q = fma(q, r, 1.0 / 120.0);
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
  return __hiloint2double(__double2hiint(res) + ((k >> 4) << 20), __double2loint(res));
#endif
}

struct N3Box {
  double Lx, Ly, Lz;
};
// d = r_i - r_j folded by the image recorded at build time (n3_code)
__device__ __forceinline__ void n3_vec(const N3Box &bx, int c, const double2 &ixy, double iz, const double2 &jxy, double jz,
                                       double &dx, double &dy, double &dz) {
  const int b = static_cast<unsigned>(c) >> kNbrIndexBits;
  dx = (ixy.x - jxy.x) + n3_fold(b, bx.Lx);
  dy = (ixy.y - jxy.y) + n3_fold(b >> 2, bx.Ly);
  dz = (iz - jz) + n3_fold(b >> 4, bx.Lz);
}

// Owner density pass of lane position t: one pair per iteration, the next list word one iteration ahead (the loops
// are kept lean on purpose: two CTAs of ~416 threads share an SM under a 72-register budget, and thread-level
// parallelism -- 26 warps -- hides the latencies).
__device__ __forceinline__ void n3_density(const PotParams &P, const EamRowDev *row, const SysView &sys, const Smem &sm,
                                           const N3Box &bx, int t, N3Lane &L, int2 cs0, double (&acc)[2]) {
  const double *T = sm.t16;
  const size_t step = static_cast<size_t>(sys.Np);
  const int2 *oc = sys.own_cs + t;
  double2 *oab = sys.own_ab + t;
  const double rc2 = row->rc2;
  const double nb1 = -row->beta1, nb2x2 = -2.0 * row->beta2, nb2x1024 = -1024.0 * row->beta2;
  const double naM = -row->alphaM, aMRm = row->alphaM * row->Rm, Dm = row->Dm;
  // sign of the reference's own pair force (EAM.cpp:203-215) or of -grad U
  const double nK2 = -(P.eam_conservative ? 1.0 : -1.0) * 2.0 * row->alphaM * row->Dm;
  const double2 ixy = sm.pxy[L.atom];
  const double iz = sm.pzw[L.atom].x;
  double dens = 0.0, e_acc = 0.0;
  int npairs = 0;
  const int nown = L.nown, nso = L.nso;
  int2 cnext = cs0;
  for (int k = 0; k < nown; k++) {
    const int2 cs = cnext;
    oc += step;
    cnext = ld_cg(oc);  // (row nown is valid memory: rows are padded)
    const int c = cs.x;
    const int j = c & kNbrIndexMask;
    double dx, dy, dz;
    n3_vec(bx, c, ixy, iz, sm.pxy[j], sm.pzw[j].x, dx, dy, dz);
    const double r2 = norm2_rn(dx, dy, dz);
    double g = 0.0, A = 0.0, nB = 0.0;
    if (EONB200_LIKELY(k < nso || !(r2 > rc2))) {  // EAM.cpp:172 decides for the uncertain class; sure-in pairs are inside
      npairs++;
      const double invr = rsqrt_fast(r2);
      const double r = r2 * invr;
      const double r4 = r2 * r2;
      const double e1 = exp_t16(nb1 * r, T);
      const double e2 = exp_t16(nb2x2 * r, T);
      const double ex = exp_t16(fma(naM, r, aMRm), T);
      const double rho_pair = fma(512.0, e2, e1);
      g = (r4 * r2) * rho_pair;  // EAM.cpp:178-183
      const double uu = fma(nb2x1024, e2, nb1 * e1);
      // g'(r)/r = [6 r^5 rho_pair + r^6 u] / r = r^4 (6 rho_pair + r u)   (EAM.cpp:186-198)
      A = r4 * fma(r, uu, 6.0 * rho_pair);
      const double dm = 1.0 - ex;
      nB = (nK2 * dm) * (ex * invr);              // -(+-phi'(r)/r), EAM.cpp:200-217
      e_acc = fma(Dm, fma(dm, dm, -1.0), e_acc);  // Dm (d^2 - 1): once per unordered pair
    }
    dens += g;
    sm.slots[cs.y] = g;
    *oab = make_double2(A, nB);
    oab += step;
  }
  L.rho_own = dens;
  acc[0] += e_acc;
  acc[1] += static_cast<double>(npairs);
}

// rho_i = own part + received slots (fixed order: four interleaved partial sums); F(rho), F'(rho).  Returns F(rho_i).
__device__ __forceinline__ double n3_embed(const EamRowDev *row, const Smem &sm, N3Lane &L) {
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
  const double dens = (s0 + s1) + (s2 + s3);
  L.dfd = eam_dembed(row->dc, dens);
  return eam_embed(row->c, dens);
}

// Owner force pass.  Pairs outside the cutoff carry zero weights (written by the density pass).  The body is a dozen
// instructions, so the list and scratch words run kN3Ring iterations ahead (an L2 round trip is ~1000 cycles).
constexpr int kN3Ring = 4;
__device__ __forceinline__ void n3_force_own(const SysView &sys, const Smem &sm, const N3Box &bx, int t, N3Lane &L,
                                             int2 (&cq)[kN3Ring], double2 (&wq)[kN3Ring]) {
  const size_t step = static_cast<size_t>(sys.Np);
  const int2 *oc = sys.own_cs + t + kN3Ring * step;
  const double2 *oab = sys.own_ab + t + kN3Ring * step;
  const double2 ixy = sm.pxy[L.atom];
  const double iz = sm.pzw[L.atom].x;
  const double dfi = L.dfd;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  const int nown = L.nown;
  for (int k0 = 0; k0 < nown; k0 += kN3Ring) {
#pragma unroll
    for (int u = 0; u < kN3Ring; u++) {
      const int2 cs = cq[u];
      const double2 w = wq[u];
      if (k0 + u + kN3Ring < nown) {  // refill this ring position
        cq[u] = ld_cg(oc + u * step);
        wq[u] = ld_cg(oab + u * step);
      }
      if (k0 + u < nown) {
        const int c = cs.x;
        const int j = c & kNbrIndexMask;
        const double2 jzw = sm.pzw[j];
        double dx, dy, dz;
        n3_vec(bx, c, ixy, iz, sm.pxy[j], jzw.x, dx, dy, dz);
        const double sc = fma(-(dfi + jzw.y), w.x, w.y);
        fx = fma(sc, dx, fx);
        fy = fma(sc, dy, fy);
        fz = fma(sc, dz, fz);
        sm.slots[cs.y] = sc;
      }
    }
    oc += kN3Ring * step;
    oab += kN3Ring * step;
  }
  L.fx = fx;
  L.fy = fy;
  L.fz = fz;
}
__device__ __forceinline__ void n3_own_prefetch(const SysView &sys, int t, int2 (&cq)[kN3Ring], double2 (&wq)[kN3Ring]) {
  const size_t step = static_cast<size_t>(sys.Np);
#pragma unroll
  for (int u = 0; u < kN3Ring; u++) {  // (rows beyond the count are valid memory: rows are padded to a multiple of 4)
    cq[u] = ld_cg(sys.own_cs + t + u * step);
    wq[u] = ld_cg(sys.own_ab + t + u * step);
  }
}

// Receiver force pass: F_i += s d' over the atom's slots in slot order, kN3RecvRing list words ahead.
constexpr int kN3RecvRing = 8;
__device__ __forceinline__ void n3_recv_prefetch(const SysView &sys, int t, int (&rq)[kN3RecvRing]) {
  const size_t step = static_cast<size_t>(sys.Np);
#pragma unroll
  for (int u = 0; u < kN3RecvRing; u++) rq[u] = ld_cg(sys.recv_code + t + u * step);  // (padded to >= kN3Pre rows)
}
__device__ __forceinline__ void n3_force_recv(const SysView &sys, const Smem &sm, const N3Box &bx, int t, const N3Lane &L,
                                              double *F, int (&rq)[kN3RecvRing]) {
  const size_t step = static_cast<size_t>(sys.Np);
  const int *rc = sys.recv_code + t + kN3RecvRing * step;
  const double *sl = sm.slots + L.rbase;
  const double2 ixy = sm.pxy[L.atom];
  const double iz = sm.pzw[L.atom].x;
  double fx = L.fx, fy = L.fy, fz = L.fz;
  const int nr = L.nrecv;
  for (int k0 = 0; k0 < nr; k0 += kN3RecvRing) {
#pragma unroll
    for (int u = 0; u < kN3RecvRing; u++) {
      const int c = rq[u];
      if (k0 + u + kN3RecvRing < nr) rq[u] = ld_cg(rc + u * step);
      if (k0 + u < nr) {
        const int j = c & kNbrIndexMask;
        double dx, dy, dz;
        n3_vec(bx, c, ixy, iz, sm.pxy[j], sm.pzw[j].x, dx, dy, dz);
        const double sc = sl[k0 + u];
        fx = fma(sc, dx, fx);
        fy = fma(sc, dy, fy);
        fz = fma(sc, dz, fz);
      }
    }
    rc += kN3RecvRing * step;
  }
  const int i = L.atom;
  F[3 * i] = fx;
  F[3 * i + 1] = fy;
  F[3 * i + 2] = fz;
}

// One lane position per thread (SINGLE: the lane state stays in registers between the passes, the first words of the
// force passes are requested while the CTA meets at the barriers) or several (state through HBM scratch: sys.dfd holds
// the own density part, the output rows the own force part).
template <bool SINGLE>
__device__ __forceinline__ void eam_n3_fast(const PotParams &P, const EamRowDev *row, const SysView &sys, const Smem &sm,
                                            const N3Box &bx, double *F, double *U, const N3Raw &raw0, int2 cs0, long long *clk) {
  const int N = sys.N;
  const int tid = threadIdx.x, nt = blockDim.x;
  double acc[2] = {0.0, 0.0};
  N3Lane L;
  N3Raw rw = raw0;
  int2 cq[kN3Ring];
  double2 wq[kN3Ring];
  int rq[kN3RecvRing];
  for (int t = tid; t < N; t += nt) {
    if (!SINGLE && t != tid) {
      n3_raw_load(sys, t, rw);
      cs0 = ld_cg(sys.own_cs + t);
    }
    n3_unpack(rw, L);
    n3_density(P, row, sys, sm, bx, t, L, cs0, acc);
    if (!SINGLE) sys.dfd[t] = L.rho_own;
  }
  if (SINGLE && tid < N) {  // first words of the two force passes
    n3_own_prefetch(sys, tid, cq, wq);
    n3_recv_prefetch(sys, tid, rq);
  }
  phase_clock(clk, 6);
  __syncthreads();
  phase_clock(clk, 7);
  for (int t = tid; t < N; t += nt) {
    if (!SINGLE) {
      n3_raw_load(sys, t, rw);
      n3_unpack(rw, L);
      L.rho_own = sys.dfd[t];
    }
    acc[0] += n3_embed(row, sm, L);
    sm.pzw[L.atom].y = L.dfd;
  }
  __syncthreads();
  phase_clock(clk, 8);
  for (int t = tid; t < N; t += nt) {
    if (!SINGLE) {
      n3_raw_load(sys, t, rw);
      n3_unpack(rw, L);
      L.dfd = sm.pzw[L.atom].y;
      n3_own_prefetch(sys, t, cq, wq);
    }
    n3_force_own(sys, sm, bx, t, L, cq, wq);
    if (!SINGLE) {
      F[3 * L.atom] = L.fx;
      F[3 * L.atom + 1] = L.fy;
      F[3 * L.atom + 2] = L.fz;
    }
  }
  __syncthreads();
  phase_clock(clk, 9);
  for (int t = tid; t < N; t += nt) {
    if (!SINGLE) {
      n3_raw_load(sys, t, rw);
      n3_unpack(rw, L);
      L.fx = F[3 * L.atom];
      L.fy = F[3 * L.atom + 1];
      L.fz = F[3 * L.atom + 2];
      n3_recv_prefetch(sys, t, rq);
    }
    n3_force_recv(sys, sm, bx, t, L, F, rq);
  }
  block_sum<2>(acc, sm.red);
  if (tid == 0) {
    *U = acc[0];
    atomicAdd(&sys.stats[2], 2ULL * static_cast<unsigned long long>(acc[1]));  // ordered in-cutoff pairs
  }
}

// The ordered-pair lane functions (filter -> density -> force over the walker's full rows, potentials.cuh) inside this
// kernel: walkers beyond skin/4 of their build positions, or whose slots do not fit.
__device__ EONB200_NOINLINE void eam_n3_ordered(const FusedCold *fc, const PotParams *P_sh, int r0, bool near_build) {
  SysView sys;
  Smem sm;
  fused_cold_views(*fc, sys, sm);
  const PotParams &P = *P_sh;
  const int N = sys.N, tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < N; i += nt) {
    const double4 p = sm.pos[i];
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
  double *F = fc->a.F + static_cast<size_t>(blockIdx.y) * fc->a.strideR;
  double *U = fc->a.U + blockIdx.y;
  if (N <= nt)
    eam_fused_fast<true>(P, ff, sys, sm, 0, N, 1u, 0u, F, U, n0_first, kend_first, cur, nullptr);
  else
    eam_fused_fast<false>(P, ff, sys, sm, 0, N, 1u, 0u, F, U, n0_first, kend_first, cur, nullptr);
}

// Smallest value over the CTA of three per-thread values (all threads get the result).  One barrier; the caller's
// next barrier protects the scratch.
__device__ __forceinline__ void cta_min3(double (&v)[3], double *scratch) {
#ifndef EONB200_EMU
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < 3; k++) v[k] = warp_min(v[k]);
  if (lane == 0) {
    scratch[warp] = v[0];
    scratch[kMaxWarps + warp] = v[1];
    scratch[2 * kMaxWarps + warp] = v[2];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 3; k++) v[k] = warp_min(lane < nwarps ? scratch[k * kMaxWarps + lane] : 1e300);
#else
  (void)scratch;
#endif
}

// grid (1, S): one CTA per walker.  Pg: the parameter block as a kernel argument (constant bank: no load latency on
// the hot path); Pp: the same block in HBM for the out-of-line cold paths.
__device__ __forceinline__ void eam_n3_body(const PotParams &P, const PotParams *__restrict__ Pp, const NlBatch &nl,
                                            const ForceBatchArgs &a) {
  EONB200_DYN_SMEM(smem_raw);
  const int s = blockIdx.y;
  __shared__ FusedCold cold;  // filled only when a cold path is about to run
  // the element rows come out of the constant bank (cold at every launch) while the positions are staged
  __shared__ __align__(16) unsigned long long rows_sh[(sizeof(EamRowDev) * kMaxEamRows + 7) / 8];
  {
    const int nw = static_cast<int>((sizeof(EamRowDev) * P.n_rows + 7) / 8);
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(P.rows);
    for (int k = threadIdx.x; k < nw; k += blockDim.x) rows_sh[k] = src[k];
  }
  const EamRowDev *rows = reinterpret_cast<const EamRowDev *>(rows_sh);
  const SysView sys = batch_sys(nl, a, s);
  Smem sm = carve_smem_n3(smem_raw, a.N, a.maxCells, sys.cellscratch);
  for (int k = threadIdx.x; k < 16; k += blockDim.x) sm.t16[k] = P.exp_t16[k];
  const int N = a.N, tid = threadIdx.x, nt = blockDim.x;
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
  // The walker's coordinates land in the tail of the position block.  A bulk copy needs 16-byte aligned ends on both
  // sides: rows of an odd 3N start on odd multiples of 8 bytes for every second walker, so the copy moves the aligned
  // interior and up to one leading and one trailing coordinate are fetched by plain loads.
  const int D = 3 * N;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(sys.R);
  const bool aligned = (addr & 7u) == 0;
  const int head = (addr & 15u) ? 1 : 0;
  const int body = (D - head) & ~1;
  double *raw = reinterpret_cast<double *>(sm.pos) + (N + (N & 1)) + head;  // &raw[head] is 16-byte aligned
  const int h_rows = ld_cg(sys.hl_hdr + 2);  // rows of the owner / receiver lists in use (largest count over the atoms)
  if (tid == 0) {
    mbar_init(sm.mbar, 1);
    if (aligned) {
      mbar_expect(sm.mbar, static_cast<uint32_t>(body) * 8u + 1024u);
      bulk_g2s_issue(raw + head, sys.R + head, static_cast<uint32_t>(body) * 8u, sm.mbar);
      bulk_g2s_issue(sm.exp_tab, P.exp_tab, 1024u, sm.mbar);
    }
  }
  if (tid == 64 % nt && h_rows > 0 && h_rows <= sys.K) {
    // the list rows this walker is about to read are contiguous (slot-major): bring them into L2 with three
    // instructions while the positions are staged
    const uint32_t rows = static_cast<uint32_t>(h_rows) * static_cast<uint32_t>(sys.Np);
    l2_prefetch_bulk(sys.own_cs, rows * 8u);
    l2_prefetch_bulk(sys.recv_code, rows * 4u);
  }
  if (aligned && tid == 32 % nt && head) raw[0] = sys.R[0];
  if (aligned && tid == 33 % nt && head + body < D) raw[D - 1] = sys.R[D - 1];
  const double Lbox[3] = {sys.box[0], sys.box[4], sys.box[8]};  // EAM.cpp:31 reads the diagonal only
  const bool have = ld_cg(sys.state) != 0;
  const int z0 = sys.Z[0];
  int h_slots = ld_cg(sys.hl_hdr);
  const int h_valid = ld_cg(sys.hl_hdr + 1);
  // this lane's atoms: list words and the first batch of its owned pairs (speculative: lists that turn out stale are
  // read again after the rebuild).  Nothing is USED before the positions have landed.
  N3Raw lraw;
  int2 cs0 = make_int2(0, 0);
  lraw.c2 = make_int2(0, 0);
  lraw.rbase = lraw.atom = 0;
  if (tid < N) {
    n3_raw_load(sys, tid, lraw);
    cs0 = ld_cg(sys.own_cs + tid);
  }
  // build positions and atomic numbers of the atoms this thread converts in the first round
  double rb[kConvAtoms][3];
  int zz[kConvAtoms];
  int odd = 0;
#pragma unroll
  for (int u = 0; u < kConvAtoms; u++) {
    const int i = u * nt + tid;
    zz[u] = z0;
    if (i < N) {
      zz[u] = sys.Z[i];
      rb[u][0] = ld_cg(sys.rbuild + 3 * i);
      rb[u][1] = ld_cg(sys.rbuild + 3 * i + 1);
      rb[u][2] = ld_cg(sys.rbuild + 3 * i + 2);
    }
  }
  fill_sh27(sm, Lbox[0], Lbox[1], Lbox[2]);
  __syncthreads();  // mbarrier initialised
  if (EONB200_LIKELY(aligned)) {
    mbar_wait(sm.mbar, 0);
  } else {
    for (int k = tid; k < D; k += nt) raw[k] = sys.R[k];
    for (int k = tid; k < 128; k += nt) sm.exp_tab[k] = P.exp_tab[k];
    __syncthreads();
  }
  phase_clock(clk, 3);
  // ---- EAM.cpp:56-83: where a coordinate is negative the whole walker moves up by the lowest one.  The first
  // round's atoms are read into registers here; they are converted in place after the reduction's barrier (pos[i]
  // overlaps only raw[i'] with i' <= i, and every later round meets at a barrier between its reads and its writes).
  double lo[3] = {0.0, 0.0, 0.0};
  double q[kConvAtoms][3];
#pragma unroll
  for (int u = 0; u < kConvAtoms; u++) {
    const int i = u * nt + tid;
    if (i < N) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        q[u][k] = raw[3 * i + k];
        lo[k] = fmin(lo[k], q[u][k]);
      }
    }
  }
  for (int i = kConvAtoms * nt + tid; i < N; i += nt) {
    lo[0] = fmin(lo[0], raw[3 * i]);
    lo[1] = fmin(lo[1], raw[3 * i + 1]);
    lo[2] = fmin(lo[2], raw[3 * i + 2]);
  }
  cta_min3(lo, sm.red);
  // ---- raw coordinates -> shifted, wrapped positions; Verlet displacement; element check ----
  double disp2_max = 0.0;
  const double dstar = 0.25 * P.skin;
  for (int base = 0; base < N; base += kConvAtoms * nt) {
    if (base > 0) {
#pragma unroll
      for (int u = 0; u < kConvAtoms; u++) {
        const int i = base + u * nt + tid;
        if (i < N) {
          zz[u] = sys.Z[i];
          rb[u][0] = ld_cg(sys.rbuild + 3 * i);
          rb[u][1] = ld_cg(sys.rbuild + 3 * i + 1);
          rb[u][2] = ld_cg(sys.rbuild + 3 * i + 2);
          q[u][0] = raw[3 * i];
          q[u][1] = raw[3 * i + 1];
          q[u][2] = raw[3 * i + 2];
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < kConvAtoms; u++) {
      const int i = base + u * nt + tid;
      if (i < N) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          if (lo[k] < 0.0) q[u][k] += -lo[k];  // EAM.cpp:78-83 (the shift is |lowest coordinate|)
          if (Lbox[k] > 0.0) {
            int guard = 0;
            while (q[u][k] > Lbox[k] && guard++ < 1000000) q[u][k] -= Lbox[k];  // EAM.cpp:86-90
          }
        }
        sm.pos[i] = make_double4(q[u][0], q[u][1], q[u][2], 0.0);  // (the cold paths' layout)
        sm.pxy[i] = make_double2(q[u][0], q[u][1]);
        sm.pzw[i] = make_double2(q[u][2], 0.0);
        const double dx = q[u][0] - rb[u][0], dy = q[u][1] - rb[u][1], dz = q[u][2] - rb[u][2];
        disp2_max = fmax(disp2_max, dx * dx + dy * dy + dz * dz);
        if (zz[u] != z0) odd |= 2;
      }
    }
  }
  if (!(disp2_max <= dstar * dstar)) odd |= 1;
  const double need = 2.0 * P.rc_list + P.skin;  // stored folds stay the reference's minimum image
  const bool tab_ok = Lbox[0] >= need && Lbox[1] >= need && Lbox[2] >= need;
  // The common case -- lists present, every atom within skin/4 of its build position, one element -- is decided by ONE
  // barrier; anything else takes the full decision below.
  const bool uniform_ok = have && a.force_rebuild == 0 && P.skin > 0.0 && tab_ok && P.eam_fast_ok != 0 && h_valid == 1;
  bool near_build = true;
  phase_clock(clk, 4);
  if (EONB200_UNLIKELY(cta_any(odd) != 0 || !uniform_ok)) {
    double red2[2] = {disp2_max, (odd & 2) ? 1.0 : 0.0};
    {
      const int op[2] = {RED_MAX, RED_MAX};
      block_reduce<2>(red2, op, sm.red);
    }
    if (tid == 0) {
      cold.nl = nl;
      cold.a = a;
      cold.Pp = Pp;
      sm.pos[N] = make_double4(1e150, 1e150, 1e150, 0.0);  // ghost atom of the full rows' padding (general path)
    }
    __syncthreads();
    const bool mixed = red2[1] != 0.0;
    const double disp2 = red2[0];
    const double lim = 0.5 * P.skin;
    const bool rebuild = a.force_rebuild != 0 || !have || P.skin <= 0.0 || !(disp2 <= lim * lim) || h_valid != 1;
    near_build = rebuild || disp2 <= dstar * dstar;  // a fresh list is built from these very positions
    if (rebuild) {
      eam_fused_rebuild(&cold, Pp);
      eam_n3_build(&cold);
      h_slots = ld_cg(sys.hl_hdr);
      if (tid < N) {
        n3_raw_load(sys, tid, lraw);
        cs0 = ld_cg(sys.own_cs + tid);
      }
      for (int i = tid; i < N; i += nt) {  // prepare_positions rewrote sm.pos
        const double4 p = sm.pos[i];
        sm.pxy[i] = make_double2(p.x, p.y);
        sm.pzw[i] = make_double2(p.z, 0.0);
      }
      __syncthreads();
    }
    if (mixed || !tab_ok || !P.eam_fast_ok) {  // general rows: full element table, geometry for mic_vec
      eam_fused_general(&cold, tab_ok, 0, N);
      return;
    }
  }
  phase_clock(clk, 12);
  double *F = a.F + static_cast<size_t>(s) * a.strideR;
  int r0 = -1;
  for (int q = 0; q < P.n_rows; q++)
    if (rows[q].Z == z0) {
      r0 = q;
      break;
    }
  if (EONB200_UNLIKELY(r0 < 0)) {  // element without a row: the reference throws 14324 (EAM.cpp:116-123)
    for (int k = tid; k < 3 * N; k += nt) F[k] = 0.0;
    if (tid == 0) a.U[s] = nan("");
    return;
  }
  if (EONB200_UNLIKELY(!near_build || h_slots > sys.slot_cap)) {  // CTA-uniform
    if (tid == 0) {
      cold.nl = nl;
      cold.a = a;
      cold.Pp = Pp;
    }
    __syncthreads();
    eam_n3_ordered(&cold, Pp, r0, near_build);
    return;
  }
  N3Box bx;
  bx.Lx = Lbox[0];
  bx.Ly = Lbox[1];
  bx.Lz = Lbox[2];
  if (EONB200_LIKELY(N <= nt))
    eam_n3_fast<true>(P, rows + r0, sys, sm, bx, F, a.U + s, lraw, cs0, clk);
  else
    eam_n3_fast<false>(P, rows + r0, sys, sm, bx, F, a.U + s, lraw, cs0, clk);
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
