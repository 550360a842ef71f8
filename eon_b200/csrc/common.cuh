// common.cuh -- CTA-level primitives shared by every kernel of libeonb200.
//
// One CTA owns one walker (one atomic configuration): all control flow is CTA-uniform, every
// vector operation is a block-strided loop and every scalar a block reduction whose result is
// broadcast to all threads.  Reductions use a fixed order (thread-strided partials -> xor
// butterfly inside the warp -> warp partials summed in warp order), so a walker's trajectory is
// reproducible from run to run.
//
// EONB200_EMU: the same sources are compiled as plain C++ (one "thread" per CTA, see
// tests/emu/emu_shim.h) by the CPU-only test-suite to exercise the algorithm logic where no
// GPU exists.  The emulation is test infrastructure; the product library never contains it.
#pragma once

#ifdef EONB200_EMU
#include "emu_shim.h"
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#endif

#include <cmath>

namespace eonb200 {

constexpr int kMaxWarps = 32;
constexpr double kPi = 3.14159265358979323846;
constexpr double kSafeEps = 1e-300;  // include/eon/SafeMath.h:19

// Branch hints: the rarely taken blocks of the hot kernels are laid out behind the hot instruction stream
// (instruction caches are cold at every launch; each jump over an inline cold block costs a fetch round trip).
#define EONB200_UNLIKELY(x) (__builtin_expect(!!(x), 0))
#define EONB200_LIKELY(x) (__builtin_expect(!!(x), 1))

#ifdef EONB200_EMU
#define __grid_constant__
#define EONB200_NOINLINE __attribute__((noinline))
#else
#define EONB200_NOINLINE __noinline__
#endif

#ifdef EONB200_EMU
#define EONB200_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
#define EONB200_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (smem), [&]() { kernel(__VA_ARGS__); })
#else
#define EONB200_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define EONB200_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// Launch with `cx` CTAs (along grid x) per thread-block cluster.  The emulation runs every CTA on its own
// (callers pass cx == 1 there).
#ifdef EONB200_EMU
#define EONB200_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, cx, ...) \
  emu::launch((grid), (smem), [&]() { kernel(__VA_ARGS__); })
#else
#define EONB200_LAUNCH_CLUSTER(kernel, grid_, block_, smem_, stream_, cx_, ...)                  \
  do {                                                                                         \
    if ((cx_) == 1) {                                                                          \
      kernel<<<(grid_), (block_), (smem_), (stream_)>>>(__VA_ARGS__);                          \
      break;                                                                                   \
    }                                                                                          \
    cudaLaunchConfig_t cfg_ = {};                                                              \
    cfg_.gridDim = (grid_);                                                                    \
    cfg_.blockDim = dim3(block_);                                                              \
    cfg_.dynamicSmemBytes = (smem_);                                                           \
    cfg_.stream = (stream_);                                                                   \
    cudaLaunchAttribute attr_[1];                                                              \
    attr_[0].id = cudaLaunchAttributeClusterDimension;                                         \
    attr_[0].val.clusterDim.x = (cx_);                                                         \
    attr_[0].val.clusterDim.y = 1;                                                             \
    attr_[0].val.clusterDim.z = 1;                                                             \
    cfg_.attrs = attr_;                                                                        \
    cfg_.numAttrs = 1;                                                                         \
    cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                                            \
  } while (0)
#endif

// ---- un-fused arithmetic for bit-exact predicates --------------------------------------------
// The neighbour predicate must round exactly like the x86-64 reference (no FMA contraction):
// client/thirdparty/vesin/vesin-single-build.cpp:41-43,138-144,1199-1202.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double norm2_rn(double x, double y, double z) {
  return add_rn(add_rn(mul_rn(x, x), mul_rn(y, y)), mul_rn(z, z));
}

// ---- warp / block reductions -------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#ifndef EONB200_EMU
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#ifndef EONB200_EMU
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#ifndef EONB200_EMU
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
  return v;
}

// Sum over `lanes` adjacent lanes (lanes a power of two <= 32); every lane of the group gets the total.
template <int LANES>
__device__ __forceinline__ double group_sum(double v) {
#ifndef EONB200_EMU
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}

enum RedOp { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };

// Block reduction of NV values at once; scratch holds NV * kMaxWarps doubles of shared memory.
// Every thread returns with the identical result in v[]: warp partials -> shared memory -> every warp folds the
// partials with the same xor butterfly (fixed order, no serial tail).  Two barriers (the second frees the scratch).
__device__ __forceinline__ double red_identity(int op) {
  return op == RED_SUM ? 0.0 : (op == RED_MAX ? -HUGE_VAL : HUGE_VAL);
}
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], const int (&op)[NV], double *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  // Callers arrive from loops whose trip count differs between lanes; without a reconvergence point the shuffles take
  // their divergent-warp path (WARPSYNC.COLLECTIVE per shuffle -- seen in ncu on the EAM kernel's energy reduction).
  __syncwarp();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double r = op[k] == RED_SUM ? warp_sum(v[k]) : (op[k] == RED_MAX ? warp_max(v[k]) : warp_min(v[k]));
    if (lane == 0) scratch[k * kMaxWarps + warp] = r;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const double x = lane < nwarps ? scratch[k * kMaxWarps + lane] : red_identity(op[k]);
    v[k] = op[k] == RED_SUM ? warp_sum(x) : (op[k] == RED_MAX ? warp_max(x) : warp_min(x));
  }
  __syncthreads();
}

template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch) {
  int op[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) op[k] = RED_SUM;
  block_reduce<NV>(v, op, scratch);
}

__device__ __forceinline__ double block_sum1(double x, double *scratch) {
  double v[1] = {x};
  block_sum<1>(v, scratch);
  return v[0];
}
__device__ __forceinline__ double block_max1(double x, double *scratch) {
  double v[1] = {x};
  const int op[1] = {RED_MAX};
  block_reduce<1>(v, op, scratch);
  return v[0];
}

// ---- thread-block clusters (one cluster per walker in the fused EAM kernel) ---------------------
// A walker whose atoms are split over T CTAs keeps those CTAs on neighbouring SMs as one cluster:
// per-atom dF/drho and the energy partials cross CTAs through distributed shared memory
// (mapa + st.shared::cluster) and one cluster barrier separates the density pass from the force pass.
__device__ __forceinline__ unsigned cluster_rank() {
#ifndef EONB200_EMU
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
#else
  return 0;
#endif
}
__device__ __forceinline__ unsigned cluster_size() {
#ifndef EONB200_EMU
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
#else
  return 1;
#endif
}
// Full cluster barrier with release / acquire semantics (covers shared::cluster and global writes).
__device__ __forceinline__ void cluster_sync_all() {
#ifndef EONB200_EMU
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
// Store a double into the same shared-memory location of CTA `rank` of this cluster.
__device__ __forceinline__ void dsmem_store(double *local_smem_ptr, unsigned rank, double v) {
#ifndef EONB200_EMU
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local_smem_ptr));
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
#else
  (void)rank;
  *local_smem_ptr = v;
#endif
}

// ---- SafeMath (include/eon/SafeMath.h:19-80) ----------------------------------------------------
__device__ __forceinline__ double safe_div(double num, double den, double fb) {
  return fabs(den) < kSafeEps ? fb : num / den;
}
__device__ __forceinline__ double safe_recip(double x, double fb) { return safe_div(1.0, x, fb); }
__device__ __forceinline__ double safe_acos(double x) { return acos(fmin(1.0, fmax(-1.0, x))); }
__device__ __forceinline__ double safe_atan_ratio(double num, double den, double fb) {
  return fabs(den) < kSafeEps ? fb : atan(num / den);
}

// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) ------------------------------------------
// Per-walker arrays (positions, forces) are contiguous and 16-byte aligned in HBM, so one elected
// thread moves a whole walker with a single bulk-copy instruction completing on an mbarrier.
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
#ifndef EONB200_EMU
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
  *bar = 0;
#endif
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
#ifndef EONB200_EMU
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
#else
  memcpy(smem_dst, gsrc, bytes);
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
#ifndef EONB200_EMU
  const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(b),
      "r"(phase)
      : "memory");
#endif
}
// Several bulk copies completing on one mbarrier: announce the total first, then issue the copies.
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes) {
#ifndef EONB200_EMU
  const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void bulk_g2s_issue(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
#ifndef EONB200_EMU
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
#else
  memcpy(smem_dst, gsrc, bytes);
#endif
}
// Ask the memory system to bring [p, p + bytes) into L2 (bytes a multiple of 16, p 16-byte aligned) / one line.
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) {
#ifndef EONB200_EMU
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void l2_prefetch_line(const void *p) {
#ifndef EONB200_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
// Loads that bypass L1 (the line may have been rewritten by another CTA of the cluster during this kernel).
__device__ __forceinline__ int ld_cg(const int *p) {
#ifndef EONB200_EMU
  return __ldcg(p);
#else
  return *p;
#endif
}
__device__ __forceinline__ double ld_cg(const double *p) {
#ifndef EONB200_EMU
  return __ldcg(p);
#else
  return *p;
#endif
}
// ---- shared-memory FP64 accumulation (half-list EAM kernel: both atoms of a pair receive its contribution) -------
// sm_100a has no native shared FP64 add: the instruction is a load / DADD / compare-and-swap loop.  add3 keeps the
// three loops of a force vector in flight together (three loads, three adds, three CAS) and retries a component only
// when its CAS lost a race, which is rare: the lanes of a warp push to distinct atoms nearly always.
__device__ __forceinline__ void smem_add_f64(double *p, double v) {
#ifdef EONB200_EMU
  *p += v;
#else
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
#endif
}
__device__ __forceinline__ void smem_add3_f64(double *p, double x, double y, double z) {
#ifdef EONB200_EMU
  p[0] += x;
  p[1] += y;
  p[2] += z;
#else
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  unsigned long long o0, o1, o2, r0, r1, r2;
  asm volatile("ld.volatile.shared.b64 %0, [%1];" : "=l"(o0) : "r"(a) : "memory");
  asm volatile("ld.volatile.shared.b64 %0, [%1+8];" : "=l"(o1) : "r"(a) : "memory");
  asm volatile("ld.volatile.shared.b64 %0, [%1+16];" : "=l"(o2) : "r"(a) : "memory");
  const unsigned long long n0 = static_cast<unsigned long long>(__double_as_longlong(__longlong_as_double(static_cast<long long>(o0)) + x));
  const unsigned long long n1 = static_cast<unsigned long long>(__double_as_longlong(__longlong_as_double(static_cast<long long>(o1)) + y));
  const unsigned long long n2 = static_cast<unsigned long long>(__double_as_longlong(__longlong_as_double(static_cast<long long>(o2)) + z));
  asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(r0) : "r"(a), "l"(o0), "l"(n0) : "memory");
  asm volatile("atom.shared.cas.b64 %0, [%1+8], %2, %3;" : "=l"(r1) : "r"(a), "l"(o1), "l"(n1) : "memory");
  asm volatile("atom.shared.cas.b64 %0, [%1+16], %2, %3;" : "=l"(r2) : "r"(a), "l"(o2), "l"(n2) : "memory");
  if (r0 != o0) asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory");
  if (r1 != o1) asm volatile("red.shared.add.f64 [%0+8], %1;" ::"r"(a), "d"(y) : "memory");
  if (r2 != o2) asm volatile("red.shared.add.f64 [%0+16], %1;" ::"r"(a), "d"(z) : "memory");
#endif
}
// Barrier + CTA-wide OR of a predicate (one barrier instruction).
__device__ __forceinline__ int cta_any(int p) {
#ifdef EONB200_EMU
  return p;
#else
  return __syncthreads_or(p);
#endif
}
__device__ __forceinline__ int2 ld_cg(const int2 *p) {
#ifndef EONB200_EMU
  return __ldcg(p);
#else
  return *p;
#endif
}
__device__ __forceinline__ int4 ld_cg(const int4 *p) {
#ifndef EONB200_EMU
  return __ldcg(p);
#else
  return *p;
#endif
}
__device__ __forceinline__ double2 ld_cg(const double2 *p) {
#ifndef EONB200_EMU
  return __ldcg(p);
#else
  return *p;
#endif
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
#ifndef EONB200_EMU
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_src));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#else
  memcpy(gdst, smem_src, bytes);
#endif
}

}  // namespace eonb200
