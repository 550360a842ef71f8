// tests/emu/emu_shim.h -- TEST INFRASTRUCTURE ONLY.
//
// Lets the CPU-only test-suite compile eon_b200/csrc/*.cu{,h} as plain C++ (g++ -x c++
// -DEONB200_EMU) and run every kernel with ONE thread per CTA, CTAs executed one after the other.
// It exists so that the control flow of the device code (cell walk, Verlet cache, dimer / Lanczos /
// optimizer state machines, the C ABI's buffer handling) can be exercised where no GPU exists.
// It is NOT a backend: the eon_b200 package never loads the emulation library, libeonb200.so does
// not contain it, and nothing timed or shipped runs through it.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline

#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct emu_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
namespace emu {
inline emu_dim3 &tidx() {
  static emu_dim3 v;
  return v;
}
inline emu_dim3 &bidx() {
  static emu_dim3 v;
  return v;
}
inline emu_dim3 &bdim() {
  static emu_dim3 v;
  return v;
}
inline emu_dim3 &gdim() {
  static emu_dim3 v;
  return v;
}
inline std::vector<unsigned char> &smem_store() {
  static std::vector<unsigned char> v;
  return v;
}
inline unsigned char *dyn_smem() {
  // 128-byte aligned view
  auto &v = smem_store();
  uintptr_t p = reinterpret_cast<uintptr_t>(v.data());
  return reinterpret_cast<unsigned char *>((p + 127) / 128 * 128);
}
}  // namespace emu
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
namespace emu {
template <typename F>
inline void launch(dim3 grid, size_t smem, F &&body) {
  smem_store().assign(smem + 256, 0);
  tidx().x = 0;
  bdim().x = 1;
  gdim().x = grid.x;
  gdim().y = grid.y;
  for (unsigned by = 0; by < grid.y; by++)
    for (unsigned b = 0; b < grid.x; b++) {
      bidx().x = b;
      bidx().y = by;
      body();
    }
}
}  // namespace emu
#define threadIdx (emu::tidx())
#define blockIdx (emu::bidx())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())

inline void __syncthreads() {}
inline void __syncwarp() {}
inline double __dmul_rn(double a, double b) {
  volatile double r = a * b;
  return r;
}
inline double __dadd_rn(double a, double b) {
  volatile double r = a + b;
  return r;
}
inline double __dsub_rn(double a, double b) {
  volatile double r = a - b;
  return r;
}
inline int atomicAdd(int *p, int v) {
  int o = *p;
  *p += v;
  return o;
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
  unsigned long long o = *p;
  *p += v;
  return o;
}
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
  unsigned long long o = *p;
  if (v > o) *p = v;
  return o;
}
inline int atomicMax(int *p, int v) {
  int o = *p;
  if (v > o) *p = v;
  return o;
}
struct double4 {
  double x, y, z, w;
};
inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
struct float4 {
  float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct int2 {
  int x, y;
};
inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct alignas(16) int4 {
  int x, y, z, w;
};
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline int atomicOr(int *p, int v) {
  int o = *p;
  *p |= v;
  return o;
}
struct double2 {
  double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
using std::max;
using std::min;

// ---- the sliver of the CUDA runtime API capi.cu uses -------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void *cudaStream_t;
typedef std::chrono::steady_clock::time_point *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
  int major = 10, minor = 0, multiProcessorCount = 1;
  size_t sharedMemPerBlockOptin = 16u << 20;
};
struct cudaFuncAttributes {
  size_t sharedSizeBytes = 0;
};
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) {
  *n = 1;
  return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  *p = cudaDeviceProp();
  return cudaSuccess;
}
inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, const void *) {
  *a = cudaFuncAttributes();
  return cudaSuccess;
}
inline cudaError_t cudaFuncSetAttribute(const void *, int, int) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) {
  *p = calloc(1, n + 64);
  return *p ? cudaSuccess : 2;
}
inline cudaError_t cudaFree(void *p) {
  free(p);
  return cudaSuccess;
}
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
  memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) {
  return cudaMemcpy(d, s, n, k);
}
inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
  for (size_t r = 0; r < h; r++) memmove(static_cast<char *>(d) + r * dp, static_cast<const char *>(s) + r * sp, w);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k,
                                     cudaStream_t) {
  return cudaMemcpy2D(d, dp, s, sp, w, h, k);
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
  memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreate(cudaStream_t *s) {
  *s = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
  *e = new std::chrono::steady_clock::time_point();
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  *e = std::chrono::steady_clock::now();
  return cudaSuccess;
}
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(*b - *a).count();
  return cudaSuccess;
}
