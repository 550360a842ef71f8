// Shared-memory accumulate micro-benchmark (development aid): cycles per warp-wide operation at SM level for
//   0 red.shared.add.f64 (CAS loop)   1 atom.shared.add.u64 (native)   2 raw atom.shared.cas.b64   3 ld.shared.b64
//   4 st.shared.b64   5 red.shared.add.f32 (native?)   6 LDS + DADD + STS (non-atomic RMW)
// address patterns: 0 lane-distinct consecutive, 1 pseudo-random distinct-ish, 2 all lanes the same word
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(int op, int pat, int iters, long long *out, double *sink) {
  extern __shared__ double sh[];
  const int n = 4096;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  unsigned idx = threadIdx.x;
  unsigned rng = threadIdx.x * 2654435761u + 12345u;
  double acc = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    unsigned a;
    if (pat == 0) a = (idx + it * 32) & (n - 1);
    else if (pat == 1) { rng = rng * 1664525u + 1013904223u; a = (rng >> 12) & (n - 1); }
    else a = (it * 7) & (n - 1);
    const uint32_t p = static_cast<uint32_t>(__cvta_generic_to_shared(sh + a));
    if (op == 0) asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(p), "d"(1.0) : "memory");
    else if (op == 1) asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(p), "l"(1ull) : "memory");
    else if (op == 2) { unsigned long long r; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(r) : "r"(p), "l"(0ull), "l"(0ull) : "memory"); acc += (double)r; }
    else if (op == 3) { double r; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(r) : "r"(p) : "memory"); acc += r; }
    else if (op == 4) asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(p), "d"(1.0) : "memory");
    else if (op == 5) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(p), "f"(1.0f) : "memory");
    else { double r; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(r) : "r"(p) : "memory"); r += 1.0; asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(p), "d"(r) : "memory"); }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 1234.5) sink[0] = acc + sh[idx & (n - 1)];
}
int main() {
  long long *out; double *sink;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&sink, 8);
  const char *opn[] = {"red.add.f64(CAS loop)", "red.add.u64", "atom.cas.b64", "ld.b64", "st.b64", "red.add.f32", "ld+dadd+st"};
  const char *pn[] = {"consecutive", "random", "same"};
  for (int threads : {32, 416, 832}) {
    for (int op = 0; op < 7; op++) for (int pat = 0; pat < 3; pat++) {
      const int iters = 2000;
      k<<<148, threads, 4096 * 8>>>(op, pat, iters, out, sink);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double cyc = (double)h[0] / iters;
      printf("threads %4d  %-22s %-12s  %.1f cycles/iter (warp 0)  -> %.2f SM cycles per warp-op\n", threads, opn[op], pn[pat], cyc, cyc / (threads / 32));
    }
  }
  return 0;
}
