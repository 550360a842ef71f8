#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, int iters) {
  double a[ILP];
  for (int u = 0; u < ILP; u++) a[u] = threadIdx.x * 1e-9 + u;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < ILP; u++) a[u] = fma(a[u], b, c);
  }
  long long t1 = clock64();
  double r = 0; for (int u = 0; u < ILP; u++) r += a[u];
  if (r == 12345.678) out[0] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int warps, double *o, long long *c) {
  int iters = 4096;
  k<ILP><<<148, warps * 32>>>(o, c, iters);
  cudaDeviceSynchronize();
  k<ILP><<<148, warps * 32>>>(o, c, iters);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("ILP=%d warps/SM=%2d  cycles/iter=%.2f  cycles per DFMA-warp-instr per SMSP=%.2f\n", ILP, warps, (double)h / iters,
         (double)h / iters / (ILP * (warps + 3) / 4 ));
}
int main() {
  double *o; long long *c; cudaMalloc(&o, 8); cudaMalloc(&c, 8);
  int ws[] = {1, 4, 8, 16, 32};
  for (int w : ws) { run<1>(w, o, c); run<2>(w, o, c); run<4>(w, o, c); run<8>(w, o, c); }
  return 0;
}
