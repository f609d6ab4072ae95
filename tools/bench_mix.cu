// do FP64 arithmetic and conversion (XU) instructions overlap on B200?  Three kernels at 1024 threads per SM:
// A: NF dependent-free DFMA per iteration; B: NX conversions (I2F.F64.U32) per iteration; C: both.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NX>
__global__ void k(double* out, long long* cyc, int iters, double a, double b, unsigned seed) {
  double x[8]; unsigned v[4];
  for (int i = 0; i < 8; ++i) x[i] = a + i + threadIdx.x * 1e-9;
  for (int i = 0; i < 4; ++i) v[i] = seed + threadIdx.x + i;
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NF; ++i) x[i & 7] = __fma_rn(x[i & 7], b, a);
#pragma unroll
    for (int i = 0; i < NX; ++i) { acc += (double)v[i & 3]; v[i & 3] = v[i & 3] * 1664525u + 1013904223u; }
  }
  long long t1 = clock64();
  double s = acc;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NF, int NX> void run() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) k<NF, NX><<<148, 1024>>>(out, cyc, iters, 1.000001, 0.999999, 12345u);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA per iter %2d, I2F(+DADD, IMAD) per iter %d : %.1f cycles per iteration per SM-scheduler (8 warps)\n", NF, NX, (double)c / iters);
  cudaFree(out); cudaFree(cyc);
}
int main() { run<16, 0>(); run<0, 2>(); run<16, 2>(); run<32, 0>(); run<0, 4>(); run<32, 4>(); run<0, 1>(); run<8, 1>(); return 0; }
