// micro-benchmark: dependent-issue latency and throughput of the FP64 / conversion instructions the
// particle loop is made of (B200).  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bench_lat.bin tools/bench_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) x[i] = __fma_rn(x[i], b, a);
      if (OP == 1) x[i] = __dmul_rn(x[i], b);
      if (OP == 2) x[i] = __dadd_rn(x[i], b);
      if (OP == 3) x[i] = (double)(long long)x[i] + a;        // F2I.S64 + I2F.S64 + DADD
      if (OP == 4) x[i] = (double)(unsigned)(int)x[i] + a;    // F2I.S32 + I2F.U32 + DADD
      if (OP == 5) x[i] = floor(x[i]) + a;                     // FRND + DADD
      if (OP == 6) x[i] = a / x[i];                            // full division
      if (OP == 7) x[i] = sqrt(x[i] + a);                      // sqrt
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP>
void run(const char* name, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<OP, ILP><<<148, threads>>>(out, cyc, iters, 1.000001, 0.999999);
  k<OP, ILP><<<148, threads>>>(out, cyc, iters, 1.000001, 0.999999);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s ILP=%d threads/SM=%4d : %.2f cycles per op per warp, %.2f warp-ops/clk/SM\n", name, ILP, threads,
         (double)c / iters / ILP, (double)iters * ILP * (threads / 32) / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 1>("DFMA", 32); run<0, 2>("DFMA", 32); run<0, 4>("DFMA", 32); run<0, 8>("DFMA", 32);
  run<0, 1>("DFMA", 256); run<0, 2>("DFMA", 256); run<0, 1>("DFMA", 1024); run<0, 2>("DFMA", 1024); run<0, 4>("DFMA", 1024);
  run<1, 1>("DMUL", 32); run<2, 1>("DADD", 32);
  run<3, 1>("F2I.S64+I2F.S64+DADD", 32); run<3, 1>("F2I.S64+I2F.S64+DADD", 1024); run<3, 4>("F2I.S64+I2F.S64+DADD", 1024);
  run<4, 1>("F2I.S32+I2F.U32+DADD", 32); run<4, 4>("F2I.S32+I2F.U32+DADD", 1024);
  run<5, 1>("FRND+DADD", 32); run<5, 4>("FRND+DADD", 1024);
  run<6, 1>("DDIV", 32); run<6, 4>("DDIV", 1024);
  run<7, 1>("DADD+DSQRT", 32); run<7, 4>("DADD+DSQRT", 1024);
  return 0;
}
