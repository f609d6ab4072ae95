// The one place where the north star allows tensor cores: the batched state transition
// X (N x m) . T' (m x m) of the multivariate / seasonal models.  This measures it at m = 13 (monthly
// bsm_ng; padded to 16 for the MMA) on B200, fp64, operands in registers / shared memory:
//   A  FP64 pipe: thread <-> particle, 13 x 13 = 169 DFMA per particle (T in shared memory)
//   B  DMMA:      mma.sync.aligned.m8n8k4.f64 (the fp64 tensor-core path that exists on sm_100a;
//                 tcgen05.mma has no fp64 kind), a warp transforms 32 particles with 32 MMAs (16 x 16 padded)
// Both variants are checked against a CPU product; the timing loop re-applies the transition ITERS times.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bench_mma.bin tools/bench_mma.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
constexpr int M = 13, MP = 16, ITERS = 256;

__global__ void k_fp64(const double* __restrict__ X, const double* __restrict__ T, double* __restrict__ Y, int N) {
  __shared__ double Ts[M * M];
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) Ts[e] = T[e];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double x[M], y[M];
#pragma unroll
  for (int d = 0; d < M; ++d) x[d] = X[(size_t)d * N + i];
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int d = 0; d < M; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) s = fma(Ts[d + M * j], x[j], s);
      y[d] = s;
    }
#pragma unroll
    for (int d = 0; d < M; ++d) x[d] = y[d];
  }
#pragma unroll
  for (int d = 0; d < M; ++d) Y[(size_t)d * N + i] = x[d];
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// a warp owns 32 particles = 4 tiles of 8; state tile kept in shared memory between iterations
// (the fragment layouts of A and D differ), T' fragments in registers
__global__ void k_dmma(const double* __restrict__ X, const double* __restrict__ T, double* __restrict__ Y, int N) {
  extern __shared__ double sm[]; // per warp: 32 x MP state tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* xs = sm + (size_t)warp * 32 * MP;
  const int i0 = (blockIdx.x * (blockDim.x >> 5) + warp) * 32;
  for (int e = lane; e < 32 * MP; e += 32) {
    const int p = e / MP, d = e % MP;
    xs[e] = (d < M && i0 + p < N) ? X[(size_t)d * N + i0 + p] : 0.0;
  }
  // B fragments: B[k][n] = T'(k, n) = T[n + M * k]; step s, n-tile q: thread holds B[4s + lane%4][8q + lane/4]
  double bf[4][2];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int k = 4 * s + (lane & 3), nn = 8 * q + (lane >> 2);
      bf[s][q] = (k < M && nn < M) ? T[nn + M * k] : 0.0;
    }
  __syncwarp();
  for (int it = 0; it < ITERS; ++it) {
    double acc[4][2][2];
#pragma unroll
    for (int tl = 0; tl < 4; ++tl)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        acc[tl][q][0] = 0.0; acc[tl][q][1] = 0.0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const double a = xs[(8 * tl + (lane >> 2)) * MP + 4 * s + (lane & 3)]; // A[row = lane/4][col = lane%4]
          dmma(acc[tl][q][0], acc[tl][q][1], a, bf[s][q]);
        }
      }
    __syncwarp();
#pragma unroll
    for (int tl = 0; tl < 4; ++tl)
#pragma unroll
      for (int q = 0; q < 2; ++q) { // D[row = lane/4][cols 2*(lane%4), +1] of n-tile q
        xs[(8 * tl + (lane >> 2)) * MP + 8 * q + 2 * (lane & 3)] = acc[tl][q][0];
        xs[(8 * tl + (lane >> 2)) * MP + 8 * q + 2 * (lane & 3) + 1] = acc[tl][q][1];
      }
    __syncwarp();
  }
  for (int e = lane; e < 32 * MP; e += 32) {
    const int p = e / MP, d = e % MP;
    if (d < M && i0 + p < N) Y[(size_t)d * N + i0 + p] = xs[e];
  }
}

int main() {
  const int N = 148 * 1024 * 4;
  std::vector<double> hX((size_t)M * N), hT(M * M), hY((size_t)M * N), hY2((size_t)M * N);
  srand(1);
  for (auto& v : hX) v = rand() / (double)RAND_MAX - 0.5;
  for (int d = 0; d < M; ++d) for (int j = 0; j < M; ++j) hT[d + M * j] = (d == j ? 0.9 : 0.0) + 0.01 * (rand() / (double)RAND_MAX - 0.5);
  double *X, *T, *Y;
  cudaMalloc(&X, sizeof(double) * M * N); cudaMalloc(&T, sizeof(double) * M * M); cudaMalloc(&Y, sizeof(double) * M * N);
  cudaMemcpy(X, hX.data(), sizeof(double) * M * N, cudaMemcpyHostToDevice);
  cudaMemcpy(T, hT.data(), sizeof(double) * M * M, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms[2] = {1e9f, 1e9f};
  const size_t smem = (size_t)8 * 32 * MP * sizeof(double);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); k_fp64<<<N / 256, 256>>>(X, T, Y, N); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float t; cudaEventElapsedTime(&t, e0, e1); ms[0] = t < ms[0] ? t : ms[0];
  }
  cudaMemcpy(hY.data(), Y, sizeof(double) * M * N, cudaMemcpyDeviceToHost);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); k_dmma<<<N / 256, 256, smem>>>(X, T, Y, N); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float t; cudaEventElapsedTime(&t, e0, e1); ms[1] = t < ms[1] ? t : ms[1];
  }
  cudaError_t err = cudaGetLastError();
  cudaMemcpy(hY2.data(), Y, sizeof(double) * M * N, cudaMemcpyDeviceToHost);
  // CPU check on a few particles
  double maxrel[2] = {0, 0};
  for (int i = 0; i < 64; ++i) {
    double x[M], y[M];
    for (int d = 0; d < M; ++d) x[d] = hX[(size_t)d * N + i];
    for (int it = 0; it < ITERS; ++it) {
      for (int d = 0; d < M; ++d) { double s = 0; for (int j = 0; j < M; ++j) s = fma(hT[d + M * j], x[j], s); y[d] = s; }
      for (int d = 0; d < M; ++d) x[d] = y[d];
    }
    for (int d = 0; d < M; ++d) {
      const double den = fabs(x[d]) + 1e-300;
      maxrel[0] = fmax(maxrel[0], fabs(hY[(size_t)d * N + i] - x[d]) / den);
      maxrel[1] = fmax(maxrel[1], fabs(hY2[(size_t)d * N + i] - x[d]) / den);
    }
  }
  const double tr = (double)N * ITERS; // particle-transitions
  printf("m = %d, N = %d particles, %d transitions each (cuda status: %s)\n", M, N, ITERS, cudaGetErrorString(err));
  printf("FP64 pipe (169 DFMA / particle)      : %.3f ms, %.3e transitions/s, %.2f TFLOP/s useful, max rel err vs CPU %.1e\n",
         ms[0], tr / (ms[0] * 1e-3), 2.0 * M * M * tr / (ms[0] * 1e-3) / 1e12, maxrel[0]);
  printf("DMMA m8n8k4.f64 (16 x 16 padded)     : %.3f ms, %.3e transitions/s, %.2f TFLOP/s useful, max rel err vs CPU %.1e\n",
         ms[1], tr / (ms[1] * 1e-3), 2.0 * M * M * tr / (ms[1] * 1e-3) / 1e12, maxrel[1]);
  return 0;
}
