// smoother_kernels.cuh -- backward ancestor tracing and particle summaries.
//
//   backtrack_kernel       filter_smoother  /root/reference/src/filter_smoother.cpp:5-18
//   summary_kernel         summary          src/summary.cpp:4-17
//   filter_summary_kernel  filter_summary   src/summary.cpp:42-79
//
// All three work on the time-major SoA layout [(t*M + d)*N + i] the filter
// kernel writes, so every access is coalesced (thread <-> particle).  The
// reference's Welford recursion is sequential in the particle index; here the
// mean and the centred second moments are two deterministic fixed-tree block
// reductions (same result on every run; equal to Welford up to rounding --
// the parity tests use a 1e-10 relative tolerance for these outputs).
#pragma once
#include "pf_device.cuh"

namespace pf {

// Kitagawa filter-smoother.  Thread i follows the ancestry of particle i of
// the last column backwards: b <- indices[b, t].  Out of place (the reference
// copies each tube before overwriting it, src/filter_smoother.cpp:10).
template <int M>
__global__ void backtrack_kernel(const double* __restrict__ alpha_f, const uint32_t* __restrict__ indices,
                                 double* __restrict__ alpha_s, uint32_t N, int n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint32_t b = i;
#pragma unroll
  for (int d = 0; d < M; ++d) alpha_s[((size_t)n * M + d) * N + i] = alpha_f[((size_t)n * M + d) * N + i];
  for (int t = n - 1; t >= 0; --t) {
    b = indices[(size_t)t * N + b];
#pragma unroll
    for (int d = 0; d < M; ++d) alpha_s[((size_t)t * M + d) * N + i] = alpha_f[((size_t)t * M + d) * N + b];
  }
}

// fixed-tree block sum of V doubles per thread; result valid in all threads
template <int V>
__device__ __forceinline__ void block_sum(double (&v)[V], double* smem /* 32*V */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int e = 0; e < V; ++e)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int e = 0; e < V; ++e) smem[warp * V + e] = v[e];
  __syncthreads();
#pragma unroll
  for (int e = 0; e < V; ++e) {
    double x = lane < nw ? smem[lane * V + e] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    v[e] = x;
  }
}

// one block per time column: mean over particles and covariance / N
template <int M>
__global__ void __launch_bounds__(1024)
summary_kernel(const double* __restrict__ alpha_tm, uint32_t N, double* __restrict__ mean_x /* M x n1 */,
               double* __restrict__ cov_x /* M x M x n1 */) {
  __shared__ double sm[32 * M * M];
  const int t = blockIdx.x;
  const double* col = alpha_tm + (size_t)t * M * N;
  double s[M];
#pragma unroll
  for (int d = 0; d < M; ++d) s[d] = 0.0;
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
#pragma unroll
    for (int d = 0; d < M; ++d) s[d] += col[(size_t)d * N + i];
  block_sum<M>(s, sm);
  double mu[M];
#pragma unroll
  for (int d = 0; d < M; ++d) mu[d] = s[d] / (double)N;
  double c[M * M];
#pragma unroll
  for (int e = 0; e < M * M; ++e) c[e] = 0.0;
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
    double dv[M];
#pragma unroll
    for (int d = 0; d < M; ++d) dv[d] = col[(size_t)d * N + i] - mu[d];
#pragma unroll
    for (int cc = 0; cc < M; ++cc)
#pragma unroll
      for (int r = 0; r < M; ++r) c[r + M * cc] += dv[r] * dv[cc];
  }
  block_sum<M * M>(c, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < M; ++d) mean_x[d + (size_t)M * t] = mu[d];
#pragma unroll
    for (int e = 0; e < M * M; ++e) cov_x[e + (size_t)M * M * t] = c[e] / (double)N;
  }
}

// one block per time column t = 0..n: predicted (at, Pt: unweighted) and, for
// t < n, filtered (att, Ptt: weighted by the normalised weights) moments.
template <int M>
__global__ void __launch_bounds__(1024)
filter_summary_kernel(const double* __restrict__ alpha_tm, const double* __restrict__ weights, uint32_t N,
                      int n, double* __restrict__ at, double* __restrict__ att, double* __restrict__ Pt,
                      double* __restrict__ Ptt) {
  constexpr int SMV = (2 * M + 1 > 2 * M * M) ? (2 * M + 1) : (2 * M * M);
  __shared__ double sm[32 * SMV];
  const int t = blockIdx.x;
  const bool filt = t < n;
  const double* col = alpha_tm + (size_t)t * M * N;
  const double* w = weights + (size_t)t * N;
  double acc[2 * M + 1];
#pragma unroll
  for (int e = 0; e < 2 * M + 1; ++e) acc[e] = 0.0;
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
    const double wi = w[i];
    acc[2 * M] += wi;
#pragma unroll
    for (int d = 0; d < M; ++d) {
      const double x = col[(size_t)d * N + i];
      acc[d] += x;
      acc[M + d] += x * wi;
    }
  }
  block_sum<2 * M + 1>(acc, sm);
  const double sw = acc[2 * M];
  double mu[M], muw[M];
#pragma unroll
  for (int d = 0; d < M; ++d) { mu[d] = acc[d] / (double)N; muw[d] = acc[M + d] / sw; }
  double c[2 * M * M];
#pragma unroll
  for (int e = 0; e < 2 * M * M; ++e) c[e] = 0.0;
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
    const double wi = w[i] / sw;
    double dv[M], dw[M];
#pragma unroll
    for (int d = 0; d < M; ++d) {
      const double x = col[(size_t)d * N + i];
      dv[d] = x - mu[d];
      dw[d] = x - muw[d];
    }
#pragma unroll
    for (int cc = 0; cc < M; ++cc)
#pragma unroll
      for (int r = 0; r < M; ++r) {
        c[r + M * cc] += dv[r] * dv[cc];
        c[M * M + r + M * cc] += wi * dw[r] * dw[cc];
      }
  }
  block_sum<2 * M * M>(c, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < M; ++d) at[d + (size_t)M * t] = mu[d];
#pragma unroll
    for (int e = 0; e < M * M; ++e) Pt[e + (size_t)M * M * t] = c[e] / (double)N;
    if (filt) {
#pragma unroll
      for (int d = 0; d < M; ++d) att[d + (size_t)M * t] = muw[d];
#pragma unroll
      for (int e = 0; e < M * M; ++e) Ptt[e + (size_t)M * M * t] = c[M * M + e];
    }
  }
}

} // namespace pf
