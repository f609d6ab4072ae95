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

// A filter that stopped early (all weights underflowed, src/model_ssm_ung.cpp:556-560) leaves the
// columns it never reached unwritten; the buffers come from a recycling pool, so they hold stale
// data.  The reference's outputs are zero-filled (arma::fill::zeros, src/R_bootstrap_filter.cpp:162-164)
// and its filter_smoother / summary then run over zeros.  Before any post-processing kernel this
// kernel zero-fills the unreached tail of every stopped filter: state rows stopped+1..n, weight
// columns stopped..n, ancestor columns stopped..n-1 (ancestor 0 is a valid index).
__global__ void zero_unreached_kernel(const int32_t* __restrict__ stopped_at, double* __restrict__ alpha_tm,
                                      double* __restrict__ weights, uint32_t* __restrict__ indices, uint32_t N,
                                      int n, int m, int keep_alpha, size_t alpha_stride, size_t weights_stride,
                                      size_t indices_stride) {
  const uint32_t f = blockIdx.y;
  const int s = stopped_at[f];
  if (s < 0) return;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (keep_alpha && alpha_tm) {
    double* a = alpha_tm + f * alpha_stride;
    for (size_t e = (size_t)(s + 1) * m * N + tid; e < (size_t)(n + 1) * m * N; e += nth) a[e] = 0.0;
  }
  if (weights) {
    double* w = weights + f * weights_stride;
    for (size_t e = (size_t)s * N + tid; e < (size_t)(n + 1) * N; e += nth) w[e] = 0.0;
  }
  if (indices) {
    uint32_t* ix = indices + f * indices_stride;
    for (size_t e = (size_t)s * N + tid; e < (size_t)n * N; e += nth) ix[e] = 0u;
  }
}


// Kitagawa filter-smoother.  Thread i follows the ancestry of particle i of
// the last column backwards: b <- indices[b, t].  Out of place (the reference
// copies each tube before overwriting it, src/filter_smoother.cpp:10).
template <int M>
__global__ void backtrack_kernel(const double* __restrict__ alpha_f, const uint32_t* __restrict__ indices,
                                 double* __restrict__ alpha_s, uint32_t N, int n, size_t alpha_stride,
                                 size_t indices_stride) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  alpha_f += (size_t)blockIdx.y * alpha_stride;   // blockIdx.y = filter
  alpha_s += (size_t)blockIdx.y * alpha_stride;
  indices += (size_t)blockIdx.y * indices_stride;
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
               double* __restrict__ cov_x /* M x M x n1 */, size_t alpha_stride) {
  __shared__ double sm[32 * M * M];
  const int t = blockIdx.x;
  const int n1 = gridDim.x;
  alpha_tm += (size_t)blockIdx.y * alpha_stride;  // blockIdx.y = filter
  mean_x += (size_t)blockIdx.y * M * n1;
  cov_x += (size_t)blockIdx.y * M * M * n1;
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
                      double* __restrict__ Ptt, size_t alpha_stride, size_t weights_stride) {
  alpha_tm += (size_t)blockIdx.y * alpha_stride;  // blockIdx.y = filter
  weights += (size_t)blockIdx.y * weights_stride;
  at += (size_t)blockIdx.y * M * (n + 1);
  att += (size_t)blockIdx.y * M * n;
  Pt += (size_t)blockIdx.y * M * M * (n + 1);
  Ptt += (size_t)blockIdx.y * M * M * n;
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


// ---- the same three kernels for a state dimension given at run time (m > 4: bsm_ng with a
// seasonal component).  The means and the covariance entries are reduced one (or a few) at a time
// with the same fixed-tree block sum, so the results equal the templated kernels' for m <= 4.
__global__ void backtrack_dyn_kernel(const double* __restrict__ alpha_f, const uint32_t* __restrict__ indices,
                                     double* __restrict__ alpha_s, uint32_t N, int n, int m, size_t alpha_stride,
                                     size_t indices_stride) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  alpha_f += (size_t)blockIdx.y * alpha_stride;
  alpha_s += (size_t)blockIdx.y * alpha_stride;
  indices += (size_t)blockIdx.y * indices_stride;
  uint32_t b = i;
  for (int d = 0; d < m; ++d) alpha_s[((size_t)n * m + d) * N + i] = alpha_f[((size_t)n * m + d) * N + i];
  for (int t = n - 1; t >= 0; --t) {
    b = indices[(size_t)t * N + b];
    for (int d = 0; d < m; ++d) alpha_s[((size_t)t * m + d) * N + i] = alpha_f[((size_t)t * m + d) * N + b];
  }
}

// one block per time column; `weights` null: plain summary (mean, cov / N) into (o_mean, o_cov);
// otherwise filter_summary: unweighted (at, Pt) into (o_mean, o_cov) and weighted (att, Ptt) into
// (o_meanw, o_covw) for t < n
__global__ void __launch_bounds__(1024)
summary_dyn_kernel(const double* __restrict__ alpha_tm, const double* __restrict__ weights, uint32_t N, int n, int m,
                   double* __restrict__ o_mean, double* __restrict__ o_cov, double* __restrict__ o_meanw,
                   double* __restrict__ o_covw, size_t alpha_stride, size_t weights_stride) {
  __shared__ double sm[32 * 3];
  __shared__ double mu[16], muw[16];
  const int t = blockIdx.x, n1 = n + 1;
  alpha_tm += (size_t)blockIdx.y * alpha_stride;
  o_mean += (size_t)blockIdx.y * m * n1;
  o_cov += (size_t)blockIdx.y * m * m * n1;
  const bool wtd = weights != nullptr;
  const bool filt = wtd && t < n;
  const double* w = wtd ? weights + (size_t)blockIdx.y * weights_stride + (size_t)t * N : nullptr;
  if (wtd) { o_meanw += (size_t)blockIdx.y * m * n; o_covw += (size_t)blockIdx.y * m * m * n; }
  const double* col = alpha_tm + (size_t)t * m * N;
  double sw = 1.0;
  if (wtd) {
    double a[1] = {0.0};
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) a[0] += w[i];
    block_sum<1>(a, sm);
    sw = a[0];
  }
  for (int d = 0; d < m; ++d) {
    double a[2] = {0.0, 0.0};
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
      const double x = col[(size_t)d * N + i];
      a[0] += x;
      if (wtd) a[1] += x * w[i];
    }
    block_sum<2>(a, sm);
    if (threadIdx.x == 0) { mu[d] = a[0] / (double)N; muw[d] = wtd ? a[1] / sw : 0.0; }
  }
  __syncthreads();
  for (int cc = 0; cc < m; ++cc)
    for (int r = 0; r < m; ++r) {
      double a[2] = {0.0, 0.0};
      for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double xr = col[(size_t)r * N + i], xc = col[(size_t)cc * N + i];
        a[0] += (xr - mu[r]) * (xc - mu[cc]);
        if (wtd) a[1] += (w[i] / sw) * (xr - muw[r]) * (xc - muw[cc]);
      }
      block_sum<2>(a, sm);
      if (threadIdx.x == 0) {
        o_cov[r + (size_t)m * cc + (size_t)m * m * t] = a[0] / (double)N;
        if (filt) o_covw[r + (size_t)m * cc + (size_t)m * m * t] = a[1];
      }
    }
  if (threadIdx.x == 0)
    for (int d = 0; d < m; ++d) {
      o_mean[d + (size_t)m * t] = mu[d];
      if (filt) o_meanw[d + (size_t)m * t] = muw[d];
    }
}

// ssm_ulg::psi_filter (src/model_ssm_ulg.cpp:646-672): N independent draws from the smoothing
// distribution, alpha_{t+1} = ahat_{t+1} + Ct_{t+1} (alpha_t - ahat_t) + L_{t+1} u.  No weights,
// no resampling: one thread per trajectory, written in the time-major layout.
template <int M>
__global__ void gaussian_psi_kernel(const double* __restrict__ alphahat, const double* __restrict__ Lt,
                                    const double* __restrict__ Ct, double* __restrict__ alpha_tm, uint32_t N,
                                    int n, uint32_t seed) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double x[M], z[M];
  for (int t = 0; t <= n; ++t) {
#pragma unroll
    for (int d = 0; d < M; ++d) z[d] = pf_normal(seed, 0u, (uint32_t)t, 0u, i * M + d);
    double nx[M];
#pragma unroll
    for (int d = 0; d < M; ++d) {
      double cd = 0.0, lu = 0.0;
      if (t > 0) {
#pragma unroll
        for (int j = 0; j < M; ++j) cd += Ct[d + M * j + (size_t)M * M * t] * (x[j] - alphahat[j + (size_t)M * (t - 1)]);
      }
#pragma unroll
      for (int j = 0; j < M; ++j) lu += Lt[d + M * j + (size_t)M * M * t] * z[j];
      nx[d] = t > 0 ? alphahat[d + (size_t)M * t] + cd + lu : alphahat[d] + lu;
    }
#pragma unroll
    for (int d = 0; d < M; ++d) { x[d] = nx[d]; alpha_tm[((size_t)t * M + d) * N + i] = nx[d]; }
  }
}

// post_correct output_type = 1 (src/approx_mcmc.cpp:388-392): after the
// backward tracing, ONE trajectory per draw is kept, chosen uniformly.  Only
// that path is traced: thread f follows the ancestry of particle
// idx = floor(u * N), u = pf_pick_uniform(seed, filter).  Output per filter is
// the (n+1) x M matrix alpha_i.slice(idx).t(), column-major.
template <int M>
__global__ void pick_trajectory_kernel(const double* __restrict__ alpha_f, const uint32_t* __restrict__ indices,
                                       double* __restrict__ traj, uint32_t N, int n, size_t alpha_stride,
                                       size_t indices_stride, uint32_t seed, uint32_t first_filter,
                                       uint32_t n_filters) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_filters) return;
  alpha_f += (size_t)f * alpha_stride;
  indices += (size_t)f * indices_stride;
  traj += (size_t)f * (n + 1) * M;
  uint32_t b = (uint32_t)(pf_pick_uniform(seed, first_filter + f) * (double)N);
  if (b > N - 1) b = N - 1;
#pragma unroll
  for (int d = 0; d < M; ++d) traj[n + (size_t)(n + 1) * d] = alpha_f[((size_t)n * M + d) * N + b];
  for (int t = n - 1; t >= 0; --t) {
    b = indices[(size_t)t * N + b];
#pragma unroll
    for (int d = 0; d < M; ++d) traj[t + (size_t)(n + 1) * d] = alpha_f[((size_t)t * M + d) * N + b];
  }
}

// post_correct output_type = 2 (src/approx_mcmc.cpp:394-424): running weighted
// mean / variance over the draws, in draw order (the reference's order with one
// thread; with several threads its critical-section order is nondeterministic).
// One thread per element of alphahat (M x n1); sequential over draws.
//   wnew = w_i * count_i; sum_w += wnew; diff = ah_i - ah; ah += wnew/sum_w*diff;
//   Valpha += wnew * diff * (ah_i - ah)'; Vt += wnew/sum_w * (Vt_i - Vt)
// state: sum_w, alphahat (M x n1), Vt (M x M x n1), Valpha (M x M x n1); updated in place so
// that successive chunks of draws continue the same recursion.
template <int M>
__global__ void postcorr_merge_kernel(const double* __restrict__ mean_i, const double* __restrict__ cov_i,
                                      const double* __restrict__ loglik, const uint32_t* __restrict__ counts,
                                      uint32_t n_filters, int n1, const double* __restrict__ sum_w_in,
                                      double* __restrict__ sum_w_out,
                                      double* __restrict__ alphahat, double* __restrict__ Vt,
                                      double* __restrict__ Valpha) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1) return;
  double sum_w = *sum_w_in;
  double ah[M], V[M * M], Va[M * M];
#pragma unroll
  for (int d = 0; d < M; ++d) ah[d] = alphahat[d + (size_t)M * t];
#pragma unroll
  for (int e = 0; e < M * M; ++e) { V[e] = Vt[e + (size_t)M * M * t]; Va[e] = Valpha[e + (size_t)M * M * t]; }
  for (uint32_t i = 0; i < n_filters; ++i) {
    const double wnew = pf_exp(loglik[i]) * (double)counts[i];
    sum_w += wnew;
    const double* mi = mean_i + (size_t)i * M * n1 + (size_t)M * t;
    const double* ci = cov_i + (size_t)i * M * M * n1 + (size_t)M * M * t;
    double diff[M], diff2[M];
    const double f = wnew / sum_w;
#pragma unroll
    for (int d = 0; d < M; ++d) { diff[d] = mi[d] - ah[d]; ah[d] += f * diff[d]; diff2[d] = mi[d] - ah[d]; }
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int r = 0; r < M; ++r) {
        Va[r + M * c] += wnew * diff[r] * diff2[c];
        V[r + M * c] += f * (ci[r + M * c] - V[r + M * c]);
      }
  }
#pragma unroll
  for (int d = 0; d < M; ++d) alphahat[d + (size_t)M * t] = ah[d];
#pragma unroll
  for (int e = 0; e < M * M; ++e) { Vt[e + (size_t)M * M * t] = V[e]; Valpha[e + (size_t)M * M * t] = Va[e]; }
  if (t == n1 - 1) *sum_w_out = sum_w;  // every thread computed the same sum_w
}

// The same two for a run-time state dimension (m > 3).  The merge runs one thread per element (r, c, t) of the
// covariance; each thread carries the recursions of alphahat_r and alphahat_c itself (same operations, same
// order as above), reads the state from *_in and writes *_out so that no thread reads what another has updated.
__global__ void pick_trajectory_dyn_kernel(const double* __restrict__ alpha_f, const uint32_t* __restrict__ indices,
                                           double* __restrict__ traj, uint32_t N, int n, int m, size_t alpha_stride,
                                           size_t indices_stride, uint32_t seed, uint32_t first_filter,
                                           uint32_t n_filters) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_filters) return;
  alpha_f += (size_t)f * alpha_stride;
  indices += (size_t)f * indices_stride;
  traj += (size_t)f * (n + 1) * m;
  uint32_t b = (uint32_t)(pf_pick_uniform(seed, first_filter + f) * (double)N);
  if (b > N - 1) b = N - 1;
  for (int d = 0; d < m; ++d) traj[n + (size_t)(n + 1) * d] = alpha_f[((size_t)n * m + d) * N + b];
  for (int t = n - 1; t >= 0; --t) {
    b = indices[(size_t)t * N + b];
    for (int d = 0; d < m; ++d) traj[t + (size_t)(n + 1) * d] = alpha_f[((size_t)t * m + d) * N + b];
  }
}
__global__ void postcorr_merge_dyn_kernel(const double* __restrict__ mean_i, const double* __restrict__ cov_i,
                                          const double* __restrict__ loglik, const uint32_t* __restrict__ counts,
                                          uint32_t n_filters, int n1, int m, const double* __restrict__ sum_w_in,
                                          double* __restrict__ sum_w_out, const double* __restrict__ ah_in,
                                          double* __restrict__ ah_out, double* __restrict__ Vt,
                                          double* __restrict__ Valpha) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x, total = (size_t)m * m * n1;
  if (idx >= total) return;
  const int r = (int)(idx % m), c = (int)((idx / m) % m);
  const size_t t = idx / ((size_t)m * m);
  double sum_w = *sum_w_in, ar = ah_in[r + (size_t)m * t], ac = ah_in[c + (size_t)m * t], V = Vt[idx], Va = Valpha[idx];
  for (uint32_t i = 0; i < n_filters; ++i) {
    const double wnew = pf_exp(loglik[i]) * (double)counts[i];
    sum_w += wnew;
    const double* mi = mean_i + (size_t)i * m * n1 + (size_t)m * t;
    const double ci = cov_i[(size_t)i * m * m * n1 + idx];
    const double f = wnew / sum_w;
    const double dr = mi[r] - ar;
    ar += f * dr;
    const double dc = mi[c] - ac;
    ac += f * dc;
    const double d2c = mi[c] - ac;
    Va += wnew * dr * d2c;
    V += f * (ci - V);
  }
  Vt[idx] = V; Valpha[idx] = Va;
  if (c == 0) ah_out[r + (size_t)m * t] = ar;
  if (idx == total - 1) *sum_w_out = sum_w;
}

// ------------------------------------------------ layout change for output
// alpha_tm [(t*M+d)*N + i]  ->  cube [d + M*(t + n1*i)]  (Armadillo cube
// m x (n+1) x N, src/R_bootstrap_filter.cpp:162).  A plain tiled transpose of
// the (n1*M) x N matrix; both sides coalesced.
// Rows [rbeg, rend) only: the streaming fetch transposes a block of finished
// time points while the filter kernel is still running.
__global__ void transpose_to_cube_kernel(const double* __restrict__ tm, double* __restrict__ cube,
                                         uint32_t rows /* n1*M */, uint32_t N, uint32_t rbeg, uint32_t rend) {
  __shared__ double tile[32][33];
  const uint32_t r0 = rbeg + blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    uint32_t r = r0 + k, c = c0 + threadIdx.x;
    if (r < rend && c < N) tile[k][threadIdx.x] = tm[(size_t)r * N + c];
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    uint32_t c = c0 + k, r = r0 + threadIdx.x;
    if (r < rend && c < N) cube[(size_t)c * rows + r] = tile[threadIdx.x][k];
  }
}

} // namespace pf
