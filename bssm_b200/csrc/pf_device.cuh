// pf_device.cuh -- the per-timestep particle loop as one persistent kernel.
//
// One FILTER is run by a group of `ranks` CTAs: a single CTA (small N), a
// thread-block cluster (<= 16 CTAs, hardware cluster barrier) or the whole
// grid (cooperative launch, software barrier).  Every thread owns P contiguous
// particles whose state, log-weights and weights stay in registers for the
// whole time step; the only per-step global traffic is the coalesced write of
// the outputs (state row, weight column, ancestor column) and the ancestor
// gather, which reads the just-written state row back from L2.
//
// Replaces the scalar loops of ssm_ung::bsf_filter / psi_filter
// (/root/reference/src/model_ssm_ung.cpp:611-712, :511-609) and their
// ssm_nlg / ssm_sde clones, including stratified_sample
// (src/stratified_sample.cpp:9-28).
//
// Arithmetic that depends on summation order in the reference (sum of weights,
// cumulative sum of normalised weights) is done in the order-independent
// fixed-point form of include/bssm_b200/pf_math.h, so results do not depend on
// P, block size, number of ranks or number of GPUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "bssm_b200/pf_math.h"

namespace pf {

enum SyncMode { SYNC_BLOCK = 0, SYNC_CLUSTER = 1, SYNC_GRID = 2 };

struct FilterParams {
  uint32_t N;            // particles per filter
  int n;                 // time points
  uint32_t seed;
  uint32_t first_filter; // random-stream id of filter 0 of this launch
  uint32_t ranks;        // CTAs per filter
  int sync_mode;
  int keep_alpha;        // 1: alpha_tm has n+1 rows, 0: 2-row ring
  // outputs (device).  alpha_tm: time-major SoA [(t*M + d) * N + i] -- the
  // coalesced native layout; the reference's m x (n+1) x N cube is produced by
  // transpose_to_cube_kernel when asked for.
  double* alpha_tm;  size_t alpha_stride;    // per-filter stride (elements)
  double* weights;   size_t weights_stride;  // N x (n+1), may be null
  uint32_t* indices; size_t indices_stride;  // N x n, or a 1-column scratch
  int keep_indices;
  double* loglik;        // [n_filters]
  int32_t* stopped_at;   // [n_filters]
  // exchange area: per filter 4 * ranks 8-byte words
  unsigned long long* xch;
  unsigned int* barrier; // grid mode only (one counter)
};

// ---------------------------------------------------------------- barriers
struct Sync {
  int mode;
  unsigned int* counter;
  unsigned int nblocks;
  unsigned int target;
  __device__ __forceinline__ void all() {
    if (mode == SYNC_BLOCK) {
      __syncthreads();
    } else if (mode == SYNC_CLUSTER) {
      // hardware cluster barrier; release/acquire at cluster scope also orders
      // the global-memory exchange (state rows, ancestors, partials)
      asm volatile("barrier.cluster.arrive.release.aligned;\n"
                   "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    } else {
      __syncthreads();
      if (threadIdx.x == 0) {
        target += nblocks;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while ((int)(v - target) < 0);
        __threadfence();
      }
      __syncthreads();
    }
  }
};

// NaN-propagating max (a NaN log-weight must stop the filter, see oracle normalise())
__device__ __forceinline__ double nanmax(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  return a > b ? a : b;
}

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ unsigned long long shfl_up_u64(unsigned long long v, int d) {
  return __shfl_up_sync(0xffffffffu, v, d);
}
__device__ __forceinline__ long long shfl_xor_i64(long long v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}

struct CollSmem {
  double wd[32];
  long long wa[32], wb[32];
  unsigned long long wu[32];
  double rd;
  long long ra, rb;
  unsigned long long ru, ru_total;
};

// Collectives over all threads of a filter.  Exchange between CTAs goes through
// `xch` in global memory (L2); rank-level partials are combined by warp 0.
struct Coll {
  CollSmem* sm;
  Sync* sync;
  unsigned long long* xch; // this filter's exchange area
  uint32_t ranks, rank;
  int lane, warp, nwarps;

  __device__ __forceinline__ double allmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, shfl_xor_d(v, o));
    if (lane == 0) sm->wd[warp] = v;
    __syncthreads();
    if (warp == 0) {
      double x = lane < nwarps ? sm->wd[lane] : -pf_inf();
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x = nanmax(x, shfl_xor_d(x, o));
      if (ranks == 1) { if (lane == 0) sm->rd = x; }
      else if (lane == 0) ((double*)xch)[rank] = x;
    }
    sync->all();
    if (ranks > 1) {
      if (warp == 0) {
        double x = -pf_inf();
        for (uint32_t r = lane; r < ranks; r += 32) x = nanmax(x, __ldcg(((const double*)xch) + r));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = nanmax(x, shfl_xor_d(x, o));
        if (lane == 0) sm->rd = x;
      }
      __syncthreads();
    }
    return sm->rd;
  }

  __device__ __forceinline__ void allsum2(long long& a, long long& b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += shfl_xor_i64(a, o); b += shfl_xor_i64(b, o); }
    if (lane == 0) { sm->wa[warp] = a; sm->wb[warp] = b; }
    __syncthreads();
    if (warp == 0) {
      long long x = lane < nwarps ? sm->wa[lane] : 0, y = lane < nwarps ? sm->wb[lane] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { x += shfl_xor_i64(x, o); y += shfl_xor_i64(y, o); }
      if (ranks == 1) { if (lane == 0) { sm->ra = x; sm->rb = y; } }
      else if (lane == 0) { ((long long*)xch)[ranks + rank] = x; ((long long*)xch)[2 * ranks + rank] = y; }
    }
    sync->all();
    if (ranks > 1) {
      if (warp == 0) {
        long long x = 0, y = 0;
        for (uint32_t r = lane; r < ranks; r += 32) {
          x += __ldcg(((const long long*)xch) + ranks + r);
          y += __ldcg(((const long long*)xch) + 2 * ranks + r);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { x += shfl_xor_i64(x, o); y += shfl_xor_i64(y, o); }
        if (lane == 0) { sm->ra = x; sm->rb = y; }
      }
      __syncthreads();
    }
    a = sm->ra; b = sm->rb;
  }

  // exclusive prefix of per-thread totals in (rank, thread) order
  __device__ __forceinline__ unsigned long long exscan(unsigned long long tot) {
    unsigned long long inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long o = shfl_up_u64(inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) sm->wu[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned long long v = lane < nwarps ? sm->wu[lane] : 0ull, w = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = shfl_up_u64(w, d);
        if (lane >= d) w += o;
      }
      sm->wu[lane] = w - v; // exclusive warp prefix
      if (lane == 31) {
        if (ranks == 1) sm->ru = 0ull;
        else xch[3 * ranks + rank] = w;
      }
    }
    sync->all();
    if (ranks > 1) {
      if (warp == 0) {
        unsigned long long x = 0ull;
        for (uint32_t r = lane; r < rank; r += 32) x += __ldcg(xch + 3 * ranks + r);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm->ru = x;
      }
      __syncthreads();
    }
    return sm->ru + sm->wu[warp] + (inc - tot);
  }
};

// --------------------------------------------------- stratified thresholds
// count(p) = #{ j in [0,N) : (r_j + j) * (1.0/N) <= p }, the position the
// reference's two-pointer merge (src/stratified_sample.cpp:17-22) reaches when
// it leaves cumulative weight p.  u_j is non-decreasing in j, so the count is
// found from the guess floor(p*N) with one or (near an integer) a few Philox
// evaluations -- no r[] array is ever materialised.
struct Strata {
  uint32_t seed, filter, t, N;
  double Nd, invN;
  __device__ __forceinline__ double u(uint32_t j) const {
    return (pf_resample_uniform(seed, filter, t, j) + (double)j) * invN;
  }
  __device__ __forceinline__ uint32_t count(double p) const {
    if (p >= 1.0) return N;
    double x = p * Nd;
    uint32_t jj = (uint32_t)x;
    if (jj > N - 1) jj = N - 1;
    const double tol = x * 4.440892098500626e-16; // 2^-51 relative guard
    const bool lo_safe = (x - (double)jj) >= tol;
    const bool hi_safe = ((double)(jj + 1) - x) >= tol;
    if (u(jj) <= p) {
      jj++;
      if (!hi_safe) while (jj < N && u(jj) <= p) jj++;
    } else {
      if (!lo_safe) while (jj > 0 && !(u(jj - 1) <= p)) jj--;
    }
    return jj;
  }
};

// -------------------------------------------------------- the filter kernel
// Model policy:
//   static constexpr int M, KI, KP;   state dim, normals at init, normals/step
//   Model(const Params&, uint32_t filter_index_in_launch)
//   double loglik0() const;           initial log-likelihood (psi: approx_loglik)
//   double loglik_end() const;        added once at the end (bsf: constants)
//   void init(const double* z, double* x) const;
//   bool observed(int t) const;
//   void begin_weight(int t);         load per-step constants
//   double logw(const double* x) const;
//   void begin_step(int t);           load per-step constants for t -> t+1
//   void propagate(const double* xa, const double* z, double* x) const;
template <class Model, int P>
__global__ void __launch_bounds__(1024, 1)
filter_kernel(const FilterParams fp, const typename Model::Params mp) {
  constexpr int M = Model::M;
  constexpr int KI = Model::KI;
  constexpr int KP = Model::KP;
  static_assert(P % 2 == 0, "Box-Muller pairs must be thread-local");
  __shared__ CollSmem csm;

  const uint32_t ranks = fp.ranks;
  const uint32_t fidx = blockIdx.x / ranks;
  const uint32_t rank = blockIdx.x - fidx * ranks;
  const uint32_t filter = fp.first_filter + fidx;
  const uint32_t N = fp.N;
  const int n = fp.n;
  const double Nd = (double)N;
  const double invN = 1.0 / Nd;

  Sync sync;
  sync.mode = fp.sync_mode;
  sync.counter = fp.barrier;
  sync.nblocks = gridDim.x;
  sync.target = 0;
  Coll coll;
  coll.sm = &csm;
  coll.sync = &sync;
  coll.xch = fp.xch + (size_t)fidx * 4 * ranks;
  coll.ranks = ranks;
  coll.rank = rank;
  coll.lane = threadIdx.x & 31;
  coll.warp = threadIdx.x >> 5;
  coll.nwarps = (blockDim.x + 31) >> 5;

  const uint32_t i0 = (rank * blockDim.x + threadIdx.x) * P;
  double* alpha = fp.alpha_tm + (size_t)fidx * fp.alpha_stride;
  double* wout = fp.weights ? fp.weights + (size_t)fidx * fp.weights_stride : nullptr;
  uint32_t* iout = fp.indices + (size_t)fidx * fp.indices_stride;

  Model model(mp, fidx);
  double x[P][M];

  // ---- initial state, column 0 (reference :621-627 / :538-544)
  {
    double z[P * KI];
    const uint32_t pair0 = (i0 * KI) >> 1;
#pragma unroll
    for (int q = 0; q < (P * KI) / 2; ++q)
      pf_normal_pair(fp.seed, filter, 0u, 0u, pair0 + q, &z[2 * q], &z[2 * q + 1]);
#pragma unroll
    for (int p = 0; p < P; ++p) model.init(&z[p * KI], x[p]);
#pragma unroll
    for (int p = 0; p < P; ++p)
      if (i0 + p < N) {
#pragma unroll
        for (int d = 0; d < M; ++d) alpha[(size_t)d * N + i0 + p] = x[p][d];
      }
  }

  double loglik = model.loglik0();
  int stopped = -1;
  double W[P];

  for (int t = 0;; ++t) {
    // ---- weights of column t (reference :633-647, :672-687).  Column n is a
    // pure prediction: unit weights (the reference tests t < n-1 before
    // weighting column t+1).
    const bool obs = (t < n) && model.observed(t);
    if (obs) {
      model.begin_weight(t);
      double lw[P];
      double mx = -pf_inf();
#pragma unroll
      for (int p = 0; p < P; ++p) {
        lw[p] = (i0 + p < N) ? model.logw(x[p]) : -pf_inf();
        mx = nanmax(mx, lw[p]);
      }
      mx = coll.allmax(mx);
      if (!pf_isfinite(mx)) { stopped = t; break; } // uniform over the filter
      long long hi = 0, lo = 0;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        double w = pf_exp(lw[p] - mx);
        W[p] = w;
        pf_fixsum tm = pf_fixsum_term(w);
        hi += tm.hi;
        lo += tm.lo;
      }
      if (wout) {
#pragma unroll
        for (int p = 0; p < P; ++p)
          if (i0 + p < N) wout[(size_t)t * N + i0 + p] = W[p];
      }
      coll.allsum2(hi, lo);
      const double S = pf_fixsum_value(hi, lo);
#pragma unroll
      for (int p = 0; p < P; ++p) W[p] = W[p] / S;
      loglik += mx + pf_log(S / Nd);
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) W[p] = (i0 + p < N) ? invN : 0.0;
      if (wout) {
#pragma unroll
        for (int p = 0; p < P; ++p)
          if (i0 + p < N) wout[(size_t)t * N + i0 + p] = 1.0;
      }
    }
    if (t == n) break;

    // ---- stratified resampling -> ancestors of column t+1 (reference :650-655)
    uint32_t* anc = iout + (fp.keep_indices ? (size_t)t * N : 0);
    {
      unsigned long long q[P], tot = 0ull;
#pragma unroll
      for (int p = 0; p < P; ++p) { q[p] = pf_cumsum_term(W[p]); tot += q[p]; }
      unsigned long long run = coll.exscan(tot);
      Strata st;
      st.seed = fp.seed; st.filter = filter; st.t = (uint32_t)t; st.N = N;
      st.Nd = Nd; st.invN = invN;
      // offspring range of particle k is [count(p_{k-1}), count(p_k)); p_{N-1} := 1
      uint32_t lo = (i0 == 0) ? 0u : ((i0 - 1 >= N - 1) ? N : st.count(pf_cumsum_value(run)));
#pragma unroll
      for (int p = 0; p < P; ++p) {
        run += q[p];
        const uint32_t k = i0 + p;
        uint32_t c = (k >= N - 1) ? N : st.count(pf_cumsum_value(run));
        for (uint32_t j = lo; j < c; ++j) anc[j] = k;
        lo = c > lo ? c : lo;
      }
    }
    sync.all();

    // ---- gather ancestors, propagate to column t+1 (reference :657-670)
    {
      model.begin_step(t);
      double z[P * KP];
      const uint32_t pair0 = (i0 * KP) >> 1;
#pragma unroll
      for (int q = 0; q < (P * KP) / 2; ++q)
        pf_normal_pair(fp.seed, filter, (uint32_t)(t + 1), 0u, pair0 + q, &z[2 * q], &z[2 * q + 1]);
      const double* src = alpha + (size_t)(fp.keep_alpha ? t : (t & 1)) * M * N;
      double* dst = alpha + (size_t)(fp.keep_alpha ? t + 1 : ((t + 1) & 1)) * M * N;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        if (i0 + p < N) {
          const uint32_t a = __ldcg(anc + i0 + p);
          double xa[M];
#pragma unroll
          for (int d = 0; d < M; ++d) xa[d] = __ldcg(src + (size_t)d * N + a);
          model.propagate(xa, &z[p * KP], x[p]);
#pragma unroll
          for (int d = 0; d < M; ++d) dst[(size_t)d * N + i0 + p] = x[p][d];
        }
      }
    }
  }

  if (stopped >= 0) loglik = -pf_inf();
  else loglik += model.loglik_end();
  if (rank == 0 && threadIdx.x == 0) {
    fp.loglik[fidx] = loglik;
    fp.stopped_at[fidx] = stopped;
  }
}

// ------------------------------------------------ layout change for output
// alpha_tm [(t*M+d)*N + i]  ->  cube [d + M*(t + n1*i)]  (Armadillo cube
// m x (n+1) x N, src/R_bootstrap_filter.cpp:162).  A plain tiled transpose of
// the (n1*M) x N matrix; both sides coalesced.
__global__ void transpose_to_cube_kernel(const double* __restrict__ tm, double* __restrict__ cube,
                                         uint32_t rows /* n1*M */, uint32_t N) {
  __shared__ double tile[32][33];
  const uint32_t r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    uint32_t r = r0 + k, c = c0 + threadIdx.x;
    if (r < rows && c < N) tile[k][threadIdx.x] = tm[(size_t)r * N + c];
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    uint32_t c = c0 + k, r = r0 + threadIdx.x;
    if (r < rows && c < N) cube[(size_t)c * rows + r] = tile[threadIdx.x][k];
  }
}

} // namespace pf
