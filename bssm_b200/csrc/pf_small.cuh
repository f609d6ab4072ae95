// pf_small.cuh -- the latency path: one CTA per filter, N <= 1024 particles.
//
// Inside an MCMC loop (mcmc::pm_mcmc / da_mcmc, src/mcmc.cpp:338-491) the particle filter runs
// once per iteration with 10-1000 particles: nothing to amortise, the run time is the length of
// the dependency chain of ONE step times n.  The general kernel (pf_device.cuh) spends that chain
// on (a) the Philox + Box-Muller normals of the next column (a log, a square root, a sine and a
// cosine in sequence), (b) the weights (one or two exponentials), (c) three filter-wide
// collectives, and (d) three trips to L2 (thresholds, ancestors, ancestor states), all on the
// same warps.  Here the CTA is split in two halves:
//   consumers  (threads 0 .. CT-1, two particles each)  weights -> max -> scan -> ancestors -> propagate
//   producers  (threads CT .. 2CT-1)                     thresholds of step t+1 and normals of column t+2
// The producers run one step ahead and hand their results over through shared memory (double
// buffered, one CTA barrier per step), so (a) leaves the critical path.  Producer 0 also keeps the
// log-likelihood recursion (a log and a division per step, thread 0's alone in the general kernel) off the
// consumers' chain: consumer 0 posts (max, sum, constant), producer 0 accumulates.  (Measured and dropped:
// issuing the model's per-step loads one step early, from either half -- the step is a chain of dependent
// FP64 operations, not of cache misses.)  Thresholds, ancestors
// and the current state column live in shared memory, so (d) does too; the collectives are
// warp shuffles plus one named barrier over the consumer half each.  The arithmetic per particle
// is the model policy's (the same functions the general kernel calls), the random numbers are
// the same Philox stream, the weight sums the same fixed-point integers: the results are
// bit-identical to the general kernel and to the oracle.
#pragma once
#include "pf_device.cuh"

namespace pf {

enum { SYNC_SMALL = 3 }; // Geometry::sync_mode value that selects this kernel (launch.cuh)

__device__ __forceinline__ void bar_named(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// the hand-over barrier between the two halves: they arrive from different places in the code, so it is a
// named barrier with an explicit thread count (bar.sync 2, blockDim.x), not __syncthreads()
__device__ __forceinline__ void bar_all() { bar_named(2, (int)blockDim.x); }
__device__ __forceinline__ long long shfl_up_i64(long long v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

// Strata (pf_device.cuh) over thresholds in shared memory
struct StrataS {
  const double* u_;
  uint32_t N;
  double Nd;
  __device__ __forceinline__ double u(uint32_t j) const { return u_[j]; }
  __device__ __forceinline__ uint32_t guess(double p) const {
    const uint32_t jj = (uint32_t)(p * Nd);
    return jj > N - 1 ? N - 1 : jj;
  }
  __device__ __forceinline__ uint32_t resolve(double p, uint32_t jj, double uj) const {
    const double x = p * Nd;
    const bool le = uj <= p;
    uint32_t c = jj + (le ? 1u : 0u);
    const double dc = pf_u32_to_f64(c);
    const double d = le ? dc - x : x - dc;
    if (d < x * 4.440892098500626e-16) { // p * N within 2^-51 of an integer: walk (rare)
      if (le) { while (c < N && u_[c] <= p) c++; }
      else { while (c > 0 && !(u_[c - 1] <= p)) c--; }
    }
    return p >= 1.0 ? N : c;
  }
};

// bytes of dynamic shared memory for N particles, P per consumer thread
template <class Model, int P>
__host__ __device__ constexpr size_t small_noise_bytes() {
  return (sizeof(typename Model::template Noise<P>) + 7) / 8 * 8;
}
__host__ __device__ inline size_t small_consumers(uint32_t N, int P) { // consumer threads, a multiple of 32
  return (size_t)((((N + P - 1) / P) + 31) / 32) * 32;
}
template <class Model, int P>
__host__ __device__ inline size_t small_smem_bytes(uint32_t N) {
  const size_t CT = small_consumers(N, P), Np = P * CT;
  return 2 * Model::M * Np * 8 + 2 * Np * 8 + 2 * CT * small_noise_bytes<Model, P>() + Np * 4;
}

// P particles per consumer thread.  Measured: P = 2 is the fastest -- fewer warps with four or eight particles
// each (one consumer warp for N = 100, collectives by shuffles only) are SLOWER in proportion to P (2.45 ->
// 3.3 us per step at N = 100, 4.96 -> 7.1 at N = 1000): the step is not a chain of latencies that more
// independent work per thread could fill, it is bounded by fetching the straight-line code of one step
// (~2 500 instructions per warp and step at P = 2, beyond the L0 instruction cache), and unrolling over
// P multiplies it.  MAXT = 512 (N <= 512) lets the compiler use up to 128 registers.
template <class Model, int P, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
small_filter_kernel(const FilterParams fp, const typename Model::Params mp) {
  constexpr int M = Model::M;
  typedef typename Model::template Noise<P> Noise;
  constexpr size_t NZB = small_noise_bytes<Model, P>();
  extern __shared__ double smem_d[];
  __shared__ double s_max[2][32];
  __shared__ long long s_h[2][32], s_l[2][32];
  __shared__ int s_stop;
  __shared__ double s_ll[2][4]; // per step: obs flag, max, sum of weights, step constant (consumer 0 -> producer 0)

  const uint32_t N = fp.N;
  const int n = fp.n;
  const double Nd = (double)N, invN = 1.0 / Nd;
  const uint32_t CT = blockDim.x >> 1, Np = P * CT;
  const bool producer = threadIdx.x >= CT;
  const uint32_t j = producer ? threadIdx.x - CT : threadIdx.x;
  const uint32_t i0 = P * j;
  const uint32_t lane = threadIdx.x & 31, warp = j >> 5, nwarps = CT >> 5;
  const uint32_t fidx = blockIdx.x;
  const uint32_t filter = fp.first_filter + fidx;

  double* xs = smem_d;                          // [2][M][Np] state columns
  double* ubs = xs + (size_t)2 * M * Np;        // [2][Np] thresholds
  unsigned char* nzs = reinterpret_cast<unsigned char*>(ubs + 2 * Np); // [2][CT] Noise
  uint32_t* ancs = reinterpret_cast<uint32_t*>(nzs + 2 * CT * NZB);    // [Np] ancestors of the step

  double* alpha = fp.alpha_tm + (size_t)fidx * fp.alpha_stride;
  double* wout = fp.weights ? fp.weights + (size_t)fidx * fp.weights_stride : nullptr;
  uint32_t* iout = fp.indices + (size_t)fidx * fp.indices_stride;

  Model model(mp, fidx);
  if (threadIdx.x == 0) s_stop = 0;

  auto thresholds = [&](int t, double* ub) {
#pragma unroll
    for (int h = 0; h < P / 2; ++h) {
      const uint32_t jj = i0 + 2 * h;
      pf_u32x4 rx = pf_philox4x32_10(jj >> 1, 0u, (uint32_t)t, filter, fp.seed, (uint32_t)PF_RNG_RESAMPLE);
      ub[jj] = pf_u53_plus(rx.v[0], rx.v[1], pf_u32_to_f64(jj)) * invN;
      ub[jj + 1] = pf_u53_plus(rx.v[2], rx.v[3], pf_u32_to_f64(jj + 1)) * invN;
    }
  };

  auto produce = [&](int t) { // thresholds of step t and normals of column t + 1 into buffer t & 1
    const int b = t & 1;
    thresholds(t, ubs + (size_t)b * Np);
    StepCtx ctx;
    ctx.seed = fp.seed; ctx.filter = filter; ctx.t = (uint32_t)(t + 1); ctx.i0 = i0;
    Noise nz;
    model.template draw<P>(ctx, nz);
    *reinterpret_cast<Noise*>(nzs + ((size_t)b * CT + j) * NZB) = nz;
  };
  // producer 0: log-likelihood of step s from what consumer 0 posted (same operations, same order as pf_device.cuh)
  double loglik = model.loglik0();
  auto account = [&](int s) {
    const double* q = s_ll[s & 1];
    if (q[0] != 0.0) {
      if (s == 0) loglik = ((q[1] + loglik) + pf_log_normal(q[2] / Nd)) + q[3];
      else loglik += (q[1] + pf_log_normal(q[2] / Nd)) + q[3];
    }
  };

  // ---- prologue: column 0 of the state (consumers), buffer 0 (producers)
  double x[P][M];
  double carry[P];
#pragma unroll
  for (int p = 0; p < P; ++p) carry[p] = 0.0;
  if (producer) {
    if (n > 0) produce(0);
  } else {
    StepCtx c0;
    c0.seed = fp.seed; c0.filter = filter; c0.t = 0u; c0.i0 = i0;
    model.template init_all<P>(c0, x);
#pragma unroll
    for (int d = 0; d < M; ++d)
#pragma unroll
      for (int p = 0; p < P; ++p) {
        xs[(size_t)d * Np + i0 + p] = x[p][d];
        if (i0 + p < N) alpha[(size_t)d * N + i0 + p] = x[p][d];
      }
  }
  bar_all();

#if PF_PHASE_TIMING
  // consumer thread 0: [0] weights + max, [1] exp + fixed-point scan, [2] ancestors, [3] gather + propagate,
  // [4] waiting at the hand-over; [5] is the producers' busy time (thread CT)
  unsigned long long ph[6] = {0, 0, 0, 0, 0, 0};
  const bool timing = fp.phase_cycles != nullptr && (threadIdx.x == 0 || threadIdx.x == CT);
  long long tc = timing ? clock64() : 0;
#define PFS_PHASE(i) if (timing) { long long now_ = clock64(); ph[i] += (unsigned long long)(now_ - tc); tc = now_; }
#else
#define PFS_PHASE(i)
#endif

  // ---- time loop: ONE hand-over barrier per step, reached by both halves at the same place
  int stopped = -1;
  bool halted = false;
  for (int t = 0; t < n; ++t) {
    const int b = t & 1;
    if (producer) {
      if (t + 1 < n) produce(t + 1);
      if (j == 0 && t >= 1) account(t - 1); // posted before the hand-over that ended step t - 1
      PFS_PHASE(5)
    } else {
      const bool obs = model.observed(t);
      double w[P], lw[P], mx = 0.0, sc_t = 0.0;
      bool stop_now = false;
      if (obs) {
        model.begin_weight(t);
        mx = -pf_inf();
        bool bad = false;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          lw[p] = (i0 + p < N) ? model.finish_logw(model.logw(x[p]), carry[p]) : -pf_inf();
          bad = bad || (lw[p] != lw[p]);
          mx = lw[p] > mx ? lw[p] : mx;
        }
        if (bad) mx = pf_inf(); // a NaN log-weight must stop the filter
        sc_t = model.step_const();
        mx = warp_max_d(mx);
        if (nwarps == 1) {
          // one consumer warp: the shuffles were the whole collective
        } else if (lane == 0) s_max[b][warp] = mx;
        if (nwarps > 1) bar_named(1, CT);
        if (nwarps == 1) {
        } else if (nwarps <= 4) { // a few warps: every thread takes the maximum of the slots itself
          double m2 = s_max[b][0];
          for (uint32_t wv = 1; wv < nwarps; ++wv) { const double v = s_max[b][wv]; m2 = v > m2 ? v : m2; }
          mx = m2;
        } else {
          mx = warp_max_d(lane < nwarps ? s_max[b][lane] : -pf_inf());
        }
        stop_now = !pf_isfinite(mx); // uniform over the filter
#pragma unroll
        for (int p = 0; p < P; ++p) w[p] = pf_exp_nonpos(lw[p] - mx);
      } else {
#pragma unroll
        for (int p = 0; p < P; ++p) w[p] = (i0 + p < N) ? 1.0 : 0.0;
      }
      if (stop_now) {
        stopped = t;
        if (threadIdx.x == 0) s_stop = t + 1; // tagged with the step: a late reader of the previous hand-over cannot mistake it
      } else {
        if (wout) {
#pragma unroll
          for (int p = 0; p < P; ++p)
            if (i0 + p < N) wout[(size_t)t * N + i0 + p] = w[p];
        }
        PFS_PHASE(0)

        // ---- one scan of the fixed-point weights: sum and cumulative sums
        long long qh[P], ql[P], th = 0, tl = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          pf_fixsum tm = pf_fixsum_term(w[p]);
          qh[p] = tm.hi; ql[p] = tm.lo;
          th += tm.hi; tl += tm.lo;
        }
        long long ih = th, il = tl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const long long vh = shfl_up_i64(ih, o), vl = shfl_up_i64(il, o);
          if ((int)lane >= o) { ih += vh; il += vl; }
        }
        long long rh = ih - th, rl = il - tl, Th = 0, Tl = 0;
        if (nwarps == 1) { // the warp's inclusive scan already holds the totals in lane 31
          Th = __shfl_sync(0xffffffffu, ih, 31); Tl = __shfl_sync(0xffffffffu, il, 31);
        } else {
          if (lane == 31) { s_h[b][warp] = ih; s_l[b][warp] = il; }
          bar_named(1, CT);
          for (uint32_t wv = 0; wv < nwarps; ++wv) {
            const long long h = s_h[b][wv], l = s_l[b][wv];
            Th += h; Tl += l;
            if (wv < warp) { rh += h; rl += l; }
          }
        }
        const double S = pf_fixsum_value(Th, Tl);
        const double invS = 1.0 / S;
        if (threadIdx.x == 0) { // the log-likelihood recursion is producer 0's
          double* q = s_ll[b];
          q[0] = obs ? 1.0 : 0.0; q[1] = mx; q[2] = S; q[3] = sc_t;
        }
        PFS_PHASE(1)

        // ---- stratified resampling: offspring range of particle k is [count(p_{k-1}), count(p_k))
        StrataS st;
        st.u_ = ubs + (size_t)b * Np; st.N = N; st.Nd = Nd;
        {
          const double pk0 = (i0 == 0) ? 0.0 : ((i0 - 1 >= N - 1) ? 1.0 : pf_fixsum_value(rh, rl) * invS);
          uint32_t lo = (i0 == 0) ? 0u : st.resolve(pk0, st.guess(pk0), st.u(st.guess(pk0)));
          constexpr int CH = P > 4 ? 4 : P; // chunks keep the live cumulative probabilities / thresholds few
#pragma unroll
          for (int c0 = 0; c0 < P; c0 += CH) {
            double pk[CH];
            uint32_t jj[CH];
            double uj[CH];
#pragma unroll
            for (int q = 0; q < CH; ++q) {
              const int p = c0 + q;
              rh += qh[p]; rl += ql[p];
              pk[q] = (i0 + p >= N - 1) ? 1.0 : pf_fixsum_value(rh, rl) * invS;
              jj[q] = st.guess(pk[q]);
              uj[q] = st.u(jj[q]);
            }
#pragma unroll
            for (int q = 0; q < CH; ++q) {
              const uint32_t c = st.resolve(pk[q], jj[q], uj[q]);
              const uint32_t val = i0 + c0 + q;
              const uint32_t cnt = c > lo ? c - lo : 0u;
              if (cnt) {
                uint32_t* qa = ancs + lo;
                qa[0] = val;
                if (cnt > 1) {
                  qa[1] = val;
                  if (cnt > 2) {
                    qa[2] = val;
                    if (cnt > 3) qa[3] = val;
                  }
                }
              }
              const uint32_t ce = lo + PF_SCATTER_CAP;
              const unsigned heavy = __ballot_sync(0xffffffffu, cnt > PF_SCATTER_CAP);
              if (heavy) scatter_heavy(ancs, heavy, ce, c, val);
              lo = c > lo ? c : lo;
            }
          }
        }
        if (nwarps == 1) __syncwarp(); else bar_named(1, CT);
        PFS_PHASE(2)

        // ---- gather the ancestors from shared memory, propagate to column t + 1
        {
          StepCtx ctx;
          ctx.seed = fp.seed; ctx.filter = filter; ctx.t = (uint32_t)(t + 1); ctx.i0 = i0;
          model.begin_step(t);
          const double* src = xs + (size_t)b * M * Np;
          double* dsts = xs + (size_t)(b ^ 1) * M * Np;
          double* dst = alpha + (size_t)(fp.keep_alpha ? t + 1 : ((t + 1) & 1)) * M * N;
          uint32_t a[P];
#pragma unroll
          for (int p = 0; p < P; ++p) a[p] = (i0 + p < N) ? ancs[i0 + p] : 0u;
          if (fp.keep_indices) {
#pragma unroll
            for (int p = 0; p < P; ++p)
              if (i0 + p < N) iout[(size_t)t * N + i0 + p] = a[p];
          }
          double xa[P][M];
#pragma unroll
          for (int p = 0; p < P; ++p)
#pragma unroll
            for (int d = 0; d < M; ++d) xa[p][d] = src[(size_t)d * Np + a[p]];
          const Noise nz = *reinterpret_cast<const Noise*>(nzs + ((size_t)b * CT + j) * NZB);
          model.template propagate_all<P>(ctx, nz, xa, x, carry);
#pragma unroll
          for (int d = 0; d < M; ++d)
#pragma unroll
            for (int p = 0; p < P; ++p) {
              dsts[(size_t)d * Np + i0 + p] = x[p][d];
              if (i0 + p < N) dst[(size_t)d * N + i0 + p] = x[p][d];
            }
        }
        PFS_PHASE(3)
      }
    }
    // hand-over: column t + 1 of the state is complete, the producers have filled buffer (t + 1) & 1,
    // consumer 0 has posted step t
    bar_all();
    if (!producer) { PFS_PHASE(4) }
    if (s_stop == t + 1) { halted = true; break; }
  }
#undef PFS_PHASE
#if PF_PHASE_TIMING
  if (timing) {
    if (threadIdx.x == 0) for (int i = 0; i < 5; ++i) fp.phase_cycles[(size_t)blockIdx.x * 6 + i] = ph[i];
    else fp.phase_cycles[(size_t)blockIdx.x * 6 + 5] = ph[5];
  }
#endif

  if (halted) { // every thread saw it after the hand-over of the step that stopped
    if (threadIdx.x == 0) { fp.loglik[fidx] = -pf_inf(); fp.stopped_at[fidx] = stopped; }
    return;
  }
  if (producer) {
    if (j == 0) {
      if (n >= 1) account(n - 1);
      fp.loglik[fidx] = loglik + model.loglik_end();
      fp.stopped_at[fidx] = -1;
    }
  } else if (wout) { // column n is a pure prediction: unit weights
#pragma unroll
    for (int p = 0; p < P; ++p)
      if (i0 + p < N) wout[(size_t)n * N + i0 + p] = 1.0;
  }
}

} // namespace pf
