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
#if !defined(__CUDACC_RTC__)
#include <cuda_runtime.h>
#include <stdint.h>
#endif
#include "bssm_b200/pf_math.h"

// Measured alternatives that were dropped (tools/build_variants.py, C2 batch of 18 filters):
// recomputing the fixed-point terms after the scan instead of keeping them in registers (no
// change), resolving all P + 1 counts at once instead of in chunks of 4 (-5%), drawing the
// normals with the threshold loads or inside the scan collective instead of inside the scatter
// barrier (-4% / -1%), letting the compiler unroll the scatter loop (-3%).

#ifndef PF_PHASE_TIMING
#define PF_PHASE_TIMING 0 // 1: per-phase cycle counters (cost 14 registers per thread; diagnostics builds only)
#endif
#ifndef PF_POLL_NS
#define PF_POLL_NS 32
#endif
#ifndef PF_SCATTER_CAP
#define PF_SCATTER_CAP 4 // offspring written by the owner lane itself; longer runs are written by the whole warp
#endif

namespace pf {

// the barrier counter of every filter sits in its own 128-byte line: the counters of a batch used to
// share one line, and every arrival and every poll of every CTA of the launch went to one L2 slice
constexpr int COUNTER_STRIDE = 32;

enum SyncMode { SYNC_BLOCK = 0, SYNC_CLUSTER = 1, SYNC_GROUP = 2 };

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
  // exchange area for the partials: per filter 4 * ranks 8-byte words
  unsigned long long* xch;
  // SYNC_GROUP: one barrier counter per filter, zeroed before the launch
  unsigned int* counters;
  // stratum thresholds u_j = (r_j + j) * (1/N) of the current step, N per filter
  double* ubuf;
  // diagnostics: when non-null, thread 0 of every CTA accumulates the SM cycles
  // spent up to each phase boundary: 6 counters per CTA
  //   [0] log-weights  [1] max collective  [2] exp + thresholds + terms
  //   [3] scan collective  [4] counts + normals + scatter + barrier  [5] gather + propagate
  unsigned long long* phase_cycles;
  // streaming fetch: when non-null (host-mapped memory, one int per filter), rank 0 publishes
  // the number of complete columns every `progress_every` steps with a system-scope fence
  int* progress;
  int progress_every;
};

// ---------------------------------------------------------- synchronisation
// A filter that spans several CTAs synchronises three times per time step
// (max, scan, scatter).  Partials are exchanged through L2 (`xch`), separated by
// a barrier over the CTAs of the filter:
//
//  SYNC_GROUP (default): software barrier -- one monotonically increasing
//    counter per filter, one atomicAdd and one polling thread per CTA, fences on
//    both sides.  Works for any number of CTAs per filter and any number of
//    filters per launch; all CTAs of a launch must be co-resident (cooperative
//    launch; the host splits a batch into waves).  Measured 1.2 us (8 CTAs) to
//    2.6 us (148 CTAs) per collective including the read of the partials
//    (tools/bench_sync.cu); publishing {value, epoch} descriptors polled by a
//    whole warp was tried and is slower (1.6 - 4.2 us: L2 contention).
//  SYNC_CLUSTER: hardware barrier.cluster (release/acquire); needs a cluster
//    launch, <= 16 CTAs per filter and a GPC with that many free SMs, which
//    limits how many filters run concurrently.  Kept for comparison.
//
// The release / acquire of the barrier also orders the bulk exchange through
// global memory (state rows, thresholds, ancestors); consumers read with ld.cg.
// Tried and rejected (measured on the C2 workloads, tools/build_variants.py):
//  * named barriers (bar.arrive for the producer warps, bar.sync for warp 0) so that
//    producers run on into the overlap window: no gain;
//  * warp 0 polling AND combining the rank partials before a single CTA barrier
//    (one __syncthreads less per collective): 13% slower on both workloads.
struct Sync {
  int mode;
  unsigned int* counter; // SYNC_GROUP: this filter's counter (zeroed before the launch)
  unsigned int nranks;
  unsigned int target;
  // Split-phase barrier.  arrive(): every write of this CTA so far is released;
  // wait(): the writes of all CTAs of the filter released by their arrive() are
  // visible.  Work placed between the two overlaps the barrier latency; it must
  // not depend on what the barrier orders.
  // `joined`: the caller has just passed a __syncthreads() that every thread
  // reaches after its last write the barrier has to publish, and thread 0 wrote
  // (or is ordered after) the exchange partials.
  __device__ __forceinline__ void arrive(bool joined = false) {
    if (mode == SYNC_BLOCK) {
    } else if (mode == SYNC_CLUSTER) {
      asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    } else {
      if (!joined) __syncthreads();
      if (threadIdx.x == 0) {
        target += nranks;
        // release only (MEMBAR.ALL.GPU): __threadfence() is fence.sc + CCTL.IVALL, and the L1
        // invalidation evicts the spill lines of every CTA resident on the SM (round 2)
        asm volatile("fence.release.gpu;" ::: "memory");
        atomicAdd(counter, 1u);
      }
    }
  }
  // SYNC_GROUP, warp 0 only: returns when every CTA of the filter has arrived
  __device__ __forceinline__ void poll_warp0() {
    if (threadIdx.x == 0) {
      // relaxed polling and NO acquire fence: a gpu-scope acquire is a CCTL.IVALL, which evicts
      // the spill lines of all resident CTAs; it is not needed because everything another CTA
      // wrote (partials, state row, thresholds, ancestors) is read with L2 loads (ld.cg) issued
      // behind the CTA barrier that follows, never through a possibly stale L1 line
      unsigned int v;
      for (;;) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
#ifdef PF_ABL_NOWAIT
        break;
#endif
        if ((int)(v - target) >= 0) break;
        // the polling warps of all CTAs of an SM share one scheduler (warp 0 of every CTA): back off
        // so that the spin does not eat its issue slots
        if (PF_POLL_NS > 0) __nanosleep(PF_POLL_NS);
      }
    }
    __syncwarp();
  }
  __device__ __forceinline__ void wait() {
    if (mode == SYNC_BLOCK) {
      __syncthreads();
    } else if (mode == SYNC_CLUSTER) {
      asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    } else {
      if (threadIdx.x < 32) poll_warp0();
      __syncthreads();
    }
  }
  __device__ __forceinline__ void all() { arrive(); wait(); }
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
__device__ __forceinline__ long long shfl_xor_i64(long long v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}

// warp maximum of a double (no NaN) with two 32-bit REDUX instead of five shuffle / compare /
// select rounds: the bit pattern is mapped to an unsigned key with the same order
__device__ __forceinline__ double warp_max_d(double v) {
  unsigned long long b = pf_d2u(v);
  b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
  const uint32_t hi = (uint32_t)(b >> 32), lo = (uint32_t)b;
  const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  b = ((unsigned long long)mh << 32) | ml;
  b = (b >> 63) ? (b & 0x7fffffffffffffffull) : ~b;
  return pf_u2d(b);
}

struct CollSmem {
  double wd[32];
  long long wa[32], wb[32];
  double rd;
  long long ra, rb, ta, tb;
};

// Collectives over all threads of a filter: warp shuffle -> shared memory ->
// (ranks > 1) exchange between CTAs; rank-level partials are combined by warp 0.
struct Coll {
  CollSmem* sm;
  Sync* sync;
  unsigned long long* xch; // this filter's exchange area, 3 * ranks words
  uint32_t ranks, rank;
  int lane, warp, nwarps;

  __device__ __forceinline__ void barrier() { sync->all(); }
  __device__ __forceinline__ void barrier_arrive() { sync->arrive(); }
  __device__ __forceinline__ void barrier_wait() { sync->wait(); }

  // max over the filter; the caller maps NaN to +inf beforehand, so a plain max
  // suffices.  Split in two: independent work between _begin and _end hides the
  // latency of the exchange between CTAs.
  __device__ __forceinline__ void allmax_begin(double v) {
    v = warp_max_d(v);
    if (lane == 0) sm->wd[warp] = v;
    __syncthreads();
    if (warp == 0) {
      const double x = warp_max_d(lane < nwarps ? sm->wd[lane] : -pf_inf());
      if (ranks == 1) { if (lane == 0) sm->rd = x; }
      else if (lane == 0) ((double*)xch)[rank] = x;
    }
    // thread 0 wrote the partial itself; nothing else has to be published here
    sync->arrive(true);
  }
  __device__ __forceinline__ double allmax_end() {
    sync->wait();
    if (ranks > 1) {
      if (warp == 0) {
        double x = -pf_inf();
        for (uint32_t r = lane; r < ranks; r += 32) { const double y = __ldcg(((const double*)xch) + r); x = y > x ? y : x; }
        x = warp_max_d(x);
        if (lane == 0) sm->rd = x;
      }
      __syncthreads();
    }
    return sm->rd;
  }
  __device__ __forceinline__ double allmax(double v) { allmax_begin(v); return allmax_end(); }

  // Exclusive prefix, in (rank, thread) order, of the per-thread totals of the
  // two fixed-point words, plus the grand totals: one collective serves both the
  // sum of the weights and their cumulative sum.  _begin returns the thread's
  // inclusive prefix within its warp (needed by _end).
  __device__ __forceinline__ void exscan2_begin(long long th, long long tl, long long& ih, long long& il) {
    ih = th; il = tl;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      long long oh = __shfl_up_sync(0xffffffffu, ih, d);
      long long ol = __shfl_up_sync(0xffffffffu, il, d);
      if (lane >= d) { ih += oh; il += ol; }
    }
    if (lane == 31) { sm->wa[warp] = ih; sm->wb[warp] = il; }
    __syncthreads();
    if (warp == 0) {
      long long vh = lane < nwarps ? sm->wa[lane] : 0, vl = lane < nwarps ? sm->wb[lane] : 0;
      long long wh = vh, wl = vl;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        long long oh = __shfl_up_sync(0xffffffffu, wh, d);
        long long ol = __shfl_up_sync(0xffffffffu, wl, d);
        if (lane >= d) { wh += oh; wl += ol; }
      }
      sm->wa[lane] = wh - vh; // exclusive warp prefixes
      sm->wb[lane] = wl - vl;
      if (lane == 31) {
        if (ranks == 1) { sm->ra = 0; sm->rb = 0; sm->ta = wh; sm->tb = wl; }
        else { ((long long*)xch)[ranks + rank] = wh; ((long long*)xch)[2 * ranks + rank] = wl; }
      }
      __syncwarp(); // lane 31's partials before thread 0 releases them
    }
    // every thread joined warp 0 above after its stores of this step (weights
    // row, thresholds): thread 0's release publishes them too
    sync->arrive(true);
  }
  __device__ __forceinline__ void exscan2_end(long long th, long long tl, long long ih, long long il,
                                              long long& eh, long long& el, long long& Th, long long& Tl) {
    sync->wait();
    if (ranks > 1) {
      // rank partials are combined by warp 0 only: letting every warp read them
      // (saving the CTA barrier) was measured 1-8% slower
      if (warp == 0) {
        long long xh = 0, xl = 0, sh = 0, sl = 0;
        for (uint32_t r = lane; r < ranks; r += 32) {
          const long long h = __ldcg(((const long long*)xch) + ranks + r);
          const long long l = __ldcg(((const long long*)xch) + 2 * ranks + r);
          sh += h; sl += l;
          if (r < rank) { xh += h; xl += l; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          xh += shfl_xor_i64(xh, o); xl += shfl_xor_i64(xl, o);
          sh += shfl_xor_i64(sh, o); sl += shfl_xor_i64(sl, o);
        }
        if (lane == 0) { sm->ra = xh; sm->rb = xl; sm->ta = sh; sm->tb = sl; }
      }
      __syncthreads();
    }
    eh = sm->ra + sm->wa[warp] + (ih - th);
    el = sm->rb + sm->wb[warp] + (il - tl);
    Th = sm->ta; Tl = sm->tb;
  }
  __device__ __forceinline__ void exscan2(long long th, long long tl, long long& eh, long long& el,
                                          long long& Th, long long& Tl) {
    long long ih, il;
    exscan2_begin(th, tl, ih, il);
    exscan2_end(th, tl, ih, il, eh, el, Th, Tl);
  }
};

// --------------------------------------------------- stratified thresholds
// count(p) = #{ j in [0,N) : u_j <= p } with u_j = (r_j + j) * (1.0/N): the
// position the reference's two-pointer merge (src/stratified_sample.cpp:17-22)
// reaches when it leaves cumulative weight p.  u_j is non-decreasing in j, so
// the count follows from the guess floor(p*N) with one threshold look-up (a few
// only when p*N is within 2^-51 of an integer).  The thresholds of the step are
// generated once by the threads that own the strata (one Philox call per two
// strata) and parked in `ubuf` (L2-resident, 8 B per particle).
// p * N within 2^-51 (relative) of an integer: walk the thresholds (out of line, rare)
static __device__ __noinline__ uint32_t strata_walk(const double* ubuf, uint32_t N, double p, uint32_t c, bool up) {
  if (up) { while (c < N && __ldcg(ubuf + c) <= p) c++; }
  else { while (c > 0 && !(__ldcg(ubuf + c - 1) <= p)) c--; }
  return c;
}
struct Strata {
  const double* ubuf;
  uint32_t N;
  double Nd;
  __device__ __forceinline__ double u(uint32_t j) const { return __ldcg(ubuf + j); }
  __device__ __forceinline__ uint32_t guess(double p) const {
    const uint32_t jj = (uint32_t)(p * Nd);
    return jj > N - 1 ? N - 1 : jj;
  }
  // resolve the count given the pre-loaded threshold uj = u(guess(p))
  __device__ __forceinline__ uint32_t resolve(double p, uint32_t jj, double uj) const {
    const double x = p * Nd;
    const bool le = uj <= p;
    uint32_t c = jj + (le ? 1u : 0u);
    const double dc = pf_u32_to_f64(c);
    const double d = le ? dc - x : x - dc;  // distance of p * N to the integer that would change the count
    if (d < x * 4.440892098500626e-16) c = strata_walk(ubuf, N, p, c, le);
    return p >= 1.0 ? N : c;
  }
};

// offspring runs longer than PF_SCATTER_CAP: every lane of the warp calls this with the same
// `heavy` mask; the lanes in it broadcast their run [s, e) and value, the whole warp writes it
static __device__ __noinline__ void scatter_heavy(uint32_t* anc, unsigned heavy, uint32_t s, uint32_t e, uint32_t val) {
  const uint32_t lane = threadIdx.x & 31u;
  while (heavy) {
    const int src = __ffs(heavy) - 1;
    heavy &= heavy - 1;
    const uint32_t s0 = __shfl_sync(0xffffffffu, s, src), e0 = __shfl_sync(0xffffffffu, e, src);
    const uint32_t v0 = __shfl_sync(0xffffffffu, val, src);
#pragma unroll 1
    for (uint32_t j = s0 + lane; j < e0; j += 32) anc[j] = v0;
  }
}

// ------------------------------------------------------- per-step context
struct StepCtx {
  uint32_t seed, filter, t, i0; // t = time column being produced (0 = initial state)
};

// Noise handling shared by the models whose step consumes K standard normals per
// particle (ssm_ung, ssm_nlg): the thread's P*K normals are P*K/2 Box-Muller
// pairs of the flat index (i*K + d), generated BEFORE the ancestors are known.
template <class Derived, int M, int KI, int KP>
struct GaussStep {
  template <int P> struct Noise { double z[P * KP]; };
  template <int P>
  __device__ __forceinline__ void init_all(const StepCtx& c, double (&x)[P][M]) const {
    double z[P * KI];
    const uint32_t pair0 = (c.i0 * KI) >> 1;
#pragma unroll
    for (int q = 0; q < (P * KI) / 2; ++q)
      pf_normal_pair(c.seed, c.filter, 0u, 0u, pair0 + q, &z[2 * q], &z[2 * q + 1]);
#pragma unroll
    for (int p = 0; p < P; ++p) static_cast<const Derived*>(this)->init(&z[p * KI], x[p]);
  }
  template <int P>
  __device__ __forceinline__ void draw(const StepCtx& c, Noise<P>& nz) const {
    const uint32_t pair0 = (c.i0 * KP) >> 1;
#ifdef PF_ABL_NONORMAL
#pragma unroll
    for (int q = 0; q < P * KP; ++q) nz.z[q] = 0.3 + 1e-9 * (double)(pair0 + q + c.t);
#else
#pragma unroll
    for (int q = 0; q < (P * KP) / 2; ++q)
      pf_normal_pair(c.seed, c.filter, c.t, 0u, pair0 + q, &nz.z[2 * q], &nz.z[2 * q + 1]);
#endif
  }
  // `carry` is one double per particle that survives from the propagation into the next
  // weighting: the psi-APF of ssm_nlg needs the ancestor's state in its weights (ratio of
  // true and approximate transition densities, src/model_ssm_nlg.cpp:915-931).
  template <int P>
  __device__ __forceinline__ void propagate_all(const StepCtx&, const Noise<P>& nz, const double (&xa)[P][M],
                                                double (&x)[P][M], double (&carry)[P]) const {
#pragma unroll
    for (int p = 0; p < P; ++p) {
      static_cast<const Derived*>(this)->propagate(xa[p], &nz.z[p * KP], x[p]);
      carry[p] = static_cast<const Derived*>(this)->trans_term(xa[p], x[p]);
    }
  }
  __device__ __forceinline__ double trans_term(const double*, const double*) const { return 0.0; }
  __device__ __forceinline__ double step_const() const { return 0.0; }
  __device__ __forceinline__ double finish_logw(double w, double) const { return w; }
};

// ------------------------------------------------ per-thread row accesses
// A thread's P particles are contiguous in every per-time row (weights column,
// state row, threshold row, ancestor column): when the row base is 16-byte
// aligned (`vec`: N even, resp. N % 4 == 0 for the 32-bit ancestors) and the
// thread is not the ragged last one, P doubles go out as P/2 128-bit stores.
// STREAM: an output column that the kernel never reads again leaves with an evict-first hint
// (st.global.cs), so that it does not push the exchange buffers (thresholds, state row, ancestors of
// the current step) out of L2 before they are consumed.
template <int P, bool STREAM = false>
__device__ __forceinline__ void store_row(double* row, uint32_t i0, uint32_t N, bool vec, const double (&v)[P]) {
  if (vec && i0 + P <= N) {
    double2* q = reinterpret_cast<double2*>(row + i0);
#pragma unroll
    for (int h = 0; h < P / 2; ++h) {
#ifndef PF_NO_STREAM_WEIGHTS // measured (ncu, bench launch): DRAM traffic 34.3 -> 30.3 GB, same time
      if (STREAM) __stcs(q + h, make_double2(v[2 * h], v[2 * h + 1]));
      else
#endif
        q[h] = make_double2(v[2 * h], v[2 * h + 1]);
    }
  } else {
#pragma unroll
    for (int p = 0; p < P; ++p)
      if (i0 + p < N) row[i0 + p] = v[p];
  }
}
template <int P>
__device__ __forceinline__ void load_anc(const uint32_t* anc, uint32_t i0, uint32_t N, bool vec4, uint32_t (&a)[P]) {
  if (vec4 && i0 + P <= N) {
    if (P == 2) {
      const uint2 v = __ldcg(reinterpret_cast<const uint2*>(anc + i0));
      a[0] = v.x; a[1] = v.y;
    } else {
#pragma unroll
      for (int h = 0; h < P / 4; ++h) {
        const uint4 v = __ldcg(reinterpret_cast<const uint4*>(anc + i0) + h);
        a[4 * h] = v.x; a[4 * h + 1] = v.y; a[4 * h + 2] = v.z; a[4 * h + 3] = v.w;
      }
    }
  } else {
#pragma unroll
    for (int p = 0; p < P; ++p) a[p] = (i0 + p < N) ? __ldcg(anc + i0 + p) : 0u;
  }
}

// -------------------------------------------------------- the filter kernel
// Model policy:
//   static constexpr int M;           state dimension
//   Model(const Params&, uint32_t filter_index_in_launch)
//   double loglik0() const;           initial log-likelihood (psi: approx_loglik)
//   double loglik_end() const;        added once at the end (bsf: constants)
//   bool observed(int t) const;
//   void begin_weight(int t);         load per-step constants
//   double logw(const double* x) const;
//   void begin_step(int t);           load per-step constants for t -> t+1
//   init_all<P>(ctx, x), Noise<P>, draw<P>(ctx, noise), propagate_all<P>(ctx, noise, xa, x):
//     the random part of the step for the thread's P particles (GaussStep above for
//     models with K normals per particle; the SDE model runs its Milstein sub-steps)
// Model::DYN (models_ung_dyn.cuh): the state dimension is model.m <= Model::M, chosen at run time
template <class T, class = void> struct ModelDyn { static constexpr bool value = false; };
template <class T> struct ModelDyn<T, decltype((void)T::DYN)> { static constexpr bool value = T::DYN; };
template <class Model> __device__ __forceinline__ int model_dim(const Model& model) {
  if constexpr (ModelDyn<Model>::value) return model.m;
  else return Model::M;
}

template <class Model, int P, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
filter_kernel(const FilterParams fp, const typename Model::Params mp) {
  constexpr int M = Model::M;
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
  sync.counter = fp.counters + (size_t)fidx * COUNTER_STRIDE;
  sync.nranks = ranks;
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

  // row bases are multiples of N elements from 256-byte aligned allocations
  const bool vec2 = (N & 1u) == 0;  // 16-byte alignment of fp64 rows
  const bool vec4 = (N & 3u) == 0;  // 16-byte alignment of the uint32 ancestor column (8-byte for P = 2 needs N even)
  Model model(mp, fidx);
  const int Md = model_dim(model); // rows of the state per time point (== M unless Model::DYN)
  double x[P][M];

  // ---- initial state, column 0 (reference :621-627 / :538-544)
  {
    StepCtx c0;
    c0.seed = fp.seed; c0.filter = filter; c0.t = 0u; c0.i0 = i0;
    model.template init_all<P>(c0, x);
#pragma unroll
    for (int d = 0; d < M; ++d) {
      if (d >= Md) break;
      double col[P];
#pragma unroll
      for (int p = 0; p < P; ++p) col[p] = x[p][d];
      store_row<P>(alpha + (size_t)d * N, i0, N, vec2, col);
    }
  }

  double loglik = model.loglik0();
  int stopped = -1;
  double carry[P];
#pragma unroll
  for (int p = 0; p < P; ++p) carry[p] = 0.0;
  double* ub = fp.ubuf + (size_t)fidx * N;
  Strata st;
  st.ubuf = ub; st.N = N; st.Nd = Nd;

#if PF_PHASE_TIMING
  unsigned long long ph[6] = {0, 0, 0, 0, 0, 0};
  const bool timing = fp.phase_cycles != nullptr && threadIdx.x == 0;
  long long tc = timing ? clock64() : 0;
#define PF_PHASE(i) if (timing) { long long now_ = clock64(); ph[i] += (unsigned long long)(now_ - tc); tc = now_; }
#else
#define PF_PHASE(i)
#endif

  for (int t = 0;; ++t) {
    // ---- weights of column t (reference :633-647, :672-687).  Column n is a
    // pure prediction: unit weights (the reference tests t < n-1 before
    // weighting column t+1).
    const bool obs = (t < n) && model.observed(t);
    double w[P], lw[P];
    double mx = 0.0;
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
      // a NaN log-weight must stop the filter (the reference's sum of weights becomes NaN):
      // +inf survives the plain max of the collective and fails the finiteness test below
      if (bad) mx = pf_inf();
      PF_PHASE(0)
      coll.allmax_begin(mx);
    }
    // ---- stratified resampling -> ancestors of column t+1 (reference :650-655)
    uint32_t* anc = iout + (fp.keep_indices ? (size_t)t * N : 0);
    if (t < n) {
      // thresholds of the strata this thread owns, generated while the maxima
      // travel between the CTAs (visible to the filter after the scan barrier)
      double uth[P];
#pragma unroll
      for (int h = 0; h < P / 2; ++h) {
        const uint32_t j = i0 + 2 * h;
        pf_u32x4 rx = pf_philox4x32_10(j >> 1, 0u, (uint32_t)t, filter, fp.seed, (uint32_t)PF_RNG_RESAMPLE);
        uth[2 * h] = pf_u53_plus(rx.v[0], rx.v[1], pf_u32_to_f64(j)) * invN;
        uth[2 * h + 1] = pf_u53_plus(rx.v[2], rx.v[3], pf_u32_to_f64(j + 1)) * invN;
      }
      store_row<P>(ub, i0, N, vec2, uth);
    }
    if (obs) {
      mx = coll.allmax_end();
      PF_PHASE(1)
      if (!pf_isfinite(mx)) { stopped = t; break; } // uniform over the filter
#pragma unroll
      for (int p = 0; p < P; ++p) w[p] = pf_exp_nonpos(lw[p] - mx);
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) w[p] = (i0 + p < N) ? 1.0 : 0.0;
    }
    if (wout) store_row<P, true>(wout + (size_t)t * N, i0, N, vec2, w);
    if (t == n) break;

    StepCtx ctx;
    ctx.seed = fp.seed; ctx.filter = filter; ctx.t = (uint32_t)(t + 1); ctx.i0 = i0;
    typename Model::template Noise<P> noise;
    {
      // one scan of the fixed-point weights gives their sum and cumulative sums
      long long qh[P], ql[P], th = 0, tl = 0;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        pf_fixsum tm = pf_fixsum_term(w[p]);
        qh[p] = tm.hi; ql[p] = tm.lo;
        th += tm.hi; tl += tm.lo;
      }
      long long rh, rl, Th, Tl;
      PF_PHASE(2)
      coll.exscan2(th, tl, rh, rl, Th, Tl);
      PF_PHASE(3)
      const double S = pf_fixsum_value(Th, Tl);
      const double invS = 1.0 / S;
      if (obs && rank == 0 && threadIdx.x == 0) { // the log-likelihood is thread 0's business (a log and a division)
        // step_const: per-step normalising constant (ssm_ulg adds norm_log_const(H_t) every
        // step, src/model_ssm_ulg.cpp:588, :631); zero for the other models
        if (t == 0) loglik = ((mx + loglik) + pf_log_normal(S / Nd)) + model.step_const(); // reference :561
        else loglik += (mx + pf_log_normal(S / Nd)) + model.step_const();                  // reference :601
      }
      // offspring range of particle k is [count(p_{k-1}), count(p_k)); p_{N-1} := 1.
      // Particles are resolved in chunks of CH so that only CH+1 cumulative
      // probabilities / threshold loads are live at a time.
      constexpr int CH = P > 4 ? 4 : P;
      double pk0 = (i0 == 0) ? 0.0 : ((i0 - 1 >= N - 1) ? 1.0 : pf_fixsum_value(rh, rl) * invS);
      uint32_t j0 = st.guess(pk0);
      double u0 = st.u(j0);
      uint32_t lo = (i0 == 0) ? 0u : st.resolve(pk0, j0, u0);
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
          // the owner writes the first PF_SCATTER_CAP offspring itself (predicated stores, no loop),
          // a longer run (a heavy particle) is written by the whole warp, coalesced
          const uint32_t cnt = c > lo ? c - lo : 0u;
          if (cnt) {
            uint32_t* q = anc + lo;
            q[0] = val;
            if (cnt > 1) {
              q[1] = val;
              if (cnt > 2) {
                q[2] = val;
                if (cnt > 3) q[3] = val;
              }
            }
          }
          const uint32_t ce = lo + PF_SCATTER_CAP;
          const unsigned heavy = __ballot_sync(0xffffffffu, cnt > PF_SCATTER_CAP);
          if (heavy) scatter_heavy(anc, heavy, ce, c, val);
          lo = c > lo ? c : lo;
        }
      }
    }
    coll.barrier_arrive();
    // normals of the next column, generated while the scatter barrier completes
    model.template draw<P>(ctx, noise);
    coll.barrier_wait();
    PF_PHASE(4)
    // columns 0..t of state, weights and ancestors are complete and visible to this thread
    if (fp.progress && rank == 0 && threadIdx.x == 0 && ((t + 1) % fp.progress_every) == 0) {
      __threadfence_system();
      *((volatile int*)fp.progress + fidx) = t + 1;
    }

    // ---- gather ancestors, propagate to column t+1 (reference :657-670)
    {
      model.begin_step(t);
      const double* src = alpha + (size_t)(fp.keep_alpha ? t : (t & 1)) * Md * N;
      double* dst = alpha + (size_t)(fp.keep_alpha ? t + 1 : ((t + 1) & 1)) * Md * N;
      uint32_t a[P];
      load_anc<P>(anc, i0, N, P == 2 ? vec2 : vec4, a);
      double xa[P][M];
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int d = 0; d < M; ++d) xa[p][d] = d < Md ? __ldcg(src + (size_t)d * N + a[p]) : 0.0;
      model.template propagate_all<P>(ctx, noise, xa, x, carry);
#pragma unroll
      for (int d = 0; d < M; ++d) {
        if (d >= Md) break;
        double col[P];
#pragma unroll
        for (int p = 0; p < P; ++p) col[p] = x[p][d];
        store_row<P>(dst + (size_t)d * N, i0, N, vec2, col);
      }
    }
    PF_PHASE(5)
  }
#undef PF_PHASE
#if PF_PHASE_TIMING
  if (timing)
    for (int i = 0; i < 6; ++i) fp.phase_cycles[(size_t)blockIdx.x * 6 + i] = ph[i];
#endif

  if (rank == 0 && threadIdx.x == 0) {
    if (stopped >= 0) loglik = -pf_inf();
    else loglik += model.loglik_end();
    fp.loglik[fidx] = loglik;
    fp.stopped_at[fidx] = stopped;
  }
}

} // namespace pf
