// sim_smoother.cuh -- the simulation smoother of the approximating Gaussian model and the SPDK importance
// weights (non-sequential importance sampling, Shephard-Pitt / Durbin-Koopman), one thread per simulated
// trajectory (per antithetic PAIR of trajectories):
//
//   ssm_ulg::simulate_states        src/model_ssm_ulg.cpp:141-247   (nsim > 1 branch, antithetic variable)
//   ssm_ulg::fast_smoother(F,K,L)   src/model_ssm_ulg.cpp:311-357   (the gains come from the host, once per model)
//   ssm_ung::importance_weights     src/model_ssm_ung.cpp:372-378   = sum_t log_weights(t, alpha) (:385-439)
//
// behind importance_sample() (src/R_importance_sample.cpp:9-75) and logLik(method = "spdk")
// (src/model_ssm_ung.cpp:199-211).  Every thread runs the three sweeps of the reference over its own
// trajectory: (1) simulate alpha+ and y+ forwards and filter y+ with the stored gains in the same sweep,
// (2) the backward recursion for r_t, (3) the forward recursion for the smoothed mean of the simulated series,
// combined on the fly with alpha+ and the smoothed mean of the data into the two antithetic draws, whose
// log-weights accumulate in the same sweep.  Storage is time-major (trajectory fastest), so that all
// accesses of a warp are coalesced: alpha (n+1) x m x nsim, r n x m x threads, v n x threads.
#pragma once
#include <cstdint>
#include "bssm_b200/pf_math.h"
#include "models_ung.cuh"

namespace pf {

// one model's inputs of the simulation smoother as offsets into its block of doubles (the host packs the model
// arrays; the approximation and the gains come from the host for one model and from spdk_gains_kernel for a
// batch of stored draws)
struct SimBlock {
  size_t y, u, xbeta, Z, T, R, C, D, a1, P1, LP1, RR, ay, aH, rH, yphi2, F, K, Lt, ahat, phi, total;
  int has_xbeta;
};

struct SimParams {
  int n, m, k, distribution;
  int Ztv, Ttv, Rtv, Ctv, Dtv;
  const double *y, *u, *xbeta; // xbeta: n values or nullptr
  const double *Z, *T, *R, *C, *D, *a1, *P1, *LP1, *RR;
  const double *ay, *aH, *rH, *yphi2; // pseudo-observations, their sd, 1 / sd, (y / phi)^2
  const double *F, *K, *Lt, *ahat;    // fast_precomputing_smoother of the data
  double phi;
  uint32_t nsim, nsim2, threads, seed;
  int antithetic;
  double* alpha; // out: (n+1) x m x nsim
  double* logw;  // out: nsim, sum_t log_weights(t, alpha_i)
  double *v, *r; // scratch: n x threads, n x m x threads
  // a batch of models (is_correction_spdk, src/approx_mcmc.cpp:606-678: one model and one approximation per
  // stored draw): every input pointer above points into the block of draw 0 and draw f's block lies
  // f * fstride doubles further on (phi: phi_f[f * fstride]); toff / soff (nf + 1 each) are the running sums of
  // the draws' threads and trajectories, which also place their alpha / logw / v / r regions; draw f uses the
  // random stream `first_filter + f`.  nf = 0: one model, the scalar members above.
  uint32_t nf, first_filter, total_threads;
  size_t fstride;
  const uint32_t *toff, *soff;
  const double* phi_f;
};

// the members of draw f of a batch, as a single-model parameter block
__device__ __forceinline__ SimParams sim_params_of(const SimParams& p, uint32_t f) {
  SimParams q = p;
  const size_t o = (size_t)f * p.fstride;
  q.y += o; q.u += o; if (q.xbeta) q.xbeta += o;
  q.Z += o; q.T += o; q.R += o; q.C += o; q.D += o; q.a1 += o; q.P1 += o; q.LP1 += o; q.RR += o;
  q.ay += o; q.aH += o; q.rH += o; q.yphi2 += o; q.F += o; q.K += o; q.Lt += o; q.ahat += o;
  q.phi = p.phi_f[o];
  const uint32_t t0 = p.toff[f], s0 = p.soff[f];
  q.nsim = p.soff[f + 1] - s0;
  q.threads = p.toff[f + 1] - t0;
  q.nsim2 = p.antithetic ? q.nsim / 2 : q.nsim;
  q.alpha += (size_t)s0 * (p.n + 1) * p.m;
  q.logw += s0;
  q.v += (size_t)t0 * p.n;
  q.r += (size_t)t0 * p.n * p.m;
  return q;
}

template <int MAXM>
__device__ void sim_smoother_thread(const SimParams& p, uint32_t i, uint32_t filter);

template <int MAXM>
__global__ void __launch_bounds__(128) sim_smoother_kernel(const SimParams p) {
  const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (p.nf == 0) {
    if (gi < p.threads) sim_smoother_thread<MAXM>(p, gi, p.first_filter);
    return;
  }
  if (gi >= p.total_threads) return;
  uint32_t lo = 0, hi = p.nf; // the draw whose thread range holds gi
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (p.toff[mid] <= gi) lo = mid; else hi = mid;
  }
  const SimParams q = sim_params_of(p, lo);
  sim_smoother_thread<MAXM>(q, gi - p.toff[lo], p.first_filter + lo);
}

template <int MAXM>
__device__ void sim_smoother_thread(const SimParams& p, const uint32_t i, const uint32_t filter) {
  const int n = p.n, m = p.m, k = p.k, K1 = p.k + 1;
  const size_t NS = p.nsim, NT = p.threads;
  // the trajectory without a partner: the last one of an odd nsim (:195-217)
  const bool single = p.antithetic ? (i == p.nsim2) : true;
  const bool lone = p.antithetic && i == p.nsim2;
  const uint32_t slot = lone ? p.nsim - 1 : i;
  constexpr double zero_tol = 1e-12; // src/model_ssm_ulg.h:19
  double a[MAXM], fa[MAXM], nx[MAXM], tmp[MAXM], z[MAXM + 1];

  // consecutive flat indices share a Box-Muller pair: one Philox call serves both (same values as pf_sim_normal)
  uint32_t pair_t = 0xffffffffu, pair_id = 0xffffffffu;
  double pz0 = 0.0, pz1 = 0.0;
  auto sim_normal = [&](uint32_t t, uint32_t f) {
    const uint32_t pr = f >> 1;
    if (pr != pair_id || t != pair_t) {
      pf_normal_pair_of((uint32_t)PF_RNG_SIM, p.seed, filter, t, 0u, pr, &pz0, &pz1);
      pair_id = pr; pair_t = t;
    }
    return (f & 1u) ? pz1 : pz0;
  };

  // ---- sweep 1: alpha+, y+ and the filter of y+ (:165-182, :323-330)
  for (int d = 0; d < m; ++d) z[d] = sim_normal(0u, i * (uint32_t)m + d);
  for (int d = 0; d < m; ++d) {
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += p.LP1[d + m * j] * z[j];
    a[d] = p.a1[d] + s;
    fa[d] = p.a1[d];
    p.alpha[(size_t)d * NS + slot] = a[d];
  }
  for (int t = 0; t < n; ++t) {
    const double* Z = p.Z + (size_t)m * (t * p.Ztv);
    const double* T = p.T + (size_t)m * m * (t * p.Ttv);
    const double* R = p.R + (size_t)m * k * (t * p.Rtv);
    const double* C = p.C + (size_t)m * (t * p.Ctv);
    const double Dt = p.D[t * p.Dtv];
    const double xb = p.xbeta ? p.xbeta[t] : 0.0;
    for (int d = 0; d < K1; ++d) z[d] = sim_normal((uint32_t)(t + 1), i * (uint32_t)K1 + d);
    double yp = p.y[t]; // stays missing where the data are missing
    if (pf_isfinite(yp)) {
      double za = 0.0;
      for (int d = 0; d < m; ++d) za += Z[d] * a[d];
      yp = xb + Dt + za + p.aH[t] * z[0];
    }
    for (int d = 0; d < m; ++d) {
      double ta = 0.0, ru = 0.0;
      for (int j = 0; j < m; ++j) ta += T[d + m * j] * a[j];
      for (int j = 0; j < k; ++j) ru += R[d + m * j] * z[1 + j];
      nx[d] = C[d] + ta + ru;
    }
    for (int d = 0; d < m; ++d) {
      a[d] = nx[d];
      p.alpha[((size_t)(t + 1) * m + d) * NS + slot] = nx[d];
    }
    const double ytmp = p.xbeta ? yp - xb : yp;
    if (pf_isfinite(ytmp) && p.F[t] > zero_tol) {
      double za = 0.0;
      for (int d = 0; d < m; ++d) za += Z[d] * fa[d];
      const double vt = ytmp - Dt - za;
      p.v[(size_t)t * NT + i] = vt;
      for (int d = 0; d < m; ++d) tmp[d] = fa[d] + p.K[d + (size_t)m * t] * vt;
    } else {
      for (int d = 0; d < m; ++d) tmp[d] = fa[d];
    }
    for (int d = 0; d < m; ++d) {
      double ta = 0.0;
      for (int j = 0; j < m; ++j) ta += T[d + m * j] * tmp[j];
      nx[d] = C[d] + ta;
    }
    for (int d = 0; d < m; ++d) fa[d] = nx[d];
  }
  // fa = column n of the smoothed mean (the recursion of sweep 3 stops at n - 1, :352-354)

  // ---- sweep 2: r_{t-1} = Z_t v_t / F_t + L_t' r_t (:335-343); r_{n-1} = 0
  double r[MAXM];
  for (int d = 0; d < m; ++d) { r[d] = 0.0; p.r[((size_t)(n - 1) * m + d) * NT + i] = 0.0; }
  for (int t = n - 1; t > 0; --t) {
    const double* Z = p.Z + (size_t)m * (t * p.Ztv);
    const bool used = pf_isfinite(p.y[t]) && p.F[t] > zero_tol;
    const double* L = used ? p.Lt + (size_t)m * m * t : p.T + (size_t)m * m * (t * p.Ttv);
    for (int d = 0; d < m; ++d) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += L[j + m * d] * r[j]; // L' r
      tmp[d] = s;
    }
    if (used) {
      const double vt = p.v[(size_t)t * NT + i], Ft = p.F[t];
      for (int d = 0; d < m; ++d) tmp[d] = Z[d] / Ft * vt + tmp[d];
    }
    for (int d = 0; d < m; ++d) { r[d] = tmp[d]; p.r[((size_t)(t - 1) * m + d) * NT + i] = tmp[d]; }
  }
  // column 0 (:344-349)
  double at[MAXM];
  {
    const bool used = pf_isfinite(p.y[0]) && p.F[0] > zero_tol;
    if (used) {
      const double* L = p.Lt;
      const double vt = p.v[i], Ft = p.F[0];
      for (int d = 0; d < m; ++d) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += L[j + m * d] * r[j];
        tmp[d] = p.Z[d] / Ft * vt + s;
      }
      for (int d = 0; d < m; ++d) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += p.P1[d + m * j] * tmp[j];
        at[d] = p.a1[d] + s;
      }
    } else if (m == 1) {
      at[0] = p.a1[0] + p.P1[0] * p.T[0] * r[0]; // 1 x 1 operands associate from the left (ung_gauss_approx.h)
    } else {
      for (int d = 0; d < m; ++d) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += p.T[j + m * d] * r[j];
        tmp[d] = s;
      }
      for (int d = 0; d < m; ++d) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += p.P1[d + m * j] * tmp[j];
        at[d] = p.a1[d] + s;
      }
    }
  }

  // ---- sweep 3: smoothed mean of the simulated series (:352-354), the two draws (:184-188) and their weights
  double w0 = 0.0, w1 = 0.0;
  const uint32_t mate = i + p.nsim2;
  for (int t = 0; t <= n; ++t) {
    if (t == n) {
      for (int d = 0; d < m; ++d) at[d] = fa[d];
    } else if (t > 0) {
      const double* T = p.T + (size_t)m * m * ((t - 1) * p.Ttv);
      const double* C = p.C + (size_t)m * ((t - 1) * p.Ctv);
      const double* RR = p.RR + (size_t)m * m * ((t - 1) * p.Rtv);
      for (int d = 0; d < m; ++d) r[d] = p.r[((size_t)(t - 1) * m + d) * NT + i];
      for (int d = 0; d < m; ++d) {
        double ta = 0.0, rr = 0.0;
        for (int j = 0; j < m; ++j) ta += T[d + m * j] * at[j];
        for (int j = 0; j < m; ++j) rr += RR[d + m * j] * r[j];
        nx[d] = C[d] + ta + rr;
      }
      for (int d = 0; d < m; ++d) at[d] = nx[d];
    }
    double x0[MAXM], x1[MAXM];
    for (int d = 0; d < m; ++d) {
      const size_t o = ((size_t)t * m + d) * NS;
      const double ap = p.alpha[o + slot], ah = p.ahat[d + (size_t)m * t];
      if (lone) {
        x0[d] = ah - at[d] + ap; // :216
      } else {
        const double x = -at[d] + ap; // :184
        if (p.antithetic) { x1[d] = ah - x; p.alpha[o + mate] = x1[d]; } // :186
        x0[d] = x + ah;                                                  // :188
      }
      p.alpha[o + slot] = x0[d];
    }
    if (t < n && pf_isfinite(p.y[t])) {
      const double* Z = p.Z + (size_t)m * (t * p.Ztv);
      const double y = p.y[t], u = p.u[t], rH = p.rH[t], ay = p.ay[t], yp2 = p.yphi2[t];
      const double Dt = p.D[t * p.Dtv], xb = p.xbeta ? p.xbeta[t] : 0.0;
      double s0, s1 = 0.0;
      if (p.distribution == 0) {
        s0 = x0[0];
        if (!single) s1 = x1[0];
      } else {
        double za = 0.0;
        for (int d = 0; d < m; ++d) za += Z[d] * x0[d];
        s0 = Dt + za + xb;
        if (!single) {
          za = 0.0;
          for (int d = 0; d < m; ++d) za += Z[d] * x1[d];
          s1 = Dt + za + xb;
        }
      }
      const double q0 = (ay - s0) * rH;
      w0 += ung_log_obs(p.distribution, s0, y, u, p.phi, yp2) + 0.5 * (q0 * q0);
      if (!single) {
        const double q1 = (ay - s1) * rH;
        w1 += ung_log_obs(p.distribution, s1, y, u, p.phi, yp2) + 0.5 * (q1 * q1);
      }
    }
  }
  p.logw[slot] = w0;
  if (!single) p.logw[mate] = w1;
}

} // namespace pf
