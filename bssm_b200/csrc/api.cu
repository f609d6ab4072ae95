// api.cu -- C ABI of the engine (include/bssm_b200.h): host-side logic that
// lays the reference's model list out on the device, picks a launch geometry,
// runs pre-pass -> filter -> post-processing kernels on one stream and hands
// results back in the reference's layouts.
//
// Host-side model logic restated here (not hot path, O(n) or O(m^3)):
//   ssm_ung::approximate   /root/reference/src/model_ssm_ung.cpp:87-140
//   ssm_ulg::fast_smoother src/model_ssm_ulg.cpp:252-305 (scalar case)
//   chol of P1 sub-block   src/model_ssm_ung.cpp:614-619
//   distribution constants src/distr_consts.cpp:7-67
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include <limits>
#include <mutex>
#include <algorithm>
#include <thread>
#include <dlfcn.h>
#include <nccl.h>

#include "bssm_b200.h"
#include "pf_device.cuh"
#include "models_ung.cuh"
#include "prepass_ung.cuh"
#include "smoother_kernels.cuh"
#include "approx_device.cuh"
#include "sim_smoother.cuh"
#include "models_nlg_sde.cuh"
#include "filters.h"
#include "bssm_b200/ung_gauss_approx.h"
#include "bssm_b200/nlg_gauss_approx.h"
#include "bssm_b200/mng_gauss_approx.h"
#include "bssm_b200/pf_ekpf.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail(BSSM_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

int require_device() {
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt == 0) {
    cudaGetLastError();
    return fail(BSSM_B200_ERR_NO_DEVICE,
                "no CUDA device: bssm_b200 has no CPU fallback (cudaGetDeviceCount: " +
                    std::string(e == cudaSuccess ? "0 devices" : cudaGetErrorString(e)) + ")");
  }
  return 0;
}

// ---------------------------------------------------------------- geometry
using pfl::Geometry;

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
uint32_t round32(uint32_t a) { return cdiv(a, 32) * 32; }

// Launch geometry.  One CTA while the filter fits (N <= 4096); otherwise a group
// of `ranks` CTAs of 256 threads synchronised through L2 (SYNC_GROUP).  Few
// filters: many ranks with few particles per thread (latency); many filters:
// fewer ranks with P = 8, several CTAs of different filters per SM (throughput).
// Results never depend on the geometry (fixed-point sums), so this is free to tune.
// Environment overrides (experiments): BSSM_B200_RANKS / _P / _BLOCK / _SYNC=cluster.
int choose_geometry(uint32_t N, uint32_t n_filters, Geometry* g, bool p2_only = false) {
  // the low word of the fixed-point weight terms is < 2^43 per particle (pf_math.h): its sum stays
  // inside an int64 for N <= 2^20, which is also what one resident group of CTAs holds comfortably
  if (N > (1u << 20)) return fail(BSSM_B200_ERR_UNSUPPORTED, "at most 2^20 particles per filter");
  uint32_t ranks = 1, block;
  int P;
  int mode = pf::SYNC_BLOCK;
  // small filters: 2 particles per thread for latency; many of them (post_correct, suggest_N): 4 per
  // thread amortise the per-thread work better (C4: 102 -> 93 ms; 8 per thread: 108 ms)
  if (N <= 1024) { P = (n_filters >= 64 && N >= 128) ? 4 : 2; }
  else if (N <= 4096) { P = 4; }
  else {
    mode = pf::SYNC_GROUP;
    const uint32_t tb = 256;
    const uint32_t capacity = 148u * 4u;                   // co-resident CTAs of 256 threads at 64 registers
    const uint32_t rmin = cdiv(N, tb * 8), rmax = cdiv(N, tb * 2);
    uint32_t want = capacity / (n_filters ? n_filters : 1);
    if (want > 148u && n_filters == 1) want = 148u;        // one CTA per SM is enough for a single filter
    ranks = want < rmin ? rmin : (want > rmax ? rmax : want);
    const uint32_t per_thread = cdiv(N, ranks * tb);
    P = per_thread <= 2 ? 2 : (per_thread <= 4 ? 4 : 8);
    ranks = cdiv(N, tb * (uint32_t)P);
    if (ranks > capacity) return fail(BSSM_B200_ERR_UNSUPPORTED, "too many particles for one resident group (N <= 1.2M)");
  }
  // the latency path (pf_small.cuh): few small filters cannot fill the device, what counts is the length of
  // one step's dependency chain.  Many small filters (post_correct) are a throughput problem: general kernel.
  const uint32_t small_max = (uint32_t)env_int("BSSM_B200_SMALL_MAX_FILTERS", 148);
  const bool small_path = !p2_only && N <= 1024 && n_filters <= small_max && !env_int("BSSM_B200_NO_SMALL", 0);
  if (p2_only) { // run-time-m policies: two particles per thread, whatever the size
    P = 2;
    if (N <= 2048) { ranks = 1; mode = pf::SYNC_BLOCK; }
    else { mode = pf::SYNC_GROUP; ranks = cdiv(N, 256u * 2u); if (ranks > 148u * 4u) return fail(BSSM_B200_ERR_UNSUPPORTED, "too many particles for the general-m kernels"); }
  }
  const int forced_ranks = env_int("BSSM_B200_RANKS", 0);
  const int forced_P = env_int("BSSM_B200_P", 0);
  if (!p2_only && (forced_P == 2 || forced_P == 4 || forced_P == 8)) P = forced_P;
  if (forced_ranks > 0) { ranks = (uint32_t)forced_ranks; mode = ranks == 1 ? pf::SYNC_BLOCK : pf::SYNC_GROUP; }
  const char* sm = getenv("BSSM_B200_SYNC");
  if (sm && !strcmp(sm, "cluster") && ranks > 1 && ranks <= 16) mode = pf::SYNC_CLUSTER;
  block = round32(cdiv(cdiv(N, ranks), (uint32_t)P));
  const int forced_block = env_int("BSSM_B200_BLOCK", 0);
  if (forced_block >= 32 && (uint32_t)forced_block * P * ranks >= N) block = round32((uint32_t)forced_block);
  if (block > 1024) return fail(BSSM_B200_ERR_ARG, "geometry: too many particles per CTA");
  if (block < 32) block = 32;
  if (small_path && forced_ranks <= 0 && forced_P == 0 && forced_block == 0) {
    P = 2; ranks = 1; mode = pf::SYNC_SMALL; // two particles per consumer thread: measured best (pf_small.cuh)
    block = 2 * round32(cdiv(N, (uint32_t)P)); // consumers + producers
  }
  g->P = P; g->block = (int)block; g->ranks = ranks; g->sync_mode = mode;
  return 0;
}

// --------------------------------------------------------------- launching
// status of the pfl:: layer (launch.cuh) -> C ABI status + message
int launch_status(int rc, const char* what) {
  if (rc == 0) return 0;
  if (rc == -2) return fail(BSSM_B200_ERR_ARG, std::string(what) + ": " + pfl::user_model_error());
  if (rc < 0) return fail(BSSM_B200_ERR_UNSUPPORTED, std::string(what) + ": variant not instantiated in this build");
  return fail(BSSM_B200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString((cudaError_t)rc));
}

// ------------------------------------------------------- host model pieces
bool finite(double x) { return std::isfinite(x); }

// src/distr_consts.cpp:7-67 through the switch at src/model_ssm_ung.cpp:690-710
double ung_dist_const(const bssm_ung_model* M) {
  auto lch = [](double n, double k) {
    return -std::log(n + 1) - std::lgamma(n - k + 1) - std::lgamma(k + 1) + std::lgamma(n + 2);
  };
  double res = 0.0;
  const double phi = M->phi;
  for (int t = 0; t < M->n; ++t) {
    const double y = M->y[t];
    if (!finite(y)) continue;
    const double u = M->u[t];
    switch (M->distribution) {
      case 0: res += -0.5 * std::log(2.0 * M_PI) - std::log(phi); break;
      case 1: res += y * std::log(u) - std::lgamma(y + 1); break;
      case 2: res += lch(u, y); break;
      case 3: res += lch(y + phi - 1, y) + phi * std::log(phi) + y * std::log(u); break;
      case 4: res += phi * std::log(phi) - std::lgamma(phi) + (phi - 1) * std::log(y) - phi * std::log(u); break;
    }
  }
  return res;
}

// lower Cholesky factor of the positive-diagonal block of P1 (:614-619)
void chol_posdiag(const double* P1, int m, std::vector<double>& L) {
  L.assign((size_t)m * m, 0.0);
  std::vector<int> idx;
  for (int d = 0; d < m; ++d)
    if (P1[d + (size_t)m * d] > 0) idx.push_back(d);
  const int q = (int)idx.size();
  for (int c = 0; c < q; ++c)
    for (int r = c; r < q; ++r) {
      double s = P1[idx[r] + (size_t)m * idx[c]];
      for (int l = 0; l < c; ++l) s -= L[idx[r] + (size_t)m * idx[l]] * L[idx[c] + (size_t)m * idx[l]];
      L[idx[r] + (size_t)m * idx[c]] = (r == c) ? std::sqrt(s) : s / L[idx[c] + (size_t)m * idx[c]];
    }
}

// psd_chol, src/psd_chol.cpp:5-18
void psd_chol_host(const double* x, int m, double* L) {
  const double eps = std::numeric_limits<double>::epsilon();
  double mx = x[0];
  for (int d = 1; d < m; ++d) mx = std::max(mx, x[d + (size_t)m * d]);
  std::vector<double> sub((size_t)m * m, 0.0);
  for (int c = 0; c < m; ++c)
    for (int r = 0; r < m; ++r)
      sub[r + (size_t)m * c] = (x[r + (size_t)m * r] > std::max(eps, eps * m * mx) &&
                                x[c + (size_t)m * c] > std::max(eps, eps * m * mx)) ? x[r + (size_t)m * c] : 0.0;
  std::vector<double> Lv;
  chol_posdiag(sub.data(), m, Lv);
  for (int e = 0; e < m * m; ++e) L[e] = Lv[e];
}

// Newton/Laplace iterations for the mode of the Gaussian approximation, m = 1:
// ssm_ung::approximate (:87-140, local branch) = repeat { laplace_iter (:302-342);
// fast_smoother (src/model_ssm_ulg.cpp:252-305) } until the mean squared change
// of the mode is below conv_tol.  Sequential in t, O(n) per sweep: host work.
void ung_approximate_m1(const bssm_ung_model* M, std::vector<double>& mode, std::vector<double>& mode_lin) {
  const int n = M->n;
  const double zero_tol = 1e-12;
  const double phi = M->phi;
  mode.assign(M->initial_mode, M->initial_mode + n);
  mode_lin = mode;
  std::vector<double> HH(n), yt(n), at(n + 1), v(n), F(n), K(n), r(n), next(n);
  auto Z = [&](int t) { return M->Z[t * M->Ztv]; };
  auto T = [&](int t) { return M->T[t * M->Ttv]; };
  auto RR = [&](int t) { double x = M->R[t * M->Rtv]; return x * x; };
  auto C = [&](int t) { return M->C[t * M->Ctv]; };
  auto D = [&](int t) { return M->D[t * M->Dtv]; };
  auto xb = [&](int t) { return M->xbeta ? M->xbeta[t] : 0.0; };
  auto yobs = [&](int t) { return M->xbeta ? yt[t] - M->xbeta[t] : yt[t]; };
  int it = 0;
  double diff = M->conv_tol + 1;
  while (it < M->max_iter && diff > M->conv_tol) {
    it++;
    mode_lin = mode;
    for (int t = 0; t < n; ++t) { // laplace_iter
      const double y = M->y[t], u = M->u[t], s = mode[t];
      switch (M->distribution) {
        case 0: {
          double tmp = y;
          if (std::fabs(tmp) < 1e-4) tmp = 1e-4;
          double q = tmp / phi;
          HH[t] = 2.0 * pf_exp(s) / (q * q);
          yt[t] = s + 1.0 - 0.5 * HH[t];
        } break;
        case 1: {
          HH[t] = 1.0 / (pf_exp(s) * u);
          yt[t] = y * HH[t] + s - 1.0;
        } break;
        case 2: {
          double e = pf_exp(s);
          HH[t] = (1.0 + e) * (1.0 + e) / (u * e);
          yt[t] = y * HH[t] + s - 1.0 - e;
        } break;
        case 3: {
          double e = pf_exp(s) * u;
          HH[t] = (phi + e) * (phi + e) / (phi * e * (y + phi));
          yt[t] = s + (phi + e) * (y - e) / ((y + phi) * e);
        } break;
        default: {
          double e = pf_exp(s) * u;
          HH[t] = e / (y * phi);
          yt[t] = s - e / y + 1;
        } break;
      }
    }
    // fast_smoother
    at[0] = M->a1[0];
    double Pt = M->P1[0];
    for (int t = 0; t < n; ++t) {
      F[t] = Z(t) * Pt * Z(t) + HH[t];
      const double yy = yobs(t);
      if (finite(yy) && F[t] > zero_tol) {
        K[t] = Pt * Z(t) / F[t];
        v[t] = yy - D(t) - Z(t) * at[t];
        at[t + 1] = C(t) + T(t) * (at[t] + K[t] * v[t]);
        const double tmp = 1.0 - K[t] * Z(t);
        Pt = T(t) * (tmp * Pt * tmp + K[t] * HH[t] * K[t]) * T(t) + RR(t);
      } else {
        at[t + 1] = C(t) + T(t) * at[t];
        Pt = T(t) * Pt * T(t) + RR(t);
      }
    }
    r[n - 1] = 0.0;
    for (int t = n - 1; t > 0; --t) {
      if (finite(yobs(t)) && F[t] > zero_tol) {
        const double L = T(t) * (1.0 - K[t] * Z(t));
        r[t - 1] = Z(t) / F[t] * v[t] + L * r[t];
      } else {
        r[t - 1] = T(t) * r[t];
      }
    }
    if (finite(yt[0]) && F[0] > zero_tol) {
      const double L = T(0) * (1.0 - K[0] * Z(0));
      at[0] = M->a1[0] + M->P1[0] * (Z(0) / F[0] * v[0] + L * r[0]);
    } else {
      at[0] = M->a1[0] + M->P1[0] * T(0) * r[0];
    }
    for (int t = 0; t < n - 1; ++t) at[t + 1] = C(t) + T(t) * at[t] + RR(t) * r[t];
    if (M->distribution == 0) for (int t = 0; t < n; ++t) next[t] = at[t];
    else for (int t = 0; t < n; ++t) next[t] = xb(t) + D(t) + Z(t) * at[t];
    // arma::accu(square(new - old)) / n, two accumulators [UNVENDORED]
    double a1 = 0.0, a2 = 0.0;
    int j;
    for (j = 1; j < n; j += 2) {
      a1 += (next[j - 1] - mode[j - 1]) * (next[j - 1] - mode[j - 1]);
      a2 += (next[j] - mode[j]) * (next[j] - mode[j]);
    }
    if ((j - 1) < n) a1 += (next[j - 1] - mode[j - 1]) * (next[j - 1] - mode[j - 1]);
    diff = (a1 + a2) / n;
    mode = next;
  }
}

// log g(y_t | signal) up to constants, src/model_ssm_ung.cpp:447-493 (host copy of pf::ung_log_obs)
static double ung_log_obs_host(int dist, double s, double y, double u, double phi, double yphi2) {
  switch (dist) {
    case 0: return -0.5 * (s + yphi2 * pf_exp(-s));
    case 1: return y * s - u * pf_exp(s);
    case 2: return y * s - u * pf_log1p(pf_exp(s));
    case 3: return y * s - (y + phi) * pf_log(phi + u * pf_exp(s));
    default: return -phi * (s + (y * pf_exp(-s) / u));
  }
}

// The psi-APF pre-pass for m = 1 on the HOST: what pf::ung_prepass_m1_kernel (prepass_ung.cuh) does for one
// filter, expression by expression and in the same order, so that the two give the same bits
// (tests/test_gpu_parity.py::test_host_prepass_equals_device_prepass).  It is a sequential O(n) chain of
// dependent divisions: one GPU warp needs 0.9 ms for n = 1000, one host core 0.05 ms.  Jobs whose filters
// share one model (replicates, the filter inside an MCMC loop) run it here, once; jobs with one model per
// filter (post_correct: thousands of independent chains) run the kernel.
struct PrepassHost {
  std::vector<double> ah, Lt, Ct, yt, Ht, sc, rHt, yphi2; // n + 1 each (rows 0..n-1 used by the n-sized ones)
  double approx_loglik = 0.0;
};
static void ung_prepass_m1_host(const bssm_ung_model* M, const double* mode, const double* mode_lin, double dist_const,
                                PrepassHost& o) {
  const int n = M->n, dist = M->distribution;
  const double phi = M->phi, zero_tol = 1e-12, log2pi = std::log(2.0 * M_PI);
  const double a1 = M->a1[0], P1 = M->P1[0];
  auto Zt = [&](int t) { return M->Z[t * M->Ztv]; };
  auto Tt = [&](int t) { return M->T[t * M->Ttv]; };
  auto RRt = [&](int t) { double r = M->R[t * M->Rtv]; return r * r; };
  auto Ct = [&](int t) { return M->C[t * M->Ctv]; };
  auto Dt = [&](int t) { return M->D[t * M->Dtv]; };
  const size_t n1 = (size_t)n + 1;
  o.ah.assign(n1, 0.0); o.Lt.assign(n1, 0.0); o.Ct.assign(n1, 0.0); o.yt.assign(n1, 0.0); o.Ht.assign(n1, 0.0);
  o.sc.assign(n1, 0.0); o.rHt.assign(n1, 0.0); o.yphi2.assign(n1, 0.0);
  std::vector<double> vt(n), Ft(n), Kt(n);
  double *ah = o.ah.data(), *Vt = o.Lt.data(), *cc = o.Ct.data(), *ay = o.yt.data(), *aH = o.Ht.data(), *sc = o.sc.data();
  // ---- laplace_iter (:302-342) and update_scales (:241-292)
  for (int t = 0; t < n; ++t) {
    const double y = M->y[t], u = M->u[t];
    double s = mode_lin[t];
    double HH, yt;
    switch (dist) {
      case 0: {
        double tmp = y;
        if (std::fabs(tmp) < 1e-4) tmp = 1e-4;
        double q = tmp / phi;
        HH = 2.0 * pf_exp(s) / (q * q);
        yt = s + 1.0 - 0.5 * HH;
      } break;
      case 1: {
        HH = 1.0 / (pf_exp(s) * u);
        yt = y * HH + s - 1.0;
      } break;
      case 2: {
        double e = pf_exp(s);
        HH = (1.0 + e) * (1.0 + e) / (u * e);
        yt = y * HH + s - 1.0 - e;
      } break;
      case 3: {
        double e = pf_exp(s) * u;
        HH = (phi + e) * (phi + e) / (phi * e * (y + phi));
        yt = s + (phi + e) * (y - e) / ((y + phi) * e);
      } break;
      default: {
        double e = pf_exp(s) * u;
        HH = e / (y * phi);
        yt = s - e / y + 1;
      } break;
    }
    const double H = std::sqrt(HH);
    ay[t] = yt;
    aH[t] = H;
    o.rHt[t] = 1.0 / H;
    { const double qy = y / phi; o.yphi2[t] = qy * qy; }
    Ft[t] = HH;
    double scale = 0.0;
    s = mode[t];
    if (std::isfinite(y)) {
      double qy = y / phi;
      double q = (yt - s) / H;
      scale = ung_log_obs_host(dist, s, y, u, phi, qy * qy) + 0.5 * (q * q);
    }
    sc[t] = scale;
  }
  auto ytmp = [&](int t) { return M->xbeta ? ay[t] - M->xbeta[t] : ay[t]; };
  // ---- approx_loglik (:519-528) and the forward pass of smoother_ccov (src/model_ssm_ulg.cpp:419-473)
  {
    double sumH = 0.0, sumR = 0.0;
    for (int t = 0; t < n; ++t) sumH += aH[t];
    { int nR = M->Rtv ? n : 1; for (int t = 0; t < nR; ++t) sumR += M->R[t]; }
    const bool run_ll = !(sumH + sumR < zero_tol);
    double gl = !run_ll ? -pf_inf() : 0.0;
    double at = a1, Pl = P1, aht = a1, Pt = P1;
    int nobs = 0;
    double sumlog = 0.0, acc1 = 0.0, acc2 = 0.0;
    ah[0] = a1; Vt[0] = P1;
    for (int t = 0; t < n; ++t) {
      const double z = Zt(t), Tk = Tt(t), RRk = RRt(t), Ck = Ct(t), Dk = Dt(t);
      const double HH = Ft[t], yt = ytmp(t);
      const bool yfin = std::isfinite(yt);
      if (run_ll) {
        double F = z * Pl * z + HH;
        if (yfin && F > zero_tol) {
          double v = yt - Dk - z * at;
          double K = Pl * z / F;
          at = Ck + Tk * (at + K * v);
          Pl = Tk * (Pl - K * K * F) * Tk + RRk;
          gl -= 0.5 * (log2pi + pf_log(F) + v * v / F);
        } else {
          at = Ck + Tk * at;
          Pl = Tk * Pl * Tk + RRk;
        }
      }
      if (std::isfinite(M->y[t])) { nobs++; sumlog += pf_log(aH[t]); }
      if (t & 1) acc2 += sc[t]; else acc1 += sc[t];
      const double F = z * Pt * z + HH;
      double Kk = 0.0, vk = 0.0, ahn, Vn;
      if (yfin && F > zero_tol) {
        Kk = Pt * z / F;
        vk = yt - Dk - z * aht;
        ahn = Ck + Tk * (aht + Kk * vk);
        const double tmp = 1.0 - Kk * z;
        Vn = Tk * (tmp * Pt * tmp + Kk * HH * Kk) * Tk + RRk;
      } else {
        ahn = Ck + Tk * aht;
        Vn = Tk * Pt * Tk + RRk;
      }
      Ft[t] = F; Kt[t] = Kk; vt[t] = vk; ah[t + 1] = ahn; Vt[t + 1] = Vn; cc[t] = Vn;
      aht = ahn; Pt = Vn;
    }
    const double norm = -0.5 * nobs * log2pi - sumlog;
    const double ct = dist_const - norm;
    o.approx_loglik = gl + ct + (acc1 + acc2);
  }
  // ---- backward pass
  {
    double rt = 0.0, Nt = 0.0;
    for (int t = n - 1; t >= 0; --t) {
      const double z = Zt(t), Tk = Tt(t);
      const double F = Ft[t], Kk = Kt[t], vk = vt[t];
      const bool used = std::isfinite(ytmp(t)) && F > zero_tol;
      const double L = used ? Tk * (1.0 - Kk * z) : Tk;
      const double zFv = used ? z / F * vk : 0.0;
      const double zzF = used ? z * z / F : 0.0;
      const double V = Vt[t], ck = cc[t], ahk = ah[t];
      const double ccn = V * L * (1.0 - Nt * ck);
      if (used) {
        rt = zFv + L * rt;
        Nt = zzF + L * Nt * L;
      } else {
        rt = L * rt;
        Nt = L * Nt * L;
      }
      const double ahn = ahk + V * rt;
      const double Vn = V - V * Nt * V;
      cc[t] = ccn; ah[t] = ahn; Vt[t] = Vn;
    }
    cc[n] = 0.0;
  }
  // ---- conditional_cov, scalar SVD branch (src/conditional_dist.cpp:8-32): t downwards, reads t - 1 unwritten
  for (int t = n; t >= 1; --t) {
    const double s = std::fabs(Vt[t - 1]);
    const double tmp = (s > 2.220446049250313e-16 * s) ? cc[t - 1] * (1.0 / s) : 0.0;
    double Vnew = Vt[t] - tmp * cc[t - 1];
    Vnew = std::sqrt(std::fabs(Vnew));
    Vt[t] = Vnew; cc[t] = tmp;
  }
  Vt[0] = std::sqrt(std::fabs(Vt[0]));
}

// Device memory is recycled between calls: a one-shot entry point allocates
// gigabytes of outputs, and cudaMalloc / cudaFree of that size cost more than
// the kernels.  Freed blocks are kept (up to BSSM_B200_POOL_GB, default 48) and
// handed back for requests of the same size class; bssm_b200_shutdown() releases
// them.
struct Pool {
  struct Block { void* p; size_t bytes; int device; }; // blocks are only reused on the device that owns them
  std::vector<Block> free_blocks;
  size_t cached = 0;
  std::mutex mu;
  static size_t size_class(size_t b) { return b <= 4096 ? 4096 : ((b + (1u << 20) - 1) >> 20) << 20; }
  cudaError_t alloc(void** p, size_t bytes) {
    const size_t sc = size_class(bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> g(mu);
      for (size_t i = 0; i < free_blocks.size(); ++i)
        if (free_blocks[i].bytes == sc && free_blocks[i].device == dev) {
          *p = free_blocks[i].p;
          cached -= sc;
          free_blocks.erase(free_blocks.begin() + i);
          return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, sc);
    if (e != cudaSuccess) { // out of memory: drop the cache and retry once
      cudaGetLastError();
      release_all();
      e = cudaMalloc(p, sc);
    }
    return e;
  }
  void free(void* p, size_t bytes) {
    const size_t sc = size_class(bytes);
    const size_t limit = (size_t)env_int("BSSM_B200_POOL_GB", 48) << 30;
    cudaPointerAttributes at;
    int dev = 0;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess) dev = at.device; else { cudaGetLastError(); cudaGetDevice(&dev); }
    std::lock_guard<std::mutex> g(mu);
    if (cached + sc > limit) { cudaFree(p); return; }
    free_blocks.push_back({p, sc, dev});
    cached += sc;
  }
  void release_all() {
    std::lock_guard<std::mutex> g(mu);
    for (auto& b : free_blocks) cudaFree(b.p);
    free_blocks.clear();
    cached = 0;
  }
};
Pool g_pool;

// Resources of the streaming fetch (run_fetch_streaming), created once per process:
// cudaHostAlloc / cudaFreeHost synchronise the device and cost milliseconds per call.
struct StreamCtx {
  std::mutex mu;          // one streaming call at a time
  int* flag = nullptr;    // host-mapped progress word
  int* dflag = nullptr;   // its device address
  cudaStream_t s2 = nullptr;
  cudaEvent_t done = nullptr;
  int device = -1;
  cudaError_t ensure() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (flag && dev == device) return cudaSuccess;
    release();
    if ((e = cudaHostAlloc((void**)&flag, sizeof(int), cudaHostAllocMapped)) != cudaSuccess) return e;
    if ((e = cudaHostGetDevicePointer((void**)&dflag, flag, 0)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming)) != cudaSuccess) return e;
    device = dev;
    return cudaSuccess;
  }
  void release() {
    if (done) cudaEventDestroy(done);
    if (s2) cudaStreamDestroy(s2);
    if (flag) cudaFreeHost(flag);
    done = nullptr; s2 = nullptr; flag = nullptr; dflag = nullptr; device = -1;
  }
};
StreamCtx g_sctx;

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t count = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) g_pool.free(p, count * sizeof(T));
    p = nullptr;
  }
  int alloc(size_t n) {
    release();
    count = n;
    if (n == 0) return 0;
    CU(g_pool.alloc((void**)&p, n * sizeof(T)));
    return 0;
  }
  int upload(const T* h, size_t n) {
    int rc = alloc(n);
    if (rc) return rc;
    if (n) CU(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
  }
};

} // namespace

// the EKF update of (a1, P1) with y_0 and its factor: the proposal of the first column of the
// EKF particle filter (src/model_ssm_nlg.cpp:1153-1160)
template <int MM>
static void ekpf_first_column(const bssm_nlg_model* M, pf::NlgParams& np) {
  pfga::NlgFns fn{M->model_id, M->theta, M->known_params};
  double a1[MM], P1[MM * MM], att[MM], Ptt[MM * MM], Zg[MM];
  fn.a1_fn(a1);
  fn.P1_fn(P1);
  if (std::isfinite(M->y[0])) {
    fn.Z_gn(0, a1, Zg);
    const double H = fn.H_fn(0, a1);
    pf_ekf_update<MM>(M->y[0], a1, P1, Zg, fn.Z_fn(0, a1), H * H, att, Ptt);
  } else {
    for (int d = 0; d < MM; ++d) att[d] = a1[d];
    for (int e = 0; e < MM * MM; ++e) Ptt[e] = P1[e];
  }
  pf_psd_chol<MM>(Ptt, np.L1);
  for (int d = 0; d < MM; ++d) np.att1[d] = att[d];
  for (int e = 0; e < MM * MM; ++e) np.P1[e] = P1[e];
}

// ------------------------------------------------------------------- job
enum { KIND_UNG = 0, KIND_NLG = 1, KIND_SDE = 2, KIND_MNG = 3 };

// per-draw pieces of an m > 1 ssm_ung-family model (what bsm_ng::update_model changes,
// src/model_bsm_ng.cpp:17-51) plus the stored mode of each draw: host arrays, one block per filter
struct DrawModels {
  std::vector<double> R, P1, phi, xbeta, modes; // nf x (m*k), nf x (m*m), nf, nf x n (or empty), nf x n
  // parameter sets (parset_ung, src/parset_ng.h:11-40): any of the model's arrays per filter, each slice
  // shaped like the model's own array; empty = shared.  R, P1, phi, xbeta above may be empty too then.
  std::vector<double> Z, T, a1, D, C;
  bool parset = false;
};

// parset_ung::update(model, i), src/parset_ng.cpp:84-124: the model with filter f's components swapped in
static void draw_model(const bssm_ung_model* M, const DrawModels& dm, size_t f, bssm_ung_model& Mi) {
  Mi = *M;
  const size_t n = (size_t)M->n, m = (size_t)M->m, k = (size_t)M->k;
  if (!dm.Z.empty()) Mi.Z = &dm.Z[m * (M->Ztv ? n : 1) * f];
  if (!dm.T.empty()) Mi.T = &dm.T[m * m * (M->Ttv ? n : 1) * f];
  if (!dm.R.empty()) Mi.R = &dm.R[m * k * (M->Rtv ? n : 1) * f];
  if (!dm.a1.empty()) Mi.a1 = &dm.a1[m * f];
  if (!dm.P1.empty()) Mi.P1 = &dm.P1[m * m * f];
  if (!dm.D.empty()) Mi.D = &dm.D[(M->Dtv ? n : 1) * f];
  if (!dm.C.empty()) Mi.C = &dm.C[m * (M->Ctv ? n : 1) * f];
  if (!dm.phi.empty()) Mi.phi = dm.phi[f];
  if (!dm.xbeta.empty()) Mi.xbeta = &dm.xbeta[n * f];
}

struct bssm_b200_job {
  int kind = KIND_UNG;
  int method = 0, post = 0, out_flags = 0;
  uint32_t N = 0, n_filters = 0, first_filter = 0, seed = 0;
  int n = 0, m = 0, k = 0;
  Geometry geo{};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float ms[3] = {0, 0, 0};
  int launches = 0;
  bool has_run = false;
  uint32_t nlg_max_iter = 100, nlg_iekf_iter = 0; // psi-APF jobs of ssm_nlg: settings of approximate(), kept for set_theta
  double nlg_conv_tol = 1e-8;
  bool dyn = false; // run-time state dimension (models_ung_dyn.cuh): (m, k) without a specialised instantiation
  bool xbeta_per_filter = false; // m = 1 psi-APF whose filters carry their own regression part (beta in theta)
  bool host_approx_per_filter = false; // psi-APF whose filters each carry a host-side approximation (any m)
  bool mng_draws = false; // ssm_mng with one model per filter (parameter sets)
  bool host_prepass = false; // m = 1 psi-APF whose filters share one model: the pre-pass ran on the host
  // model on device
  DevBuf<double> y, Z, T, R, a1, P1, LP1, D, C, xbeta, u, mode, mode_lin;
  // psi pre-pass results
  DevBuf<double> alphahat, Lt, Ct, ytilde, Htilde, scales, approx_loglik, phi_f, lconst_f, scratch, rHt, yphi2;
  size_t pstride = 0;
  // outputs
  DevBuf<double> alpha_tm, alpha_sm, weights, cube;
  DevBuf<uint32_t> indices;
  DevBuf<double> loglik;
  DevBuf<int32_t> stopped;
  DevBuf<unsigned long long> xch;
  DevBuf<unsigned int> counters;
  int wave = 0; // filters per launch in SYNC_GROUP mode (all CTAs of a launch must be co-resident)
  DevBuf<double> ubuf;
  DevBuf<unsigned long long> phases;
  DevBuf<double> s_at, s_att, s_Pt, s_Ptt, s_mean, s_cov, traj, draw_par;
  size_t alpha_stride = 0, weights_stride = 0, indices_stride = 0;
  bool keep_alpha = false, keep_weights = false, keep_indices = false;
  pf::UngParams up{};
  pf::NlgParams np{};
  pf::SdeParams sp{};
  int sde_model_id = 0, nlg_model_id = 0;
  DevBuf<double> lgy, theta_f, theta_u, known_u, init_u;
  pf::MngParams gp{};
  pf::NlgPsiParams npp{};
  DevBuf<double> stepc;
  DevBuf<double> phi_v, HH_v, hconst_v;
  DevBuf<int> dist_v;
  pf::PrepassParams pp{};
  pf::FilterParams fp{};
  ~bssm_b200_job() {
    // on error paths a kernel or an async copy may still be running on the stream: the buffers go back
    // to the recycling pool right after this, so wait for the stream first
    if (stream) cudaStreamSynchronize(stream);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

template <int M>
int run_post_M(bssm_b200_job* J) {
  const uint32_t N = J->N, nf = J->n_filters;
  const int n = J->n;
  if (J->post == BSSM_POST_FILTER_SUMMARY) {
    pf::filter_summary_kernel<M><<<dim3(n + 1, nf), 1024, 0, J->stream>>>(
        J->alpha_tm.p, J->weights.p, N, n, J->s_at.p, J->s_att.p, J->s_Pt.p, J->s_Ptt.p,
        J->alpha_stride, J->weights_stride);
    J->launches += 1;
  } else if (J->post == BSSM_POST_SMOOTHER || J->post == BSSM_POST_SMOOTHER_MERGE) {
    pf::backtrack_kernel<M><<<dim3(cdiv(N, 256), nf), 256, 0, J->stream>>>(
        J->alpha_tm.p, J->indices.p, J->alpha_sm.p, N, n, J->alpha_stride, J->indices_stride);
    pf::summary_kernel<M><<<dim3(n + 1, nf), N >= 1024 ? 1024 : (int)round32(N), 0, J->stream>>>(
        J->alpha_sm.p, N, J->s_mean.p, J->s_cov.p, J->alpha_stride);
    J->launches += 2;
  } else if (J->post == BSSM_POST_PICK) {
    pf::pick_trajectory_kernel<M><<<cdiv(nf, 128), 128, 0, J->stream>>>(
        J->alpha_tm.p, J->indices.p, J->traj.p, N, n, J->alpha_stride, J->indices_stride, J->seed,
        J->first_filter, nf);
    J->launches += 1;
  }
  CU(cudaGetLastError());
  return 0;
}

int run_post_dyn(bssm_b200_job* J) {
  const uint32_t N = J->N, nf = J->n_filters;
  const int n = J->n, m = J->m;
  const int bs = N >= 1024 ? 1024 : (int)round32(N);
  if (J->post == BSSM_POST_FILTER_SUMMARY) {
    pf::summary_dyn_kernel<<<dim3(n + 1, nf), bs, 0, J->stream>>>(J->alpha_tm.p, J->weights.p, N, n, m, J->s_at.p, J->s_Pt.p,
                                                              J->s_att.p, J->s_Ptt.p, J->alpha_stride, J->weights_stride);
    J->launches += 1;
  } else if (J->post == BSSM_POST_SMOOTHER || J->post == BSSM_POST_SMOOTHER_MERGE) {
    pf::backtrack_dyn_kernel<<<dim3(cdiv(N, 256), nf), 256, 0, J->stream>>>(J->alpha_tm.p, J->indices.p, J->alpha_sm.p, N, n, m,
                                                                         J->alpha_stride, J->indices_stride);
    pf::summary_dyn_kernel<<<dim3(n + 1, nf), bs, 0, J->stream>>>(J->alpha_sm.p, nullptr, N, n, m, J->s_mean.p, J->s_cov.p,
                                                              nullptr, nullptr, J->alpha_stride, 0);
    J->launches += 2;
  } else if (J->post == BSSM_POST_PICK) {
    pf::pick_trajectory_dyn_kernel<<<cdiv(nf, 128), 128, 0, J->stream>>>(
        J->alpha_tm.p, J->indices.p, J->traj.p, N, n, m, J->alpha_stride, J->indices_stride, J->seed, J->first_filter, nf);
    J->launches += 1;
  }
  CU(cudaGetLastError());
  return 0;
}

int run_post(bssm_b200_job* J) {
  if (J->m > 4) return run_post_dyn(J);
  switch (J->m) {
    case 1: return run_post_M<1>(J);
    case 2: return run_post_M<2>(J);
    case 3: return run_post_M<3>(J);
    case 4: return run_post_M<4>(J);
  }
  return fail(BSSM_B200_ERR_UNSUPPORTED, "post-processing supports m <= 4");
}

int launch_ung_filter(bssm_b200_job* J, uint32_t f0, uint32_t nf) {
  pf::FilterParams fp = J->fp;
  // offset per-filter pointers when launching a sub-range (grid mode: one filter per launch)
  fp.first_filter = J->first_filter + f0;
  fp.alpha_tm = J->alpha_tm.p + (size_t)f0 * J->alpha_stride;
  fp.weights = J->weights.p ? J->weights.p + (size_t)f0 * J->weights_stride : nullptr;
  fp.indices = J->indices.p + (size_t)f0 * J->indices_stride;
  fp.loglik = J->loglik.p + f0;
  fp.stopped_at = J->stopped.p + f0;
  fp.xch = J->xch.p + (size_t)f0 * 4 * J->geo.ranks;
  fp.counters = J->counters.p + (size_t)f0 * pf::COUNTER_STRIDE;
  fp.ubuf = J->ubuf.p + (size_t)f0 * J->N;
  if (fp.phase_cycles) fp.phase_cycles = J->phases.p + (size_t)f0 * 6 * J->geo.ranks;
  if (J->kind == KIND_NLG && J->method == BSSM_METHOD_PSI) {
    pf::NlgPsiParams npp = J->npp;
    npp.approx_loglik += f0;
    npp.stepc += (size_t)f0 * npp.fstride_stepc; npp.alphahat += (size_t)f0 * npp.fstride_a;
    npp.Lt += (size_t)f0 * npp.fstride_m; npp.Ct += (size_t)f0 * npp.fstride_m;
    return launch_status(pfl::launch_nlg_psi(J->nlg_model_id, fp, npp, J->geo, nf, J->stream), "nlg psi filter");
  }
  if (J->kind == KIND_NLG && J->method == BSSM_METHOD_EKPF)
    return launch_status(pfl::launch_nlg_ekpf(J->nlg_model_id, fp, J->np, J->geo, nf, J->stream), "nlg ekpf");
  if (J->kind == KIND_NLG && J->nlg_model_id >= BSSM_USER_MODEL_BASE) {
    pf::NlgParams np = J->np;
    np.theta_u += (size_t)f0 * np.theta_u_stride;
    np.init_u += (size_t)f0 * np.init_u_stride;
    return launch_status(pfl::launch_user_nlg(J->nlg_model_id, J->m, J->k, fp, np, J->geo, nf, np.theta_u_stride != 0, J->stream),
                         "nlg bsf (user model)");
  }
  if (J->kind == KIND_NLG && J->np.draw_blk) {
    pf::NlgParams np = J->np;
    np.draw_blk += (size_t)pf::NLG_DRAW_BLOCK * f0;
    return launch_status(pfl::launch_nlg_bsf_draws(J->nlg_model_id, fp, np, J->geo, nf, J->stream), "nlg bsf (draws)");
  }
  if (J->kind == KIND_NLG) return launch_status(pfl::launch_nlg_bsf(J->nlg_model_id, fp, J->np, J->geo, nf, J->stream), "nlg bsf");
  if (J->kind == KIND_SDE) {
    pf::SdeParams sp = J->sp;
    if (sp.theta_f) sp.theta_f += (size_t)6 * f0;
    if (sp.theta_u) sp.theta_u += (size_t)f0 * sp.theta_u_stride;
    if (J->sde_model_id >= BSSM_USER_MODEL_BASE)
      return launch_status(pfl::launch_user_sde(J->sde_model_id, fp, sp, J->geo, nf, sp.theta_u_stride != 0, J->stream),
                           "sde bsf (user model)");
    return launch_status(pfl::launch_sde_bsf(J->sde_model_id, fp, sp, J->geo, nf, J->stream), "sde bsf");
  }
  if (J->kind == KIND_MNG) {
    pf::MngParams gp = J->gp;
    // per-filter strides are 0 unless the job carries one model per filter (parameter sets)
    gp.Z += (size_t)f0 * gp.fs_Z; gp.T += (size_t)f0 * gp.fs_T; gp.R += (size_t)f0 * gp.fs_R; gp.C += (size_t)f0 * gp.fs_C;
    gp.D += (size_t)f0 * gp.fs_D; gp.a1 += (size_t)f0 * gp.fs_a1; gp.LP1 += (size_t)f0 * gp.fs_LP1;
    gp.phi += (size_t)f0 * gp.fs_phi;
    if (gp.lconst_f) gp.lconst_f += f0;
    const bool psi = J->method == BSSM_METHOD_PSI;
    if (psi) {
      gp.approx_loglik += f0;
      gp.alphahat += (size_t)f0 * gp.fs_a; gp.Lt += (size_t)f0 * gp.fs_m; gp.Ct += (size_t)f0 * gp.fs_m;
      gp.ytilde += (size_t)f0 * gp.fs_np; gp.Htilde += (size_t)f0 * gp.fs_np; gp.scales += (size_t)f0 * gp.fs_n;
    }
    if (J->dyn)
      return psi ? launch_status(pfl::launch_mng_psi_dyn(fp, gp, J->geo, nf, J->stream), "ssm_mng psi filter (run-time m)")
                 : launch_status(pfl::launch_mng_bsf_dyn(fp, gp, J->geo, nf, J->stream), "ssm_mng bsf (run-time m)");
    if (J->mng_draws)
      return psi ? launch_status(pfl::launch_mng_psi_draws(J->m, fp, gp, J->geo, nf, J->stream), "ssm_mng psi filter (draws)")
                 : launch_status(pfl::launch_mng_bsf_draws(J->m, J->k, fp, gp, J->geo, nf, J->stream), "ssm_mng bsf (draws)");
    return psi ? launch_status(pfl::launch_mng_psi(J->m, fp, gp, J->geo, nf, J->stream), "ssm_mng psi filter (m)")
               : launch_status(pfl::launch_mng_bsf(J->m, J->k, fp, gp, J->geo, nf, J->stream), "ssm_mng bsf (m,k)");
  }
  pf::UngParams up = J->up;
  up.Z += (size_t)f0 * up.fs_Z; up.D += (size_t)f0 * up.fs_D; // parameter sets: one model per filter
  up.T += (size_t)f0 * up.fs_T; up.R += (size_t)f0 * up.fs_R; up.C += (size_t)f0 * up.fs_C;
  up.a1 += (size_t)f0 * up.fs_a1; up.LP1 += (size_t)f0 * up.fs_LP1;
  if (up.lconst_f) up.lconst_f += f0;
  if (J->method == BSSM_METHOD_PSI && (J->m > 1 || J->xbeta_per_filter || J->host_approx_per_filter)) {
    up.approx_loglik += f0;
    up.alphahat += (size_t)f0 * up.fs_a; up.Lt += (size_t)f0 * up.fs_m; up.Ct += (size_t)f0 * up.fs_m;
    up.ytilde += (size_t)f0 * up.fs_n; up.Htilde += (size_t)f0 * up.fs_n; up.scales += (size_t)f0 * up.fs_n;
    up.xbeta += (size_t)f0 * up.fs_x;
    if (up.phi_f) up.phi_f += f0;
    if (J->dyn) return launch_status(pfl::launch_ung_psi_dyn(fp, up, J->geo, nf, J->stream), "psi filter (run-time m)");
    return launch_status(pfl::launch_ung_psim(J->m, fp, up, J->geo, nf, J->stream), "psi filter (m)");
  }
  if (J->method == BSSM_METHOD_PSI) {
    up.alphahat += (size_t)f0 * J->pstride; up.Lt += (size_t)f0 * J->pstride; up.Ct += (size_t)f0 * J->pstride;
    up.ytilde += (size_t)f0 * J->pstride; up.Htilde += (size_t)f0 * J->pstride; up.scales += (size_t)f0 * J->pstride;
    up.rHt += (size_t)f0 * J->pstride; up.yphi2 += (size_t)f0 * J->pstride;
    up.approx_loglik += f0;
    if (up.phi_f) up.phi_f += f0;
    return launch_status(pfl::launch_ung_psi1(fp, up, J->geo, nf, J->stream), "psi filter");
  }
  if (up.bsf_par) {
    up.bsf_par += (size_t)7 * f0;
    up.xbeta += (size_t)f0 * up.fs_x;
    return launch_status(pfl::launch_ung_bsf_draws(fp, up, J->geo, nf, J->stream), "bsf (per-draw models)");
  }
  if (J->dyn) {
    up.xbeta += (size_t)f0 * up.fs_x;
    if (up.phi_f) up.phi_f += f0;
    return launch_status(pfl::launch_ung_bsf_dyn(fp, up, J->geo, nf, J->stream), "bsf (run-time m)");
  }
  if (J->m <= 2) return launch_status(pfl::launch_ung_bsf_a(J->m, J->k, fp, up, J->geo, nf, J->stream), "bsf (m,k)");
  return launch_status(pfl::launch_ung_bsf_b(J->m, J->k, fp, up, J->geo, nf, J->stream), "bsf (m,k)");
}

} // namespace

extern "C" {

void bssm_b200_release_comm(void);
const char* bssm_b200_last_error(void) { return g_err.c_str(); }
const char* bssm_b200_version(void) { return "bssm_b200 0.1 (sm_100a)"; }

int bssm_b200_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); return 0; }
  return cnt;
}

int bssm_b200_set_device(int device) {
  int rc = require_device();
  if (rc) return rc;
  CU(cudaSetDevice(device));
  return 0;
}

const char* bssm_b200_last_gather(void);
void bssm_b200_shutdown(void) {
  bssm_b200_release_comm();
  pfl::user_models_release();
  g_pool.release_all();
  {
    std::lock_guard<std::mutex> g(g_sctx.mu);
    g_sctx.release();
  }
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) == cudaSuccess && cnt > 0) cudaDeviceSynchronize();
  cudaGetLastError();
}

// allocates the outputs / exchange buffers common to every model kind and fills FilterParams
static int finish_job(bssm_b200_job* J, int out_flags, int post) {
  int rc;
  const int n = J->n, m = J->m;
  const uint32_t nsim = J->N, n_filters = J->n_filters;
  // ---- outputs
  const bool smooth = post == BSSM_POST_SMOOTHER || post == BSSM_POST_SMOOTHER_MERGE;
  J->keep_alpha = (out_flags & BSSM_OUT_ALPHA) || post != BSSM_POST_NONE;
  J->keep_weights = (out_flags & BSSM_OUT_WEIGHTS) || post == BSSM_POST_FILTER_SUMMARY;
  J->keep_indices = (out_flags & BSSM_OUT_INDICES) || smooth || post == BSSM_POST_PICK;
  J->alpha_stride = (size_t)(J->keep_alpha ? n + 1 : 2) * m * nsim;
  J->weights_stride = J->keep_weights ? (size_t)nsim * (n + 1) : 0;
  J->indices_stride = J->keep_indices ? (size_t)nsim * n : (size_t)nsim;
  if ((rc = J->alpha_tm.alloc(J->alpha_stride * n_filters))) return rc;
  if (J->keep_weights && (rc = J->weights.alloc(J->weights_stride * n_filters))) return rc;
  if ((rc = J->indices.alloc(J->indices_stride * n_filters))) return rc;
  if (smooth && (rc = J->alpha_sm.alloc(J->alpha_stride * n_filters))) return rc;
  if (post == BSSM_POST_PICK && (rc = J->traj.alloc((size_t)n_filters * (n + 1) * m))) return rc;
  if ((rc = J->loglik.alloc(n_filters)) || (rc = J->stopped.alloc(n_filters))) return rc;
  if ((rc = J->xch.alloc((size_t)4 * J->geo.ranks * n_filters)) || (rc = J->counters.alloc((size_t)n_filters * pf::COUNTER_STRIDE)) ||
      (rc = J->ubuf.alloc((size_t)nsim * n_filters))) return rc;
  if (post == BSSM_POST_FILTER_SUMMARY) {
    if ((rc = J->s_at.alloc((size_t)n_filters * m * (n + 1))) || (rc = J->s_att.alloc((size_t)n_filters * m * n)) ||
        (rc = J->s_Pt.alloc((size_t)n_filters * m * m * (n + 1))) || (rc = J->s_Ptt.alloc((size_t)n_filters * m * m * n)))
      return rc;
  } else if (smooth) {
    if ((rc = J->s_mean.alloc((size_t)n_filters * m * (n + 1))) || (rc = J->s_cov.alloc((size_t)n_filters * m * m * (n + 1))))
      return rc;
  }
  pf::FilterParams& fp = J->fp;
  fp.N = nsim; fp.n = n; fp.seed = J->seed; fp.first_filter = J->first_filter;
  fp.ranks = J->geo.ranks; fp.sync_mode = J->geo.sync_mode;
  fp.keep_alpha = J->keep_alpha ? 1 : 0;
  fp.alpha_tm = J->alpha_tm.p; fp.alpha_stride = J->alpha_stride;
  fp.weights = J->weights.p; fp.weights_stride = J->weights_stride;
  fp.indices = J->indices.p; fp.indices_stride = J->indices_stride;
  fp.keep_indices = J->keep_indices ? 1 : 0;
  fp.loglik = J->loglik.p; fp.stopped_at = J->stopped.p;
  fp.xch = J->xch.p; fp.counters = J->counters.p; fp.ubuf = J->ubuf.p;
  fp.phase_cycles = nullptr;
  if (env_int("BSSM_B200_PHASES", 0)) {
    if ((rc = J->phases.alloc((size_t)6 * J->geo.ranks * n_filters))) return rc;
    CU(cudaMemset(J->phases.p, 0, sizeof(unsigned long long) * J->phases.count));
    fp.phase_cycles = J->phases.p;
  }
  return 0;
}

// m = 1 psi-APF whose filters share one model: Newton sweeps (approximate()) and the pre-pass on the host,
// one copy of the result for all filters of the job (pstride = 0)
static int host_prepass_upload(bssm_b200_job* J, const bssm_ung_model* M) {
  std::vector<double> mode, mode_lin;
  ung_approximate_m1(M, mode, mode_lin);
  PrepassHost ph;
  ung_prepass_m1_host(M, mode.data(), mode_lin.data(), ung_dist_const(M), ph);
  const size_t n1 = (size_t)J->n + 1, nf = J->n_filters;
  // one staging buffer, one copy: [alphahat | Lt | Ct | ytilde | Htilde | scales | rHt | yphi2 | approx_loglik x nf | phi x nf]
  std::vector<double> st(8 * n1 + 2 * nf);
  const std::vector<double>* parts[8] = {&ph.ah, &ph.Lt, &ph.Ct, &ph.yt, &ph.Ht, &ph.sc, &ph.rHt, &ph.yphi2};
  for (int a = 0; a < 8; ++a) std::copy(parts[a]->begin(), parts[a]->end(), st.begin() + a * n1);
  for (size_t f = 0; f < nf; ++f) { st[8 * n1 + f] = ph.approx_loglik; st[8 * n1 + nf + f] = M->phi; }
  int rc = J->alphahat.upload(st.data(), st.size());
  if (rc) return rc;
  double* b = J->alphahat.p;
  pf::UngParams& up = J->up;
  up.alphahat = b; up.Lt = b + n1; up.Ct = b + 2 * n1; up.ytilde = b + 3 * n1; up.Htilde = b + 4 * n1;
  up.scales = b + 5 * n1; up.rHt = b + 6 * n1; up.yphi2 = b + 7 * n1;
  up.approx_loglik = b + 8 * n1; up.phi_f = b + 8 * n1 + nf;
  up.pstride = 0;
  J->pstride = 0;
  J->host_prepass = true;
  return 0;
}

// draw_par (6 x n_filters: T, R, a1, P1, C, phi) and modes (n x n_filters) are
// HOST arrays; when given, every filter has its own model and mode
// (post_correct: update_model + approximate_for_is per draw)
static int create_ung_job(const bssm_ung_model* M, int method, uint32_t nsim, uint32_t seed,
                          uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post,
                          const double* draw_par, const double* draw_modes, bssm_b200_job** job,
                          const DrawModels* dm = nullptr) {
  if (!M || !job) return fail(BSSM_B200_ERR_ARG, "null argument");
  *job = nullptr;
  if (M->n < 1 || M->m < 1 || M->k < 1 || nsim < 2 || n_filters < 1)
    return fail(BSSM_B200_ERR_ARG, "need n >= 1, m >= 1, k >= 1, nsim >= 2, n_filters >= 1");
  if (M->distribution < 0 || M->distribution > 5) return fail(BSSM_B200_ERR_ARG, "distribution must be 0..5");
  const bool gaussian = M->distribution == BSSM_DIST_GAUSSIAN;
  if (gaussian && (!M->H || method != BSSM_METHOD_BSF))
    return fail(BSSM_B200_ERR_ARG, "linear-Gaussian models: bootstrap filter only, H required "
                                   "(their psi 'filter' is bssm_b200_gaussian_psi_smoother)");
  if (method != BSSM_METHOD_PSI && method != BSSM_METHOD_BSF) return fail(BSSM_B200_ERR_ARG, "method must be 1 (psi) or 2 (bsf)");
  if (method == BSSM_METHOD_BSF && draw_par && M->m != 1)
    return fail(BSSM_B200_ERR_UNSUPPORTED, "per-draw bootstrap filters: m = 1 in this build");
  if (M->m > 16) return fail(BSSM_B200_ERR_UNSUPPORTED, "state dimension m <= 16 in this build");
  const bool parset = dm && dm->parset;
  if (method == BSSM_METHOD_PSI && !parset && ((M->m != 1 && draw_par) || (M->m == 1 && dm && !draw_par)))
    return fail(BSSM_B200_ERR_UNSUPPORTED, "psi-APF with per-draw models: m = 1 (or bsm_ng through the DrawModels path)");
  if (parset && (draw_par || gaussian)) return fail(BSSM_B200_ERR_ARG, "parameter sets: non-Gaussian ssm_ung models, no draw_par");
  if (M->k > M->m) return fail(BSSM_B200_ERR_ARG, "k > m");
  if (post < 0 || post > BSSM_POST_SMOOTHER_MERGE) return fail(BSSM_B200_ERR_ARG, "bad post mode");
  int rc = require_device();
  if (rc) return rc;

  bssm_b200_job* J = new bssm_b200_job();
  struct Guard { bssm_b200_job* j; ~Guard() { delete j; } } guard{J};
  J->method = method; J->post = post; J->out_flags = out_flags;
  J->N = nsim; J->n_filters = n_filters; J->first_filter = first_filter_id; J->seed = seed;
  J->n = M->n; J->m = M->m; J->k = M->k;
  const int n = M->n, m = M->m, k = M->k;
  {
    const int mk = M->m * 10 + M->k;
    const bool bsf_inst = mk == 11 || mk == 21 || mk == 22 || mk == 32 || mk == 33;
    // per-filter models of the bootstrap filter are read by the run-time-m policy only
    J->dyn = M->m > 3 || (method == BSSM_METHOD_BSF && (!bsf_inst || parset));
  }
  rc = choose_geometry(nsim, n_filters, &J->geo, J->dyn);
  if (rc) return rc;
  CU(cudaStreamCreateWithFlags(&J->stream, cudaStreamNonBlocking));
  for (auto& e : J->ev) CU(cudaEventCreate(&e));

  // ---- model upload (column-major, as R holds it)
  std::vector<double> xb(n, 0.0);
  if (M->xbeta) xb.assign(M->xbeta, M->xbeta + n);
  std::vector<double> LP1;
  if (gaussian) { LP1.resize((size_t)m * m); psd_chol_host(M->P1, m, LP1.data()); } // psd_chol, src/model_ssm_ulg.cpp:560
  else chol_posdiag(M->P1, m, LP1);
  std::vector<double> ones(n, 1.0);
  if ((rc = J->y.upload(M->y, n))) return rc;
  if ((rc = J->Z.upload(M->Z, (size_t)m * (M->Ztv ? n : 1)))) return rc;
  if ((rc = J->T.upload(M->T, (size_t)m * m * (M->Ttv ? n : 1)))) return rc;
  if ((rc = J->R.upload(M->R, (size_t)m * k * (M->Rtv ? n : 1)))) return rc;
  if ((rc = J->a1.upload(M->a1, m))) return rc;
  if ((rc = J->P1.upload(M->P1, (size_t)m * m))) return rc;
  if ((rc = J->LP1.upload(LP1.data(), (size_t)m * m))) return rc;
  if ((rc = J->D.upload(M->D, M->Dtv ? n : 1))) return rc;
  if ((rc = J->C.upload(M->C, (size_t)m * (M->Ctv ? n : 1)))) return rc;
  if ((rc = J->xbeta.upload(xb.data(), n))) return rc;
  if ((rc = J->u.upload(M->u ? M->u : ones.data(), n))) return rc;

  pf::UngParams& up = J->up;
  up.n = n; up.distribution = M->distribution;
  up.m_dyn = m; up.k_dyn = k;
  up.Ztv = M->Ztv; up.Ttv = M->Ttv; up.Rtv = M->Rtv; up.Dtv = M->Dtv; up.Ctv = M->Ctv;
  up.y = J->y.p; up.Z = J->Z.p; up.T = J->T.p; up.R = J->R.p; up.a1 = J->a1.p; up.LP1 = J->LP1.p;
  up.D = J->D.p; up.C = J->C.p; up.xbeta = J->xbeta.p; up.u = J->u.p;
  up.phi = M->phi;
  const double dconst = gaussian ? 0.0 : ung_dist_const(M);
  up.loglik_const = dconst;
  up.HH = nullptr; up.hconst = nullptr;
  if (gaussian) { // HH = H^2 (compute_HH) and norm_log_const(H_t) per time point (src/distr_consts.cpp:12-14)
    std::vector<double> HH(n), hc(n);
    for (int t = 0; t < n; ++t) {
      const double h = M->H[t * M->Htv];
      HH[t] = h * h;
      hc[t] = -0.5 * std::log(2.0 * M_PI) - std::log(h);
    }
    if ((rc = J->HH_v.upload(HH.data(), n)) || (rc = J->hconst_v.upload(hc.data(), n))) return rc;
    up.HH = J->HH_v.p; up.hconst = J->hconst_v.p;
  }

  if (method == BSSM_METHOD_BSF && draw_par) {
    // is_correction_bsf: one (T, R, a1, P1, C, phi) per filter; sqrt(P1) and the state-independent
    // constants of the log-likelihood (they depend on phi) per filter on the host
    if (m != 1 || gaussian) return fail(BSSM_B200_ERR_UNSUPPORTED, "per-draw bootstrap filters: m = 1, non-Gaussian observations");
    std::vector<double> bp((size_t)7 * n_filters);
    for (uint32_t f = 0; f < n_filters; ++f) {
      const double* q = draw_par + (size_t)6 * f;
      bssm_ung_model Mi = *M;
      Mi.phi = q[5];
      double* b = &bp[(size_t)7 * f];
      b[0] = q[0]; b[1] = q[1]; b[2] = q[2];
      b[3] = q[3] > 0.0 ? std::sqrt(q[3]) : 0.0; // chol of the positive-diagonal block, src/model_ssm_ung.cpp:616-618
      b[4] = q[4]; b[5] = q[5]; b[6] = ung_dist_const(&Mi);
    }
    if ((rc = J->draw_par.upload(bp.data(), bp.size()))) return rc;
    up.bsf_par = J->draw_par.p;
    if (dm && !dm->xbeta.empty()) { // beta is part of theta: one regression part per filter
      if (dm->xbeta.size() != (size_t)n * n_filters) return fail(BSSM_B200_ERR_ARG, "per-draw xbeta: n x n_filters");
      if ((rc = J->xbeta.upload(dm->xbeta.data(), dm->xbeta.size()))) return rc;
      up.xbeta = J->xbeta.p; up.fs_x = (size_t)n;
    }
  }

  // parameter sets: the per-filter model arrays the kernels read (Z, D for the psi-APF; all for the bootstrap filter)
  auto upload_slices = [&](const std::vector<double>& v, DevBuf<double>& buf, size_t slice, const double** ptr,
                           size_t* stride) -> int {
    if (v.empty()) return 0;
    if (v.size() != slice * n_filters) return fail(BSSM_B200_ERR_ARG, "parameter set: slice size x n_filters expected");
    int r = buf.upload(v.data(), v.size());
    if (r) return r;
    *ptr = buf.p; *stride = slice;
    return 0;
  };
  if (parset) {
    if ((rc = upload_slices(dm->Z, J->Z, (size_t)m * (M->Ztv ? n : 1), &up.Z, &up.fs_Z)) ||
        (rc = upload_slices(dm->D, J->D, (size_t)(M->Dtv ? n : 1), &up.D, &up.fs_D)))
      return rc;
  }
  if (method == BSSM_METHOD_BSF && parset) {
    // is_correction_bsf<ssm_ung> (src/approx_mcmc.cpp:806-921): every filter runs its own model
    if ((rc = upload_slices(dm->T, J->T, (size_t)m * m * (M->Ttv ? n : 1), &up.T, &up.fs_T)) ||
        (rc = upload_slices(dm->R, J->R, (size_t)m * k * (M->Rtv ? n : 1), &up.R, &up.fs_R)) ||
        (rc = upload_slices(dm->C, J->C, (size_t)m * (M->Ctv ? n : 1), &up.C, &up.fs_C)) ||
        (rc = upload_slices(dm->a1, J->a1, (size_t)m, &up.a1, &up.fs_a1)) ||
        (rc = upload_slices(dm->xbeta, J->xbeta, (size_t)n, &up.xbeta, &up.fs_x)))
      return rc;
    if (!dm->P1.empty()) { // chol of each draw's P1 (positive-diagonal block), src/model_ssm_ung.cpp:616-618
      std::vector<double> L((size_t)m * m * n_filters), Lf;
      for (uint32_t f = 0; f < n_filters; ++f) {
        chol_posdiag(&dm->P1[(size_t)m * m * f], m, Lf);
        std::copy(Lf.begin(), Lf.end(), &L[(size_t)m * m * f]);
      }
      if ((rc = upload_slices(L, J->LP1, (size_t)m * m, &up.LP1, &up.fs_LP1))) return rc;
    }
    if (!dm->phi.empty()) { // phi enters the state-independent constants of the log-likelihood
      std::vector<double> lc(n_filters);
      for (uint32_t f = 0; f < n_filters; ++f) {
        bssm_ung_model Mi = *M;
        Mi.phi = dm->phi[f];
        lc[f] = ung_dist_const(&Mi);
      }
      if ((rc = J->phi_f.upload(dm->phi.data(), n_filters)) || (rc = J->lconst_f.upload(lc.data(), n_filters))) return rc;
      up.phi_f = J->phi_f.p; up.lconst_f = J->lconst_f.p;
    }
  }

  if (method == BSSM_METHOD_PSI && dm && (m > 1 || parset)) {
    // approximate_for_is (src/model_ssm_ung.cpp:143-161) per draw on the host, OpenMP over draws:
    // laplace_iter at the stored mode, scales, Kalman smoother + conditional covariances
    if (M->Rtv && !parset) return fail(BSSM_B200_ERR_UNSUPPORTED, "per-draw models: time-invariant R only");
    if (dm->modes.size() != (size_t)n * n_filters) return fail(BSSM_B200_ERR_ARG, "per-draw models need per-draw modes");
    const size_t n1 = (size_t)n + 1, mm = (size_t)m * m, nf = n_filters;
    const bool has_xb = !dm->xbeta.empty();
    if (!env_int("BSSM_B200_HOST_APPROX", 0)) {
      // on the device, one thread per draw (approx_device.cuh): the draws are the parallelism.  The psi-APF reads
      // only the approximation (and Z, D, phi, xbeta), so the job's T / R / a1 / P1 / C buffers can carry the
      // per-draw arrays the approximation needs.
      pf::ApproxParams ap;
      memset(&ap, 0, sizeof(ap));
      ap.n = n; ap.m = m; ap.k = k; ap.distribution = M->distribution;
      ap.Ztv = M->Ztv; ap.Ttv = M->Ttv; ap.Rtv = M->Rtv; ap.Dtv = M->Dtv; ap.Ctv = M->Ctv;
      ap.y = J->y.p; ap.u = J->u.p;
      ap.Z = up.Z; ap.fs_Z = up.fs_Z; ap.D = up.D; ap.fs_D = up.fs_D; // shared, or per draw (parameter sets, above)
      ap.T = J->T.p; ap.R = J->R.p; ap.a1 = J->a1.p; ap.P1 = J->P1.p; ap.C = J->C.p;
      if ((rc = upload_slices(dm->T, J->T, (size_t)m * m * (M->Ttv ? n : 1), &ap.T, &ap.fs_T)) ||
          (rc = upload_slices(dm->R, J->R, (size_t)m * k * (M->Rtv ? n : 1), &ap.R, &ap.fs_R)) ||
          (rc = upload_slices(dm->C, J->C, (size_t)m * (M->Ctv ? n : 1), &ap.C, &ap.fs_C)) ||
          (rc = upload_slices(dm->a1, J->a1, (size_t)m, &ap.a1, &ap.fs_a1)) ||
          (rc = upload_slices(dm->P1, J->P1, mm, &ap.P1, &ap.fs_P1)))
        return rc;
      ap.has_xbeta = (has_xb || M->xbeta) ? 1 : 0;
      if (has_xb) {
        if ((rc = J->xbeta.upload(dm->xbeta.data(), dm->xbeta.size()))) return rc;
        up.xbeta = J->xbeta.p;
      }
      ap.xbeta = J->xbeta.p; ap.fs_x = has_xb ? (size_t)n : 0;
      ap.phi = M->phi;
      if (!dm->phi.empty()) {
        if ((rc = J->phi_f.upload(dm->phi.data(), nf))) return rc;
        ap.phi_f = J->phi_f.p;
      }
      if ((rc = J->mode.upload(dm->modes.data(), dm->modes.size()))) return rc;
      ap.modes = J->mode.p;
      if ((rc = J->alphahat.alloc((size_t)m * n1 * nf)) || (rc = J->Lt.alloc(mm * n1 * nf)) || (rc = J->Ct.alloc(mm * n1 * nf)) ||
          (rc = J->ytilde.alloc((size_t)n * nf)) || (rc = J->Htilde.alloc((size_t)n * nf)) || (rc = J->scales.alloc((size_t)n * nf)) ||
          (rc = J->approx_loglik.alloc(nf)) || (rc = J->scratch.alloc((size_t)(m + 3) * n * nf)))
        return rc;
      CU(cudaMemsetAsync(J->approx_loglik.p, 0, sizeof(double) * nf, J->stream)); // approximate_for_is: approx_loglik = 0 (:159)
      ap.alphahat = J->alphahat.p; ap.Lt = J->Lt.p; ap.Ct = J->Ct.p; ap.ytilde = J->ytilde.p; ap.Htilde = J->Htilde.p;
      ap.scales = J->scales.p; ap.scratch = J->scratch.p;
      ap.fs_a = (size_t)m * n1; ap.fs_m = mm * n1; ap.fs_n = (size_t)n;
      CU(cudaEventRecord(J->ev[0], J->stream));
      const uint32_t ab = cdiv((uint32_t)nf, 32);
      if (m == 1) pf::ung_approx_for_is_kernel<1, true><<<ab, 32, 0, J->stream>>>(ap, (uint32_t)nf);
      else if (m == 2) pf::ung_approx_for_is_kernel<2, true><<<ab, 32, 0, J->stream>>>(ap, (uint32_t)nf);
      else if (m == 3) pf::ung_approx_for_is_kernel<3, true><<<ab, 32, 0, J->stream>>>(ap, (uint32_t)nf);
      else pf::ung_approx_for_is_kernel<16, false><<<ab, 32, 0, J->stream>>>(ap, (uint32_t)nf);
      CU(cudaGetLastError());
      CU(cudaEventRecord(J->ev[1], J->stream));
      CU(cudaStreamSynchronize(J->stream));
      if (env_int("BSSM_B200_DEBUG_APPROX", 0)) {
        float ams = 0;
        cudaEventElapsedTime(&ams, J->ev[0], J->ev[1]);
        fprintf(stderr, "[approx] %zu draws, m = %d, n = %d: %.3f ms on the device\n", nf, m, n, ams);
      }
      J->pstride = 0;
      J->host_approx_per_filter = true; // = the filters carry ready-made approximations
      up.phi_f = dm->phi.empty() ? nullptr : J->phi_f.p; up.alphahat = J->alphahat.p; up.Lt = J->Lt.p; up.Ct = J->Ct.p;
      up.ytilde = J->ytilde.p; up.Htilde = J->Htilde.p; up.scales = J->scales.p;
      up.approx_loglik = J->approx_loglik.p; up.pstride = 0;
      up.fs_a = (size_t)m * n1; up.fs_m = mm * n1; up.fs_n = (size_t)n; up.fs_x = has_xb ? (size_t)n : 0;
      if ((rc = finish_job(J, out_flags, post))) return rc;
      guard.j = nullptr;
      *job = J;
      return 0;
    }
    std::vector<double> ah((size_t)m * n1 * nf), Lt(mm * n1 * nf), Ct(mm * n1 * nf), yt((size_t)n * nf),
        Ht((size_t)n * nf), sc((size_t)n * nf), al(nf, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < (int64_t)nf; ++f) {
      bssm_ung_model Mi;
      draw_model(M, *dm, (size_t)f, Mi);
      pfga::UngApprox ga(&Mi);
      ga.run(&dm->modes[(size_t)n * f]);
      std::copy(ga.alphahat.begin(), ga.alphahat.end(), &ah[(size_t)m * n1 * f]);
      std::copy(ga.L.begin(), ga.L.end(), &Lt[mm * n1 * f]);
      std::copy(ga.Ct.begin(), ga.Ct.end(), &Ct[mm * n1 * f]);
      std::copy(ga.ay.begin(), ga.ay.end(), &yt[(size_t)n * f]);
      std::copy(ga.aH.begin(), ga.aH.end(), &Ht[(size_t)n * f]);
      std::copy(ga.scales.begin(), ga.scales.end(), &sc[(size_t)n * f]);
    }
    if ((rc = J->alphahat.upload(ah.data(), ah.size())) || (rc = J->Lt.upload(Lt.data(), Lt.size())) ||
        (rc = J->Ct.upload(Ct.data(), Ct.size())) || (rc = J->ytilde.upload(yt.data(), yt.size())) ||
        (rc = J->Htilde.upload(Ht.data(), Ht.size())) || (rc = J->scales.upload(sc.data(), sc.size())) ||
        (rc = J->approx_loglik.upload(al.data(), nf)))
      return rc;
    if (!dm->phi.empty() && (rc = J->phi_f.upload(dm->phi.data(), nf))) return rc;
    if (has_xb) {
      if ((rc = J->xbeta.upload(dm->xbeta.data(), dm->xbeta.size()))) return rc;
      up.xbeta = J->xbeta.p;
    }
    J->pstride = 0;
    J->host_approx_per_filter = true;
    up.phi_f = dm->phi.empty() ? nullptr : J->phi_f.p; up.alphahat = J->alphahat.p; up.Lt = J->Lt.p; up.Ct = J->Ct.p;
    up.ytilde = J->ytilde.p; up.Htilde = J->Htilde.p; up.scales = J->scales.p;
    up.approx_loglik = J->approx_loglik.p; up.pstride = 0;
    up.fs_a = (size_t)m * n1; up.fs_m = mm * n1; up.fs_n = (size_t)n; up.fs_x = has_xb ? (size_t)n : 0;
  } else if (method == BSSM_METHOD_PSI && m > 1) {
    // Gaussian approximation on the host (O(n m^3), sequential in t), shared by all replicates
    if (!M->initial_mode) return fail(BSSM_B200_ERR_ARG, "psi-APF needs initial_mode");
    if (M->max_iter < 1) return fail(BSSM_B200_ERR_ARG, "max_iter must be >= 1");
    pfga::UngApprox ga(M);
    ga.run(nullptr);
    std::vector<double> al(n_filters, ga.approx_loglik);
    if ((rc = J->alphahat.upload(ga.alphahat.data(), ga.alphahat.size())) || (rc = J->Lt.upload(ga.L.data(), ga.L.size())) ||
        (rc = J->Ct.upload(ga.Ct.data(), ga.Ct.size())) || (rc = J->ytilde.upload(ga.ay.data(), n)) ||
        (rc = J->Htilde.upload(ga.aH.data(), n)) || (rc = J->scales.upload(ga.scales.data(), n)) ||
        (rc = J->approx_loglik.upload(al.data(), n_filters)))
      return rc;
    J->pstride = 0;
    up.phi_f = nullptr; up.alphahat = J->alphahat.p; up.Lt = J->Lt.p; up.Ct = J->Ct.p;
    up.ytilde = J->ytilde.p; up.Htilde = J->Htilde.p; up.scales = J->scales.p;
    up.approx_loglik = J->approx_loglik.p; up.pstride = 0;
  } else if (method == BSSM_METHOD_PSI) {
    if (draw_par) {
      // approximate_for_is (src/model_ssm_ung.cpp:143-161): fixed mode per draw, no iterations
      if (!draw_modes) return fail(BSSM_B200_ERR_ARG, "per-draw models need per-draw modes");
      if ((rc = J->mode.upload(draw_modes, (size_t)n * n_filters)) ||
          (rc = J->draw_par.upload(draw_par, (size_t)6 * n_filters)))
        return rc;
    } else {
      if (!M->initial_mode) return fail(BSSM_B200_ERR_ARG, "psi-APF needs initial_mode");
      if (M->max_iter < 1) return fail(BSSM_B200_ERR_ARG, "max_iter must be >= 1");
      if (!env_int("BSSM_B200_DEVICE_PREPASS", 0)) {
        // the filters of the job share one model: the pre-pass is ONE sequential chain, the host's business
        if ((rc = host_prepass_upload(J, M))) return rc;
        if ((rc = finish_job(J, out_flags, post))) return rc;
        guard.j = nullptr;
        *job = J;
        return 0;
      }
      std::vector<double> mode, mode_lin;
      ung_approximate_m1(M, mode, mode_lin);
      if ((rc = J->mode.upload(mode.data(), n)) || (rc = J->mode_lin.upload(mode_lin.data(), n))) return rc;
    }
    J->pstride = (size_t)n + 1;
    const size_t tot = J->pstride * n_filters;
    if ((rc = J->alphahat.alloc(tot)) || (rc = J->Lt.alloc(tot)) || (rc = J->Ct.alloc(tot)) ||
        (rc = J->ytilde.alloc(tot)) || (rc = J->Htilde.alloc(tot)) || (rc = J->scales.alloc(tot)) ||
        (rc = J->scratch.alloc(3 * tot)) || (rc = J->approx_loglik.alloc(n_filters)) ||
        (rc = J->phi_f.alloc(n_filters)) || (rc = J->rHt.alloc(tot)) || (rc = J->yphi2.alloc(tot)))
      return rc;
    pf::PrepassParams& pp = J->pp;
    pp.n = n; pp.distribution = M->distribution;
    pp.Ztv = M->Ztv; pp.Ttv = M->Ttv; pp.Rtv = M->Rtv; pp.Dtv = M->Dtv; pp.Ctv = M->Ctv;
    pp.y = J->y.p; pp.Z = J->Z.p; pp.T = J->T.p; pp.R = J->R.p; pp.a1 = J->a1.p; pp.P1 = J->P1.p;
    pp.D = J->D.p; pp.C = J->C.p; pp.xbeta = J->xbeta.p; pp.u = J->u.p;
    pp.has_xbeta = M->xbeta ? 1 : 0;
    pp.phi = M->phi;
    if (draw_par) {
      pp.draw_par = J->draw_par.p;
      pp.modes = J->mode.p; pp.modes_lin = J->mode.p; pp.mode_stride = (size_t)n;
      pp.compute_loglik = 0; // approx_loglik = 0 (:159)
    } else {
      pp.draw_par = nullptr;
      pp.modes = J->mode.p; pp.modes_lin = J->mode_lin.p; pp.mode_stride = 0; // replicates share the mode
      pp.compute_loglik = 1;
    }
    pp.dist_const = dconst;
    pp.log2pi = std::log(2.0 * M_PI);
    pp.alphahat = J->alphahat.p; pp.Lt = J->Lt.p; pp.Ct = J->Ct.p; pp.ytilde = J->ytilde.p;
    pp.Htilde = J->Htilde.p; pp.scales = J->scales.p; pp.approx_loglik = J->approx_loglik.p;
    pp.phi_f = J->phi_f.p; pp.scratch = J->scratch.p; pp.pstride = J->pstride;
    pp.rHt = J->rHt.p; pp.yphi2 = J->yphi2.p;
    up.rHt = J->rHt.p; up.yphi2 = J->yphi2.p;
    up.phi_f = J->phi_f.p; up.alphahat = J->alphahat.p; up.Lt = J->Lt.p; up.Ct = J->Ct.p;
    up.ytilde = J->ytilde.p; up.Htilde = J->Htilde.p; up.scales = J->scales.p;
    up.approx_loglik = J->approx_loglik.p; up.pstride = J->pstride;
    if (draw_par && dm && !dm->xbeta.empty()) {
      // beta is part of theta (ar1_ng / bsm_ng update_model -> compute_xbeta): one regression part per
      // filter; the pre-pass reads its own column, the filter runs the per-filter-pointer policy at m = 1
      if (dm->xbeta.size() != (size_t)n * n_filters) return fail(BSSM_B200_ERR_ARG, "per-draw xbeta: n x n_filters");
      if ((rc = J->xbeta.upload(dm->xbeta.data(), dm->xbeta.size()))) return rc;
      pp.xbeta = J->xbeta.p; pp.xbeta_stride = (size_t)n; pp.has_xbeta = 1;
      up.xbeta = J->xbeta.p; up.fs_x = (size_t)n;
      up.fs_a = up.fs_m = up.fs_n = J->pstride;
      J->xbeta_per_filter = true;
    }
  }

  if ((rc = finish_job(J, out_flags, post))) return rc;
  guard.j = nullptr;
  *job = J;
  return 0;
}

int bssm_b200_job_create_ung(const bssm_ung_model* M, int method, uint32_t nsim, uint32_t seed,
                             uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post,
                             bssm_b200_job** job) {
  if (post > BSSM_POST_SMOOTHER) return fail(BSSM_B200_ERR_ARG, "post must be 0, 1 or 2");
  return create_ung_job(M, method, nsim, seed, n_filters, first_filter_id, out_flags, post, nullptr, nullptr, job);
}

// enqueues the pre-pass and the filter kernel(s) on the job's stream (events 0..2)
static int job_launch_filters(bssm_b200_job* J) {
  int rc;
  J->launches = 0;
  CU(cudaEventRecord(J->ev[0], J->stream));
  if (J->kind == KIND_UNG && J->method == BSSM_METHOD_PSI && J->m == 1 && !J->host_approx_per_filter && !J->host_prepass) {
    const uint32_t warps_per_block = 4;
    pf::ung_prepass_m1_kernel<<<cdiv(J->n_filters, warps_per_block), 32 * warps_per_block, 0, J->stream>>>(
        J->pp, J->n_filters);
    CU(cudaGetLastError());
    J->launches += 1;
  }
  CU(cudaEventRecord(J->ev[1], J->stream));
  if (J->geo.sync_mode == pf::SYNC_GROUP) {
    if (J->wave == 0) {
      J->wave = bssm_b200_job_max_concurrent_filters(J);
      if (J->wave < 1) return fail(BSSM_B200_ERR_UNSUPPORTED, "filter group does not fit on the device");
    }
    CU(cudaMemsetAsync(J->counters.p, 0, sizeof(unsigned int) * J->counters.count, J->stream));
    for (uint32_t f = 0; f < J->n_filters; f += (uint32_t)J->wave) {
      const uint32_t nf = std::min<uint32_t>((uint32_t)J->wave, J->n_filters - f);
      if ((rc = launch_ung_filter(J, f, nf))) return rc;
      J->launches += 1;
    }
  } else {
    if ((rc = launch_ung_filter(J, 0, J->n_filters))) return rc;
    J->launches += 1;
  }
  CU(cudaEventRecord(J->ev[2], J->stream));
  return 0;
}

// post-processing kernels, synchronisation, timings
static int job_finish(bssm_b200_job* J, float* ms_out) {
  int rc;
  if (J->post != BSSM_POST_NONE) {
    // stopped filters: zero the columns they never reached before anything reads them
    pf::zero_unreached_kernel<<<dim3(64, J->n_filters), 256, 0, J->stream>>>(
        J->stopped.p, J->alpha_tm.p, J->weights.p, J->keep_indices ? J->indices.p : nullptr, J->N, J->n, J->m,
        J->keep_alpha ? 1 : 0, J->alpha_stride, J->weights_stride, J->indices_stride);
    CU(cudaGetLastError());
    J->launches += 1;
    if ((rc = run_post(J))) return rc;
  }
  CU(cudaEventRecord(J->ev[3], J->stream));
  CU(cudaStreamSynchronize(J->stream));
  CU(cudaEventElapsedTime(&J->ms[0], J->ev[0], J->ev[1]));
  CU(cudaEventElapsedTime(&J->ms[1], J->ev[1], J->ev[2]));
  CU(cudaEventElapsedTime(&J->ms[2], J->ev[2], J->ev[3]));
  float tot = 0;
  CU(cudaEventElapsedTime(&tot, J->ev[0], J->ev[3]));
  if (ms_out) *ms_out = tot;
  J->has_run = true;
  return 0;
}

int bssm_b200_job_run(bssm_b200_job* J, float* ms_out) {
  if (!J) return fail(BSSM_B200_ERR_ARG, "null job");
  int rc = job_launch_filters(J);
  return rc ? rc : job_finish(J, ms_out);
}

int bssm_b200_job_kernel_ms(const bssm_b200_job* J, float ms[3]) {
  if (!J || !ms) return fail(BSSM_B200_ERR_ARG, "null argument");
  ms[0] = J->ms[0]; ms[1] = J->ms[1]; ms[2] = J->ms[2];
  return 0;
}

int bssm_b200_job_launches(const bssm_b200_job* J) { return J ? J->launches : 0; }

int bssm_b200_job_phase_cycles(bssm_b200_job* J, uint32_t cta, uint64_t cycles[6]) {
  if (!J || !cycles) return fail(BSSM_B200_ERR_ARG, "null argument");
  for (int i = 0; i < 6; ++i) cycles[i] = 0;
  if (!J->phases.p) return 0;
  if ((size_t)cta * 6 + 6 > J->phases.count) return fail(BSSM_B200_ERR_ARG, "cta out of range");
  CU(cudaMemcpy(cycles, J->phases.p + (size_t)cta * 6, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return 0;
}

int bssm_b200_job_max_concurrent_filters(const bssm_b200_job* J) {
  if (!J) return 0;
  int out = 0, rc;
  if (J->kind == KIND_NLG && J->nlg_model_id >= BSSM_USER_MODEL_BASE)
    rc = pfl::occupancy_user_nlg(J->nlg_model_id, J->m, J->k, J->geo, J->np.theta_u_stride != 0, &out);
  else if (J->kind == KIND_NLG && J->method == BSSM_METHOD_PSI) rc = pfl::occupancy_nlg_psi(J->geo, &out);
  else if (J->kind == KIND_NLG) rc = pfl::occupancy_nlg_bsf(J->geo, &out);
  else if (J->kind == KIND_SDE && J->sde_model_id >= BSSM_USER_MODEL_BASE)
    rc = pfl::occupancy_user_sde(J->sde_model_id, J->geo, J->sp.theta_u_stride != 0, &out);
  else if (J->kind == KIND_SDE) rc = pfl::occupancy_sde_bsf(J->geo, &out);
  else if (J->kind == KIND_MNG && J->dyn) rc = pfl::occupancy_mng_dyn(J->geo, &out);
  else if (J->kind == KIND_MNG) rc = pfl::occupancy_mng_bsf(J->geo, &out);
  else if (J->dyn) rc = pfl::occupancy_ung_dyn(J->geo, &out);
  else if (J->method == BSSM_METHOD_PSI) rc = pfl::occupancy_ung_psi1(J->geo, &out);
  else rc = pfl::occupancy_ung_bsf(J->geo, &out);
  return rc ? 0 : out;
}

int bssm_b200_job_logliks(bssm_b200_job* J, double* loglik) {
  if (!J || !loglik) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (!J->has_run) return fail(BSSM_B200_ERR_ARG, "job has not run");
  CU(cudaMemcpy(loglik, J->loglik.p, sizeof(double) * J->n_filters, cudaMemcpyDeviceToHost));
  return 0;
}

int bssm_b200_job_logliks_dev(bssm_b200_job* J, double* loglik_dev) {
  if (!J || !loglik_dev) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (!J->has_run) return fail(BSSM_B200_ERR_ARG, "job has not run");
  CU(cudaMemcpy(loglik_dev, J->loglik.p, sizeof(double) * J->n_filters, cudaMemcpyDeviceToDevice));
  return 0;
}

int bssm_b200_job_fetch(bssm_b200_job* J, uint32_t f, bssm_filter_out* out) {
  if (!J || !out) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (!J->has_run) return fail(BSSM_B200_ERR_ARG, "job has not run");
  if (f >= J->n_filters) return fail(BSSM_B200_ERR_ARG, "filter index out of range");
  const int n = J->n, m = J->m;
  const uint32_t N = J->N;
  int32_t stopped = -1;
  CU(cudaMemcpy(&out->loglik, J->loglik.p + f, sizeof(double), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&stopped, J->stopped.p + f, sizeof(int32_t), cudaMemcpyDeviceToHost));
  out->stopped_at = stopped;
  out->approx_loglik = 0.0;
  if (J->method == BSSM_METHOD_PSI && J->host_prepass) // packed behind the pre-pass arrays (host_prepass_upload)
    CU(cudaMemcpy(&out->approx_loglik, J->up.approx_loglik + f, sizeof(double), cudaMemcpyDeviceToHost));
  else if (J->method == BSSM_METHOD_PSI && J->approx_loglik.p)
    CU(cudaMemcpy(&out->approx_loglik, J->approx_loglik.p + f, sizeof(double), cudaMemcpyDeviceToHost));
  out->device_ms = J->ms[0] + J->ms[1] + J->ms[2];
  // columns the filter never reached stay zero, as in the reference (arma::fill::zeros)
  const int cols_alpha = stopped >= 0 ? stopped + 1 : n + 1; // state columns 0..stopped exist
  const int cols_w = stopped >= 0 ? stopped : n + 1;
  const int cols_i = stopped >= 0 ? stopped : n;
  if (out->alpha) {
    if (!J->keep_alpha) return fail(BSSM_B200_ERR_ARG, "job was created without BSSM_OUT_ALPHA");
    const double* src = (J->alpha_sm.p ? J->alpha_sm.p : J->alpha_tm.p) + (size_t)f * J->alpha_stride;
    const uint32_t rows = (uint32_t)(n + 1) * m;
    int rc = J->cube.alloc((size_t)rows * N);
    if (rc) return rc;
    if (stopped >= 0) // zero the unreached tail in the time-major source first
      CU(cudaMemsetAsync((void*)(src + (size_t)cols_alpha * m * N), 0, sizeof(double) * (size_t)(n + 1 - cols_alpha) * m * N, J->stream));
    dim3 grid(cdiv(N, 32), cdiv(rows, 32));
    pf::transpose_to_cube_kernel<<<grid, dim3(32, 8), 0, J->stream>>>(src, J->cube.p, rows, N, 0u, rows);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out->alpha, J->cube.p, sizeof(double) * (size_t)rows * N, cudaMemcpyDeviceToHost, J->stream));
    CU(cudaStreamSynchronize(J->stream));
  }
  if (out->weights) {
    if (!J->keep_weights) return fail(BSSM_B200_ERR_ARG, "job was created without BSSM_OUT_WEIGHTS");
    if (stopped >= 0) memset(out->weights + (size_t)N * cols_w, 0, sizeof(double) * (size_t)N * (n + 1 - cols_w));
    CU(cudaMemcpy(out->weights, J->weights.p + (size_t)f * J->weights_stride, sizeof(double) * (size_t)N * cols_w, cudaMemcpyDeviceToHost));
  }
  if (out->indices) {
    if (!J->keep_indices) return fail(BSSM_B200_ERR_ARG, "job was created without BSSM_OUT_INDICES");
    if (stopped >= 0) memset(out->indices + (size_t)N * cols_i, 0, sizeof(uint32_t) * (size_t)N * (n - cols_i));
    CU(cudaMemcpy(out->indices, J->indices.p + (size_t)f * J->indices_stride, sizeof(uint32_t) * (size_t)N * cols_i, cudaMemcpyDeviceToHost));
  }
  if (J->post == BSSM_POST_FILTER_SUMMARY) {
    if (out->at) CU(cudaMemcpy(out->at, J->s_at.p + (size_t)f * m * (n + 1), sizeof(double) * m * (n + 1), cudaMemcpyDeviceToHost));
    if (out->att) CU(cudaMemcpy(out->att, J->s_att.p + (size_t)f * m * n, sizeof(double) * m * n, cudaMemcpyDeviceToHost));
    if (out->Pt) CU(cudaMemcpy(out->Pt, J->s_Pt.p + (size_t)f * m * m * (n + 1), sizeof(double) * m * m * (n + 1), cudaMemcpyDeviceToHost));
    if (out->Ptt) CU(cudaMemcpy(out->Ptt, J->s_Ptt.p + (size_t)f * m * m * n, sizeof(double) * m * m * n, cudaMemcpyDeviceToHost));
  } else if (J->post == BSSM_POST_SMOOTHER || J->post == BSSM_POST_SMOOTHER_MERGE) {
    if (out->alphahat) CU(cudaMemcpy(out->alphahat, J->s_mean.p + (size_t)f * m * (n + 1), sizeof(double) * m * (n + 1), cudaMemcpyDeviceToHost));
    if (out->Vt) CU(cudaMemcpy(out->Vt, J->s_cov.p + (size_t)f * m * m * (n + 1), sizeof(double) * m * m * (n + 1), cudaMemcpyDeviceToHost));
  }
  return 0;
}

void bssm_b200_job_destroy(bssm_b200_job* J) { delete J; }

// ------------------------------------------------------------ ssm_nlg / ssm_sde
namespace {

int job_common(bssm_b200_job* J, int kind, int n, int m, int k, uint32_t nsim, uint32_t seed,
               uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post) {
  int rc = require_device();
  if (rc) return rc;
  if (n < 1 || nsim < 2 || n_filters < 1) return fail(BSSM_B200_ERR_ARG, "need n >= 1, nsim >= 2, n_filters >= 1");
  if (post < 0 || post > BSSM_POST_SMOOTHER) return fail(BSSM_B200_ERR_ARG, "post must be 0, 1 or 2");
  J->kind = kind; J->method = BSSM_METHOD_BSF; J->post = post; J->out_flags = out_flags;
  J->N = nsim; J->n_filters = n_filters; J->first_filter = first_filter_id; J->seed = seed;
  J->n = n; J->m = m; J->k = k;
  if ((rc = choose_geometry(nsim, n_filters, &J->geo))) return rc;
  CU(cudaStreamCreateWithFlags(&J->stream, cudaStreamNonBlocking));
  for (auto& e : J->ev) CU(cudaEventCreate(&e));
  return 0;
}

} // namespace

// registry models of ssm_nlg: shapes and parameter counts
static bool nlg_user(const bssm_nlg_model* M) { return M->model_id >= BSSM_USER_MODEL_BASE; }
static int nlg_state_dim(const bssm_nlg_model* M) { return nlg_user(M) ? M->m : (M->model_id == BSSM_NLG_GROWTH ? 2 : 1); }
static int nlg_n_theta(const bssm_nlg_model* M) {
  if (nlg_user(M)) return 1;
  return M->model_id == BSSM_NLG_AR_EXP ? 4 : (M->model_id == BSSM_NLG_SIN_EXP ? 2 : (M->model_id == BSSM_NLG_LINEAR_GAUSSIAN ? 1 : 3));
}
// allow_user: the bootstrap-filter paths also take models registered with bssm_b200_nlg_register
static int check_nlg_model(const bssm_nlg_model* M, bool allow_user = false) {
  if (nlg_user(M)) {
    if (!allow_user)
      return fail(BSSM_B200_ERR_UNSUPPORTED, "ssm_nlg models from user source run the bootstrap filter; the psi-APF and the EKF "
                                             "particle filter need the host-side approximation (registry models)");
    if (!pfl::user_nlg_exists(M->model_id)) return fail(BSSM_B200_ERR_ARG, "ssm_nlg: model_id was not returned by bssm_b200_nlg_register");
    if (M->p != 1 || M->m < 1 || M->m > 4 || M->k < 1 || M->k > M->m || M->n_theta < 1)
      return fail(BSSM_B200_ERR_ARG, "ssm_nlg user model: p = 1, 1 <= k <= m <= 4, at least one parameter");
    return 0;
  }
  if (M->model_id == BSSM_NLG_GROWTH) {
    if (M->p != 1 || M->m != 2 || M->k != 2 || M->n_theta < 3 || M->n_known < 6)
      return fail(BSSM_B200_ERR_ARG, "growth model: p = 1, m = 2, k = 2, 3 parameters, 6 known parameters");
    return 0;
  }
  if (M->model_id == BSSM_NLG_AR_EXP) {
    if (M->p != 1 || M->m != 1 || M->k != 1 || M->n_theta < 4)
      return fail(BSSM_B200_ERR_ARG, "ar_exp model: p = 1, m = 1, k = 1, 4 parameters (mu, rho, log sigma, log H)");
    return 0;
  }
  if (M->model_id == BSSM_NLG_SIN_EXP) {
    if (M->p != 1 || M->m != 1 || M->k != 1 || M->n_theta < 2)
      return fail(BSSM_B200_ERR_ARG, "sin_exp model: p = 1, m = 1, k = 1, 2 parameters (log H, log R)");
    return 0;
  }
  if (M->model_id == BSSM_NLG_LINEAR_GAUSSIAN) {
    if (M->p != 1 || M->m != 1 || M->k != 1 || M->n_theta < 1)
      return fail(BSSM_B200_ERR_ARG, "linear_gaussian model: p = 1, m = 1, k = 1, 1 parameter (log H)");
    return 0;
  }
  return fail(BSSM_B200_ERR_UNSUPPORTED, "ssm_nlg: unknown model_id (registry: BSSM_NLG_GROWTH, BSSM_NLG_AR_EXP, BSSM_NLG_SIN_EXP, "
                                         "BSSM_NLG_LINEAR_GAUSSIAN)");
}

// everything of NlgParams that depends on theta / known_params (np.n and np.y stay): the registry
// models' a1_fn, P1_fn, H_fn, R_fn constants and, for the EKF particle filter, the first EKF update
static void fill_nlg_params(const bssm_nlg_model* M, int method, pf::NlgParams& np) {
  if (M->model_id == BSSM_NLG_AR_EXP) { // theta = (mu, rho, log sigma, log H), R/cpp_example_models.R:200-236
    const double e = pf_exp(M->theta[2]);
    const double P1 = e * e / (1 - M->theta[1] * M->theta[1]);
    np.a1[0] = M->theta[0];
    psd_chol_host(&P1, 1, np.LP1);
    np.mu = M->theta[0]; np.rho = M->theta[1];
    np.H = pf_exp(M->theta[3]); np.R0 = e;
  } else if (M->model_id == BSSM_NLG_LINEAR_GAUSSIAN) { // a1 = 0, P1 = 1, R = 1, theta = (log H), R/cpp_example_models.R:519-553
    const double P1 = 1.0;
    np.a1[0] = 0.0;
    psd_chol_host(&P1, 1, np.LP1);
    np.mu = 0.0; np.rho = 1.0; // T_fn(alpha) = alpha = mu (1 - rho) + rho alpha exactly
    np.H = pf_exp(M->theta[0]); np.R0 = 1.0;
  } else if (M->model_id == BSSM_NLG_SIN_EXP) { // a1 = 0, P1 = 1, theta = (log H, log R), R/cpp_example_models.R:640-662
    const double P1 = 1.0;
    np.a1[0] = 0.0;
    psd_chol_host(&P1, 1, np.LP1);
    np.H = pf_exp(M->theta[0]); np.R0 = pf_exp(M->theta[1]);
  } else {
    const double* kp = M->known_params;
    np.dT = kp[0]; np.K = kp[1];
    np.a1[0] = kp[2]; np.a1[1] = kp[3];
    const double P1[4] = {kp[4], 0.0, 0.0, kp[5]};
    psd_chol_host(P1, 2, np.LP1);
    np.H = pf_exp(M->theta[0]); np.R0 = pf_exp(M->theta[1]); np.R1 = pf_exp(M->theta[2]);
  }
  np.wconst = std::log(2.0 * M_PI) + 2.0 * pf_log(np.H);
  if (method == BSSM_METHOD_EKPF) {
    if (nlg_state_dim(M) == 1) ekpf_first_column<1>(M, np); else ekpf_first_column<2>(M, np);
  }
}

// a1_fn / P1_fn of a user model for `nt` parameter vectors (J->theta_u): evaluated on the device, P1 factored
// with psd_chol (src/psd_chol.cpp:5-18) on the host, (a1, L) uploaded as the filters' initial distribution
static int nlg_user_initial(bssm_b200_job* J, uint32_t nt) {
  const int m = J->m;
  const size_t blk = (size_t)m + (size_t)m * m;
  int rc;
  if ((rc = J->init_u.alloc(blk * nt))) return rc;
  if ((rc = launch_status(pfl::user_nlg_init(J->nlg_model_id, m, J->np.theta_u, J->np.theta_u_stride, J->np.known_u, nt,
                                             J->init_u.p, J->stream), "ssm_nlg user model (a1_fn, P1_fn)")))
    return rc;
  std::vector<double> h(blk * nt);
  CU(cudaMemcpyAsync(h.data(), J->init_u.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, J->stream));
  CU(cudaStreamSynchronize(J->stream));
  for (uint32_t f = 0; f < nt; ++f) {
    double L[16];
    psd_chol_host(&h[blk * f + m], m, L);
    std::copy(L, L + (size_t)m * m, &h[blk * f + m]);
  }
  CU(cudaMemcpyAsync(J->init_u.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, J->stream));
  CU(cudaStreamSynchronize(J->stream)); // h goes out of scope
  J->np.init_u = J->init_u.p;
  J->np.init_u_stride = nt > 1 ? blk : 0;
  return 0;
}

// draw_theta: n_filters columns of n_par parameters, one bootstrap filter each (is_correction_bsf)
static int create_nlg_job(const bssm_nlg_model* M, int method, uint32_t nsim, uint32_t seed, uint32_t n_filters,
                          uint32_t first_filter_id, int out_flags, int post, bssm_b200_job** job,
                          const double* draw_theta = nullptr, uint32_t n_par = 0) {
  if (!M || !job || !M->y || !M->theta) return fail(BSSM_B200_ERR_ARG, "null argument");
  *job = nullptr;
  int rc = check_nlg_model(M, method == BSSM_METHOD_BSF);
  if (rc) return rc;
  if (M->model_id == BSSM_NLG_GROWTH && !M->known_params) return fail(BSSM_B200_ERR_ARG, "null argument");
  const int m = nlg_state_dim(M);
  bssm_b200_job* J = new bssm_b200_job();
  struct Guard { bssm_b200_job* j; ~Guard() { delete j; } } guard{J};
  rc = job_common(J, KIND_NLG, M->n, m, m, nsim, seed, n_filters, first_filter_id, out_flags,
                  post > BSSM_POST_SMOOTHER ? BSSM_POST_NONE : post);
  if (rc) return rc;
  J->post = post; // PICK / SMOOTHER_MERGE: the IS-correction drivers
  J->nlg_model_id = M->model_id;
  if ((rc = J->y.upload(M->y, M->n))) return rc;
  pf::NlgParams& np = J->np;
  memset(&np, 0, sizeof(np));
  np.n = M->n; np.y = J->y.p;
  if (method == BSSM_METHOD_EKPF) J->method = BSSM_METHOD_EKPF;
  if (nlg_user(M)) {
    // functors from user source: theta (one per filter for the correction phase) and known_params on the
    // device; a1_fn / P1_fn evaluated there, psd_chol(P1) on the host
    J->k = M->k;
    const double one = 0.0;
    if ((rc = J->theta_u.upload(draw_theta ? draw_theta : M->theta, draw_theta ? (size_t)n_par * n_filters : (size_t)M->n_theta)) ||
        (rc = J->known_u.upload(M->n_known > 0 ? M->known_params : &one, M->n_known > 0 ? (size_t)M->n_known : 1)))
      return rc;
    np.theta_u = J->theta_u.p; np.theta_u_stride = draw_theta ? n_par : 0;
    np.known_u = J->known_u.p;
    if ((rc = nlg_user_initial(J, draw_theta ? n_filters : 1u))) return rc;
  } else {
    fill_nlg_params(M, method, np);
  }
  if (draw_theta && !nlg_user(M)) {
    if (method != BSSM_METHOD_BSF) return fail(BSSM_B200_ERR_UNSUPPORTED, "per-filter theta: bootstrap filter only");
    std::vector<double> blk((size_t)pf::NLG_DRAW_BLOCK * n_filters);
    for (uint32_t f = 0; f < n_filters; ++f) {
      bssm_nlg_model Mi = *M;
      Mi.theta = draw_theta + (size_t)f * n_par;
      Mi.n_theta = (int32_t)n_par;
      pf::NlgParams di = np;
      fill_nlg_params(&Mi, method, di);
      double* b = &blk[(size_t)pf::NLG_DRAW_BLOCK * f];
      b[0] = di.a1[0]; b[1] = di.a1[1];
      for (int e = 0; e < 4; ++e) b[2 + e] = di.LP1[e];
      b[6] = di.H; b[7] = di.R0; b[8] = di.R1; b[9] = di.wconst; b[10] = di.mu; b[11] = di.rho;
    }
    if ((rc = J->theta_f.upload(blk.data(), blk.size()))) return rc;
    np.draw_blk = J->theta_f.p;
  }
  if ((rc = finish_job(J, out_flags, post))) return rc;
  guard.j = nullptr;
  *job = J;
  return 0;
}

int bssm_b200_job_create_nlg(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, uint32_t n_filters,
                             uint32_t first_filter_id, int out_flags, int post, bssm_b200_job** job) {
  return create_nlg_job(M, BSSM_METHOD_BSF, nsim, seed, n_filters, first_filter_id, out_flags, post, job);
}
int bssm_b200_job_create_nlg_ekpf(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, uint32_t n_filters,
                                  uint32_t first_filter_id, int out_flags, int post, bssm_b200_job** job) {
  return create_nlg_job(M, BSSM_METHOD_EKPF, nsim, seed, n_filters, first_filter_id, out_flags, post, job);
}

// The filter inside an MCMC loop for ssm_nlg (bootstrap / EKF particle filter) and ssm_sde models:
// nonlinear_pm_mcmc / sde_pm_mcmc run the filter once per iteration at a new theta (src/R_mcmc.cpp:551-609 ->
// src/mcmc.cpp:338-491; src/R_sde.cpp:114-166 -> src/mcmc.cpp:694-820, model.bsf_filter at :710, :759).
static int nlg_psi_approximation(bssm_b200_job* J, const bssm_nlg_model* M);
int bssm_b200_job_set_theta_nlg(bssm_b200_job* J, const bssm_nlg_model* M, const double* theta, uint32_t n_par) {
  if (!J || !M || !theta) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (J->kind != KIND_NLG || M->model_id != J->nlg_model_id || M->n != J->n)
    return fail(BSSM_B200_ERR_ARG, "job_set_theta_nlg: an ssm_nlg job created from the same model and data");
  bssm_nlg_model Mi = *M;
  Mi.theta = theta;
  Mi.n_theta = (int32_t)n_par;
  int rc = check_nlg_model(&Mi, J->method == BSSM_METHOD_BSF);
  if (rc) return rc;
  if (nlg_user(M)) { // new theta on the device, a1_fn / P1_fn again
    if (J->np.theta_u_stride) return fail(BSSM_B200_ERR_ARG, "job_set_theta_nlg: IS-correction jobs carry one theta per filter already");
    if (n_par != J->theta_u.count) return fail(BSSM_B200_ERR_ARG, "job_set_theta_nlg: theta has another length than the model's");
    CU(cudaMemcpyAsync(J->theta_u.p, theta, sizeof(double) * n_par, cudaMemcpyHostToDevice, J->stream));
    CU(cudaStreamSynchronize(J->stream));
    return nlg_user_initial(J, 1u);
  }
  if (J->method == BSSM_METHOD_PSI) {
    if (J->npp.fstride_stepc != 0) return fail(BSSM_B200_ERR_ARG, "job_set_theta_nlg: post-correction jobs carry one theta per filter already");
    return nlg_psi_approximation(J, &Mi); // approximate() at the new theta on the host, same buffers
  }
  fill_nlg_params(&Mi, J->method, J->np);
  return 0;
}

static void sde_theta_block(int model_id, const double* th, double* b);
int bssm_b200_job_set_theta_sde(bssm_b200_job* J, const double* theta, uint32_t n_par) {
  if (!J || !theta) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (J->kind != KIND_SDE) return fail(BSSM_B200_ERR_ARG, "job_set_theta_sde: an ssm_sde job");
  if (J->sp.theta_f || J->sp.theta_u_stride) return fail(BSSM_B200_ERR_ARG, "job_set_theta_sde: IS-correction jobs carry one theta per filter already");
  if (J->sde_model_id >= BSSM_USER_MODEL_BASE) { // user functors read theta from the device
    if (n_par != J->theta_u.count) return fail(BSSM_B200_ERR_ARG, "job_set_theta_sde: theta has another length than the model's");
    CU(cudaMemcpyAsync(J->theta_u.p, theta, sizeof(double) * n_par, cudaMemcpyHostToDevice, J->stream));
    CU(cudaStreamSynchronize(J->stream)); // theta is the caller's buffer
    return 0;
  }
  if (n_par < 3) return fail(BSSM_B200_ERR_ARG, "ssm_sde registry models need 3 parameters");
  double b[6];
  sde_theta_block(J->sde_model_id, theta, b);
  pf::SdeParams& sp = J->sp;
  sp.th0 = theta[0]; sp.th1 = theta[1]; sp.th2 = theta[2];
  sp.logsd = b[3]; sp.e0 = b[4]; sp.e2 = b[5];
  return 0;
}

// the 32 doubles per time point that the psi-APF weights of an ssm_nlg registry model read
// (layout: NlgPsiParams in models_nlg_sde.cuh), m = 1 or 2
static void pack_nlg_step(const pfga::NlgApprox& ga, const double* y, int m, int t, double* c) {
  const double LOG2PI = std::log(2.0 * M_PI);
  const size_t mm = (size_t)m * m;
  const double Ht = ga.H[t], HHt = Ht * Ht;
  c[0] = y[t]; c[1] = LOG2PI + 2.0 * pf_log(Ht); c[2] = Ht; c[3] = LOG2PI + pf_log(HHt); c[4] = HHt;
  c[5] = ga.D[t]; c[8] = ga.scales[t];
  for (int d = 0; d < m; ++d) c[6 + d] = ga.Z[(size_t)m * t + d];
  if (t >= 1) {
    for (int d = 0; d < m; ++d) c[10 + d] = ga.C[(size_t)m * (t - 1) + d];
    for (size_t e = 0; e < mm; ++e) c[12 + e] = ga.T[mm * (t - 1) + e];
    if (m == 1) { // scalar dmvnorm branch: variances instead of inverse Cholesky factors
      c[16] = ga.f_approx[t - 1].var1;
      c[21] = ga.f_true[t - 1].var1;
      c[26] = ga.fn.theta[0]; c[27] = ga.fn.theta[1]; // ar_exp: T_fn = mu (1 - rho) + rho alpha
      if (ga.fn.model_id == BSSM_NLG_LINEAR_GAUSSIAN) { c[26] = 0.0; c[27] = 1.0; } // T_fn(alpha) = alpha
    } else {
      for (size_t e = 0; e < mm; ++e) {
        c[16 + e] = ga.f_approx[t - 1].rooti[e];
        c[21 + e] = ga.f_true[t - 1].rooti[e];
      }
    }
    c[20] = ga.f_approx[t - 1].cst;
    c[25] = ga.f_true[t - 1].cst;
  }
}

// Gaussian approximation of an ssm_nlg model at M->theta on the host ((iterated) EKF start, relinearisation
// sweeps, Kalman smoother), packed and uploaded for the job's replicate filters (which share it)
static int nlg_psi_approximation(bssm_b200_job* J, const bssm_nlg_model* M) {
  const int n = M->n, m = nlg_state_dim(M);
  const uint32_t n_filters = J->n_filters;
  int rc;
  pfga::NlgApprox ga(M);
  ga.max_iter = (int)J->nlg_max_iter;
  ga.conv_tol = J->nlg_conv_tol;
  ga.iekf_iter = (int)J->nlg_iekf_iter;
  ga.run(nullptr);
  if (!ga.finite_mode) return fail(BSSM_B200_ERR_ARG, "Approximation did not converge.");
  std::vector<double> sc((size_t)32 * n, 0.0);
  for (int t = 0; t < n; ++t) pack_nlg_step(ga, M->y, m, t, &sc[(size_t)32 * t]);
  std::vector<double> al(n_filters, ga.approx_loglik);
  if ((rc = J->stepc.upload(sc.data(), sc.size())) || (rc = J->alphahat.upload(ga.alphahat.data(), ga.alphahat.size())) ||
      (rc = J->Lt.upload(ga.L.data(), ga.L.size())) || (rc = J->Ct.upload(ga.Ct.data(), ga.Ct.size())) ||
      (rc = J->approx_loglik.upload(al.data(), n_filters)))
    return rc;
  pf::NlgPsiParams& pp = J->npp;
  pp.n = n; pp.stepc = J->stepc.p; pp.alphahat = J->alphahat.p; pp.Lt = J->Lt.p; pp.Ct = J->Ct.p;
  pp.approx_loglik = J->approx_loglik.p;
  pp.dT = 0.0; pp.K = 0.0;
  if (M->model_id == BSSM_NLG_GROWTH) { pp.dT = M->known_params[0]; pp.K = M->known_params[1]; }
  pp.fstride_stepc = 0; pp.fstride_a = 0; pp.fstride_m = 0; // replicates share the approximation
  return 0;
}

static int create_nlg_psi_job(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, uint32_t max_iter,
                              double conv_tol, uint32_t iekf_iter, uint32_t n_filters, uint32_t first_filter_id,
                              int out_flags, int post, bssm_b200_job** job) {
  if (!M || !job || !M->y || !M->theta) return fail(BSSM_B200_ERR_ARG, "null argument");
  *job = nullptr;
  int rc = check_nlg_model(M);
  if (rc) return rc;
  if (M->model_id == BSSM_NLG_GROWTH && !M->known_params) return fail(BSSM_B200_ERR_ARG, "null argument");
  const int n = M->n, m = nlg_state_dim(M);
  bssm_b200_job* J = new bssm_b200_job();
  struct Guard { bssm_b200_job* j; ~Guard() { delete j; } } guard{J};
  rc = job_common(J, KIND_NLG, n, m, m, nsim, seed, n_filters, first_filter_id, out_flags, post);
  if (rc) return rc;
  J->method = BSSM_METHOD_PSI;
  J->nlg_model_id = M->model_id;
  J->nlg_max_iter = max_iter; J->nlg_conv_tol = conv_tol; J->nlg_iekf_iter = iekf_iter;
  if ((rc = nlg_psi_approximation(J, M))) return rc;
  if ((rc = finish_job(J, out_flags, post))) return rc;
  guard.j = nullptr;
  *job = J;
  return 0;
}

int bssm_b200_job_create_nlg_psi(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, uint32_t max_iter,
                                 double conv_tol, uint32_t n_filters, uint32_t first_filter_id,
                                 int out_flags, int post, bssm_b200_job** job) {
  return create_nlg_psi_job(M, nsim, seed, max_iter, conv_tol, 0, n_filters, first_filter_id, out_flags, post, job);
}

// the derived parameters the device functors read, for one theta: {th0, th1, th2, logsd, e0, e2}
static void sde_theta_block(int model_id, const double* th, double* b) {
  b[0] = th[0]; b[1] = th[1]; b[2] = th[2]; b[3] = 0.0; b[4] = 0.0; b[5] = 0.0;
  if (model_id == BSSM_SDE_GBM) b[3] = pf_log(th[2]);
  else { b[4] = pf_exp(th[0]); b[5] = pf_exp(th[2]); } // Poisson-OU: theta = (log rho, nu, log sigma)
}

// draw_theta: null (all filters share model->theta) or n_par x n_filters, one theta per filter
static int create_sde_job(const bssm_sde_model* M, uint32_t nsim, uint32_t L, uint32_t seed,
                          uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post,
                          const double* draw_theta, uint32_t n_par, bssm_b200_job** job) {
  if (!M || !job || !M->y || !M->theta) return fail(BSSM_B200_ERR_ARG, "null argument");
  *job = nullptr;
  const bool user = M->model_id >= BSSM_USER_MODEL_BASE;
  if (user && !pfl::user_sde_exists(M->model_id)) return fail(BSSM_B200_ERR_ARG, "ssm_sde: model_id was not returned by bssm_b200_sde_register");
  if (!user && M->model_id != BSSM_SDE_GBM && M->model_id != BSSM_SDE_POISSON_OU)
    return fail(BSSM_B200_ERR_UNSUPPORTED, "ssm_sde: unknown model_id (registry: BSSM_SDE_GBM, BSSM_SDE_POISSON_OU; "
                                           "or a model registered with bssm_b200_sde_register)");
  if (!user && M->n_theta < 3) return fail(BSSM_B200_ERR_ARG, "ssm_sde registry models need 3 parameters");
  if (user && M->n_theta < 1) return fail(BSSM_B200_ERR_ARG, "ssm_sde: empty theta");
  if (L < 1 || L > 20) return fail(BSSM_B200_ERR_ARG, "discretisation level L must be in 1..20");
  bssm_b200_job* J = new bssm_b200_job();
  struct Guard { bssm_b200_job* j; ~Guard() { delete j; } } guard{J};
  int rc = job_common(J, KIND_SDE, M->n, 1, 1, nsim, seed, n_filters, first_filter_id, out_flags,
                      post > BSSM_POST_SMOOTHER ? BSSM_POST_NONE : post);
  if (rc) return rc;
  J->post = post; // PICK / SMOOTHER_MERGE: the IS-correction drivers
  if ((rc = J->y.upload(M->y, M->n))) return rc;
  pf::SdeParams& sp = J->sp;
  sp.n = M->n; sp.y = J->y.p; sp.x0 = M->x0; sp.positive = M->positive ? 1 : 0;
  sp.nsub = 1 << L;
  sp.dt = 1.0 / (double)sp.nsub;      // t / n with t = 1, src/milstein.cpp:23-24
  sp.sqdt = std::sqrt(sp.dt);
  J->sde_model_id = M->model_id;
  sp.lgy = nullptr; sp.theta_f = nullptr; sp.theta_u = nullptr; sp.theta_u_stride = 0;
  if (user) {
    // the functors read the theta vector itself (src/model_ssm_sde.h:8-11), from device memory
    if (draw_theta) {
      if ((rc = J->theta_u.upload(draw_theta, (size_t)n_par * n_filters))) return rc;
      sp.theta_u_stride = n_par;
    } else if ((rc = J->theta_u.upload(M->theta, (size_t)M->n_theta))) {
      return rc;
    }
    sp.theta_u = J->theta_u.p;
  } else {
    sp.th0 = M->theta[0]; sp.th1 = M->theta[1]; sp.th2 = M->theta[2];
    double b[6];
    sde_theta_block(M->model_id, M->theta, b);
    sp.logsd = b[3]; sp.e0 = b[4]; sp.e2 = b[5];
  }
  if (draw_theta && !user) {
    if (n_par < 3) return fail(BSSM_B200_ERR_ARG, "ssm_sde registry models need 3 parameters");
    std::vector<double> blk((size_t)6 * n_filters);
    for (uint32_t f = 0; f < n_filters; ++f) sde_theta_block(M->model_id, draw_theta + (size_t)f * n_par, &blk[(size_t)6 * f]);
    if ((rc = J->theta_f.upload(blk.data(), blk.size()))) return rc;
    sp.theta_f = J->theta_f.p;
  }
  if (M->model_id == BSSM_SDE_POISSON_OU) { // log y! per time point, R/cpp_example_models.R:37-39
    std::vector<double> lg(M->n, 0.0);
    for (int t = 0; t < M->n; ++t) if (std::isfinite(M->y[t])) lg[t] = std::lgamma(M->y[t] + 1.0);
    if ((rc = J->lgy.upload(lg.data(), M->n))) return rc;
    sp.lgy = J->lgy.p;
  }
  if ((rc = finish_job(J, out_flags, post))) return rc;
  guard.j = nullptr;
  *job = J;
  return 0;
}

int bssm_b200_nlg_register(const char* cuda_source, int32_t m, int32_t k, int32_t* model_id) {
  if (!cuda_source || !model_id) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (m < 1 || m > 4 || k < 1 || k > m) return fail(BSSM_B200_ERR_ARG, "ssm_nlg user model: 1 <= k <= m <= 4");
  int rc = require_device();
  if (rc) return rc;
  int id = -1;
  rc = pfl::user_nlg_register(cuda_source, m, k, &id);
  if (rc) return launch_status(rc, "ssm_nlg user model");
  *model_id = id;
  return 0;
}

int bssm_b200_sde_register(const char* cuda_source, int32_t* model_id) {
  if (!cuda_source || !model_id) return fail(BSSM_B200_ERR_ARG, "null argument");
  int rc = require_device();
  if (rc) return rc;
  int id = -1;
  rc = pfl::user_sde_register(cuda_source, &id);
  if (rc) return launch_status(rc, "ssm_sde user model");
  *model_id = id;
  return 0;
}

int bssm_b200_job_create_sde(const bssm_sde_model* M, uint32_t nsim, uint32_t L, uint32_t seed,
                             uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post,
                             bssm_b200_job** job) {
  return create_sde_job(M, nsim, L, seed, n_filters, first_filter_id, out_flags, post, nullptr, 0, job);
}

// state-independent constants of the ssm_mng log-likelihood, src/model_ssm_mng.cpp:636-666 (pre-summed)
static double mng_dist_const(const bssm_mng_model* M) {
  const int n = M->n, p = M->p;
  auto lch = [](double nn, double kk) {
    return -std::log(nn + 1) - std::lgamma(nn - kk + 1) - std::lgamma(kk + 1) + std::lgamma(nn + 2);
  };
  double cst = 0.0;
  for (int i = 0; i < p; ++i) {
    const double phi = M->phi[i];
    for (int t = 0; t < n; ++t) {
      const double y = M->y[t + (size_t)n * i], u = M->u[t + (size_t)n * i];
      if (!std::isfinite(y)) continue;
      switch (M->distribution[i]) {
        case 1: cst += -std::lgamma(y + 1) + y * std::log(u); break;
        case 2: cst += lch(u, y); break;
        case 3: cst += lch(y + phi - 1, y) + phi * std::log(phi) + y * std::log(u); break;
        case 4: cst += phi * std::log(phi) - std::lgamma(phi) + (phi - 1) * std::log(y) - phi * std::log(u); break;
        case 5: cst += -0.5 * std::log(2.0 * M_PI) - std::log(phi); break;
      }
    }
  }
  return cst;
}

// parameter sets of ssm_mng (parset_mng, src/parset_ng.h:42-62) for the filters of one job: slices shaped like
// the model's own arrays, empty = shared; modes p x n per filter (psi-APF)
struct MngDraws {
  std::vector<double> phi, Z, T, R, a1, P1, D, C, modes;
};
// parset_mng::update(model, i), src/parset_ng.cpp:185-215
static void draw_model_mng(const bssm_mng_model* M, const MngDraws& dm, size_t f, bssm_mng_model& Mi) {
  Mi = *M;
  const size_t n = (size_t)M->n, m = (size_t)M->m, k = (size_t)M->k, p = (size_t)M->p;
  if (!dm.phi.empty()) Mi.phi = &dm.phi[p * f];
  if (!dm.Z.empty()) Mi.Z = &dm.Z[p * m * (M->Ztv ? n : 1) * f];
  if (!dm.T.empty()) Mi.T = &dm.T[m * m * (M->Ttv ? n : 1) * f];
  if (!dm.R.empty()) Mi.R = &dm.R[m * k * (M->Rtv ? n : 1) * f];
  if (!dm.a1.empty()) Mi.a1 = &dm.a1[m * f];
  if (!dm.P1.empty()) Mi.P1 = &dm.P1[m * m * f];
  if (!dm.D.empty()) Mi.D = &dm.D[p * (M->Dtv ? n : 1) * f];
  if (!dm.C.empty()) Mi.C = &dm.C[m * (M->Ctv ? n : 1) * f];
}

static int create_mng_job(const bssm_mng_model* M, int method, uint32_t nsim, uint32_t seed, uint32_t n_filters,
                          uint32_t first_filter_id, int out_flags, int post, bssm_b200_job** job,
                          const MngDraws* dm = nullptr) {
  if (!M || !job || !M->y || !M->u || !M->Z || !M->T || !M->R || !M->a1 || !M->P1 || !M->D || !M->C ||
      !M->phi || !M->distribution)
    return fail(BSSM_B200_ERR_ARG, "null argument");
  *job = nullptr;
  if (M->p < 1 || M->p > 16 || M->m < 1 || M->m > 16 || M->k < 1 || M->k > M->m)
    return fail(BSSM_B200_ERR_ARG, "ssm_mng: 1 <= p <= 16, 1 <= k <= m <= 16");
  for (int j = 0; j < M->p; ++j)
    if (M->distribution[j] < 1 || M->distribution[j] > 5) return fail(BSSM_B200_ERR_ARG, "ssm_mng: distribution codes 1..5");
  bssm_b200_job* J = new bssm_b200_job();
  struct Guard { bssm_b200_job* j; ~Guard() { delete j; } } guard{J};
  const int n = M->n, m = M->m, k = M->k, p = M->p;
  int rc = job_common(J, KIND_MNG, n, m, k, nsim, seed, n_filters, first_filter_id, out_flags,
                      post > BSSM_POST_SMOOTHER ? BSSM_POST_NONE : post);
  if (rc) return rc;
  J->post = post; // PICK / SMOOTHER_MERGE: the IS-correction drivers
  {
    // (m, k) without a specialised instantiation (k_mng.cu): the run-time-m policies, two particles per thread
    const int mk = m * 10 + k;
    const bool inst = method == BSSM_METHOD_PSI ? m <= 3 : (mk == 11 || mk == 21 || mk == 22 || mk == 33);
    J->dyn = !inst;
    if (J->dyn && (rc = choose_geometry(nsim, n_filters, &J->geo, true))) return rc;
  }
  std::vector<double> LP1;
  chol_posdiag(M->P1, m, LP1);
  std::vector<int> dist(M->distribution, M->distribution + p);
  if ((rc = J->y.upload(M->y, (size_t)n * p)) || (rc = J->u.upload(M->u, (size_t)n * p)) ||
      (rc = J->Z.upload(M->Z, (size_t)p * m * (M->Ztv ? n : 1))) ||
      (rc = J->T.upload(M->T, (size_t)m * m * (M->Ttv ? n : 1))) ||
      (rc = J->R.upload(M->R, (size_t)m * k * (M->Rtv ? n : 1))) || (rc = J->a1.upload(M->a1, m)) ||
      (rc = J->LP1.upload(LP1.data(), (size_t)m * m)) || (rc = J->D.upload(M->D, (size_t)p * (M->Dtv ? n : 1))) ||
      (rc = J->C.upload(M->C, (size_t)m * (M->Ctv ? n : 1))) || (rc = J->phi_v.upload(M->phi, p)) ||
      (rc = J->dist_v.upload(dist.data(), p)))
    return rc;
  pf::MngParams& gp = J->gp;
  gp.n = n; gp.p = p; gp.m_dyn = m; gp.k_dyn = k;
  gp.Ztv = M->Ztv; gp.Ttv = M->Ttv; gp.Rtv = M->Rtv; gp.Dtv = M->Dtv; gp.Ctv = M->Ctv;
  gp.y = J->y.p; gp.u = J->u.p; gp.Z = J->Z.p; gp.T = J->T.p; gp.R = J->R.p; gp.a1 = J->a1.p; gp.LP1 = J->LP1.p;
  gp.D = J->D.p; gp.C = J->C.p; gp.phi = J->phi_v.p; gp.distribution = J->dist_v.p;
  gp.loglik_const = mng_dist_const(M);
  if (dm) {
    // one model per filter: upload the per-filter slices of what the parameter set carries
    J->mng_draws = true;
    auto slices = [&](const std::vector<double>& v, DevBuf<double>& buf, size_t slice, const double** ptr,
                      size_t* stride) -> int {
      if (v.empty()) return 0;
      if (v.size() != slice * n_filters) return fail(BSSM_B200_ERR_ARG, "parameter set: slice size x n_filters expected");
      int r = buf.upload(v.data(), v.size());
      if (r) return r;
      *ptr = buf.p; *stride = slice;
      return 0;
    };
    if ((rc = slices(dm->Z, J->Z, (size_t)p * m * (M->Ztv ? n : 1), &gp.Z, &gp.fs_Z)) ||
        (rc = slices(dm->T, J->T, (size_t)m * m * (M->Ttv ? n : 1), &gp.T, &gp.fs_T)) ||
        (rc = slices(dm->R, J->R, (size_t)m * k * (M->Rtv ? n : 1), &gp.R, &gp.fs_R)) ||
        (rc = slices(dm->C, J->C, (size_t)m * (M->Ctv ? n : 1), &gp.C, &gp.fs_C)) ||
        (rc = slices(dm->D, J->D, (size_t)p * (M->Dtv ? n : 1), &gp.D, &gp.fs_D)) ||
        (rc = slices(dm->a1, J->a1, (size_t)m, &gp.a1, &gp.fs_a1)) ||
        (rc = slices(dm->phi, J->phi_v, (size_t)p, &gp.phi, &gp.fs_phi)))
      return rc;
    if (!dm->P1.empty()) {
      std::vector<double> L((size_t)m * m * n_filters), Lf;
      for (uint32_t f = 0; f < n_filters; ++f) {
        chol_posdiag(&dm->P1[(size_t)m * m * f], m, Lf);
        std::copy(Lf.begin(), Lf.end(), &L[(size_t)m * m * f]);
      }
      if ((rc = slices(L, J->LP1, (size_t)m * m, &gp.LP1, &gp.fs_LP1))) return rc;
    }
    if (!dm->phi.empty() && method == BSSM_METHOD_BSF) {
      std::vector<double> lc(n_filters);
      for (uint32_t f = 0; f < n_filters; ++f) {
        bssm_mng_model Mi = *M;
        Mi.phi = &dm->phi[(size_t)p * f];
        lc[f] = mng_dist_const(&Mi);
      }
      if ((rc = J->lconst_f.upload(lc.data(), n_filters))) return rc;
      gp.lconst_f = J->lconst_f.p;
    }
  }
  if (method == BSSM_METHOD_PSI && dm) {
    // approximate_for_is (src/model_ssm_mng.cpp:119-135) per draw on the host, OpenMP over draws
    if (dm->modes.size() != (size_t)p * n * n_filters) return fail(BSSM_B200_ERR_ARG, "parameter set: p x n modes per filter");
    J->method = BSSM_METHOD_PSI;
    const size_t n1 = (size_t)n + 1, mm = (size_t)m * m, np = (size_t)n * p, nf = n_filters;
    if (!env_int("BSSM_B200_HOST_APPROX", 0)) {
      // one thread per draw on the device (approx_device.cuh); the per-draw model arrays are those uploaded above
      pf::MngApproxParams ap;
      memset(&ap, 0, sizeof(ap));
      ap.n = n; ap.m = m; ap.k = k; ap.p = p;
      ap.Ztv = M->Ztv; ap.Ttv = M->Ttv; ap.Rtv = M->Rtv; ap.Dtv = M->Dtv; ap.Ctv = M->Ctv;
      ap.y = J->y.p; ap.u = J->u.p; ap.distribution = J->dist_v.p;
      ap.Z = gp.Z; ap.fs_Z = gp.fs_Z; ap.T = gp.T; ap.fs_T = gp.fs_T; ap.R = gp.R; ap.fs_R = gp.fs_R;
      ap.C = gp.C; ap.fs_C = gp.fs_C; ap.D = gp.D; ap.fs_D = gp.fs_D; ap.a1 = gp.a1; ap.fs_a1 = gp.fs_a1;
      ap.phi = gp.phi; ap.fs_phi = gp.fs_phi;
      if (!dm->P1.empty()) {
        if ((rc = J->P1.upload(dm->P1.data(), dm->P1.size()))) return rc;
        ap.fs_P1 = mm;
      } else if ((rc = J->P1.upload(M->P1, mm))) {
        return rc;
      }
      ap.P1 = J->P1.p;
      const size_t per_t = (size_t)(3 * p + 1 + 2 * p) + 3 * (size_t)p * m;
      if ((rc = J->mode.upload(dm->modes.data(), dm->modes.size())) || (rc = J->alphahat.alloc((size_t)m * n1 * nf)) ||
          (rc = J->Lt.alloc(mm * n1 * nf)) || (rc = J->Ct.alloc(mm * n1 * nf)) || (rc = J->ytilde.alloc(np * nf)) ||
          (rc = J->Htilde.alloc(np * nf)) || (rc = J->scales.alloc((size_t)n * nf)) || (rc = J->approx_loglik.alloc(nf)) ||
          (rc = J->scratch.alloc(per_t * n * nf)))
        return rc;
      CU(cudaMemsetAsync(J->approx_loglik.p, 0, sizeof(double) * nf, J->stream));
      ap.modes = J->mode.p;
      ap.alphahat = J->alphahat.p; ap.Lt = J->Lt.p; ap.Ct = J->Ct.p; ap.ytilde = J->ytilde.p; ap.Htilde = J->Htilde.p;
      ap.scales = J->scales.p; ap.scratch = J->scratch.p;
      const uint32_t ab = cdiv((uint32_t)nf, 32);
      if (m <= 3 && p <= 8) pf::mng_approx_for_is_kernel<3, 8><<<ab, 32, 0, J->stream>>>(ap, (uint32_t)nf);
      else pf::mng_approx_for_is_kernel<16, 16><<<ab, 32, 0, J->stream>>>(ap, (uint32_t)nf);
      CU(cudaGetLastError());
      CU(cudaStreamSynchronize(J->stream));
    } else {
    std::vector<double> ah((size_t)m * n1 * nf), Lt(mm * n1 * nf), Ct(mm * n1 * nf), yt(np * nf), Ht(np * nf),
        sc((size_t)n * nf), al(nf, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < (int64_t)nf; ++f) {
      bssm_mng_model Mi;
      draw_model_mng(M, *dm, (size_t)f, Mi);
      pfga::MngApprox ga(&Mi);
      ga.run(&dm->modes[np * f]);
      al[f] = ga.approx_loglik;
      std::copy(ga.alphahat.begin(), ga.alphahat.end(), &ah[(size_t)m * n1 * f]);
      std::copy(ga.L.begin(), ga.L.end(), &Lt[mm * n1 * f]);
      std::copy(ga.Ct.begin(), ga.Ct.end(), &Ct[mm * n1 * f]);
      std::copy(ga.scales.begin(), ga.scales.end(), &sc[(size_t)n * f]);
      for (int t = 0; t < n; ++t)
        for (int j = 0; j < p; ++j) { // pseudo-observations in the layout of y (n x p)
          yt[np * f + t + (size_t)n * j] = ga.ay[j + (size_t)p * t];
          Ht[np * f + t + (size_t)n * j] = ga.aH[j + (size_t)p * t];
        }
    }
    if ((rc = J->alphahat.upload(ah.data(), ah.size())) || (rc = J->Lt.upload(Lt.data(), Lt.size())) ||
        (rc = J->Ct.upload(Ct.data(), Ct.size())) || (rc = J->ytilde.upload(yt.data(), yt.size())) ||
        (rc = J->Htilde.upload(Ht.data(), Ht.size())) || (rc = J->scales.upload(sc.data(), sc.size())) ||
        (rc = J->approx_loglik.upload(al.data(), nf)))
      return rc;
    }
    gp.alphahat = J->alphahat.p; gp.Lt = J->Lt.p; gp.Ct = J->Ct.p; gp.ytilde = J->ytilde.p;
    gp.Htilde = J->Htilde.p; gp.scales = J->scales.p; gp.approx_loglik = J->approx_loglik.p;
    gp.fs_a = (size_t)m * n1; gp.fs_m = mm * n1; gp.fs_np = np; gp.fs_n = (size_t)n;
  } else if (method == BSSM_METHOD_PSI) {
    // Gaussian approximation on the host (O(n (m^3 + p^3)), sequential in t), shared by all replicates
    if (!M->initial_mode) return fail(BSSM_B200_ERR_ARG, "psi-APF needs initial_mode");
    if (M->max_iter < 1) return fail(BSSM_B200_ERR_ARG, "max_iter must be >= 1");
    J->method = BSSM_METHOD_PSI;
    pfga::MngApprox ga(M);
    ga.run(nullptr);
    // pseudo-observations in the layout of y (n x p)
    std::vector<double> yt((size_t)n * p), Ht((size_t)n * p);
    for (int t = 0; t < n; ++t)
      for (int j = 0; j < p; ++j) {
        yt[t + (size_t)n * j] = ga.ay[j + (size_t)p * t];
        Ht[t + (size_t)n * j] = ga.aH[j + (size_t)p * t];
      }
    std::vector<double> al(n_filters, ga.approx_loglik);
    if ((rc = J->alphahat.upload(ga.alphahat.data(), ga.alphahat.size())) || (rc = J->Lt.upload(ga.L.data(), ga.L.size())) ||
        (rc = J->Ct.upload(ga.Ct.data(), ga.Ct.size())) || (rc = J->ytilde.upload(yt.data(), yt.size())) ||
        (rc = J->Htilde.upload(Ht.data(), Ht.size())) || (rc = J->scales.upload(ga.scales.data(), n)) ||
        (rc = J->approx_loglik.upload(al.data(), n_filters)))
      return rc;
    gp.alphahat = J->alphahat.p; gp.Lt = J->Lt.p; gp.Ct = J->Ct.p; gp.ytilde = J->ytilde.p;
    gp.Htilde = J->Htilde.p; gp.scales = J->scales.p; gp.approx_loglik = J->approx_loglik.p;
  }
  if ((rc = finish_job(J, out_flags, post))) return rc;
  guard.j = nullptr;
  *job = J;
  return 0;
}

int bssm_b200_job_create_mng(const bssm_mng_model* M, uint32_t nsim, uint32_t seed, uint32_t n_filters,
                             uint32_t first_filter_id, int out_flags, int post, bssm_b200_job** job) {
  return create_mng_job(M, BSSM_METHOD_BSF, nsim, seed, n_filters, first_filter_id, out_flags, post, job);
}
int bssm_b200_job_create_mng_psi(const bssm_mng_model* M, uint32_t nsim, uint32_t seed, uint32_t n_filters,
                                 uint32_t first_filter_id, int out_flags, int post, bssm_b200_job** job) {
  return create_mng_job(M, BSSM_METHOD_PSI, nsim, seed, n_filters, first_filter_id, out_flags, post, job);
}

// One-shot calls with large host outputs: ship finished blocks of time points to the host
// WHILE the persistent filter kernel is still running.  The kernel publishes its progress
// (complete columns) through host-mapped memory every `every` steps; this thread polls it and
// enqueues, on a second stream, the transpose of the block into the cube layout and the strided
// device-to-host copies.  For C2 (1.05 GB of outputs, PCIe-bound) the call then takes about
// max(kernel, copy) instead of kernel + copy.  Falls back to the plain path when there is nothing
// to stream; a filter that stopped early is re-fetched the plain way (zero-filled tail).
static int run_fetch_streaming(bssm_b200_job* J, bssm_filter_out* out) {
  const bool smoothed = J->post == BSSM_POST_SMOOTHER || J->post == BSSM_POST_SMOOTHER_MERGE;
  const bool s_alpha = out->alpha && J->keep_alpha && !smoothed; // the smoother rewrites alpha afterwards
  const bool s_w = out->weights && J->keep_weights;
  const bool s_i = out->indices && J->keep_indices;
  const int n = J->n, m = J->m;
  const uint32_t N = J->N;
  const size_t bytes = (size_t)N * (n + 1) * 8 * ((s_alpha ? m : 0) + (s_w ? 1 : 0));
  const size_t min_bytes = (size_t)env_int("BSSM_B200_STREAM_MIN_MB", 32) << 20; // tests force 0
  if (J->n_filters != 1 || bytes == 0 || bytes < min_bytes || env_int("BSSM_B200_NO_STREAMING", 0) ||
      J->geo.sync_mode == pf::SYNC_SMALL) { // the latency kernel does not publish its progress
    float ms = 0;
    int rc = bssm_b200_job_run(J, &ms);
    return rc ? rc : bssm_b200_job_fetch(J, 0, out);
  }
  int rc;
  std::lock_guard<std::mutex> sguard(g_sctx.mu);
  CU(g_sctx.ensure());
  StreamCtx& R = g_sctx;
  // whatever path leaves this function, copies that read the job's buffers on the second stream have finished
  // before the caller may destroy the job (its buffers go back to the recycling pool)
  struct S2Guard { cudaStream_t s; ~S2Guard() { cudaStreamSynchronize(s); } } s2guard{R.s2};
  *R.flag = 0;
  int* dflag = R.dflag;
  const int every = std::max(8, (n + 1) / std::max(1, env_int("BSSM_B200_STREAM_CHUNKS", 6)));
  const bool dbg = env_int("BSSM_B200_STREAM_DEBUG", 0) != 0;
  auto now_ms = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
  const double t_start = now_ms();
  J->fp.progress = dflag; J->fp.progress_every = every;
  const uint32_t rows = (uint32_t)(n + 1) * m;
  if (s_alpha && (rc = J->cube.alloc((size_t)rows * N))) return rc;
  if ((rc = job_launch_filters(J))) return rc;
  CU(cudaEventRecord(R.done, J->stream));
  J->fp.progress = nullptr;

  int shipped = 0; // columns [0, shipped) are on their way to the host
  auto ship = [&](int c1) -> int { // columns [shipped, c1)
    const int c0 = shipped;
    if (c1 <= c0) return 0;
    if (s_alpha) {
      const uint32_t rb = (uint32_t)c0 * m, re = (uint32_t)c1 * m;
      dim3 grid(cdiv(N, 32), cdiv(re - rb, 32));
      pf::transpose_to_cube_kernel<<<grid, dim3(32, 8), 0, R.s2>>>(J->alpha_tm.p, J->cube.p, rows, N, rb, re);
      CU(cudaGetLastError());
      CU(cudaMemcpy2DAsync(out->alpha + rb, (size_t)rows * 8, J->cube.p + rb, (size_t)rows * 8, (size_t)(re - rb) * 8, N,
                           cudaMemcpyDeviceToHost, R.s2));
    }
    if (s_w)
      CU(cudaMemcpyAsync(out->weights + (size_t)N * c0, J->weights.p + (size_t)N * c0, sizeof(double) * (size_t)N * (c1 - c0),
                         cudaMemcpyDeviceToHost, R.s2));
    if (s_i) {
      const int i1 = std::min(c1, n);
      if (i1 > c0)
        CU(cudaMemcpyAsync(out->indices + (size_t)N * c0, J->indices.p + (size_t)N * c0,
                           sizeof(uint32_t) * (size_t)N * (i1 - c0), cudaMemcpyDeviceToHost, R.s2));
    }
    if (dbg) fprintf(stderr, "[stream] %.2f ms: shipped columns [%d, %d)\n", now_ms() - t_start, c0, c1);
    shipped = c1;
    return 0;
  };
  for (;;) {
    const int p = *(volatile int*)R.flag;
    if (p > shipped) { if ((rc = ship(p))) return rc; continue; }
    const cudaError_t q = cudaEventQuery(R.done);
    if (q == cudaSuccess) break;
    if (q != cudaErrorNotReady) return fail(BSSM_B200_ERR_CUDA, std::string("filter kernel: ") + cudaGetErrorString(q));
  }
  if (dbg) fprintf(stderr, "[stream] %.2f ms: filter kernel done\n", now_ms() - t_start);
  int32_t stopped = -1;
  CU(cudaMemcpyAsync(&stopped, J->stopped.p, sizeof(int32_t), cudaMemcpyDeviceToHost, J->stream));
  CU(cudaStreamSynchronize(J->stream));
  float ms = 0;
  if (stopped >= 0) { // rare: let the plain path zero-fill what the filter never reached
    CU(cudaStreamSynchronize(R.s2));
    if ((rc = job_finish(J, &ms))) return rc;
    return bssm_b200_job_fetch(J, 0, out);
  }
  if ((rc = ship(n + 1))) return rc;
  if ((rc = job_finish(J, &ms))) return rc; // filter summary / smoother on the job's stream, overlapping the last copies
  bssm_filter_out rest = *out;
  if (s_alpha) rest.alpha = nullptr;
  if (s_w) rest.weights = nullptr;
  if (s_i) rest.indices = nullptr;
  rc = bssm_b200_job_fetch(J, 0, &rest);
  out->loglik = rest.loglik; out->stopped_at = rest.stopped_at; out->approx_loglik = rest.approx_loglik;
  out->device_ms = rest.device_ms;
  CU(cudaStreamSynchronize(R.s2));
  if (dbg) fprintf(stderr, "[stream] %.2f ms: all copies done\n", now_ms() - t_start);
  return rc;
}

static int run_fetch_destroy(bssm_b200_job* J, bssm_filter_out* out) {
  int rc = run_fetch_streaming(J, out);
  bssm_b200_job_destroy(J);
  return rc;
}

static int out_flags_of(const bssm_filter_out* out) {
  return (out->alpha ? BSSM_OUT_ALPHA : 0) | (out->weights ? BSSM_OUT_WEIGHTS : 0) | (out->indices ? BSSM_OUT_INDICES : 0);
}

int bssm_b200_bsf_nlg(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_nlg(M, nsim, seed, 1, 0, out_flags_of(out), BSSM_POST_FILTER_SUMMARY, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_bsf_smoother_nlg(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_nlg(M, nsim, seed, 1, 0, out_flags_of(out), BSSM_POST_SMOOTHER, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_ekpf_nlg(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_nlg_ekpf(M, nsim, seed, 1, 0, out_flags_of(out), BSSM_POST_FILTER_SUMMARY, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_ekpf_smoother_nlg(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_nlg_ekpf(M, nsim, seed, 1, 0, out_flags_of(out), BSSM_POST_SMOOTHER, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_psi_smoother_nlg(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, uint32_t max_iter,
                               double conv_tol, uint32_t iekf_iter, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = create_nlg_psi_job(M, nsim, seed, max_iter, conv_tol, iekf_iter, 1, 0, out_flags_of(out), BSSM_POST_SMOOTHER, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_psi_filter_nlg(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, uint32_t max_iter,
                             double conv_tol, uint32_t iekf_iter, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = create_nlg_psi_job(M, nsim, seed, max_iter, conv_tol, iekf_iter, 1, 0, out_flags_of(out), BSSM_POST_NONE, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_gaussian_psi_smoother(const bssm_ung_model* M, uint32_t nsim, uint32_t seed, double* alpha) {
  if (!M || !alpha) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (M->distribution != BSSM_DIST_GAUSSIAN || !M->H) return fail(BSSM_B200_ERR_ARG, "gaussian_psi_smoother: linear-Gaussian model with H required");
  if (M->m > 3) return fail(BSSM_B200_ERR_UNSUPPORTED, "gaussian_psi_smoother: m <= 3 in this build");
  int rc = require_device();
  if (rc) return rc;
  const int n = M->n, m = M->m;
  // smoother_ccov + conditional_cov of the model itself (src/model_ssm_ulg.cpp:646-652): host, O(n m^3)
  pfga::UngApprox ga(M);
  for (int t = 0; t < n; ++t) {
    const double h = M->H[t * M->Htv];
    ga.ay[t] = M->y[t]; ga.aH[t] = h; ga.aHH[t] = h * h;
  }
  ga.smoother_and_conditional();
  DevBuf<double> ah, Lt, Ct, tm, cube;
  const uint32_t rows = (uint32_t)(n + 1) * m;
  if ((rc = ah.upload(ga.alphahat.data(), ga.alphahat.size())) || (rc = Lt.upload(ga.L.data(), ga.L.size())) ||
      (rc = Ct.upload(ga.Ct.data(), ga.Ct.size())) || (rc = tm.alloc((size_t)rows * nsim)) ||
      (rc = cube.alloc((size_t)rows * nsim)))
    return rc;
  const dim3 g1(cdiv(nsim, 128));
  if (m == 1) pf::gaussian_psi_kernel<1><<<g1, 128>>>(ah.p, Lt.p, Ct.p, tm.p, nsim, n, seed);
  else if (m == 2) pf::gaussian_psi_kernel<2><<<g1, 128>>>(ah.p, Lt.p, Ct.p, tm.p, nsim, n, seed);
  else pf::gaussian_psi_kernel<3><<<g1, 128>>>(ah.p, Lt.p, Ct.p, tm.p, nsim, n, seed);
  CU(cudaGetLastError());
  pf::transpose_to_cube_kernel<<<dim3(cdiv(nsim, 32), cdiv(rows, 32)), dim3(32, 8)>>>(tm.p, cube.p, rows, nsim, 0u, rows);
  CU(cudaGetLastError());
  CU(cudaMemcpy(alpha, cube.p, sizeof(double) * (size_t)rows * nsim, cudaMemcpyDeviceToHost));
  return 0;
}

// ---------------------------------------------------------------- SPDK (simulation smoother + importance weights)
namespace {
// one model's inputs of pf::sim_smoother_kernel as offsets into its block of doubles
struct SpdkLayout : pf::SimBlock {
  SpdkLayout(const bssm_ung_model* M, bool xb) {
    has_xbeta = xb ? 1 : 0;
    const size_t n = M->n, m = M->m, k = M->k, mm = m * m;
    size_t o = 0;
    auto take = [&](size_t cnt) { size_t at = o; o += cnt; return at; };
    y = take(n); u = take(n); xbeta = take(xb ? n : 0);
    Z = take(m * (M->Ztv ? n : 1)); T = take(mm * (M->Ttv ? n : 1)); R = take(m * k * (M->Rtv ? n : 1));
    C = take(m * (M->Ctv ? n : 1)); D = take(M->Dtv ? n : 1); a1 = take(m); P1 = take(mm); LP1 = take(mm);
    RR = take(mm * (M->Rtv ? n : 1));
    ay = take(n); aH = take(n); rH = take(n); yphi2 = take(n); F = take(n); K = take(m * n); Lt = take(mm * n);
    ahat = take(m * (n + 1)); phi = take(1);
    total = o;
  }
};

// the model part of a block: the arrays of the model itself and the factor of P1
void spdk_fill_model(const bssm_ung_model* M, const SpdkLayout& L, double* b) {
  const int n = M->n, m = M->m, k = M->k;
  const size_t mm = (size_t)m * m;
  auto put = [&](size_t at, const double* src, size_t cnt) { std::copy(src, src + cnt, b + at); };
  put(L.y, M->y, n); put(L.u, M->u, n);
  if (L.has_xbeta) put(L.xbeta, M->xbeta, n);
  put(L.Z, M->Z, (size_t)m * (M->Ztv ? n : 1)); put(L.T, M->T, mm * (M->Ttv ? n : 1));
  put(L.R, M->R, (size_t)m * k * (M->Rtv ? n : 1)); put(L.C, M->C, (size_t)m * (M->Ctv ? n : 1));
  put(L.D, M->D, M->Dtv ? n : 1); put(L.a1, M->a1, m); put(L.P1, M->P1, mm);
  psd_chol_host(M->P1, m, b + L.LP1); // src/model_ssm_ulg.cpp:145
  b[L.phi] = M->phi;
}

// The host part for one model: approximation (mode_in == nullptr: approximate(), src/model_ssm_ung.cpp:87-140;
// else approximate_for_is at the stored mode, :143-161), scales, the smoothed mean of the pseudo-observations
// and the gains of its smoother (fast_precomputing_smoother, src/model_ssm_ulg.cpp:359-413).
void spdk_fill(const bssm_ung_model* M, const double* mode_in, const SpdkLayout& L, double* b, double* approx_ll,
               double* ssum) {
  const int n = M->n, m = M->m;
  const size_t mm = (size_t)m * m;
  pfga::UngApprox ga(M);
  ga.run_head(mode_in);
  pfga::UngApprox::SimGains sg;
  std::vector<double> ahat;
  ga.fast_smoother(ahat, &sg);
  auto put = [&](size_t at, const double* src, size_t cnt) { std::copy(src, src + cnt, b + at); };
  spdk_fill_model(M, L, b);
  put(L.RR, sg.RR.data(), sg.RR.size());
  put(L.ay, ga.ay.data(), n); put(L.aH, ga.aH.data(), n);
  for (int t = 0; t < n; ++t) {
    b[L.rH + t] = 1.0 / ga.aH[t];
    const double q = M->y[t] / M->phi;
    b[L.yphi2 + t] = q * q;
  }
  put(L.F, sg.F.data(), n); put(L.K, sg.K.data(), (size_t)m * n); put(L.Lt, sg.Lt.data(), mm * n);
  put(L.ahat, ahat.data(), (size_t)m * (n + 1));
  *approx_ll = ga.approx_loglik;
  *ssum = pfga::accu2(ga.scales.data(), (size_t)n);
}

// nf blocks (pack) -> the log-weights sum_t log_weights(t, alpha_i) of all trajectories, draw after draw
// (nsim_f[f] each); alpha_out: nf == 1 only, m x (n+1) x nsim
int spdk_launch(const bssm_ung_model* M, const SpdkLayout& L, const double* dblocks, double phi0, uint32_t nf,
                const uint32_t* nsim_f, int antithetic, uint32_t seed, uint32_t first_filter,
                std::vector<double>& logw, double* alpha_out) {
  const int n = M->n, m = M->m, k = M->k;
  const size_t rows = (size_t)(n + 1) * m;
  std::vector<uint32_t> toff(nf + 1, 0u), soff(nf + 1, 0u);
  for (uint32_t f = 0; f < nf; ++f) {
    const uint32_t ns = nsim_f[f];
    soff[f + 1] = soff[f] + ns;
    toff[f + 1] = toff[f] + (antithetic ? ns / 2 + (ns & 1u) : ns);
  }
  const size_t S = soff[nf], Tn = toff[nf];
  pf::SimParams sp{};
  sp.n = n; sp.m = m; sp.k = k; sp.distribution = M->distribution;
  sp.Ztv = M->Ztv; sp.Ttv = M->Ttv; sp.Rtv = M->Rtv; sp.Ctv = M->Ctv; sp.Dtv = M->Dtv;
  sp.seed = seed; sp.antithetic = antithetic ? 1 : 0; sp.first_filter = first_filter;
  DevBuf<double> tm, lw, v, r, cube;
  DevBuf<uint32_t> doff;
  int rc;
  if ((rc = tm.alloc(rows * S)) || (rc = lw.alloc(S)) || (rc = v.alloc((size_t)n * Tn)) ||
      (rc = r.alloc((size_t)n * m * Tn)))
    return rc;
  const double* b = dblocks;
  sp.y = b + L.y; sp.u = b + L.u; sp.xbeta = L.has_xbeta ? b + L.xbeta : nullptr;
  sp.Z = b + L.Z; sp.T = b + L.T; sp.R = b + L.R; sp.C = b + L.C; sp.D = b + L.D; sp.a1 = b + L.a1;
  sp.P1 = b + L.P1; sp.LP1 = b + L.LP1; sp.RR = b + L.RR; sp.ay = b + L.ay; sp.aH = b + L.aH; sp.rH = b + L.rH;
  sp.yphi2 = b + L.yphi2; sp.F = b + L.F; sp.K = b + L.K; sp.Lt = b + L.Lt; sp.ahat = b + L.ahat;
  sp.phi_f = b + L.phi; sp.phi = phi0;
  sp.alpha = tm.p; sp.logw = lw.p; sp.v = v.p; sp.r = r.p;
  if (nf == 1) { // one model: the scalar members, no search
    sp.nf = 0; sp.nsim = nsim_f[0]; sp.nsim2 = antithetic ? sp.nsim / 2 : sp.nsim; sp.threads = toff[1];
  } else {
    std::vector<uint32_t> both(toff);
    both.insert(both.end(), soff.begin(), soff.end());
    if ((rc = doff.upload(both.data(), both.size()))) return rc;
    sp.nf = nf; sp.fstride = L.total; sp.toff = doff.p; sp.soff = doff.p + (nf + 1);
  }
  sp.total_threads = (uint32_t)Tn;
  const dim3 grid(cdiv((uint32_t)Tn, 128));
  const int mk = m > k ? m : k;
  if (mk == 1) pf::sim_smoother_kernel<1><<<grid, 128>>>(sp);
  else if (mk <= 4) pf::sim_smoother_kernel<4><<<grid, 128>>>(sp);
  else pf::sim_smoother_kernel<16><<<grid, 128>>>(sp);
  CU(cudaGetLastError());
  logw.resize(S);
  CU(cudaMemcpy(logw.data(), lw.p, sizeof(double) * S, cudaMemcpyDeviceToHost));
  if (alpha_out && nf == 1) {
    const uint32_t nsim = nsim_f[0];
    if ((rc = cube.alloc(rows * nsim))) return rc;
    pf::transpose_to_cube_kernel<<<dim3(cdiv(nsim, 32), cdiv((uint32_t)rows, 32)), dim3(32, 8)>>>(
        tm.p, cube.p, (uint32_t)rows, nsim, 0u, (uint32_t)rows);
    CU(cudaGetLastError());
    CU(cudaMemcpy(alpha_out, cube.p, sizeof(double) * rows * nsim, cudaMemcpyDeviceToHost));
  }
  return 0;
}

int spdk_check_model(const bssm_ung_model* M, const char* who) {
  if (!M) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (M->n < 1 || M->m < 1 || M->k < 1 || M->k > M->m) return fail(BSSM_B200_ERR_ARG, "need n >= 1, 1 <= k <= m");
  if (M->distribution < 0 || M->distribution > 4)
    return fail(BSSM_B200_ERR_ARG, std::string(who) + ": non-Gaussian ssm_ung models (distribution 0..4)");
  if (!M->u || !M->initial_mode) return fail(BSSM_B200_ERR_ARG, std::string(who) + ": u and initial_mode required");
  if (M->m > 16) return fail(BSSM_B200_ERR_UNSUPPORTED, "state dimension m <= 16 in this build");
  return 0;
}
} // namespace

// importance_sample_ng for the ssm_ung family (src/R_importance_sample.cpp:9-75) and the SPDK branch of
// ssm_ung::log_likelihood (src/model_ssm_ung.cpp:199-211): the Gaussian approximation and the gains of its
// smoother on the host, once (O(n m^3)); the nsim simulated trajectories and their weights on the device.
int bssm_b200_importance_sample_ung(const bssm_ung_model* M, uint32_t nsim, int use_antithetic, uint32_t seed,
                                    double* alpha, double* weights, double* loglik) {
  int rc = spdk_check_model(M, "importance_sample");
  if (rc) return rc;
  // nsim = 1 is the reference's separate single-draw shortcut (src/model_ssm_ulg.cpp:219-241): not a device job
  if (nsim < 2) return fail(BSSM_B200_ERR_UNSUPPORTED, "importance_sample: nsim >= 2");
  if ((rc = require_device())) return rc;
  if ((size_t)(M->n + 1) * M->m * nsim > ((size_t)1 << 32)) return fail(BSSM_B200_ERR_ARG, "importance_sample: (n + 1) m nsim <= 2^32");
  const SpdkLayout L(M, M->xbeta != nullptr);
  std::vector<double> pack(L.total), w;
  double approx_ll, ssum;
  spdk_fill(M, nullptr, L, pack.data(), &approx_ll, &ssum);
  DevBuf<double> dpack;
  if ((rc = dpack.upload(pack.data(), pack.size()))) return rc;
  if ((rc = spdk_launch(M, L, dpack.p, M->phi, 1, &nsim, use_antithetic, seed, 0u, w, alpha))) return rc;
  // w - accu(scales), exp(w - max), mean (src/model_ssm_ung.cpp:206-210): nsim values, on the host
  double maxw = -std::numeric_limits<double>::infinity();
  for (uint32_t i = 0; i < nsim; ++i) { w[i] -= ssum; if (w[i] > maxw) maxw = w[i]; }
  for (uint32_t i = 0; i < nsim; ++i) w[i] = pf_exp(w[i] - maxw);
  if (weights) std::copy(w.begin(), w.end(), weights);
  if (loglik) {
    loglik[0] = approx_ll + pf_log(pfga::accu2(w.data(), nsim) / nsim) + maxw;
    loglik[1] = approx_ll;
  }
  return 0;
}

// approx_mcmc::is_correction_spdk (src/approx_mcmc.cpp:606-678) for ssm_ung models given as parameter sets
// (parset_ung::update, src/parset_ng.cpp:84-124), output_type 3: the host gathers every stored draw's model arrays
// into its block; the approximation at the stored mode (approximate_for_is) and the smoother gains follow on the
// device, one thread per draw (pf::spdk_gains_kernel), then ALL simulated trajectories of a chunk of draws in one
// launch of the simulation smoother.  weight_i = mean(exp(w - accu(scales))) (:637-639).
int bssm_b200_is_correction_spdk_ung_parset(const bssm_ung_model* M, const bssm_parset* ps, uint32_t nsim,
                                            uint32_t seed, int is_type, const uint32_t* counts, const double* modes,
                                            const double* approx_loglik, uint32_t n_stored, double* weights,
                                            double* posterior) {
  int rc = spdk_check_model(M, "is_correction_spdk");
  if (rc) return rc;
  if (!ps || !counts || !modes || !weights || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  if (ps->beta && (!ps->xreg || ps->n_xreg < 1)) return fail(BSSM_B200_ERR_ARG, "parset: beta needs xreg and n_xreg");
  if (is_type < 1 || is_type > 3) return fail(BSSM_B200_ERR_ARG, "is_type must be 1, 2 or 3");
  if (nsim < 2) return fail(BSSM_B200_ERR_UNSUPPORTED, "is_correction_spdk: nsim >= 2");
  for (uint32_t i = 0; i < n_stored; ++i)
    if (counts[i] < 1) return fail(BSSM_B200_ERR_ARG, "is_correction_spdk: counts must be >= 1");
  if ((rc = require_device())) return rc;
  const size_t n = M->n, m = M->m, k = M->k;
  // expand() for is_type 3 (src/approx_mcmc.cpp:49-88): one draw per repetition, each with its own stream
  std::vector<uint32_t> src, ns;
  for (uint32_t i = 0; i < n_stored; ++i) {
    if (is_type == 3) for (uint32_t c = 0; c < counts[i]; ++c) { src.push_back(i); ns.push_back(nsim); }
    else { src.push_back(i); ns.push_back(is_type == 1 ? nsim * counts[i] : nsim); }
  }
  const size_t nd = src.size();
  const SpdkLayout L(M, M->xbeta != nullptr || ps->beta != nullptr);
  // BSSM_B200_HOST_APPROX=1: per-draw approximation and gains in an OpenMP loop on the host (A/B tests)
  const bool host_approx = env_int("BSSM_B200_HOST_APPROX", 0) != 0;
  // chunks of draws: at most 2^31 doubles of trajectories and 2^28 doubles of model blocks per launch
  size_t d0 = 0;
  while (d0 < nd) {
    size_t d1 = d0, traj = 0;
    while (d1 < nd) {
      const bool fits = traj + ns[d1] <= ((size_t)1 << 31) / ((n + 1) * m) && (d1 - d0 + 1) * L.total <= ((size_t)1 << 28);
      if (d1 > d0 && !fits) break;
      traj += ns[d1];
      ++d1;
    }
    const uint32_t nf = (uint32_t)(d1 - d0);
    // the blocks of a chunk are tens of MB: the staging buffer is kept between calls (no first-touch faults again),
    // unless it grew beyond 256 MB -- such a buffer is handed back when the call ends
    static thread_local std::vector<double> pack;
    struct Trim { std::vector<double>& v; ~Trim() { if (v.capacity() > ((size_t)1 << 25)) std::vector<double>().swap(v); } } trim{pack};
    if (pack.size() < L.total * nf) pack.resize(L.total * nf);
    double* const blocks = pack.data(); // (`pack` itself is thread-local: the OpenMP workers below must not name it)
    std::vector<double> al(nf), ss(nf), w, md(host_approx ? 0 : n * nf); // al: 0 by definition of approximate_for_is
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < (int64_t)nf; ++f) {
      const uint32_t c = src[d0 + f];
      bssm_ung_model Mi = *M; // parset_ung::update(model, c)
      std::vector<double> xb;
      if (ps->phi) Mi.phi = ps->phi[c];
      if (ps->Z) Mi.Z = ps->Z + m * (M->Ztv ? n : 1) * c;
      if (ps->T) Mi.T = ps->T + m * m * (M->Ttv ? n : 1) * c;
      if (ps->R) Mi.R = ps->R + m * k * (M->Rtv ? n : 1) * c;
      if (ps->a1) Mi.a1 = ps->a1 + m * c;
      if (ps->P1) Mi.P1 = ps->P1 + m * m * c;
      if (ps->D) Mi.D = ps->D + (M->Dtv ? n : 1) * c;
      if (ps->C) Mi.C = ps->C + m * (M->Ctv ? n : 1) * c;
      if (ps->beta) { // compute_xbeta, src/model_ssm_ung.h:83
        xb.assign(n, 0.0);
        for (size_t t = 0; t < n; ++t) {
          double sum = 0.0;
          for (int j = 0; j < ps->n_xreg; ++j) sum += ps->xreg[t + n * j] * ps->beta[(size_t)ps->n_xreg * c + j];
          xb[t] = sum;
        }
        Mi.xbeta = xb.data();
      }
      if (host_approx) {
        spdk_fill(&Mi, modes + n * c, L, blocks + L.total * f, &al[f], &ss[f]);
      } else { // the model arrays only: approximation and gains follow on the device, one thread per draw
        spdk_fill_model(&Mi, L, blocks + L.total * f);
        std::copy(modes + n * c, modes + n * (c + 1), &md[n * f]);
      }
    }
    DevBuf<double> dpack;
    if ((rc = dpack.upload(blocks, L.total * nf))) return rc;
    if (!host_approx) {
      DevBuf<double> dmodes, dss, scratch;
      if ((rc = dmodes.upload(md.data(), md.size())) || (rc = dss.alloc(nf)) || (rc = scratch.alloc((m + 2) * n * nf)))
        return rc;
      pf::SpdkGainParams gp{};
      gp.n = M->n; gp.m = M->m; gp.k = M->k; gp.distribution = M->distribution;
      gp.Ztv = M->Ztv; gp.Ttv = M->Ttv; gp.Rtv = M->Rtv; gp.Dtv = M->Dtv; gp.Ctv = M->Ctv;
      gp.L = L; gp.blocks = dpack.p; gp.modes = dmodes.p; gp.ssum = dss.p; gp.scratch = scratch.p;
      const dim3 grid(cdiv(nf, 32));
      if (m == 1) pf::spdk_gains_kernel<1, true><<<grid, 32>>>(gp, nf);
      else if (m == 2) pf::spdk_gains_kernel<2, true><<<grid, 32>>>(gp, nf);
      else if (m == 3) pf::spdk_gains_kernel<3, true><<<grid, 32>>>(gp, nf);
      else pf::spdk_gains_kernel<16, false><<<grid, 32>>>(gp, nf);
      CU(cudaGetLastError());
      CU(cudaMemcpy(ss.data(), dss.p, sizeof(double) * nf, cudaMemcpyDeviceToHost));
    }
    if ((rc = spdk_launch(M, L, dpack.p, blocks[L.phi], nf, &ns[d0], 1, seed, (uint32_t)d0, w, nullptr))) return rc;
    size_t o = 0;
    for (uint32_t f = 0; f < nf; ++f) {
      const uint32_t cnt = ns[d0 + f];
      for (uint32_t i = 0; i < cnt; ++i) w[o + i] = pf_exp(w[o + i] - ss[f]);
      weights[d0 + f] = pfga::accu2(&w[o], cnt) / cnt;
      o += cnt;
    }
    d0 = d1;
  }
  if (posterior)
    for (size_t i = 0; i < nd; ++i) posterior[i] = (approx_loglik ? approx_loglik[src[i]] : 0.0) + pf_log(weights[i]);
  return 0;
}

int bssm_b200_bsf_mng(const bssm_mng_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_mng(M, nsim, seed, 1, 0, out_flags_of(out), BSSM_POST_FILTER_SUMMARY, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_bsf_smoother_mng(const bssm_mng_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_mng(M, nsim, seed, 1, 0, out_flags_of(out), BSSM_POST_SMOOTHER, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_psi_smoother_mng(const bssm_mng_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_mng_psi(M, nsim, seed, 1, 0, out_flags_of(out), BSSM_POST_SMOOTHER, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_psi_filter_mng(const bssm_mng_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_mng_psi(M, nsim, seed, 1, 0, out_flags_of(out), BSSM_POST_NONE, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_bsf_sde(const bssm_sde_model* M, uint32_t nsim, uint32_t L, uint32_t seed, bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_sde(M, nsim, L, seed, 1, 0, out_flags_of(out), BSSM_POST_FILTER_SUMMARY, &J);
  if (rc) return rc;
  const int n = M->n;
  rc = run_fetch_destroy(J, out);
  // the reference allocates att (1, n+1) and Ptt (1, 1, n+1) here and filter_summary
  // leaves their last column zero (src/R_sde.cpp:56-60)
  if (!rc && out->att) out->att[n] = 0.0;
  if (!rc && out->Ptt) out->Ptt[n] = 0.0;
  return rc;
}
int bssm_b200_bsf_smoother_sde(const bssm_sde_model* M, uint32_t nsim, uint32_t L, uint32_t seed,
                               bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_sde(M, nsim, L, seed, 1, 0, out_flags_of(out), BSSM_POST_SMOOTHER, &J);
  return rc ? rc : run_fetch_destroy(J, out);
}
int bssm_b200_loglik_sde(const bssm_sde_model* M, uint32_t nsim, uint32_t L, uint32_t seed, double* loglik) {
  if (!loglik) return fail(BSSM_B200_ERR_ARG, "null loglik");
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_sde(M, nsim, L, seed, 1, 0, 0, BSSM_POST_NONE, &J);
  if (rc) return rc;
  float ms = 0;
  rc = bssm_b200_job_run(J, &ms);
  if (!rc) rc = bssm_b200_job_logliks(J, loglik);
  bssm_b200_job_destroy(J);
  return rc;
}

// --------------------------------------------------------- post-correction
namespace {

// update_model(theta) restricted to the scalars the m = 1 psi-APF reads:
// par = {T, R, a1, P1, C, phi}.  svm: src/model_svm.cpp:12-29; ar1_ng:
// src/model_ar1_ng.cpp:12-36 (sigma and phi are on the log scale there).
int theta_to_par(const bssm_ung_model* M, const bssm_theta_map* map, const double* th, uint32_t n_par,
                 double par[6]) {
  double T = M->T[0], R = M->R[0], a1 = M->a1[0], P1 = M->P1[0], C = M->C[0], phi = M->phi;
  if (map->model_type == 3) {
    if (n_par < 3) return fail(BSSM_B200_ERR_ARG, "svm needs 3 parameters");
    if (map->svm_type == 0) {
      phi = th[2];
    } else {
      a1 = th[2];
      C = th[2] * (1.0 - th[0]);
    }
    T = th[0];
    R = th[1];
    P1 = th[1] * th[1] / (1 - th[0] * th[0]);
  } else if (map->model_type == 4) {
    if (n_par < (uint32_t)(2 + map->mu_est + map->phi_est)) return fail(BSSM_B200_ERR_ARG, "ar1_ng: too few parameters");
    const double rho = th[0];
    const double sigma = pf_exp(th[1]);
    T = rho;
    R = sigma;
    const double RR = sigma * sigma;
    if (map->mu_est) {
      a1 = th[2];
      C = th[2] * (1.0 - rho);
    }
    P1 = RR / (1.0 - rho * rho);
    if (map->phi_est) phi = pf_exp(th[2 + map->mu_est]);
  } else if (map->model_type == 2) {
    // bsm_ng with a level only (m = 1), src/model_bsm_ng.cpp:17-45
    if (n_par < (uint32_t)(map->level_est + map->phi_est)) return fail(BSSM_B200_ERR_ARG, "bsm_ng: too few parameters");
    if (map->level_est) R = pf_exp(th[0]);
    if (map->phi_est) phi = pf_exp(th[map->level_est]);
  } else {
    return fail(BSSM_B200_ERR_UNSUPPORTED, "post-correction: model_type must be 2 (bsm_ng), 3 (svm) or 4 (ar1_ng) in this build");
  }
  par[0] = T; par[1] = R; par[2] = a1; par[3] = P1; par[4] = C; par[5] = phi;
  return 0;
}

// bsm_ng::update_model (src/model_bsm_ng.cpp:17-51) for any m: R, P1 (m x k, m x m copies of the
// model's), phi and xbeta (n, when the model has regressors) of one draw
int theta_to_bsm(const bssm_ung_model* M, const bssm_theta_map* map, const double* th, uint32_t n_par,
                 double* R, double* P1, double* phi, double* xbeta) {
  const int m = M->m, k = M->k, n = M->n;
  const int need = map->level_est + map->slope_est + map->seasonal_est + map->noise + map->phi_est + map->n_xreg;
  if ((int)n_par < need) return fail(BSSM_B200_ERR_ARG, "bsm_ng: too few parameters");
  if (1 + map->slope + map->seasonal + map->noise > k || 1 + map->slope > m)
    return fail(BSSM_B200_ERR_ARG, "bsm_ng: component flags do not fit the model's R");
  std::copy(M->R, M->R + (size_t)m * k, R);
  std::copy(M->P1, M->P1 + (size_t)m * m, P1);
  if (map->level_est) R[0] = pf_exp(th[0]);
  if (map->slope_est) R[1 + (size_t)m * 1] = pf_exp(th[map->level_est]);
  if (map->seasonal_est) {
    const int r = 1 + map->slope;
    R[r + (size_t)m * r] = pf_exp(th[map->level_est + map->slope_est]);
  }
  if (map->noise) {
    const int c = 1 + map->slope + map->seasonal;
    const double sd = pf_exp(th[map->level_est + map->slope_est + map->seasonal_est]);
    R[(m - 1) + (size_t)m * c] = sd;
    P1[(m - 1) + (size_t)m * (m - 1)] = sd * sd;
  }
  *phi = map->phi_est ? pf_exp(th[map->level_est + map->slope_est + map->seasonal_est + map->noise]) : M->phi;
  if (map->n_xreg > 0) { // compute_xbeta, src/model_ssm_ung.h:83
    const double* beta = th + (n_par - map->n_xreg);
    for (int t = 0; t < n; ++t) {
      double sum = 0.0;
      for (int j = 0; j < map->n_xreg; ++j) sum += map->xreg[t + (size_t)n * j] * beta[j];
      xbeta[t] = sum;
    }
  }
  return 0;
}

// compute_xbeta (src/model_ssm_ung.h:83) of one draw: beta = the last n_xreg entries of theta
void theta_to_xbeta(const bssm_theta_map* map, int n, const double* th, uint32_t n_par, double* xbeta) {
  const double* beta = th + (n_par - map->n_xreg);
  for (int t = 0; t < n; ++t) {
    double sum = 0.0;
    for (int j = 0; j < map->n_xreg; ++j) sum += map->xreg[t + (size_t)n * j] * beta[j];
    xbeta[t] = sum;
  }
}

struct DrawList {
  std::vector<uint32_t> src;   // expanded draw -> column of theta / modes
  std::vector<uint32_t> count; // effective count of the expanded draw
};

int build_draws(int is_type, const uint32_t* counts, uint32_t n_stored, DrawList* dl) {
  if (is_type < 1 || is_type > 3) return fail(BSSM_B200_ERR_ARG, "is_type must be 1, 2 or 3");
  dl->src.clear(); dl->count.clear();
  for (uint32_t i = 0; i < n_stored; ++i) {
    if (is_type == 3) { // expand(), src/approx_mcmc.cpp:49-88
      for (uint32_t c = 0; c < counts[i]; ++c) { dl->src.push_back(i); dl->count.push_back(1u); }
    } else {
      dl->src.push_back(i); dl->count.push_back(counts[i]);
    }
  }
  return 0;
}

int check_postcorr_args(const bssm_ung_model* M, const bssm_theta_map* map, const uint32_t* counts,
                        const double* theta, const double* modes, uint32_t n_stored) {
  if (!M || !map || !counts || !theta || !modes || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  const bool general = map->model_type == 2 && M->m > 1;
  if (M->m != 1 && !general)
    return fail(BSSM_B200_ERR_UNSUPPORTED, "post-correction with m > 1 is implemented for bsm_ng (model_type 2) in this build");
  if (general && (M->m < 2 || M->m > 16))
    return fail(BSSM_B200_ERR_UNSUPPORTED, "post-correction for bsm_ng: m <= 16 in this build");
  if (map->n_xreg > 0 && !map->xreg) return fail(BSSM_B200_ERR_ARG, "n_xreg > 0 needs xreg");
  if (M->xbeta && map->n_xreg <= 0) return fail(BSSM_B200_ERR_ARG, "the model has regressors: theta_map needs n_xreg and xreg");
  if (map->n_xreg > 0 && map->model_type == 3) return fail(BSSM_B200_ERR_ARG, "svm has no regression part");
  return 0;
}

// gathers the per-draw inputs of draws [d0, d1) and creates their job
int make_draw_job(const bssm_ung_model* M, const bssm_theta_map* map, const DrawList& dl, uint32_t d0, uint32_t d1,
                  uint32_t N, uint32_t seed, const double* theta, uint32_t n_par, const double* modes,
                  int out_flags, int post, bssm_b200_job** job) {
  const uint32_t nf = d1 - d0;
  const size_t n = (size_t)M->n;
  if (map->model_type == 2 && M->m > 1) {
    const size_t mk = (size_t)M->m * M->k, mm = (size_t)M->m * M->m;
    DrawModels dm;
    dm.R.resize(mk * nf); dm.P1.resize(mm * nf); dm.phi.resize(nf); dm.modes.resize(n * nf);
    if (map->n_xreg > 0) dm.xbeta.resize(n * nf);
    for (uint32_t f = 0; f < nf; ++f) {
      const uint32_t c = dl.src[d0 + f];
      int rc = theta_to_bsm(M, map, theta + (size_t)c * n_par, n_par, &dm.R[mk * f], &dm.P1[mm * f], &dm.phi[f],
                            map->n_xreg > 0 ? &dm.xbeta[n * f] : nullptr);
      if (rc) return rc;
      memcpy(&dm.modes[n * f], modes + n * c, n * sizeof(double));
    }
    return create_ung_job(M, BSSM_METHOD_PSI, N, seed, nf, d0, out_flags, post, nullptr, nullptr, job, &dm);
  }
  std::vector<double> par((size_t)6 * nf), md(n * nf);
  DrawModels dm; // m = 1: only the regression part of each draw
  if (map->n_xreg > 0) {
    if ((int)n_par < map->n_xreg) return fail(BSSM_B200_ERR_ARG, "too few parameters for n_xreg");
    dm.xbeta.resize(n * nf);
  }
  for (uint32_t f = 0; f < nf; ++f) {
    const uint32_t c = dl.src[d0 + f];
    int rc = theta_to_par(M, map, theta + (size_t)c * n_par, n_par, &par[(size_t)6 * f]);
    if (rc) return rc;
    if (map->n_xreg > 0) theta_to_xbeta(map, (int)n, theta + (size_t)c * n_par, n_par, &dm.xbeta[n * f]);
    memcpy(&md[n * f], modes + n * c, n * sizeof(double));
  }
  return create_ung_job(M, BSSM_METHOD_PSI, N, seed, nf, d0, out_flags, post, par.data(), md.data(), job,
                        map->n_xreg > 0 ? &dm : nullptr);
}

} // namespace

int bssm_b200_approximate_mode(const bssm_ung_model* M, double* mode) {
  if (!M || !mode || !M->initial_mode) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (M->m != 1) return fail(BSSM_B200_ERR_UNSUPPORTED, "approximate_mode: m = 1 only in this build");
  if (M->max_iter < 1) return fail(BSSM_B200_ERR_ARG, "max_iter must be >= 1");
  std::vector<double> md, lin;
  ung_approximate_m1(M, md, lin);
  memcpy(mode, md.data(), sizeof(double) * M->n);
  return 0;
}

// A resident job inside an MCMC loop: new theta, same data and buffers (pm_mcmc / da_mcmc call the
// particle filter once per iteration at a new theta, src/mcmc.cpp:338-491).  update_model(theta) +
// approximate() on the host (O(n) per Newton sweep), then only the mode (n doubles) and the model
// scalars cross PCIe; the next run() redoes the device pre-pass and the filters.
int bssm_b200_job_set_theta(bssm_b200_job* J, const bssm_ung_model* M, const bssm_theta_map* map,
                            const double* theta, uint32_t n_par) {
  if (!J || !M || !map || !theta) return fail(BSSM_B200_ERR_ARG, "null argument");
  if (J->kind != KIND_UNG || J->m != 1 || M->m != 1 || M->n != J->n)
    return fail(BSSM_B200_ERR_UNSUPPORTED, "job_set_theta: m = 1 ssm_ung-family jobs created from the same data");
  if (M->xbeta && map->n_xreg <= 0) return fail(BSSM_B200_ERR_ARG, "the model has regressors: theta_map needs n_xreg and xreg");
  if (J->xbeta_per_filter) return fail(BSSM_B200_ERR_ARG, "job_set_theta: post-correction jobs carry one theta per filter already");
  double par[6];
  int rc = theta_to_par(M, map, theta, n_par, par);
  if (rc) return rc;
  bssm_ung_model Mi = *M;
  Mi.T = &par[0]; Mi.R = &par[1]; Mi.a1 = &par[2]; Mi.P1 = &par[3]; Mi.C = &par[4]; Mi.phi = par[5];
  Mi.Ttv = 0; Mi.Rtv = 0; Mi.Ctv = 0;
  std::vector<double> xb;
  if (map->n_xreg > 0) { // beta = the tail of theta: new regression part, shared by the job's replicates
    if (!map->xreg || (int)n_par < map->n_xreg || map->model_type == 3)
      return fail(BSSM_B200_ERR_ARG, "n_xreg > 0 needs xreg, beta in theta and a model with a regression part");
    xb.resize(J->n);
    theta_to_xbeta(map, J->n, theta, n_par, xb.data());
    Mi.xbeta = xb.data();
    if ((rc = J->xbeta.upload(xb.data(), xb.size()))) return rc;
    J->up.xbeta = J->xbeta.p; J->pp.xbeta = J->xbeta.p; J->pp.has_xbeta = 1;
  }
  const uint32_t nf = J->n_filters;
  if (J->method == BSSM_METHOD_PSI) {
    if (J->pp.draw_par && J->pp.mode_stride != 0)
      return fail(BSSM_B200_ERR_ARG, "job_set_theta: post-correction jobs carry one theta per filter already");
    if (!M->initial_mode || M->max_iter < 1) return fail(BSSM_B200_ERR_ARG, "psi-APF needs initial_mode and max_iter >= 1");
    if (J->host_prepass) return host_prepass_upload(J, &Mi); // update_model(theta), approximate(), pre-pass: all on the host
    std::vector<double> mode, mode_lin, dp((size_t)6 * nf);
    ung_approximate_m1(&Mi, mode, mode_lin);
    for (uint32_t f = 0; f < nf; ++f) memcpy(&dp[(size_t)6 * f], par, sizeof(par));
    if ((rc = J->mode.upload(mode.data(), J->n)) || (rc = J->mode_lin.upload(mode_lin.data(), J->n)) ||
        (rc = J->draw_par.upload(dp.data(), dp.size())))
      return rc;
    pf::PrepassParams& pp = J->pp;
    pp.draw_par = J->draw_par.p;
    pp.modes = J->mode.p; pp.modes_lin = J->mode_lin.p; pp.mode_stride = 0;
    pp.compute_loglik = 1;
    pp.dist_const = ung_dist_const(&Mi);
  } else {
    std::vector<double> bp((size_t)7 * nf);
    for (uint32_t f = 0; f < nf; ++f) {
      double* b = &bp[(size_t)7 * f];
      b[0] = par[0]; b[1] = par[1]; b[2] = par[2]; b[3] = par[3] > 0.0 ? std::sqrt(par[3]) : 0.0;
      b[4] = par[4]; b[5] = par[5]; b[6] = ung_dist_const(&Mi);
    }
    if ((rc = J->draw_par.upload(bp.data(), bp.size()))) return rc;
    J->up.bsf_par = J->draw_par.p;
  }
  return 0;
}

int bssm_b200_job_create_postcorr(const bssm_ung_model* M, const bssm_theta_map* map, uint32_t nsim,
                                  uint32_t seed, int is_type, const uint32_t* counts, const double* theta,
                                  uint32_t n_par, const double* modes, uint32_t n_stored,
                                  uint32_t draw_begin, uint32_t draw_end, bssm_b200_job** job) {
  int rc = check_postcorr_args(M, map, counts, theta, modes, n_stored);
  if (rc) return rc;
  if (is_type == 1) return fail(BSSM_B200_ERR_ARG, "resident post-correction jobs need equal N per draw (is_type 2 or 3)");
  DrawList dl;
  if ((rc = build_draws(is_type, counts, n_stored, &dl))) return rc;
  if (draw_begin == 0 && draw_end == 0) draw_end = (uint32_t)dl.src.size();
  if (draw_begin >= draw_end || draw_end > dl.src.size()) return fail(BSSM_B200_ERR_ARG, "bad draw range");
  return make_draw_job(M, map, dl, draw_begin, draw_end, nsim, seed, theta, n_par, modes, 0, BSSM_POST_NONE, job);
}

} // extern "C"

// The draw loop of approx_mcmc::is_correction_psi (src/approx_mcmc.cpp:353-427) for any model
// kind: `make_job(d0, d1, N, post, &job)` builds the resident batch of draws [d0, d1).
template <class MakeJob>
static int postcorr_driver(int n, int m, int output_type, uint32_t nsim, int is_type, const uint32_t* counts,
                           uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end, bssm_postcorr_out* out,
                           MakeJob&& make_job, const double* approx_ll = nullptr, bool posterior_adds_approx = false) {
  // approx_ll (per stored draw): is_correction_bsf weights exp(loglik - approx_loglik), src/approx_mcmc.cpp:478
  int rc;
  if (!out || !out->weights || !out->posterior) return fail(BSSM_B200_ERR_ARG, "weights and posterior outputs are required");
  if (output_type < 1 || output_type > 3) return fail(BSSM_B200_ERR_ARG, "output_type must be 1, 2 or 3");
  if (output_type == 1 && !out->alpha) return fail(BSSM_B200_ERR_ARG, "output_type 1 needs out->alpha");
  if (output_type == 2 && (!out->alphahat || !out->Vt)) return fail(BSSM_B200_ERR_ARG, "output_type 2 needs out->alphahat and out->Vt");
  if ((rc = require_device())) return rc;
  DrawList dl;
  if ((rc = build_draws(is_type, counts, n_stored, &dl))) return rc;
  if (draw_begin == 0 && draw_end == 0) draw_end = (uint32_t)dl.src.size();
  if (draw_begin >= draw_end || draw_end > dl.src.size()) return fail(BSSM_B200_ERR_ARG, "bad draw range");
  const size_t n1 = (size_t)n + 1;
  out->device_ms = 0.f; out->launches = 0; out->sum_w = 0.0;

  const int post = output_type == 1 ? BSSM_POST_PICK : (output_type == 2 ? BSSM_POST_SMOOTHER_MERGE : BSSM_POST_NONE);
  const double budget = 1e9 * (double)env_int("BSSM_B200_CHUNK_GB", 8);
  // running state of the weighted merge (output_type 2)
  DevBuf<double> d_sumw, d_ah, d_Vt, d_Va;
  DevBuf<uint32_t> d_cnt;
  if (output_type == 2) {
    // d_ah: two copies for the run-time-m merge (m > 3), which reads one and writes the other
    if ((rc = d_sumw.alloc(2)) || (rc = d_ah.alloc((size_t)2 * m * n1)) || (rc = d_Vt.alloc((size_t)m * m * n1)) ||
        (rc = d_Va.alloc((size_t)m * m * n1)))
      return rc;
    CU(cudaMemset(d_sumw.p, 0, 2 * sizeof(double)));
    CU(cudaMemset(d_ah.p, 0, sizeof(double) * 2 * m * n1));
    CU(cudaMemset(d_Vt.p, 0, sizeof(double) * m * m * n1));
    CU(cudaMemset(d_Va.p, 0, sizeof(double) * m * m * n1));
  }
  int chunk_no = 0;
  uint32_t d0 = draw_begin;
  while (d0 < draw_end) {
    // a chunk = consecutive draws with the same particle count (is_type 1: nsim * count)
    const uint32_t N = is_type == 1 ? nsim * dl.count[d0] : nsim;
    double per_filter = 8.0 * n1 * (2.0 * m * m + 2.0 * m + 8.0); // approximation per draw: alphahat, L, C, pseudo-observations, scratch
    if (post != BSSM_POST_NONE) per_filter += (double)N * n1 * m * 8.0 * (output_type == 2 ? 2 : 1) + (double)N * n * 4.0;
    else per_filter += (double)N * (2 * m * 8.0 + 4.0 + 8.0);
    uint32_t maxc = (uint32_t)std::max(1.0, budget / per_filter);
    uint32_t d1 = d0 + 1;
    while (d1 < draw_end && d1 - d0 < maxc && (is_type != 1 || dl.count[d1] == dl.count[d0])) ++d1;
    const uint32_t nf = d1 - d0;
    bssm_b200_job* J = nullptr;
    if ((rc = make_job(dl, d0, d1, N, post, &J))) return rc;
    struct JG { bssm_b200_job* j; ~JG() { delete j; } } jg{J};
    float ms = 0;
    if ((rc = bssm_b200_job_run(J, &ms))) return rc;
    out->device_ms += ms;
    out->launches += J->launches;
    std::vector<double> ll(nf);
    CU(cudaMemcpy(ll.data(), J->loglik.p, sizeof(double) * nf, cudaMemcpyDeviceToHost));
    for (uint32_t f = 0; f < nf; ++f) {
      const size_t o = (size_t)(d0 - draw_begin) + f;
      const double off = approx_ll ? approx_ll[dl.src[d0 + f]] : 0.0;
      if (out->loglik) out->loglik[o] = ll[f];
      ll[f] = ll[f] - off;
      const double w = pf_exp(ll[f]);         // weight_storage(i) = exp(loglik [- approx_loglik]), :385, :478
      out->weights[o] = w;
      // prior (0: the caller adds it) + approx_loglik + log(weight), :425-426, :520, :590
      out->posterior[o] = 0.0 + (posterior_adds_approx ? off : 0.0) + pf_log(w);
    }
    if (approx_ll && output_type == 2) // the merge below weights the draws by exp(J->loglik)
      CU(cudaMemcpyAsync(J->loglik.p, ll.data(), sizeof(double) * nf, cudaMemcpyHostToDevice, J->stream));
    if (output_type == 1) {
      CU(cudaMemcpy(out->alpha + (size_t)(d0 - draw_begin) * n1 * m, J->traj.p, sizeof(double) * nf * n1 * m, cudaMemcpyDeviceToHost));
    } else if (output_type == 2) {
      if ((rc = d_cnt.upload(&dl.count[d0], nf))) return rc;
      cudaEvent_t e0, e1;
      CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
      CU(cudaEventRecord(e0, J->stream));
      double* sw_in = d_sumw.p + (chunk_no & 1);
      double* sw_out = d_sumw.p + ((chunk_no + 1) & 1);
      CU(cudaMemcpyAsync(sw_out, sw_in, sizeof(double), cudaMemcpyDeviceToDevice, J->stream));
      if (m == 1)
        pf::postcorr_merge_kernel<1><<<cdiv((uint32_t)n1, 128), 128, 0, J->stream>>>(
            J->s_mean.p, J->s_cov.p, J->loglik.p, d_cnt.p, nf, (int)n1, sw_in, sw_out, d_ah.p, d_Vt.p, d_Va.p);
      else if (m == 2)
        pf::postcorr_merge_kernel<2><<<cdiv((uint32_t)n1, 128), 128, 0, J->stream>>>(
            J->s_mean.p, J->s_cov.p, J->loglik.p, d_cnt.p, nf, (int)n1, sw_in, sw_out, d_ah.p, d_Vt.p, d_Va.p);
      else if (m == 3)
        pf::postcorr_merge_kernel<3><<<cdiv((uint32_t)n1, 128), 128, 0, J->stream>>>(
            J->s_mean.p, J->s_cov.p, J->loglik.p, d_cnt.p, nf, (int)n1, sw_in, sw_out, d_ah.p, d_Vt.p, d_Va.p);
      else { // ping-pong over the chunks: chunk c reads copy c & 1, writes the other
        const size_t total = (size_t)m * m * n1;
        pf::postcorr_merge_dyn_kernel<<<(unsigned)((total + 127) / 128), 128, 0, J->stream>>>(
            J->s_mean.p, J->s_cov.p, J->loglik.p, d_cnt.p, nf, (int)n1, m, sw_in, sw_out,
            d_ah.p + (size_t)(chunk_no & 1) * m * n1, d_ah.p + (size_t)((chunk_no + 1) & 1) * m * n1, d_Vt.p, d_Va.p);
      }
      CU(cudaGetLastError());
      CU(cudaEventRecord(e1, J->stream));
      CU(cudaStreamSynchronize(J->stream));
      float mms = 0;
      CU(cudaEventElapsedTime(&mms, e0, e1));
      cudaEventDestroy(e0); cudaEventDestroy(e1);
      out->device_ms += mms;
      out->launches += 1;
    }
    d0 = d1;
    ++chunk_no;
  }
  if (output_type == 2) {
    double sw = 0;
    CU(cudaMemcpy(&sw, d_sumw.p + (chunk_no & 1), sizeof(double), cudaMemcpyDeviceToHost));
    out->sum_w = sw;
    std::vector<double> ah((size_t)m * n1), Vt((size_t)m * m * n1), Va((size_t)m * m * n1);
    // m <= 3 updates copy 0 in place; the run-time-m merge left its result in copy (number of chunks) & 1
    CU(cudaMemcpy(ah.data(), d_ah.p + (m > 3 ? (size_t)(chunk_no & 1) * m * n1 : 0), sizeof(double) * ah.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(Vt.data(), d_Vt.p, sizeof(double) * Vt.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(Va.data(), d_Va.p, sizeof(double) * Va.size(), cudaMemcpyDeviceToHost));
    // Vt += Valpha / sum_w (:422-424); alphahat is returned transposed, (n+1) x m
    for (size_t e = 0; e < Vt.size(); ++e) out->Vt[e] = Vt[e] + Va[e] / sw;
    for (size_t t = 0; t < n1; ++t)
      for (int d = 0; d < m; ++d) out->alphahat[t + n1 * d] = ah[d + (size_t)m * t];
  }
  return 0;
}



// ------------------------------------------------- multi-GPU post-correction
// The reference fans the draws out over OpenMP threads inside ONE call (n_threads,
// src/approx_mcmc.cpp:364-369).  Here n_threads is the number of GPUs: one host thread per device,
// contiguous static shards of the draw list (schedule(static)), each shard run by postcorr_driver on
// its device, then ONE ncclAllGather of the per-draw log-likelihoods over a communicator that is
// created once per process (ncclCommInitAll) and kept.  NCCL is loaded with dlopen so that the
// library has no link-time dependency on it; every rank ends up with all log-likelihoods, rank 0's
// copy becomes the weights / posterior of the call.  Draw d uses random stream d whatever the
// sharding, so the result is bit-identical to the one-GPU call.
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::vector<ncclComm_t> comms;
  std::vector<int> devs;
  std::mutex mu;
  bool tried = false;
  std::string why;
  bool load() {
    if (tried) return h != nullptr;
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) { why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
    CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !AllGather || !GetErrorString) { why = "libnccl: missing symbols"; h = nullptr; return false; }
    return true;
  }
  // persistent communicator over `d` (re-created only when the device set changes)
  bool ensure(const std::vector<int>& d) {
    std::lock_guard<std::mutex> g(mu);
    if (!load()) return false;
    if (d == devs && !comms.empty()) return true;
    release_locked();
    comms.assign(d.size(), nullptr);
    ncclResult_t r = CommInitAll(comms.data(), (int)d.size(), d.data());
    if (r != ncclSuccess) { why = std::string("ncclCommInitAll: ") + GetErrorString(r); comms.clear(); return false; }
    devs = d;
    return true;
  }
  void release_locked() {
    for (auto c : comms) if (c) CommDestroy(c);
    comms.clear(); devs.clear();
  }
  void release() { std::lock_guard<std::mutex> g(mu); if (h) release_locked(); }
};
NcclApi g_nccl;
thread_local std::string g_gather_note; // how the last multi-GPU call gathered: "nccl: G ranks" / "host: <why>" / ""

template <class MakeJob>
static int postcorr_multi(uint32_t n_threads, int n, int m, int output_type, uint32_t nsim, int is_type,
                          const uint32_t* counts, uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                          bssm_postcorr_out* out, MakeJob&& make_job, const double* approx_ll = nullptr,
                          bool posterior_adds_approx = false) {
  g_gather_note.clear();
  int rc = require_device();
  if (rc) return rc;
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  int G = (int)std::min<uint32_t>(n_threads ? n_threads : 1u, (uint32_t)ndev);
  G = std::min(G, std::max(1, env_int("BSSM_B200_MAX_GPUS", 64)));
  DrawList dl;
  if ((rc = build_draws(is_type, counts, n_stored, &dl))) return rc;
  uint32_t b = draw_begin, e = draw_end;
  if (b == 0 && e == 0) e = (uint32_t)dl.src.size();
  if (b >= e || e > dl.src.size()) return fail(BSSM_B200_ERR_ARG, "bad draw range");
  const uint32_t len = e - b;
  // the summary merge (output_type 2) is a sequential weighted recursion over the draws: one device
  if (G <= 1 || output_type == 2 || len < (uint32_t)(2 * G) || !out || !out->weights || !out->posterior)
    return postcorr_driver(n, m, output_type, nsim, is_type, counts, n_stored, draw_begin, draw_end, out, make_job,
                           approx_ll, posterior_adds_approx);
  if (output_type == 1 && !out->alpha) return fail(BSSM_B200_ERR_ARG, "output_type 1 needs out->alpha");
  int dev0 = 0;
  CU(cudaGetDevice(&dev0));
  std::vector<int> devs(G);
  for (int g = 0; g < G; ++g) devs[g] = (dev0 + g) % ndev;
  const bool use_nccl = !env_int("BSSM_B200_NO_NCCL", 0) && g_nccl.ensure(devs);
  const size_t n1 = (size_t)n + 1;
  std::vector<uint32_t> sb(G + 1);
  for (int g = 0; g <= G; ++g) sb[g] = b + (uint32_t)(((uint64_t)len * g) / G); // contiguous static blocks
  uint32_t maxlen = 0;
  for (int g = 0; g < G; ++g) maxlen = std::max(maxlen, sb[g + 1] - sb[g]);
  std::vector<std::vector<double>> ll(G), wg(G), pg(G);
  std::vector<std::vector<double>> gathered(G);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  std::vector<float> ms(G, 0.f);
  std::vector<int> launches(G, 0);
  auto worker = [&](int g) {
    if (cudaSetDevice(devs[g]) != cudaSuccess) { rcs[g] = BSSM_B200_ERR_CUDA; errs[g] = "cudaSetDevice failed"; }
    const uint32_t lg = sb[g + 1] - sb[g];
    ll[g].assign(lg, 0.0); wg[g].assign(lg, 0.0); pg[g].assign(lg, 0.0);
    bssm_postcorr_out og;
    memset(&og, 0, sizeof(og));
    og.weights = wg[g].data(); og.posterior = pg[g].data(); og.loglik = ll[g].data();
    if (output_type == 1) og.alpha = out->alpha + (size_t)(sb[g] - b) * n1 * m; // disjoint slices of the caller's buffer
    if (!rcs[g]) {
      rcs[g] = postcorr_driver(n, m, output_type, nsim, is_type, counts, n_stored, sb[g], sb[g + 1], &og, make_job,
                               approx_ll, posterior_adds_approx);
      if (rcs[g]) errs[g] = g_err;
    }
    ms[g] = og.device_ms; launches[g] = og.launches;
    if (use_nccl) {
      // every rank takes part in the collective even if its shard failed (a missing rank would hang the others)
      DevBuf<double> send, recv;
      cudaStream_t st = nullptr;
      bool ok = send.alloc(maxlen) == 0 && recv.alloc((size_t)maxlen * G) == 0 &&
                cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
      if (ok) {
        cudaMemsetAsync(send.p, 0, sizeof(double) * maxlen, st);
        cudaMemcpyAsync(send.p, ll[g].data(), sizeof(double) * lg, cudaMemcpyHostToDevice, st);
      }
      // NOTE: a rank that could not allocate still has to call the collective; with 8 KB buffers this does not happen
      ncclResult_t r = ok ? g_nccl.AllGather(send.p, recv.p, maxlen, ncclDouble, g_nccl.comms[g], st) : ncclInternalError;
      if (r == ncclSuccess && g == 0) {
        gathered[0].resize((size_t)maxlen * G);
        cudaMemcpyAsync(gathered[0].data(), recv.p, sizeof(double) * maxlen * G, cudaMemcpyDeviceToHost, st);
      }
      if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
      if (r != ncclSuccess && !rcs[g]) { rcs[g] = BSSM_B200_ERR_CUDA; errs[g] = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); }
    }
  };
  std::vector<std::thread> th;
  for (int g = 1; g < G; ++g) th.emplace_back(worker, g);
  worker(0);
  for (auto& t : th) t.join();
  cudaSetDevice(dev0);
  for (int g = 0; g < G; ++g)
    if (rcs[g]) return fail(rcs[g], "post-correction shard on device " + std::to_string(devs[g]) + ": " + errs[g]);
  out->device_ms = *std::max_element(ms.begin(), ms.end()); // the shards run concurrently
  out->launches = 0;
  for (int g = 0; g < G; ++g) out->launches += launches[g];
  out->sum_w = 0.0;
  for (int g = 0; g < G; ++g) {
    const uint32_t lg = sb[g + 1] - sb[g];
    for (uint32_t f = 0; f < lg; ++f) {
      const size_t o = (size_t)(sb[g] - b) + f;
      // with NCCL the log-likelihoods of the call are the gathered ones (rank 0's receive buffer)
      const double l = use_nccl ? gathered[0][(size_t)g * maxlen + f] : ll[g][f];
      if (out->loglik) out->loglik[o] = l;
      const double off = approx_ll ? approx_ll[dl.src[sb[g] + f]] : 0.0;
      const double lw = l - off;
      const double w = pf_exp(lw);
      out->weights[o] = w;
      out->posterior[o] = 0.0 + (posterior_adds_approx ? off : 0.0) + pf_log(w);
    }
  }
  g_gather_note = use_nccl ? ("nccl: " + std::to_string(G) + " ranks, one ncclAllGather of " + std::to_string(maxlen) + " doubles per rank")
                           : ("host: " + (g_nccl.why.empty() ? std::string("BSSM_B200_NO_NCCL") : g_nccl.why));
  return 0;
}

extern "C" {

// how the last multi-GPU post-correction of this thread gathered its log-likelihoods:
// "nccl: G ranks, ..." / "host: <why NCCL was not used>" / "" (single device)
const char* bssm_b200_last_gather(void) { return g_gather_note.c_str(); }
void bssm_b200_release_comm(void) { g_nccl.release(); }

int bssm_b200_postcorrection_nongaussian(const bssm_ung_model* M, const bssm_theta_map* map, int output_type,
                                         uint32_t nsim, uint32_t seed, uint32_t n_threads, int is_type,
                                         const uint32_t* counts, const double* theta, uint32_t n_par,
                                         const double* modes, uint32_t n_stored, uint32_t draw_begin,
                                         uint32_t draw_end, bssm_postcorr_out* out) {
  int rc = check_postcorr_args(M, map, counts, theta, modes, n_stored);
  if (rc) return rc;
  return postcorr_multi(n_threads, M->n, M->m, output_type, nsim, is_type, counts, n_stored, draw_begin, draw_end, out,
                         [&](const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N, int post, bssm_b200_job** J) {
                           return make_draw_job(M, map, dl, d0, d1, N, seed, theta, n_par, modes, 0, post, J);
                         });
}

// per-draw Gaussian approximation of an ssm_nlg model: approximate_for_is at the stored mode
// (src/model_ssm_nlg.cpp:153-186) + smoother, for the draws [d0, d1).  Host work, OpenMP over
// draws (the reference does this inside its OpenMP region too, src/approx_mcmc.cpp:372-373).
static int make_nlg_draw_job(const bssm_nlg_model* M, const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N,
                             uint32_t seed, const double* theta, uint32_t n_par, const double* modes, int post,
                             bssm_b200_job** job, int out_flags = 0) {
  *job = nullptr;
  const uint32_t nf = d1 - d0;
  const int n = M->n, m = nlg_state_dim(M);
  const size_t n1 = (size_t)n + 1, mm = (size_t)m * m;
  bssm_b200_job* J = new bssm_b200_job();
  struct Guard { bssm_b200_job* j; ~Guard() { delete j; } } guard{J};
  int rc = job_common(J, KIND_NLG, n, m, m, N, seed, nf, d0, 0, BSSM_POST_NONE);
  if (rc) return rc;
  J->method = BSSM_METHOD_PSI;
  J->post = post;
  J->nlg_model_id = M->model_id;
  if (!env_int("BSSM_B200_HOST_APPROX", 0)) {
    // one thread per draw on the device (approx_device.cuh): relinearise at the stored mode, smooth, pack
    const int k = m;
    std::vector<double> th((size_t)n_par * nf), md((size_t)m * n * nf);
    for (uint32_t f = 0; f < nf; ++f) {
      const uint32_t c = dl.src[d0 + f];
      memcpy(&th[(size_t)n_par * f], theta + (size_t)c * n_par, sizeof(double) * n_par);
      memcpy(&md[(size_t)m * n * f], modes + (size_t)m * n * c, sizeof(double) * m * n);
    }
    const double zero = 0.0;
    const size_t per = (size_t)(2 * m + mm + (size_t)m * k + 3 + m + 2) * n;
    if ((rc = J->y.upload(M->y, n)) || (rc = J->theta_u.upload(th.data(), th.size())) || (rc = J->mode.upload(md.data(), md.size())) ||
        (rc = J->known_u.upload(M->n_known > 0 ? M->known_params : &zero, M->n_known > 0 ? (size_t)M->n_known : 1)) ||
        (rc = J->stepc.alloc((size_t)32 * n * nf)) || (rc = J->alphahat.alloc((size_t)m * n1 * nf)) ||
        (rc = J->Lt.alloc(mm * n1 * nf)) || (rc = J->Ct.alloc(mm * n1 * nf)) || (rc = J->approx_loglik.alloc(nf)) ||
        (rc = J->scratch.alloc(per * nf)))
      return rc;
    CU(cudaMemsetAsync(J->approx_loglik.p, 0, sizeof(double) * nf, J->stream)); // the correction weight only (src/approx_mcmc.cpp:385)
    pf::NlgApproxParams ap;
    ap.n = n; ap.model_id = M->model_id; ap.n_par = (int)n_par;
    ap.y = J->y.p; ap.theta = J->theta_u.p; ap.known = J->known_u.p; ap.modes = J->mode.p;
    ap.stepc = J->stepc.p; ap.alphahat = J->alphahat.p; ap.Lt = J->Lt.p; ap.Ct = J->Ct.p; ap.scratch = J->scratch.p;
    const uint32_t ab = cdiv(nf, 32);
    if (m == 2) pf::nlg_approx_for_is_kernel<2, 2><<<ab, 32, 0, J->stream>>>(ap, nf);
    else pf::nlg_approx_for_is_kernel<1, 1><<<ab, 32, 0, J->stream>>>(ap, nf);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(J->stream));
  } else {
  std::vector<double> sc((size_t)32 * n * nf, 0.0), ah((size_t)m * n1 * nf), Lt(mm * n1 * nf), Ct(mm * n1 * nf),
      al(nf, 0.0);
  int bad = 0;
#pragma omp parallel for schedule(static)
  for (int64_t f = 0; f < (int64_t)nf; ++f) {
    const uint32_t c = dl.src[d0 + f];
    bssm_nlg_model Mi = *M;
    Mi.theta = theta + (size_t)c * n_par;
    Mi.n_theta = (int32_t)n_par;
    pfga::NlgApprox ga(&Mi);
    ga.run(modes + (size_t)m * n * c);
    if (!ga.finite_mode) { bad = 1; continue; }
    al[f] = 0.0; // approximate_for_is: the correction weight only (src/approx_mcmc.cpp:385)
    double* scf = &sc[(size_t)32 * n * f];
    for (int t = 0; t < n; ++t) pack_nlg_step(ga, M->y, m, t, scf + (size_t)32 * t);
    std::copy(ga.alphahat.begin(), ga.alphahat.end(), &ah[(size_t)m * n1 * f]);
    std::copy(ga.L.begin(), ga.L.end(), &Lt[mm * n1 * f]);
    std::copy(ga.Ct.begin(), ga.Ct.end(), &Ct[mm * n1 * f]);
  }
  if (bad) return fail(BSSM_B200_ERR_ARG, "post-correction: non-finite Gaussian approximation for a draw");
  if ((rc = J->stepc.upload(sc.data(), sc.size())) || (rc = J->alphahat.upload(ah.data(), ah.size())) ||
      (rc = J->Lt.upload(Lt.data(), Lt.size())) || (rc = J->Ct.upload(Ct.data(), Ct.size())) ||
      (rc = J->approx_loglik.upload(al.data(), nf)))
    return rc;
  }
  pf::NlgPsiParams& pp = J->npp;
  pp.n = n; pp.stepc = J->stepc.p; pp.alphahat = J->alphahat.p; pp.Lt = J->Lt.p; pp.Ct = J->Ct.p;
  pp.approx_loglik = J->approx_loglik.p;
  if (M->model_id == BSSM_NLG_GROWTH) { pp.dT = M->known_params[0]; pp.K = M->known_params[1]; }
  pp.fstride_stepc = (size_t)32 * n; pp.fstride_a = (size_t)m * n1; pp.fstride_m = mm * n1;
  if ((rc = finish_job(J, out_flags, post))) return rc;
  guard.j = nullptr;
  *job = J;
  return 0;
}

int bssm_b200_job_create_postcorr_nlg(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, int is_type,
                                      const uint32_t* counts, const double* theta, uint32_t n_par,
                                      const double* modes, uint32_t n_stored, uint32_t draw_begin,
                                      uint32_t draw_end, int out_flags, bssm_b200_job** job) {
  if (!M || !counts || !theta || !modes || !job || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  {
    int rcm = check_nlg_model(M);
    if (rcm) return rcm;
    if (n_par < (uint32_t)nlg_n_theta(M)) return fail(BSSM_B200_ERR_ARG, "post-correction (nonlinear): too few parameters");
  }
  if (is_type == 1) return fail(BSSM_B200_ERR_ARG, "resident post-correction jobs need equal N per draw (is_type 2 or 3)");
  int rc = require_device();
  if (rc) return rc;
  DrawList dl;
  if ((rc = build_draws(is_type, counts, n_stored, &dl))) return rc;
  if (draw_begin == 0 && draw_end == 0) draw_end = (uint32_t)dl.src.size();
  if (draw_begin >= draw_end || draw_end > dl.src.size()) return fail(BSSM_B200_ERR_ARG, "bad draw range");
  return make_nlg_draw_job(M, dl, draw_begin, draw_end, nsim, seed, theta, n_par, modes, BSSM_POST_NONE, job, out_flags);
}

int bssm_b200_postcorrection_nonlinear(const bssm_nlg_model* M, int output_type, uint32_t nsim, uint32_t seed,
                                       uint32_t n_threads, int is_type, const uint32_t* counts,
                                       const double* theta, uint32_t n_par, const double* modes,
                                       uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                                       bssm_postcorr_out* out) {
  if (!M || !counts || !theta || !modes || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  {
    int rcm = check_nlg_model(M);
    if (rcm) return rcm;
    if (n_par < (uint32_t)nlg_n_theta(M)) return fail(BSSM_B200_ERR_ARG, "post-correction (nonlinear): too few parameters");
  }
  return postcorr_multi(n_threads, M->n, nlg_state_dim(M), output_type, nsim, is_type, counts, n_stored,
                         draw_begin, draw_end, out,
                         [&](const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N, int post, bssm_b200_job** J) {
                           return make_nlg_draw_job(M, dl, d0, d1, N, seed, theta, n_par, modes, post, J);
                         });
}

// ---------------------------------------------- is_correction_bsf (ung family)
int bssm_b200_is_correction_bsf(const bssm_ung_model* M, const bssm_theta_map* map, int output_type,
                                uint32_t nsim, uint32_t seed, uint32_t n_threads, int is_type,
                                const uint32_t* counts, const double* theta, uint32_t n_par,
                                const double* approx_loglik, uint32_t n_stored, uint32_t draw_begin,
                                uint32_t draw_end, bssm_postcorr_out* out) {
  if (!M || !map || !counts || !theta || !approx_loglik || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  if (M->m != 1) return fail(BSSM_B200_ERR_UNSUPPORTED, "is_correction_bsf: m = 1 models in this build");
  if (M->xbeta && map->n_xreg <= 0) return fail(BSSM_B200_ERR_ARG, "the model has regressors: theta_map needs n_xreg and xreg");
  if (map->n_xreg > 0 && (!map->xreg || (int)n_par < map->n_xreg || map->model_type == 3))
    return fail(BSSM_B200_ERR_ARG, "n_xreg > 0 needs xreg, beta in theta and a model with a regression part");
  return postcorr_multi(
      n_threads, M->n, 1, output_type, nsim, is_type, counts, n_stored, draw_begin, draw_end, out,
      [&](const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N, int post, bssm_b200_job** J) {
        const uint32_t nf = d1 - d0;
        std::vector<double> par((size_t)6 * nf);
        DrawModels dm;
        if (map->n_xreg > 0) dm.xbeta.resize((size_t)M->n * nf);
        for (uint32_t f = 0; f < nf; ++f) {
          const double* th = theta + (size_t)dl.src[d0 + f] * n_par;
          int rc = theta_to_par(M, map, th, n_par, &par[(size_t)6 * f]);
          if (rc) return rc;
          if (map->n_xreg > 0) theta_to_xbeta(map, M->n, th, n_par, &dm.xbeta[(size_t)M->n * f]);
        }
        return create_ung_job(M, BSSM_METHOD_BSF, N, seed, nf, d0, 0, post, par.data(), nullptr, J,
                              map->n_xreg > 0 ? &dm : nullptr);
      },
      approx_loglik, /*posterior_adds_approx=*/false);
}

// ------------------------------------------ parameter sets (ssm_ung with an R update_fn)
// slices of the stored draws [d0, d1) of one parset component (nullptr -> empty = shared)
static void gather_slices(const double* src, size_t slice, const DrawList& dl, uint32_t d0, uint32_t d1,
                          std::vector<double>& dst) {
  dst.clear();
  if (!src) return;
  dst.resize(slice * (d1 - d0));
  for (uint32_t f = 0; f < d1 - d0; ++f)
    memcpy(&dst[slice * f], src + slice * dl.src[d0 + f], sizeof(double) * slice);
}

static int check_ung_parset(const bssm_ung_model* M, const bssm_parset* ps) {
  if (M->distribution == BSSM_DIST_GAUSSIAN) return fail(BSSM_B200_ERR_ARG, "parameter sets: non-Gaussian ssm_ung models");
  if (ps->beta && (!ps->xreg || ps->n_xreg < 1)) return fail(BSSM_B200_ERR_ARG, "parset: beta needs xreg and n_xreg");
  return 0;
}

// parset_ung::update for the draws [d0, d1): gathers the components update_fn returned
static void ung_parset_draws(const bssm_ung_model* M, const bssm_parset* ps, const DrawList& dl, uint32_t d0,
                             uint32_t d1, DrawModels& dm) {
  const size_t n = (size_t)M->n, m = (size_t)M->m, k = (size_t)M->k;
  dm.parset = true;
  gather_slices(ps->phi, 1, dl, d0, d1, dm.phi);
  gather_slices(ps->Z, m * (M->Ztv ? n : 1), dl, d0, d1, dm.Z);
  gather_slices(ps->T, m * m * (M->Ttv ? n : 1), dl, d0, d1, dm.T);
  gather_slices(ps->R, m * k * (M->Rtv ? n : 1), dl, d0, d1, dm.R);
  gather_slices(ps->a1, m, dl, d0, d1, dm.a1);
  gather_slices(ps->P1, m * m, dl, d0, d1, dm.P1);
  gather_slices(ps->D, M->Dtv ? n : 1, dl, d0, d1, dm.D);
  gather_slices(ps->C, m * (M->Ctv ? n : 1), dl, d0, d1, dm.C);
  if (ps->beta) { // compute_xbeta, src/model_ssm_ung.h:83
    dm.xbeta.resize(n * (d1 - d0));
    for (uint32_t f = 0; f < d1 - d0; ++f) {
      const double* beta = ps->beta + (size_t)ps->n_xreg * dl.src[d0 + f];
      for (size_t t = 0; t < n; ++t) {
        double sum = 0.0;
        for (int j = 0; j < ps->n_xreg; ++j) sum += ps->xreg[t + n * j] * beta[j];
        dm.xbeta[n * f + t] = sum;
      }
    }
  }
}

int bssm_b200_postcorrection_ung_parset(const bssm_ung_model* M, const bssm_parset* ps, int output_type,
                                        uint32_t nsim, uint32_t seed, uint32_t n_threads, int is_type,
                                        const uint32_t* counts, const double* modes, uint32_t n_stored,
                                        uint32_t draw_begin, uint32_t draw_end, bssm_postcorr_out* out) {
  if (!M || !ps || !counts || !modes || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  int rc = check_ung_parset(M, ps);
  if (rc) return rc;
  return postcorr_multi(
      n_threads, M->n, M->m, output_type, nsim, is_type, counts, n_stored, draw_begin, draw_end, out,
      [&](const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N, int post, bssm_b200_job** J) {
        DrawModels dm;
        ung_parset_draws(M, ps, dl, d0, d1, dm);
        gather_slices(modes, (size_t)M->n, dl, d0, d1, dm.modes);
        return create_ung_job(M, BSSM_METHOD_PSI, N, seed, d1 - d0, d0, 0, post, nullptr, nullptr, J, &dm);
      });
}

int bssm_b200_is_correction_bsf_ung_parset(const bssm_ung_model* M, const bssm_parset* ps, int output_type,
                                           uint32_t nsim, uint32_t seed, uint32_t n_threads, int is_type,
                                           const uint32_t* counts, const double* approx_loglik, uint32_t n_stored,
                                           uint32_t draw_begin, uint32_t draw_end, bssm_postcorr_out* out) {
  if (!M || !ps || !counts || !approx_loglik || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  int rc = check_ung_parset(M, ps);
  if (rc) return rc;
  return postcorr_multi(
      n_threads, M->n, M->m, output_type, nsim, is_type, counts, n_stored, draw_begin, draw_end, out,
      [&](const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N, int post, bssm_b200_job** J) {
        DrawModels dm;
        ung_parset_draws(M, ps, dl, d0, d1, dm);
        return create_ung_job(M, BSSM_METHOD_BSF, N, seed, d1 - d0, d0, 0, post, nullptr, nullptr, J, &dm);
      },
      approx_loglik, /*posterior_adds_approx=*/false);
}

// ------------------------------------------ parameter sets (ssm_mng with an R update_fn)
static void mng_parset_draws(const bssm_mng_model* M, const bssm_parset* ps, const DrawList& dl, uint32_t d0,
                             uint32_t d1, MngDraws& dm) {
  const size_t n = (size_t)M->n, m = (size_t)M->m, k = (size_t)M->k, p = (size_t)M->p;
  gather_slices(ps->phi, p, dl, d0, d1, dm.phi);
  gather_slices(ps->Z, p * m * (M->Ztv ? n : 1), dl, d0, d1, dm.Z);
  gather_slices(ps->T, m * m * (M->Ttv ? n : 1), dl, d0, d1, dm.T);
  gather_slices(ps->R, m * k * (M->Rtv ? n : 1), dl, d0, d1, dm.R);
  gather_slices(ps->a1, m, dl, d0, d1, dm.a1);
  gather_slices(ps->P1, m * m, dl, d0, d1, dm.P1);
  gather_slices(ps->D, p * (M->Dtv ? n : 1), dl, d0, d1, dm.D);
  gather_slices(ps->C, m * (M->Ctv ? n : 1), dl, d0, d1, dm.C);
}

int bssm_b200_postcorrection_mng_parset(const bssm_mng_model* M, const bssm_parset* ps, int output_type,
                                        uint32_t nsim, uint32_t seed, uint32_t n_threads, int is_type,
                                        const uint32_t* counts, const double* modes, uint32_t n_stored,
                                        uint32_t draw_begin, uint32_t draw_end, bssm_postcorr_out* out) {
  if (!M || !ps || !counts || !modes || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  if (ps->beta) return fail(BSSM_B200_ERR_ARG, "ssm_mng has no regression part");
  return postcorr_multi(
      n_threads, M->n, M->m, output_type, nsim, is_type, counts, n_stored, draw_begin, draw_end, out,
      [&](const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N, int post, bssm_b200_job** J) {
        MngDraws dm;
        mng_parset_draws(M, ps, dl, d0, d1, dm);
        gather_slices(modes, (size_t)M->p * M->n, dl, d0, d1, dm.modes);
        return create_mng_job(M, BSSM_METHOD_PSI, N, seed, d1 - d0, d0, 0, post, J, &dm);
      });
}

int bssm_b200_is_correction_bsf_mng_parset(const bssm_mng_model* M, const bssm_parset* ps, int output_type,
                                           uint32_t nsim, uint32_t seed, uint32_t n_threads, int is_type,
                                           const uint32_t* counts, const double* approx_loglik, uint32_t n_stored,
                                           uint32_t draw_begin, uint32_t draw_end, bssm_postcorr_out* out) {
  if (!M || !ps || !counts || !approx_loglik || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  if (ps->beta) return fail(BSSM_B200_ERR_ARG, "ssm_mng has no regression part");
  return postcorr_multi(
      n_threads, M->n, M->m, output_type, nsim, is_type, counts, n_stored, draw_begin, draw_end, out,
      [&](const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N, int post, bssm_b200_job** J) {
        MngDraws dm;
        mng_parset_draws(M, ps, dl, d0, d1, dm);
        return create_mng_job(M, BSSM_METHOD_BSF, N, seed, d1 - d0, d0, 0, post, J, &dm);
      },
      approx_loglik, /*posterior_adds_approx=*/false);
}

// ------------------------------------------------- is_correction_bsf (ssm_nlg)
// approx_mcmc::is_correction_bsf<ssm_nlg>, src/approx_mcmc.cpp:442-518: update_model(theta_i), bsf_filter,
// weight = exp(loglik - approx_loglik_i), posterior = prior + log(weight)
int bssm_b200_is_correction_bsf_nlg(const bssm_nlg_model* M, int output_type, uint32_t nsim, uint32_t seed,
                                    uint32_t n_threads, int is_type, const uint32_t* counts, const double* theta,
                                    uint32_t n_par, const double* approx_loglik, uint32_t n_stored,
                                    uint32_t draw_begin, uint32_t draw_end, bssm_postcorr_out* out) {
  if (!M || !counts || !theta || !approx_loglik || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  {
    int rcm = check_nlg_model(M, true);
    if (rcm) return rcm;
    if (n_par < (uint32_t)nlg_n_theta(M)) return fail(BSSM_B200_ERR_ARG, "is_correction_bsf (nonlinear): too few parameters");
  }
  return postcorr_multi(
      n_threads, M->n, nlg_state_dim(M), output_type, nsim, is_type, counts, n_stored, draw_begin, draw_end, out,
      [&](const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N, int post, bssm_b200_job** J) {
        const uint32_t nf = d1 - d0;
        std::vector<double> th((size_t)n_par * nf);
        for (uint32_t f = 0; f < nf; ++f)
          memcpy(&th[(size_t)n_par * f], theta + (size_t)dl.src[d0 + f] * n_par, sizeof(double) * n_par);
        int flags = post == BSSM_POST_NONE ? 0 : (BSSM_OUT_ALPHA | BSSM_OUT_INDICES);
        return create_nlg_job(M, BSSM_METHOD_BSF, N, seed, nf, d0, flags, post, J, th.data(), n_par);
      },
      approx_loglik, /*posterior_adds_approx=*/false);
}

// ------------------------------------------------- is_correction_bsf (ssm_sde)
int bssm_b200_is_correction_bsf_sde(const bssm_sde_model* M, uint32_t L_f, int output_type, uint32_t nsim,
                                    uint32_t seed, uint32_t n_threads, int is_type, const uint32_t* counts,
                                    const double* theta, uint32_t n_par, const double* approx_loglik,
                                    uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                                    bssm_postcorr_out* out) {
  if (!M || !counts || !theta || !approx_loglik || n_stored == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  return postcorr_multi(
      n_threads, M->n, 1, output_type, nsim, is_type, counts, n_stored, draw_begin, draw_end, out,
      [&](const DrawList& dl, uint32_t d0, uint32_t d1, uint32_t N, int post, bssm_b200_job** J) {
        const uint32_t nf = d1 - d0;
        std::vector<double> th((size_t)n_par * nf);
        for (uint32_t f = 0; f < nf; ++f)
          memcpy(&th[(size_t)n_par * f], theta + (size_t)dl.src[d0 + f] * n_par, sizeof(double) * n_par);
        int flags = post == BSSM_POST_NONE ? 0 : (BSSM_OUT_ALPHA | BSSM_OUT_INDICES);
        return create_sde_job(M, N, L_f, seed, nf, d0, flags, post, th.data(), n_par, J);
      },
      approx_loglik, /*posterior_adds_approx=*/true);
}

// ------------------------------------------------------------- suggest_N
// sd of `replications` psi-APF log-likelihood estimates for one particle count: a batch of
// replicate filters at fixed theta; the streams of candidate c are c*replications ... so that
// candidates do not share random numbers (the reference continues one engine, src/R_postcorrection.cpp:27-35)
static int loglik_sd(bssm_b200_job* J, uint32_t replications, double* sd) {
  struct JG { bssm_b200_job* j; ~JG() { delete j; } } jg{J};
  float ms = 0;
  int rc = bssm_b200_job_run(J, &ms);
  if (rc) return rc;
  std::vector<double> ll(replications);
  if ((rc = bssm_b200_job_logliks(J, ll.data()))) return rc;
  // arma::stddev: two-pass, N - 1 normalisation
  double mean = 0.0;
  for (double v : ll) mean += v;
  mean /= (double)replications;
  double acc = 0.0;
  for (double v : ll) acc += (v - mean) * (v - mean);
  *sd = std::sqrt(acc / (double)(replications - 1));
  return 0;
}

int bssm_b200_suggest_n_nongaussian(const bssm_ung_model* M, const bssm_theta_map* map, const double* theta,
                                    uint32_t n_par, const uint32_t* candidates, uint32_t n_candidates,
                                    uint32_t replications, uint32_t seed, double* sds) {
  if (!M || !map || !theta || !candidates || !sds || n_candidates == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  if (replications < 2) return fail(BSSM_B200_ERR_ARG, "replications must be >= 2");
  // model.update_model(theta)
  bssm_ung_model Mi = *M;
  double par[6];
  std::vector<double> R, P1, xb;
  if (map->model_type == 1) {
    // ssm_ung: theta maps onto the model through an R closure (update_model(theta, update_fn),
    // src/R_postcorrection.cpp:37-40); the caller has applied it, `M` IS the model at theta
  } else if (M->m == 1 && !(map->model_type == 2 && map->n_xreg > 0)) {
    if (M->xbeta && map->n_xreg <= 0) return fail(BSSM_B200_ERR_ARG, "the model has regressors: theta_map needs n_xreg and xreg");
    int rc = theta_to_par(M, map, theta, n_par, par);
    if (rc) return rc;
    if (map->n_xreg > 0) { // ar1_ng with regressors: beta = the tail of theta
      if (!map->xreg || (int)n_par < map->n_xreg) return fail(BSSM_B200_ERR_ARG, "n_xreg > 0 needs xreg and beta in theta");
      xb.resize(M->n);
      theta_to_xbeta(map, M->n, theta, n_par, xb.data());
      Mi.xbeta = xb.data();
    }
    Mi.T = &par[0]; Mi.R = &par[1]; Mi.a1 = &par[2]; Mi.P1 = &par[3]; Mi.C = &par[4]; Mi.phi = par[5];
    Mi.Ttv = 0; Mi.Rtv = 0; Mi.Ctv = 0;
  } else {
    if (map->model_type != 2) return fail(BSSM_B200_ERR_UNSUPPORTED, "suggest_N with m > 1: bsm_ng (model_type 2) in this build");
    if (M->Rtv) return fail(BSSM_B200_ERR_UNSUPPORTED, "time-invariant R only");
    R.resize((size_t)M->m * M->k); P1.resize((size_t)M->m * M->m);
    if (map->n_xreg > 0) xb.resize(M->n);
    int rc = theta_to_bsm(M, map, theta, n_par, R.data(), P1.data(), &Mi.phi, xb.empty() ? nullptr : xb.data());
    if (rc) return rc;
    Mi.R = R.data(); Mi.P1 = P1.data();
    if (!xb.empty()) Mi.xbeta = xb.data();
  }
  for (uint32_t c = 0; c < n_candidates; ++c) {
    bssm_b200_job* J = nullptr;
    int rc = create_ung_job(&Mi, BSSM_METHOD_PSI, candidates[c], seed, replications, c * replications, 0,
                            BSSM_POST_NONE, nullptr, nullptr, &J);
    if (rc) return rc;
    if ((rc = loglik_sd(J, replications, &sds[c]))) return rc;
  }
  return 0;
}

int bssm_b200_suggest_n_mng(const bssm_mng_model* M, const uint32_t* candidates, uint32_t n_candidates,
                            uint32_t replications, uint32_t seed, double* sds) {
  if (!M || !candidates || !sds || n_candidates == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  if (replications < 2) return fail(BSSM_B200_ERR_ARG, "replications must be >= 2");
  for (uint32_t c = 0; c < n_candidates; ++c) {
    bssm_b200_job* J = nullptr;
    int rc = create_mng_job(M, BSSM_METHOD_PSI, candidates[c], seed, replications, c * replications, 0, BSSM_POST_NONE, &J);
    if (rc) return rc;
    if ((rc = loglik_sd(J, replications, &sds[c]))) return rc;
  }
  return 0;
}

int bssm_b200_suggest_n_nonlinear(const bssm_nlg_model* M, const double* theta, uint32_t n_par,
                                  const uint32_t* candidates, uint32_t n_candidates, uint32_t replications,
                                  uint32_t seed, uint32_t max_iter, double conv_tol, double* sds) {
  if (!M || !theta || !candidates || !sds || n_candidates == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  if (replications < 2) return fail(BSSM_B200_ERR_ARG, "replications must be >= 2");
  bssm_nlg_model Mi = *M;
  Mi.theta = theta;
  Mi.n_theta = (int32_t)n_par;
  for (uint32_t c = 0; c < n_candidates; ++c) {
    bssm_b200_job* J = nullptr;
    int rc = bssm_b200_job_create_nlg_psi(&Mi, candidates[c], seed, max_iter, conv_tol, replications,
                                          c * replications, 0, BSSM_POST_NONE, &J);
    if (rc) return rc;
    if ((rc = loglik_sd(J, replications, &sds[c]))) return rc;
  }
  return 0;
}

// ---------------------------------------------------------- one-shot calls
static int one_shot(const bssm_ung_model* M, int method, uint32_t nsim, uint32_t seed, int post,
                    bssm_filter_out* out) {
  if (!out) return fail(BSSM_B200_ERR_ARG, "null out");
  int flags = 0;
  if (out->alpha) flags |= BSSM_OUT_ALPHA;
  if (out->weights) flags |= BSSM_OUT_WEIGHTS;
  if (out->indices) flags |= BSSM_OUT_INDICES;
  bssm_b200_job* J = nullptr;
  int rc = bssm_b200_job_create_ung(M, method, nsim, seed, 1, 0, flags, post, &J);
  if (rc) return rc;
  return run_fetch_destroy(J, out);
}

int bssm_b200_bsf(const bssm_ung_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  return one_shot(M, BSSM_METHOD_BSF, nsim, seed, BSSM_POST_FILTER_SUMMARY, out);
}
int bssm_b200_bsf_smoother(const bssm_ung_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  return one_shot(M, BSSM_METHOD_BSF, nsim, seed, BSSM_POST_SMOOTHER, out);
}
int bssm_b200_psi_smoother(const bssm_ung_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  return one_shot(M, BSSM_METHOD_PSI, nsim, seed, BSSM_POST_SMOOTHER, out);
}
int bssm_b200_psi_filter(const bssm_ung_model* M, uint32_t nsim, uint32_t seed, bssm_filter_out* out) {
  return one_shot(M, BSSM_METHOD_PSI, nsim, seed, BSSM_POST_NONE, out);
}

// ------------------------------------------------------------- unit hook
namespace {
// single-CTA stratified resampling with caller-supplied uniforms: the same
// fixed-point cumsum and threshold counting the filter kernel uses, r[] read
// from memory instead of Philox.
__global__ void __launch_bounds__(1024)
stratified_kernel(const double* __restrict__ p, uint32_t n_p, const double* __restrict__ r, uint32_t N,
                  uint32_t* __restrict__ xp) {
  __shared__ pf::CollSmem csm;
  pf::Sync sync; sync.mode = pf::SYNC_BLOCK; sync.counter = nullptr; sync.nranks = 1; sync.target = 0;
  pf::Coll coll;
  coll.sm = &csm; coll.sync = &sync; coll.xch = nullptr; coll.ranks = 1; coll.rank = 0;
  coll.lane = threadIdx.x & 31; coll.warp = threadIdx.x >> 5; coll.nwarps = (blockDim.x + 31) >> 5;
  const uint32_t chunk = (n_p + blockDim.x - 1) / blockDim.x;
  const uint32_t k0 = threadIdx.x * chunk;
  long long th = 0, tl = 0;
  for (uint32_t k = k0; k < k0 + chunk && k < n_p; ++k) {
    pf_fixsum tm = pf_fixsum_term(p[k]);
    th += tm.hi; tl += tm.lo;
  }
  long long rh, rl, Th, Tl;
  coll.exscan2(th, tl, rh, rl, Th, Tl);
  const double invS = 1.0 / pf_fixsum_value(Th, Tl);
  const double Nd = (double)N, invN = 1.0 / Nd;
  auto u = [&](uint32_t j) { return (r[j] + (double)j) * invN; };
  auto count = [&](double pk) -> uint32_t {
    // plain binary search here (unit hook, not performance critical)
    uint32_t lo = 0, hi = N;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (u(mid) <= pk) lo = mid + 1; else hi = mid; }
    return lo;
  };
  uint32_t lo = (k0 == 0) ? 0u : ((k0 - 1 >= n_p - 1) ? N : count(pf_fixsum_value(rh, rl) * invS));
  for (uint32_t k = k0; k < k0 + chunk && k < n_p; ++k) {
    pf_fixsum tm = pf_fixsum_term(p[k]);
    rh += tm.hi; rl += tm.lo;
    uint32_t c = (k >= n_p - 1) ? N : count(pf_fixsum_value(rh, rl) * invS);
    for (uint32_t j = lo; j < c; ++j) xp[j] = k;
    lo = c > lo ? c : lo;
  }
}
} // namespace

namespace {
__global__ void math_probe_kernel(int op, const double* __restrict__ x, double* __restrict__ y, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = x[i], s, c;
  switch (op) {
    case 0: v = pf_exp(v); break;
    case 1: v = pf_log(v); break;
    case 2: v = pf_log1p(v); break;
    case 3: pf_sincos2pi(v, &s, &c); v = s; break;
    case 4: pf_sincos2pi(v, &s, &c); v = c; break;
    case 5: v = pf_sin(v); break;
    default: v = pf_cos(v); break;
  }
  y[i] = v;
}
__global__ void rng_probe_kernel(uint32_t seed, uint32_t filter, uint32_t t, uint32_t count,
                                 double* __restrict__ z, double* __restrict__ r) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  z[i] = pf_normal(seed, filter, t, 0u, i);
  r[i] = pf_resample_uniform(seed, filter, t, i);
}
} // namespace

int bssm_b200_math_probe(int op, const double* x, double* y, uint32_t n) {
  if (!x || !y || n == 0 || op < 0 || op > 6) return fail(BSSM_B200_ERR_ARG, "bad argument");
  int rc = require_device();
  if (rc) return rc;
  DevBuf<double> dx, dy;
  if ((rc = dx.upload(x, n)) || (rc = dy.alloc(n))) return rc;
  math_probe_kernel<<<cdiv(n, 256), 256>>>(op, dx.p, dy.p, n);
  CU(cudaGetLastError());
  CU(cudaMemcpy(y, dy.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return 0;
}

int bssm_b200_rng_probe(uint32_t seed, uint32_t filter, uint32_t t, uint32_t count, double* z, double* r) {
  if (!z || !r || count == 0) return fail(BSSM_B200_ERR_ARG, "bad argument");
  int rc = require_device();
  if (rc) return rc;
  DevBuf<double> dz, dr;
  if ((rc = dz.alloc(count)) || (rc = dr.alloc(count))) return rc;
  rng_probe_kernel<<<cdiv(count, 256), 256>>>(seed, filter, t, count, dz.p, dr.p);
  CU(cudaGetLastError());
  CU(cudaMemcpy(z, dz.p, sizeof(double) * count, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(r, dr.p, sizeof(double) * count, cudaMemcpyDeviceToHost));
  return 0;
}

int bssm_b200_stratified_sample(const double* p, uint32_t n_p, const double* r, uint32_t N, uint32_t* xp) {
  if (!p || !r || !xp || n_p == 0 || N == 0) return fail(BSSM_B200_ERR_ARG, "null or empty argument");
  int rc = require_device();
  if (rc) return rc;
  DevBuf<double> dp, dr;
  DevBuf<uint32_t> dx;
  if ((rc = dp.upload(p, n_p)) || (rc = dr.upload(r, N)) || (rc = dx.alloc(N))) return rc;
  stratified_kernel<<<1, 1024>>>(dp.p, n_p, dr.p, N, dx.p);
  CU(cudaGetLastError());
  CU(cudaMemcpy(xp, dx.p, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost));
  return 0;
}

} // extern "C"
