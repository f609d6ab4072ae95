// filters.h -- the filter-kernel launchers, one translation unit per model family
// (see launch.cuh for the status convention).
#pragma once
#include "launch.cuh"
#include "models_ung.cuh"
#include "models_nlg_sde.cuh"

namespace pfl {
using Stream = cudaStream_t;
// k_ung_psi1.cu: psi-APF, m = 1 (svm, ar1_ng, univariate ssm_ung / ssm_ulg)
int launch_ung_psi1(const pf::FilterParams&, const pf::UngParams&, const Geometry&, uint32_t nf, Stream);
int occupancy_ung_psi1(const Geometry&, int* out);
// k_ung_psim.cu: psi-APF with per-filter approximations: m in {2, 3}, and m = 1 when each filter has its own xbeta
int launch_ung_psim(int m, const pf::FilterParams&, const pf::UngParams&, const Geometry&, uint32_t nf, Stream);
// k_ung_bsf_a.cu / k_ung_bsf_b.cu: bootstrap filter, (m, k) in {(1,1),(2,1),(2,2)} / {(3,2),(3,3)}
int launch_ung_bsf_a(int m, int k, const pf::FilterParams&, const pf::UngParams&, const Geometry&, uint32_t nf, Stream);
int launch_ung_bsf_b(int m, int k, const pf::FilterParams&, const pf::UngParams&, const Geometry&, uint32_t nf, Stream);
int launch_ung_bsf_draws(const pf::FilterParams&, const pf::UngParams&, const Geometry&, uint32_t nf, Stream); // m = 1, one model per filter
int occupancy_ung_bsf(const Geometry&, int* out);
// k_ung_dyn.cu: bootstrap filter and psi-APF with a run-time state dimension (m <= 16, k <= m), P = 2 only
int launch_ung_bsf_dyn(const pf::FilterParams&, const pf::UngParams&, const Geometry&, uint32_t nf, Stream);
int launch_ung_psi_dyn(const pf::FilterParams&, const pf::UngParams&, const Geometry&, uint32_t nf, Stream);
int occupancy_ung_dyn(const Geometry&, int* out);
int launch_mng_bsf_dyn(const pf::FilterParams&, const pf::MngParams&, const Geometry&, uint32_t nf, Stream); // ssm_mng, m <= 16
int launch_mng_psi_dyn(const pf::FilterParams&, const pf::MngParams&, const Geometry&, uint32_t nf, Stream);
int occupancy_mng_dyn(const Geometry&, int* out);
// k_mng.cu: multivariate-observation bootstrap filter and psi-APF (m <= 3)
int launch_mng_bsf(int m, int k, const pf::FilterParams&, const pf::MngParams&, const Geometry&, uint32_t nf, Stream);
int launch_mng_psi(int m, const pf::FilterParams&, const pf::MngParams&, const Geometry&, uint32_t nf, Stream);
int occupancy_mng_bsf(const Geometry&, int* out);
// k_mng_draws.cu: the same with one model (and one approximation) per filter -- parameter sets
int launch_mng_bsf_draws(int m, int k, const pf::FilterParams&, const pf::MngParams&, const Geometry&, uint32_t nf, Stream);
int launch_mng_psi_draws(int m, const pf::FilterParams&, const pf::MngParams&, const Geometry&, uint32_t nf, Stream);
// k_nlg.cu: ssm_nlg functor registry (growth, ar_exp), bootstrap and psi-APF
int launch_nlg_bsf(int model_id, const pf::FilterParams&, const pf::NlgParams&, const Geometry&, uint32_t nf, Stream);
int launch_nlg_bsf_draws(int model_id, const pf::FilterParams&, const pf::NlgParams&, const Geometry&, uint32_t nf, Stream); // k_nlg_draws.cu: one theta per filter
int launch_nlg_psi(int model_id, const pf::FilterParams&, const pf::NlgPsiParams&, const Geometry&, uint32_t nf, Stream);
int launch_nlg_ekpf(int model_id, const pf::FilterParams&, const pf::NlgParams&, const Geometry&, uint32_t nf, Stream); // EKF particle filter
int occupancy_nlg_bsf(const Geometry&, int* out);
int occupancy_nlg_psi(const Geometry&, int* out);
// k_sde.cu: Milstein bootstrap filter; model_id 0 = GBM, 1 = Poisson-OU (device functor registry)
int launch_sde_bsf(int model_id, const pf::FilterParams&, const pf::SdeParams&, const Geometry&, uint32_t nf, Stream);
int occupancy_sde_bsf(const Geometry&, int* out);
// user_models.cu: ssm_sde functors given as CUDA source, compiled at run time (NVRTC) into the same kernels.
// Status: as above, and -2 = registration / compilation failed (text in user_model_error()).
int user_sde_register(const char* source, int* model_id);
bool user_sde_exists(int model_id);
const char* user_model_error();
void user_models_release();
int launch_user_sde(int model_id, const pf::FilterParams&, const pf::SdeParams&, const Geometry&, uint32_t nf, bool per_filter, Stream);
int occupancy_user_sde(int model_id, const Geometry&, bool per_filter, int* out);
// ssm_nlg bootstrap filter from user source (Z_fn, H_fn, T_fn, R_fn, a1_fn, P1_fn; p = 1, state dimension m, k etas)
int user_nlg_register(const char* source, int m, int k, int* model_id);
bool user_nlg_exists(int model_id);
int launch_user_nlg(int model_id, int m, int k, const pf::FilterParams&, const pf::NlgParams&, const Geometry&, uint32_t nf,
                    bool per_filter, Stream);
int user_nlg_init(int model_id, int m, const double* theta, size_t stride, const double* known, uint32_t nf, double* out, Stream);
int occupancy_user_nlg(int model_id, int m, int k, const Geometry&, bool per_filter, int* out);
} // namespace pfl
