#include "filters.h"
namespace pfl {
// model_id: 0 = growth, 1 = ar_exp, 2 = sin_exp, 3 = linear_gaussian (device functor registry, include/bssm_b200.h);
// the linear-Gaussian policies are instantiated in k_nlg_lin.cu (build time)
int launch_nlg_lin(int which, const pf::FilterParams&, const void* mp, const Geometry&, uint32_t nf, Stream);
int launch_nlg_bsf(int model_id, const pf::FilterParams& fp, const pf::NlgParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (model_id) {
    case 0: return launch<pf::NlgGrowthBsf>(fp, mp, g, nf, s);
    case 1: return launch<pf::NlgArExpBsf>(fp, mp, g, nf, s);
    case 2: return launch<pf::NlgSinExpBsf>(fp, mp, g, nf, s);
    case 3: return launch_nlg_lin(0, fp, &mp, g, nf, s);
  }
  return -1;
}
int launch_nlg_psi(int model_id, const pf::FilterParams& fp, const pf::NlgPsiParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (model_id) {
    case 0: return launch<pf::NlgGrowthPsi>(fp, mp, g, nf, s);
    case 1: return launch<pf::NlgArExpPsi>(fp, mp, g, nf, s);
    case 2: return launch<pf::NlgSinExpPsi>(fp, mp, g, nf, s);
    case 3: return launch_nlg_lin(1, fp, &mp, g, nf, s);
  }
  return -1;
}
int launch_nlg_ekpf(int model_id, const pf::FilterParams& fp, const pf::NlgParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (model_id) {
    case 0: return launch<pf::NlgEkpf<0>>(fp, mp, g, nf, s);
    case 1: return launch<pf::NlgEkpf<1>>(fp, mp, g, nf, s);
    case 2: return launch<pf::NlgEkpf<2>>(fp, mp, g, nf, s);
    case 3: return launch_nlg_lin(2, fp, &mp, g, nf, s);
  }
  return -1;
}
int occupancy_nlg_bsf(const Geometry& g, int* out) { return occupancy<pf::NlgGrowthBsf>(g, out); }
int occupancy_nlg_psi(const Geometry& g, int* out) { return occupancy<pf::NlgGrowthPsi>(g, out); }
} // namespace pfl
