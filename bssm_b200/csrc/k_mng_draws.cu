#include "filters.h"
// ssm_mng with one model per filter (parameter sets): is_correction_psi / is_correction_bsf for ssm_mng,
// src/approx_mcmc.cpp:1038-1157 and :1159-1273
namespace pfl {
int launch_mng_bsf_draws(int m, int k, const pf::FilterParams& fp, const pf::MngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (m * 10 + k) {
    case 11: return launch<pf::MngBsf<1, 1, true>>(fp, mp, g, nf, s);
    case 21: return launch<pf::MngBsf<2, 1, true>>(fp, mp, g, nf, s);
    case 22: return launch<pf::MngBsf<2, 2, true>>(fp, mp, g, nf, s);
    case 33: return launch<pf::MngBsf<3, 3, true>>(fp, mp, g, nf, s);
  }
  return -1;
}
int launch_mng_psi_draws(int m, const pf::FilterParams& fp, const pf::MngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (m) {
    case 1: return launch<pf::MngPsi<1, true>>(fp, mp, g, nf, s);
    case 2: return launch<pf::MngPsi<2, true>>(fp, mp, g, nf, s);
    case 3: return launch<pf::MngPsi<3, true>>(fp, mp, g, nf, s);
  }
  return -1;
}
} // namespace pfl
