#include "filters.h"
// ssm_nlg bootstrap filter with one theta per filter: is_correction_bsf, src/approx_mcmc.cpp:447-518
namespace pfl {
int launch_nlg_bsf_draws(int model_id, const pf::FilterParams& fp, const pf::NlgParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (model_id) {
    case 0: return launch<pf::NlgGrowthBsfDraws>(fp, mp, g, nf, s);
    case 1: return launch<pf::NlgArExpBsfDraws>(fp, mp, g, nf, s);
    case 2: return launch<pf::NlgSinExpBsfDraws>(fp, mp, g, nf, s);
    case 3: return launch<pf::NlgLinearBsfDraws>(fp, mp, g, nf, s);
  }
  return -1;
}
} // namespace pfl
