#include "filters.h"
namespace pfl {
int launch_nlg_bsf(const pf::FilterParams& fp, const pf::NlgParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  return launch<pf::NlgGrowthBsf>(fp, mp, g, nf, s);
}
int launch_nlg_psi(const pf::FilterParams& fp, const pf::NlgPsiParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  return launch<pf::NlgGrowthPsi>(fp, mp, g, nf, s);
}
int occupancy_nlg_bsf(const Geometry& g, int* out) { return occupancy<pf::NlgGrowthBsf>(g, out); }
int occupancy_nlg_psi(const Geometry& g, int* out) { return occupancy<pf::NlgGrowthPsi>(g, out); }
} // namespace pfl
