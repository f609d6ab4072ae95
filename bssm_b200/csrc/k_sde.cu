#include "filters.h"
namespace pfl {
int launch_sde_bsf(int model_id, const pf::FilterParams& fp, const pf::SdeParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  const bool per_filter = mp.theta_f != nullptr;
  switch (model_id) {
    case 0: return per_filter ? launch<pf::SdeGbmBsfDraws>(fp, mp, g, nf, s) : launch<pf::SdeGbmBsf>(fp, mp, g, nf, s);
    case 1: return per_filter ? launch<pf::SdePoissonOuBsfDraws>(fp, mp, g, nf, s)
                              : launch<pf::SdePoissonOuBsf>(fp, mp, g, nf, s);
  }
  return -1;
}
int occupancy_sde_bsf(const Geometry& g, int* out) { return occupancy<pf::SdeGbmBsf>(g, out); }
} // namespace pfl
