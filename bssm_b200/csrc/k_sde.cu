#include "filters.h"
namespace pfl {
int launch_sde_bsf(int model_id, const pf::FilterParams& fp, const pf::SdeParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (model_id) {
    case 0: return launch<pf::SdeGbmBsf>(fp, mp, g, nf, s);
    case 1: return launch<pf::SdePoissonOuBsf>(fp, mp, g, nf, s);
  }
  return -1;
}
int occupancy_sde_bsf(const Geometry& g, int* out) { return occupancy<pf::SdeGbmBsf>(g, out); }
} // namespace pfl
