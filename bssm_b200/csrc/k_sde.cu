#include "filters.h"
namespace pfl {
int launch_sde_bsf(const pf::FilterParams& fp, const pf::SdeParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  return launch<pf::SdeGbmBsf>(fp, mp, g, nf, s);
}
int occupancy_sde_bsf(const Geometry& g, int* out) { return occupancy<pf::SdeGbmBsf>(g, out); }
} // namespace pfl
