#include "filters.h"
namespace pfl {
int launch_ung_psi1(const pf::FilterParams& fp, const pf::UngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  return launch<pf::UngPsi1>(fp, mp, g, nf, s);
}
int occupancy_ung_psi1(const Geometry& g, int* out) { return occupancy<pf::UngPsi1>(g, out); }
} // namespace pfl
