#include "filters.h"
namespace pfl {
int launch_ung_bsf_a(int m, int k, const pf::FilterParams& fp, const pf::UngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (m * 10 + k) {
    case 11: return launch<pf::UngBsf<1, 1>>(fp, mp, g, nf, s);
    case 21: return launch<pf::UngBsf<2, 1>>(fp, mp, g, nf, s);
    case 22: return launch<pf::UngBsf<2, 2>>(fp, mp, g, nf, s);
  }
  return -1;
}
// same register budget for all (m, k)
int occupancy_ung_bsf(const Geometry& g, int* out) { return occupancy<pf::UngBsf<1, 1>>(g, out); }
} // namespace pfl
