#include "filters.h"
namespace pfl {
int launch_mng_bsf(int m, int k, const pf::FilterParams& fp, const pf::MngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (m * 10 + k) {
    case 11: return launch<pf::MngBsf<1, 1>>(fp, mp, g, nf, s);
    case 21: return launch<pf::MngBsf<2, 1>>(fp, mp, g, nf, s);
    case 22: return launch<pf::MngBsf<2, 2>>(fp, mp, g, nf, s);
    case 33: return launch<pf::MngBsf<3, 3>>(fp, mp, g, nf, s);
  }
  return -1;
}
int launch_mng_psi(int m, const pf::FilterParams& fp, const pf::MngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (m) {
    case 1: return launch<pf::MngPsi<1>>(fp, mp, g, nf, s);
    case 2: return launch<pf::MngPsi<2>>(fp, mp, g, nf, s);
    case 3: return launch<pf::MngPsi<3>>(fp, mp, g, nf, s);
  }
  return -1;
}
int occupancy_mng_bsf(const Geometry& g, int* out) { return occupancy<pf::MngBsf<1, 1>>(g, out); }
} // namespace pfl
