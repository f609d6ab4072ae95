#include "filters.h"
namespace pfl {
int launch_ung_bsf_b(int m, int k, const pf::FilterParams& fp, const pf::UngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (m * 10 + k) {
    case 32: return launch<pf::UngBsf<3, 2>>(fp, mp, g, nf, s);
    case 33: return launch<pf::UngBsf<3, 3>>(fp, mp, g, nf, s);
  }
  return -1;
}
int launch_ung_bsf_draws(const pf::FilterParams& fp, const pf::UngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  return launch<pf::UngBsfDraw>(fp, mp, g, nf, s);
}
} // namespace pfl
