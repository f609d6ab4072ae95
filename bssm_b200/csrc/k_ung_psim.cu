#include "filters.h"
namespace pfl {
int launch_ung_psim(int m, const pf::FilterParams& fp, const pf::UngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  if (m == 1) return launch<pf::UngPsiM<1>>(fp, mp, g, nf, s); // m = 1 with one regression part per filter
  if (m == 2) return launch<pf::UngPsiM<2>>(fp, mp, g, nf, s);
  if (m == 3) return launch<pf::UngPsiM<3>>(fp, mp, g, nf, s);
  return -1;
}
} // namespace pfl
