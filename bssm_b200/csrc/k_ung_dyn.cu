#include "filters.h"
#include "models_ung_dyn.cuh"
// run-time state dimension (m <= 16): two particles per thread only -- the per-particle arrays of these
// policies live in local memory, more particles per thread only add to it
namespace pfl {
template <class Model>
static int launch_p2(const pf::FilterParams& fp, const pf::UngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  if (g.P != 2) return -1;
  return g.block <= 512 ? launch_V<Model, 2, PF_SMALL_MAXT, PF_SMALL_MINB>(fp, mp, g, nf, s)
                        : launch_V<Model, 2, 1024, 1>(fp, mp, g, nf, s);
}
int launch_ung_bsf_dyn(const pf::FilterParams& fp, const pf::UngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  return launch_p2<pf::UngBsfDyn<16>>(fp, mp, g, nf, s);
}
int launch_ung_psi_dyn(const pf::FilterParams& fp, const pf::UngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  return launch_p2<pf::UngPsiDyn<16>>(fp, mp, g, nf, s);
}
// ssm_mng at run-time m (also the parameter-set jobs: the per-filter strides are honoured)
template <class Model>
static int launch_p2m(const pf::FilterParams& fp, const pf::MngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  if (g.P != 2) return -1;
  return g.block <= 512 ? launch_V<Model, 2, PF_SMALL_MAXT, PF_SMALL_MINB>(fp, mp, g, nf, s)
                        : launch_V<Model, 2, 1024, 1>(fp, mp, g, nf, s);
}
int launch_mng_bsf_dyn(const pf::FilterParams& fp, const pf::MngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  return launch_p2m<pf::MngBsfDyn<16>>(fp, mp, g, nf, s);
}
int launch_mng_psi_dyn(const pf::FilterParams& fp, const pf::MngParams& mp, const Geometry& g, uint32_t nf, Stream s) {
  return launch_p2m<pf::MngPsiDyn<16>>(fp, mp, g, nf, s);
}
int occupancy_mng_dyn(const Geometry& g, int* out) {
  return g.block <= 512 ? occupancy_V<pf::MngBsfDyn<16>, 2, PF_SMALL_MAXT, PF_SMALL_MINB>(g, out)
                        : occupancy_V<pf::MngBsfDyn<16>, 2, 1024, 1>(g, out);
}
int occupancy_ung_dyn(const Geometry& g, int* out) {
  return g.block <= 512 ? occupancy_V<pf::UngBsfDyn<16>, 2, PF_SMALL_MAXT, PF_SMALL_MINB>(g, out)
                        : occupancy_V<pf::UngBsfDyn<16>, 2, 1024, 1>(g, out);
}
} // namespace pfl
