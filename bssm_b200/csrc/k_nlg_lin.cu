#include "filters.h"
// nlg_linear_gaussian (R/cpp_example_models.R:512-634): bootstrap filter, psi-APF and EKF particle filter
namespace pfl {
int launch_nlg_lin(int which, const pf::FilterParams& fp, const void* mp, const Geometry& g, uint32_t nf, Stream s) {
  switch (which) {
    case 0: return launch<pf::NlgLinearBsf>(fp, *static_cast<const pf::NlgParams*>(mp), g, nf, s);
    case 1: return launch<pf::NlgLinearPsi>(fp, *static_cast<const pf::NlgPsiParams*>(mp), g, nf, s);
    case 2: return launch<pf::NlgEkpf<3>>(fp, *static_cast<const pf::NlgParams*>(mp), g, nf, s);
  }
  return -1;
}
} // namespace pfl
