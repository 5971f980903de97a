// Link-time stub for the ONE reference translation unit the oracle build leaves out:
// modules/npdm/npdm_spin_adaptation.C needs boost::spirit (absent from this image) and is reached only by the
// NPDM calc types, never by the DMRG energy sweep that the oracle exists to reproduce.  Calling it aborts.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "npdm_spin_adaptation.h"

namespace SpinAdapted {
namespace Npdm {
void Npdm_spin_adaptation::npdm_set_up_linear_equations(const int, const std::string&, const std::vector<int>&,
                                                        const std::vector<double>&, Matrix&, ColumnVector&,
                                                        std::vector<std::vector<int> >&) {
  std::fprintf(stderr, "oracle build: NPDM spin adaptation is not part of the oracle (needs boost::spirit)\n");
  std::abort();
}
}  // namespace Npdm
}  // namespace SpinAdapted
