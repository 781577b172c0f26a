// tcgen05 / TMEM 3xTF32 flash-style phi for large N (placeholder until the kernel lands:
// reports "unsupported" so that svgd_large.cu's SIMT tiles are used).
#include "common.cuh"

namespace dust {
bool phi_tc_supported(const dust_phi_args*) { return false; }
int phi_tc(const dust_phi_args*, cudaStream_t) {
  set_error("tcgen05 phi kernel not built");
  return DUST_ERR_UNSUPPORTED;
}
}  // namespace dust
