// Included by gen/fem3dt_mk*.cu with DN_MK defined: one translation unit per Dirichlet set so
// the kernels compile in parallel.
#include "fem3d_tma.cuh"
#include "fem3d_tma_combos.h"
namespace dn {
#define DN_INST(MK, NU, F, NMK, LK)                                                                 \
  template cudaError_t launch3t<MK, NU, F, NMK, LK>(const P3T&, dim3, dim3, size_t, cudaStream_t);  \
  template int occ3t<MK, NU, F, NMK, LK>(int, size_t);
#if DN_MK >= 5
DN3T_COMBOS_OP(DN_INST, DN_MK)
#else
DN3T_COMBOS(DN_INST, DN_MK)
#endif
#undef DN_INST
}  // namespace dn
