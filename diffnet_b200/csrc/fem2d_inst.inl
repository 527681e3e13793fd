// Included by gen/fem2d_v*_mk*.cu with DN_V and DN_MK defined: one translation unit per
// (V, MK) so the kernels compile in parallel.
#include "fem2d.cuh"
#include "fem2d_combos.h"
namespace dn {
#define DN_INST(V, MK, NU, FM, NMK, GN) \
  template cudaError_t launch2d<V, MK, NU, FM, NMK, GN>(const P2D&, dim3, dim3, cudaStream_t);
DN2D_COMBOS(DN_INST, DN_V, DN_MK)
#undef DN_INST
}  // namespace dn
