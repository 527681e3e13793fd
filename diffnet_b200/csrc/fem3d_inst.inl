// Included by gen/fem3d_v*_mk*.cu with DN_V and DN_MK defined.
#include "fem3d.cuh"
#include "fem3d_combos.h"
namespace dn {
#define DN_INST(V, MK, NU, FM, NMK) \
  template cudaError_t launch3d<V, MK, NU, FM, NMK>(const P3D&, dim3, dim3, size_t, cudaStream_t);
DN3D_COMBOS(DN_INST, DN_V, DN_MK)
#undef DN_INST
}  // namespace dn
