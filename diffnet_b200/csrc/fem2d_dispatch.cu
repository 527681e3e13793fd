// Runtime -> compile-time dispatch of the 2-D kernel family (instantiated in gen/fem2d_*.cu).
#include "fem2d.cuh"
#include "fem2d_combos.h"
namespace dn {
#define DN_EXT(V, MK, NU, FM, NMK, GN) \
  extern template cudaError_t launch2d<V, MK, NU, FM, NMK, GN>(const P2D&, dim3, dim3, cudaStream_t);
DN2D_ALL(DN_EXT)
#undef DN_EXT

launch2d_fn get_launch2d(int V, int MK, int NU, int FM, int NUMASK, int GN) {
#define DN_CASE(V_, MK_, NU_, FM_, NMK_, GN_)                                          \
  if (V == V_ && MK == MK_ && NU == (int)NU_ && FM == FM_ && NUMASK == (int)NMK_ &&   \
      GN == (int)GN_)                                                                  \
    return &launch2d<V_, MK_, NU_, FM_, NMK_, GN_>;
  DN2D_ALL(DN_CASE)
#undef DN_CASE
  return nullptr;
}
}  // namespace dn
