// Runtime -> compile-time dispatch of the streaming 2-D kernel family (gen/fem2dt_mk*.cu).
#include "fem2d_tma.cuh"
#include "fem2d_tma_combos.h"
namespace dn {
#define DN_EXT(MK, NU, F, NMK)                                                                        \
  extern template cudaError_t launch2t<MK, NU, F, NMK>(const P2T&, dim3, dim3, size_t, cudaStream_t); \
  extern template int occ2t<MK, NU, F, NMK>(int, size_t);
DN2T_ALL(DN_EXT)
#undef DN_EXT

launch2t_fn get_launch2t(int MK, int NU, int F, int NUMASK) {
#define DN_CASE(MK_, NU_, F_, NMK_)                                                          \
  if (MK == MK_ && NU == (int)NU_ && F == (int)F_ && NUMASK == (int)NMK_)     \
    return &launch2t<MK_, NU_, F_, NMK_>;
  DN2T_ALL(DN_CASE)
#undef DN_CASE
  return nullptr;
}

occ2t_fn get_occ2t(int MK, int NU, int F, int NUMASK) {
#define DN_CASE(MK_, NU_, F_, NMK_)                                                          \
  if (MK == MK_ && NU == (int)NU_ && F == (int)F_ && NUMASK == (int)NMK_)     \
    return &occ2t<MK_, NU_, F_, NMK_>;
  DN2T_ALL(DN_CASE)
#undef DN_CASE
  return nullptr;
}
}  // namespace dn
