// Instantiation list of k_fem3d_tma: X(MK, NUK, HAS_F, NUMASK).
//   MK    Dirichlet set (0..3 scalar-valued masks, 4 = one mask with a nodal value field,
//         5..7 = 1..3 masks with mask_input = 0)
//   NUK   0: nu == 1, 1: nodal nu, 2: nodal nu on an isotropic grid (hx == hy == hz: k applied per node)
#pragma once
#define DN3T_COMBOS(X, MK)                                                          \
  X(MK, 0, false, false) X(MK, 0, true, false) X(MK, 1, false, false)               \
  X(MK, 1, true, false) X(MK, 1, false, true) X(MK, 1, true, true)                  \
  X(MK, 2, false, false) X(MK, 2, true, false) X(MK, 2, false, true) X(MK, 2, true, true)
// MK 5..7 (mask_input = 0, the resmin backward operator): no source term, no nu mask
#define DN3T_COMBOS_OP(X, MK) X(MK, 0, false, false) X(MK, 1, false, false) X(MK, 2, false, false)
#define DN3T_ALL(X)                                                                              \
  DN3T_COMBOS(X, 0) DN3T_COMBOS(X, 1) DN3T_COMBOS(X, 2) DN3T_COMBOS(X, 3) DN3T_COMBOS(X, 4)             \
  DN3T_COMBOS_OP(X, 5) DN3T_COMBOS_OP(X, 6) DN3T_COMBOS_OP(X, 7)
