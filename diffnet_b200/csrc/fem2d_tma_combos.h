// Instantiation list of k_fem2d_tma: X(MK, HAS_NU, FK, NUMASK).
//   FK    0 no source, 1 (true) nodal source, 2 assembled load vector (no nu mask variants)
//   MK    Dirichlet set (0..3 scalar-valued masks, 4 = one mask with a nodal value field,
//         5..7 = 1..3 masks with mask_input = 0)
#pragma once
#define DN2T_COMBOS(X, MK)                                                       \
  X(MK, false, false, false) X(MK, false, true, false) X(MK, true, false, false) \
  X(MK, true, true, false) X(MK, true, false, true) X(MK, true, true, true)        \
  X(MK, false, 2, false) X(MK, true, 2, false)
// MK 5..7 (mask_input = 0, the resmin backward operator): no source term, no nu mask
#define DN2T_COMBOS_OP(X, MK) X(MK, false, false, false) X(MK, true, false, false)
#define DN2T_ALL(X)                                                                              \
  DN2T_COMBOS(X, 0) DN2T_COMBOS(X, 1) DN2T_COMBOS(X, 2) DN2T_COMBOS(X, 3) DN2T_COMBOS(X, 4)             \
  DN2T_COMBOS_OP(X, 5) DN2T_COMBOS_OP(X, 6) DN2T_COMBOS_OP(X, 7)
