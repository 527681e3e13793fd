// Launch planning and runtime -> compile-time dispatch of the streaming 3-D kernel family
// (instantiated in gen/fem3dt_mk*.cu).
#include <cstdlib>
#include <cstring>

#include "fem3d.cuh"
#include "fem3d_tma.cuh"
#include "fem3d_tma_combos.h"

namespace dn {
#define DN_EXT(MK, NU, F, NMK)                                                                        \
  extern template cudaError_t launch3t<MK, NU, F, NMK>(const P3T&, dim3, dim3, size_t, cudaStream_t); \
  extern template int occ3t<MK, NU, F, NMK>(int, size_t);
DN3T_ALL(DN_EXT)
#undef DN_EXT

launch3t_fn get_launch3t(int MK, int NU, int F, int NUMASK) {
#define DN_CASE(MK_, NU_, F_, NMK_)                                                \
  if (MK == MK_ && NU == (int)NU_ && F == (int)F_ && NUMASK == (int)NMK_)          \
    return &launch3t<MK_, NU_, F_, NMK_>;
  DN3T_ALL(DN_CASE)
#undef DN_CASE
  return nullptr;
}

occ3t_fn get_occ3t(int MK, int NU, int F, int NUMASK) {
#define DN_CASE(MK_, NU_, F_, NMK_)                                                \
  if (MK == MK_ && NU == (int)NU_ && F == (int)F_ && NUMASK == (int)NMK_)          \
    return &occ3t<MK_, NU_, F_, NMK_>;
  DN3T_ALL(DN_CASE)
#undef DN_CASE
  return nullptr;
}

static int env_i3(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

struct Plan3T { int ok, LX, TY, threads, nty, ZC, nzc, S; long long grid; size_t smem; };

static size_t smem_3t(int S, int nf, int TY, int nx, int threads) {
  return (size_t)S * nf * (TY + 2) * nx * 4 + 16 + (size_t)S * 8 + (size_t)2 * 4 * threads * 4;
}

// Launch shape search.  Candidates: thread rows per tile (threads = rows * LX <= 512) x number of
// z-chunks.  Cost model (the kernel is FP32-pipe bound, CTAs on one SM share the pipe):
//   SM time ~ ceil(grid / SMs) * rows * (ZC + 1)   [thread-layers executed by the busiest SM]
// with a penalty when fewer than 12 warps per SM are resident (latency hiding) and a preference
// for a single wave.  `occ` may be null (workspace sizing): then the register bound is assumed.
static Plan3T plan3t(const dn_geom* g, int nf, int sms, occ3t_fn occ) {
  Plan3T best;
  memset(&best, 0, sizeof(best));
  if (g->nx % 4 != 0 || g->nx < 8 || g->nx > 256) return best;
  const int LX = g->nx / 2;
  const int maxt = env_i3("DN_T3_THREADS", DN_T3_MAXT);
  const int step = 32 / gcd_i(LX, 32);               // thread rows come in multiples of this
  const int rows_cap = (maxt / LX) / step * step;
  if (rows_cap < 2) return best;
  const int ty_forced = env_i3("DN_T3_TY", 0), zc_forced = env_i3("DN_T3_ZC", 0);
  int zmin = env_i3("DN_T3_ZCMIN", 4);
  if (zmin < 1) zmin = 1;
  int S0 = env_i3("DN_T3_STAGES", 3);
  if (S0 < 2) S0 = 2;
  if (S0 > 8) S0 = 8;
  double best_cost = 0.0;
  for (int rows = step; rows <= rows_cap; rows += step) {
    if (rows < 2) continue;
    int TYmax = rows - 1;
    if (ty_forced > 0) { if (ty_forced > TYmax) continue; TYmax = ty_forced; }
    const int nty = (g->ny + TYmax - 1) / TYmax;
    const int TY = (g->ny + nty - 1) / nty;
    if (ty_forced <= 0 && (TY + 1 + step - 1) / step * step != rows) continue;   // a smaller block covers it
    const int threads = rows * LX;
    int S = S0;
    while (S > 2 && smem_3t(S, nf, TY, g->nx, threads) > (size_t)kMaxDynSmem) --S;
    const size_t smem = smem_3t(S, nf, TY, g->nx, threads);
    if (smem > (size_t)kMaxDynSmem) continue;
    int cps = occ ? occ(threads, smem) : (int)(65536 / (threads * 128));
    if (cps < 1) continue;
    const long long tiles = (long long)g->batch * nty;
    const int nzc_max = zc_forced > 0 ? 1 : (g->nz + zmin - 1) / zmin;
    for (int nzc_try = 1; nzc_try <= nzc_max; ++nzc_try) {
      int ZC = zc_forced > 0 ? zc_forced : (g->nz + nzc_try - 1) / nzc_try;
      if (ZC > g->nz) ZC = g->nz;
      if (ZC < 1) ZC = 1;
      const int nzc = (g->nz + ZC - 1) / ZC;
      if (zc_forced <= 0 && nzc != nzc_try) continue;
      const long long grid = tiles * nzc;
      const long long per_sm = (grid + sms - 1) / sms;               // CTAs the busiest SM runs
      const long long resident = per_sm < cps ? per_sm : cps;
      const double waves = (double)((per_sm + cps - 1) / cps);
      const double warps = (double)resident * threads / 32.0;
      double cost = (double)per_sm * rows * (ZC + 1);
      if (warps < 12.0) cost *= 1.0 + 0.4 * (12.0 - warps) / 12.0;
      cost *= 1.0 + 0.05 * (waves - 1.0);
      if (!best.ok || cost < best_cost) {
        best_cost = cost;
        best.ok = 1; best.LX = LX; best.TY = TY; best.threads = threads; best.nty = nty;
        best.ZC = ZC; best.nzc = nzc; best.S = S; best.grid = grid; best.smem = smem;
      }
      if (grid > 64LL * sms) break;                                   // finer chunks only add seams
    }
  }
  return best;
}

long long plan3t_max_ctas(const dn_geom* g) {
  // workspace sizing: the default plan's grid, with head-room for the env knobs (a knob setting
  // that needs more is refused with DN_EWORKSPACE, never silently truncated)
  if (g->nx % 4 != 0 || g->nx < 8 || g->nx > 256) return 0;
  // the finest shape the planner may pick: tiles owning one node row... bounded in practice by
  // one thread-row step and chunks of DN_T3_ZCMIN (>= 1, default 4) planes
  const int LX = g->nx / 2;
  const int step = 32 / gcd_i(LX, 32);
  const int ty_min = step > 1 ? step - 1 : 1;
  const long long nty = (g->ny + ty_min - 1) / ty_min;
  const long long nzc = (g->nz + 3) / 4;
  return (long long)g->batch * nty * nzc;
}

int run3t(const Field& u, const Field& nu, const Field& f, const Field& fgp, const Field& numask,
          const Mask* mk, int nmasks, int MK, const Consts& k, bool vec4, const dn_geom* g, float* grad,
          int mode, int mask_input, void* workspace, size_t wsb, double* loss_out, float* loss_f32,
          void* stream, int sms, bool* handled) {
  *handled = false;
  const char* path = getenv("DN_3D_PATH");
  if (path && !strcmp(path, "tile")) return DN_OK;
  if (!vec4 || fgp.p || !mask_input || ((uintptr_t)grad % 16 != 0)) return DN_OK;
  const int NU = nu.p ? 1 : 0, F = f.p ? 1 : 0, NMK = numask.p ? 1 : 0;
  launch3t_fn fn = get_launch3t(MK, NU, F, NMK);
  occ3t_fn occ = get_occ3t(MK, NU, F, NMK);
  if (!fn || !occ) return DN_OK;
  P3T p;
  memset(&p, 0, sizeof(p));
  int nf = 0;
  p.fld[nf++] = u;
  if (NU) p.fld[nf++] = nu;
  if (F) p.fld[nf++] = f;
  if (NMK) p.fld[nf++] = numask;
  for (int i = 0; i < nmasks; ++i) { p.fld[nf++] = mk[i].m; p.mval[i] = mk[i].v; }
  if (MK == 4) p.fld[nf++] = mk[0].vf;
  for (int i = 0; i < nf; ++i)
    if (p.fld[i].sy != g->nx) return DN_OK;      // bulk copies take whole runs of rows
  Plan3T pl = plan3t(g, nf, sms, occ);
  if (!pl.ok) { cudaGetLastError(); return DN_OK; }
  if (pl.grid > 0x7fffffffLL) return DN_OK;
  const size_t need = 64 + 8 * (size_t)pl.grid;
  if (!workspace || wsb < need) return fail(DN_EWORKSPACE, "workspace too small: %zu < %zu", wsb, need);
  if ((uintptr_t)workspace % 16) return fail(DN_EWORKSPACE, "workspace must be 16-byte aligned");
  p.nf = nf;
  p.B = g->batch; p.nx = g->nx; p.ny = g->ny; p.nz = g->nz;
  p.LX = pl.LX; p.TY = pl.TY; p.nty = pl.nty; p.ZC = pl.ZC; p.nzc = pl.nzc; p.S = pl.S;
  if (g->z_own_hi > g->z_own_lo) { p.zloss_lo = g->z_own_lo; p.zloss_hi = g->z_own_hi; }
  else { p.zloss_lo = 0; p.zloss_hi = g->nz; }
  const float t = k.t;
  auto pr = [](float v) { return make_float2(v, v); };
  p.k3.kx = pr(k.kx); p.k3.ky = pr(k.ky); p.k3.kz = pr(k.kz);
  p.k3.kxt = pr(k.kx * t); p.k3.kyt = pr(k.ky * t); p.k3.kzt = pr(k.kz * t);
  p.k3.kxtt = pr(k.kx * t * t); p.k3.kytt = pr(k.ky * t * t); p.k3.kztt = pr(k.kz * t * t);
  p.k3.t = pr(t); p.k3.tt = pr(t * t);
  p.k3.nkf = pr(-k.kf); p.k3.nkft = pr(-k.kf * t); p.k3.nkftt = pr(-k.kf * t * t);
  p.k3.nkfttt = pr(-k.kf * t * t * t);
  p.k3.c0x = pr(8.f * k.kx); p.k3.c0y = pr(8.f * k.ky); p.k3.c0z = pr(8.f * k.kz);
  p.grad = grad;
  p.red.counter = (unsigned int*)workspace;
  p.red.partials = (double*)((char*)workspace + 64);
  p.red.loss_out = loss_out; p.red.loss_f32 = loss_f32;
  p.mode = mode;
  *handled = true;
  return check_cuda(fn(p, dim3((unsigned)pl.grid), dim3(pl.threads), pl.smem, (cudaStream_t)stream),
                    "fem3d_tma launch");
}

}  // namespace dn
