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

// `occ` may be null (workspace sizing): then one CTA per SM is assumed.
static Plan3T plan3t(const dn_geom* g, int nf, int sms, occ3t_fn occ) {
  Plan3T pl;
  memset(&pl, 0, sizeof(pl));
  if (g->nx % 4 != 0 || g->nx < 8 || g->nx > 256) return pl;
  pl.LX = g->nx / 2;
  const int maxt = env_i3("DN_T3_THREADS", DN_T3_MAXT);
  const int step = 32 / gcd_i(pl.LX, 32);               // thread rows come in multiples of this
  int rows_max = (maxt / pl.LX) / step * step;           // element rows per tile (one thread row each)
  if (rows_max < 2) return pl;
  int TYmax = rows_max - 1;                              // owned node rows per tile (one halo element row)
  TYmax = env_i3("DN_T3_TY", TYmax);
  if (TYmax < 1) TYmax = 1;
  if (TYmax > rows_max - 1) TYmax = rows_max - 1;
  pl.nty = (g->ny + TYmax - 1) / TYmax;
  pl.TY = (g->ny + pl.nty - 1) / pl.nty;
  int rows = (pl.TY + 1 + step - 1) / step * step;
  pl.threads = rows * pl.LX;
  if (pl.threads > DN_T3_MAXT) return pl;
  // the kernel indexes its exchange buffers by thread: TY is what the tile OWNS, rows what it runs
  int S = env_i3("DN_T3_STAGES", 3);
  if (S < 2) S = 2;
  if (S > 8) S = 8;
  while (S > 2 && smem_3t(S, nf, pl.TY, g->nx, pl.threads) > (size_t)kMaxDynSmem) --S;
  pl.S = S;
  pl.smem = smem_3t(S, nf, pl.TY, g->nx, pl.threads);
  if (pl.smem > (size_t)kMaxDynSmem) return pl;
  int cps = occ ? occ(pl.threads, pl.smem) : 1;
  if (cps < 1) return pl;
  // one wave: tiles * chunks <= resident slots; chunks of >= ZCmin planes
  const long long tiles = (long long)g->batch * pl.nty;
  const long long slots = (long long)sms * cps;
  long long nzc = slots / tiles;
  if (nzc < 1) nzc = 1;
  int ZC = (int)((g->nz + nzc - 1) / nzc);
  int zmin = env_i3("DN_T3_ZCMIN", 4);
  if (zmin < 1) zmin = 1;
  if (ZC < zmin) ZC = zmin;
  ZC = env_i3("DN_T3_ZC", ZC);
  if (ZC < 1) ZC = 1;
  if (ZC > g->nz) ZC = g->nz;
  pl.ZC = ZC;
  pl.nzc = (g->nz + ZC - 1) / ZC;
  pl.grid = tiles * pl.nzc;
  pl.ok = 1;
  return pl;
}

long long plan3t_max_ctas(const dn_geom* g) {
  // workspace sizing: the default plan's grid, with head-room for the env knobs (a knob setting
  // that needs more is refused with DN_EWORKSPACE, never silently truncated)
  Plan3T pl = plan3t(g, DN_T2_MAXF, 148, nullptr);
  if (!pl.ok) return 0;
  const long long tiles = (long long)g->batch * pl.nty;
  const long long zc4 = (g->nz + 3) / 4;
  return 4 * tiles * zc4;
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
