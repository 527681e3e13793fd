// gp_eval.cu -- see gp_eval.cuh.
#include "gp_eval.cuh"

namespace dn {

// out[b, G, (k,) j, i] = sum_a T[G][a] * in[b, (k+kb,) j+jb, i+ib],  G = (kg*n + jg)*n + ig.
// One thread per element, x fastest (coalesced loads and stores); each nodal value is read
// by the 4 (8) elements around it through L1.
template <int NSD>
__global__ void __launch_bounds__(256) k_gp_eval(Field in, int B, int nx, int ny, int nz,
                                                 GpTables tb, float* __restrict__ out) {
  const int nelx = nx - 1, nely = ny - 1, nelz = (NSD == 3) ? nz - 1 : 1;
  const long long nel = (long long)nelx * nely * nelz;
  const long long total = nel * B;
  const int n = tb.n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % nelx);
    long long r = idx / nelx;
    const int j = (int)(r % nely);
    r /= nely;
    const int k = (int)(r % nelz);
    const int b = (int)(r / nelz);
    const float* base = in.p + (long long)b * in.sb + (long long)k * (NSD == 3 ? in.sz : 0) +
                        (long long)j * in.sy + i;
    float v[2][2][2];
#pragma unroll
    for (int kb = 0; kb < (NSD == 3 ? 2 : 1); ++kb)
#pragma unroll
      for (int jb = 0; jb < 2; ++jb)
#pragma unroll
        for (int ib = 0; ib < 2; ++ib)
          v[kb][jb][ib] = __ldg(base + (long long)kb * in.sz + (long long)jb * in.sy + ib);
    float* o = out + (long long)b * (NSD == 3 ? n * n * n : n * n) * nel +
               ((long long)k * nely + j) * nelx + i;
    for (int kg = 0; kg < (NSD == 3 ? n : 1); ++kg) {
      // collapse z
      float w[2][2];
#pragma unroll
      for (int jb = 0; jb < 2; ++jb)
#pragma unroll
        for (int ib = 0; ib < 2; ++ib)
          w[jb][ib] = (NSD == 3) ? tb.c[2][kg][0] * v[0][jb][ib] + tb.c[2][kg][1] * v[1][jb][ib]
                                 : v[0][jb][ib];
      for (int jg = 0; jg < n; ++jg) {
        const float r0 = tb.c[1][jg][0] * w[0][0] + tb.c[1][jg][1] * w[1][0];
        const float r1 = tb.c[1][jg][0] * w[0][1] + tb.c[1][jg][1] * w[1][1];
        for (int ig = 0; ig < n; ++ig) {
          const int G = (kg * n + jg) * n + ig;
          o[(long long)G * nel] = tb.c[0][ig][0] * r0 + tb.c[0][ig][1] * r1;
        }
      }
    }
  }
}

// gin[b, node] = sum over the elements e around the node and all Gauss points G of
// T[G][a(node in e)] * gout[b, G, e]  (gather form: no atomics, deterministic).
template <int NSD>
__global__ void __launch_bounds__(256) k_gp_eval_adj(const float* __restrict__ gout, int B, int nx,
                                                     int ny, int nz, GpTables tb,
                                                     float* __restrict__ gin) {
  const int nelx = nx - 1, nely = ny - 1, nelz = (NSD == 3) ? nz - 1 : 1;
  const int nzz = (NSD == 3) ? nz : 1;
  const long long nel = (long long)nelx * nely * nelz;
  const long long total = (long long)B * nzz * ny * nx;
  const int n = tb.n;
  const int ngp = (NSD == 3) ? n * n * n : n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % nx);
    long long r = idx / nx;
    const int y = (int)(r % ny);
    r /= ny;
    const int z = (int)(r % nzz);
    const int b = (int)(r / nzz);
    const float* gb = gout + (long long)b * ngp * nel;
    float acc = 0.f;
    for (int kb = 0; kb < (NSD == 3 ? 2 : 1); ++kb) {
      const int ek = z - kb;
      if (NSD == 3 && (ek < 0 || ek >= nelz)) continue;
      for (int jb = 0; jb < 2; ++jb) {
        const int ej = y - jb;
        if (ej < 0 || ej >= nely) continue;
        for (int ib = 0; ib < 2; ++ib) {
          const int ei = x - ib;
          if (ei < 0 || ei >= nelx) continue;
          const float* ge = gb + ((long long)(NSD == 3 ? ek : 0) * nely + ej) * nelx + ei;
          for (int kg = 0; kg < (NSD == 3 ? n : 1); ++kg) {
            const float cz = (NSD == 3) ? tb.c[2][kg][kb] : 1.f;
            for (int jg = 0; jg < n; ++jg) {
              const float czy = cz * tb.c[1][jg][jb];
              float s = 0.f;
              for (int ig = 0; ig < n; ++ig)
                s += tb.c[0][ig][ib] * __ldg(ge + (long long)((kg * n + jg) * n + ig) * nel);
              acc += czy * s;
            }
          }
        }
      }
    }
    gin[idx] = acc;
  }
}

__global__ void __launch_bounds__(256) k_scale(float* __restrict__ x, size_t n,
                                               const float* __restrict__ factor) {
  const float f = __ldg(factor);
  if (f == 1.0f) return;   // loss.backward() with the default grad_output: nothing to do
  const size_t n4 = (((uintptr_t)x % 16) == 0) ? n / 4 : 0;
  float4* x4 = reinterpret_cast<float4*>(x);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = x4[i];
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    x4[i] = v;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] *= f;
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

cudaError_t launch_gp_eval(Field in, int B, int nx, int ny, int nz, int nsd, const GpTables& tb,
                           float* out, cudaStream_t s) {
  const long long total = (long long)B * (nx - 1) * (ny - 1) * (nsd == 3 ? nz - 1 : 1);
  if (nsd == 2)
    k_gp_eval<2><<<grid_for(total, 256), 256, 0, s>>>(in, B, nx, ny, 1, tb, out);
  else
    k_gp_eval<3><<<grid_for(total, 256), 256, 0, s>>>(in, B, nx, ny, nz, tb, out);
  return cudaGetLastError();
}

cudaError_t launch_gp_eval_adj(const float* gout, int B, int nx, int ny, int nz, int nsd,
                               const GpTables& tb, float* gin, cudaStream_t s) {
  const long long total = (long long)B * nx * ny * (nsd == 3 ? nz : 1);
  if (nsd == 2)
    k_gp_eval_adj<2><<<grid_for(total, 256), 256, 0, s>>>(gout, B, nx, ny, 1, tb, gin);
  else
    k_gp_eval_adj<3><<<grid_for(total, 256), 256, 0, s>>>(gout, B, nx, ny, nz, tb, gin);
  return cudaGetLastError();
}

cudaError_t launch_scale(float* x, size_t n, const float* factor_dev, cudaStream_t s) {
  k_scale<<<grid_for((long long)(n + 3) / 4, 256), 256, 0, s>>>(x, n, factor_dev);
  return cudaGetLastError();
}

}  // namespace dn
