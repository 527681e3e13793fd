// producers.cu -- device-side producers of the loss path's INPUT tensors, so that a training
// step needs no per-step host->device copy of fp32 fields.  Each kernel restates what a reference dataset class
// builds on the host with numpy and writes the same (B, 3, [D,] H, W) `inputs` layout [nu | domain, bc1, bc2]
// plus the (B, 1, ...) zero forcing:
//   * KL log-normal diffusivity + column masks   DiffNet/gen_input_calc.py:74-181, datasets/parametric/klsum.py:10-46
//   * image masks (immersed-boundary inputs)     datasets/parametric/images.py:9-49 (ImageIMBack)
//   * raw voxel files                            datasets/single_instances/voxels.py:8-61 (load_raw, VoxelIMBackRAW)
// Pure streaming writes: one thread per 4 consecutive nodes (x fastest), 16-byte stores; the KL sum is evaluated in
// fp64 from 1-D factor tables (the reference works in fp64 and rounds to fp32 at the end).
#include <cuda_runtime.h>
#include <stdint.h>

#include "dn_common.cuh"

namespace dn {

constexpr int kKlMaxTerms = 10;      // the reference's omega tables hold 10 roots; its sums use the first 6

// 1-D factor tables: phi[d][i][n] = eta w_i cos(w_i t_n) + sin(w_i t_n),  t_n = n / (size_d - 1)   (np.linspace(0, 1, size))
struct KlParams {
  double omega[kKlMaxTerms];
  double sqrt_lambda[kKlMaxTerms];
  double eta;
  int nterms;
};

__global__ void k_kl_tables(KlParams kp, int size, double* __restrict__ phi /* [nterms][size] */) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= size) return;
  // np.linspace(0, 1, size): start + n * step with step = 1 / (size - 1); the last point is set to 1 exactly
  const double step = 1.0 / (double)(size - 1);
  const double t = (n == size - 1) ? 1.0 : (double)n * step;
  for (int i = 0; i < kp.nterms; ++i) {
    const double w = kp.omega[i];
    // no fma contraction: the reference rounds every product and sum (numpy fp64)
    phi[(size_t)i * size + n] = __dadd_rn(__dmul_rn(__dmul_rn(kp.eta, w), cos(__dmul_rn(w, t))), sin(__dmul_rn(w, t)));
  }
}

// inputs[b, 0] = exp(sum_i ((((a_bi sqrt(lx_i)) sqrt(ly_i)) phi_i(x)) phi_i(y))), inputs[b, 1] = [x == 0], inputs[b, 2] = [x == last]
// (2-D: tensor (H, W), x = last axis -- np.meshgrid(x, y), gen_input_calc.py:117-122).  forcing[b, 0] = 0.
__global__ void __launch_bounds__(256) k_kl_inputs_2d(const double* __restrict__ coeffs, KlParams kp, int B, int size,
                                                      const double* __restrict__ phi, float* __restrict__ inputs,
                                                      float* __restrict__ forcing) {
  const int q4 = size / 4;                                   // float4 per row
  const long long per = (long long)size * q4;
  const long long total = (long long)B * per;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per);
    const long long r = idx - (long long)b * per;
    const int j = (int)(r / q4), i0 = 4 * (int)(r - (long long)j * q4);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = 0; t < kp.nterms; ++t) {
      const double a = __dmul_rn(__dmul_rn(coeffs[(size_t)b * kp.nterms + t], kp.sqrt_lambda[t]), kp.sqrt_lambda[t]);
      const double py = phi[(size_t)t * size + j];
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] = __dadd_rn(s[e], __dmul_rn(__dmul_rn(a, phi[(size_t)t * size + i0 + e]), py));
    }
    const size_t plane = (size_t)size * size;
    float* o = inputs + (size_t)b * 3 * plane + (size_t)j * size + i0;
    *reinterpret_cast<float4*>(o) = make_float4((float)exp(s[0]), (float)exp(s[1]), (float)exp(s[2]), (float)exp(s[3]));
    *reinterpret_cast<float4*>(o + plane) = make_float4(i0 == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(o + 2 * plane) = make_float4(0.f, 0.f, 0.f, i0 + 4 == size ? 1.f : 0.f);
    if (forcing) *reinterpret_cast<float4*>(forcing + (size_t)b * plane + (size_t)j * size + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// 3-D: np.meshgrid(x, y, z) has shape (ny, nx, nz) -- the tensor's axes are (y, x, z), its LAST axis carries z
// (gen_input_calc.py:124-130); the product order is ((((a sx) sy) sz) phi(x)) phi(y)) phi(z)  (:112).
__global__ void __launch_bounds__(256) k_kl_field_3d(const double* __restrict__ coeffs, KlParams kp, int B, int size,
                                                     const double* __restrict__ phi, float* __restrict__ out,
                                                     long long out_stride_b) {
  const int q4 = size / 4;
  const long long per = (long long)size * size * q4;
  const long long total = (long long)B * per;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per);
    long long r = idx - (long long)b * per;
    const int a2 = 4 * (int)(r % q4);          // last axis  (z coordinate)
    r /= q4;
    const int a1 = (int)(r % size);            // middle axis (x coordinate)
    const int a0 = (int)(r / size);            // first axis  (y coordinate)
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = 0; t < kp.nterms; ++t) {
      const double a = __dmul_rn(__dmul_rn(__dmul_rn(coeffs[(size_t)b * kp.nterms + t], kp.sqrt_lambda[t]), kp.sqrt_lambda[t]), kp.sqrt_lambda[t]);
      const double pxy = __dmul_rn(__dmul_rn(a, phi[(size_t)t * size + a1]), phi[(size_t)t * size + a0]);
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] = __dadd_rn(s[e], __dmul_rn(pxy, phi[(size_t)t * size + a2 + e]));
    }
    float* o = out + (size_t)b * out_stride_b + ((size_t)a0 * size + a1) * size + a2;
    *reinterpret_cast<float4*>(o) = make_float4((float)exp(s[0]), (float)exp(s[1]), (float)exp(s[2]), (float)exp(s[3]));
  }
}

// ImageIMBack (images.py:9-49): img8 = the greyscale image bytes; domain = 1 - (img > 0), bc1 = (img > 0),
// bc2 = the four edges; forcing = 0.  H, W arbitrary (scalar stores; one thread per pixel).
__global__ void __launch_bounds__(256) k_image_inputs(const uint8_t* __restrict__ img, int B, int H, int W,
                                                      float* __restrict__ inputs, float* __restrict__ forcing) {
  const long long plane = (long long)H * W, total = (long long)B * plane;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / plane);
    const long long r = idx - (long long)b * plane;
    const int j = (int)(r / W), i = (int)(r - (long long)j * W);
    const float obj = img[idx] > 0 ? 1.f : 0.f;
    float* o = inputs + (size_t)b * 3 * plane + r;
    o[0] = 1.f - obj;
    o[plane] = obj;
    o[2 * plane] = (i == 0 || i == W - 1 || j == 0 || j == H - 1) ? 1.f : 0.f;
    if (forcing) forcing[idx] = 0.f;
  }
}

// VoxelIMBackRAW (voxels.py:8-61): raw = the bytes of <name>inouts.raw, Fortran order over numDiv = (d0, d1, d2):
// vox[i0, i1, i2] = (raw[i0 + d0 (i1 + d1 i2)] / 254.0 > 0.25); domain = 1 everywhere except
// domain[off + i0, off + i1, off + i2] = 1 - vox; bc1 = 1 - domain; bc2 = the six faces.  One sample (1, 3, N, N, N).
__global__ void __launch_bounds__(256) k_voxel_inputs(const uint8_t* __restrict__ raw, int d0, int d1, int d2, int N,
                                                      int off, float* __restrict__ inputs, float* __restrict__ forcing) {
  const long long vol = (long long)N * N * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < vol;
       idx += (long long)gridDim.x * blockDim.x) {
    const int a2 = (int)(idx % N);
    const long long r = idx / N;
    const int a1 = (int)(r % N), a0 = (int)(r / N);
    const int i0 = a0 - off, i1 = a1 - off, i2 = a2 - off;
    float vox = 0.f;
    if (i0 >= 0 && i0 < d0 && i1 >= 0 && i1 < d1 && i2 >= 0 && i2 < d2)
      vox = ((double)raw[(size_t)i0 + (size_t)d0 * ((size_t)i1 + (size_t)d1 * i2)] / 254.0 > 0.25) ? 1.f : 0.f;
    inputs[idx] = 1.f - vox;
    inputs[vol + idx] = vox;
    inputs[2 * vol + idx] = (a0 == 0 || a0 == N - 1 || a1 == 0 || a1 == N - 1 || a2 == 0 || a2 == N - 1) ? 1.f : 0.f;
    if (forcing) forcing[idx] = 0.f;
  }
}

// Synthetic immersed geometries for the bench / tests (stand-ins for the reference's image and SIMP datasets):
// star-shaped silhouettes r(theta) = r0 (1 + sum_k a_k cos(k theta + p_k)), k = 2..5, and unions of boxes.
struct StarParams { float cx, cy, r0, a[4], ph[4]; };
__global__ void __launch_bounds__(256) k_star_inputs(const StarParams* __restrict__ sp, int B, int N,
                                                     float* __restrict__ inputs, float* __restrict__ forcing) {
  const long long plane = (long long)N * N, total = (long long)B * plane;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / plane);
    const long long r = idx - (long long)b * plane;
    const int j = (int)(r / N), i = (int)(r - (long long)j * N);
    const StarParams s = sp[b];
    const float x = -1.f + 2.f * (float)i / (float)(N - 1) - s.cx, y = -1.f + 2.f * (float)j / (float)(N - 1) - s.cy;
    const float th = atan2f(y, x);
    float rad = s.r0;
#pragma unroll
    for (int k = 0; k < 4; ++k) rad += s.r0 * s.a[k] * cosf((float)(k + 2) * th + s.ph[k]);
    const float obj = (sqrtf(x * x + y * y) < rad) ? 1.f : 0.f;
    float* o = inputs + (size_t)b * 3 * plane + r;
    o[0] = 1.f - obj;
    o[plane] = obj;
    o[2 * plane] = (i == 0 || i == N - 1 || j == 0 || j == N - 1) ? 1.f : 0.f;
    if (forcing) forcing[idx] = 0.f;
  }
}

struct BoxParams { int n; int lo[3][3], hi[3][3]; };   // up to 3 boxes per sample, [box][axis], half-open [lo, hi)
__global__ void __launch_bounds__(256) k_box_masks_3d(const BoxParams* __restrict__ bp, int B, int N,
                                                      float* __restrict__ source, float* __restrict__ sink,
                                                      float* __restrict__ forcing) {
  const long long vol = (long long)N * N * N, total = (long long)B * vol;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / vol);
    long long r = idx - (long long)b * vol;
    const int a2 = (int)(r % N);
    r /= N;
    const int a1 = (int)(r % N), a0 = (int)(r / N);
    const BoxParams p = bp[b];
    bool in = false;
    for (int q = 0; q < p.n; ++q)
      in = in || (a0 >= p.lo[q][0] && a0 < p.hi[q][0] && a1 >= p.lo[q][1] && a1 < p.hi[q][1] && a2 >= p.lo[q][2] && a2 < p.hi[q][2]);
    source[idx] = in ? 1.f : 0.f;
    sink[idx] = (a0 == 0 || a0 == N - 1 || a1 == 0 || a1 == N - 1 || a2 == 0 || a2 == N - 1) ? 1.f : 0.f;
    if (forcing) forcing[idx] = 0.f;
  }
}

static int grid_of(long long work, int block) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = (work + block - 1) / block;
  const long long cap = 16LL * sms;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dn

using namespace dn;

extern "C" {

size_t dn_gen_kl_table_bytes(int nterms, int size) { return (size_t)(nterms > 0 ? nterms : 0) * (size_t)(size > 0 ? size : 0) * sizeof(double); }

int dn_gen_kl_inputs_f32(const double* coeffs, int batch, int nterms, const double* omega, double eta, int nsd, int size,
                         void* tables, size_t table_bytes, float* inputs, float* forcing, void* stream) {
  if (int rc = dn_device_check()) return rc;
  if (!coeffs || !omega || !inputs || !tables) return fail(DN_EINVAL, "dn_gen_kl_inputs: NULL argument");
  if (nterms < 1 || nterms > kKlMaxTerms) return fail(DN_EINVAL, "dn_gen_kl_inputs: nterms=%d (1..%d)", nterms, kKlMaxTerms);
  if (batch < 1 || size < 4 || size % 4 || (nsd != 2 && nsd != 3))
    return fail(DN_EINVAL, "dn_gen_kl_inputs: batch >= 1, size %% 4 == 0, nsd 2 or 3");
  if (table_bytes < dn_gen_kl_table_bytes(nterms, size)) return fail(DN_EWORKSPACE, "dn_gen_kl_inputs: table buffer too small");
  if (((uintptr_t)inputs % 16) || (forcing && ((uintptr_t)forcing % 16))) return fail(DN_EINVAL, "dn_gen_kl_inputs: outputs must be 16-byte aligned");
  KlParams kp;
  kp.nterms = nterms;
  kp.eta = eta;
  for (int i = 0; i < nterms; ++i) {
    kp.omega[i] = omega[i];
    // lambda = 2 eta sigma / (1 + (eta omega)^2), sigma = 1 (gen_input_calc.py:83-84); np.sqrt(lambda) per direction
    kp.sqrt_lambda[i] = sqrt(2.0 * eta / (1.0 + (eta * omega[i]) * (eta * omega[i])));
  }
  cudaStream_t s = (cudaStream_t)stream;
  double* phi = (double*)tables;
  k_kl_tables<<<(size + 127) / 128, 128, 0, s>>>(kp, size, phi);
  if (nsd == 2) {
    const long long work = (long long)batch * size * (size / 4);
    k_kl_inputs_2d<<<grid_of(work, 256), 256, 0, s>>>(coeffs, kp, batch, size, phi, inputs, forcing);
  } else {
    // 3-D: the reference ships only the diffusivity generator (no dataset class wraps it): `inputs` = (B, 1, N, N, N)
    const long long work = (long long)batch * size * size * (size / 4);
    k_kl_field_3d<<<grid_of(work, 256), 256, 0, s>>>(coeffs, kp, batch, size, phi, inputs, (long long)size * size * size);
    if (forcing) {
      cudaError_t e = cudaMemsetAsync(forcing, 0, (size_t)batch * size * size * size * sizeof(float), s);
      if (e != cudaSuccess) return check_cuda(e, "dn_gen_kl_inputs: memset");
    }
  }
  return check_cuda(cudaGetLastError(), "dn_gen_kl_inputs launch");
}

int dn_gen_image_inputs_f32(const unsigned char* img, int batch, int height, int width, float* inputs, float* forcing,
                            void* stream) {
  if (int rc = dn_device_check()) return rc;
  if (!img || !inputs || batch < 1 || height < 2 || width < 2) return fail(DN_EINVAL, "dn_gen_image_inputs: bad arguments");
  const long long work = (long long)batch * height * width;
  k_image_inputs<<<grid_of(work, 256), 256, 0, (cudaStream_t)stream>>>(img, batch, height, width, inputs, forcing);
  return check_cuda(cudaGetLastError(), "dn_gen_image_inputs launch");
}

int dn_gen_voxel_inputs_f32(const unsigned char* raw, int d0, int d1, int d2, int domain_size, int offset, float* inputs,
                            float* forcing, void* stream) {
  if (int rc = dn_device_check()) return rc;
  if (!raw || !inputs || d0 < 1 || d1 < 1 || d2 < 1 || domain_size < 2 || offset < 0)
    return fail(DN_EINVAL, "dn_gen_voxel_inputs: bad arguments");
  if (offset + d0 > domain_size || offset + d1 > domain_size || offset + d2 > domain_size)
    return fail(DN_EINVAL, "dn_gen_voxel_inputs: the voxel block (%d, %d, %d) at offset %d does not fit %d^3", d0, d1, d2, offset,
                domain_size);
  const long long work = (long long)domain_size * domain_size * domain_size;
  k_voxel_inputs<<<grid_of(work, 256), 256, 0, (cudaStream_t)stream>>>(raw, d0, d1, d2, domain_size, offset, inputs, forcing);
  return check_cuda(cudaGetLastError(), "dn_gen_voxel_inputs launch");
}

int dn_gen_star_inputs_f32(const float* params, int batch, int size, float* inputs, float* forcing, void* stream) {
  if (int rc = dn_device_check()) return rc;
  if (!params || !inputs || batch < 1 || size < 2) return fail(DN_EINVAL, "dn_gen_star_inputs: bad arguments");
  static_assert(sizeof(StarParams) == 11 * sizeof(float), "StarParams layout");
  const long long work = (long long)batch * size * size;
  k_star_inputs<<<grid_of(work, 256), 256, 0, (cudaStream_t)stream>>>((const StarParams*)params, batch, size, inputs, forcing);
  return check_cuda(cudaGetLastError(), "dn_gen_star_inputs launch");
}

int dn_gen_box_masks_3d_f32(const int* params, int batch, int size, float* source, float* sink, float* forcing, void* stream) {
  if (int rc = dn_device_check()) return rc;
  if (!params || !source || !sink || batch < 1 || size < 2) return fail(DN_EINVAL, "dn_gen_box_masks_3d: bad arguments");
  static_assert(sizeof(BoxParams) == 19 * sizeof(int), "BoxParams layout");
  const long long work = (long long)batch * size * size * size;
  k_box_masks_3d<<<grid_of(work, 256), 256, 0, (cudaStream_t)stream>>>((const BoxParams*)params, batch, size, source, sink, forcing);
  return check_cuda(cudaGetLastError(), "dn_gen_box_masks_3d launch");
}

}  // extern "C"
