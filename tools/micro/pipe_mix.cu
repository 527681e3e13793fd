// Do ALU-pipe (LOP3/IADD3/FSEL), LSU (LDS) and FMA-pipe (scalar and packed FP32) instructions overlap on one
// sm_100a sub-partition, or do they serialise at dispatch?  clock64() inside the kernel, one CTA of 1024
// threads per SM (8 warps per sub-partition), register operands only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix pipe_mix.cu && ./pipe_mix
#include <cstdio>
#include <cuda_runtime.h>
#define NA 8
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, const float* in, int iters, long long* cyc) {
  __shared__ float sh[2048];
  float2 acc[NA], X[4];
  unsigned ia[NA], ib[4];
  float sa[NA];
  for (int i = threadIdx.x; i < 2048; i += 1024) sh[i] = in[i & 1023];
#pragma unroll
  for (int i = 0; i < NA; ++i) { acc[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 2 * i + 1]); ia[i] = __float_as_uint(in[threadIdx.x + 5 * i]); sa[i] = in[threadIdx.x + 3 * i + 2]; }
#pragma unroll
  for (int i = 0; i < 4; ++i) { X[i] = make_float2(in[threadIdx.x + 64 + i], in[threadIdx.x + 80 + i]); ib[i] = __float_as_uint(in[threadIdx.x + 100 + i]); }
  __syncthreads();
  const float* sp = sh + (threadIdx.x & 1023);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (MODE == 0) acc[i] = __fadd2_rn(acc[i], X[i & 3]);
      if (MODE == 1) ia[i] = (ia[i] ^ ib[i & 3]) + ib[(i + 1) & 3];                                   // LOP3 + IADD3 (2 ALU)
      if (MODE == 2) { acc[i] = __fadd2_rn(acc[i], X[i & 3]); ia[i] = (ia[i] ^ ib[i & 3]) + ib[(i + 1) & 3]; }
      if (MODE == 3) { sa[i] += X[i & 3].x; ia[i] = (ia[i] ^ ib[i & 3]) + ib[(i + 1) & 3]; }
      if (MODE == 4) { acc[i] = __ffma2_rn(acc[i], X[i & 3], X[(i + 1) & 3]); ia[i] = (ia[i] ^ ib[i & 3]) + ib[(i + 1) & 3]; }
      if (MODE == 5) { sa[i] += sp[(i * 32 + it) & 1023]; }                                            // LDS + FADD
      if (MODE == 6) { acc[i] = __fadd2_rn(acc[i], X[i & 3]); sa[i] += sp[(i * 32 + it) & 1023]; }     // FADD2 + LDS + FADD
      if (MODE == 7) { acc[i] = __fadd2_rn(acc[i], X[i & 3]); ia[i] = ia[i] ^ ib[i & 3]; }             // FADD2 + 1 LOP3
      if (MODE == 8) { acc[i] = __fadd2_rn(acc[i], X[i & 3]); sa[i] = (ia[i] & 1u) ? X[i & 3].y : sa[i]; ia[i] += ib[i & 3]; }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NA; ++i) s += acc[i].x + acc[i].y + sa[i] + __uint_as_float(ia[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, float* in, long long* cyc, int sms) {
  const int iters = 2048;
  k<MODE><<<sms, 1024>>>(out, in, 16, cyc);
  k<MODE><<<sms, 1024>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  long long* h = new long long[sms];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
  delete[] h;
  printf("%-40s %9.0f cyc  %.2f cycles per unrolled slot per warp (8 warps/SMSP)\n", name, mx, mx / (8.0 * iters * NA));
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float *out, *in; long long* cyc;
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 1024 * 4);
  cudaMalloc(&in, 8192); cudaMemset(in, 0, 8192);
  cudaMalloc(&cyc, sizeof(long long) * p.multiProcessorCount);
  const int s = p.multiProcessorCount;
  run<0>("FADD2", out, in, cyc, s);
  run<1>("LOP3+IADD3", out, in, cyc, s);
  run<2>("FADD2 + LOP3+IADD3", out, in, cyc, s);
  run<3>("FADD + LOP3+IADD3", out, in, cyc, s);
  run<4>("FFMA2(3reg) + LOP3+IADD3", out, in, cyc, s);
  run<5>("LDS + FADD", out, in, cyc, s);
  run<6>("FADD2 + LDS + FADD", out, in, cyc, s);
  run<7>("FADD2 + LOP3", out, in, cyc, s);
  run<8>("FADD2 + LOP3.. + FSEL + IADD3", out, in, cyc, s);
  return 0;
}
