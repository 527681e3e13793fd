// FP32 issue/pipe rates on sm_100a with REGISTER operands (the patterns real element code has), and mixes:
// scalar vs packed (FFMA2/FADD2/FMUL2), packed + scalar interleaved, packed + ALU (FSEL) interleaved.
// Reports warp-instructions per clock per SM sub-partition, from clock64() inside the kernel
// (frequency-independent).  One CTA of 1024 threads per SM = 8 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_mix fp32_mix.cu && ./fp32_mix
#include <cstdio>
#include <cuda_runtime.h>

#define NA 8

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, const float* in, int iters, long long* cyc) {
  float2 acc[NA], X[4], Y[4];
  float sa[2 * NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 2 * i + 1]);
#pragma unroll
  for (int i = 0; i < 2 * NA; ++i) sa[i] = in[threadIdx.x + 3 * i + 2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    X[i] = make_float2(in[threadIdx.x + 64 + i], in[threadIdx.x + 80 + i]);
    Y[i] = make_float2(in[threadIdx.x + 96 + i], in[threadIdx.x + 112 + i]);
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (MODE == 0) { sa[2 * i] = fmaf(sa[2 * i], X[i & 3].x, Y[i & 3].x); sa[2 * i + 1] = fmaf(sa[2 * i + 1], X[i & 3].y, Y[i & 3].y); }
      if (MODE == 1) acc[i] = __ffma2_rn(acc[i], X[i & 3], Y[i & 3]);
      if (MODE == 2) acc[i] = __fadd2_rn(acc[i], X[i & 3]);
      if (MODE == 3) acc[i] = __fmul2_rn(acc[i], X[i & 3]);
      if (MODE == 4) { sa[2 * i] += X[i & 3].x; sa[2 * i + 1] += X[i & 3].y; }
      if (MODE == 5) { acc[i] = __fadd2_rn(acc[i], X[i & 3]); sa[2 * i] += Y[i & 3].x; sa[2 * i + 1] += Y[i & 3].y; }
      if (MODE == 6) { acc[i] = __ffma2_rn(acc[i], X[i & 3], Y[i & 3]); sa[2 * i] = fmaf(sa[2 * i], X[i & 3].x, Y[i & 3].y); sa[2 * i + 1] = fmaf(sa[2 * i + 1], X[i & 3].y, Y[i & 3].x); }
      if (MODE == 7) { acc[i] = __fadd2_rn(acc[i], X[i & 3]); sa[2 * i] = (sa[2 * i + 1] > 0.5f) ? Y[i & 3].x : sa[2 * i]; }
      if (MODE == 8) { acc[i] = __ffma2_rn(acc[i], X[i & 3], Y[i & 3]); sa[2 * i] += Y[i & 3].x; }
      if (MODE == 9) { acc[i] = __fadd2_rn(acc[i], X[i & 3]); sa[2 * i] += Y[i & 3].x; }
      if (MODE == 10) { acc[i] = __ffma2_rn(acc[i], X[i & 3], acc[i]); }                     // 2 distinct register operands
      if (MODE == 11) { sa[2 * i] = (sa[2 * i + 1] > 0.5f) ? Y[i & 3].x : sa[2 * i]; sa[2 * i + 1] = (sa[2 * i] > 0.25f) ? Y[i & 3].y : sa[2 * i + 1]; }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NA; ++i) s += acc[i].x + acc[i].y + sa[2 * i] + sa[2 * i + 1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter_packed, int per_iter_scalar, float* out, float* in, long long* cyc, int sms) {
  const int iters = 2048, grid = sms;
  k<MODE><<<grid, 1024>>>(out, in, 16, cyc);
  k<MODE><<<grid, 1024>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  long long* h = new long long[grid];
  cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  delete[] h;
  // 8 warps per sub-partition, each issues iters * (packed + scalar) instructions of interest
  const double winst = 8.0 * iters * (per_iter_packed + per_iter_scalar);
  const double lane = 8.0 * iters * (per_iter_packed * 64.0 + per_iter_scalar * 32.0);
  printf("%-28s %9.0f cyc  %.3f inst/clk/SMSP  %.1f lane-ops/clk/SM  (%.2f cyc per packed-equivalent)\n", name, mx,
         winst / mx, 4.0 * lane / mx, mx / (8.0 * iters * (per_iter_packed + 0.5 * per_iter_scalar)));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float *out, *in;
  long long* cyc;
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 4 * 256 * 4);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0, 4096);
  cudaMalloc(&cyc, sizeof(long long) * p.multiProcessorCount * 4);
  const int s = p.multiProcessorCount;
  run<0>("FFMA 3reg x16", 0, 16, out, in, cyc, s);
  run<1>("FFMA2 3reg x8", 8, 0, out, in, cyc, s);
  run<2>("FADD2 2reg x8", 8, 0, out, in, cyc, s);
  run<3>("FMUL2 2reg x8", 8, 0, out, in, cyc, s);
  run<4>("FADD 2reg x16", 0, 16, out, in, cyc, s);
  run<5>("FADD2 x8 + FADD x16", 8, 16, out, in, cyc, s);
  run<6>("FFMA2 x8 + FFMA x16", 8, 16, out, in, cyc, s);
  run<7>("FADD2 x8 + FSETP/FSEL x8", 8, 16, out, in, cyc, s);
  run<8>("FFMA2 x8 + FADD x8", 8, 8, out, in, cyc, s);
  run<9>("FADD2 x8 + FADD x8", 8, 8, out, in, cyc, s);
  run<10>("FFMA2 2reg x8", 8, 0, out, in, cyc, s);
  run<11>("FSETP+FSEL x16", 0, 32, out, in, cyc, s);
  return 0;
}
