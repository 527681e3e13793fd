// Does the operand-reuse cache (.reuse) lower the dispatch cost of FP32 instructions on sm_100a?
// Same harness as fp32_mix.cu: one CTA of 1024 threads per SM (8 warps per sub-partition), clock64().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o reuse reuse.cu && ./reuse
#include <cstdio>
#include <cuda_runtime.h>

#define NA 8
__constant__ float2 cK[4];

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
      if (MODE == 0) acc[i] = __ffma2_rn(acc[i], X[i & 3], Y[i & 3]);          // 3 distinct registers, no reuse
      if (MODE == 1) acc[i] = __ffma2_rn(acc[i], X[0], Y[0]);                  // two operands repeat -> .reuse x2
      if (MODE == 2) acc[i] = __ffma2_rn(X[0], Y[i & 3], acc[i]);              // one operand repeats
      if (MODE == 3) acc[i] = __fadd2_rn(acc[i], X[i & 3]);                    // FADD2 no reuse
      if (MODE == 4) acc[i] = __fadd2_rn(acc[i], X[0]);                        // FADD2 one operand repeats
      if (MODE == 5) acc[i] = __ffma2_rn(acc[i], cK[i & 3], Y[i & 3]);         // one constant-bank operand
      if (MODE == 6) acc[i] = __ffma2_rn(acc[i], cK[i & 3], cK[(i + 1) & 3]);  // two constant-bank operands?
      if (MODE == 7) { sa[2 * i] = fmaf(sa[2 * i], X[0].x, Y[0].x); sa[2 * i + 1] = fmaf(sa[2 * i + 1], X[0].x, Y[0].x); }   // scalar, 2 repeat
      if (MODE == 8) { sa[2 * i] = fmaf(sa[2 * i], X[i & 3].x, Y[i & 3].x); sa[2 * i + 1] = fmaf(sa[2 * i + 1], X[i & 3].y, Y[i & 3].y); }
      if (MODE == 9) acc[i] = __fmul2_rn(acc[i], cK[i & 3]);                   // FMUL2 by a constant
      if (MODE == 10) acc[i] = __fadd2_rn(acc[i], acc[i]);                     // FADD2, one distinct register
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
void run(const char* name, int per_iter, float* out, float* in, long long* cyc, int sms) {
  const int iters = 2048, grid = sms;
  k<MODE><<<grid, 1024>>>(out, in, 16, cyc);
  k<MODE><<<grid, 1024>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  long long* h = new long long[grid];
  cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  delete[] h;
  printf("%-44s %9.0f cyc  %.2f dispatch cycles per instruction\n", name, mx, mx / (8.0 * iters * per_iter));
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
  run<0>("FFMA2 3 registers, no repeat", 8, out, in, cyc, s);
  run<1>("FFMA2 two operands repeat (reuse)", 8, out, in, cyc, s);
  run<2>("FFMA2 one operand repeats (reuse)", 8, out, in, cyc, s);
  run<3>("FADD2 2 registers, no repeat", 8, out, in, cyc, s);
  run<4>("FADD2 one operand repeats (reuse)", 8, out, in, cyc, s);
  run<5>("FFMA2 one constant-bank operand", 8, out, in, cyc, s);
  run<6>("FFMA2 two constant operands (as compiled)", 8, out, in, cyc, s);
  run<7>("FFMA scalar, two operands repeat", 16, out, in, cyc, s);
  run<8>("FFMA scalar, 3 registers", 16, out, in, cyc, s);
  run<9>("FMUL2 by a constant", 8, out, in, cyc, s);
  run<10>("FADD2 x + x", 8, out, in, cyc, s);
  return 0;
}
