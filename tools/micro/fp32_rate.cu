// FP32 pipe rates on sm_100a: scalar FFMA/FADD vs packed FFMA2/FADD2 (lane-ops per clock per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rate fp32_rate.cu && ./fp32_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  const float2 A = make_float2(a, a), Bv = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a, b); acc[i].y = fmaf(acc[i].y, a, b); }
      if (MODE == 1) acc[i] = __ffma2_rn(acc[i], A, Bv);
      if (MODE == 2) { acc[i].x = acc[i].x + b; acc[i].y = acc[i].y + b; }
      if (MODE == 3) acc[i] = __fadd2_rn(acc[i], Bv);
      if (MODE == 4) { acc[i].x = acc[i].x * a; acc[i].y = acc[i].y * a; }
      if (MODE == 5) acc[i] = __fmul2_rn(acc[i], A);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, int sms, float clk_ghz) {
  const int iters = 4096, grid = sms * 8;
  k<MODE><<<grid, 256>>>(out, 16, 1.0001f, 1e-4f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 1e-4f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double laneops = (double)grid * 256 * iters * 32;
  printf("%-8s %.3f ms  %.1f lane-ops/clk/SM (at %.3f GHz)  %.2f T lane-ops/s\n", name, ms,
         laneops / (ms * 1e-3) / (clk_ghz * 1e9) / sms, clk_ghz, laneops / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float ghz = khz * 1e-6f;
  float* out;
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
  run<0>("FFMA", out, p.multiProcessorCount, ghz);
  run<1>("FFMA2", out, p.multiProcessorCount, ghz);
  run<2>("FADD", out, p.multiProcessorCount, ghz);
  run<3>("FADD2", out, p.multiProcessorCount, ghz);
  run<4>("FMUL", out, p.multiProcessorCount, ghz);
  run<5>("FMUL2", out, p.multiProcessorCount, ghz);
  return 0;
}
