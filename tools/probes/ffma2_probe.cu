// Is the packed FP32 FMA (fma.rn.f32x2, sm_100) one issue slot for two FMAs?  Times 16 independent accumulator
// chains per thread: scalar FFMA vs FFMA2, same number of FMAs.  nvcc -arch=sm_100a -O3 ffma2_probe.cu && ./a.out
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 A = make_float2(a, a * 0.999f), B = make_float2(b, b * 1.001f);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) {
        acc[i].x = fmaf(acc[i].x, A.x, B.x);
        acc[i].y = fmaf(acc[i].y, A.y, B.y);
      } else {
        acc[i] = __ffma2_rn(acc[i], A, B);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* d;
  const int blocks = 148 * 8, threads = 256, iters = 20000;
  cudaMalloc(&d, blocks * threads * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
      else k<1><<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)blocks * threads * iters * 16;
      if (rep == 2) printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
