// ffma2_lab.cu - does FFMA2 (packed fp32x2) relieve ISSUE pressure on B200? Each variant performs the same
// fp32 FMAs per thread, mixed with independent integer (ALU pipe) work so that the scalar form is issue-bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_lab ffma2_lab.cu && ./ffma2_lab
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: scalar FFMA x8 + 8 LOP/IADD, 1: FFMA2 x4 + 8 LOP/IADD, 2: scalar FFMA x8 only, 3: FFMA2 x4 only
__global__ void k(float* out, int n, float a, float b) {
  float2 acc[4];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = threadIdx.x * 2654435761u + i;
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (MODE == 1 || MODE == 3) acc[i] = __ffma2_rn(acc[i], A, B);
      else { acc[i].x = fmaf(acc[i].x, a, b); acc[i].y = fmaf(acc[i].y, a, b); }
    }
    if (MODE < 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = (u[i] ^ (u[i] >> 7)) + 0x9e3779b9u;
    }
  }
  float s = 0.f;
  unsigned x = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += acc[i].x + acc[i].y;
#pragma unroll
  for (int i = 0; i < 8; ++i) x ^= u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(x & 0x7fffff);
}

template <int MODE>
float run(float* out, int n) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 512>>>(out, 16, 0.999f, 0.001f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 512>>>(out, n, 0.999f, 0.001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 512 * sizeof(float));
  const int n = 1 << 16;
  const char* names[4] = {"8 FFMA + 24 int ops", "4 FFMA2 + 24 int ops", "8 FFMA", "4 FFMA2"};
  float ms[4] = {run<0>(out, n), run<1>(out, n), run<2>(out, n), run<3>(out, n)};
  for (int i = 0; i < 4; ++i)
    printf("{\"variant\": \"%s\", \"ms\": %.3f, \"ns_per_iter_per_warp_slot\": %.3f}\n", names[i], ms[i], ms[i] * 1e6 / n);
  return cudaGetLastError() != cudaSuccess;
}
