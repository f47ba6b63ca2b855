// mathlab.cu - accuracy of the SFU approximations (lg2/rcp/rsqrt .approx) on this GPU, and of three
// candidate evaluations of |ln p - ln t|, against fp64. Decides which forms the fast metric / loss
// math may use within the 1e-5 tolerance of BASELINE.json.   nvcc -arch=sm_100a -O3 -o mathlab mathlab.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqa(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_nr(float x) { float y = rcpa(x); return fmaf(y, fmaf(-x, y, 1.0f), y); }

struct Stat { double max_abs, sum, sum2, max_rel; unsigned long long n; };
__device__ void acc(Stat* s, double err, double ref) {
  // per-thread local then atomics (coarse: this is a lab tool, not a product kernel)
  atomicAdd(&s->sum, err); atomicAdd(&s->sum2, err * err); atomicAdd(&s->n, 1ull);
  unsigned long long* m = reinterpret_cast<unsigned long long*>(&s->max_abs);
  unsigned long long v = __double_as_longlong(fabs(err));
  atomicMax(m, v);
  if (ref != 0.0) { unsigned long long* mr = reinterpret_cast<unsigned long long*>(&s->max_rel); atomicMax(mr, (unsigned long long)__double_as_longlong(fabs(err / ref))); }
}

// op: 0 lg2 on [lo,hi) (log-uniform), 1 rcp, 2 rsqrt, 3 ex2 on [-hi, 0]
__global__ void unary(int op, float lo, float hi, int n, Stat* s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double u = (i + 0.5) / n;
  float x = (op == 3) ? (float)(-hi * u) : (float)(lo * exp(u * log((double)hi / lo)));
  double err, ref;
  if (op == 0) { ref = log2((double)x); err = (double)lg2a(x) - ref; }
  else if (op == 1) { ref = 1.0 / (double)x; err = ((double)rcpa(x) - ref) / ref; }
  else if (op == 2) { ref = 1.0 / sqrt((double)x); err = ((double)rsqa(x) - ref) / ref; }
  else { ref = exp2((double)x); err = ((double)ex2a(x) - ref) / ref; }
  acc(s, err, op == 0 ? ref : 0.0);
}

// |ln p - ln t| three ways on synthetic depth pairs; also aggregate sums
__global__ void pipelines(int n, unsigned seed, Stat* sA, Stat* sB, Stat* sC, double* sums) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // cheap hash RNG
  unsigned h = (i + 1) * 2654435761u ^ seed; h ^= h >> 16; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
  unsigned h2 = h * 1664525u + 1013904223u; h2 ^= h2 >> 15; h2 *= 2246822519u; h2 ^= h2 >> 13;
  float t = 0.5f + 9.5f * (h >> 8) * (1.0f / 16777216.0f);
  float g = ((h2 >> 8) * (1.0f / 16777216.0f) - 0.5f) * 3.46f * 0.5f;   // uniform with sigma 0.5
  float p = fmaxf(t + g, 1e-3f);
  float hi = fmaxf(p, t), lo = fminf(p, t);
  double ref = log((double)hi / (double)lo);
  float LA = logf(hi / lo);                                   // IEEE divide + libdevice log
  float LB = lg2a(hi * rcp_nr(lo)) * 0.69314718055994531f;    // rcp + Newton, MUFU.LG2
  float LC = lg2a(hi * rcpa(lo)) * 0.69314718055994531f;      // raw rcp, MUFU.LG2
  acc(sA, (double)LA - ref, ref); acc(sB, (double)LB - ref, ref); acc(sC, (double)LC - ref, ref);
  atomicAdd(&sums[0], ref); atomicAdd(&sums[1], (double)LA); atomicAdd(&sums[2], (double)LB); atomicAdd(&sums[3], (double)LC);
  atomicAdd(&sums[4], ref * ref); atomicAdd(&sums[5], (double)LA * LA); atomicAdd(&sums[6], (double)LB * LB); atomicAdd(&sums[7], (double)LC * LC);
  // |p-t|/t with raw rcp vs exact
  double ar = fabs((double)p - t) / t;
  atomicAdd(&sums[8], ar); atomicAdd(&sums[9], (double)(fabsf(p - t) * rcpa(t)));
  atomicAdd(&sums[10], fabs((double)p - t) / sqrt((double)t)); atomicAdd(&sums[11], (double)(fabsf(p - t) * rsqa(t)));
}

static void show(const char* name, Stat* d) {
  Stat h; cudaMemcpy(&h, d, sizeof(Stat), cudaMemcpyDeviceToHost);
  printf("{\"test\": \"%s\", \"n\": %llu, \"max_abs_err\": %.4g, \"mean_err\": %.4g, \"rms_err\": %.4g, \"max_rel_err\": %.4g}\n", name, h.n,
         h.max_abs, h.sum / h.n, sqrt(h.sum2 / h.n), h.max_rel);
  cudaMemset(d, 0, sizeof(Stat));
}

int main() {
  Stat* s; cudaMalloc(&s, 4 * sizeof(Stat)); cudaMemset(s, 0, 4 * sizeof(Stat));
  double* sums; cudaMalloc(&sums, 16 * sizeof(double)); cudaMemset(sums, 0, 16 * sizeof(double));
  const int n = 1 << 22;
  unary<<<n / 256, 256>>>(0, 1.0f, 2.0f, n, s); cudaDeviceSynchronize(); show("lg2.approx on [1,2) (abs err, log2 units)", s);
  unary<<<n / 256, 256>>>(0, 1.0f, 1.01f, n, s); cudaDeviceSynchronize(); show("lg2.approx on [1,1.01)", s);
  unary<<<n / 256, 256>>>(0, 1.0f, 64.0f, n, s); cudaDeviceSynchronize(); show("lg2.approx on [1,64)", s);
  unary<<<n / 256, 256>>>(0, 0.015625f, 1.0f, n, s); cudaDeviceSynchronize(); show("lg2.approx on [1/64,1)", s);
  unary<<<n / 256, 256>>>(1, 0.01f, 100.0f, n, s); cudaDeviceSynchronize(); show("rcp.approx rel err on [0.01,100)", s);
  unary<<<n / 256, 256>>>(2, 0.01f, 100.0f, n, s); cudaDeviceSynchronize(); show("rsqrt.approx rel err on [0.01,100)", s);
  unary<<<n / 256, 256>>>(3, 0.0f, 30.0f, n, s); cudaDeviceSynchronize(); show("ex2.approx rel err on [-30,0]", s);
  pipelines<<<n / 256, 256>>>(n, 12345u, s, s + 1, s + 2, sums); cudaDeviceSynchronize();
  show("|ln p-ln t| A: logf(fdiv_rn)", s); show("|ln p-ln t| B: lg2.approx(hi*rcp_nr(lo))", s + 1); show("|ln p-ln t| C: lg2.approx(hi*rcp.approx(lo))", s + 2);
  double h[16]; cudaMemcpy(h, sums, sizeof(h), cudaMemcpyDeviceToHost);
  printf("{\"test\": \"aggregate sum|L| rel err\", \"A\": %.3g, \"B\": %.3g, \"C\": %.3g}\n", h[1] / h[0] - 1, h[2] / h[0] - 1, h[3] / h[0] - 1);
  printf("{\"test\": \"aggregate sum L^2 rel err\", \"A\": %.3g, \"B\": %.3g, \"C\": %.3g}\n", h[5] / h[4] - 1, h[6] / h[4] - 1, h[7] / h[4] - 1);
  printf("{\"test\": \"aggregate absrel (raw rcp) / rsq (raw rsqrt) rel err\", \"absrel\": %.3g, \"rsq\": %.3g}\n", h[9] / h[8] - 1, h[11] / h[10] - 1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
