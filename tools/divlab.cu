// divlab.cu - are the unconditional fast-path sequences of vnl.cu (Fast::div, Fast::sqrt) bit-identical to div.rn.f32 /
// sqrt.rn.f32 inside the operand ranges the kernel checks?  nvcc -arch=sm_100a -o tools/divlab tools/divlab.cu && tools/divlab
// Prints one JSON line: operands tried and mismatches per class.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float fast_div(float a, float b) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(b));
  const float e = fmaf(-b, y, 1.0f);
  y = fmaf(e, y, y);
  const float q = a * y;
  const float r = fmaf(-b, q, a);
  return fmaf(r, y, q);
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  const float g = x * y, h = 0.5f * y;
  const float r = fmaf(-g, g, x);
  return fmaf(r, h, g);
}
__device__ __forceinline__ uint32_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return static_cast<uint32_t>(x);
}
// a float with a uniformly random mantissa and sign and an exponent in [elo, ehi] (binary exponents)
__device__ __forceinline__ float rnd_float(uint64_t key, int elo, int ehi) {
  const uint32_t m = mix(key), e = mix(key ^ 0x9e3779b97f4a7c15ULL);
  const int ex = elo + static_cast<int>(e % static_cast<uint32_t>(ehi - elo + 1));
  const uint32_t bits = (m & 0x807fffffu) | (static_cast<uint32_t>(ex + 127) << 23);
  return __uint_as_float(bits);
}
template <int LO_EXP>
__global__ void lab(unsigned long long n, unsigned long long* out) {
  unsigned long long bad_div = 0, bad_div_sub = 0, bad_sqrt = 0, tried = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    // divisor in [2^-93, 2^83] (the kernel's 1e-28 .. 1e25), numerator up to 2^20 times larger or 2^60 times smaller
    const float b = rnd_float(3 * i, -93, 83);
    int eb;
    frexpf(b, &eb);
    int lo = eb - 60, hi = eb + 20;
    if (lo < LO_EXP) lo = LO_EXP;   // -120: numerators whose residual a - b q underflows; -93 (1e-28): the kernel's range
    if (hi > 100) hi = 100;
    const float a = rnd_float(3 * i + 1, lo, hi);
    const float q0 = __fdiv_rn(a, b), q1 = fast_div(a, b);
    if (__float_as_uint(q0) != __float_as_uint(q1)) {
      if (fabsf(q0) < 1.17549435e-38f) ++bad_div_sub; else ++bad_div;
    }
    const float x = fabsf(rnd_float(3 * i + 2, -99, 93));   // 1e-30 .. 1e28
    if (__float_as_uint(sqrtf(x)) != __float_as_uint(fast_sqrt(x))) ++bad_sqrt;
    ++tried;
  }
  atomicAdd(out + 0, tried); atomicAdd(out + 1, bad_div); atomicAdd(out + 2, bad_div_sub); atomicAdd(out + 3, bad_sqrt);
}
int main() {
  unsigned long long* d; unsigned long long h[4] = {0, 0, 0, 0};
  cudaMalloc(&d, sizeof(h)); cudaMemset(d, 0, sizeof(h));
  int rc = 0;
  for (int pass = 0; pass < 2; ++pass) {
    cudaMemset(d, 0, sizeof(h));
    if (pass == 0) lab<-93><<<148 * 8, 256>>>(4000000000ULL, d);
    else lab<-120><<<148 * 8, 256>>>(4000000000ULL, d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"lab\": \"fast-path div / sqrt against div.rn / sqrt.rn\", \"numerator_min_exp\": %d, \"operand_sets\": %llu, "
           "\"div_mismatch_normal_quotient\": %llu, \"div_mismatch_subnormal_quotient\": %llu, \"sqrt_mismatch\": %llu}\n",
           pass == 0 ? -93 : -120, h[0], h[1], h[2], h[3]);
    if (pass == 0 && (h[1] || h[3])) rc = 2;
  }
  return rc;
}
