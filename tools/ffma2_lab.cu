// ffma2_lab.cu - does FFMA2 (packed fp32x2) relieve ISSUE pressure on B200 in a metric-like instruction mix?
// Per "pixel pair": 28 fp32 add/mul/fma lane-ops (14 per pixel), 18 ALU-pipe ops (FSETP/FSEL/FMNMX/SEL), 6 MUFU.
// Variant 0 issues the fp32 work as 28 scalar instructions, variant 1 as 14 packed ones.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_lab ffma2_lab.cu && ./ffma2_lab
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqa(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int PACKED>
__global__ void k(const float2* __restrict__ in, float* out, int n) {
  float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0;
  unsigned cnt = 0;
  float2 p = in[threadIdx.x], t = in[threadIdx.x + 512];
  for (int it = 0; it < n; ++it) {
    // ALU-pipe work (scalar in both variants)
    const bool v0 = t.x > 0.f, v1 = t.y > 0.f;
    float2 pp = make_float2(v0 ? p.x : 1.f, v1 ? p.y : 1.f), tt = make_float2(v0 ? t.x : 1.f, v1 ? t.y : 1.f);
    pp.x = fmaxf(pp.x, 1e-7f); pp.y = fmaxf(pp.y, 1e-7f);
    float2 lo = make_float2(fminf(pp.x, tt.x), fminf(pp.y, tt.y)), nhi = make_float2(fminf(-pp.x, -tt.x), fminf(-pp.y, -tt.y));
    float2 lp = make_float2(lg2a(pp.x), lg2a(pp.y)), lt = make_float2(lg2a(tt.x), lg2a(tt.y));
    float2 rs = make_float2(rsqa(tt.x), rsqa(tt.y));
    float2 ad, e, s1, s2, s3, dl, rc, ar;
    const float2 T1 = make_float2(1.25f, 1.25f), T2 = make_float2(1.5625f, 1.5625f), T3 = make_float2(1.953125f, 1.953125f);
    const float2 E = make_float2(5.96e-8f, 5.96e-8f);
    if (PACKED) {
      ad = __fadd2_rn(nhi, lo); a0 = __fadd2_rn(a0, ad); a1 = __ffma2_rn(ad, ad, a1);
      e = __fmul2_rn(lo, E); s1 = __ffma2_rn(lo, T1, nhi); s2 = __ffma2_rn(lo, T2, nhi); s3 = __ffma2_rn(lo, T3, nhi);
      dl = __fadd2_rn(lp, make_float2(-lt.x, -lt.y)); a2 = __ffma2_rn(dl, dl, a2);
      rc = __fmul2_rn(rs, rs); ar = __fmul2_rn(ad, rc); a3 = __fadd2_rn(a3, ar); a4 = __ffma2_rn(ar, ad, a4); a5 = __ffma2_rn(ad, rs, a5);
    } else {
      ad.x = nhi.x + lo.x; ad.y = nhi.y + lo.y; a0.x += ad.x; a0.y += ad.y; a1.x = fmaf(ad.x, ad.x, a1.x); a1.y = fmaf(ad.y, ad.y, a1.y);
      e.x = lo.x * E.x; e.y = lo.y * E.y;
      s1.x = fmaf(lo.x, 1.25f, nhi.x); s1.y = fmaf(lo.y, 1.25f, nhi.y); s2.x = fmaf(lo.x, 1.5625f, nhi.x); s2.y = fmaf(lo.y, 1.5625f, nhi.y);
      s3.x = fmaf(lo.x, 1.953125f, nhi.x); s3.y = fmaf(lo.y, 1.953125f, nhi.y);
      dl.x = lp.x - lt.x; dl.y = lp.y - lt.y; a2.x = fmaf(dl.x, dl.x, a2.x); a2.y = fmaf(dl.y, dl.y, a2.y);
      rc.x = rs.x * rs.x; rc.y = rs.y * rs.y; ar.x = ad.x * rc.x; ar.y = ad.y * rc.y; a3.x += ar.x; a3.y += ar.y;
      a4.x = fmaf(ar.x, ad.x, a4.x); a4.y = fmaf(ar.y, ad.y, a4.y); a5.x = fmaf(ad.x, rs.x, a5.x); a5.y = fmaf(ad.y, rs.y, a5.y);
    }
    cnt += (s1.x > e.x ? 1u : (s2.x > e.x ? 0x100u : (s3.x > e.x ? 0x10000u : 0x1000000u)));
    cnt += (s1.y > e.y ? 1u : (s2.y > e.y ? 0x100u : (s3.y > e.y ? 0x10000u : 0x1000000u)));
    // next "pixel": keep the inputs changing without memory traffic
    p.x = p.x * 1.0001f + 1e-3f; p.y = p.y * 0.9999f + 2e-3f; t.x += 1e-3f; t.y += 2e-3f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + a4.x + a4.y + a5.x + a5.y + cnt;
}

template <int PACKED>
float run(const float2* in, float* out, int n) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PACKED><<<148 * 2, 512>>>(in, out, 16);
  cudaEventRecord(e0);
  k<PACKED><<<148 * 2, 512>>>(in, out, n);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float2* in; float* out;
  cudaMalloc(&in, 1024 * sizeof(float2)); cudaMalloc(&out, 148 * 2 * 512 * sizeof(float));
  float2 h[1024]; for (int i = 0; i < 1024; ++i) h[i] = make_float2(0.5f + i * 0.01f, 0.7f + i * 0.013f);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int n = 1 << 15;
  const float a = run<0>(in, out, n), b = run<1>(in, out, n);
  // 8 warps per SMSP (2 CTAs x 512 threads), 1.965 GHz
  printf("{\"variant\": \"scalar fp32\", \"ms\": %.3f, \"cycles_per_warp_pair\": %.1f}\n", a, a * 1e-3 * 1.965e9 / n / 8);
  printf("{\"variant\": \"packed f32x2\", \"ms\": %.3f, \"cycles_per_warp_pair\": %.1f}\n", b, b * 1e-3 * 1.965e9 / n / 8);
  return cudaGetLastError() != cudaSuccess;
}
