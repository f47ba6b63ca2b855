// pipe_lab.cu - measured issue cost of the instruction classes the metric / loss kernels are made of, on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipe_lab tools/pipe_lab.cu && ./gpurun_out/pipe_lab
// Every variant runs 148 x 2 CTAs x 512 threads (8 warps per SM sub-partition, as the product kernels do) through
// ITERS iterations of 16 INDEPENDENT instances of one instruction (no dependency stalls: the number is the pipe /
// issue cost, not the latency). Reported: SM cycles per warp-instruction per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int U = 16;

enum Op { FADD, FFMA_RRR, FFMA_IMM, FMUL_SAT, FSEL, FSET, FSETP_SEL, FMNMX, FMNMX3, LG2, RSQ, RCP, IADD, LOP3, FADD2, FFMA2, FMUL2,
          MIX_LEAN, MIX_LEAN2, NOPS };
const char* kNames[] = {"FADD", "FFMA rrr", "FFMA imm", "FMUL.SAT", "FSEL (selp)", "FSET (set.gt.f32)", "FSETP+FSEL", "FMNMX", "FMNMX3",
                        "MUFU.LG2", "MUFU.RSQ", "MUFU.RCP", "IADD", "LOP3", "FADD2 (f32x2)", "FFMA2 (f32x2)", "FMUL2 (f32x2)",
                        "lean metric px (scalar)", "lean metric px (f32x2 sums)"};

template <int OP>
__global__ void __launch_bounds__(512, 2) k(float* out, float a0, float b0, unsigned long long* cyc) {
  float x[U];
  unsigned long long xx[U / 2];
#pragma unroll
  for (int i = 0; i < U; ++i) x[i] = a0 + i * 0.001f + threadIdx.x * 1e-6f;
#pragma unroll
  for (int i = 0; i < U / 2; ++i) xx[i] = (static_cast<unsigned long long>(__float_as_uint(x[2 * i])) << 32) | __float_as_uint(x[2 * i + 1]);
  float y = b0, z = b0 * 0.5f;
  unsigned long long yy = (static_cast<unsigned long long>(__float_as_uint(y)) << 32) | __float_as_uint(z);
  int pi = threadIdx.x & 1;
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (OP == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y));
      if (OP == FFMA_RRR) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y), "f"(z));
      if (OP == FFMA_IMM) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, %1;" : "+f"(x[i]) : "f"(y));
      if (OP == FMUL_SAT) asm volatile("mul.sat.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y));
      if (OP == FSEL) asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %0, %1, p;}" : "+f"(x[i]) : "f"(y), "r"(pi));
      if (OP == FSET) asm volatile("set.gt.f32.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y));
      if (OP == FSETP_SEL) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %2, p;}" : "+f"(x[i]) : "f"(y), "f"(z));
      if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y));
      if (OP == FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y), "f"(z));
      if (OP == LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(*reinterpret_cast<int*>(&x[i])) : "r"(pi));
      if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*reinterpret_cast<int*>(&x[i])) : "r"(pi), "r"(it));
    }
#pragma unroll
    for (int i = 0; i < U / 2; ++i) {
      if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(xx[i]) : "l"(yy));
      if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(xx[i]) : "l"(yy));
      if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(xx[i]) : "l"(yy));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < U; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < U / 2; ++i) s += __uint_as_float(static_cast<unsigned>(xx[i])) + __uint_as_float(static_cast<unsigned>(xx[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) atomicMax(cyc, static_cast<unsigned long long>(t1 - t0));
}

// the lean per-pixel body of metric_math.cuh (log + rsq groups, SILog share), 4 pixels per iteration
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqa(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float maxnan(float x, float y) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(y)); return r; }

template <int PACK>
__global__ void __launch_bounds__(512, 2) kmix(float* out, const float4* __restrict__ in, unsigned long long* cyc, int iters) {
  float s_abs = 0.f, s_sq = 0.f, s_l10 = 0.f, s_ln = 0.f, s_rsq = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, nv = 0.f, s0 = 0.f;
  float2 S_abs = {0.f, 0.f}, S_sq = {0.f, 0.f}, S_l10 = {0.f, 0.f}, S_ln = {0.f, 0.f}, S_rsq = {0.f, 0.f}, C1 = {0.f, 0.f}, C2 = {0.f, 0.f},
         C3 = {0.f, 0.f}, NV = {0.f, 0.f}, S0 = {0.f, 0.f};
  float4 p4 = in[threadIdx.x], t4 = in[512 + threadIdx.x];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, tv[4] = {t4.x, t4.y, t4.z, t4.w};
    if (PACK == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = pv[j], t = tv[j];
        const bool v = t > 0.f;
        const float tt = v ? t : 1.0f, pp = v ? p : 1.0f;
        const float hi = maxnan(pp, tt), lo = fminf(pp, tt);
        const float ad = hi - lo;
        s_abs += ad; s_sq = fmaf(ad, ad, s_sq);
        const float e = lo * 5.9604644775390625e-08f;
        c1 += (fmaf(lo, 1.25f, -hi) > e) ? 1.f : 0.f;
        c2 += (fmaf(lo, 1.5625f, -hi) > e) ? 1.f : 0.f;
        c3 += (fmaf(lo, 1.953125f, -hi) > e) ? 1.f : 0.f;
        nv += __saturatef(t * 8.507059173023462e37f);
        const float dl = lg2a(pp) - lg2a(tt);
        s_l10 += fabsf(dl); s_ln = fmaf(dl, dl, s_ln); s0 += dl;
        s_rsq = fmaf(ad, rsqa(tt), s_rsq);
      }
    } else {
      // pairs (0,1) and (2,3): selects / min / max / compares / MUFU scalar, every sum and product packed
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float pa = pv[2 * h], pb = pv[2 * h + 1], ta = tv[2 * h], tb = tv[2 * h + 1];
        const bool va = ta > 0.f, vb = tb > 0.f;
        const float tta = va ? ta : 1.0f, ppa = va ? pa : 1.0f, ttb = vb ? tb : 1.0f, ppb = vb ? pb : 1.0f;
        const float2 hi = make_float2(maxnan(ppa, tta), maxnan(ppb, ttb)), lo = make_float2(fminf(ppa, tta), fminf(ppb, ttb));
        const float2 nhi = make_float2(-hi.x, -hi.y);
        const float2 ad = __fadd2_rn(hi, make_float2(-lo.x, -lo.y));
        S_abs = __fadd2_rn(S_abs, ad); S_sq = __ffma2_rn(ad, ad, S_sq);
        const float2 e = __fmul2_rn(lo, make_float2(5.9604644775390625e-08f, 5.9604644775390625e-08f));
        const float2 x1 = __ffma2_rn(lo, make_float2(1.25f, 1.25f), nhi), x2 = __ffma2_rn(lo, make_float2(1.5625f, 1.5625f), nhi),
                     x3 = __ffma2_rn(lo, make_float2(1.953125f, 1.953125f), nhi);
        C1 = __fadd2_rn(C1, make_float2(x1.x > e.x ? 1.f : 0.f, x1.y > e.y ? 1.f : 0.f));
        C2 = __fadd2_rn(C2, make_float2(x2.x > e.x ? 1.f : 0.f, x2.y > e.y ? 1.f : 0.f));
        C3 = __fadd2_rn(C3, make_float2(x3.x > e.x ? 1.f : 0.f, x3.y > e.y ? 1.f : 0.f));
        NV = __fadd2_rn(NV, make_float2(__saturatef(ta * 8.507059173023462e37f), __saturatef(tb * 8.507059173023462e37f)));
        const float2 dl = __fadd2_rn(make_float2(lg2a(ppa), lg2a(ppb)), make_float2(-lg2a(tta), -lg2a(ttb)));
        S_l10 = __fadd2_rn(S_l10, make_float2(fabsf(dl.x), fabsf(dl.y))); S_ln = __ffma2_rn(dl, dl, S_ln); S0 = __fadd2_rn(S0, dl);
        S_rsq = __ffma2_rn(ad, make_float2(rsqa(tta), rsqa(ttb)), S_rsq);
      }
    }
    // next "pixels": keep the inputs changing without memory traffic (4 FFMA per quad, both variants)
    p4.x = fmaf(p4.x, 1.0001f, 1e-3f); p4.y = fmaf(p4.y, 0.9999f, 2e-3f); t4.x += 1e-3f; t4.z += 2e-3f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s_abs + s_sq + s_l10 + s_ln + s_rsq + c1 + c2 + c3 + nv + s0 + S_abs.x + S_abs.y + S_sq.x + S_sq.y +
      S_l10.x + S_l10.y + S_ln.x + S_ln.y + S_rsq.x + S_rsq.y + C1.x + C1.y + C2.x + C2.y + C3.x + C3.y + NV.x + NV.y + S0.x + S0.y;
  if (threadIdx.x == 0) atomicMax(cyc, static_cast<unsigned long long>(t1 - t0));
}

template <int OP>
void run(float* out, unsigned long long* cyc) {
  cudaMemset(cyc, 0, 8);
  k<OP><<<148 * 2, 512>>>(out, 1.0f, 1.0000001f, cyc);
  cudaMemset(cyc, 0, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<148 * 2, 512>>>(out, 1.0f, 1.0000001f, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  if (OP == FADD) printf("{\"clock_check\": \"FADD kernel: %llu clock64 cycles in %.2f us (event) -> %.0f MHz\"}\n", c, ms * 1e3, c / (ms * 1e3));
  const int per_iter = (OP >= FADD2 && OP <= FMUL2) ? U / 2 : U;
  // 8 warps per sub-partition each issue ITERS * per_iter instructions
  printf("{\"op\": \"%s\", \"cycles_per_warp_instr_per_smsp\": %.2f}\n", kNames[OP], static_cast<double>(c) / (8.0 * ITERS * per_iter));
}

int main() {
  float* out; unsigned long long* cyc; float4* in;
  cudaMalloc(&out, 148 * 2 * 512 * sizeof(float)); cudaMalloc(&cyc, 8); cudaMalloc(&in, 1024 * sizeof(float4));
  float4 h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = make_float4(0.5f + i * 0.01f, 0.7f + i * 0.013f, 1.5f + i * 0.002f, (i % 5) ? 2.5f : 0.f);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<FADD>(out, cyc); run<FFMA_RRR>(out, cyc); run<FFMA_IMM>(out, cyc); run<FMUL_SAT>(out, cyc); run<FSEL>(out, cyc); run<FSET>(out, cyc);
  run<FSETP_SEL>(out, cyc); run<FMNMX>(out, cyc); run<FMNMX3>(out, cyc); run<LG2>(out, cyc); run<RSQ>(out, cyc); run<RCP>(out, cyc);
  run<IADD>(out, cyc); run<LOP3>(out, cyc); run<FADD2>(out, cyc); run<FFMA2>(out, cyc); run<FMUL2>(out, cyc);
  const int iters = 4096;
  for (int pack = 0; pack < 2; ++pack) {
    cudaMemset(cyc, 0, 8);
    if (pack) kmix<1><<<148 * 2, 512>>>(out, in, cyc, iters); else kmix<0><<<148 * 2, 512>>>(out, in, cyc, iters);
    unsigned long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"op\": \"%s\", \"cycles_per_warp_pixel_per_smsp\": %.2f}\n", kNames[MIX_LEAN + pack], static_cast<double>(c) / (8.0 * iters * 4));
  }
  return cudaGetLastError() != cudaSuccess;
}
