// metric_math.cuh - per-pixel arithmetic of the masked error metrics.
//
// Two evaluation modes of the SAME quantities (include/mde_b200.h, MDE_Q_*):
//   Ref  : the reference's own op sequence (metrics.py:75-109 + torchmetrics closed forms):
//          log10f(p)-log10f(t), log1pf(p)-log1pf(t), IEEE divides everywhere.
//   Fast : algebraically equal forms built on the SFU (MUFU.RCP / LG2 / RSQ); used by default
//          because the suite in Ref form is issue-bound, not HBM-bound, on B200.
//
// Bit-exact delta counts in BOTH modes. The reference counts max(fl(p/t), fl(t/p)) < 1.25^k, which
// equals r = fl(max(p,t)/min(p,t)) < 1.25^k (round-to-nearest is monotone). Ref mode evaluates r
// with an IEEE divide. Fast mode evaluates u = log_1.25(q), q = hi * rcp.approx(lo): |u - log_1.25(x)|
// <= 2.5e-6 for the true quotient x (rcp.approx: 1 ulp, lg2.approx: <= 3.3e-7 absolute, both bounds
// measured on B200 by tools/mathlab.cu and documented in the PTX ISA), so u < k decides r < 1.25^k
// whenever u is farther than 1e-5 from the integers 1, 2, 3; inside that window (6e-5 of the pixels)
// the pixel takes the exact IEEE divide. The integer counts are therefore identical to the reference.
#pragma once

#include <cuda_runtime.h>

namespace mde {

// metric groups a launch needs (template mask: only the requested work is compiled in)
constexpr unsigned kGrpLog = 1u;    // MDE_Q_LOG10, MDE_Q_LNSQ
constexpr unsigned kGrpLog1p = 2u;  // MDE_Q_SLE
constexpr unsigned kGrpRel = 4u;    // MDE_Q_ABSREL, MDE_Q_SQREL, MDE_Q_RSQ
constexpr unsigned kGrpAll = 7u;

// ln(x) for finite x >= 2^-126 (used with x >= 1): exponent/mantissa split with the mantissa in
// [2/3, 4/3) and ln(1+f) = f - f^2/2 + f^3 g(f), g = own degree-7 near-minimax fit on [-1/3,1/3]
// (max abs error 6.1e-9, max rel error 1.6e-8 before fp32 rounding). NaN and +inf pass through.
__device__ __forceinline__ float ln_pos(float x) {
  const int ix = __float_as_int(x);
  const int e = (ix - 0x3f2aaaab) & 0xff800000;
  const float f = __int_as_float(ix - e) - 1.0f;
  const float fe = static_cast<float>(e) * 1.1920928955078125e-07f;  // exponent as float
  float g = -0.1242986634938813f;
  g = fmaf(g, f, 0.13433774844250851f);
  g = fmaf(g, f, -0.12287236520750586f);
  g = fmaf(g, f, 0.14116977926319346f);
  g = fmaf(g, f, -0.16673398305540069f);
  g = fmaf(g, f, 0.20003835432812128f);
  g = fmaf(g, f, -0.24999943056923857f);
  g = fmaf(g, f, 0.33333319832784086f);
  const float f2 = f * f;
  float r = fmaf(f2, fmaf(g, f, -0.5f), f);  // f - f^2/2 + f^3 g
  r = fmaf(fe, 0.69314718055994531f, r);
  return (x < __int_as_float(0x7f800000)) ? r : x;  // +inf, NaN unchanged
}

// 1/x with one Newton step on MUFU.RCP: relative error ~1e-7, unbiased enough for sums
__device__ __forceinline__ float rcp_nr(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return fmaf(y, fmaf(-x, y, 1.0f), y);
}

// ln(x) for any x: polynomial path for normal positive x, libdevice for the rest (0, negatives,
// denormals, inf, NaN) so that the reference's -inf / NaN results are reproduced
__device__ __forceinline__ float ln_any(float x) {
  if (x >= 1.17549435e-38f && x < __int_as_float(0x7f800000)) return ln_pos(x);
  return logf(x);
}

// out-of-line slow path of log_ratio (rare: non-positive, denormal, inf or NaN operands)
static __device__ __noinline__ float log_ratio_slow(float p, float t) { return logf(p) - logf(t); }

// ln(p) - ln(t) with ONE logarithm: ln(p * (1/t)). |error| <= ~2e-7 absolute for normal operands
// (the reference's own two logf calls carry ~1e-7 * |ln| each); anything else takes the slow path.
__device__ __forceinline__ float log_ratio(float p, float t) {
  const float q = p * rcp_nr(t);
  float d = ln_pos(q);
  if (!(q >= 1.17549435e-38f && q < __int_as_float(0x7f800000) && p > 0.f && t > 0.f)) d = log_ratio_slow(p, t);
  return d;
}

// fp32 tile accumulators of one thread: 8 float sums + 4 integer counts
struct MetricTile {
  float s_abs, s_sq, s_log10, s_sle, s_absrel, s_sqrel, s_rsq, s_lnsq;
  __device__ __forceinline__ void zero() {
    s_abs = s_sq = s_log10 = s_sle = s_absrel = s_sqrel = s_rsq = s_lnsq = 0.f;
  }
};
// n / c1 / c2 / c3 plus (fast mode) a packed 4 x 8-bit histogram of the threshold level of the
// valid pixels since the last unpack(): byte 0 = ratio < 1.25, byte 1 = [1.25, 1.5625),
// byte 2 = [1.5625, 1.953125), byte 3 = the rest. At most 255 pixels may be packed between unpacks.
struct MetricCounts {
  int n, c1, c2, c3;
  unsigned pk;
  __device__ __forceinline__ void zero() { n = c1 = c2 = c3 = 0; pk = 0u; }
  __device__ __forceinline__ void unpack() {
    const int b0 = pk & 0xffu, b1 = (pk >> 8) & 0xffu, b2 = (pk >> 16) & 0xffu, b3 = pk >> 24;
    c1 += b0;
    c2 += b0 + b1;
    c3 += b0 + b1 + b2;
    n += b0 + b1 + b2 + b3;
    pk = 0u;
  }
};

__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ln(q) for q >= 1 (NaN passes through): the SFU log2 away from 1, a 4-term series of ln(1+f) below
// 1 + 2^-5 where lg2.approx's +4e-8 absolute bias would otherwise dominate a small result.
__device__ __forceinline__ float ln_ge1_fast(float q, float l2) {
  const float f = q - 1.0f;
  const float ser = f * fmaf(f, fmaf(f, fmaf(f, -0.25f, 0.33333333f), -0.5f), 1.0f);
  return (f < 0.03125f) ? ser : l2 * 0.69314718055994531f;
}

// One pixel of the metric suite. Besides accumulating, returns through `L_out` the value
// |ln p - ln t| = ln(max/min) (fast mode with kGrpLog only; 0 otherwise) so that a fused loss can
// reuse the logarithm, and through `d_out` the difference p - t on the clamped prediction.
template <unsigned G, bool Ref>
__device__ __forceinline__ void metric_px_ex(float p, float t, MetricTile& a, MetricCounts& c, float& L_out,
                                             float& d_out) {
  const bool v = t > 0.f;                         // metrics.py:60
  p = (p < 1e-7f) ? 1e-7f : p;                    // clamp_min(pred, 1e-7), NaN preserved (metrics.py:59)
  // invalid pixels are replaced by p = t = 1: every float contribution below is then exactly 0
  const float pp = v ? p : 1.0f;
  const float tt = v ? t : 1.0f;
  const float d = pp - tt;
  const float hi = fmaxf(pp, tt), lo = fminf(pp, tt);
  const float ad = fabsf(d);
  a.s_abs += ad;
  a.s_sq = fmaf(d, d, a.s_sq);
  L_out = 0.f;
  d_out = d;
  if (Ref) {
    // == max(fl(p/t), fl(t/p)) (metrics.py:76); the fma adds 0*d so that a NaN prediction, which
    // fmaxf/fminf drop, still poisons r as torch.max would (exact no-op for finite d)
    const float r = fmaf(d, 0.0f, __fdiv_rn(hi, lo));
    c.n += v ? 1 : 0;
    c.c1 += (v && r < 1.25f) ? 1 : 0;             // strict '<' (metrics.py:77,82,87)
    c.c2 += (v && r < 1.5625f) ? 1 : 0;
    c.c3 += (v && r < 1.953125f) ? 1 : 0;
    if (G & kGrpLog) {
      a.s_log10 += fabsf(log10f(pp) - log10f(tt));            // metrics.py:90-91
      const float dl = logf(pp) - logf(tt);
      a.s_lnsq = fmaf(dl, dl, a.s_lnsq);
    }
    if (G & kGrpLog1p) {
      const float dl = log1pf(pp) - log1pf(tt);               // torchmetrics msle
      a.s_sle = fmaf(dl, dl, a.s_sle);
    }
    if (G & kGrpRel) {
      const float sq = d * d;
      a.s_absrel += __fdiv_rn(ad, tt);                        // metrics.py:97
      const float sr = __fdiv_rn(sq, tt);                     // metrics.py:103
      a.s_sqrel += sr;
      a.s_rsq += __fsqrt_rn(sr);                              // metrics.py:109
    }
  } else {
    // q ~ max/min (1.5 ulp), NaN predictions propagate through the 0*d term
    const float q = fmaf(d, 0.0f, hi * mufu_rcp(lo));
    const float l2 = mufu_lg2(q);
    // threshold level in the log domain: u = log_1.25(q); NaN -> 4 (beyond every threshold)
    const float u = fminf(l2 * 3.1062837195f, 4.0f);
    bool lt1 = u < 1.0f, lt2 = u < 2.0f, lt3 = u < 3.0f;
    const float kf = (u + 12582912.0f) - 12582912.0f;        // rint(u) via the 1.5*2^23 trick
    if (fabsf(u - kf) < 1e-5f && u > 0.5f && u < 3.5f) {      // within 1e-5 of a threshold: decide exactly
      const float r = __fdiv_rn(hi, lo);
      lt1 = r < 1.25f;
      lt2 = r < 1.5625f;
      lt3 = r < 1.953125f;
    }
    const unsigned inc = lt1 ? 1u : (lt2 ? 0x100u : (lt3 ? 0x10000u : 0x1000000u));
    c.pk += v ? inc : 0u;
    if (G & kGrpLog) {
      const float L = ln_ge1_fast(q, l2);                     // |ln p - ln t| = ln(max/min)
      L_out = L;
      a.s_log10 = fmaf(L, 0.43429448190325182f, a.s_log10);
      a.s_lnsq = fmaf(L, L, a.s_lnsq);
    }
    if (G & kGrpLog1p) {
      // |log1p p - log1p t| = ln((1+hi)/(1+lo))
      const float s1 = fmaf(d, 0.0f, (1.0f + hi) * mufu_rcp(1.0f + lo));
      const float L = ln_ge1_fast(s1 < 1.0f ? 1.0f : s1, mufu_lg2(s1));   // NaN stays NaN
      a.s_sle = fmaf(L, L, a.s_sle);
    }
    if (G & kGrpRel) {
      const float ar = ad * mufu_rcp(tt);
      a.s_absrel += ar;
      a.s_sqrel = fmaf(ar, ad, a.s_sqrel);
      a.s_rsq = fmaf(ad, mufu_rsq(tt), a.s_rsq);              // sqrt((p-t)^2/t) = |p-t| / sqrt(t)
    }
  }
}

template <unsigned G, bool Ref>
__device__ __forceinline__ void metric_px(float p, float t, MetricTile& a, MetricCounts& c) {
  float L, d;
  metric_px_ex<G, Ref>(p, t, a, c, L, d);
}

// finished values from raw sums (shared by the device finaliser and mde_metrics_finalize_host)
__host__ __device__ inline void metric_values(const double* raw, double* val) {
  const double n = raw[MDE_Q_NVALID];
  val[MDE_M_DELTA1] = raw[MDE_Q_D1] / n;
  val[MDE_M_DELTA2] = raw[MDE_Q_D2] / n;
  val[MDE_M_DELTA3] = raw[MDE_Q_D3] / n;
  val[MDE_M_MAE] = raw[MDE_Q_ABS] / n;
  val[MDE_M_MSE] = raw[MDE_Q_SQ] / n;
  val[MDE_M_LOG10] = raw[MDE_Q_LOG10] / n;
  val[MDE_M_MSLE] = raw[MDE_Q_SLE] / n;
  val[MDE_M_ABSREL] = raw[MDE_Q_ABSREL] / n;
  val[MDE_M_SQREL] = raw[MDE_Q_SQREL] / n;
  val[MDE_M_RMSE] = raw[MDE_Q_RSQ] / n;
  val[MDE_M_RMSE_TRUE] = sqrt(raw[MDE_Q_SQ] / n);
  val[MDE_M_RMSE_LOG] = sqrt(raw[MDE_Q_LNSQ] / n);
}

}  // namespace mde
