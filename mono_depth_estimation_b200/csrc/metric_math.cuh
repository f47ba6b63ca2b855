// metric_math.cuh - per-pixel arithmetic of the masked error metrics.
//
// Two evaluation modes of the SAME quantities (include/mde_b200.h, MDE_Q_*):
//   Ref  : the reference's own op sequence (metrics.py:75-109 + torchmetrics closed forms):
//          log10f(p)-log10f(t), log1pf(p)-log1pf(t), IEEE divides everywhere.
//   Fast : algebraically equal forms built on the SFU (MUFU.LG2 / RSQ), ~35 issue slots per pixel; used
//          by default because the suite in Ref form is issue-bound, not HBM-bound, on B200.
//          Logarithmic sums are carried in log2 units and scaled once per flush (tile_scale()).
//
// Bit-exact delta counts in BOTH modes, and in Fast mode without a divide or a branch. The reference
// counts max(fl(p/t), fl(t/p)) < T for T = 1.25^k, which equals RN(hi/lo) < T with hi = max, lo = min
// (round-to-nearest is monotone). All three T lie in (1,2) and have an even last mantissa bit, so
//     RN(x) < T   <=>   x < T - 2^-24   <=>   lo*T - hi > lo * 2^-24            (x = hi/lo >= 1, lo > 0)
// (the midpoint between T and its predecessor rounds to T). fmaf(lo, T, -hi) rounds lo*T - hi once; that
// value is a multiple of 2^-6 ulp(lo), hence exactly representable whenever it is within 2^18 ulp(lo) of
// the right-hand side, and lo*2^-24 is exact - so the fp32 comparison decides the real inequality for
// every input (tiny lo with hi/lo astronomically large cannot be near a threshold). NaN predictions make
// hi NaN (max.NaN) and fail every comparison, as NaN < T does in the reference.
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

__device__ __forceinline__ float fmax_nan(float x, float y) {   // NaN if either operand is NaN
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(y));
  return r;
}

// Factor that turns fp32 tile sum q (MetricTile order: abs, sq, log10, sle, absrel, sqrel, rsq, lnsq)
// into the unit of the raw quantity: Fast mode carries its logarithms in log2 units.
template <bool Ref>
__host__ __device__ constexpr float tile_scale(int q) {
  return Ref ? 1.0f : (q == 2 ? 0.30102999566398120f : ((q == 3 || q == 7) ? 0.48045301391820142f : 1.0f));
}

// Fast mode relies on MUFU.LG2/RSQ with flush-to-zero: a quad that holds a valid SUBNORMAL target must
// be evaluated in Ref mode instead (one predicate per quad: 3 FMNMX + 1 FSETP). Never true for depth
// in metres; kept so that the fast path has no input it silently gets wrong.
__device__ __forceinline__ bool metric_quad_needs_ref(const float4& t) {
  const float a = (t.x > 0.f) ? t.x : 1.0f, b = (t.y > 0.f) ? t.y : 1.0f;
  const float c = (t.z > 0.f) ? t.z : 1.0f, d = (t.w > 0.f) ? t.w : 1.0f;
  return fminf(fminf(a, b), fminf(c, d)) < 1.17549435e-38f;
}

// One pixel of the metric suite. Besides accumulating, returns through `L_out` (fast mode with kGrpLog
// only; 0 otherwise) the signed value log2(p) - log2(t) of the clamped prediction, so that a fused loss
// can reuse the logarithms, and through `d_out` the difference p - t on the clamped prediction.
template <unsigned G, bool Ref>
__device__ __forceinline__ void metric_px_ex(float p, float t, MetricTile& a, MetricCounts& c, float& L_out,
                                             float& d_out) {
  const bool v = t > 0.f;                         // metrics.py:60
  // invalid pixels are replaced by p = t = 1: every float contribution below is then exactly 0
  const float ps = v ? p : 1.0f;
  const float tt = v ? t : 1.0f;
  const float pp = (ps < 1e-7f) ? 1e-7f : ps;     // clamp_min(pred, 1e-7), NaN preserved (metrics.py:59)
  const float d = pp - tt;
  const float hi = Ref ? fmaxf(pp, tt) : fmax_nan(pp, tt), lo = fminf(pp, tt);
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
    // level of the ratio among the thresholds, exact (see the header): 1 FMUL + 3 FFMA + 3 FSETP
    const float e = lo * 5.9604644775390625e-08f;
    const bool lt1 = fmaf(lo, 1.25f, -hi) > e;                 // strict '<' (metrics.py:77,82,87)
    const bool lt2 = fmaf(lo, 1.5625f, -hi) > e;
    const bool lt3 = fmaf(lo, 1.953125f, -hi) > e;
    const unsigned inc = lt1 ? 1u : (lt2 ? 0x100u : (lt3 ? 0x10000u : 0x1000000u));
    c.pk += v ? inc : 0u;
    if (G & kGrpLog) {
      const float dl = mufu_lg2(pp) - mufu_lg2(tt);           // log2 p - log2 t, signed
      L_out = dl;
      a.s_log10 += fabsf(dl);                                 // x log10(2) at the flush
      a.s_lnsq = fmaf(dl, dl, a.s_lnsq);                      // x ln(2)^2 at the flush
    }
    if (G & kGrpLog1p) {
      const float d1 = mufu_lg2(1.0f + pp) - mufu_lg2(1.0f + tt);
      a.s_sle = fmaf(d1, d1, a.s_sle);                        // x ln(2)^2 at the flush
    }
    if (G & kGrpRel) {
      const float rs = mufu_rsq(tt);
      const float ar = ad * (rs * rs);                        // |p-t| / t
      a.s_absrel += ar;
      a.s_sqrel = fmaf(ar, ad, a.s_sqrel);
      a.s_rsq = fmaf(ad, rs, a.s_rsq);                        // sqrt((p-t)^2/t) = |p-t| / sqrt(t)
    }
  }
}

// Rare-quad path of the fast kernels: one pixel in Ref arithmetic, returned by value (nothing of the
// caller's register-resident accumulators has its address taken) and added in the fast tile's units.
struct MetricContrib {
  MetricTile s;
  MetricCounts c;
};
template <unsigned G>
static __device__ __noinline__ MetricContrib metric_px_ref_contrib(float p, float t) {
  MetricContrib r;
  r.s.zero();
  r.c.zero();
  float L, d;
  metric_px_ex<G, true>(p, t, r.s, r.c, L, d);
  return r;
}
__device__ __forceinline__ void metric_add_contrib(const MetricContrib& r, MetricTile& a, MetricCounts& c) {
  a.s_abs += r.s.s_abs; a.s_sq += r.s.s_sq;
  a.s_log10 += r.s.s_log10 * (1.0f / tile_scale<false>(2));
  a.s_sle += r.s.s_sle * (1.0f / tile_scale<false>(3));
  a.s_absrel += r.s.s_absrel; a.s_sqrel += r.s.s_sqrel; a.s_rsq += r.s.s_rsq;
  a.s_lnsq += r.s.s_lnsq * (1.0f / tile_scale<false>(7));
  c.n += r.c.n; c.c1 += r.c.c1; c.c2 += r.c.c2; c.c3 += r.c.c3;
}
template <unsigned G>
__device__ __forceinline__ void metric_px_ref_into_fast(float p, float t, MetricTile& a, MetricCounts& c) {
  metric_add_contrib(metric_px_ref_contrib<G>(p, t), a, c);
}

template <unsigned G, bool Ref>
__device__ __forceinline__ void metric_px(float p, float t, MetricTile& a, MetricCounts& c) {
  float L, d;
  metric_px_ex<G, Ref>(p, t, a, c, L, d);
}

// ---- "lean" fast form (round 2) ---------------------------------------------------------------------
// The fast form above costs ~20 ALU-pipe instructions per pixel (FSETP / FSEL / FMNMX / SEL / IADD issue at
// half the rate of the FMA pipe on B200) and was issue-bound. The lean form moves everything it can to the
// FMA pipe and drops what the caller can guarantee:
//   * |p - t| = hi - lo (the same subtraction, sign dropped): no separate difference, no abs;
//   * the three threshold counts are FLOAT counters fed by FSET/FADD (exact below 2^24) instead of a packed
//     byte histogram (3 SEL + SEL + IADD); invalid pixels (replaced by p = t = 1, ratio 1) are counted by all
//     three; the valid pixels are counted the same way (FMUL.SAT + FADD) and the caller removes the
//     (pixels seen - valid) surplus once per flush;
//   * NOCLAMP: the caller's per-quad rare test already sent p < 1e-7 (and NaN) to the exact path, so
//     clamp_min(pred, 1e-7) is the identity here;
//   * kGrpRsq: only the 'rmse' quirk sum of the REL group (the reference's default train metrics need no
//     absrel / sqrel).
// Per pixel with {log, rsq}: 11 ALU-pipe + 15 FMA-pipe + 3 MUFU instructions.
constexpr unsigned kGrpRsq = 8u;    // MDE_Q_RSQ only (implied by kGrpRel)
constexpr unsigned kGrpMask = 15u;

struct MetricAcc {
  float s_abs, s_sq, s_log10, s_sle, s_absrel, s_sqrel, s_rsq, s_lnsq;   // MetricTile order
  float c1, c2, c3;   // pixels below each threshold, INCLUDING the invalid ones seen by the lean form
  float nval;         // valid pixels seen by the lean form (float counter)
  int n_x, c1_x, c2_x, c3_x;   // exact counts booked by the rare (reference-arithmetic) path
  __device__ __forceinline__ void zero() {
    s_abs = s_sq = s_log10 = s_sle = s_absrel = s_sqrel = s_rsq = s_lnsq = 0.f;
    c1 = c2 = c3 = nval = 0.f;
    n_x = c1_x = c2_x = c3_x = 0;
  }
  // exact counts after `lean_px` pixels went through metric_px_lean
  __device__ __forceinline__ int n_valid(int) const { return static_cast<int>(nval) + n_x; }
  __device__ __forceinline__ int count(int k, int lean_px) const {
    const float c = (k == 1) ? c1 : (k == 2) ? c2 : c3;
    const int x = (k == 1) ? c1_x : (k == 2) ? c2_x : c3_x;
    return static_cast<int>(c) - (lean_px - static_cast<int>(nval)) + x;
  }
  __device__ __forceinline__ float sum(int q) const {
    return (q == 0) ? s_abs : (q == 1) ? s_sq : (q == 2) ? s_log10 : (q == 3) ? s_sle
         : (q == 4) ? s_absrel : (q == 5) ? s_sqrel : (q == 6) ? s_rsq : s_lnsq;
  }
};

// One pixel, lean form. Returns log2(p) - log2(t) (0 off the mask) when G has kGrpLog.
template <unsigned G, bool NOCLAMP>
__device__ __forceinline__ float metric_px_lean(float p, float t, MetricAcc& a) {
  const bool v = t > 0.f;                          // metrics.py:60
  const float tt = v ? t : 1.0f;
  float pp = v ? p : 1.0f;
  if (!NOCLAMP) pp = fmax_nan(pp, 1e-7f);          // clamp_min(pred, 1e-7), NaN preserved (metrics.py:59)
  const float hi = fmax_nan(pp, tt), lo = fminf(pp, tt);
  const float ad = hi - lo;                        // |p - t|
  a.s_abs += ad;
  a.s_sq = fmaf(ad, ad, a.s_sq);
  const float e = lo * 5.9604644775390625e-08f;
  a.c1 += (fmaf(lo, 1.25f, -hi) > e) ? 1.0f : 0.0f;      // strict '<' (metrics.py:77,82,87), exact (header)
  a.c2 += (fmaf(lo, 1.5625f, -hi) > e) ? 1.0f : 0.0f;
  a.c3 += (fmaf(lo, 1.953125f, -hi) > e) ? 1.0f : 0.0f;
  // valid count on the FMA pipe: sat(t * 2^126) is exactly 1 for every normal t > 0 and 0 for t <= 0 / NaN (a valid
  // SUBNORMAL target never reaches the lean form: the caller's rare test sends its quad to the exact path)
  a.nval += __saturatef(t * 8.507059173023462e37f);
  float dl = 0.f;
  // 1/sqrt(t) serves the REL / RSQ sums AND the logarithm: log2 p - log2 t = log2((p rs) rs) is ONE MUFU.LG2 instead of
  // two (MUFU issues at 1/8 of the FMA rate: the SFU ops are most of this pixel's issue cycles). The SFU reciprocal
  // square root is good to 1.2e-7 relative (profiles/r01_mathlab_sfu_accuracy.jsonl), so the ratio carries 2.5e-7 and
  // the logarithm ~7e-7 absolute in log2 units against ~5e-7 for the difference of two MUFU.LG2 - the sums move by
  // ~3e-7 relative (measured for the same form with MUFU.RCP: aggregate 2.8e-7), tolerance 1e-5.
  float rs = 0.f;
  if (G & (kGrpRel | kGrpRsq)) rs = mufu_rsq(tt);
  if (G & kGrpLog) {
    dl = (G & (kGrpRel | kGrpRsq)) ? mufu_lg2((pp * rs) * rs) : mufu_lg2(pp) - mufu_lg2(tt);
    a.s_log10 += fabsf(dl);                        // x log10(2) at the flush
    a.s_lnsq = fmaf(dl, dl, a.s_lnsq);             // x ln(2)^2 at the flush
  }
  if (G & kGrpLog1p) {
    const float d1 = mufu_lg2(1.0f + pp) - mufu_lg2(1.0f + tt);
    a.s_sle = fmaf(d1, d1, a.s_sle);
  }
  if (G & kGrpRel) {
    const float ar = ad * (rs * rs);
    a.s_absrel += ar;
    a.s_sqrel = fmaf(ar, ad, a.s_sqrel);
    a.s_rsq = fmaf(ad, rs, a.s_rsq);
  } else if (G & kGrpRsq) {
    a.s_rsq = fmaf(ad, rs, a.s_rsq);               // sqrt((p-t)^2/t) = |p-t| / sqrt(t)
  }
  return dl;
}

// rare path for the lean accumulators: one pixel in reference arithmetic, booked in the lean units
__device__ __forceinline__ void metric_add_contrib(const MetricContrib& r, MetricAcc& a) {
  a.s_abs += r.s.s_abs; a.s_sq += r.s.s_sq;
  a.s_log10 += r.s.s_log10 * (1.0f / tile_scale<false>(2));
  a.s_sle += r.s.s_sle * (1.0f / tile_scale<false>(3));
  a.s_absrel += r.s.s_absrel; a.s_sqrel += r.s.s_sqrel; a.s_rsq += r.s.s_rsq;
  a.s_lnsq += r.s.s_lnsq * (1.0f / tile_scale<false>(7));
  a.n_x += r.c.n; a.c1_x += r.c.c1; a.c2_x += r.c.c2; a.c3_x += r.c.c3;
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
