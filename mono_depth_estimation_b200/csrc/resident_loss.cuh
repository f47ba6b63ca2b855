// resident_loss.cuh - small inputs: L1 / MSE / berHu / Laina berHu / SILog (+ metric suite) forward+backward with every
// pixel of the call held in REGISTERS between the phases. Included by resident_loss.cu only.
//
//   MaskedL1Loss criteria.py:80-90, MaskedMSELoss :67-77, berHuLoss :111-133, LainaBerHuLoss :476-506, silog_loss :724-732,
//   MetricComputation.compute metrics.py:58-67 (pooled, fused into the same pass)
//
// The generic kernel (losses_kernel.cuh) is a persistent 296 x 512 grid that walks its tiles once per phase: at the
// reference's own C1 size (8x1x228x304 = 6.6 MB, 1 us of HBM time) that is 2-3 sweeps through L2, one or two ticket
// barriers (four dependent L2 round trips each) and tile claims: 10.9 us (L1) to 20 us (Laina). Here one CTA of 1024
// threads per SM loads at most R quads per thread ONCE (148 x 1024 x 4 x R pixels: 606 k with R = 1, 1.2 M with R = 2),
// and everything that follows - the global max of berHu / Laina, the masked sums, the metric suite, the gradient - is
// evaluated from those registers. The totals travel through the slot exchange of the SS SILog kernel (every CTA
// publishes self-validating words and gathers all slots itself: one store, one load, one block reduction after the
// last CTA is ready); the maxima through a one-word-per-CTA exchange of the same kind in a separate slot range.
#pragma once
#include "losses_kernel.cuh"

namespace mde {
namespace {

constexpr int kRsThreads = 1024;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsMaxBase = 192;                       // the max exchange uses slots [192, 384): the sum exchange of the
constexpr int kRsMaxGrid = kSlotCtas - kRsMaxBase;    // same launch may overwrite a slot a slow CTA still polls otherwise

template <int KIND, unsigned MG, int R>
__global__ void __launch_bounds__(kRsThreads, 1) resident_loss_kernel(LossArgs a) {
  __shared__ double sm_d[(MG ? 16 : 1) * kRsWarps];
  __shared__ double sm_met[16];
  __shared__ double sm_own[4 * kRsWarps];
  __shared__ double sm_gather[kRsWarps * 4];
  __shared__ double sm_tot[4];
  __shared__ float sm_k[4];
  __shared__ float sm_f[kRsWarps];
  __shared__ float sm_f2[kRsWarps];
  __shared__ unsigned sm_epoch;
  constexpr bool kNeedMax = (KIND == MDE_LOSS_BERHU || KIND == MDE_LOSS_LAINA_BERHU);
  constexpr unsigned kRefG = MG & 7u;

  const float* __restrict__ pred = static_cast<const float*>(a.pred);
  const float* __restrict__ gt = a.gt;
  float* grad = static_cast<float*>(a.grad);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = static_cast<int>(gridDim.x), cta = static_cast<int>(blockIdx.x);
  const int nq = static_cast<int>(a.n >> 2);
  const int q0 = cta * kRsThreads + tid;                          // this thread's quad k: q0 + k * qs
  const int qs = G * kRsThreads;

  pdl_wait();   // launched with launch_pdl: nothing a predecessor wrote may be read before this point
  Ws ws = ws_view(a.ws);
  unsigned epoch_reg = 0u;
  if (tid == 0) epoch_reg = __ldcg(&ws.hdr->epoch);               // issued first: needed by the first exchange
  float4 p4[R], t4[R];
  bool has[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int q = q0 + k * qs;
    has[k] = q < nq;
    if (has[k]) {
      p4[k] = __ldcs(reinterpret_cast<const float4*>(pred) + q);
      t4[k] = __ldcs(reinterpret_cast<const float4*>(gt) + q);
    } else {
      p4[k] = make_float4(1.f, 1.f, 1.f, 1.f);
      t4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // n % 4 tail: one pixel per thread of the last CTA
  const int64_t ti = (static_cast<int64_t>(nq) << 2) + tid;
  const bool has_tail = (cta == G - 1) && ti < a.n;
  float tp = 1.f, tt = 0.f;
  if (has_tail) {
    tp = __ldg(pred + ti);
    tt = __ldg(gt + ti);
  }
  if (tid == 0) sm_epoch = epoch_reg;

  // The metric suite's per-pixel arithmetic (~45 instructions per pixel: 0.7 us of issue slots per quad and thread at
  // this size) runs while the FIRST exchange of the kernel travels (one L2 round trip with nothing else to do): behind
  // the publication of the CTA's maximum (berHu / Laina) or of its loss sums (the other kinds).
  MetricAcc acc;
  acc.zero();
  int lean_q = 0;
  auto metric_eval = [&] {
    if constexpr (MG != 0) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (has[k]) {
          if (metric_quad_needs_ref(t4[k])) {                     // a valid subnormal target: exact arithmetic
            metric_add_contrib(metric_px_ref_contrib<kRefG>(p4[k].x, t4[k].x), acc);
            metric_add_contrib(metric_px_ref_contrib<kRefG>(p4[k].y, t4[k].y), acc);
            metric_add_contrib(metric_px_ref_contrib<kRefG>(p4[k].z, t4[k].z), acc);
            metric_add_contrib(metric_px_ref_contrib<kRefG>(p4[k].w, t4[k].w), acc);
          } else {
            metric_px_lean<MG, false>(p4[k].x, t4[k].x, acc);
            metric_px_lean<MG, false>(p4[k].y, t4[k].y, acc);
            metric_px_lean<MG, false>(p4[k].z, t4[k].z, acc);
            metric_px_lean<MG, false>(p4[k].w, t4[k].w, acc);
            ++lean_q;
          }
        }
      }
      if (has_tail) metric_add_contrib(metric_px_ref_contrib<kRefG>(tp, tt), acc);
    }
  };

  // ---------------- phase A0: global max (berHu: max(p - t) over ALL pixels; Laina: max n_i) -------------------
  float cthr = 0.f, gmax = 0.f;
  float4 r4[R];                // Laina: the signed residuals r_i (criteria.py:488-494), one logarithm per pixel and call
  float tr = 0.f;
#pragma unroll
  for (int k = 0; k < R; ++k) r4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (kNeedMax) {
    float mx = -INFINITY;
    bool saw_nan = false;
    auto px_max = [&](float p, float t) -> float {
      float x, r = 0.f;
      if constexpr (KIND == MDE_LOSS_BERHU) {
        x = p - t;                                                // criteria.py:118 - signed, unmasked
      } else {
        x = laina_resid(p, t, t > 0.f, a.use_logs != 0, a.clamp_val, r);
      }
      saw_nan |= (x != x);
      mx = fmaxf(mx, x);
      return r;
    };
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (has[k]) {
        r4[k].x = px_max(p4[k].x, t4[k].x); r4[k].y = px_max(p4[k].y, t4[k].y);
        r4[k].z = px_max(p4[k].z, t4[k].z); r4[k].w = px_max(p4[k].w, t4[k].w);
      }
    }
    if (has_tail) tr = px_max(tp, tt);
    const float qnan = __int_as_float(0x7fc00000);
    mx = warp_max(mx);
    const bool wn = __any_sync(0xffffffffu, saw_nan);
    if (lane == 0) sm_f[warp] = wn ? qnan : mx;
    __syncthreads();
    const unsigned seq1 = sm_epoch * 4u + 1u;
    if (warp == 0) {
      float x = sm_f[lane];
      bool xn = x != x;
      x = warp_max(xn ? -INFINITY : x);
      xn = __any_sync(0xffffffffu, xn);
      if (lane == 0)
        st_relaxed_u64(ws.slots + static_cast<size_t>(kRsMaxBase + cta) * 8,
                       (static_cast<unsigned long long>(seq1) << 32) | __float_as_uint(xn ? qnan : x));
    }
    metric_eval();                                                // while the maxima travel
    float gm = -INFINITY;
    bool gn = false;
    if (tid < G) {                                                // one slot per thread, every CTA gathers all of them
      const unsigned long long* w = ws.slots + static_cast<size_t>(kRsMaxBase + tid) * 8;
      unsigned long long v;
      do {
        v = ld_relaxed_u64(w);
      } while (static_cast<unsigned>(v >> 32) != seq1);
      const float f = __uint_as_float(static_cast<unsigned>(v));
      gn = f != f;
      gm = gn ? -INFINITY : f;
    }
    gm = warp_max(gm);
    gn = __any_sync(0xffffffffu, gn);
    if (lane == 0) sm_f2[warp] = gn ? qnan : gm;
    __syncthreads();
    {
      float x = sm_f2[lane];
      bool xn = x != x;
      x = warp_max(xn ? -INFINITY : x);
      xn = __any_sync(0xffffffffu, xn);
      gmax = xn ? qnan : x;
    }
    cthr = 0.2f * gmax;                                           // criteria.py:119 / :496 (fp32 product)
  }

  // ---------------- phase A1: masked sums and counts from the registers ------------------------------------------
  float s0 = 0.f, s1 = 0.f, c0 = 0.f, c1 = 0.f;                   // a thread sees <= 4 R + 1 pixels: float counts are exact
  // Laina's two quotients have the SAME divisor for every pixel: one IEEE reciprocal per thread instead of two divisions
  // per pixel (a division is a ~30-instruction subroutine: 8 per thread were ~1 us of issue slots at this size)
  const float laD = 2.f * cthr + 1e-9f, laInvD = 1.f / laD, laInvD2 = laInvD * laInvD;
  auto px_sum = [&](float p, float t, float r) {
    if constexpr (KIND == MDE_LOSS_L1) {
      const bool v = t > 0.f;
      s0 += v ? fabsf(t - p) : 0.f;
      c0 += v ? 1.f : 0.f;
    } else if constexpr (KIND == MDE_LOSS_MSE) {
      const bool v = t > 0.f;
      const float d = t - p;
      s0 += v ? d * d : 0.f;
      c0 += v ? 1.f : 0.f;
    } else if constexpr (KIND == MDE_LOSS_SILOG) {
      bool v;
      const float d = silog_resid(p, t, v);                       // criteria.py:730-731, 0 off the mask t > 0.01
      s0 += d;
      s1 = fmaf(d, d, s1);
      c0 += v ? 1.f : 0.f;
    } else if constexpr (KIND == MDE_LOSS_BERHU) {
      const bool v = t > 0.f;
      const float ad = fabsf(t - p);
      const bool hub = v && (ad > cthr);                          // criteria.py:126
      s0 += v ? ad : 0.f;
      s1 += hub ? ad * ad : 0.f;
      c0 += v ? 1.f : 0.f;
      c1 += hub ? 1.f : 0.f;
    } else {  // LAINA
      const bool m = t > 0.f;
      const float ni = fabsf(r) * (m ? 1.f : 0.f);                // n_i = |r_i| m_i
      const bool big = !(ni < cthr);                              // criteria.py:497-498
      const float num = fmaf(ni, ni, cthr * cthr);
      s0 += big ? num * laInvD : ni;
      s1 += big ? (2.f * cthr * laD - 2.f * num) * laInvD2 : 0.f;  // d/dc of the quadratic branch
      c0 += m ? 1.f : 0.f;
      c1 += (ni == gmax) ? 1.f : 0.f;
    }
  };
#pragma unroll
  for (int k = 0; k < R; ++k) {
    if (has[k]) {
      px_sum(p4[k].x, t4[k].x, r4[k].x); px_sum(p4[k].y, t4[k].y, r4[k].y);
      px_sum(p4[k].z, t4[k].z, r4[k].z); px_sum(p4[k].w, t4[k].w, r4[k].w);
    }
  }
  if (has_tail) px_sum(tp, tt, tr);
  // without a max exchange: two quads per thread hide their metric arithmetic behind the sum exchange (8x300x400, L1:
  // 9.3 -> 8.4 us); with one quad the publication would only be delayed (C1: 6.8 -> 7.1 us), so it comes first
  constexpr bool kEvalBehindSums = !kNeedMax && R > 1;
  if constexpr (!kNeedMax && !kEvalBehindSums) metric_eval();
  {
    float run[4] = {s0, s1, c0, c1};
    const float tot = warp_multi_sum<4>(run);                     // quantity (lane >> 3) & 3
    if ((lane & 7) == 0) sm_own[(lane >> 3) * kRsWarps + warp] = static_cast<double>(tot);
    __syncthreads();
  }
  // launch parity: this launch uses workspace set `par`; CTA 0 cleans the OTHER set (used by the previous cooperative
  // launch, which has completed) for the next one - what coop_prologue does
  const unsigned epoch = sm_epoch;
  const int par = static_cast<int>(epoch & 1u);
  double* gacc = ws.gacc + par * kGacc;
  unsigned* ukey = ws.ukey + par * kUkey;
  if (cta == 0) {
    const int o = par ^ 1;
    for (int i = tid; i < kGacc; i += kRsThreads) ws.gacc[o * kGacc + i] = 0.0;
    for (int i = tid; i < kUkey; i += kRsThreads) ws.ukey[o * kUkey + i] = 0u;
    if (tid == 0) {
      const unsigned dirty = __ldcg(&ws.hdr->dirty[o]);
      if (dirty) {
        const unsigned cap = __ldcg(&ws.hdr->max_images);
        double* rows = ws.iacc + static_cast<size_t>(1 + o) * cap * kIacc;
        for (size_t i = 0; i < static_cast<size_t>(dirty) * kIacc; ++i) rows[i] = 0.0;
        ws.hdr->dirty[o] = 0u;
      }
    }
  }
  // pooled metric sums of this CTA -> 12 doubles published as self-validating words in the CTA's row of ws.mslots (the
  // finalising CTA gathers the rows at the end of the kernel). No atomics: 148 CTAs adding into the same 12 addresses
  // serialise in L2 (~2 us, with a fence and an arrival counter behind them), which a kernel this short cannot hide.
  const unsigned seq3 = epoch * 4u + 3u;
  auto flush_metrics = [&] {
    if constexpr (MG != 0) {
      const int lean_px = 4 * lean_q;
      const int r0 = __reduce_add_sync(0xffffffffu, acc.n_valid(lean_px)), r1 = __reduce_add_sync(0xffffffffu, acc.count(1, lean_px));
      const int r2 = __reduce_add_sync(0xffffffffu, acc.count(2, lean_px)), r3 = __reduce_add_sync(0xffffffffu, acc.count(3, lean_px));
      if (lane == 0) {
        sm_d[0 * kRsWarps + warp] = r0; sm_d[1 * kRsWarps + warp] = r1;
        sm_d[2 * kRsWarps + warp] = r2; sm_d[3 * kRsWarps + warp] = r3;
      }
      float v8[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v8[q] = acc.sum(q);
      const float sq = warp_multi_sum<8>(v8);                     // quantity (lane >> 2) & 7
      if ((lane & 3) == 0) {
        const int q = lane >> 2;
        sm_d[(4 + q) * kRsWarps + warp] = static_cast<double>(sq * tile_scale<false>(q));
      }
      __syncthreads();
      if (warp < 12) {                                            // warp q sums quantity q over the 32 warps (a serial pass
        const double tot = warp_sum(sm_d[warp * kRsWarps + lane]);  // by 12 threads cost ~0.5 us in front of the gather)
        if (lane == 0) {
          const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(tot));
          const unsigned long long tag = static_cast<unsigned long long>(seq3) << 32;
          unsigned long long* w2 = ws.mslots + static_cast<size_t>(cta) * kMetSlotWords + warp * 2;
          st_relaxed_u64(w2, tag | (b >> 32));
          st_relaxed_u64(w2 + 1, tag | (b & 0xffffffffull));
        }
      }
    }
  };

  // ---------------- all-reduce of the totals; every CTA derives the coefficients itself --------------------------
  grid_sum4_counted<kRsWarps>(ws.slots, ukey + 2, epoch * 4u + 2u, sm_own, sm_gather, sm_tot, [&] {
    if constexpr (kEvalBehindSums) metric_eval();                 // while the loss sums travel
    flush_metrics();
  });
  if (tid < 32) __syncwarp();   // sm_tot was written by threads of warp 0
  if (tid == 0) {
    const double S0 = sm_tot[0], S1 = sm_tot[1], N0 = sm_tot[2], N1 = sm_tot[3];
    const float gs = a.grad_scale;
    double loss;
    float k1 = 0.f, k2 = 0.f, k3 = 0.f;
    if constexpr (KIND == MDE_LOSS_L1) {
      const double inv = 1.0 / N0;
      loss = S0 * inv;
      k1 = gs * static_cast<float>(inv);
    } else if constexpr (KIND == MDE_LOSS_MSE) {
      const double inv = 1.0 / N0;
      loss = S0 * inv;
      k1 = 2.0f * gs * static_cast<float>(inv);
    } else if constexpr (KIND == MDE_LOSS_SILOG) {
      const double inv = 1.0 / N0;
      const double dm = S0 * inv, qm = S1 * inv;
      const double var = qm - static_cast<double>(a.vf) * dm * dm;   // the cancellation stays in fp64
      const float sd = sqrtf(static_cast<float>(var));
      loss = 10.0 * static_cast<double>(sd);
      k1 = 10.0f * gs * static_cast<float>(inv) / sd;                 // dL/dd_i = k1 * (d_i - k2)
      k2 = a.vf * static_cast<float>(dm);
    } else if constexpr (KIND == MDE_LOSS_BERHU) {
      const double inv = 1.0 / (N0 + N1);
      loss = (S0 + S1) * inv;                                     // mean of the concatenation (criteria.py:131)
      k1 = gs * static_cast<float>(inv);
    } else {
      const double inv = a.size_average ? 1.0 / N0 : 1.0;
      loss = S0 * inv;
      k1 = gs * static_cast<float>(inv);
      k2 = gs * 0.2f * static_cast<float>(S1 * inv) / static_cast<float>(N1);  // share of dL/dc per tied maximum
      k3 = 2.f * cthr + 1e-9f;
    }
    sm_k[0] = k1; sm_k[1] = k2; sm_k[2] = k3; sm_k[3] = 0.f;
    if (cta == 0) {
      *a.loss_out = static_cast<float>(loss);
      if (a.totals_out) {
        a.totals_out[0] = S0; a.totals_out[1] = S1; a.totals_out[2] = N0; a.totals_out[3] = N1;
        a.totals_out[4] = static_cast<double>(gmax); a.totals_out[5] = loss;
      }
      ws.hdr->epoch = epoch + 1u;
    }
  }
  __syncthreads();
  pdl_trigger();   // a dependent launch may start filling the SMs this grid leaves
  const float k1 = sm_k[0], k2 = sm_k[1], k3 = sm_k[2];

  // ---------------- gradient from the registers ------------------------------------------------------------------
  if (grad != nullptr) {
    const float inv_k3 = (KIND == MDE_LOSS_LAINA_BERHU) ? 1.f / k3 : 0.f;
    auto px_grad = [&](float p, float t, float r) -> float {
      if constexpr (KIND == MDE_LOSS_L1) {
        return (t > 0.f) ? -sgn(t - p) * k1 : 0.f;
      } else if constexpr (KIND == MDE_LOSS_MSE) {
        return (t > 0.f) ? -(t - p) * k1 : 0.f;
      } else if constexpr (KIND == MDE_LOSS_SILOG) {
        bool v;
        const float d = silog_resid(p, t, v);
        return v ? k1 * (d - k2) * rcp_nr(p) : 0.f;
      } else if constexpr (KIND == MDE_LOSS_BERHU) {
        const bool v = t > 0.f;
        const float d = t - p;
        const float ad = fabsf(d);
        const bool hub = v && (ad > cthr);
        return v ? -sgn(d) * (hub ? fmaf(2.f, ad, 1.f) : 1.f) * k1 : 0.f;
      } else {
        const bool m = t > 0.f;
        const float ni = fabsf(r) * (m ? 1.f : 0.f);
        const bool big = !(ni < cthr);
        float dn = (big ? 2.f * ni * inv_k3 : 1.f) * k1;
        if (ni == gmax) dn += k2;
        float dp = m ? sgn(r) : 0.f;                              // dn_i/dp = sign(r) m [p >= cv] / p
        if (a.use_logs) dp = (p >= a.clamp_val) ? dp * rcp_nr(p) : 0.f;
        return dn * dp;
      }
    };
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (has[k]) {
        float4 g;
        g.x = px_grad(p4[k].x, t4[k].x, r4[k].x); g.y = px_grad(p4[k].y, t4[k].y, r4[k].y);
        g.z = px_grad(p4[k].z, t4[k].z, r4[k].z); g.w = px_grad(p4[k].w, t4[k].w, r4[k].w);
        __stcs(reinterpret_cast<float4*>(grad) + (q0 + k * qs), g);
      }
    }
    if (has_tail) grad[ti] = px_grad(tp, tt, tr);
  }

  // pooled metric values (one mean over all valid pixels of the call, metrics.py:58-67): the LAST CTA gathers every
  // CTA's row of 12 sums (one quantity per thread: tid & 15; <= 3 rows per thread, all loads in flight before the first
  // is looked at), reduces them in a fixed order and its first warp forms the values
  if constexpr (MG != 0) {
    if (cta == G - 1) {
      const int q = tid & 15;
      constexpr int kIt = (kMetSlotCtas * 16 + kRsThreads - 1) / kRsThreads;
      unsigned long long hi[kIt], lo[kIt];
#pragma unroll
      for (int u = 0; u < kIt; ++u) {
        const int c = (tid >> 4) + (kRsThreads / 16) * u;
        if (q < 12 && c < G) {
          hi[u] = ld_relaxed_u64(ws.mslots + static_cast<size_t>(c) * kMetSlotWords + q * 2);
          lo[u] = ld_relaxed_u64(ws.mslots + static_cast<size_t>(c) * kMetSlotWords + q * 2 + 1);
        }
      }
      double accq = 0.0;
#pragma unroll
      for (int u = 0; u < kIt; ++u) {
        const int c = (tid >> 4) + (kRsThreads / 16) * u;
        if (q < 12 && c < G) {
          while ((hi[u] >> 32) != seq3 || (lo[u] >> 32) != seq3) {
            hi[u] = ld_relaxed_u64(ws.mslots + static_cast<size_t>(c) * kMetSlotWords + q * 2);
            lo[u] = ld_relaxed_u64(ws.mslots + static_cast<size_t>(c) * kMetSlotWords + q * 2 + 1);
          }
          accq += __longlong_as_double(static_cast<long long>((hi[u] << 32) | (lo[u] & 0xffffffffull)));
        }
      }
      accq += __shfl_xor_sync(0xffffffffu, accq, 16);
      if (lane < 16) sm_d[lane * kRsWarps + warp] = accq;
      __syncthreads();
      if (warp < 12) {                                            // warp q: quantity q over the 32 warps' partials
        const double tot = warp_sum(sm_d[warp * kRsWarps + lane]);
        if (lane == 0) sm_met[(warp < 4) ? warp : kTileToQ[warp - 4]] = tot;
      }
      __syncthreads();
      if (tid < 32) {
        const bool own = lane < MDE_METRIC_NM;
        const double P = own ? sm_met[lane] : 0.0;
        const double nn = __shfl_sync(0xffffffffu, P, MDE_Q_NVALID);
        const double num = __shfl_sync(0xffffffffu, P, own ? kValNumL[lane] : 0);
        double val = num / nn;
        if (lane >= MDE_M_RMSE_TRUE) val = sqrt(val);
        if (own) {
          const double im = (a.n_img == 1) ? val : __longlong_as_double(0x7ff8000000000000LL);
          a.met_f64[lane] = val;
          a.met_f64[MDE_METRIC_NM + lane] = im;   // per-image means are not formed by the fused path
          a.met_f64[2 * MDE_METRIC_NM + lane] = P;
          a.met_f64[2 * MDE_METRIC_NM + MDE_METRIC_NQ + 1 + lane] = im;   // per-image value sums (one image: the values)
          if (a.met_f32) {
            a.met_f32[lane] = static_cast<float>(val);
            a.met_f32[MDE_METRIC_NM + lane] = static_cast<float>(im);
          }
          if (a.met_accum) a.met_accum[lane] += static_cast<float>(val);   // MetricComputation's running sums
          if (a.met_raw_accum) a.met_raw_accum[lane] += P;
        }
        if (lane == 0) a.met_f64[2 * MDE_METRIC_NM + MDE_METRIC_NQ] = (a.n_img == 1 && nn > 0.0) ? 1.0 : __longlong_as_double(0x7ff8000000000000LL);
      }
    }
  }
}

template <int KIND, unsigned MG, int R>
int launch_resident_r(LossArgs& a, int grid, cudaStream_t st) {
  const void* fn = reinterpret_cast<const void*>(&resident_loss_kernel<KIND, MG, R>);
  void* args[] = {&a};
  MDE_CUDA_TRY(launch_pdl(fn, dim3(static_cast<unsigned>(grid)), dim3(kRsThreads), args, 0, st, true));
  count_launch();
  return MDE_OK;
}

// fp32, 128-bit aligned, default mask, at most 2 quads per thread of one 1024-thread CTA per SM; otherwise `taken`
// stays false and the generic kernel runs
template <int KIND, unsigned MG>
int launch_resident_mg(LossArgs& a, cudaStream_t st, bool& taken) {
  taken = false;
  static const bool off = [] { const char* e = getenv("MDE_NO_RESIDENT"); return e && atoi(e) != 0; }();
  if (off || a.mask != nullptr) return MDE_OK;
  int cap = coop_grid(reinterpret_cast<const void*>(&resident_loss_kernel<KIND, MG, 2>), kRsThreads, 0);
  const int cap1 = coop_grid(reinterpret_cast<const void*>(&resident_loss_kernel<KIND, MG, 1>), kRsThreads, 0);
  if (cap <= 0 || cap1 <= 0) return MDE_OK;
  if (cap1 < cap) cap = cap1;
  const int sms = sm_count();
  if (cap > sms) cap = sms;                 // one CTA per SM
  if (cap > kRsMaxGrid) cap = kRsMaxGrid;
  const int64_t nq = a.n >> 2;
  if (nq < 1) return MDE_OK;
  const int64_t nt = (nq + kRsThreads - 1) / kRsThreads;
  if (nt > static_cast<int64_t>(cap) * 2) return MDE_OK;
  const int grid = static_cast<int>(nt < cap ? nt : cap);
  // SILog with the metric suite and two quads per thread: the shared-memory variant shares the logarithm between loss
  // and metrics and is faster there (8x300x400: 9.4 against 11.6 us)
  if (KIND == MDE_LOSS_SILOG && MG != 0 && nt > grid && a.grad != nullptr) return MDE_OK;
  a.chunk = make_chunking(nq, 8, grid);
  taken = true;
  return (nt <= grid) ? launch_resident_r<KIND, MG, 1>(a, grid, st) : launch_resident_r<KIND, MG, 2>(a, grid, st);
}

template <unsigned MG>
int launch_resident_kind(LossArgs& a, int kind, cudaStream_t st, bool& taken) {
  switch (kind) {
    case MDE_LOSS_L1: return launch_resident_mg<MDE_LOSS_L1, MG>(a, st, taken);
    case MDE_LOSS_MSE: return launch_resident_mg<MDE_LOSS_MSE, MG>(a, st, taken);
    case MDE_LOSS_BERHU: return launch_resident_mg<MDE_LOSS_BERHU, MG>(a, st, taken);
    case MDE_LOSS_LAINA_BERHU: return launch_resident_mg<MDE_LOSS_LAINA_BERHU, MG>(a, st, taken);
    case MDE_LOSS_SILOG: return launch_resident_mg<MDE_LOSS_SILOG, MG>(a, st, taken);
    default: taken = false; return MDE_OK;
  }
}

}  // namespace
}  // namespace mde
