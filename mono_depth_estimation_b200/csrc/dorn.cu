// dorn.cu - DORN ordinal head: pairwise softmax + decode, SID <-> depth, ordinal losses.
//
//   OrdinalRegressionLayer.forward   reference network/Dorn.py:292-321
//   label_to_depth / depth_to_label  reference modules/dorn.py:95-107
//   ordLoss.forward                  reference criteria.py:744-787
//   OrdinalRegressionLoss.__call__   reference criteria.py:789-836
//
// Layout: logits x [n, 2K, hw] with pair k = channels (2k, 2k+1) (Dorn.py:305-306). One thread
// owns one pixel and walks the K pairs in groups (dorn_group: 8 pairs = 16 loads in flight for the
// supervision step, 4 pairs for the plain layer); a warp therefore reads 32 consecutive pixels of one
// channel plane per load (fully coalesced, also when hw is odd and the planes are only 4-byte
// aligned). Everything the step needs - P, decode, depth, the loss term and both logit gradients -
// is produced from that single read of the logits:
// 8K (logits) + 4 (gt) + 8K (grad) + 8 (decode) + 4 (depth) = 1104 B/px at K = 68.
//
// Per pair the kernel issues 42 instructions (60 before the groups lost their per-pair predicates and
// 64-bit stride multiplies): P through MUFU.EX2 + MUFU.RCP (one Newton step), the
// log of the clamped probability through MUFU.LG2 accumulated in the log2 domain (one multiply by
// ln 2 per pixel), predicated selects instead of branches. Accuracy of these forms on B200 is in
// profiles/r01_mathlab_sfu_accuracy.jsonl (relative error ~1e-7, far inside the 1e-5 tolerance).
//
// Bit-exact decode. The reference decides on softmax(clamp(a), clamp(b))[1] > 0.5 evaluated in
// fp32: with d = fl(b' - a') > 0, P = 1/fl(1 + e), e = fl(exp(-d)). P > 0.5 <=> fl(1+e) < 2 <=>
// e <= 1 - 2^-23 <=> exp(-d) <= 1 - 1.5*2^-24 (ties-to-even at the midpoint) <=> d > 1.5*2^-24
// (d = 1.5*2^-24 itself gives exp(-d) just above the midpoint). So decode counts pairs with
// fl(b' - a') > 0x1.8p-24f; equal or non-positive logit pairs (clamped to 1e-8) tie and do not count.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "metric_math.cuh"

namespace mde {
namespace {

constexpr int kDBlock = 256;
constexpr int kDWarps = kDBlock / 32;
// pairs per group = half the loads a thread keeps in flight. Measured on one box at C3 (tools/dorn_probe.py, profiles/
// r02_dorn_ab.jsonl): supervision step 155 us with 4, 151 us with 8; decode only 66.2 us with 4, 67.9 us with 8.
#ifndef MDE_DORN_U_LOSS
#define MDE_DORN_U_LOSS 8
#endif
// resident CTAs of 256 threads per SM the register budget is set for (grid = SMs x this). Measured at C3 on one box
// (profiles/r02_dorn_ab.jsonl, last block): supervision step 151 us with 4, 150 with 5, 187 with 6; decode only 65.5 / 67.2 / 78.8;
// layer forward WITH the probabilities (the module path: 12 K B/px, a third of them stores) 116.5 / 104.3 / 106.0.
constexpr int dorn_ctas(bool has_prob, bool want_loss) { return (has_prob && !want_loss) ? 5 : 4; }
#ifndef MDE_DORN_U_PLAIN
#define MDE_DORN_U_PLAIN 4
#endif
constexpr float kTieMargin = 8.940696716308594e-08f;  // 1.5 * 2^-24, exactly representable

__device__ __forceinline__ float clamp_logit(float v) {
  // torch.clamp(v, 1e-8, 1e4): NaN propagates
  v = (v < 1e-8f) ? 1e-8f : v;
  return (v > 1e4f) ? 1e4f : v;
}

// P = softmax(a', b')[1] = sigmoid(b' - a'), evaluated as softmax does (exp(x - max) / sum) on the SFU:
// e = 2^(-|d| log2 e), P = 1/(1+e) or e/(1+e). Relative error ~2e-7.
__device__ __forceinline__ float pair_prob(float d) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-fabsf(d) * 1.4426950408889634f));
  const float rs = rcp_nr(1.0f + e);
  return (d >= 0.f) ? rs : e * rs;
}

// SID / UD label of a metric depth, op for op as modules/dorn.py:102-107 evaluates it in fp32
__device__ __forceinline__ float depth_label(float t, float alpha, float beta, int K, int disc) {
  if (disc == MDE_DISC_SID) {
    const float num = static_cast<float>(K) * logf(__fdiv_rn(t, alpha));
    return __fdiv_rn(num, logf(__fdiv_rn(beta, alpha)));
  }
  return __fdiv_rn(static_cast<float>(K) * (t - alpha), beta - alpha);
}
// modules/dorn.py:95-100
__device__ __forceinline__ float label_depth(float label, float alpha, float beta, int K, int disc) {
  if (disc == MDE_DISC_SID) {
    const float e = logf(alpha) + __fdiv_rn(logf(__fdiv_rn(beta, alpha)) * label, static_cast<float>(K));
    return expf(e);
  }
  return alpha + __fdiv_rn((beta - alpha) * label, static_cast<float>(K));
}

// block sum of one double -> atomicAdd to *dst (thread 0)
__device__ __forceinline__ void publish_one(double v, double* dst, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kDWarps; ++w) t += sm[w];
    if (t != 0.0) atomicAdd(dst, t);
  }
}

// true (block-uniform) in the last CTA of the grid to get here; all earlier CTAs' atomics are visible
__device__ __forceinline__ bool last_cta(unsigned* ticket) {
  __shared__ bool sm_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sm_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (sm_last) __threadfence();
  return sm_last;
}

// U consecutive pairs of ONE pixel: all 2U logits in flight before the first is used. FULL: every pair exists (no
// predicates, no branches between the pairs - the first version tested k < K per pair and per load, 60 instructions per
// pair; this form has 40); otherwise only the first `npair` < U do.
//   pa / ga / pp: channel 2 k0 of the pixel in the logits / their gradient, plane k0 in the probabilities
//   kf0 = float(k0); y: the pixel's label. LABELLED = false: y is NaN - neither k <= y nor k > y holds (criteria.py:769-770),
//   the pixel has no loss term and a zero gradient (its own instantiation: the common path carries no select for it)
template <typename XT, bool HAS_PROB, bool WANT_LOSS, bool HAS_GRAD, int U, bool FULL, bool LABELLED>
__device__ __forceinline__ void dorn_group(const XT* __restrict__ pa, XT* __restrict__ ga, float* __restrict__ pp, unsigned e,
                                           unsigned ep, unsigned hwu, int npair, float kf0, float y, float inv_nhw, int& cnt,
                                           float& l2sum) {
  float av[U], bv[U];
  // Element of plane c of the pixel: pa[e + c * hwu] with the index formed in 32 bits. Two callers: tensors of fewer than
  // 2^32 elements pass the tensor's base pointers (uniform) and the pixel's element index e, so an access is one 32-bit
  // add (shared by the logit and its gradient) and one IMAD.WIDE; larger tensors pass per-pixel pointers and e = 0
  // (c * hwu < 2^32: hw < 2^27 is checked on the host, c < 32). The first version multiplied 64-bit plane strides.
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (FULL || u < npair) {
      av[u] = Elem<XT>::ld1(pa + (e + static_cast<unsigned>(2 * u) * hwu));
      bv[u] = Elem<XT>::ld1(pa + (e + static_cast<unsigned>(2 * u + 1) * hwu));
    } else {
      av[u] = 0.f;
      bv[u] = 0.f;
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!FULL && u >= npair) break;
    const float ca = clamp_logit(av[u]), cb = clamp_logit(bv[u]);
    const float d = cb - ca;
    const float P = pair_prob(d);
    cnt += (d > kTieMargin) ? 1 : 0;  // == (P > 0.5) of the reference, see header
    if constexpr (HAS_PROB) __stcs(pp + (ep + static_cast<unsigned>(u) * hwu), P);
    if constexpr (WANT_LOSS && LABELLED) {
      const float kf = kf0 + static_cast<float>(u);
      const bool le = kf <= y;
      const float q = 1.0f - P;
      const float xsel = le ? P : q;                     // criteria.py:777-778
      const bool clamped = xsel < 1e-8f;                 // clamp(., 1e-8, 1e8); the upper bound cannot bind
      const float xc = clamped ? 1e-8f : xsel;           // NaN stays NaN
      l2sum += mufu_lg2(xc);
      // dloss/d(b'-a') = -(1-P)/NHW for k <= y, +P/NHW for k > y, zero where the clamp binds
      float gz = le ? -q : P;
      gz = clamped ? 0.f : gz * inv_nhw;
      if constexpr (HAS_GRAD) {
        // the clamp of a logit passes the gradient on [1e-8, 1e4]: exactly where the clamp returned the logit itself
        // (a NaN logit compares unequal and gets 0, as `v >= 1e-8 && v <= 1e4` gave it)
        Elem<XT>::st1(ga + (e + static_cast<unsigned>(2 * u) * hwu), (ca == av[u]) ? -gz : 0.f);
        Elem<XT>::st1(ga + (e + static_cast<unsigned>(2 * u + 1) * hwu), (cb == bv[u]) ? gz : 0.f);
      }
    } else if constexpr (WANT_LOSS && HAS_GRAD) {
      Elem<XT>::st1(ga + (e + static_cast<unsigned>(2 * u) * hwu), 0.f);
      Elem<XT>::st1(ga + (e + static_cast<unsigned>(2 * u + 1) * hwu), 0.f);
    }
  }
}

struct DornArgs {
  const void* x;
  const float* gt;     // metric depth [n,hw] (fused) or SID label [n,hw] (label_mode) or null
  int label_mode;      // gt already holds the float label (ordLoss entry point)
  int index32;         // n 2K hw < 2^32: element indices fit 32 bits (set by the launcher)
  int64_t n, hw;
  int K;
  float alpha, beta;
  int disc;
  float grad_scale;
  void* ws;
  float* loss_out;
  float* prob;
  int64_t* decode;
  float* depth;
  void* grad_x;
};

// One kernel for: layer forward only (gt == null), fused supervision step (gt != null).
// Compile-time switches (which outputs exist) keep the inner loop free of branches.
template <typename XT, bool HAS_PROB, bool WANT_LOSS, bool HAS_GRAD>
__global__ void __launch_bounds__(kDBlock, dorn_ctas(HAS_PROB, WANT_LOSS)) dorn_kernel(DornArgs a) {
  __shared__ double sm[kDWarps];
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  XT* __restrict__ gx = static_cast<XT*>(a.grad_x);
  const int K = a.K;
  const int64_t hw = a.hw;
  const int64_t npx = a.n * hw;
  constexpr bool want_loss = WANT_LOSS;
  const float inv_nhw = a.grad_scale / static_cast<float>(npx);

  double loss_acc = 0.0;
  for (int64_t px = static_cast<int64_t>(blockIdx.x) * kDBlock + threadIdx.x; px < npx;
       px += static_cast<int64_t>(gridDim.x) * kDBlock) {
    const int64_t img = px / hw;
    const int64_t off = px - img * hw;
    const XT* pa = x + img * (2 * static_cast<int64_t>(K)) * hw + off;   // channel 0 of this pixel
    XT* ga = HAS_GRAD ? gx + img * (2 * static_cast<int64_t>(K)) * hw + off : nullptr;
    float* pp = HAS_PROB ? a.prob + img * static_cast<int64_t>(K) * hw + off : nullptr;

    float y = 0.f;
    if constexpr (want_loss) {
      const float t = __ldg(a.gt + px);
      y = a.label_mode ? t : depth_label(t, a.alpha, a.beta, K, a.disc);
    }
    int cnt = 0;
    float l2sum = 0.f;   // sum of log2(clamped probability); times ln 2 at the end
    const unsigned hwu = static_cast<unsigned>(hw);
    constexpr int kDornU = WANT_LOSS ? MDE_DORN_U_LOSS : MDE_DORN_U_PLAIN;
    auto walk = [&](auto labelled, auto index32) {
      constexpr bool L = decltype(labelled)::value;
      constexpr bool I32 = decltype(index32)::value;
      // I32: uniform base pointers + 32-bit element indices; otherwise per-pixel pointers that advance, indices from 0
      const XT* qa = I32 ? x : pa;
      XT* qg = I32 ? gx : ga;
      float* qp = I32 ? a.prob : pp;
      unsigned e = I32 ? static_cast<unsigned>(pa - x) : 0u;
      unsigned ep = (I32 && HAS_PROB) ? static_cast<unsigned>(pp - a.prob) : 0u;
      int k0 = 0;
      for (; k0 + kDornU <= K; k0 += kDornU) {   // full groups: no per-pair predicates, kDornU pairs (2 kDornU loads) in flight
        dorn_group<XT, HAS_PROB, WANT_LOSS, HAS_GRAD, kDornU, true, L>(qa, qg, qp, e, ep, hwu, kDornU, static_cast<float>(k0), y, inv_nhw, cnt, l2sum);
        if constexpr (I32) {
          e += 2u * kDornU * hwu;
          ep += kDornU * hwu;
        } else {
          qa += 2 * kDornU * hw;
          if constexpr (HAS_GRAD) qg += 2 * kDornU * hw;
          if constexpr (HAS_PROB) qp += kDornU * hw;
        }
      }
      if (k0 < K)                                 // the last K % kDornU pairs
        dorn_group<XT, HAS_PROB, WANT_LOSS, HAS_GRAD, kDornU, false, L>(qa, qg, qp, e, ep, hwu, K - k0, static_cast<float>(k0), y, inv_nhw, cnt, l2sum);
    };
    const bool labelled = !want_loss || y == y;   // NaN label (a NaN or negative target under SID): rare, own code path
    if (a.index32) {
      if (labelled) walk(std::true_type{}, std::true_type{});
      else walk(std::false_type{}, std::true_type{});
    } else {
      if (labelled) walk(std::true_type{}, std::false_type{});
      else walk(std::false_type{}, std::false_type{});
    }
    if (a.decode) a.decode[px] = static_cast<int64_t>(cnt);
    if (a.depth) a.depth[px] = label_depth(static_cast<float>(cnt), a.alpha, a.beta, K, a.disc);
    loss_acc += static_cast<double>(l2sum);
  }

  if constexpr (!want_loss) return;
  Ws ws = ws_view(a.ws);
  publish_one(loss_acc, &ws.hdr->tacc[0], sm);
  if (last_cta(&ws.hdr->ticket)) {
    if (threadIdx.x == 0) {
      const double s = __ldcg(&ws.hdr->tacc[0]) * 0.69314718055994531;   // log2 -> ln
      *a.loss_out = static_cast<float>(-s / static_cast<double>(npx));  // criteria.py:784-785
      ws.hdr->tacc[0] = 0.0;
      ws.hdr->ticket = 0u;
    }
  }
}

// ordLoss(P, y): elementwise over [n,K,hw] with the label broadcast over K. Same walk as dorn_group: full groups of U
// planes without per-plane predicates, 32-bit element indices (`wide`: per-pixel pointers, indices from 0), pixels with
// a NaN label on their own code path (no term, zero gradient).
template <bool HAS_GRAD, int U, bool FULL, bool LABELLED>
__device__ __forceinline__ void ord_loss_group(const float* __restrict__ pp, float* __restrict__ gp, unsigned e, unsigned hwu,
                                               int nplane, float kf0, float y, float inv_nhw, float& l2sum) {
  float pv[U];
#pragma unroll
  for (int u = 0; u < U; ++u) pv[u] = (FULL || u < nplane) ? __ldcs(pp + (e + static_cast<unsigned>(u) * hwu)) : 0.5f;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!FULL && u >= nplane) break;
    if constexpr (LABELLED) {
      const float P = pv[u];
      const bool le = (kf0 + static_cast<float>(u)) <= y;
      const float xsel = le ? P : 1.0f - P;
      // clamp(x, 1e-8, 1e8) with NaN kept; the gradient passes inside the closed interval
      const bool lo = xsel < 1e-8f, hi = xsel > 1e8f;
      const float xc = lo ? 1e-8f : (hi ? 1e8f : xsel);
      l2sum += mufu_lg2(xc);
      if constexpr (HAS_GRAD) {
        // d/dP: -1/(P NHW) for k <= y, +1/((1-P) NHW) for k > y
        const float g = (le ? -inv_nhw : inv_nhw) * rcp_nr(xc);
        __stcs(gp + (e + static_cast<unsigned>(u) * hwu), (lo || hi) ? 0.f : g);
      }
    } else if constexpr (HAS_GRAD) {
      __stcs(gp + (e + static_cast<unsigned>(u) * hwu), 0.f);
    }
  }
}

#ifndef MDE_ORD_U
#define MDE_ORD_U 16   // planes per group (C3, one box: 89.9 us with 8, 86.5 with 16; 6 CTAs per SM: 87.4 / 104.6)
#endif
#ifndef MDE_ORD_CTAS
#define MDE_ORD_CTAS 4
#endif
template <bool HAS_GRAD>
__global__ void __launch_bounds__(kDBlock, MDE_ORD_CTAS)
ord_loss_kernel(const float* __restrict__ prob, const float* __restrict__ label, int64_t n, int K, int64_t hw,
                float grad_scale, void* ws_raw, float* loss_out, float* __restrict__ grad, int index32) {
  __shared__ double sm[kDWarps];
  constexpr int U = MDE_ORD_U;
  const int64_t npx = n * hw;
  const float inv_nhw = grad_scale / static_cast<float>(npx);
  const unsigned hwu = static_cast<unsigned>(hw);
  double loss_acc = 0.0;
  for (int64_t px = static_cast<int64_t>(blockIdx.x) * kDBlock + threadIdx.x; px < npx;
       px += static_cast<int64_t>(gridDim.x) * kDBlock) {
    const int64_t img = px / hw;
    const int64_t first = img * static_cast<int64_t>(K) * hw + (px - img * hw);   // plane 0 of this pixel
    const float y = __ldg(label + px);
    float l2sum = 0.f;
    auto walk = [&](auto labelled, auto index32_t) {
      constexpr bool L = decltype(labelled)::value;
      constexpr bool I32 = decltype(index32_t)::value;
      const float* qp = I32 ? prob : prob + first;
      float* qg = HAS_GRAD ? (I32 ? grad : grad + first) : nullptr;
      unsigned e = I32 ? static_cast<unsigned>(first) : 0u;
      int k0 = 0;
      for (; k0 + U <= K; k0 += U) {
        ord_loss_group<HAS_GRAD, U, true, L>(qp, qg, e, hwu, U, static_cast<float>(k0), y, inv_nhw, l2sum);
        if constexpr (I32) {
          e += U * hwu;
        } else {
          qp += U * hw;
          if constexpr (HAS_GRAD) qg += U * hw;
        }
      }
      if (k0 < K) ord_loss_group<HAS_GRAD, U, false, L>(qp, qg, e, hwu, K - k0, static_cast<float>(k0), y, inv_nhw, l2sum);
    };
    const bool labelled = (y == y);
    if (index32) {
      if (labelled) walk(std::true_type{}, std::true_type{});
      else walk(std::false_type{}, std::true_type{});
    } else {
      if (labelled) walk(std::true_type{}, std::false_type{});
      else walk(std::false_type{}, std::false_type{});
    }
    loss_acc += static_cast<double>(l2sum);
  }
  Ws ws = ws_view(ws_raw);
  publish_one(loss_acc, &ws.hdr->tacc[0], sm);
  if (last_cta(&ws.hdr->ticket)) {
    if (threadIdx.x == 0) {
      const double s = __ldcg(&ws.hdr->tacc[0]) * 0.69314718055994531;
      *loss_out = static_cast<float>(-s / static_cast<double>(npx));
      ws.hdr->tacc[0] = 0.0;
      ws.hdr->ticket = 0u;
    }
  }
}

// backward of OrdinalRegressionLayer: grad_x from grad_P. Groups of U pairs (3 U loads in flight), 32-bit element
// indices where the logits have fewer than 2^32 elements.
template <typename XT, int U, bool FULL>
__device__ __forceinline__ void layer_bwd_group(const XT* __restrict__ px_x, XT* __restrict__ px_g, const float* __restrict__ px_p,
                                                unsigned e, unsigned ep, unsigned hwu, int npair) {
  float av[U], bv[U], gv[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (FULL || u < npair) {
      av[u] = Elem<XT>::ld1(px_x + (e + static_cast<unsigned>(2 * u) * hwu));
      bv[u] = Elem<XT>::ld1(px_x + (e + static_cast<unsigned>(2 * u + 1) * hwu));
      gv[u] = __ldcs(px_p + (ep + static_cast<unsigned>(u) * hwu));
    } else {
      av[u] = bv[u] = gv[u] = 0.f;
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!FULL && u >= npair) break;
    const float ca = clamp_logit(av[u]), cb = clamp_logit(bv[u]);
    const float P = pair_prob(cb - ca);
    const float gz = gv[u] * P * (1.0f - P);  // softmax backward for the 2-way case
    // the clamp passes the gradient exactly where it returned the logit itself ([1e-8, 1e4]; NaN compares unequal)
    Elem<XT>::st1(px_g + (e + static_cast<unsigned>(2 * u) * hwu), (ca == av[u]) ? -gz : 0.f);
    Elem<XT>::st1(px_g + (e + static_cast<unsigned>(2 * u + 1) * hwu), (cb == bv[u]) ? gz : 0.f);
  }
}

template <typename XT>
__global__ void __launch_bounds__(kDBlock, 4)
ordinal_layer_bwd_kernel(const XT* __restrict__ x, const float* __restrict__ gp, int64_t n, int K, int64_t hw,
                         XT* __restrict__ gx, int index32) {
  constexpr int U = 4;
  const int64_t npx = n * hw;
  const unsigned hwu = static_cast<unsigned>(hw);
  for (int64_t px = static_cast<int64_t>(blockIdx.x) * kDBlock + threadIdx.x; px < npx;
       px += static_cast<int64_t>(gridDim.x) * kDBlock) {
    const int64_t img = px / hw;
    const int64_t off = px - img * hw;
    const int64_t base = img * (2 * static_cast<int64_t>(K)) * hw + off;
    const int64_t pbase = img * static_cast<int64_t>(K) * hw + off;
    auto walk = [&](auto index32_t) {
      constexpr bool I32 = decltype(index32_t)::value;
      const XT* qx = I32 ? x : x + base;
      XT* qg = I32 ? gx : gx + base;
      const float* qp = I32 ? gp : gp + pbase;
      unsigned e = I32 ? static_cast<unsigned>(base) : 0u, ep = I32 ? static_cast<unsigned>(pbase) : 0u;
      int k0 = 0;
      for (; k0 + U <= K; k0 += U) {
        layer_bwd_group<XT, U, true>(qx, qg, qp, e, ep, hwu, U);
        if constexpr (I32) {
          e += 2u * U * hwu;
          ep += U * hwu;
        } else {
          qx += 2 * U * hw;
          qg += 2 * U * hw;
          qp += U * hw;
        }
      }
      if (k0 < K) layer_bwd_group<XT, U, false>(qx, qg, qp, e, ep, hwu, K - k0);
    };
    if (index32) walk(std::true_type{});
    else walk(std::false_type{});
  }
}

// OrdinalRegressionLoss: pass 1 counts valid pixels into tacc[8]
__global__ void __launch_bounds__(kDBlock, 4) orl_count_kernel(const float* __restrict__ gt, int64_t npx, void* ws_raw) {
  __shared__ double sm[kDWarps];
  double c = 0.0;
  for (int64_t px = static_cast<int64_t>(blockIdx.x) * kDBlock + threadIdx.x; px < npx;
       px += static_cast<int64_t>(gridDim.x) * kDBlock)
    c += (__ldg(gt + px) > 0.f) ? 1.0 : 0.0;
  Ws ws = ws_view(ws_raw);
  publish_one(c, &ws.hdr->tacc[8], sm);
}

__global__ void __launch_bounds__(kDBlock, 4)
orl_kernel(const float* __restrict__ prob, const float* __restrict__ gt, int64_t n, int K, int64_t hw, float alpha,
           float beta, int disc, float grad_scale, void* ws_raw, float* loss_out, float* __restrict__ grad, int index32) {
  __shared__ double sm[kDWarps];
  Ws ws = ws_view(ws_raw);
  const unsigned hwu = static_cast<unsigned>(hw);
  const double n_valid = __ldcg(&ws.hdr->tacc[8]);
  const float gcoef = -grad_scale / static_cast<float>(n_valid);
  const int64_t npx = n * hw;
  double loss_acc = 0.0;
  for (int64_t px = static_cast<int64_t>(blockIdx.x) * kDBlock + threadIdx.x; px < npx;
       px += static_cast<int64_t>(gridDim.x) * kDBlock) {
    const int64_t img = px / hw;
    const int64_t base = img * (2 * static_cast<int64_t>(K)) * hw + (px - img * hw);
    const float t = __ldg(gt + px);
    const bool valid = t > 0.f;  // criteria.py:829
    long long label = 0;
    if (valid) label = __float2ll_rz(depth_label(t, alpha, beta, K, disc));  // .long(): toward zero (:805)
    float lsum = 0.f;
    // planes 0..K-1 hold the '<=' probabilities, K..2K-1 the '>' ones: pixel (k <= label) reads plane k, otherwise
    // plane K + k. Groups of 8 selected planes in flight, 32-bit element indices (index32) or offsets from the pixel.
    const bool i32 = index32 != 0;
    const float* qp = i32 ? prob : prob + base;
    float* qg = grad ? (i32 ? grad : grad + base) : nullptr;
    const unsigned e0 = i32 ? static_cast<unsigned>(base) : 0u;
    const unsigned Khw = static_cast<unsigned>(K) * hwu;
    // ord_c0 = 1 where not (k > label) (:809-810): k <= label, clamped to the walk's range
    const int n_le = (label < 0) ? 0 : ((label >= static_cast<long long>(K)) ? K : static_cast<int>(label) + 1);
    for (int k0 = 0; k0 < K; k0 += 8) {
      float pv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u;
        const unsigned ek = e0 + static_cast<unsigned>(k) * hwu;
        pv[u] = (valid && k < K) ? __ldcs(qp + (k < n_le ? ek : ek + Khw)) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u;
        if (k >= K) break;
        const bool le = k < n_le;
        if (valid) lsum += pv[u];
        if (qg) {
          const unsigned ek = e0 + static_cast<unsigned>(k) * hwu;
          __stcs(qg + ek, (valid && le) ? gcoef : 0.f);
          __stcs(qg + (ek + Khw), (valid && !le) ? gcoef : 0.f);
        }
      }
    }
    loss_acc -= static_cast<double>(lsum);
  }
  publish_one(loss_acc, &ws.hdr->tacc[0], sm);
  if (last_cta(&ws.hdr->ticket)) {
    if (threadIdx.x == 0) {
      const double s = __ldcg(&ws.hdr->tacc[0]);
      *loss_out = static_cast<float>(s / n_valid);
      ws.hdr->tacc[0] = 0.0;
      ws.hdr->tacc[8] = 0.0;
      ws.hdr->ticket = 0u;
    }
  }
}

template <typename LT>
__global__ void __launch_bounds__(kDBlock) label_to_depth_kernel(const LT* __restrict__ label, int64_t n, float alpha,
                                                                 float beta, int K, int disc, float* __restrict__ depth) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kDBlock + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * kDBlock)
    depth[i] = label_depth(static_cast<float>(label[i]), alpha, beta, K, disc);
}

__global__ void __launch_bounds__(kDBlock) depth_to_label_kernel(const float* __restrict__ depth, int64_t n, float alpha,
                                                                 float beta, int K, int disc, float* __restrict__ label) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kDBlock + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * kDBlock)
    label[i] = depth_label(__ldg(depth + i), alpha, beta, K, disc);
}

// element indices of a tensor fit 32 bits; MDE_DORN_NO_INDEX32=1 sends small tensors down the 64-bit pointer path too
// (a >= 16 GB tensor is not a test case) - read per call so that a test can switch it
inline int index32_ok(int64_t n_elem) {
  if (const char* e = getenv("MDE_DORN_NO_INDEX32")) {
    if (atoi(e) != 0) return 0;
  }
  return n_elem < (int64_t(1) << 32) ? 1 : 0;
}

inline unsigned px_grid(int64_t npx, int ctas_per_sm) {
  int64_t g = (npx + kDBlock - 1) / kDBlock;
  const int64_t cap = static_cast<int64_t>(sm_count()) * ctas_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

template <typename XT>
int launch_dorn_t(DornArgs& a, cudaStream_t st) {
  const bool p = a.prob != nullptr, l = a.gt != nullptr, g = a.grad_x != nullptr && l;
  if (a.hw >= (int64_t(1) << 27)) {
    set_error("DORN head: more than 2^27 pixels per image");
    return MDE_ETOOBIG;
  }
  a.index32 = index32_ok(a.n * 2 * static_cast<int64_t>(a.K) * a.hw);
#define MDE_DORN(P, L, G) dorn_kernel<XT, P, L, G><<<px_grid(a.n * a.hw, dorn_ctas(P, L)), kDBlock, 0, st>>>(a)
  if (!l) { if (p) MDE_DORN(true, false, false); else MDE_DORN(false, false, false); }
  else if (g) { if (p) MDE_DORN(true, true, true); else MDE_DORN(false, true, true); }
  else { if (p) MDE_DORN(true, true, false); else MDE_DORN(false, true, false); }
#undef MDE_DORN
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int launch_dorn(DornArgs& a, int x_dtype, cudaStream_t st) {
  switch (x_dtype) {
    case MDE_F32: return launch_dorn_t<float>(a, st);
    case MDE_F16: return launch_dorn_t<__half>(a, st);
    case MDE_BF16: return launch_dorn_t<__nv_bfloat16>(a, st);
    default: set_error("dorn: unknown x_dtype %d", x_dtype); return MDE_EINVAL;
  }
}

}  // namespace
}  // namespace mde

using namespace mde;

extern "C" int mde_ordinal_layer_fwd(const void* x, int x_dtype, int64_t n, int64_t K, int64_t hw, float* prob,
                                     int64_t* decode, void* stream) {
  MDE_REQUIRE(x != nullptr, MDE_EINVAL, "null logits");
  MDE_REQUIRE(n > 0 && K > 0 && hw > 0 && K < 32768, MDE_EINVAL, "bad shape");
  DornArgs a{};
  a.x = x; a.gt = nullptr; a.label_mode = 0; a.n = n; a.hw = hw; a.K = static_cast<int>(K);
  a.alpha = 1.f; a.beta = 2.f; a.disc = MDE_DISC_UD; a.grad_scale = 1.f;
  a.ws = nullptr; a.loss_out = nullptr; a.prob = prob; a.decode = decode; a.depth = nullptr; a.grad_x = nullptr;
  return launch_dorn(a, x_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int mde_ordinal_layer_bwd(const void* x, int x_dtype, const float* grad_prob, int64_t n, int64_t K,
                                     int64_t hw, void* grad_x, void* stream) {
  MDE_REQUIRE(x && grad_prob && grad_x, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n > 0 && K > 0 && hw > 0 && K < 32768, MDE_EINVAL, "bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = px_grid(n * hw, 4);
  const int k = static_cast<int>(K);
  MDE_REQUIRE(hw < (int64_t(1) << 27), MDE_ETOOBIG, "more than 2^27 pixels per image");
  const int i32 = index32_ok(n * 2 * K * hw);
  switch (x_dtype) {
    case MDE_F32: ordinal_layer_bwd_kernel<float><<<grid, kDBlock, 0, st>>>(static_cast<const float*>(x), grad_prob, n, k, hw, static_cast<float*>(grad_x), i32); break;
    case MDE_F16: ordinal_layer_bwd_kernel<__half><<<grid, kDBlock, 0, st>>>(static_cast<const __half*>(x), grad_prob, n, k, hw, static_cast<__half*>(grad_x), i32); break;
    case MDE_BF16: ordinal_layer_bwd_kernel<__nv_bfloat16><<<grid, kDBlock, 0, st>>>(static_cast<const __nv_bfloat16*>(x), grad_prob, n, k, hw, static_cast<__nv_bfloat16*>(grad_x), i32); break;
    default: set_error("mde_ordinal_layer_bwd: unknown x_dtype %d", x_dtype); return MDE_EINVAL;
  }
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

extern "C" int mde_label_to_depth_i64(const int64_t* label, int64_t n, float alpha, float beta, int ord_num,
                                      int discretization, float* depth, void* stream) {
  MDE_REQUIRE(label && depth, MDE_EINVAL, "null pointer");
  if (n <= 0) return MDE_OK;
  label_to_depth_kernel<int64_t><<<px_grid(n, 8), kDBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      label, n, alpha, beta, ord_num, discretization, depth);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

extern "C" int mde_label_to_depth_f32(const float* label, int64_t n, float alpha, float beta, int ord_num,
                                      int discretization, float* depth, void* stream) {
  MDE_REQUIRE(label && depth, MDE_EINVAL, "null pointer");
  if (n <= 0) return MDE_OK;
  label_to_depth_kernel<float><<<px_grid(n, 8), kDBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      label, n, alpha, beta, ord_num, discretization, depth);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

extern "C" int mde_depth_to_label(const float* depth, int64_t n, float alpha, float beta, int ord_num,
                                  int discretization, float* label, void* stream) {
  MDE_REQUIRE(depth && label, MDE_EINVAL, "null pointer");
  if (n <= 0) return MDE_OK;
  depth_to_label_kernel<<<px_grid(n, 8), kDBlock, 0, static_cast<cudaStream_t>(stream)>>>(depth, n, alpha, beta, ord_num,
                                                                                          discretization, label);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

extern "C" int mde_ord_loss(const float* prob, const float* target_label, int64_t n, int64_t K, int64_t hw,
                            float grad_scale, void* ws, float* loss_out, float* grad_prob, void* stream) {
  MDE_REQUIRE(prob && target_label && ws && loss_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n > 0 && K > 0 && hw > 0 && K < 32768, MDE_EINVAL, "bad shape");
  MDE_REQUIRE(hw < (int64_t(1) << 27), MDE_ETOOBIG, "more than 2^27 pixels per image");
  const int i32 = index32_ok(n * K * hw);
  if (grad_prob)
    ord_loss_kernel<true><<<px_grid(n * hw, MDE_ORD_CTAS), kDBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        prob, target_label, n, static_cast<int>(K), hw, grad_scale, ws, loss_out, grad_prob, i32);
  else
    ord_loss_kernel<false><<<px_grid(n * hw, MDE_ORD_CTAS), kDBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        prob, target_label, n, static_cast<int>(K), hw, grad_scale, ws, loss_out, nullptr, i32);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

extern "C" int mde_dorn_fused(const void* x, int x_dtype, const float* gt_depth, int64_t n, int64_t K, int64_t hw,
                              float alpha, float beta, int discretization, float grad_scale, void* ws,
                              float* loss_out, float* prob, int64_t* decode, float* depth, void* grad_x,
                              void* stream) {
  MDE_REQUIRE(x && gt_depth && ws && loss_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n > 0 && K > 0 && hw > 0 && K < 32768, MDE_EINVAL, "bad shape");
  DornArgs a{};
  a.x = x; a.gt = gt_depth; a.label_mode = 0; a.n = n; a.hw = hw; a.K = static_cast<int>(K);
  a.alpha = alpha; a.beta = beta; a.disc = discretization; a.grad_scale = grad_scale;
  a.ws = ws; a.loss_out = loss_out; a.prob = prob; a.decode = decode; a.depth = depth; a.grad_x = grad_x;
  return launch_dorn(a, x_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int mde_ordinal_regression_loss(const float* prob, const float* gt_depth, int64_t n, int64_t K, int64_t hw,
                                           float alpha, float beta, int discretization, float grad_scale, void* ws,
                                           float* loss_out, float* grad_prob, void* stream) {
  MDE_REQUIRE(prob && gt_depth && ws && loss_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n > 0 && K > 0 && hw > 0 && K < 32768, MDE_EINVAL, "bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  orl_count_kernel<<<px_grid(n * hw, 4), kDBlock, 0, st>>>(gt_depth, n * hw, ws);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  MDE_REQUIRE(2 * K * hw < (int64_t(1) << 32), MDE_ETOOBIG, "more than 2^32 probabilities per image");
  orl_kernel<<<px_grid(n * hw, 4), kDBlock, 0, st>>>(prob, gt_depth, n, static_cast<int>(K), hw, alpha, beta,
                                                     discretization, grad_scale, ws, loss_out, grad_prob,
                                                     index32_ok(n * 2 * K * hw));
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
