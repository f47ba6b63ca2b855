// losses.cu - masked scalar depth losses, forward and backward fused in ONE cooperative launch.
//
//   MDE_LOSS_L1           MaskedL1Loss      reference criteria.py:80-90
//   MDE_LOSS_MSE          MaskedMSELoss     reference criteria.py:67-77
//   MDE_LOSS_BERHU        berHuLoss         reference criteria.py:111-133
//   MDE_LOSS_LAINA_BERHU  LainaBerHuLoss    reference criteria.py:476-506
//   MDE_LOSS_SILOG        silog_loss        reference criteria.py:724-732
//   (MDE_LOSS_EIGEN lives in eigen.cu)
//
// Kernel shape: persistent grid (<= 148 SMs x 2 CTAs), each CTA owns one contiguous 128-byte
// aligned chunk. Phase A0 (berHu / Laina only) finds the global max, phase A1 reduces the masked
// sums and counts (fp32 per tile, fp64 across tiles, warp shuffle + shared memory per CTA, one
// fp64 atomic per CTA and quantity), a grid-wide barrier publishes the totals, and phase B
// writes dloss/dpred walking the SAME chunk backwards so that the lines read last are re-read
// first (they are still in L2). HBM traffic is 8 B/px read + 4 B/px written when pred+target
// fit in the 126 MB L2.
#include "common.cuh"
#include "metric_math.cuh"

namespace mde {
namespace {

struct LossArgs {
  const void* pred;
  const float* gt;
  const uint8_t* mask;
  int64_t n;
  float vf, clamp_val;
  int use_logs, size_average;
  float grad_scale;
  void* ws;
  float* loss_out;
  double* totals_out;
  void* grad;
};

// ---- chunk iteration ----------------------------------------------------------------------------
// forward: body(idx, p, t) for every element of this CTA's chunk; fold() after every batch of <= 8
// (VEC) / 4 (scalar) elements per thread.
template <typename PT, bool VEC, typename Body, typename Fold>
__device__ __forceinline__ void chunk_forward(const PT* __restrict__ pred, const float* __restrict__ gt,
                                              int64_t n, Body&& body, Fold&& fold) {
  if constexpr (VEC) {
    const int64_t nq = n >> 2;
    int64_t qb, qe;
    cta_chunk(nq, 8, blockIdx.x, gridDim.x, qb, qe);
    int64_t q = qb + threadIdx.x;
    for (; q + kBlock < qe; q += 2 * kBlock) {
      const float4 p0 = Elem<PT>::template ld4<true>(pred + 4 * q);
      const float4 t0 = Elem<float>::template ld4<true>(gt + 4 * q);
      const float4 p1 = Elem<PT>::template ld4<true>(pred + 4 * (q + kBlock));
      const float4 t1 = Elem<float>::template ld4<true>(gt + 4 * (q + kBlock));
      body(4 * q + 0, p0.x, t0.x); body(4 * q + 1, p0.y, t0.y);
      body(4 * q + 2, p0.z, t0.z); body(4 * q + 3, p0.w, t0.w);
      const int64_t q1 = q + kBlock;
      body(4 * q1 + 0, p1.x, t1.x); body(4 * q1 + 1, p1.y, t1.y);
      body(4 * q1 + 2, p1.z, t1.z); body(4 * q1 + 3, p1.w, t1.w);
      fold();
    }
    if (q < qe) {
      const float4 p0 = Elem<PT>::template ld4<true>(pred + 4 * q);
      const float4 t0 = Elem<float>::template ld4<true>(gt + 4 * q);
      body(4 * q + 0, p0.x, t0.x); body(4 * q + 1, p0.y, t0.y);
      body(4 * q + 2, p0.z, t0.z); body(4 * q + 3, p0.w, t0.w);
      fold();
    }
    if (blockIdx.x == gridDim.x - 1) {  // n % 4 tail
      const int64_t i = (nq << 2) + threadIdx.x;
      if (i < n) {
        body(i, Elem<PT>::ld1(pred + i), __ldg(gt + i));
        fold();
      }
    }
  } else {
    int64_t b, e;
    cta_chunk(n, 32, blockIdx.x, gridDim.x, b, e);
    int64_t i = b + threadIdx.x;
    for (; i + 3 * kBlock < e; i += 4 * kBlock) {
      float p[4], t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        p[k] = Elem<PT>::ld1(pred + i + k * kBlock);
        t[k] = __ldg(gt + i + k * kBlock);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) body(i + k * kBlock, p[k], t[k]);
      fold();
    }
    for (; i < e; i += kBlock) {
      body(i, Elem<PT>::ld1(pred + i), __ldg(gt + i));
      fold();
    }
  }
}

// backward walk: out[idx] = body(idx, p, t), last-read lines first.
template <typename PT, bool VEC, typename Body>
__device__ __forceinline__ void chunk_map_reverse(const PT* __restrict__ pred, const float* __restrict__ gt,
                                                  PT* __restrict__ out, int64_t n, Body&& body) {
  if constexpr (VEC) {
    const int64_t nq = n >> 2;
    int64_t qb, qe;
    cta_chunk(nq, 8, blockIdx.x, gridDim.x, qb, qe);
    if (blockIdx.x == gridDim.x - 1) {
      const int64_t i = (nq << 2) + threadIdx.x;
      if (i < n) Elem<PT>::st1(out + i, body(i, Elem<PT>::ld1(pred + i), __ldg(gt + i)));
    }
    int64_t q = qe - 1 - threadIdx.x;
    for (; q - kBlock >= qb; q -= 2 * kBlock) {
      const int64_t q1 = q - kBlock;
      const float4 p0 = Elem<PT>::template ld4<false>(pred + 4 * q);
      const float4 t0 = Elem<float>::template ld4<false>(gt + 4 * q);
      const float4 p1 = Elem<PT>::template ld4<false>(pred + 4 * q1);
      const float4 t1 = Elem<float>::template ld4<false>(gt + 4 * q1);
      float4 g0, g1;
      g0.x = body(4 * q + 0, p0.x, t0.x); g0.y = body(4 * q + 1, p0.y, t0.y);
      g0.z = body(4 * q + 2, p0.z, t0.z); g0.w = body(4 * q + 3, p0.w, t0.w);
      g1.x = body(4 * q1 + 0, p1.x, t1.x); g1.y = body(4 * q1 + 1, p1.y, t1.y);
      g1.z = body(4 * q1 + 2, p1.z, t1.z); g1.w = body(4 * q1 + 3, p1.w, t1.w);
      Elem<PT>::st4(out + 4 * q, g0);
      Elem<PT>::st4(out + 4 * q1, g1);
    }
    if (q >= qb) {
      const float4 p0 = Elem<PT>::template ld4<false>(pred + 4 * q);
      const float4 t0 = Elem<float>::template ld4<false>(gt + 4 * q);
      float4 g0;
      g0.x = body(4 * q + 0, p0.x, t0.x); g0.y = body(4 * q + 1, p0.y, t0.y);
      g0.z = body(4 * q + 2, p0.z, t0.z); g0.w = body(4 * q + 3, p0.w, t0.w);
      Elem<PT>::st4(out + 4 * q, g0);
    }
  } else {
    int64_t b, e;
    cta_chunk(n, 32, blockIdx.x, gridDim.x, b, e);
    for (int64_t i = e - 1 - threadIdx.x; i >= b; i -= kBlock)
      Elem<PT>::st1(out + i, body(i, Elem<PT>::ld1(pred + i), __ldg(gt + i)));
  }
}

// block-reduce up to 4 doubles and add them to gacc[0..N)
template <int N>
__device__ __forceinline__ void publish_sums(const double (&v)[N], double* gacc, double* sm) {
  const double tot = block_sum<N>(v, sm);
  if (threadIdx.x < N && tot != 0.0) atomicAdd(&gacc[threadIdx.x], tot);
}

// block max -> order-preserving atomicMax; NaN anywhere sets the flag word
__device__ __forceinline__ void publish_max(float m, bool saw_nan, unsigned* ukey, float* sm_f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  m = warp_max(m);
  const bool any_nan = __any_sync(0xffffffffu, saw_nan);
  if (lane == 0) {
    sm_f[warp] = m;
    if (any_nan) atomicOr(&ukey[1], 1u);
  }
  __syncthreads();
  if (warp == 0) {
    float x = (lane < kWarps) ? sm_f[lane] : -INFINITY;
    x = warp_max(x);
    if (lane == 0) atomicMax(&ukey[0], float_key(x));
  }
  __syncthreads();
}

__device__ __forceinline__ float read_max(const unsigned* ukey) {
  const unsigned k = __ldcg(&ukey[0]);
  const unsigned f = __ldcg(&ukey[1]);
  if (f) return __int_as_float(0x7fc00000);
  return k == 0u ? -INFINITY : key_float(k);
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// ln(x) for any x: fast path for normal positive x, libdevice for the rest (0, negatives, denormals, inf, NaN)
__device__ __forceinline__ float ln_any(float x) {
  if (x >= 1.17549435e-38f && x < __int_as_float(0x7f800000)) return ln_pos(x);
  return logf(x);
}

// Laina residual n_i = |ln max(p,cv) - ln max(t,cv)| * m   (criteria.py:488-494)
__device__ __forceinline__ float laina_resid(float p, float t, bool m, bool use_logs, float cv, float& r_out) {
  float r;
  if (use_logs) {
    const float pc = (p < cv) ? cv : p;
    const float tc = (t < cv) ? cv : t;
    r = logf(pc) - logf(tc);
  } else {
    r = p - t;
  }
  r_out = r;
  return fabsf(r) * (m ? 1.f : 0.f);
}

template <int KIND, typename PT, bool VEC>
__global__ void __launch_bounds__(kBlock, kCtasPerSm) masked_loss_kernel(LossArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm_d[4 * kWarps];
  __shared__ float sm_f[kWarps];

  const PT* __restrict__ pred = static_cast<const PT*>(a.pred);
  const float* __restrict__ gt = a.gt;
  const uint8_t* __restrict__ mask = a.mask;
  PT* __restrict__ grad = static_cast<PT*>(a.grad);
  const int64_t n = a.n;

  Ws ws = ws_view(a.ws);
  unsigned epoch;
  const int par = coop_prologue(ws, epoch);
  double* gacc = ws.gacc + par * kGacc;
  unsigned* ukey = ws.ukey + par * kUkey;

  // ---------------- phase A0: global max (berHu: max(p - t) over ALL pixels; Laina: max n_i) ----
  float cthr = 0.f, gmax = 0.f;
  if constexpr (KIND == MDE_LOSS_BERHU || KIND == MDE_LOSS_LAINA_BERHU) {
    float mx = -INFINITY;
    bool saw_nan = false;
    chunk_forward<PT, VEC>(
        pred, gt, n,
        [&](int64_t i, float p, float t) {
          float x;
          if constexpr (KIND == MDE_LOSS_BERHU) {
            x = p - t;  // criteria.py:118 - signed, unmasked
          } else {
            const bool m = mask ? (mask[i] != 0) : (t > 0.f);
            float r;
            x = laina_resid(p, t, m, a.use_logs != 0, a.clamp_val, r);
          }
          saw_nan |= (x != x);
          mx = fmaxf(mx, x);
        },
        [] {});
    publish_max(mx, saw_nan, ukey, sm_f);
    grid.sync();
    gmax = read_max(ukey);
    cthr = 0.2f * gmax;  // criteria.py:119 / :496 (fp32 product)
  }

  // ---------------- phase A1: masked sums and counts ------------------------------------------------
  {
    float s0 = 0.f, s1 = 0.f;
    int c0 = 0, c1 = 0;
    double run[4] = {0.0, 0.0, 0.0, 0.0};
    auto fold = [&] {
      run[0] += s0;
      run[1] += s1;
      s0 = 0.f;
      s1 = 0.f;
    };
    chunk_forward<PT, VEC>(
        pred, gt, n,
        [&](int64_t i, float p, float t) {
          if constexpr (KIND == MDE_LOSS_L1) {
            const bool v = t > 0.f;
            s0 += v ? fabsf(t - p) : 0.f;
            c0 += v ? 1 : 0;
          } else if constexpr (KIND == MDE_LOSS_MSE) {
            const bool v = t > 0.f;
            const float d = t - p;
            s0 += v ? d * d : 0.f;
            c0 += v ? 1 : 0;
          } else if constexpr (KIND == MDE_LOSS_SILOG) {
            const bool v = t > 0.01f;  // criteria.py:730
            const float d = v ? (ln_any(p) - ln_any(t)) : 0.f;
            s0 += d;
            s1 = fmaf(d, d, s1);
            c0 += v ? 1 : 0;
          } else if constexpr (KIND == MDE_LOSS_BERHU) {
            const bool v = t > 0.f;
            const float ad = fabsf(t - p);
            const bool hub = v && (ad > cthr);  // criteria.py:126
            s0 += v ? ad : 0.f;
            s1 += hub ? ad * ad : 0.f;
            c0 += v ? 1 : 0;
            c1 += hub ? 1 : 0;
          } else {  // LAINA
            const bool m = mask ? (mask[i] != 0) : (t > 0.f);
            float r;
            const float ni = laina_resid(p, t, m, a.use_logs != 0, a.clamp_val, r);
            const bool big = !(ni < cthr);  // criteria.py:497-498
            const float D = 2.f * cthr + 1e-9f;
            const float num = fmaf(ni, ni, cthr * cthr);
            s0 += big ? num / D : ni;
            s1 += big ? (2.f * cthr * D - 2.f * num) / (D * D) : 0.f;  // d/dc of the quadratic branch
            c0 += m ? 1 : 0;
            c1 += (ni == gmax) ? 1 : 0;
          }
        },
        fold);
    run[2] = static_cast<double>(c0);
    run[3] = static_cast<double>(c1);
    publish_sums<4>(run, gacc, sm_d);
  }
  grid.sync();

  const double S0 = __ldcg(&gacc[0]), S1 = __ldcg(&gacc[1]);
  const double N0 = __ldcg(&gacc[2]), N1 = __ldcg(&gacc[3]);

  // ---------------- loss value and gradient coefficients (every thread, fp64) --------------------
  double loss;
  float k1 = 0.f, k2 = 0.f, k3 = 0.f;
  const double gs = static_cast<double>(a.grad_scale);
  if constexpr (KIND == MDE_LOSS_L1) {
    loss = S0 / N0;
    k1 = static_cast<float>(gs / N0);
  } else if constexpr (KIND == MDE_LOSS_MSE) {
    loss = S0 / N0;
    k1 = static_cast<float>(2.0 * gs / N0);
  } else if constexpr (KIND == MDE_LOSS_SILOG) {
    const double dm = S0 / N0, q = S1 / N0;
    const double s = sqrt(q - static_cast<double>(a.vf) * dm * dm);
    loss = 10.0 * s;
    k1 = static_cast<float>(10.0 * gs / (s * N0));     // dL/dd_i = k1 * (d_i - k2)
    k2 = static_cast<float>(static_cast<double>(a.vf) * dm);
  } else if constexpr (KIND == MDE_LOSS_BERHU) {
    loss = (S0 + S1) / (N0 + N1);                        // mean of the concatenation (criteria.py:131)
    k1 = static_cast<float>(gs / (N0 + N1));
  } else {
    const double Mdiv = a.size_average ? N0 : 1.0;
    loss = S0 / Mdiv;
    k1 = static_cast<float>(gs / Mdiv);
    k2 = static_cast<float>(gs * 0.2 * S1 / (N1 * Mdiv));  // share of dL/dc per tied maximum
    k3 = 2.f * cthr + 1e-9f;
  }

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *a.loss_out = static_cast<float>(loss);
    if (a.totals_out) {
      a.totals_out[0] = S0;
      a.totals_out[1] = S1;
      a.totals_out[2] = N0;
      a.totals_out[3] = N1;
      a.totals_out[4] = static_cast<double>(gmax);
      a.totals_out[5] = loss;
    }
    ws.hdr->epoch = epoch + 1u;
  }
  if (grad == nullptr) return;

  // ---------------- phase B: gradient, chunk walked backwards ----------------------------------------
  chunk_map_reverse<PT, VEC>(pred, gt, grad, n, [&](int64_t i, float p, float t) -> float {
    if constexpr (KIND == MDE_LOSS_L1) {
      const bool v = t > 0.f;
      return v ? -sgn(t - p) * k1 : 0.f;
    } else if constexpr (KIND == MDE_LOSS_MSE) {
      const bool v = t > 0.f;
      return v ? -(t - p) * k1 : 0.f;
    } else if constexpr (KIND == MDE_LOSS_SILOG) {
      const bool v = t > 0.01f;
      const float d = ln_any(p) - ln_any(t);
      return v ? k1 * (d - k2) / p : 0.f;
    } else if constexpr (KIND == MDE_LOSS_BERHU) {
      const bool v = t > 0.f;
      const float d = t - p;
      const float ad = fabsf(d);
      const bool hub = v && (ad > cthr);
      return v ? -sgn(d) * (hub ? fmaf(2.f, ad, 1.f) : 1.f) * k1 : 0.f;
    } else {
      const bool m = mask ? (mask[i] != 0) : (t > 0.f);
      float r;
      const float ni = laina_resid(p, t, m, a.use_logs != 0, a.clamp_val, r);
      const bool big = !(ni < cthr);
      float dn = (big ? 2.f * ni / k3 : 1.f) * k1;
      if (ni == gmax) dn += k2;
      // dn_i/dp = sign(r) * m * [p >= cv] / p   (clamp passes the gradient where p >= cv)
      float dp = m ? sgn(r) : 0.f;
      if (a.use_logs) dp = (p >= a.clamp_val) ? dp / p : 0.f;
      return dn * dp;
    }
  });
}

template <int KIND, typename PT, bool VEC>
int launch_loss(LossArgs& a, cudaStream_t st) {
  const void* fn = reinterpret_cast<const void*>(&masked_loss_kernel<KIND, PT, VEC>);
  const int64_t units = VEC ? (a.n >> 2) : a.n;
  int64_t grid = (units + kBlock - 1) / kBlock;
  const int cap = coop_grid(fn, kBlock, 0);
  if (cap <= 0) return MDE_ECUDA;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  void* args[] = {&a};
  MDE_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(static_cast<unsigned>(grid)), dim3(kBlock), args, 0, st));
  count_launch();
  return MDE_OK;
}

template <int KIND, typename PT>
int launch_loss_vec(LossArgs& a, cudaStream_t st) {
  const bool vec = aligned_to(a.pred, 4 * sizeof(PT)) && aligned_to(a.gt, 16) &&
                   (a.grad == nullptr || aligned_to(a.grad, 4 * sizeof(PT)));
  return vec ? launch_loss<KIND, PT, true>(a, st) : launch_loss<KIND, PT, false>(a, st);
}

template <int KIND>
int launch_loss_dtype(LossArgs& a, int dtype, cudaStream_t st) {
  switch (dtype) {
    case MDE_F32: return launch_loss_vec<KIND, float>(a, st);
    case MDE_F16: return launch_loss_vec<KIND, __half>(a, st);
    case MDE_BF16: return launch_loss_vec<KIND, __nv_bfloat16>(a, st);
    default: set_error("mde_masked_loss: unknown pred_dtype %d", dtype); return MDE_EINVAL;
  }
}

}  // namespace

int eigen_loss_launch(const void* pred, int pred_dtype, const float* target, int64_t n_img, int64_t h, int64_t w,
                      float grad_scale, void* ws, float* loss_out, double* totals_out, void* grad,
                      cudaStream_t st);  // eigen.cu

}  // namespace mde

extern "C" int mde_masked_loss(int kind, const void* pred, int pred_dtype, const float* target,
                               const uint8_t* mask_u8, int64_t n_img, int64_t h, int64_t w,
                               const mde_loss_params* params, float grad_scale, void* ws, float* loss_out,
                               double* totals_out, void* grad, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && ws && loss_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && h > 0 && w > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(aligned_to(target, 4), MDE_EALIGN, "misaligned target");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kind == MDE_LOSS_EIGEN)
    return eigen_loss_launch(pred, pred_dtype, target, n_img, h, w, grad_scale, ws, loss_out, totals_out, grad, st);
  LossArgs a;
  a.pred = pred;
  a.gt = target;
  a.mask = mask_u8;
  a.n = n_img * h * w;
  a.vf = params ? params->variance_focus : 0.85f;
  a.clamp_val = params ? params->clamp_val : 1e-9f;
  a.use_logs = params ? params->use_logs : 1;
  a.size_average = params ? params->size_average : 1;
  a.grad_scale = grad_scale;
  a.ws = ws;
  a.loss_out = loss_out;
  a.totals_out = totals_out;
  a.grad = grad;
  switch (kind) {
    case MDE_LOSS_L1: return launch_loss_dtype<MDE_LOSS_L1>(a, pred_dtype, st);
    case MDE_LOSS_MSE: return launch_loss_dtype<MDE_LOSS_MSE>(a, pred_dtype, st);
    case MDE_LOSS_BERHU: return launch_loss_dtype<MDE_LOSS_BERHU>(a, pred_dtype, st);
    case MDE_LOSS_LAINA_BERHU: return launch_loss_dtype<MDE_LOSS_LAINA_BERHU>(a, pred_dtype, st);
    case MDE_LOSS_SILOG: return launch_loss_dtype<MDE_LOSS_SILOG>(a, pred_dtype, st);
    default: set_error("mde_masked_loss: unknown kind %d", kind); return MDE_EINVAL;
  }
}
