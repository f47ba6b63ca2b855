// losses_split.cu - the masked losses in SPLIT-PHASE form for global-batch training over several GPUs
// (SURVEY 8e, row "Loss, global-batch mode"; north_star: "loss batches shard by image ... NCCL used only for the
// tiny allreduce of partial sums and counts").
//
// A loss of the family (MaskedL1Loss criteria.py:80-90, MaskedMSELoss :67-77, berHuLoss :111-133,
// LainaBerHuLoss :476-506, silog_loss :724-732) is a map over pixels followed by a handful of scalar totals; the
// gradient of every pixel depends on the input only through those totals. With the batch sharded by image:
//   stage 0 (berHu / Laina only)  partials <- max over the shard                 -> all-reduce(MAX) of 1 double
//   stage 1                      partials <- {S0, S1, N0, N1} given the global max -> all-reduce(SUM) of 4 doubles
//   finish                       loss (identical on every rank) and dloss/dpred of the shard from the global totals
// so that N GPUs produce exactly the loss and the gradient of the single-GPU full-batch call. These are plain
// streaming kernels (no co-residency requirement, nothing waits on another rank): the exchange happens between
// launches, on the caller's communicator. Per-pixel arithmetic = losses_kernel.cuh (shared helpers).
#include "losses_kernel.cuh"

namespace mde {
namespace {

constexpr int kSpBlock = 256;

__device__ __forceinline__ void atomic_max_nan(double* addr, double v) {
  // max that propagates NaN (torch.max does): NaN wins and stays
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  for (;;) {
    const double o = __longlong_as_double(static_cast<long long>(old));
    if (o != o) return;
    if (v == v && o >= v) return;
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
    if (old == assumed) return;
  }
}

struct SplitArgs {
  const void* pred;
  const float* gt;
  const uint8_t* mask;
  int64_t n;
  float vf, clamp_val;
  int use_logs, size_average;
  const double* gmax_in;   // device, stage 1 / finish of berHu and Laina: the GLOBAL max
  double* partials;        // device [8]: {S0, S1, N0, N1, max, -, -, -}; stage kernels ADD (max: MAX) into it
  const double* totals;    // device [8]: finish
  float grad_scale;
  float* loss_out;
  void* grad;
};

// stage 0: max(p - t) over all pixels (berHu, criteria.py:118-119) or max n_i (Laina, :495-496)
template <int KIND, typename PT>
__global__ void __launch_bounds__(kSpBlock) split_max_kernel(SplitArgs a) {
  __shared__ float sm_m[kSpBlock / 32];
  __shared__ int sm_nan;
  const PT* __restrict__ pred = static_cast<const PT*>(a.pred);
  float mx = -INFINITY;
  bool nan = false;
  if (threadIdx.x == 0) sm_nan = 0;
  __syncthreads();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kSpBlock + threadIdx.x; i < a.n; i += static_cast<int64_t>(gridDim.x) * kSpBlock) {
    const float p = Elem<PT>::ld1(pred + i), t = __ldg(a.gt + i);
    float x;
    if constexpr (KIND == MDE_LOSS_BERHU) {
      x = p - t;
    } else {
      const bool m = a.mask ? (a.mask[i] != 0) : (t > 0.f);
      float r;
      x = laina_resid(p, t, m, a.use_logs != 0, a.clamp_val, r);
    }
    nan |= (x != x);
    mx = fmaxf(mx, x);
  }
  mx = warp_max(mx);
  if (__any_sync(0xffffffffu, nan) && (threadIdx.x & 31) == 0) atomicOr(&sm_nan, 1);
  if ((threadIdx.x & 31) == 0) sm_m[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = sm_m[0];
    for (int w = 1; w < kSpBlock / 32; ++w) m = fmaxf(m, sm_m[w]);
    atomic_max_nan(a.partials + 4, sm_nan ? static_cast<double>(__int_as_float(0x7fc00000)) : static_cast<double>(m));
  }
}

// per-pixel contribution to the totals, identical to phase A1 of masked_loss_kernel
template <int KIND>
__device__ __forceinline__ void split_px(const SplitArgs& a, int64_t i, float p, float t, float cthr, float gmax, float& s0, float& s1,
                                         int& c0, int& c1) {
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
    bool v;
    const float d = silog_resid(p, t, v);
    s0 += d;
    s1 = fmaf(d, d, s1);
    c0 += v ? 1 : 0;
  } else if constexpr (KIND == MDE_LOSS_BERHU) {
    const bool v = t > 0.f;
    const float ad = fabsf(t - p);
    const bool hub = v && (ad > cthr);
    s0 += v ? ad : 0.f;
    s1 += hub ? ad * ad : 0.f;
    c0 += v ? 1 : 0;
    c1 += hub ? 1 : 0;
  } else {
    const bool m = a.mask ? (a.mask[i] != 0) : (t > 0.f);
    float r;
    const float ni = laina_resid(p, t, m, a.use_logs != 0, a.clamp_val, r);
    const bool big = !(ni < cthr);
    const float D = 2.f * cthr + 1e-9f;
    const float num = fmaf(ni, ni, cthr * cthr);
    s0 += big ? num / D : ni;
    s1 += big ? (2.f * cthr * D - 2.f * num) / (D * D) : 0.f;
    c0 += m ? 1 : 0;
    c1 += (ni == gmax) ? 1 : 0;
  }
}

// stage 1: {S0, S1, N0, N1} of the shard (fp32 over <= 64 pixels per thread and fold, fp64 beyond)
template <int KIND, typename PT>
__global__ void __launch_bounds__(kSpBlock) split_sums_kernel(SplitArgs a) {
  __shared__ double sm[4 * (kSpBlock / 32)];
  const PT* __restrict__ pred = static_cast<const PT*>(a.pred);
  float gmax = 0.f, cthr = 0.f;
  if constexpr (KIND == MDE_LOSS_BERHU || KIND == MDE_LOSS_LAINA_BERHU) {
    gmax = static_cast<float>(__ldg(a.gmax_in));
    cthr = 0.2f * gmax;
  }
  double run[4] = {0.0, 0.0, 0.0, 0.0};
  float s0 = 0.f, s1 = 0.f;
  int c0 = 0, c1 = 0, it = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kSpBlock + threadIdx.x; i < a.n; i += static_cast<int64_t>(gridDim.x) * kSpBlock) {
    split_px<KIND>(a, i, Elem<PT>::ld1(pred + i), __ldg(a.gt + i), cthr, gmax, s0, s1, c0, c1);
    if ((++it & 63) == 0) {
      run[0] += s0; run[1] += s1; s0 = 0.f; s1 = 0.f;
    }
  }
  run[0] += s0; run[1] += s1; run[2] = c0; run[3] = c1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double s = warp_sum(run[q]);
    if (lane == 0) sm[q * (kSpBlock / 32) + warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double tot = 0.0;
    for (int w = 0; w < kSpBlock / 32; ++w) tot += sm[threadIdx.x * (kSpBlock / 32) + w];
    if (tot != 0.0) atomicAdd(a.partials + threadIdx.x, tot);
  }
}

// finish: loss from the GLOBAL totals (every rank computes the same value) and the gradient of this shard,
// identical to the coefficient step and phase B of masked_loss_kernel
template <int KIND, typename PT>
__global__ void __launch_bounds__(kSpBlock) split_grad_kernel(SplitArgs a) {
  const PT* __restrict__ pred = static_cast<const PT*>(a.pred);
  PT* grad = static_cast<PT*>(a.grad);
  const double S0 = __ldg(a.totals + 0), S1 = __ldg(a.totals + 1), N0 = __ldg(a.totals + 2), N1 = __ldg(a.totals + 3);
  const float gmax = static_cast<float>(__ldg(a.totals + 4));
  const float cthr = 0.2f * gmax;
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
    const double dm = S0 * inv, q = S1 * inv;
    const double var = q - static_cast<double>(a.vf) * dm * dm;
    const float s = sqrtf(static_cast<float>(var));
    loss = 10.0 * static_cast<double>(s);
    k1 = 10.0f * gs * static_cast<float>(inv) / s;
    k2 = a.vf * static_cast<float>(dm);
  } else if constexpr (KIND == MDE_LOSS_BERHU) {
    const double inv = 1.0 / (N0 + N1);
    loss = (S0 + S1) * inv;
    k1 = gs * static_cast<float>(inv);
  } else {
    const double inv = a.size_average ? 1.0 / N0 : 1.0;
    loss = S0 * inv;
    k1 = gs * static_cast<float>(inv);
    k2 = gs * 0.2f * static_cast<float>(S1 * inv) / static_cast<float>(N1);
    k3 = 2.f * cthr + 1e-9f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.loss_out = static_cast<float>(loss);
  if (grad == nullptr) return;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kSpBlock + threadIdx.x; i < a.n; i += static_cast<int64_t>(gridDim.x) * kSpBlock) {
    const float p = Elem<PT>::ld1(pred + i), t = __ldg(a.gt + i);
    float g;
    if constexpr (KIND == MDE_LOSS_L1) {
      g = (t > 0.f) ? -sgn(t - p) * k1 : 0.f;
    } else if constexpr (KIND == MDE_LOSS_MSE) {
      g = (t > 0.f) ? -(t - p) * k1 : 0.f;
    } else if constexpr (KIND == MDE_LOSS_SILOG) {
      bool v;
      const float d = silog_resid(p, t, v);
      g = v ? k1 * (d - k2) * rcp_nr(p) : 0.f;
    } else if constexpr (KIND == MDE_LOSS_BERHU) {
      const bool v = t > 0.f;
      const float d = t - p, ad = fabsf(d);
      const bool hub = v && (ad > cthr);
      g = v ? -sgn(d) * (hub ? fmaf(2.f, ad, 1.f) : 1.f) * k1 : 0.f;
    } else {
      const bool m = a.mask ? (a.mask[i] != 0) : (t > 0.f);
      float r;
      const float ni = laina_resid(p, t, m, a.use_logs != 0, a.clamp_val, r);
      const bool big = !(ni < cthr);
      float dn = (big ? 2.f * ni / k3 : 1.f) * k1;
      if (ni == gmax) dn += k2;
      float dp = m ? sgn(r) : 0.f;
      if (a.use_logs) dp = (p >= a.clamp_val) ? dp / p : 0.f;
      g = dn * dp;
    }
    Elem<PT>::st1(grad + i, g);
  }
}

template <int KIND, typename PT>
int split_launch(int stage, SplitArgs& a, cudaStream_t st) {
  int64_t grid = (a.n + kSpBlock - 1) / kSpBlock;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const unsigned g = static_cast<unsigned>(grid);
  if (stage == 0) {
    if constexpr (KIND == MDE_LOSS_BERHU || KIND == MDE_LOSS_LAINA_BERHU) split_max_kernel<KIND, PT><<<g, kSpBlock, 0, st>>>(a);
    else return MDE_OK;   // no max stage for this loss
  } else if (stage == 1) {
    split_sums_kernel<KIND, PT><<<g, kSpBlock, 0, st>>>(a);
  } else {
    split_grad_kernel<KIND, PT><<<g, kSpBlock, 0, st>>>(a);
  }
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

template <int KIND>
int split_dtype(int stage, SplitArgs& a, int dtype, cudaStream_t st) {
  switch (dtype) {
    case MDE_F32: return split_launch<KIND, float>(stage, a, st);
    case MDE_F16: return split_launch<KIND, __half>(stage, a, st);
    case MDE_BF16: return split_launch<KIND, __nv_bfloat16>(stage, a, st);
    default: set_error("mde_masked_loss_split: unknown pred_dtype %d", dtype); return MDE_EINVAL;
  }
}

int split_kind(int kind, int stage, SplitArgs& a, int dtype, cudaStream_t st) {
  switch (kind) {
    case MDE_LOSS_L1: return split_dtype<MDE_LOSS_L1>(stage, a, dtype, st);
    case MDE_LOSS_MSE: return split_dtype<MDE_LOSS_MSE>(stage, a, dtype, st);
    case MDE_LOSS_BERHU: return split_dtype<MDE_LOSS_BERHU>(stage, a, dtype, st);
    case MDE_LOSS_LAINA_BERHU: return split_dtype<MDE_LOSS_LAINA_BERHU>(stage, a, dtype, st);
    case MDE_LOSS_SILOG: return split_dtype<MDE_LOSS_SILOG>(stage, a, dtype, st);
    default: set_error("mde_masked_loss_split: kind %d has no split-phase form", kind); return MDE_EINVAL;
  }
}

SplitArgs split_args(const void* pred, const float* target, const uint8_t* mask_u8, int64_t n, const mde_loss_params* params) {
  SplitArgs a{};
  a.pred = pred;
  a.gt = target;
  a.mask = mask_u8;
  a.n = n;
  a.vf = params ? params->variance_focus : 0.85f;
  a.clamp_val = params ? params->clamp_val : 1e-9f;
  a.use_logs = params ? params->use_logs : 1;
  a.size_average = params ? params->size_average : 1;
  a.grad_scale = 1.0f;
  return a;
}

}  // namespace
}  // namespace mde

extern "C" int mde_masked_loss_partials(int kind, int stage, const void* pred, int pred_dtype, const float* target,
                                        const uint8_t* mask_u8, int64_t n, const mde_loss_params* params,
                                        const double* gmax_in, double* partials, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && partials, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(stage == 0 || stage == 1, MDE_EINVAL, "stage must be 0 (max) or 1 (sums)");
  const bool needs_max = (kind == MDE_LOSS_BERHU || kind == MDE_LOSS_LAINA_BERHU);
  MDE_REQUIRE(!(stage == 1 && needs_max && gmax_in == nullptr), MDE_EINVAL, "berHu / Laina sums need the global max (gmax_in)");
  SplitArgs a = split_args(pred, target, mask_u8, n, params);
  a.gmax_in = gmax_in;
  a.partials = partials;
  return split_kind(kind, stage, a, pred_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int mde_masked_loss_from_totals(int kind, const void* pred, int pred_dtype, const float* target,
                                           const uint8_t* mask_u8, int64_t n, const mde_loss_params* params,
                                           const double* totals, float grad_scale, float* loss_out, void* grad,
                                           void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && totals && loss_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n > 0, MDE_EINVAL, "empty input");
  SplitArgs a = split_args(pred, target, mask_u8, n, params);
  a.totals = totals;
  a.grad_scale = grad_scale;
  a.loss_out = loss_out;
  a.grad = grad;
  return split_kind(kind, 2, a, pred_dtype, static_cast<cudaStream_t>(stream));
}
