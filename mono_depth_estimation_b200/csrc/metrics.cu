// metrics.cu - fused masked error-metric kernel (replaces MetricComputation.compute,
// reference metrics.py:58-67, and the metric functions metrics.py:75-109,116-122).
//
// One pass over pred/target (8 B/px algorithmic): every CTA owns a contiguous, 128-byte aligned
// chunk of the batch, accumulates the 4 integer counts and up to 8 float sums PER IMAGE
// (fp32 over <= 16 pixels, fp64 across tiles), flushes with one warp-shuffle/shared-memory block
// reduction per image it touched and one fp64 atomic per quantity, and the last CTA to finish
// turns the per-image sums into pooled values, per-image values and the mean over images.
#include "common.cuh"
#include "metric_math.cuh"

namespace mde {

namespace {

constexpr int kNQ = MDE_METRIC_NQ;
constexpr int kNM = MDE_METRIC_NM;

template <unsigned G, bool Ref>
struct MetricThread {
  MetricTile tile;
  MetricCounts cnt;
  double run[8];

  __device__ __forceinline__ void reset() {
    tile.zero();
    cnt.zero();
#pragma unroll
    for (int i = 0; i < 8; ++i) run[i] = 0.0;
  }
  __device__ __forceinline__ void px(float p, float t) { metric_px<G, Ref>(p, t, tile, cnt); }
  __device__ __forceinline__ void quad(const float4& p, const float4& t) {
    px(p.x, t.x);
    px(p.y, t.y);
    px(p.z, t.z);
    px(p.w, t.w);
  }
  // fold the fp32 tile sums into the fp64 running sums
  __device__ __forceinline__ void fold() {
    run[0] += tile.s_abs;
    run[1] += tile.s_sq;
    if (G & kGrpLog) { run[2] += tile.s_log10; run[7] += tile.s_lnsq; }
    if (G & kGrpLog1p) run[3] += tile.s_sle;
    if (G & kGrpRel) { run[4] += tile.s_absrel; run[5] += tile.s_sqrel; run[6] += tile.s_rsq; }
    tile.zero();
  }
};

// raw-quantity index of run[i]
__constant__ int kRunToQ[8] = {MDE_Q_ABS, MDE_Q_SQ, MDE_Q_LOG10, MDE_Q_SLE,
                               MDE_Q_ABSREL, MDE_Q_SQREL, MDE_Q_RSQ, MDE_Q_LNSQ};

// Run by the LAST CTA only: per-image values, pooled values, mean over images; re-zero the
// per-image accumulators so the workspace is clean for the next call. Kept out of line so its
// fp64 register arrays do not constrain the streaming loop's register allocation.
__device__ __noinline__ void metrics_finalize(Ws ws, int64_t n_img, double* __restrict__ out_f64,
                                              float* __restrict__ out_f32, double* __restrict__ per_image_values,
                                              double* __restrict__ per_image_raw, double* sm_d) {
  double* iacc = ws.iacc;
  double pooled[kNQ];
  double vsum[kNM];
  double nimg_valid = 0.0;
#pragma unroll
  for (int q = 0; q < kNQ; ++q) pooled[q] = 0.0;
#pragma unroll
  for (int m = 0; m < kNM; ++m) vsum[m] = 0.0;

  for (int64_t b = threadIdx.x; b < n_img; b += kBlock) {
    double raw[kNQ], val[kNM];
#pragma unroll
    for (int q = 0; q < kNQ; ++q) {
      raw[q] = __ldcg(&iacc[b * kIacc + q]);
      iacc[b * kIacc + q] = 0.0;  // leave the workspace clean for the next call
      pooled[q] += raw[q];
    }
    metric_values(raw, val);
    if (raw[MDE_Q_NVALID] > 0.0) {
      nimg_valid += 1.0;
#pragma unroll
      for (int m = 0; m < kNM; ++m) vsum[m] += val[m];
    }
    if (per_image_values) {
#pragma unroll
      for (int m = 0; m < kNM; ++m) per_image_values[b * kNM + m] = val[m];
    }
    if (per_image_raw) {
#pragma unroll
      for (int q = 0; q < kNQ; ++q) per_image_raw[b * kNQ + q] = raw[q];
    }
  }
  const double p_tot = block_sum<kNQ>(pooled, sm_d);   // thread q holds pooled total q
  __shared__ double sm_pooled[kNQ];
  if (threadIdx.x < kNQ) sm_pooled[threadIdx.x] = p_tot;
  const double v_tot = block_sum<kNM>(vsum, sm_d);     // thread m holds sum of per-image value m
  double one[1] = {nimg_valid};
  const double n_valid_img = block_sum<1>(one, sm_d);  // thread 0
  __shared__ double sm_nimg;
  if (threadIdx.x == 0) sm_nimg = n_valid_img;
  __syncthreads();
  if (threadIdx.x < kNM) {
    const double mean_v = v_tot / sm_nimg;
    out_f64[kNM + threadIdx.x] = mean_v;
    if (out_f32) out_f32[kNM + threadIdx.x] = static_cast<float>(mean_v);
  }
  if (threadIdx.x < kNQ) out_f64[2 * kNM + threadIdx.x] = sm_pooled[threadIdx.x];
  if (threadIdx.x == 0) {
    double val[kNM];
    metric_values(sm_pooled, val);
    for (int m = 0; m < kNM; ++m) {
      out_f64[m] = val[m];
      if (out_f32) out_f32[m] = static_cast<float>(val[m]);
    }
    out_f64[2 * kNM + kNQ] = sm_nimg;
    ws.hdr->ticket = 0;
  }
}

template <typename PT, int VEC, unsigned G, bool Ref>
__global__ void __launch_bounds__(kBlock, kCtasPerSm)
metrics_kernel(const PT* __restrict__ pred, const float* __restrict__ gt, int64_t n_img, int64_t hw,
               void* ws_raw, double* __restrict__ out_f64, float* __restrict__ out_f32,
               double* __restrict__ per_image_values, double* __restrict__ per_image_raw) {
  __shared__ double sm_d[12 * kWarps];
  __shared__ int sm_i[4 * kWarps];
  __shared__ bool sm_last;

  Ws ws = ws_view(ws_raw);
  double* iacc = ws.iacc;  // parity set 0: [n_img][kIacc]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t units_per_img = hw / VEC;  // VEC == 4 requires hw % 4 == 0 (checked on the host)
  const int64_t total_units = n_img * units_per_img;
  int64_t ub, ue;
  cta_chunk(total_units, 32 / VEC, blockIdx.x, gridDim.x, ub, ue);

  MetricThread<G, Ref> th;
  int64_t u = ub;
  while (u < ue) {
    const int64_t img = u / units_per_img;
    int64_t seg_end = (img + 1) * units_per_img;
    if (seg_end > ue) seg_end = ue;
    th.reset();

    int64_t i = u + threadIdx.x;
    if (VEC == 4) {
      // 2 quads of pred and of target in flight per thread (4 x 16 B)
      for (; i + kBlock < seg_end; i += 2 * kBlock) {
        const float4 p0 = Elem<PT>::template ld4<false>(pred + 4 * i);
        const float4 t0 = Elem<float>::template ld4<false>(gt + 4 * i);
        const float4 p1 = Elem<PT>::template ld4<false>(pred + 4 * (i + kBlock));
        const float4 t1 = Elem<float>::template ld4<false>(gt + 4 * (i + kBlock));
        th.quad(p0, t0);
        th.quad(p1, t1);
        th.fold();
      }
      if (i < seg_end) {
        const float4 p0 = Elem<PT>::template ld4<false>(pred + 4 * i);
        const float4 t0 = Elem<float>::template ld4<false>(gt + 4 * i);
        th.quad(p0, t0);
        th.fold();
      }
    } else {
      for (; i + 3 * kBlock < seg_end; i += 4 * kBlock) {
        float p[4], t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          p[k] = Elem<PT>::ld1(pred + i + k * kBlock);
          t[k] = __ldg(gt + i + k * kBlock);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) th.px(p[k], t[k]);
        th.fold();
      }
      for (; i < seg_end; i += kBlock) {
        th.px(Elem<PT>::ld1(pred + i), __ldg(gt + i));
        th.fold();
      }
    }

    // ---- flush this image's partial sums: warp shuffle -> shared memory -> 12 fp64 atomics ----
    {
      const int c0 = __reduce_add_sync(0xffffffffu, th.cnt.n);
      const int c1 = __reduce_add_sync(0xffffffffu, th.cnt.c1);
      const int c2 = __reduce_add_sync(0xffffffffu, th.cnt.c2);
      const int c3 = __reduce_add_sync(0xffffffffu, th.cnt.c3);
      if (lane == 0) {
        sm_i[0 * kWarps + warp] = c0;
        sm_i[1 * kWarps + warp] = c1;
        sm_i[2 * kWarps + warp] = c2;
        sm_i[3 * kWarps + warp] = c3;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double s = warp_sum(th.run[q]);
        if (lane == 0) sm_d[q * kWarps + warp] = s;
      }
      __syncthreads();
      if (threadIdx.x < 12) {
        double tot = 0.0;
        int qidx;
        if (threadIdx.x < 4) {
          long long ci = 0;
          for (int w = 0; w < kWarps; ++w) ci += sm_i[threadIdx.x * kWarps + w];
          tot = static_cast<double>(ci);
          qidx = threadIdx.x;  // MDE_Q_NVALID, D1, D2, D3
        } else {
          const int q = threadIdx.x - 4;
          for (int w = 0; w < kWarps; ++w) tot += sm_d[q * kWarps + w];
          qidx = kRunToQ[q];
        }
        if (tot != 0.0) atomicAdd(&iacc[img * kIacc + qidx], tot);
      }
      __syncthreads();
    }
    u = seg_end;
  }

  // ---- last CTA finishes: per-image values, pooled values, mean over images; re-zero ws ----
  __threadfence();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&ws.hdr->ticket, 1u);
    sm_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!sm_last) return;
  __threadfence();

  metrics_finalize(ws, n_img, out_f64, out_f32, per_image_values, per_image_raw, sm_d);
}

template <typename PT, int VEC, unsigned G, bool Ref>
int launch_metrics(const void* pred, const float* gt, int64_t n_img, int64_t hw, void* ws, double* out_f64,
                   float* out_f32, double* piv, double* pir, cudaStream_t st) {
  const int64_t units = n_img * (hw / VEC);
  const int64_t per_cta_min = static_cast<int64_t>(kBlock);  // at least one unit per thread
  int64_t grid = (units + per_cta_min - 1) / per_cta_min;
  const int64_t cap = static_cast<int64_t>(sm_count()) * kCtasPerSm;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  metrics_kernel<PT, VEC, G, Ref><<<static_cast<unsigned>(grid), kBlock, 0, st>>>(
      static_cast<const PT*>(pred), gt, n_img, hw, ws, out_f64, out_f32, piv, pir);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

template <typename PT>
int dispatch_metrics(const void* pred, const float* gt, int64_t n_img, int64_t hw, unsigned flags, void* ws,
                     double* out_f64, float* out_f32, double* piv, double* pir, cudaStream_t st) {
  const bool ref = (flags & MDE_METRICS_REFERENCE_MATH) != 0;
  unsigned g = (flags >> 8) & kGrpAll;
  if (g == 0) g = kGrpAll;
  const bool vec = (hw % 4 == 0) && aligned_to(pred, 4 * sizeof(PT)) && aligned_to(gt, 16);
#define MDE_CASE(V, GG, R) return launch_metrics<PT, V, GG, R>(pred, gt, n_img, hw, ws, out_f64, out_f32, piv, pir, st)
  if (!vec) {  // odd image sizes: scalar (still coalesced) path, all groups
    if (ref) MDE_CASE(1, kGrpAll, true);
    MDE_CASE(1, kGrpAll, false);
  }
  if (ref) MDE_CASE(4, kGrpAll, true);
  switch (g) {
    case 1: MDE_CASE(4, 1u, false);
    case 2: MDE_CASE(4, 2u, false);
    case 3: MDE_CASE(4, 3u, false);
    case 4: MDE_CASE(4, 4u, false);
    case 5: MDE_CASE(4, 5u, false);
    case 6: MDE_CASE(4, 6u, false);
    default: MDE_CASE(4, 7u, false);
  }
#undef MDE_CASE
}

}  // namespace
}  // namespace mde

extern "C" int mde_metrics(const void* pred, int pred_dtype, const float* target, int64_t n_img, int64_t hw,
                           unsigned flags, void* ws, double* out_f64, float* out_f32, double* per_image_values,
                           double* per_image_raw, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && ws && out_f64, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(aligned_to(target, 4) && aligned_to(out_f64, 8), MDE_EALIGN, "misaligned pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (pred_dtype) {
    case MDE_F32:
      return dispatch_metrics<float>(pred, target, n_img, hw, flags, ws, out_f64, out_f32, per_image_values,
                                     per_image_raw, st);
    case MDE_F16:
      return dispatch_metrics<__half>(pred, target, n_img, hw, flags, ws, out_f64, out_f32, per_image_values,
                                      per_image_raw, st);
    case MDE_BF16:
      return dispatch_metrics<__nv_bfloat16>(pred, target, n_img, hw, flags, ws, out_f64, out_f32,
                                             per_image_values, per_image_raw, st);
    default:
      set_error("mde_metrics: unknown pred_dtype %d", pred_dtype);
      return MDE_EINVAL;
  }
}

extern "C" void mde_metrics_finalize_host(const double* raw, double* values) { mde::metric_values(raw, values); }
