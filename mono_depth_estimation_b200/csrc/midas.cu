// midas.cu - scale-and-shift alignment of a relative-depth prediction to the target (the
// evaluation-side part of the MiDaS family).
//
//   compute_scale_and_shift(prediction, target, mask)   reference criteria.py:154-176
//   MidasModule.scale_shift(pred, target)               reference modules/midas.py:56-62  (s * pred + t)
//
// Per image: the masked sums a00 = sum m p^2, a01 = sum m p, a11 = sum m, b0 = sum m p t, b1 = sum m t, then the
// 2x2 solve x = A^-1 b (zeros where det == 0, criteria.py:170-174). The reference makes 5 masked full-tensor
// products and 5 reductions; here one pass over pred/target (8 B/px) accumulates all five in fp64, and the last
// CTA solves every image in fp64 and re-zeroes the per-image
// accumulators (workspace region 0, shared with the metrics kernel: zero on entry, zero on exit).
#include "common.cuh"

namespace mde {
namespace {

constexpr int kMBlock = 256;
constexpr int kMWarps = kMBlock / 32;
constexpr int kMChunk = kMBlock * 16;   // pixels per CTA step

template <typename PT>
__global__ void __launch_bounds__(kMBlock) scale_shift_kernel(const PT* __restrict__ pred, const float* __restrict__ gt,
                                                             const uint8_t* __restrict__ mask, int64_t n_img, int64_t hw,
                                                             int chunks_per_img, void* ws_raw, float* __restrict__ scale_out,
                                                             float* __restrict__ shift_out) {
  __shared__ double sm[5 * kMWarps];
  __shared__ bool sm_last;
  Ws ws = ws_view(ws_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_work = n_img * chunks_per_img;
  const int64_t per_chunk = ((hw + chunks_per_img - 1) / chunks_per_img + kMBlock - 1) / kMBlock * kMBlock;
  for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
    const int64_t img = wk / chunks_per_img;
    const int64_t c0 = (wk - img * chunks_per_img) * per_chunk;
    int64_t c1 = c0 + per_chunk;
    if (c1 > hw) c1 = hw;
    const PT* p_img = pred + img * hw;
    const float* t_img = gt + img * hw;
    const uint8_t* m_img = mask ? mask + img * hw : nullptr;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    // fp64 throughout: p*p and p*t are exact in fp64, so a system that is singular in exact arithmetic (one
    // valid pixel: a00 a11 - a01^2 = p^2 - p^2) stays exactly singular, as it does in the reference's fp32 chain
    for (int64_t base = c0; base < c1; base += kMChunk) {
#pragma unroll 4
      for (int k = 0; k < 16; ++k) {
        const int64_t i = base + k * kMBlock + threadIdx.x;
        if (i < c1) {
          const float p = Elem<PT>::ld1(p_img + i), t = __ldg(t_img + i);
          const bool m = m_img ? (m_img[i] != 0) : (t > 0.f);     // criteria.py:155-156
          const double pm = m ? static_cast<double>(p) : 0.0, tm = m ? static_cast<double>(t) : 0.0;
          acc[0] = fma(pm, pm, acc[0]); acc[1] += pm; acc[2] += m ? 1.0 : 0.0;
          acc[3] = fma(pm, tm, acc[3]); acc[4] += tm;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const double s = warp_sum(acc[q]);
      if (lane == 0) sm[q * kMWarps + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      double tot = 0.0;
      for (int w = 0; w < kMWarps; ++w) tot += sm[threadIdx.x * kMWarps + w];
      if (tot != 0.0) atomicAdd(&ws.iacc[img * kIacc + threadIdx.x], tot);
    }
    __syncthreads();
  }
  __threadfence();
  if (threadIdx.x == 0) sm_last = (atomicAdd(&ws.hdr->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!sm_last) return;
  __threadfence();
  for (int64_t b = threadIdx.x; b < n_img; b += kMBlock) {
    double* r = ws.iacc + b * kIacc;
    const double a00 = __ldcg(r + 0), a01 = __ldcg(r + 1), a11 = __ldcg(r + 2), b0 = __ldcg(r + 3), b1 = __ldcg(r + 4);
    const double det = a00 * a11 - a01 * a01;
    double x0 = 0.0, x1 = 0.0;
    if (det != 0.0) {                                      // criteria.py:170-174
      x0 = (a11 * b0 - a01 * b1) / det;
      x1 = (-a01 * b0 + a00 * b1) / det;
    }
    scale_out[b] = static_cast<float>(x0);
    shift_out[b] = static_cast<float>(x1);
#pragma unroll
    for (int q = 0; q < 5; ++q) r[q] = 0.0;
  }
  if (threadIdx.x == 0) ws.hdr->ticket = 0u;
}

// out = scale[img] * pred + shift[img], multiply and add rounded separately as the reference's two ops are
template <typename PT>
__global__ void __launch_bounds__(kMBlock) apply_scale_shift_kernel(const PT* __restrict__ pred, const float* __restrict__ scale,
                                                                   const float* __restrict__ shift, int64_t n_img, int64_t hw,
                                                                   float* __restrict__ out) {
  const int64_t total = n_img * hw;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kMBlock + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kMBlock) {
    const int64_t b = i / hw;
    __stcs(out + i, __fadd_rn(__fmul_rn(__ldg(scale + b), Elem<PT>::ld1(pred + i)), __ldg(shift + b)));
  }
}

template <typename PT>
int launch_scale_shift(const void* pred, const float* gt, const uint8_t* mask, int64_t n_img, int64_t hw, void* ws,
                       float* scale_out, float* shift_out, cudaStream_t st) {
  const int64_t cap = static_cast<int64_t>(sm_count()) * 4;
  int64_t cpi = cap / n_img;
  const int64_t max_cpi = (hw + kMChunk - 1) / kMChunk;
  if (cpi > max_cpi) cpi = max_cpi;
  if (cpi < 1) cpi = 1;
  int64_t grid = n_img * cpi;
  if (grid > cap) grid = cap;
  scale_shift_kernel<PT><<<static_cast<unsigned>(grid), kMBlock, 0, st>>>(static_cast<const PT*>(pred), gt, mask, n_img, hw,
                                                                         static_cast<int>(cpi), ws, scale_out, shift_out);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

template <typename PT>
int launch_apply(const void* pred, const float* scale, const float* shift, int64_t n_img, int64_t hw, float* out,
                 cudaStream_t st) {
  int64_t grid = (n_img * hw + kMBlock - 1) / kMBlock;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (grid > cap) grid = cap;
  apply_scale_shift_kernel<PT><<<static_cast<unsigned>(grid), kMBlock, 0, st>>>(static_cast<const PT*>(pred), scale, shift, n_img,
                                                                               hw, out);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

}  // namespace
}  // namespace mde

extern "C" int mde_scale_and_shift(const void* pred, int pred_dtype, const float* target, const uint8_t* mask_u8,
                                   int64_t n_img, int64_t hw, void* ws, float* scale_out, float* shift_out, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && ws && scale_out && shift_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (pred_dtype) {
    case MDE_F32: return launch_scale_shift<float>(pred, target, mask_u8, n_img, hw, ws, scale_out, shift_out, st);
    case MDE_F16: return launch_scale_shift<__half>(pred, target, mask_u8, n_img, hw, ws, scale_out, shift_out, st);
    case MDE_BF16: return launch_scale_shift<__nv_bfloat16>(pred, target, mask_u8, n_img, hw, ws, scale_out, shift_out, st);
    default: set_error("mde_scale_and_shift: unknown pred_dtype %d", pred_dtype); return MDE_EINVAL;
  }
}

extern "C" int mde_apply_scale_shift(const void* pred, int pred_dtype, const float* scale, const float* shift, int64_t n_img,
                                     int64_t hw, float* out, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && scale && shift && out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (pred_dtype) {
    case MDE_F32: return launch_apply<float>(pred, scale, shift, n_img, hw, out, st);
    case MDE_F16: return launch_apply<__half>(pred, scale, shift, n_img, hw, out, st);
    case MDE_BF16: return launch_apply<__nv_bfloat16>(pred, scale, shift, n_img, hw, out, st);
    default: set_error("mde_apply_scale_shift: unknown pred_dtype %d", pred_dtype); return MDE_EINVAL;
  }
}
