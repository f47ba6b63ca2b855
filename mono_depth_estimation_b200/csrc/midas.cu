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
#include <cstdlib>

#include "common.cuh"

namespace mde {
namespace {

constexpr int kMBlock = 256;
constexpr int kMWarps = kMBlock / 32;
constexpr int kMChunk = kMBlock * 16;   // pixels per CTA step

// Elementwise pass over [n_img, hw] with per-image coefficients: a CTA takes tiles of 8 * kMBlock pixels of ONE image
// (one division per tile instead of one 64-bit division per pixel), eight independent elements per thread.
// f(img, global index).
template <typename F>
__device__ __forceinline__ void for_each_pixel_by_image(int64_t n_img, int64_t hw, F f) {
  constexpr int64_t kTile = 8 * kMBlock;
  const int64_t tiles_per_img = (hw + kTile - 1) / kTile;
  const int64_t n_tiles = n_img * tiles_per_img;
  for (int64_t w = blockIdx.x; w < n_tiles; w += gridDim.x) {
    const int64_t img = w / tiles_per_img;
    const int64_t c0 = (w - img * tiles_per_img) * kTile + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t i = c0 + k * kMBlock;
      if (i < hw) f(img, img * hw + i);
    }
  }
}
inline unsigned by_image_grid(int64_t n_img, int64_t hw, int per_sm) {
  int64_t g = n_img * ((hw + 8 * kMBlock - 1) / (8 * kMBlock));
  const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

template <typename PT>
__global__ void __launch_bounds__(kMBlock) scale_shift_kernel(const PT* __restrict__ pred, const float* __restrict__ gt,
                                                             const uint8_t* __restrict__ mask, int64_t n_img, int64_t hw,
                                                             int chunks_per_img, void* ws_raw, float* __restrict__ scale_out,
                                                             float* __restrict__ shift_out, double* __restrict__ sums_out) {
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
    if (sums_out) {   // {a00, a01, a11, det, valid} for the backward through the solve (mde_midas_ssi_backward)
      sums_out[b * 5 + 0] = a00; sums_out[b * 5 + 1] = a01; sums_out[b * 5 + 2] = a11;
      sums_out[b * 5 + 3] = det; sums_out[b * 5 + 4] = (det != 0.0) ? 1.0 : 0.0;
    }
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
  for_each_pixel_by_image(n_img, hw, [&](int64_t b, int64_t i) {
    __stcs(out + i, __fadd_rn(__fmul_rn(__ldg(scale + b), Elem<PT>::ld1(pred + i)), __ldg(shift + b)));
  });
}

template <typename PT>
int launch_scale_shift(const void* pred, const float* gt, const uint8_t* mask, int64_t n_img, int64_t hw, void* ws,
                       float* scale_out, float* shift_out, double* sums_out, cudaStream_t st) {
  const int64_t cap = static_cast<int64_t>(sm_count()) * 4;
  int64_t cpi = cap / n_img;
  const int64_t max_cpi = (hw + kMChunk - 1) / kMChunk;
  if (cpi > max_cpi) cpi = max_cpi;
  if (cpi < 1) cpi = 1;
  int64_t grid = n_img * cpi;
  if (grid > cap) grid = cap;
  scale_shift_kernel<PT><<<static_cast<unsigned>(grid), kMBlock, 0, st>>>(static_cast<const PT*>(pred), gt, mask, n_img, hw,
                                                                         static_cast<int>(cpi), ws, scale_out, shift_out, sums_out);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

template <typename PT>
int launch_apply(const void* pred, const float* scale, const float* shift, int64_t n_img, int64_t hw, float* out,
                 cudaStream_t st) {
  apply_scale_shift_kernel<PT><<<by_image_grid(n_img, hw, 8), kMBlock, 0, st>>>(static_cast<const PT*>(pred), scale, shift, n_img,
                                                                                 hw, out);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

}  // namespace
}  // namespace mde

extern "C" int mde_scale_and_shift(const void* pred, int pred_dtype, const float* target, const uint8_t* mask_u8,
                                   int64_t n_img, int64_t hw, void* ws, float* scale_out, float* shift_out,
                                   double* sums_out, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && ws && scale_out && shift_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (pred_dtype) {
    case MDE_F32: return launch_scale_shift<float>(pred, target, mask_u8, n_img, hw, ws, scale_out, shift_out, sums_out, st);
    case MDE_F16: return launch_scale_shift<__half>(pred, target, mask_u8, n_img, hw, ws, scale_out, shift_out, sums_out, st);
    case MDE_BF16: return launch_scale_shift<__nv_bfloat16>(pred, target, mask_u8, n_img, hw, ws, scale_out, shift_out, sums_out, st);
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

// ---- MidasLoss without scale/shift: data term + multi-scale gradient matching, forward and backward -----------
//
//   MidasLoss(alpha, scales, loss in {'mse','l1','trim'}, reduction='batch-based')   reference criteria.py:306-332
//     data   mse_loss :219-223  sum m (p-t)^2 / sum(2 M)       l1_loss :201-206  sum m |t-p| / sum(2 M)
//            trimmed_mae_loss :208-217 slices the (values, indices) TUPLE of torch.sort, i.e. trims nothing:
//            numerically the l1 data term
//     reg    GradientLoss :283-303: for step = 1, 2, 4, ...: on the [::step, ::step] grid
//            (sum |d[x+step] - d[x]| m m' + sum |d[y+step] - d[y]| m m') / sum M_step, d = m (p - t)   (:226-244)
//     reduction_batch_based :179-188: a zero divisor gives 0
//   loss = data + alpha * reg (alpha > 0). This is the criterion of the registered method `my`
//   (modules/my.py:39: MidasLoss(alpha=0.5, loss='mse', reduction='batch-based')).
//
// One cooperative launch: phase A reduces {S_data, N, S_s, N_s (s < scales)} with a 3-point stencil per scale, a grid
// sync publishes them, phase B writes the gradient with the 5-point stencil per scale. The reference runs ~15 masked
// full-tensor ops per scale plus 4 strided slicing copies.
//   Rows of whole quads (W % 4 == 0, 16-byte aligned tensors): a thread owns 4 consecutive pixels of a row; the rows
//   above / below arrive as 128-bit pairs, the pixels beside the quad as scalars, and the coarse scales ride along
//   (pixels 0 and 2 of a quad on even rows are the stride-2 grid, pixel 0 is on every coarser grid whose step divides
//   its row and column): one sweep per phase.
//   Other widths: scalar sweep for scale 0, dense loops over the grid points of every coarser scale, and for the
//   gradient a second pass (after one more grid sync) in which the thread owning a stride-2 pixel adds the share of
//   all the coarse scales.
//   Positions (image, row, column) advance by a host-computed stride: no division per pixel.
namespace mde {
namespace {

constexpr int kMaxScales = 8;

struct MidasArgs {
  const void* pred;
  const float* gt;
  const float* vsrc;    // nullable: validity comes from vsrc > 0 instead of gt > 0 (TrimmedProcrustesLoss: gt is normalised)
  int n_img, h, w;
  unsigned dj, di, dimg;   // grid stride in (columns, rows, images); filled by launch_midas
  unsigned djq, diq, dimgq; // the same for a stride counted in quads (4 pixels of a row); vec4 path
  int vec4;                // w % 4 == 0 and 16-byte aligned tensors: scale 0 runs on quads with 128-bit accesses
  const float* scale;   // nullable: per-image alignment p^ = scale * p + shift (the 'ssi' variants)
  const float* shift;
  int kind;        // 0: mse, 1: l1 (= trim)
  int scales;
  float alpha, grad_scale;
  void* ws;
  float* loss_out;
  void* grad;
};

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// Position of a flat pixel index inside its image, advanced by a constant stride without dividing: the stride's
// (images, rows, columns) decomposition comes from the host.
struct PixPos {
  unsigned img, i, j;
};
__device__ __forceinline__ void pos_advance(PixPos& q, unsigned dj, unsigned di, unsigned dimg, unsigned w, unsigned h) {
  q.j += dj;
  unsigned carry = 0u;
  if (q.j >= w) { q.j -= w; carry = 1u; }
  q.i += di + carry;
  carry = 0u;
  if (q.i >= h) { q.i -= h; carry = 1u; }
  q.img += dimg + carry;
}
__device__ __forceinline__ void pos_advance(PixPos& q, const MidasArgs& a) {
  pos_advance(q, a.dj, a.di, a.dimg, static_cast<unsigned>(a.w), static_cast<unsigned>(a.h));
}

template <typename PT, bool VS, int MAXS, bool SSI>
__global__ void __launch_bounds__(kBlock, kCtasPerSm) midas_loss_kernel(MidasArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm_d[2 * kWarps];
  __shared__ float sm_c[1 + kMaxScales];
  const PT* __restrict__ pred = static_cast<const PT*>(a.pred);
  const float* __restrict__ gt = a.gt;
  PT* __restrict__ grad = static_cast<PT*>(a.grad);
  const unsigned H = a.h, W = a.w;
  const int S = a.scales;
  const unsigned HW = H * W;
  const unsigned total = static_cast<unsigned>(a.n_img) * HW;
  const unsigned tid = blockIdx.x * kBlock + threadIdx.x, nthr = gridDim.x * kBlock;
  constexpr bool ssi = SSI;                 // a.scale != nullptr: per-image alignment on load
  const float* __restrict__ vs = a.vsrc;
  // prediction as the loss sees it: aligned with two separately rounded ops, as `scale * prediction + shift` is
  auto align = [&](float p, float sc, float sh) -> float { return ssi ? __fadd_rn(__fmul_rn(sc, p), sh) : p; };
  // residual of one pixel and its validity (criteria.py:322: mask = target > 0 on the ORIGINAL target)
  auto residual = [&](unsigned idx, float sc, float sh, bool& v) -> float {
    const float t = __ldg(gt + idx);
    const float p = Elem<PT>::ld1(pred + idx);
    v = VS ? (__ldg(vs + idx) > 0.f) : (t > 0.f);
    return align(p, sc, sh) - t;
  };
  // the same for the quad of pixels idx .. idx + 3 of one row (128-bit accesses; idx % 4 == 0)
  auto residual4 = [&](unsigned idx, float sc, float sh, float (&r)[4], bool (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(gt + idx));
    const float4 p = Elem<PT>::template ld4<true>(pred + idx);
    float4 m = t;
    if (VS) m = __ldg(reinterpret_cast<const float4*>(vs + idx));
    r[0] = align(p.x, sc, sh) - t.x; r[1] = align(p.y, sc, sh) - t.y;
    r[2] = align(p.z, sc, sh) - t.z; r[3] = align(p.w, sc, sh) - t.w;
    v[0] = m.x > 0.f; v[1] = m.y > 0.f; v[2] = m.z > 0.f; v[3] = m.w > 0.f;
  };
  // grid point g of scale s (the [::2^s, ::2^s] slicing of criteria.py:298-300) -> image, row, column
  auto grid_point = [&](unsigned g, int s, unsigned Ws, unsigned HWs, PixPos& q) {
    q.img = g / HWs;
    const unsigned rem = g - q.img * HWs;
    const unsigned is = rem / Ws;
    q.i = is << s;
    q.j = (rem - is * Ws) << s;
  };

  Ws ws = ws_view(a.ws);
  unsigned epoch;
  const int par = coop_prologue(ws, epoch);
  double* gacc = ws.gacc + par * kGacc;

  PixPos q0;
  {
    const unsigned first = (tid < total) ? tid : 0u;
    q0.img = first / HW;
    const unsigned rem = first - q0.img * HW;
    q0.i = rem / W;
    q0.j = rem - q0.i * W;
  }

  // ---------------- phase A: data term and scale 0 over every pixel ---------------------------------------------
  {
    // fp32 within a thread for at most 64 pixels, then fp64; counts are exact integers
    float f_data = 0.f, f_s0 = 0.f;
    unsigned c_data = 0u;
    double d_data = 0.0, d_s0 = 0.0;
    PixPos q = q0;
    unsigned it = 0u;
    if (a.vec4) {
      // quads: the row below comes as a second 128-bit pair, the pixel right of the quad as one scalar pair. The coarse
      // scales ride along: on even rows pixels 0 and 2 of the quad are on the stride-2 grid (their neighbours are in
      // the quad, right of it, or two rows down: one more 128-bit pair), pixel 0 is on the grid of every scale >= 2
      // whose step divides its row and column (scalar neighbours).
      const unsigned Wq = W >> 2, nquads = total >> 2;
      float fs[MAXS];
      unsigned cs[MAXS];
      double ds[MAXS];
#pragma unroll
      for (int s = 0; s < MAXS; ++s) { fs[s] = 0.f; cs[s] = 0u; ds[s] = 0.0; }
      PixPos qq;
      {
        const unsigned first = (tid < nquads) ? tid : 0u;
        qq.img = first / (H * Wq);
        const unsigned rem = first - qq.img * (H * Wq);
        qq.i = rem / Wq;
        qq.j = rem - qq.i * Wq;
      }
      for (unsigned qd = tid; qd < nquads; qd += nthr, pos_advance(qq, a.djq, a.diq, a.dimgq, Wq, H)) {
        const unsigned idx = qd << 2, j0 = qq.j << 2;
        const float sc = ssi ? __ldg(a.scale + qq.img) : 1.f, sh = ssi ? __ldg(a.shift + qq.img) : 0.f;
        const bool has_re = j0 + 4u < W, has_d = qq.i + 1u < H;
        float r[4], rd[4];
        bool v[4], vd[4], v_re;
        residual4(idx, sc, sh, r, v);
        residual4(has_d ? idx + W : idx, sc, sh, rd, vd);
        const float r_re = residual(has_re ? idx + 4u : idx, sc, sh, v_re);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float res = v[k] ? r[k] : 0.f;
          f_data += (a.kind == 0) ? res * res : fabsf(res);
          c_data += v[k] ? 1u : 0u;
          const bool vr = (k < 3) ? v[(k + 1) & 3] : (has_re && v_re);
          const float rr = (k < 3) ? r[(k + 1) & 3] : r_re;
          float e = 0.f;
          if (v[k] && vr) e += fabsf(rr - res);
          if (v[k] && has_d && vd[k]) e += fabsf(rd[k] - res);
          fs[0] += e;
        }
        if (S > 1 && (qq.i & 1u) == 0u) {
          const bool has_d2 = qq.i + 2u < H;
          float rd2[4];
          bool vd2[4];
          residual4(has_d2 ? idx + 2u * W : idx, sc, sh, rd2, vd2);
          cs[1] += (v[0] ? 1u : 0u) + (v[2] ? 1u : 0u);
          float e = 0.f;
          if (v[0]) {
            if (v[2]) e += fabsf(r[2] - r[0]);
            if (has_d2 && vd2[0]) e += fabsf(rd2[0] - r[0]);
          }
          if (v[2]) {
            if (has_re && v_re) e += fabsf(r_re - r[2]);
            if (has_d2 && vd2[2]) e += fabsf(rd2[2] - r[2]);
          }
          fs[1] += e;
#pragma unroll
          for (int s = 2; s < MAXS; ++s) {
            if (s >= S) break;
            const unsigned step = 1u << s;
            if (((qq.i | j0) & (step - 1u)) != 0u) break;        // off this grid: off every coarser grid too
            cs[s] += v[0] ? 1u : 0u;
            if (!v[0]) continue;
            const bool has_r = j0 + step < W, has_dn = qq.i + step < H;
            bool vr, vdn;
            const float rr = residual(has_r ? idx + step : idx, sc, sh, vr);
            const float rdn = residual(has_dn ? idx + step * W : idx, sc, sh, vdn);
            float es = 0.f;
            if (has_r && vr) es += fabsf(rr - r[0]);
            if (has_dn && vdn) es += fabsf(rdn - r[0]);
            fs[s] += es;
          }
        }
        if ((++it & 15u) == 0u) {
          d_data += static_cast<double>(f_data); f_data = 0.f;
#pragma unroll
          for (int s = 0; s < MAXS; ++s) { ds[s] += static_cast<double>(fs[s]); fs[s] = 0.f; }
        }
      }
      cs[0] = c_data;
      {
        const double pair[2] = {d_data + static_cast<double>(f_data), static_cast<double>(c_data)};
        const double tot = block_sum<2>(pair, sm_d);
        if (threadIdx.x < 2 && tot != 0.0) atomicAdd(&gacc[threadIdx.x], tot);
      }
#pragma unroll
      for (int s = 0; s < MAXS; ++s) {
        if (s < S) {                                             // uniform over the grid
          const double pair[2] = {ds[s] + static_cast<double>(fs[s]), static_cast<double>(cs[s])};
          const double tot = block_sum<2>(pair, sm_d);
          if (threadIdx.x < 2 && tot != 0.0) atomicAdd(&gacc[2 + 2 * s + threadIdx.x], tot);
        }
      }
    } else {   // rows that are not whole quads: scalar sweep, then dense loops over the coarse grids
      for (unsigned idx = tid; idx < total; idx += nthr, pos_advance(q, a)) {
        const float sc = ssi ? __ldg(a.scale + q.img) : 1.f, sh = ssi ? __ldg(a.shift + q.img) : 0.f;
        const bool has_r = q.j + 1u < W, has_d = q.i + 1u < H;
        bool v, v_r, v_d;                                        // centre, right and lower neighbour requested together
        const float res_c = residual(idx, sc, sh, v);
        const float res_r = residual(has_r ? idx + 1u : idx, sc, sh, v_r);
        const float res_d = residual(has_d ? idx + W : idx, sc, sh, v_d);
        const float res = v ? res_c : 0.f;
        f_data += (a.kind == 0) ? res * res : fabsf(res);
        c_data += v ? 1u : 0u;
        float e = 0.f;
        if (v && has_r && v_r) e += fabsf(res_r - res);
        if (v && has_d && v_d) e += fabsf(res_d - res);
        f_s0 += e;
        if ((++it & 63u) == 0u) {
          d_data += static_cast<double>(f_data); f_data = 0.f;
          d_s0 += static_cast<double>(f_s0); f_s0 = 0.f;
        }
      }
      {
        const double pair[2] = {d_data + static_cast<double>(f_data), static_cast<double>(c_data)};
        const double tot = block_sum<2>(pair, sm_d);
        if (threadIdx.x < 2 && tot != 0.0) atomicAdd(&gacc[threadIdx.x], tot);
      }
      if (S > 0) {
        const double pair[2] = {d_s0 + static_cast<double>(f_s0), static_cast<double>(c_data)};
        const double tot = block_sum<2>(pair, sm_d);
        if (threadIdx.x < 2 && tot != 0.0) atomicAdd(&gacc[2 + threadIdx.x], tot);
      }
      // coarser scales: a dense loop over the grid points of each scale (every lane has work)
#pragma unroll 1
      for (int s = 1; s < S; ++s) {
        const unsigned step = 1u << s;
        const unsigned Hs = (H + step - 1u) >> s, Ws = (W + step - 1u) >> s, HWs = Hs * Ws;
        const unsigned count = static_cast<unsigned>(a.n_img) * HWs;
        float f = 0.f;
        unsigned c = 0u;
        double d = 0.0;
        it = 0u;
        for (unsigned g = tid; g < count; g += nthr) {
          grid_point(g, s, Ws, HWs, q);
          const unsigned idx = q.img * HW + q.i * W + q.j;
          const float sc = ssi ? __ldg(a.scale + q.img) : 1.f, sh = ssi ? __ldg(a.shift + q.img) : 0.f;
          const bool has_r = q.j + step < W, has_d = q.i + step < H;
          bool v, v_r, v_d;
          const float res = residual(idx, sc, sh, v);
          const float res_r = residual(has_r ? idx + step : idx, sc, sh, v_r);
          const float res_d = residual(has_d ? idx + step * W : idx, sc, sh, v_d);
          c += v ? 1u : 0u;
          float e = 0.f;
          if (v && has_r && v_r) e += fabsf(res_r - res);
          if (v && has_d && v_d) e += fabsf(res_d - res);
          f += e;
          if ((++it & 63u) == 0u) { d += static_cast<double>(f); f = 0.f; }
        }
        const double pair[2] = {d + static_cast<double>(f), static_cast<double>(c)};
        const double tot = block_sum<2>(pair, sm_d);
        if (threadIdx.x < 2 && tot != 0.0) atomicAdd(&gacc[2 + 2 * s + threadIdx.x], tot);
      }
    }
  }
  grid.sync();

  // ---------------- totals -> loss and coefficients ---------------------------------------------------------
  if (threadIdx.x == 0) {
    const double Sd = __ldcg(&gacc[0]), N = __ldcg(&gacc[1]);
    double loss = (N > 0.0) ? Sd / (2.0 * N) : 0.0;                 // reduction_batch_based(image_loss, 2 M)
    const double gs = static_cast<double>(a.grad_scale);
    sm_c[0] = (N > 0.0) ? static_cast<float>(gs * (a.kind == 0 ? 1.0 / N : 0.5 / N)) : 0.f;
    for (int s = 0; s < kMaxScales; ++s) sm_c[1 + s] = 0.f;
    for (int s = 0; s < S; ++s) {
      const double Ss = __ldcg(&gacc[2 + 2 * s]), Ns = __ldcg(&gacc[3 + 2 * s]);
      const bool on = (a.alpha > 0.f) && (Ns > 0.0);
      if (on) loss += static_cast<double>(a.alpha) * Ss / Ns;
      sm_c[1 + s] = on ? static_cast<float>(gs * static_cast<double>(a.alpha) / Ns) : 0.f;
    }
    if (blockIdx.x == 0) {
      *a.loss_out = static_cast<float>(loss);
      ws.hdr->epoch = epoch + 1u;
    }
  }
  __syncthreads();
  if (grad == nullptr) return;

  // ---------------- phase B: gradient of the data term and of scale 0, every pixel --------------------------------
  const float cd = sm_c[0], c0 = sm_c[1];
  if (a.vec4) {
    const unsigned Wq = W >> 2, nquads = total >> 2;
    PixPos qq;
    {
      const unsigned first = (tid < nquads) ? tid : 0u;
      qq.img = first / (H * Wq);
      const unsigned rem = first - qq.img * (H * Wq);
      qq.i = rem / Wq;
      qq.j = rem - qq.i * Wq;
    }
    const bool on = S > 0;
    for (unsigned qd = tid; qd < nquads; qd += nthr, pos_advance(qq, a.djq, a.diq, a.dimgq, Wq, H)) {
      const unsigned idx = qd << 2, j0 = qq.j << 2;
      const float sc = ssi ? __ldg(a.scale + qq.img) : 1.f, sh = ssi ? __ldg(a.shift + qq.img) : 0.f;
      const bool has_re = on && (j0 + 4u < W), has_le = on && (j0 >= 1u), has_d = on && (qq.i + 1u < H), has_u = on && (qq.i >= 1u);
      float r[4], rd[4], ru[4];
      bool v[4], vd[4], vu[4], v_re, v_le;
      residual4(idx, sc, sh, r, v);                            // three rows of the quad, then the two pixels beside it
      residual4(has_d ? idx + W : idx, sc, sh, rd, vd);
      residual4(has_u ? idx - W : idx, sc, sh, ru, vu);
      const float r_re = residual(has_re ? idx + 4u : idx, sc, sh, v_re);
      const float r_le = residual(has_le ? idx - 1u : idx, sc, sh, v_le);
      float g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float res = r[k];
        const bool vr = (k < 3) ? (on && v[(k + 1) & 3]) : (has_re && v_re);
        const float rr = (k < 3) ? r[(k + 1) & 3] : r_re;
        const bool vl = (k > 0) ? (on && v[(k + 3) & 3]) : (has_le && v_le);
        const float rl = (k > 0) ? r[(k + 3) & 3] : r_le;
        float sg = 0.f;   // sum over the four pairs of d|.|/d(res of this pixel)
        if (vr) sg -= sgnf(rr - res);
        if (vl) sg += sgnf(res - rl);
        if (has_d && vd[k]) sg -= sgnf(rd[k] - res);
        if (has_u && vu[k]) sg += sgnf(res - ru[k]);
        g[k] = v[k] ? fmaf(c0, sg, cd * (a.kind == 0 ? res : sgnf(res))) : 0.f;
      }
      if (S > 1 && (qq.i & 1u) == 0u) {                        // the coarse scales of pixels 0 and 2 (see phase A)
        const bool has_d2 = qq.i + 2u < H, has_u2 = qq.i >= 2u, has_l2 = j0 >= 2u;
        float rd2[4], ru2[4];
        bool vd2[4], vu2[4], v_l2;
        residual4(has_d2 ? idx + 2u * W : idx, sc, sh, rd2, vd2);
        residual4(has_u2 ? idx - 2u * W : idx, sc, sh, ru2, vu2);
        const float r_l2 = residual(has_l2 ? idx - 2u : idx, sc, sh, v_l2);
        const float c1 = sm_c[2];
        if (v[0]) {
          float sg = 0.f;
          if (v[2]) sg -= sgnf(r[2] - r[0]);
          if (has_l2 && v_l2) sg += sgnf(r[0] - r_l2);
          if (has_d2 && vd2[0]) sg -= sgnf(rd2[0] - r[0]);
          if (has_u2 && vu2[0]) sg += sgnf(r[0] - ru2[0]);
          g[0] = fmaf(c1, sg, g[0]);
        }
        if (v[2]) {
          float sg = 0.f;
          if (has_re && v_re) sg -= sgnf(r_re - r[2]);
          if (v[0]) sg += sgnf(r[2] - r[0]);
          if (has_d2 && vd2[2]) sg -= sgnf(rd2[2] - r[2]);
          if (has_u2 && vu2[2]) sg += sgnf(r[2] - ru2[2]);
          g[2] = fmaf(c1, sg, g[2]);
        }
        if (v[0]) {
#pragma unroll 1
          for (int s = 2; s < S; ++s) {
            const unsigned step = 1u << s;
            if (((qq.i | j0) & (step - 1u)) != 0u) break;
            const bool has_r = j0 + step < W, has_l = j0 >= step, has_dn = qq.i + step < H, has_up = qq.i >= step;
            bool vr, vl, vdn, vup;
            const float rr = residual(has_r ? idx + step : idx, sc, sh, vr);
            const float rl = residual(has_l ? idx - step : idx, sc, sh, vl);
            const float rdn = residual(has_dn ? idx + step * W : idx, sc, sh, vdn);
            const float rup = residual(has_up ? idx - step * W : idx, sc, sh, vup);
            float sg = 0.f;
            if (has_r && vr) sg -= sgnf(rr - r[0]);
            if (has_l && vl) sg += sgnf(r[0] - rl);
            if (has_dn && vdn) sg -= sgnf(rdn - r[0]);
            if (has_up && vup) sg += sgnf(r[0] - rup);
            g[0] = fmaf(sm_c[1 + s], sg, g[0]);
          }
        }
      }
      Elem<PT>::st4(grad + idx, make_float4(g[0], g[1], g[2], g[3]));
    }
    return;                                                    // every scale is in: no second pass
  } else {
    PixPos q = q0;
    for (unsigned idx = tid; idx < total; idx += nthr, pos_advance(q, a)) {
      const float sc = ssi ? __ldg(a.scale + q.img) : 1.f, sh = ssi ? __ldg(a.shift + q.img) : 0.f;
      const bool on = S > 0;
      const bool has_r = on && (q.j + 1u < W), has_l = on && (q.j >= 1u), has_d = on && (q.i + 1u < H), has_u = on && (q.i >= 1u);
      bool v, v_r, v_l, v_d, v_u;                              // the five points of the stencil requested together
      const float res = residual(idx, sc, sh, v);
      const float res_r = residual(has_r ? idx + 1u : idx, sc, sh, v_r);
      const float res_l = residual(has_l ? idx - 1u : idx, sc, sh, v_l);
      const float res_d = residual(has_d ? idx + W : idx, sc, sh, v_d);
      const float res_u = residual(has_u ? idx - W : idx, sc, sh, v_u);
      float g = 0.f;
      if (v) {
        float sg = 0.f;   // sum over the four pairs of d|.|/d(res of this pixel)
        if (has_r && v_r) sg -= sgnf(res_r - res);
        if (has_l && v_l) sg += sgnf(res - res_l);
        if (has_d && v_d) sg -= sgnf(res_d - res);
        if (has_u && v_u) sg += sgnf(res - res_u);
        g = fmaf(c0, sg, cd * (a.kind == 0 ? res : sgnf(res)));
      }
      Elem<PT>::st1(grad + idx, g);
    }
  }
  if (S <= 1) return;
  // coarser scales: every pixel of the stride-2 grid adds its share for all the scales it is on (one thread per pixel,
  // hence no race), after every CTA has written the pass above
  __threadfence();
  grid.sync();
  {
    const unsigned Hs = (H + 1u) >> 1, Ws = (W + 1u) >> 1, HWs = Hs * Ws;
    const unsigned count = static_cast<unsigned>(a.n_img) * HWs;
    PixPos q;
    for (unsigned gp = tid; gp < count; gp += nthr) {
      grid_point(gp, 1, Ws, HWs, q);
      const unsigned idx = q.img * HW + q.i * W + q.j;
      const float sc = ssi ? __ldg(a.scale + q.img) : 1.f, sh = ssi ? __ldg(a.shift + q.img) : 0.f;
      bool v;
      const float res = residual(idx, sc, sh, v);
      if (!v) continue;
      float add = 0.f;
#pragma unroll 1
      for (int s = 1; s < S; ++s) {
        const unsigned step = 1u << s;
        if (((q.i | q.j) & (step - 1u)) != 0u) break;          // off this grid: off every coarser grid too
        const bool has_r = q.j + step < W, has_l = q.j >= step, has_d = q.i + step < H, has_u = q.i >= step;
        bool v_r, v_l, v_d, v_u;
        const float res_r = residual(has_r ? idx + step : idx, sc, sh, v_r);
        const float res_l = residual(has_l ? idx - step : idx, sc, sh, v_l);
        const float res_d = residual(has_d ? idx + step * W : idx, sc, sh, v_d);
        const float res_u = residual(has_u ? idx - step * W : idx, sc, sh, v_u);
        float sg = 0.f;
        if (has_r && v_r) sg -= sgnf(res_r - res);
        if (has_l && v_l) sg += sgnf(res - res_l);
        if (has_d && v_d) sg -= sgnf(res_d - res);
        if (has_u && v_u) sg += sgnf(res - res_u);
        add = fmaf(sm_c[1 + s], sg, add);
      }
      if (add != 0.f) {
        const float g0 = static_cast<float>(__ldcg(grad + idx));
        Elem<PT>::st1(grad + idx, g0 + add);
      }
    }
  }
}

template <typename PT>
int launch_midas(MidasArgs& a, cudaStream_t st) {
  // validity source (TrimmedProcrustes) and alignment (the 'ssi' variants) never come together
  const void* fn;
  if (a.scales <= 4) {
    fn = a.vsrc ? reinterpret_cast<const void*>(&midas_loss_kernel<PT, true, 4, false>)
                : (a.scale ? reinterpret_cast<const void*>(&midas_loss_kernel<PT, false, 4, true>)
                           : reinterpret_cast<const void*>(&midas_loss_kernel<PT, false, 4, false>));
  } else {
    fn = a.vsrc ? reinterpret_cast<const void*>(&midas_loss_kernel<PT, true, kMaxScales, false>)
                : (a.scale ? reinterpret_cast<const void*>(&midas_loss_kernel<PT, false, kMaxScales, true>)
                           : reinterpret_cast<const void*>(&midas_loss_kernel<PT, false, kMaxScales, false>));
  }
  const int64_t n = static_cast<int64_t>(a.n_img) * a.h * a.w;
  int64_t grid = (n + kBlock - 1) / kBlock;
  const int cap = coop_grid(fn, kBlock, 0);
  if (cap <= 0) return MDE_ECUDA;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  {   // the grid stride as (images, rows, columns), so that the kernel advances its position without dividing
    const int64_t nthr = grid * kBlock;
    a.dj = static_cast<unsigned>(nthr % a.w);
    const int64_t t1 = nthr / a.w;
    a.di = static_cast<unsigned>(t1 % a.h);
    a.dimg = static_cast<unsigned>(t1 / a.h);
    const int64_t wq = (a.w % 4 == 0) ? a.w / 4 : 1;
    a.djq = static_cast<unsigned>(nthr % wq);
    const int64_t t2 = nthr / wq;
    a.diq = static_cast<unsigned>(t2 % a.h);
    a.dimgq = static_cast<unsigned>(t2 / a.h);
    a.vec4 = (a.w % 4 == 0) && aligned_to(a.pred, 16) && aligned_to(a.gt, 16) && (a.vsrc == nullptr || aligned_to(a.vsrc, 16)) &&
             (a.grad == nullptr || aligned_to(a.grad, 16)) && (getenv("MDE_MIDAS_SCALAR") == nullptr);
  }
  void* args[] = {&a};
  MDE_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(static_cast<unsigned>(grid)), dim3(kBlock), args, 0, st));
  count_launch();
  return MDE_OK;
}

}  // namespace
}  // namespace mde

namespace mde {
namespace {
int midas_loss_entry(const void* pred, int pred_dtype, const float* target, const float* vsrc, const float* scale,
                     const float* shift, int64_t n_img, int64_t h, int64_t w, int data_kind, float alpha, int scales,
                     float grad_scale, void* ws, float* loss_out, void* grad, void* stream) {
  MDE_REQUIRE(pred && target && ws && loss_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && h > 0 && w > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(n_img * h * w < (int64_t(1) << 31), MDE_ETOOBIG, "more than 2^31 pixels");
  MDE_REQUIRE(data_kind == 0 || data_kind == 1, MDE_EINVAL, "data_kind: 0 (mse) or 1 (l1 / trim)");
  MDE_REQUIRE(scales >= 0 && scales <= kMaxScales, MDE_EINVAL, "scales must be in [0, 8]");
  MDE_REQUIRE((scale == nullptr) == (shift == nullptr), MDE_EINVAL, "scale and shift come together");
  MidasArgs a;
  a.scale = scale; a.shift = shift; a.vsrc = vsrc;
  a.pred = pred; a.gt = target; a.n_img = static_cast<int>(n_img); a.h = static_cast<int>(h); a.w = static_cast<int>(w);
  a.kind = data_kind; a.scales = scales; a.alpha = alpha; a.grad_scale = grad_scale; a.ws = ws; a.loss_out = loss_out;
  a.grad = grad;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (pred_dtype) {
    case MDE_F32: return launch_midas<float>(a, st);
    case MDE_F16: return launch_midas<__half>(a, st);
    case MDE_BF16: return launch_midas<__nv_bfloat16>(a, st);
    default: set_error("mde_midas_loss: unknown pred_dtype %d", pred_dtype); return MDE_EINVAL;
  }
}
}  // namespace
}  // namespace mde

extern "C" int mde_midas_loss(const void* pred, int pred_dtype, const float* target, const float* scale, const float* shift,
                              int64_t n_img, int64_t h, int64_t w, int data_kind, float alpha, int scales, float grad_scale,
                              void* ws, float* loss_out, void* grad, void* stream) {
  return mde::midas_loss_entry(pred, pred_dtype, target, nullptr, scale, shift, n_img, h, w, data_kind, alpha, scales,
                               grad_scale, ws, loss_out, grad, stream);
}

// ---- backward through the alignment: dL/dp from g = dL/dp^ ---------------------------------------------------
// p^_j = s p_j + t with (s, t) the least-squares solution of A [s t]^T = b (A, b: the masked sums above). Perturbing
// p_i changes a00, a01, b0, hence (s, t):  A [ds dt]^T = m_i [y_i - 2 s p_i - t, -s]^T dp_i. With G0 = sum_j g_j and
// G1 = sum_j g_j p_j per image,
//     dL/dp_i = s g_i + m_i (U y_i - 2 s U p_i - t U - s V),   U = (a11 G1 - a01 G0) / det,  V = (a00 G0 - a01 G1) / det
// (zero correction where det == 0: the reference leaves scale = shift = 0 there as constants). Two launches: the
// per-image reduction of (g, g p), then the elementwise update in place on g.
namespace mde {
namespace {

__global__ void __launch_bounds__(kMBlock) ssi_reduce_kernel(const float* __restrict__ pred, const float* __restrict__ g,
                                                            int64_t n_img, int64_t hw, int chunks_per_img, void* ws_raw,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const double* __restrict__ sums, float* __restrict__ coef) {
  __shared__ double sm[2 * kMWarps];
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
    double G0 = 0.0, G1 = 0.0;
    for (int64_t i = c0 + threadIdx.x; i < c1; i += kMBlock) {
      const double gv = static_cast<double>(__ldg(g + img * hw + i));
      G0 += gv;
      G1 = fma(gv, static_cast<double>(__ldg(pred + img * hw + i)), G1);
    }
    const double s0 = warp_sum(G0), s1 = warp_sum(G1);
    if (lane == 0) { sm[warp] = s0; sm[kMWarps + warp] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
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
    const double G0 = __ldcg(r + 0), G1 = __ldcg(r + 1);
    r[0] = 0.0; r[1] = 0.0;
    const double a00 = sums[b * 5 + 0], a01 = sums[b * 5 + 1], a11 = sums[b * 5 + 2], det = sums[b * 5 + 3];
    const double s = static_cast<double>(scale[b]), t = static_cast<double>(shift[b]);
    double U = 0.0, V = 0.0;
    if (det != 0.0) {
      U = (a11 * G1 - a01 * G0) / det;
      V = (a00 * G0 - a01 * G1) / det;
    }
    coef[b * 4 + 0] = static_cast<float>(s);                 // s g_i
    coef[b * 4 + 1] = static_cast<float>(U);                 // U y_i
    coef[b * 4 + 2] = static_cast<float>(-2.0 * s * U);      // -2 s U p_i
    coef[b * 4 + 3] = static_cast<float>(-(t * U + s * V));  // constant
  }
  if (threadIdx.x == 0) ws.hdr->ticket = 0u;
}

__global__ void __launch_bounds__(kMBlock) ssi_update_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                            const float* __restrict__ coef, int64_t n_img, int64_t hw,
                                                            float* __restrict__ g) {
  for_each_pixel_by_image(n_img, hw, [&](int64_t b, int64_t i) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(coef) + b);
    const float y = __ldg(gt + i);
    float out = c.x * g[i];
    if (y > 0.f) out += fmaf(c.y, y, fmaf(c.z, __ldg(pred + i), c.w));
    g[i] = out;
  });
}

}  // namespace
}  // namespace mde

extern "C" int mde_midas_ssi_backward(const float* pred, const float* target, const float* scale, const float* shift,
                                      const double* sums, int64_t n_img, int64_t hw, void* ws, float* coef_scratch,
                                      float* grad_inout, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && scale && shift && sums && ws && coef_scratch && grad_inout, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(aligned_to(coef_scratch, 16), MDE_EALIGN, "coef_scratch must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 4;
  int64_t cpi = cap / n_img;
  const int64_t max_cpi = (hw + kMChunk - 1) / kMChunk;
  if (cpi > max_cpi) cpi = max_cpi;
  if (cpi < 1) cpi = 1;
  int64_t grid = n_img * cpi;
  if (grid > cap) grid = cap;
  ssi_reduce_kernel<<<static_cast<unsigned>(grid), kMBlock, 0, st>>>(pred, grad_inout, n_img, hw, static_cast<int>(cpi), ws, scale,
                                                                    shift, sums, coef_scratch);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  ssi_update_kernel<<<by_image_grid(n_img, hw, 8), kMBlock, 0, st>>>(pred, target, coef_scratch, n_img, hw, grad_inout);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}


// ---- TrimmedProcrustesLoss: robust per-image normalisation (median / mean absolute deviation) -----------------
//
//   normalize_prediction_robust(x, mask)    reference criteria.py:135-152
//     m_b = median over ALL pixels of (mask * x)_b (torch.median: the lower median, the zeros of the masked-out
//           pixels included), 0 for images without a valid pixel;  x' = x - m_b
//     s_b = clamp(sum mask |x'| / sum mask, min 1e-6), 1 for images without a valid pixel;  result x' / s_b
//   TrimmedProcrustesLoss.forward            reference criteria.py:335-363 (the `midas` method's default criterion
//     '--loss ssitrim', modules/midas.py:36-37): trimmed_mae (= l1 as written, :208-217) + alpha * GradientLoss on
//     the two normalised tensors, the mask still being `target > 0` of the ORIGINAL target (:351).
//
// Statistics: the median is found EXACTLY with a 3-round radix select on the order-preserving key of the fp32 bit
// pattern (11 + 11 + 10 bits; shared-memory histograms per unit, one global histogram per (image, tensor)), then one
// more sweep finds the first index that holds it (the element the median's gradient goes to), the deviation sum and
// Z = sum mask sign(x - m). stats row (8 floats): {m, s, n, k as int bits, mask_k, Z, clamped, 0}.
namespace mde {
namespace {

constexpr int kStatW = 8;
constexpr int kRBlock = 1024;
constexpr int kRWarps = kRBlock / 32;

__device__ __forceinline__ unsigned okey(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// scratch (caller-owned, any content on entry): per pair (image, tensor) 4 doubles {n, deviation, Z, -}, then per pair
// 2048 + 4 words {histogram, prefix, rank, first index, -}
constexpr int kRBins = 2048;
constexpr int kRWords = kRBins + 4;

struct RobustArgs {
  const float* pred;
  const float* gt;
  unsigned hw, chunk;     // pixels per image, pixels per unit (multiple of 4)
  int pairs, parts;       // pairs = 2 * n_img; every pair is split into `parts` units
  double* facc;           // [pairs][4]
  unsigned* words;        // [pairs][kRWords]
  float* stats_pred;
  float* stats_gt;
};

// One cooperative launch: every (image, tensor) pair is cut into `parts` units so that all SMs work even for a small
// batch. Round r: each unit histograms digit r of the keys that match the prefix found so far (shared-memory
// atomics) and flushes the non-empty bins to the pair's global histogram;
// the last unit of a pair to arrive picks the bin that holds the rank (per-pair ticket), one grid sync per round hands
// the prefix to the next digit. Then one more sweep for the index, deviation and sign sum.
__global__ void __launch_bounds__(kRBlock, 1) robust_stats_kernel(RobustArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ unsigned hist[kRBins];
  __shared__ unsigned sm_u[4];
  __shared__ double sm_d[3 * kRWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned n = a.hw;
  const int units = a.pairs * a.parts;

  // ---- phase 0: scratch to its initial state ----
  {
    const size_t nw = static_cast<size_t>(a.pairs) * kRWords;
    for (size_t i = static_cast<size_t>(blockIdx.x) * kRBlock + threadIdx.x; i < nw; i += static_cast<size_t>(gridDim.x) * kRBlock) {
      const unsigned w = static_cast<unsigned>(i % kRWords);
      a.words[i] = (w == kRBins + 1) ? ((n - 1u) >> 1) : ((w == kRBins + 2) ? 0xffffffffu : 0u);   // rank of the lower median
    }
    for (int i = blockIdx.x * kRBlock + threadIdx.x; i < a.pairs * 4; i += gridDim.x * kRBlock) a.facc[i] = 0.0;
  }
  __threadfence();
  grid.sync();

  // ---- three radix rounds: 11 + 11 + 10 bits of the order-preserving key ----
  unsigned fixed = 0u;
#pragma unroll 1
  for (int r = 0; r < 3; ++r) {
    const int shift = (r == 0) ? 21 : ((r == 1) ? 10 : 0);
    const unsigned bins = (r == 2) ? 1024u : 2048u;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int pair = u / a.parts, part = u - pair * a.parts;
      const int64_t img = pair >> 1;
      const float* x = ((pair & 1) ? a.gt : a.pred) + img * n;
      const float* t = a.gt + img * n;
      unsigned* w = a.words + static_cast<size_t>(pair) * kRWords;
      const unsigned prefix = __ldcg(w + kRBins);
      const unsigned lo = static_cast<unsigned>(part) * a.chunk;
      unsigned hi = lo + a.chunk;
      if (hi > n) hi = n;
      for (unsigned i = threadIdx.x; i < kRBins; i += kRBlock) hist[i] = 0u;
      __syncthreads();
      for (unsigned i0 = lo; i0 < hi; i0 += 8u * kRBlock) {
        float xv[8], tv[8];                                      // eight elements requested before the first use
        bool ok[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const unsigned i = i0 + static_cast<unsigned>(k) * kRBlock + threadIdx.x;
          ok[k] = i < hi;
          xv[k] = ok[k] ? __ldg(x + i) : 0.f;
          tv[k] = ok[k] ? __ldg(t + i) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const unsigned key = okey((tv[k] > 0.f) ? xv[k] : 0.f * xv[k]);     // mask * x as the reference forms it
          const bool take = ok[k] && ((key & fixed) == prefix);
          const unsigned bin = (key >> shift) & (bins - 1u);
          if (take) atomicAdd(&hist[bin], 1u);
        }
      }
      __syncthreads();
      for (unsigned b = threadIdx.x; b < bins; b += kRBlock) {
        const unsigned c = hist[b];
        if (c) atomicAdd(w + b, c);
      }
      // the last unit of this pair to get here (per-pair ticket) picks the bin that holds the rank
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) sm_u[3] = (atomicAdd(w + kRBins + 3, 1u) == static_cast<unsigned>(a.parts - 1)) ? 1u : 0u;
      __syncthreads();
      if (sm_u[3] == 0u) continue;                             // uniform over the CTA
      __threadfence();
      const unsigned rank = __ldcg(w + kRBins + 1);
      {   // CTA-wide scan, two counters per thread
        const unsigned b0 = 2u * threadIdx.x;
        const unsigned c0 = (b0 < bins) ? __ldcg(w + b0) : 0u, c1 = (b0 + 1u < bins) ? __ldcg(w + b0 + 1u) : 0u;
        const unsigned mine = c0 + c1;
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += y;
        }
        if (lane == 31) hist[warp] = incl;                   // hist is free between the rounds
        __syncthreads();
        if (warp == 0) {
          const unsigned tot = hist[lane];
          unsigned wi = tot;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
          }
          hist[32 + lane] = wi - tot;                        // exclusive prefix of the warp totals
        }
        __syncthreads();
        const unsigned excl = hist[32 + warp] + incl - mine;
        if (excl <= rank && rank < excl + mine) {            // exactly one thread
          const bool second = rank >= excl + c0;
          sm_u[0] = b0 + (second ? 1u : 0u);
          sm_u[1] = rank - excl - (second ? c0 : 0u);
        }
      }
      __syncthreads();
      for (unsigned b = threadIdx.x; b < bins; b += kRBlock) w[b] = 0u;
      if (threadIdx.x == 0) {
        w[kRBins] = prefix | (sm_u[0] << shift);
        w[kRBins + 1] = sm_u[1];
        w[kRBins + 3] = 0u;                                      // ticket for the next round
      }
      __syncthreads();
    }
    fixed |= (bins - 1u) << shift;
    __threadfence();
    grid.sync();                                                 // one per round: the next digit needs every pair's prefix
  }

  // ---- final sweep: first index holding the median, valid count, deviation sum, sign sum ----
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int pair = u / a.parts, part = u - pair * a.parts;
    const int64_t img = pair >> 1;
    const float* x = ((pair & 1) ? a.gt : a.pred) + img * n;
    const float* t = a.gt + img * n;
    unsigned* w = a.words + static_cast<size_t>(pair) * kRWords;
    const unsigned prefix = __ldcg(w + kRBins);
    const float med = okey_inv(prefix);
    const unsigned lo = static_cast<unsigned>(part) * a.chunk;
    unsigned hi = lo + a.chunk;
    if (hi > n) hi = n;
    unsigned kmin = 0xffffffffu;
    float cnt = 0.f, zs = 0.f;                     // exact small integers in fp32 (< 2^24 per thread)
    double dev = 0.0;
    for (unsigned i = lo + threadIdx.x; i < hi; i += kRBlock) {
      const bool v = __ldg(t + i) > 0.f;
      const float xv = __ldg(x + i);
      if (i < kmin && okey(v ? xv : 0.f * xv) == prefix) kmin = i;
      if (v) {
        const float d = xv - med;
        cnt += 1.f;
        dev += static_cast<double>(fabsf(d));
        zs += (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
      }
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    const double c2 = warp_sum(static_cast<double>(cnt)), d2 = warp_sum(dev), z2 = warp_sum(static_cast<double>(zs));
    if (lane == 0) {
      if (kmin != 0xffffffffu) atomicMin(w + kRBins + 2, kmin);
      sm_d[warp] = c2; sm_d[kRWarps + warp] = d2; sm_d[2 * kRWarps + warp] = z2;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double tot = 0.0;
      for (int q = 0; q < kRWarps; ++q) tot += sm_d[threadIdx.x * kRWarps + q];
      if (tot != 0.0) atomicAdd(a.facc + pair * 4 + threadIdx.x, tot);
    }
    // the last unit of the pair writes its statistics row
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(w + kRBins + 3, 1u) == static_cast<unsigned>(a.parts - 1)) {
      __threadfence();
      float* out = ((pair & 1) ? a.stats_gt : a.stats_pred) + img * kStatW;
      const double nn = __ldcg(a.facc + pair * 4), dv = __ldcg(a.facc + pair * 4 + 1), z = __ldcg(a.facc + pair * 4 + 2);
      const unsigned k = __ldcg(w + kRBins + 2);
      float m = 0.f, s = 1.f, clamped = 1.f, mk = 0.f;
      if (nn > 0.0) {                                         // criteria.py:139, :144-150
        m = med;
        const float raw = static_cast<float>(dv / nn);
        clamped = (raw < 1e-6f) ? 1.f : 0.f;
        s = (raw < 1e-6f) ? 1e-6f : raw;
        mk = (k < n && __ldg(t + k) > 0.f) ? 1.f : 0.f;
      }
      out[0] = m; out[1] = s; out[2] = static_cast<float>(nn); out[3] = __uint_as_float(k);
      out[4] = mk; out[5] = static_cast<float>(z); out[6] = clamped; out[7] = 0.f;
    }
    __syncthreads();
  }
}

// x' = (x - m) / s for both tensors: subtraction and IEEE division rounded separately, as the reference's two ops are
__global__ void __launch_bounds__(kMBlock) robust_apply_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                              const float* __restrict__ stats_pred,
                                                              const float* __restrict__ stats_gt, int64_t n_img, int64_t hw,
                                                              float* __restrict__ pred_out, float* __restrict__ gt_out) {
  for_each_pixel_by_image(n_img, hw, [&](int64_t b, int64_t i) {
    const float mp = __ldg(stats_pred + b * kStatW), sp = __ldg(stats_pred + b * kStatW + 1);
    const float mt = __ldg(stats_gt + b * kStatW), st = __ldg(stats_gt + b * kStatW + 1);
    pred_out[i] = __fdiv_rn(__fsub_rn(__ldg(pred + i), mp), sp);
    gt_out[i] = __fdiv_rn(__fsub_rn(__ldg(gt + i), mt), st);
  });
}

// Backward through the normalisation of the prediction, x' = (x - m) / s with m = x_k mask_k (median element k) and
// s = sum mask |x - m| / n (constant where clamped). With g = dL/dx', G = sum g, Gx = sum g x' per image:
//   dL/dx_j = g_j / s - mask_j sign(x_j - m) Gx / (s n)  +  [j == k] mask_k (Gx Z / (s n) - G / s),  Z = sum mask sign(x - m)
// (the Gx terms vanish where s was clamped, every correction where the image has no valid pixel).
__global__ void __launch_bounds__(kMBlock) robust_reduce_kernel(const float* __restrict__ xn, const float* __restrict__ g,
                                                               int64_t n_img, int64_t hw, int chunks_per_img, void* ws_raw,
                                                               const float* __restrict__ stats, float* __restrict__ coef) {
  __shared__ double sm[2 * kMWarps];
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
    double G0 = 0.0, G1 = 0.0;
    for (int64_t i = c0 + threadIdx.x; i < c1; i += kMBlock) {
      const double gv = static_cast<double>(__ldg(g + img * hw + i));
      G0 += gv;
      G1 = fma(gv, static_cast<double>(__ldg(xn + img * hw + i)), G1);
    }
    const double s0 = warp_sum(G0), s1 = warp_sum(G1);
    if (lane == 0) { sm[warp] = s0; sm[kMWarps + warp] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
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
    const double G = __ldcg(r + 0), Gx = __ldcg(r + 1);
    r[0] = 0.0; r[1] = 0.0;
    const float* sr = stats + b * kStatW;
    const double s = static_cast<double>(sr[1]), n = static_cast<double>(sr[2]), mk = static_cast<double>(sr[4]);
    const double Z = static_cast<double>(sr[5]);
    const bool live = (n > 0.0) && (sr[6] == 0.f);
    const double q = live ? Gx / (s * n) : 0.0;
    coef[b * 4 + 0] = static_cast<float>(1.0 / s);
    coef[b * 4 + 1] = static_cast<float>(-q);
    coef[b * 4 + 2] = (n > 0.0) ? static_cast<float>(mk * (q * Z - G / s)) : 0.f;
    coef[b * 4 + 3] = sr[3];                                  // k (bit pattern of an unsigned)
  }
  if (threadIdx.x == 0) ws.hdr->ticket = 0u;
}

__global__ void __launch_bounds__(kMBlock) robust_update_kernel(const float* __restrict__ xn, const float* __restrict__ gt,
                                                               const float* __restrict__ coef, int64_t n_img, int64_t hw,
                                                               float* __restrict__ g) {
  for_each_pixel_by_image(n_img, hw, [&](int64_t b, int64_t i) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(coef) + b);
    float out = c.x * g[i];
    if (__ldg(gt + i) > 0.f) out = fmaf(c.y, sgnf(__ldg(xn + i)), out);     // sign(x - m) = sign(x')
    if (static_cast<unsigned>(i - b * hw) == __float_as_uint(c.w)) out += c.z;
    g[i] = out;
  });
}

}  // namespace
}  // namespace mde

extern "C" size_t mde_robust_scratch_bytes(int64_t n_img) {
  if (n_img < 1) n_img = 1;
  return static_cast<size_t>(2 * n_img) * (4 * sizeof(double) + mde::kRWords * sizeof(unsigned));
}

extern "C" int mde_robust_normalize(const float* pred, const float* target, int64_t n_img, int64_t hw, void* scratch,
                                    float* stats_pred, float* stats_target, float* pred_out, float* target_out, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && scratch && stats_pred && stats_target && pred_out && target_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(hw < (int64_t(1) << 31) && n_img < (int64_t(1) << 20), MDE_ETOOBIG, "image or batch too large");
  MDE_REQUIRE(aligned_to(scratch, 8), MDE_EALIGN, "scratch must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    const void* fn = reinterpret_cast<const void*>(&robust_stats_kernel);
    const int cap = coop_grid(fn, kRBlock, 0);
    if (cap <= 0) return MDE_ECUDA;
    RobustArgs a;
    a.pred = pred; a.gt = target; a.hw = static_cast<unsigned>(hw); a.pairs = static_cast<int>(2 * n_img);
    int64_t parts = cap / a.pairs;                       // fill the GPU, but keep >= 8192 pixels per unit
    const int64_t max_parts = (hw + 8191) / 8192;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    a.parts = static_cast<int>(parts);
    a.chunk = static_cast<unsigned>(((hw + parts - 1) / parts + 3) / 4 * 4);
    a.facc = static_cast<double*>(scratch);
    a.words = reinterpret_cast<unsigned*>(static_cast<char*>(scratch) + static_cast<size_t>(a.pairs) * 4 * sizeof(double));
    a.stats_pred = stats_pred; a.stats_gt = stats_target;
    int64_t grid = static_cast<int64_t>(a.pairs) * a.parts;
    if (grid > cap) grid = cap;
    void* args[] = {&a};
    MDE_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(static_cast<unsigned>(grid)), dim3(kRBlock), args, 0, st));
    count_launch();
  }
  robust_apply_kernel<<<by_image_grid(n_img, hw, 8), kMBlock, 0, st>>>(pred, target, stats_pred, stats_target, n_img, hw, pred_out,
                                                                        target_out);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

extern "C" int mde_midas_loss_masked(const void* pred, int pred_dtype, const float* target, const float* valid_src,
                                     int64_t n_img, int64_t h, int64_t w, int data_kind, float alpha, int scales,
                                     float grad_scale, void* ws, float* loss_out, void* grad, void* stream) {
  using namespace mde;
  MDE_REQUIRE(valid_src, MDE_EINVAL, "null pointer");
  return midas_loss_entry(pred, pred_dtype, target, valid_src, nullptr, nullptr, n_img, h, w, data_kind, alpha, scales,
                          grad_scale, ws, loss_out, grad, stream);
}

extern "C" int mde_robust_backward(const float* pred_norm, const float* target, const float* stats_pred, int64_t n_img,
                                   int64_t hw, void* ws, float* coef_scratch, float* grad_inout, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred_norm && target && stats_pred && ws && coef_scratch && grad_inout, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(aligned_to(coef_scratch, 16), MDE_EALIGN, "coef_scratch must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 4;
  int64_t cpi = cap / n_img;
  const int64_t max_cpi = (hw + kMChunk - 1) / kMChunk;
  if (cpi > max_cpi) cpi = max_cpi;
  if (cpi < 1) cpi = 1;
  int64_t grid = n_img * cpi;
  if (grid > cap) grid = cap;
  robust_reduce_kernel<<<static_cast<unsigned>(grid), kMBlock, 0, st>>>(pred_norm, grad_inout, n_img, hw, static_cast<int>(cpi), ws,
                                                                       stats_pred, coef_scratch);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  robust_update_kernel<<<by_image_grid(n_img, hw, 8), kMBlock, 0, st>>>(pred_norm, target, coef_scratch, n_img, hw, grad_inout);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
