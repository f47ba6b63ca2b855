// eigen.cu - MaskedDepthLoss (Eigen scale-invariant + gradient term), reference criteria.py:17-64,
// forward and backward in one cooperative launch.
//
//   mask = target > 0; per image b: n_b, S1_b = sum d, S2_b = sum d^2 with d = (pred - target)*mask
//   depth = (sum_b n_b S2_b - 0.5 sum_b S1_b^2) / sum_b n_b^2                       (criteria.py:38-41)
//   grad  = sum m_y e_y^2 / sum m_y + sum m_x e_x^2 / sum m_x,  e = forward difference of
//           (pred - target) along H / W, pair mask m = mask[i] & mask[i+1]          (criteria.py:51-60)
//
// Phase A reduces {n_b,S1_b,S2_b} per image and {My,Ey,Mx,Ex} globally with a 3-point stencil
// (centre, right, down - neighbours come from L1/L2), the grid barrier publishes them, phase B
// writes the gradient with the 5-point stencil. Pixels are walked flat, one contiguous chunk per
// CTA; the CTA flushes its per-image sums whenever its chunk crosses an image boundary.
#include "common.cuh"

namespace mde {
namespace {

struct EigenArgs {
  const void* pred;
  const float* gt;
  int n_img, h, w;
  Chunking chunk;  // elements
  float grad_scale;
  void* ws;
  float* loss_out;
  double* totals_out;
  void* grad;
};

template <typename PT>
__global__ void __launch_bounds__(kBlock, kCtasPerSm) eigen_loss_kernel(EigenArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm_d[4 * kWarps];
  __shared__ double sm_bc[2];

  const PT* __restrict__ pred = static_cast<const PT*>(a.pred);
  const float* __restrict__ gt = a.gt;
  PT* __restrict__ grad = static_cast<PT*>(a.grad);
  const int H = a.h, W = a.w;
  const unsigned HW = static_cast<unsigned>(H) * static_cast<unsigned>(W);

  Ws ws = ws_view(a.ws);
  const unsigned cap = __ldcg(&ws.hdr->max_images);
  if (static_cast<unsigned>(a.n_img) > cap) {  // workspace too small: flag and bail out (grid-uniform)
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.hdr->error = 1u;
    return;
  }
  unsigned epoch;
  const int par = coop_prologue(ws, epoch);
  double* gacc = ws.gacc + par * kGacc;
  double* irow = ws.iacc + static_cast<size_t>(1 + par) * cap * kIacc;

  int64_t cb64, ce64;
  cta_chunk(a.chunk, blockIdx.x, cb64, ce64);
  const unsigned cb = static_cast<unsigned>(cb64), ce = static_cast<unsigned>(ce64);

  // ---------------- phase A ---------------------------------------------------------------------------
  {
    double g_run[4] = {0.0, 0.0, 0.0, 0.0};  // My, Ey, Mx, Ex over the whole chunk
    unsigned u = cb;
    while (u < ce) {
      const unsigned img = u / HW;
      unsigned seg_end = (img + 1) * HW;
      if (seg_end > ce) seg_end = ce;
      double i_run[3] = {0.0, 0.0, 0.0};
      float s1 = 0.f, s2 = 0.f, ey = 0.f, ex = 0.f;
      int cn = 0, cy = 0, cx = 0;
      int it = 0;
      for (unsigned idx = u + threadIdx.x; idx < seg_end; idx += kBlock) {
        const unsigned rem = idx - img * HW;
        const unsigned i = rem / W, j = rem - i * W;
        const float p = Elem<PT>::ld1(pred + idx);
        const float t = __ldg(gt + idx);
        const bool v = t > 0.f;
        const float mf = v ? 1.f : 0.f;
        const float d = p * mf - t * mf;  // criteria.py:32-35
        s1 += d;
        s2 = fmaf(d, d, s2);
        cn += v ? 1 : 0;
        if (i + 1 < static_cast<unsigned>(H)) {
          const float pd = Elem<PT>::ld1(pred + idx + W);
          const float td = __ldg(gt + idx + W);
          const bool m = v && (td > 0.f);
          const float e = (pd - p) - (td - t);
          ey += m ? e * e : 0.f;
          cy += m ? 1 : 0;
        }
        if (j + 1 < static_cast<unsigned>(W)) {
          const float pr = Elem<PT>::ld1(pred + idx + 1);
          const float tr = __ldg(gt + idx + 1);
          const bool m = v && (tr > 0.f);
          const float e = (pr - p) - (tr - t);
          ex += m ? e * e : 0.f;
          cx += m ? 1 : 0;
        }
        if ((++it & 7) == 0) {  // fold fp32 tile sums into fp64 every 8 pixels
          i_run[1] += s1; i_run[2] += s2; g_run[1] += ey; g_run[3] += ex;
          s1 = s2 = ey = ex = 0.f;
        }
      }
      i_run[0] = static_cast<double>(cn);
      i_run[1] += s1;
      i_run[2] += s2;
      g_run[0] += static_cast<double>(cy);
      g_run[1] += ey;
      g_run[2] += static_cast<double>(cx);
      g_run[3] += ex;
      const double tot = block_sum<3>(i_run, sm_d);
      if (threadIdx.x < 3 && tot != 0.0) atomicAdd(&irow[static_cast<size_t>(img) * kIacc + threadIdx.x], tot);
      u = seg_end;
    }
    const double gt_tot = block_sum<4>(g_run, sm_d);
    if (threadIdx.x < 4 && gt_tot != 0.0) atomicAdd(&gacc[threadIdx.x], gt_tot);
  }
  grid.sync();

  // ---------------- totals: D = sum n_b^2, A = sum n_b S2_b - 0.5 sum S1_b^2 -------------------------------
  double da[2] = {0.0, 0.0};
  for (int b = threadIdx.x; b < a.n_img; b += kBlock) {
    const double nb = __ldcg(&irow[static_cast<size_t>(b) * kIacc + 0]);
    const double s1b = __ldcg(&irow[static_cast<size_t>(b) * kIacc + 1]);
    const double s2b = __ldcg(&irow[static_cast<size_t>(b) * kIacc + 2]);
    da[0] += nb * nb;
    da[1] += nb * s2b - 0.5 * s1b * s1b;
  }
  const double da_tot = block_sum<2>(da, sm_d);
  if (threadIdx.x < 2) sm_bc[threadIdx.x] = da_tot;
  __syncthreads();
  const double D = sm_bc[0], A = sm_bc[1];
  const double My = __ldcg(&gacc[0]), Ey = __ldcg(&gacc[1]);
  const double Mx = __ldcg(&gacc[2]), Ex = __ldcg(&gacc[3]);
  const double loss = A / D + Ey / My + Ex / Mx;

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *a.loss_out = static_cast<float>(loss);
    if (a.totals_out) {
      a.totals_out[0] = D; a.totals_out[1] = A; a.totals_out[2] = My;
      a.totals_out[3] = Ey; a.totals_out[4] = Mx; a.totals_out[5] = Ex;
    }
    ws.hdr->dirty[par] = static_cast<unsigned>(a.n_img);
    ws.hdr->epoch = epoch + 1u;
  }
  if (grad == nullptr) return;

  // ---------------- phase B: gradient (5-point stencil), chunk walked backwards --------------------------
  const double gs = static_cast<double>(a.grad_scale);
  const float kD = static_cast<float>(gs / D);
  const float kY = static_cast<float>(2.0 * gs / My);
  const float kX = static_cast<float>(2.0 * gs / Mx);
  unsigned cur_img = 0xffffffffu;
  float two_nb = 0.f, s1b = 0.f;
  if (ce > cb) {
    for (int64_t idx64 = static_cast<int64_t>(ce) - 1 - threadIdx.x; idx64 >= static_cast<int64_t>(cb); idx64 -= kBlock) {
      const unsigned idx = static_cast<unsigned>(idx64);
      const unsigned img = idx / HW;
      if (img != cur_img) {
        cur_img = img;
        two_nb = static_cast<float>(2.0 * __ldcg(&irow[static_cast<size_t>(img) * kIacc + 0]));
        s1b = static_cast<float>(__ldcg(&irow[static_cast<size_t>(img) * kIacc + 1]));
      }
      const unsigned rem = idx - img * HW;
      const unsigned i = rem / W, j = rem - i * W;
      const float p = Elem<PT>::ld1(pred + idx);
      const float t = __ldg(gt + idx);
      const bool v = t > 0.f;
      float g = 0.f;
      if (v) {
        const float c = p - t;
        g = (two_nb * c - s1b) * kD;
        if (i + 1 < static_cast<unsigned>(H)) {
          const float td = __ldg(gt + idx + W);
          if (td > 0.f) g -= kY * ((Elem<PT>::ld1(pred + idx + W) - td) - c);
        }
        if (i >= 1) {
          const float tu = __ldg(gt + idx - W);
          if (tu > 0.f) g += kY * (c - (Elem<PT>::ld1(pred + idx - W) - tu));
        }
        if (j + 1 < static_cast<unsigned>(W)) {
          const float tr = __ldg(gt + idx + 1);
          if (tr > 0.f) g -= kX * ((Elem<PT>::ld1(pred + idx + 1) - tr) - c);
        }
        if (j >= 1) {
          const float tl = __ldg(gt + idx - 1);
          if (tl > 0.f) g += kX * (c - (Elem<PT>::ld1(pred + idx - 1) - tl));
        }
      }
      Elem<PT>::st1(grad + idx, g);
    }
  }
}

template <typename PT>
int launch_eigen(EigenArgs& a, cudaStream_t st) {
  const void* fn = reinterpret_cast<const void*>(&eigen_loss_kernel<PT>);
  const int64_t n = static_cast<int64_t>(a.n_img) * a.h * a.w;
  int64_t grid = (n + kBlock - 1) / kBlock;
  const int cap = coop_grid(fn, kBlock, 0);
  if (cap <= 0) return MDE_ECUDA;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  a.chunk = make_chunking(n, 32, static_cast<int>(grid));
  void* args[] = {&a};
  MDE_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(static_cast<unsigned>(grid)), dim3(kBlock), args, 0, st));
  count_launch();
  return MDE_OK;
}

}  // namespace

int eigen_loss_launch(const void* pred, int pred_dtype, const float* target, int64_t n_img, int64_t h, int64_t w,
                      float grad_scale, void* ws, float* loss_out, double* totals_out, void* grad,
                      cudaStream_t st) {
  if (n_img * h * w >= (int64_t(1) << 31)) {
    set_error("mde_masked_loss(EIGEN): more than 2^31 pixels");
    return MDE_ETOOBIG;
  }
  EigenArgs a;
  a.pred = pred;
  a.gt = target;
  a.n_img = static_cast<int>(n_img);
  a.h = static_cast<int>(h);
  a.w = static_cast<int>(w);
  a.grad_scale = grad_scale;
  a.ws = ws;
  a.loss_out = loss_out;
  a.totals_out = totals_out;
  a.grad = grad;
  switch (pred_dtype) {
    case MDE_F32: return launch_eigen<float>(a, st);
    case MDE_F16: return launch_eigen<__half>(a, st);
    case MDE_BF16: return launch_eigen<__nv_bfloat16>(a, st);
    default: set_error("mde_masked_loss(EIGEN): unknown pred_dtype %d", pred_dtype); return MDE_EINVAL;
  }
}

}  // namespace mde
