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
// Small inputs (fp32, W % 4 == 0, <= 606 k pixels) run in eigen_resident_kernel further down instead: every pixel stays in
// registers from the first load to the gradient store, no grid barrier and no atomics (C1: 16.3 -> 8.75 us).
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


// ---- small inputs: every pixel in registers (round 2, session 4) --------------------------------------------------
// The cooperative kernel above walks its chunk twice through L2 with scalar loads, a 32-bit division per pixel, per-image
// fp64 atomics and a grid barrier: at the reference's own C1 size (8x1x228x304, 1 us of HBM time) that is 19 us. Here one
// CTA of 1024 threads per SM holds ONE quad per thread (with the quads above and below it and the two pixels beside it:
// 6 x 128 bit + 4 x 32 bit, all in flight at once) from the first load to the gradient store, like resident_loss.cuh:
//   * the four global sums {My, Ey, Mx, Ex} travel through the slot exchange (grid_sum4_counted: every CTA publishes
//     self-validating words and gathers all slots itself);
//   * a CTA covers 4096 consecutive pixels, i.e. at most TWO images (H W >= 4096 is required), so its per-image sums are
//     two triples {n, S1, S2}: published in the CTA's row of ws.mslots, and every CTA gathers, per image, the rows of the
//     contiguous CTA range that overlaps that image (one warp per image, fixed order: bit-reproducible);
//   * every CTA then forms D, A and the loss itself, and the gradient comes from the registers (5-point stencil).
// Taken for fp32 predictions with W % 4 == 0 (a quad never straddles a row), 128-bit aligned pointers, at most one quad
// per thread of one CTA per SM (606 k pixels on 148 SMs) and at most kEsMaxImg images; everything else runs above.
constexpr int kEsThreads = 1024;
constexpr int kEsWarps = kEsThreads / 32;
constexpr unsigned kEsPix = kEsThreads * 4;   // pixels of a CTA
constexpr int kEsMaxImg = 128;
constexpr int kEsMaxGrid = 192;               // rows of ws.mslots (kMetSlotCtas)
static_assert(kEsMaxGrid <= kMetSlotCtas && 12 <= kMetSlotWords, "per-image triples must fit the CTA's mslots row");

__global__ void __launch_bounds__(kEsThreads, 1) eigen_resident_kernel(EigenArgs a) {
  __shared__ double sm_own[4 * kEsWarps];
  __shared__ double sm_two[8 * kEsWarps];
  __shared__ double sm_gather[kEsWarps * 4];
  __shared__ double sm_tot[4];
  __shared__ double sm_img[kEsMaxImg * 3];
  __shared__ float sm_k[4];
  __shared__ unsigned sm_epoch;

  const float* __restrict__ pred = static_cast<const float*>(a.pred);
  const float* __restrict__ gt = a.gt;
  float* grad = static_cast<float*>(a.grad);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = static_cast<int>(gridDim.x), cta = static_cast<int>(blockIdx.x);
  const unsigned W = static_cast<unsigned>(a.w), H = static_cast<unsigned>(a.h), HW = H * W;
  const unsigned nq = (static_cast<unsigned>(a.n_img) * HW) >> 2;
  const unsigned q = static_cast<unsigned>(cta) * kEsThreads + tid;
  const bool has = q < nq;
  const unsigned idx = q << 2;
  const unsigned img_a = (static_cast<unsigned>(cta) * kEsPix) / HW;   // first image of this CTA; the other one is img_a + 1

  pdl_wait();   // launched with launch_pdl: nothing a predecessor wrote may be read before this point
  Ws ws = ws_view(a.ws);
  unsigned epoch_reg = 0u;
  if (tid == 0) epoch_reg = __ldcg(&ws.hdr->epoch);

  unsigned img = 0u, i = 0u, j = 0u;
  if (has) {
    img = idx / HW;
    const unsigned rem = idx - img * HW;
    i = rem / W;
    j = rem - i * W;
  }
  const bool has_d = has && (i + 1u < H), has_u = has && (i >= 1u), has_l = has && (j >= 1u), has_r = has && (j + 4u < W);
  const unsigned wq = W >> 2;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // a neighbour that does not exist reads as target 0: its pair mask is false and its gradient term vanishes
  float4 p4 = z4, t4 = z4, pd4 = z4, td4 = z4, pu4 = z4, tu4 = z4;
  float pl = 0.f, tl = 0.f, pr = 0.f, tr = 0.f;
  const float4* pred4 = reinterpret_cast<const float4*>(pred);
  const float4* gt4 = reinterpret_cast<const float4*>(gt);
  if (has) { p4 = __ldg(pred4 + q); t4 = __ldg(gt4 + q); }
  if (has_d) { pd4 = __ldg(pred4 + q + wq); td4 = __ldg(gt4 + q + wq); }
  if (has_u) { pu4 = __ldg(pred4 + q - wq); tu4 = __ldg(gt4 + q - wq); }
  if (has_l) { pl = __ldg(pred + idx - 1u); tl = __ldg(gt + idx - 1u); }
  if (has_r) { pr = __ldg(pred + idx + 4u); tr = __ldg(gt + idx + 4u); }
  if (tid == 0) sm_epoch = epoch_reg;

  // ---------------- sums from the registers (criteria.py:32-41, :51-60) ------------------------------------------------
  float s1 = 0.f, s2 = 0.f, ey = 0.f, ex = 0.f, cn = 0.f, cy = 0.f, cx = 0.f;   // <= 4 pixels per thread: float counts are exact
  auto px_sum = [&](float p, float t, float pdn, float tdn, float prt, float trt) {
    const bool v = t > 0.f;
    const float mf = v ? 1.f : 0.f;
    const float d = p * mf - t * mf;
    s1 += d;
    s2 = fmaf(d, d, s2);
    cn += mf;
    const bool my = v && (tdn > 0.f);
    const float e1 = (pdn - p) - (tdn - t);
    ey += my ? e1 * e1 : 0.f;
    cy += my ? 1.f : 0.f;
    const bool mx = v && (trt > 0.f);
    const float e2 = (prt - p) - (trt - t);
    ex += mx ? e2 * e2 : 0.f;
    cx += mx ? 1.f : 0.f;
  };
  if (has) {
    px_sum(p4.x, t4.x, pd4.x, td4.x, p4.y, t4.y);
    px_sum(p4.y, t4.y, pd4.y, td4.y, p4.z, t4.z);
    px_sum(p4.z, t4.z, pd4.z, td4.z, p4.w, t4.w);
    px_sum(p4.w, t4.w, pd4.w, td4.w, pr, tr);
  }
  {
    double g4[4] = {static_cast<double>(cy), static_cast<double>(ey), static_cast<double>(cx), static_cast<double>(ex)};
    const double tot = warp_multi_sum<4>(g4);                      // quantity (lane >> 3) & 3
    if ((lane & 7) == 0) sm_own[(lane >> 3) * kEsWarps + warp] = tot;
    const bool second = has && (img != img_a);
    const double dn = static_cast<double>(cn), d1 = static_cast<double>(s1), d2 = static_cast<double>(s2);
    double v8[8] = {second ? 0.0 : dn, second ? 0.0 : d1, second ? 0.0 : d2, 0.0,
                    second ? dn : 0.0, second ? d1 : 0.0, second ? d2 : 0.0, 0.0};
    const double tw = warp_multi_sum<8>(v8);                       // quantity (lane >> 2) & 7
    if ((lane & 3) == 0) sm_two[(lane >> 2) * kEsWarps + warp] = tw;
  }
  __syncthreads();

  // launch parity: CTA 0 cleans the OTHER workspace set for the next cooperative launch (what coop_prologue does)
  const unsigned epoch = sm_epoch;
  const int par = static_cast<int>(epoch & 1u);
  unsigned* ukey = ws.ukey + par * kUkey;
  if (cta == 0) {
    const int o = par ^ 1;
    for (int k = tid; k < kGacc; k += kEsThreads) ws.gacc[o * kGacc + k] = 0.0;
    for (int k = tid; k < kUkey; k += kEsThreads) ws.ukey[o * kUkey + k] = 0u;
    if (tid == 0) {
      const unsigned dirty = __ldcg(&ws.hdr->dirty[o]);
      if (dirty) {
        const unsigned cap = __ldcg(&ws.hdr->max_images);
        double* rows = ws.iacc + static_cast<size_t>(1 + o) * cap * kIacc;
        for (size_t k = 0; k < static_cast<size_t>(dirty) * kIacc; ++k) rows[k] = 0.0;
        ws.hdr->dirty[o] = 0u;
      }
    }
  }

  // ---------------- exchange: four global sums through the slots, two per-image triples through the CTA's mslots row ---
  const unsigned seq3 = epoch * 4u + 3u;
  grid_sum4_counted<kEsWarps>(ws.slots, ukey + 2, epoch * 4u + 2u, sm_own, sm_gather, sm_tot, [&] {
    if (warp < 8 && (warp & 3) != 3) {                              // warps 0-2: first image {n, S1, S2}; 4-6: second image
      const double tot = warp_sum(sm_two[warp * kEsWarps + lane]);
      if (lane == 0) {
        const int k = (warp >> 2) * 3 + (warp & 3);
        const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(tot));
        const unsigned long long tag = static_cast<unsigned long long>(seq3) << 32;
        unsigned long long* w2 = ws.mslots + static_cast<size_t>(cta) * kMetSlotWords + k * 2;
        st_relaxed_u64(w2, tag | (b >> 32));
        st_relaxed_u64(w2 + 1, tag | (b & 0xffffffffull));
      }
    }
  });
  // per-image totals: warp b, b + 32, ... sums image b over the CTAs [c_lo, c_hi] that overlap it (lanes stride the range)
  for (int b = warp; b < a.n_img; b += kEsWarps) {
    const unsigned first_px = static_cast<unsigned>(b) * HW;
    const int c_lo = static_cast<int>(first_px / kEsPix);
    int c_hi = static_cast<int>((first_px + HW - 1u) / kEsPix);
    if (c_hi > G - 1) c_hi = G - 1;
    double x[3] = {0.0, 0.0, 0.0};
    for (int c = c_lo + lane; c <= c_hi; c += 32) {
      const unsigned ia = (static_cast<unsigned>(c) * kEsPix) / HW;
      const unsigned long long* w = ws.mslots + static_cast<size_t>(c) * kMetSlotWords + ((ia == static_cast<unsigned>(b)) ? 0 : 6);
      unsigned long long hi[3], lo[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        hi[k] = ld_relaxed_u64(w + 2 * k);
        lo[k] = ld_relaxed_u64(w + 2 * k + 1);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        while ((hi[k] >> 32) != seq3 || (lo[k] >> 32) != seq3) {
          hi[k] = ld_relaxed_u64(w + 2 * k);
          lo[k] = ld_relaxed_u64(w + 2 * k + 1);
        }
        x[k] += __longlong_as_double(static_cast<long long>((hi[k] << 32) | (lo[k] & 0xffffffffull)));
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double tot = warp_sum(x[k]);
      if (lane == 0) sm_img[b * 3 + k] = tot;
    }
  }
  __syncthreads();

  // ---------------- totals: D = sum n_b^2, A = sum n_b S2_b - 0.5 sum S1_b^2; every CTA derives the coefficients ------
  if (warp == 0) {
    double d0 = 0.0, d1 = 0.0;
    for (int b = lane; b < a.n_img; b += 32) {
      const double nb = sm_img[b * 3 + 0], s1b = sm_img[b * 3 + 1], s2b = sm_img[b * 3 + 2];
      d0 += nb * nb;
      d1 += nb * s2b - 0.5 * s1b * s1b;
    }
    const double D = warp_sum(d0), A = warp_sum(d1);
    if (lane == 0) {
      const double My = sm_tot[0], Ey = sm_tot[1], Mx = sm_tot[2], Ex = sm_tot[3];
      const double gs = static_cast<double>(a.grad_scale);
      sm_k[0] = static_cast<float>(gs / D);
      sm_k[1] = static_cast<float>(2.0 * gs / My);
      sm_k[2] = static_cast<float>(2.0 * gs / Mx);
      if (cta == 0) {
        *a.loss_out = static_cast<float>(A / D + Ey / My + Ex / Mx);
        if (a.totals_out) {
          a.totals_out[0] = D; a.totals_out[1] = A; a.totals_out[2] = My;
          a.totals_out[3] = Ey; a.totals_out[4] = Mx; a.totals_out[5] = Ex;
        }
        ws.hdr->epoch = epoch + 1u;
      }
    }
  }
  __syncthreads();
  pdl_trigger();   // a dependent launch may start filling the SMs this grid leaves
  if (grad == nullptr || !has) return;

  // ---------------- gradient from the registers (5-point stencil) ------------------------------------------------------
  const float kD = sm_k[0], kY = sm_k[1], kX = sm_k[2];
  const float two_nb = static_cast<float>(2.0 * sm_img[img * 3 + 0]);
  const float s1b = static_cast<float>(sm_img[img * 3 + 1]);
  auto px_grad = [&](float p, float t, float pdn, float tdn, float pup, float tup, float prt, float trt, float plf,
                     float tlf) -> float {
    if (!(t > 0.f)) return 0.f;
    const float c = p - t;
    float g = (two_nb * c - s1b) * kD;
    if (tdn > 0.f) g -= kY * ((pdn - tdn) - c);
    if (tup > 0.f) g += kY * (c - (pup - tup));
    if (trt > 0.f) g -= kX * ((prt - trt) - c);
    if (tlf > 0.f) g += kX * (c - (plf - tlf));
    return g;
  };
  float4 g;
  g.x = px_grad(p4.x, t4.x, pd4.x, td4.x, pu4.x, tu4.x, p4.y, t4.y, pl, tl);
  g.y = px_grad(p4.y, t4.y, pd4.y, td4.y, pu4.y, tu4.y, p4.z, t4.z, p4.x, t4.x);
  g.z = px_grad(p4.z, t4.z, pd4.z, td4.z, pu4.z, tu4.z, p4.w, t4.w, p4.y, t4.y);
  g.w = px_grad(p4.w, t4.w, pd4.w, td4.w, pu4.w, tu4.w, pr, tr, p4.z, t4.z);
  __stcs(reinterpret_cast<float4*>(grad) + q, g);
}

// fp32, W % 4 == 0, 128-bit aligned, H W >= 4096, one quad per thread of one 1024-thread CTA per SM; otherwise `taken`
// stays false and the cooperative kernel runs
int launch_eigen_resident(EigenArgs& a, cudaStream_t st, bool& taken) {
  taken = false;
  static const bool off = [] { const char* e = getenv("MDE_NO_RESIDENT"); return e && atoi(e) != 0; }();
  if (off) return MDE_OK;
  if ((a.w & 3) != 0 || !aligned_to(a.pred, 16) || !aligned_to(a.gt, 16) || (a.grad && !aligned_to(a.grad, 16))) return MDE_OK;
  const int64_t hw = static_cast<int64_t>(a.h) * a.w;
  if (hw < static_cast<int64_t>(kEsPix) || a.n_img > kEsMaxImg || a.n_img < 1) return MDE_OK;
  const void* fn = reinterpret_cast<const void*>(&eigen_resident_kernel);
  int cap = coop_grid(fn, kEsThreads, 0);
  if (cap <= 0) return MDE_OK;
  const int sms = sm_count();
  if (cap > sms) cap = sms;                 // one CTA per SM
  if (cap > kEsMaxGrid) cap = kEsMaxGrid;
  const int64_t nq = (static_cast<int64_t>(a.n_img) * hw) >> 2;
  const int64_t nt = (nq + kEsThreads - 1) / kEsThreads;
  if (nt > cap) return MDE_OK;
  taken = true;
  void* args[] = {&a};
  MDE_CUDA_TRY(launch_pdl(fn, dim3(static_cast<unsigned>(nt)), dim3(kEsThreads), args, 0, st, true));
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
  if (pred_dtype == MDE_F32) {
    bool taken = false;
    const int rc = launch_eigen_resident(a, st, taken);
    if (rc != MDE_OK || taken) return rc;
  }
  switch (pred_dtype) {
    case MDE_F32: return launch_eigen<float>(a, st);
    case MDE_F16: return launch_eigen<__half>(a, st);
    case MDE_BF16: return launch_eigen<__nv_bfloat16>(a, st);
    default: set_error("mde_masked_loss(EIGEN): unknown pred_dtype %d", pred_dtype); return MDE_EINVAL;
  }
}

}  // namespace mde
