// stdepth.cu - the layered-depth ("stdepth") base criterion of the reference's method modules, forward + backward in
// one cooperative launch (the criterion of the registered methods `bts` and `laina`, which inherit
// BaseModule.setup_criterion).
//
//   BaseModule.setup_criterion -> _loss(pred, targ, rgba)        reference modules/base_module.py:124-208
//     mask1 = rgba[:, 3] > 0 (per pixel)                                                            :133
//     maskD = targ[:, D] > 0 (per element of the depth channels D = 8:10, or 16:20 with 20 channels) :137-138
//     'silma' / 'silms': depth_w * nan_to_num(silog(pred[D][maskD], targ[D][maskD]))                 :157, :160
//                        + l1 / mse over the 8 colour channels of the mask1 pixels                    :158, :161
//     'mse' / 'mae'    : mse / l1 over ALL channels of the mask1 pixels + depth_w * the same over maskD :162-167
//     'fbdivergence'   : fbdiv_w * mean over mask1 pixels of the two front/back cosine terms          :184-194
//     loss = sum of the terms present                                                                  :196
//   silog = criteria.silog_loss (criteria.py:724-732): its own mask gt > 1e-2 on top of maskD.
//   The SSIM and compositing terms (:168-183, stdepth_utils.py) are out of scope.
//
// The reference gathers every masked tensor (4-8 boolean gathers of [B,8..20,H,W] tensors with a D->H sync each)
// and runs one loss per gather. Here a thread owns one pixel and walks its C channel planes (a warp reads 32
// consecutive pixels of a plane: full 128-byte lines), phase A reduces 12 totals, a grid sync publishes them,
// phase B writes the gradient of every term at once. Algorithmic traffic: (4C + 4C + 4) read + 4C written per pixel.
#include "common.cuh"

namespace mde {
namespace {

enum : int { ST_SILOG = 1, ST_CMAE = 2, ST_CMSE = 4, ST_ALLMSE = 8, ST_ALLMAE = 16, ST_FBDIV = 32 };

struct StdArgs {
  const void* pred;     // [n_img, C, hw]
  const float* targ;    // [n_img, C, hw]
  const float* alpha;   // plane 3 of image 0 of rgba [n_img, rgba_c, hw]
  int64_t alpha_stride; // rgba_c * hw
  int n_img, flags;
  unsigned hw;
  float depth_w, fbdiv_w, lambda, grad_scale;
  void* ws;
  float* out;           // [8]: total, depth_silog, color, all_mse, all_mae, fb_divergence, n_mask1_px, n_maskD
  void* grad;           // nullable, dtype of pred
};

// accumulator slots
enum : int { A_N1 = 0, A_CABS, A_CSQ, A_AABS, A_ASQ, A_ND, A_DABS, A_DSQ, A_NS, A_SD, A_SDD, A_FB, A_COUNT };

__device__ __forceinline__ float sgn0(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

struct FbTerm { float f, np, nt, dot, mag; };
// one of the two cosine-like terms: sum_c P_c T_c / (|P| |T| + 1e-3)          base_module.py:187-193
__device__ __forceinline__ FbTerm fb_term(const float (&P)[3], const float (&T)[3]) {
  FbTerm r;
  r.np = sqrtf(P[0] * P[0] + P[1] * P[1] + P[2] * P[2]);
  r.nt = sqrtf(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]);
  r.mag = r.np * r.nt + 1e-3f;
  r.dot = P[0] * T[0] + P[1] * T[1] + P[2] * T[2];
  r.f = P[0] * T[0] / r.mag + P[1] * T[1] / r.mag + P[2] * T[2] / r.mag;
  return r;
}

template <typename PT, int C>
__global__ void __launch_bounds__(kBlock, 1) stdepth_loss_kernel(StdArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm_d[2 * kWarps];
  __shared__ float sm_c[8];
  const PT* __restrict__ pred = static_cast<const PT*>(a.pred);
  const float* __restrict__ targ = a.targ;
  PT* __restrict__ grad = static_cast<PT*>(a.grad);
  const unsigned HW = a.hw;
  const unsigned npx = static_cast<unsigned>(a.n_img) * HW;
  const unsigned tid = blockIdx.x * kBlock + threadIdx.x, nthr = gridDim.x * kBlock;
  const int flags = a.flags;
  constexpr int d0 = (C == 10) ? 8 : 16, d1 = C;                // depth channels (base_module.py:137)
  constexpr unsigned PX = (C == 10) ? 2u : 1u;                  // pixels in flight per thread

  Ws ws = ws_view(a.ws);
  unsigned epoch;
  const int par = coop_prologue(ws, epoch);
  double* gacc = ws.gacc + par * kGacc;

  // ---------------- phase A: totals ----------------------------------------------------------------------
  {
    double acc[A_COUNT];
#pragma unroll
    for (int q = 0; q < A_COUNT; ++q) acc[q] = 0.0;
    for (unsigned px0 = tid; px0 < npx; px0 += PX * nthr) {
      float pva[PX][C], tva[PX][C];                              // all 2C loads of PX pixels are requested before any use
      bool m1a[PX], ona[PX];
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        const unsigned px = px0 + static_cast<unsigned>(u) * nthr;
        ona[u] = px < npx;
        const unsigned pxc = ona[u] ? px : px0;
        const unsigned b = pxc / HW, pix = pxc - b * HW;
        const size_t base = static_cast<size_t>(b) * C * HW + pix;
        m1a[u] = __ldg(a.alpha + static_cast<size_t>(b) * a.alpha_stride + pix) > 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          pva[u][c] = Elem<PT>::ld1(pred + base + static_cast<size_t>(c) * HW);
          tva[u][c] = __ldg(targ + base + static_cast<size_t>(c) * HW);
        }
      }
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        if (!ona[u]) continue;
        const bool m1 = m1a[u];
        const float (&pv)[C] = pva[u];
        const float (&tv)[C] = tva[u];
        float cabs = 0.f, csq = 0.f, aabs = 0.f, asq = 0.f, dabs = 0.f, dsq = 0.f, sd = 0.f, sdd = 0.f, nd = 0.f, ns = 0.f;
        float pf[3], pb[3], tf[3], tb[3];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float p = pv[c], t = tv[c];
          const float diff = p - t;
          if (c < 3) { pf[c] = p; tf[c] = t; }
          if (c >= 4 && c < 7) { pb[c - 4] = p; tb[c - 4] = t; }
          if (m1) {
            aabs += fabsf(diff); asq = fmaf(diff, diff, asq);
            if (c < 8) { cabs += fabsf(diff); csq = fmaf(diff, diff, csq); }
          }
          if (c >= d0 && c < d1 && t > 0.f) {                      // maskD                      base_module.py:138
            nd += 1.f; dabs += fabsf(diff); dsq = fmaf(diff, diff, dsq);
            if ((flags & ST_SILOG) && t > 1e-2f) {                 // silog's own mask           criteria.py:729
              const float dl = logf(p) - logf(t);
              ns += 1.f; sd += dl; sdd = fmaf(dl, dl, sdd);
            }
          }
        }
        if (m1) {
          acc[A_N1] += 1.0;
          acc[A_CABS] += static_cast<double>(cabs); acc[A_CSQ] += static_cast<double>(csq);
          acc[A_AABS] += static_cast<double>(aabs); acc[A_ASQ] += static_cast<double>(asq);
          if (flags & ST_FBDIV) acc[A_FB] += static_cast<double>(fb_term(pf, tb).f + fb_term(pb, tf).f);
        }
        acc[A_ND] += static_cast<double>(nd); acc[A_DABS] += static_cast<double>(dabs); acc[A_DSQ] += static_cast<double>(dsq);
        acc[A_NS] += static_cast<double>(ns); acc[A_SD] += static_cast<double>(sd); acc[A_SDD] += static_cast<double>(sdd);
      }
    }
#pragma unroll
    for (int q = 0; q < A_COUNT; q += 2) {
      const double pair[2] = {acc[q], acc[q + 1]};
      const double tot = block_sum<2>(pair, sm_d);
      if (threadIdx.x < 2 && tot != 0.0) atomicAdd(&gacc[q + threadIdx.x], tot);   // a NaN total is != 0 and is added
    }
  }
  grid.sync();

  // ---------------- totals -> loss terms and gradient coefficients -------------------------------------------
  if (threadIdx.x == 0) {
    const double n1 = __ldcg(&gacc[A_N1]), nD = __ldcg(&gacc[A_ND]), ns = __ldcg(&gacc[A_NS]);
    const double N8 = 8.0 * n1, NC = static_cast<double>(C) * n1;
    const double gs = static_cast<double>(a.grad_scale), dw = static_cast<double>(a.depth_w);
    double t_sil = 0.0, t_col = 0.0, t_mse = 0.0, t_mae = 0.0, t_fb = 0.0, total = 0.0;
    float k_sil = 0.f, k_mean = 0.f;
    if (flags & ST_SILOG) {
      const double mean = __ldcg(&gacc[A_SD]) / ns, q = __ldcg(&gacc[A_SDD]) / ns;
      const double lam = static_cast<double>(a.lambda);
      const double s = sqrt(q - lam * mean * mean);
      double val = 10.0 * s;                                  // criteria.py:732
      const bool ok = (val == val) && (val <= 3.4028234663852886e38);
      if (val != val) val = 0.0;                              // nan_to_num             base_module.py:126-127
      else if (val > 3.4028234663852886e38) val = 3.4028234663852886e38;
      t_sil = dw * val;
      total += t_sil;
      if (ok && s > 0.0) {                                    // d/dp = dw * 10 / (s n) * (d - lam mean) / p
        k_sil = static_cast<float>(gs * dw * 10.0 / (s * ns));
        k_mean = static_cast<float>(lam * mean);
      }
    }
    if (flags & ST_CMAE) { t_col = __ldcg(&gacc[A_CABS]) / N8; total += t_col; }
    if (flags & ST_CMSE) { t_col = __ldcg(&gacc[A_CSQ]) / N8; total += t_col; }
    if (flags & ST_ALLMSE) { t_mse = __ldcg(&gacc[A_ASQ]) / NC + dw * (__ldcg(&gacc[A_DSQ]) / nD); total += t_mse; }
    if (flags & ST_ALLMAE) { t_mae = __ldcg(&gacc[A_AABS]) / NC + dw * (__ldcg(&gacc[A_DABS]) / nD); total += t_mae; }
    if (flags & ST_FBDIV) { t_fb = static_cast<double>(a.fbdiv_w) * (__ldcg(&gacc[A_FB]) / n1); total += t_fb; }
    sm_c[0] = k_sil; sm_c[1] = k_mean;
    sm_c[2] = static_cast<float>(gs / N8);                    // colour terms (x 2 for mse)
    sm_c[3] = static_cast<float>(gs / NC);                    // all-channel terms
    sm_c[4] = static_cast<float>(gs * dw / nD);               // depth part of the all-channel terms
    sm_c[5] = static_cast<float>(gs * static_cast<double>(a.fbdiv_w) / n1);
    if (blockIdx.x == 0) {
      a.out[0] = static_cast<float>(total); a.out[1] = static_cast<float>(t_sil); a.out[2] = static_cast<float>(t_col);
      a.out[3] = static_cast<float>(t_mse); a.out[4] = static_cast<float>(t_mae); a.out[5] = static_cast<float>(t_fb);
      a.out[6] = static_cast<float>(n1); a.out[7] = static_cast<float>(nD);
      ws.hdr->epoch = epoch + 1u;
    }
  }
  __syncthreads();
  if (grad == nullptr) return;

  // ---------------- phase B: gradient of every term --------------------------------------------------------
  const float k_sil = sm_c[0], k_mean = sm_c[1], k_col = sm_c[2], k_all = sm_c[3], k_dep = sm_c[4], k_fb = sm_c[5];
  for (unsigned px0 = tid; px0 < npx; px0 += PX * nthr) {
    float pva[PX][C], tva[PX][C];
    bool m1a[PX], ona[PX];
    size_t basea[PX];
#pragma unroll
    for (int u = 0; u < PX; ++u) {
      const unsigned px = px0 + static_cast<unsigned>(u) * nthr;
      ona[u] = px < npx;
      const unsigned pxc = ona[u] ? px : px0;
      const unsigned b = pxc / HW, pix = pxc - b * HW;
      basea[u] = static_cast<size_t>(b) * C * HW + pix;
      m1a[u] = __ldg(a.alpha + static_cast<size_t>(b) * a.alpha_stride + pix) > 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        pva[u][c] = Elem<PT>::ld1(pred + basea[u] + static_cast<size_t>(c) * HW);
        tva[u][c] = __ldg(targ + basea[u] + static_cast<size_t>(c) * HW);
      }
    }
#pragma unroll
    for (int u = 0; u < PX; ++u) {
      if (!ona[u]) continue;
      const bool m1 = m1a[u];
      const size_t base = basea[u];
      const float (&pv)[C] = pva[u];
      const float (&tv)[C] = tva[u];
      float g[C];
      float pf[3], pb[3], tf[3], tb[3];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float p = pv[c], t = tv[c];
        const float diff = p - t;
        if (c < 3) { pf[c] = p; tf[c] = t; }
        if (c >= 4 && c < 7) { pb[c - 4] = p; tb[c - 4] = t; }
        float gc = 0.f;
        if (m1) {
          if (c < 8) {
            if (flags & ST_CMAE) gc = fmaf(k_col, sgn0(diff), gc);
            if (flags & ST_CMSE) gc = fmaf(2.f * k_col, diff, gc);
          }
          if (flags & ST_ALLMSE) gc = fmaf(2.f * k_all, diff, gc);
          if (flags & ST_ALLMAE) gc = fmaf(k_all, sgn0(diff), gc);
        }
        if (c >= d0 && c < d1 && t > 0.f) {
          if (flags & ST_ALLMSE) gc = fmaf(2.f * k_dep, diff, gc);
          if (flags & ST_ALLMAE) gc = fmaf(k_dep, sgn0(diff), gc);
          if ((flags & ST_SILOG) && t > 1e-2f) gc += __fdividef(k_sil * ((logf(p) - logf(t)) - k_mean), p);
        }
        g[c] = gc;
      }
      if ((flags & ST_FBDIV) && m1) {
        const FbTerm f1 = fb_term(pf, tb), f2 = fb_term(pb, tf);
        const float r1 = (f1.np > 0.f) ? f1.dot * f1.nt / (f1.mag * f1.mag * f1.np) : 0.f;
        const float r2 = (f2.np > 0.f) ? f2.dot * f2.nt / (f2.mag * f2.mag * f2.np) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          g[c] = fmaf(k_fb, tb[c] / f1.mag - r1 * pf[c], g[c]);
          g[4 + c] = fmaf(k_fb, tf[c] / f2.mag - r2 * pb[c], g[4 + c]);
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) Elem<PT>::st1(grad + base + static_cast<size_t>(c) * HW, g[c]);
    }
  }
}

template <typename PT, int C>
int launch_stdepth(StdArgs& a, cudaStream_t st) {
  const void* fn = reinterpret_cast<const void*>(&stdepth_loss_kernel<PT, C>);
  const int64_t n = static_cast<int64_t>(a.n_img) * a.hw;
  int64_t grid = (n + kBlock - 1) / kBlock;
  const int cap = coop_grid(fn, kBlock, 0);
  if (cap <= 0) return MDE_ECUDA;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  void* args[] = {&a};
  MDE_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(static_cast<unsigned>(grid)), dim3(kBlock), args, 0, st));
  count_launch();
  return MDE_OK;
}

template <typename PT>
int launch_stdepth_c(StdArgs& a, int64_t C, cudaStream_t st) {
  return (C == 10) ? launch_stdepth<PT, 10>(a, st) : launch_stdepth<PT, 20>(a, st);
}

}  // namespace
}  // namespace mde

extern "C" int mde_stdepth_loss(const void* pred, int pred_dtype, const float* targ, const float* rgba, int64_t rgba_c,
                                int64_t n_img, int64_t C, int64_t hw, int flags, float depth_w, float fbdiv_w,
                                float variance_focus, float grad_scale, void* ws, float* out8, void* grad, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && targ && rgba && ws && out8, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(C == 10 || C == 20, MDE_EINVAL, "C must be 10 (single layer, depth 8:10) or 20 (three layers, depth 16:20)");
  MDE_REQUIRE(rgba_c >= 4, MDE_EINVAL, "rgba needs an alpha plane (channel 3)");
  MDE_REQUIRE(n_img * C * hw < (int64_t(1) << 31), MDE_ETOOBIG, "more than 2^31 elements");
  MDE_REQUIRE(flags > 0 && flags < 64, MDE_EINVAL, "flags: bit set of SILOG 1, CMAE 2, CMSE 4, ALLMSE 8, ALLMAE 16, FBDIV 32");
  StdArgs a;
  a.pred = pred; a.targ = targ; a.alpha = rgba + 3 * hw; a.alpha_stride = rgba_c * hw;
  a.n_img = static_cast<int>(n_img); a.hw = static_cast<unsigned>(hw);
  a.flags = flags; a.depth_w = depth_w; a.fbdiv_w = fbdiv_w; a.lambda = variance_focus; a.grad_scale = grad_scale;
  a.ws = ws; a.out = out8; a.grad = grad;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (pred_dtype) {
    case MDE_F32: return launch_stdepth_c<float>(a, C, st);
    case MDE_F16: return launch_stdepth_c<__half>(a, C, st);
    case MDE_BF16: return launch_stdepth_c<__nv_bfloat16>(a, C, st);
    default: set_error("mde_stdepth_loss: unknown pred_dtype %d", pred_dtype); return MDE_EINVAL;
  }
}
