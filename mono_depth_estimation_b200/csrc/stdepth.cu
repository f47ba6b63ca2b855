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
// consecutive pixels of a plane: full 128-byte lines). Algorithmic traffic: (4C + 4C + 4) read + 4C written per pixel.
//   forward only   one sweep reduces the 12 totals, a grid sync publishes them, CTA 0 forms the terms.
//   with gradient  the gradient coefficients depend on COUNTS and on the SILog sums only - i.e. on the alpha plane and
//                  on the D depth channels of pred / targ (20 of the 84 B/px read at C = 10). Phase A reads just those,
//                  a grid sync publishes them, and phase B is the ONE sweep over all channels: it writes the gradient of
//                  every term and accumulates the colour / all-channel / front-back sums of the loss VALUE on the way;
//                  the last CTA to finish (ticket) forms the terms. 144 B/px moved instead of 208 (round 1: both phases
//                  read everything), 124 compulsory.
#include <cstdlib>

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

  // loss terms from the totals of parity set `gacc`; thread 0 of one CTA. Returns the SILog coefficients through k_*.
  auto terms = [&](bool write_out, float& k_sil, float& k_mean, double& n1_o, double& nD_o) {
    const double n1 = __ldcg(&gacc[A_N1]), nD = __ldcg(&gacc[A_ND]), ns = __ldcg(&gacc[A_NS]);
    const double N8 = 8.0 * n1, NC = static_cast<double>(C) * n1;
    const double gs = static_cast<double>(a.grad_scale), dw = static_cast<double>(a.depth_w);
    double t_sil = 0.0, t_col = 0.0, t_mse = 0.0, t_mae = 0.0, t_fb = 0.0, total = 0.0;
    k_sil = 0.f; k_mean = 0.f;
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
    n1_o = n1; nD_o = nD;
    if (!write_out) return;
    if (flags & ST_CMAE) { t_col = __ldcg(&gacc[A_CABS]) / N8; total += t_col; }
    if (flags & ST_CMSE) { t_col = __ldcg(&gacc[A_CSQ]) / N8; total += t_col; }
    if (flags & ST_ALLMSE) { t_mse = __ldcg(&gacc[A_ASQ]) / NC + dw * (__ldcg(&gacc[A_DSQ]) / nD); total += t_mse; }
    if (flags & ST_ALLMAE) { t_mae = __ldcg(&gacc[A_AABS]) / NC + dw * (__ldcg(&gacc[A_DABS]) / nD); total += t_mae; }
    if (flags & ST_FBDIV) { t_fb = static_cast<double>(a.fbdiv_w) * (__ldcg(&gacc[A_FB]) / n1); total += t_fb; }
    a.out[0] = static_cast<float>(total); a.out[1] = static_cast<float>(t_sil); a.out[2] = static_cast<float>(t_col);
    a.out[3] = static_cast<float>(t_mse); a.out[4] = static_cast<float>(t_mae); a.out[5] = static_cast<float>(t_fb);
    a.out[6] = static_cast<float>(n1); a.out[7] = static_cast<float>(nD);
    ws.hdr->epoch = epoch + 1u;
  };

  if (grad == nullptr) {
      // ---------------- forward only: one sweep, all totals ---------------------------------------------------
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

    if (blockIdx.x == 0 && threadIdx.x == 0) {
      float k0, k1;
      double n1, nD;
      terms(true, k0, k1, n1, nD);
    }
    return;
  }

  // ---------------- phase A: what the gradient coefficients need (alpha plane + depth channels only) ---------
  {
    constexpr unsigned PXA = 4u;                                  // pixels in flight per thread (1 + 2 D loads each)
    constexpr int D = d1 - d0;
    double acc[A_COUNT];
#pragma unroll
    for (int q = 0; q < A_COUNT; ++q) acc[q] = 0.0;
    for (unsigned px0 = tid; px0 < npx; px0 += PXA * nthr) {
      float pva[PXA][D], tva[PXA][D], al[PXA];
      bool ona[PXA];
#pragma unroll
      for (int u = 0; u < PXA; ++u) {
        const unsigned px = px0 + static_cast<unsigned>(u) * nthr;
        ona[u] = px < npx;
        const unsigned pxc = ona[u] ? px : px0;
        const unsigned b = pxc / HW, pix = pxc - b * HW;
        const size_t base = static_cast<size_t>(b) * C * HW + pix;
        al[u] = __ldg(a.alpha + static_cast<size_t>(b) * a.alpha_stride + pix);
#pragma unroll
        for (int c = 0; c < D; ++c) {
          pva[u][c] = Elem<PT>::ld1(pred + base + static_cast<size_t>(d0 + c) * HW);
          tva[u][c] = __ldg(targ + base + static_cast<size_t>(d0 + c) * HW);
        }
      }
#pragma unroll
      for (int u = 0; u < PXA; ++u) {
        if (!ona[u]) continue;
        float dabs = 0.f, dsq = 0.f, sd = 0.f, sdd = 0.f, nd = 0.f, ns = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const float p = pva[u][c], t = tva[u][c];
          const float diff = p - t;
          if (t > 0.f) {                                           // maskD                      base_module.py:138
            nd += 1.f; dabs += fabsf(diff); dsq = fmaf(diff, diff, dsq);
            if ((flags & ST_SILOG) && t > 1e-2f) {                 // silog's own mask           criteria.py:729
              const float dl = logf(p) - logf(t);
              ns += 1.f; sd += dl; sdd = fmaf(dl, dl, sdd);
            }
          }
        }
        if (al[u] > 0.f) acc[A_N1] += 1.0;
        acc[A_ND] += static_cast<double>(nd); acc[A_DABS] += static_cast<double>(dabs); acc[A_DSQ] += static_cast<double>(dsq);
        acc[A_NS] += static_cast<double>(ns); acc[A_SD] += static_cast<double>(sd); acc[A_SDD] += static_cast<double>(sdd);
      }
    }
    {
      const double pair[2] = {acc[A_N1], 0.0};
      const double tot = block_sum<2>(pair, sm_d);
      if (threadIdx.x == 0 && tot != 0.0) atomicAdd(&gacc[A_N1], tot);
    }
    static_assert(A_ND == 5 && A_SDD == 10, "the depth totals are slots 5..10");
#pragma unroll
    for (int q = A_ND; q <= A_SDD; q += 2) {
      const double pair[2] = {acc[q], acc[q + 1]};
      const double tot = block_sum<2>(pair, sm_d);
      if (threadIdx.x < 2 && tot != 0.0) atomicAdd(&gacc[q + threadIdx.x], tot);   // a NaN total is != 0 and is added
    }
  }
  grid.sync();

  // ---------------- counts and SILog sums -> gradient coefficients -----------------------------------------------
  if (threadIdx.x == 0) {
    float k_sil, k_mean;
    double n1, nD;
    terms(false, k_sil, k_mean, n1, nD);
    const double N8 = 8.0 * n1, NC = static_cast<double>(C) * n1;
    const double gs = static_cast<double>(a.grad_scale), dw = static_cast<double>(a.depth_w);
    sm_c[0] = k_sil; sm_c[1] = k_mean;
    sm_c[2] = static_cast<float>(gs / N8);                    // colour terms (x 2 for mse)
    sm_c[3] = static_cast<float>(gs / NC);                    // all-channel terms
    sm_c[4] = static_cast<float>(gs * dw / nD);               // depth part of the all-channel terms
    sm_c[5] = static_cast<float>(gs * static_cast<double>(a.fbdiv_w) / n1);
  }
  __syncthreads();

  // ---------------- phase B: the one sweep over all channels: gradient of every term + the value sums -------------
  const float k_sil = sm_c[0], k_mean = sm_c[1], k_col = sm_c[2], k_all = sm_c[3], k_dep = sm_c[4], k_fb = sm_c[5];
  double v_cabs = 0.0, v_csq = 0.0, v_aabs = 0.0, v_asq = 0.0, v_fb = 0.0;
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
      float cabs = 0.f, csq = 0.f, aabs = 0.f, asq = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float p = pv[c], t = tv[c];
        const float diff = p - t;
        if (c < 3) { pf[c] = p; tf[c] = t; }
        if (c >= 4 && c < 7) { pb[c - 4] = p; tb[c - 4] = t; }
        float gc = 0.f;
        if (m1) {
          aabs += fabsf(diff); asq = fmaf(diff, diff, asq);
          if (c < 8) { cabs += fabsf(diff); csq = fmaf(diff, diff, csq); }
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
      if (m1) {
        v_cabs += static_cast<double>(cabs); v_csq += static_cast<double>(csq);
        v_aabs += static_cast<double>(aabs); v_asq += static_cast<double>(asq);
      }
      if ((flags & ST_FBDIV) && m1) {
        const FbTerm f1 = fb_term(pf, tb), f2 = fb_term(pb, tf);
        v_fb += static_cast<double>(f1.f + f2.f);
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
  // value sums of this sweep -> totals; the last CTA to arrive forms the loss terms
  {
    const double p0[2] = {v_cabs, v_csq}, p1[2] = {v_aabs, v_asq}, p2[2] = {v_fb, 0.0};
    const double t0 = block_sum<2>(p0, sm_d);
    if (threadIdx.x < 2 && t0 != 0.0) atomicAdd(&gacc[A_CABS + threadIdx.x], t0);
    const double t1 = block_sum<2>(p1, sm_d);
    if (threadIdx.x < 2 && t1 != 0.0) atomicAdd(&gacc[A_AABS + threadIdx.x], t1);
    const double t2 = block_sum<2>(p2, sm_d);
    if (threadIdx.x == 0 && t2 != 0.0) atomicAdd(&gacc[A_FB], t2);
  }
  static_assert(A_CSQ == A_CABS + 1 && A_ASQ == A_AABS + 1, "value sums are stored in pairs");
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ws.hdr->ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      float k0, k1;
      double n1, nD;
      terms(true, k0, k1, n1, nD);
      ws.hdr->ticket = 0u;
    }
  }
}

// ---- the same criterion on QUADS of a channel plane (hw % 4 == 0, 16-byte aligned tensors) ----------------------------
// Every term but 'fbdivergence' is elementwise given the pixel's alpha (and that one needs only channels 0-2 and 4-6 of
// the pixel, which sit in the first channel group's registers: FB = true): a thread owns 4 consecutive pixels and walks the
// channel planes with 128-bit accesses - a warp touches 512 contiguous bytes per plane visit instead of 128 (the
// write-locality finding of DESIGN 4.11: 5.1 instead of 3.0 TB/s on a plane-strided store pattern) and executes a quarter
// of the memory instructions. Channels go through in groups of <= 10 (all loads of a group requested before the first is
// used: 20 x 16 B per thread in flight). Same two phases as the scalar kernel: counts + SILog sums from the alpha plane
// and the depth channels, grid sync, then ONE sweep that writes every gradient and accumulates the value sums.
template <typename PT, int C, bool FB>
__global__ void __launch_bounds__(kBlock, 1) stdepth_vec_kernel(StdArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm_d[2 * kWarps];
  __shared__ float sm_c[8];
  const PT* __restrict__ pred = static_cast<const PT*>(a.pred);
  const float* __restrict__ targ = a.targ;
  PT* __restrict__ grad = static_cast<PT*>(a.grad);
  const unsigned HW = a.hw, HQ = HW >> 2;
  const unsigned nq = static_cast<unsigned>(a.n_img) * HQ;
  const unsigned tid = blockIdx.x * kBlock + threadIdx.x, nthr = gridDim.x * kBlock;
  const int flags = a.flags;
  constexpr int d0 = (C == 10) ? 8 : 16, d1 = C, D = d1 - d0;

  Ws ws = ws_view(a.ws);
  unsigned epoch;
  const int par = coop_prologue(ws, epoch);
  double* gacc = ws.gacc + par * kGacc;
  const double gs = static_cast<double>(a.grad_scale), dw = static_cast<double>(a.depth_w);

  auto comps = [](const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; };

  // ---------------- phase A: alpha plane + depth channels -> counts, depth sums, SILog sums ----------------------
  {
    double acc[A_COUNT];
#pragma unroll
    for (int q = 0; q < A_COUNT; ++q) acc[q] = 0.0;
    for (unsigned q = tid; q < nq; q += nthr) {
      const unsigned b = q / HQ, pix = (q - b * HQ) << 2;
      const size_t base = static_cast<size_t>(b) * C * HW + pix;
      const float4 al4 = __ldcs(reinterpret_cast<const float4*>(a.alpha + static_cast<size_t>(b) * a.alpha_stride + pix));
      float4 p4[D], t4[D];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        p4[c] = Elem<PT>::template ld4<true>(pred + base + static_cast<size_t>(d0 + c) * HW);
        t4[c] = Elem<float>::template ld4<true>(targ + base + static_cast<size_t>(d0 + c) * HW);
      }
      float dabs = 0.f, dsq = 0.f, sd = 0.f, sdd = 0.f, nd = 0.f, ns = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        float pv[4], tv[4];
        comps(p4[c], pv); comps(t4[c], tv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p = pv[j], t = tv[j], diff = p - t;
          if (t > 0.f) {                                           // maskD                      base_module.py:138
            nd += 1.f; dabs += fabsf(diff); dsq = fmaf(diff, diff, dsq);
            if ((flags & ST_SILOG) && t > 1e-2f) {                 // silog's own mask           criteria.py:729
              const float dl = logf(p) - logf(t);
              ns += 1.f; sd += dl; sdd = fmaf(dl, dl, sdd);
            }
          }
        }
      }
      acc[A_N1] += static_cast<double>((al4.x > 0.f ? 1.f : 0.f) + (al4.y > 0.f ? 1.f : 0.f) + (al4.z > 0.f ? 1.f : 0.f) + (al4.w > 0.f ? 1.f : 0.f));
      acc[A_ND] += static_cast<double>(nd); acc[A_DABS] += static_cast<double>(dabs); acc[A_DSQ] += static_cast<double>(dsq);
      acc[A_NS] += static_cast<double>(ns); acc[A_SD] += static_cast<double>(sd); acc[A_SDD] += static_cast<double>(sdd);
    }
    {
      const double pair[2] = {acc[A_N1], 0.0};
      const double tot = block_sum<2>(pair, sm_d);
      if (threadIdx.x == 0 && tot != 0.0) atomicAdd(&gacc[A_N1], tot);
    }
#pragma unroll
    for (int q = A_ND; q <= A_SDD; q += 2) {
      const double pair[2] = {acc[q], acc[q + 1]};
      const double tot = block_sum<2>(pair, sm_d);
      if (threadIdx.x < 2 && tot != 0.0) atomicAdd(&gacc[q + threadIdx.x], tot);   // a NaN total is != 0 and is added
    }
  }
  grid.sync();

  // loss terms from the totals (thread 0 of one CTA); the SILog coefficients come back through k_*
  auto terms = [&](bool write_out, float& k_sil, float& k_mean, double& n1_o, double& nD_o) {
    const double n1 = __ldcg(&gacc[A_N1]), nD = __ldcg(&gacc[A_ND]), ns = __ldcg(&gacc[A_NS]);
    const double N8 = 8.0 * n1, NC = static_cast<double>(C) * n1;
    double t_sil = 0.0, t_col = 0.0, t_mse = 0.0, t_mae = 0.0, total = 0.0;
    k_sil = 0.f; k_mean = 0.f;
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
      if (ok && s > 0.0) {
        k_sil = static_cast<float>(gs * dw * 10.0 / (s * ns));
        k_mean = static_cast<float>(lam * mean);
      }
    }
    n1_o = n1; nD_o = nD;
    if (!write_out) return;
    if (flags & ST_CMAE) { t_col = __ldcg(&gacc[A_CABS]) / N8; total += t_col; }
    if (flags & ST_CMSE) { t_col = __ldcg(&gacc[A_CSQ]) / N8; total += t_col; }
    if (flags & ST_ALLMSE) { t_mse = __ldcg(&gacc[A_ASQ]) / NC + dw * (__ldcg(&gacc[A_DSQ]) / nD); total += t_mse; }
    if (flags & ST_ALLMAE) { t_mae = __ldcg(&gacc[A_AABS]) / NC + dw * (__ldcg(&gacc[A_DABS]) / nD); total += t_mae; }
    double t_fb = 0.0;
    if (FB && (flags & ST_FBDIV)) { t_fb = static_cast<double>(a.fbdiv_w) * (__ldcg(&gacc[A_FB]) / n1); total += t_fb; }
    a.out[0] = static_cast<float>(total); a.out[1] = static_cast<float>(t_sil); a.out[2] = static_cast<float>(t_col);
    a.out[3] = static_cast<float>(t_mse); a.out[4] = static_cast<float>(t_mae); a.out[5] = static_cast<float>(t_fb);
    a.out[6] = static_cast<float>(n1); a.out[7] = static_cast<float>(nD);
    ws.hdr->epoch = epoch + 1u;
  };
  if (threadIdx.x == 0) {
    float k_sil, k_mean;
    double n1, nD;
    terms(false, k_sil, k_mean, n1, nD);
    sm_c[0] = k_sil; sm_c[1] = k_mean;
    sm_c[2] = static_cast<float>(gs / (8.0 * n1));                       // colour terms (x 2 for mse)
    sm_c[3] = static_cast<float>(gs / (static_cast<double>(C) * n1));   // all-channel terms
    sm_c[4] = static_cast<float>(gs * dw / nD);                          // depth part of the all-channel terms
    sm_c[5] = static_cast<float>(gs * static_cast<double>(a.fbdiv_w) / n1);
  }
  __syncthreads();

  // ---------------- phase B: one sweep over all channels ----------------------------------------------------------
  // The loss-name flags are folded into four per-class coefficients up front, so that the element code carries no flag
  // test (the first version spent 62 instructions per element, most of them uniform flag tests and their selects):
  //   colour channel (c < 8), mask1:   g = a_col sgn(d) + b_col d        other channel, mask1:  g = a_oth sgn(d) + b_oth d
  //   depth channel, maskD:            g += a_dep sgn(d) + b_dep d  (+ the SILog term)
  // A masked-out element enters as d = 0 (a select, not a product: a NaN there must not reach the sums).
  const float k_sil = sm_c[0], k_mean = sm_c[1], k_col = sm_c[2], k_all = sm_c[3], k_dep = sm_c[4];
  const float a_all = (flags & ST_ALLMAE) ? k_all : 0.f, b_all = (flags & ST_ALLMSE) ? 2.f * k_all : 0.f;
  const float a_col = a_all + ((flags & ST_CMAE) ? k_col : 0.f), b_col = b_all + ((flags & ST_CMSE) ? 2.f * k_col : 0.f);
  const float a_dep = (flags & ST_ALLMAE) ? k_dep : 0.f, b_dep = (flags & ST_ALLMSE) ? 2.f * k_dep : 0.f;
  const bool sil = (flags & ST_SILOG) != 0, dep_terms = (flags & (ST_ALLMAE | ST_ALLMSE)) != 0;
  auto sgn_times = [](float k, float d) { return (d > 0.f) ? k : ((d < 0.f) ? -k : 0.f); };
  const float k_fb = sm_c[5];
  double v_cabs = 0.0, v_csq = 0.0, v_aabs = 0.0, v_asq = 0.0, v_fb = 0.0;
  constexpr int GRP = 10;                                         // channels per group (C = 10: one, C = 20: two)
  for (unsigned q = tid; q < nq; q += nthr) {
    const unsigned b = q / HQ, pix = (q - b * HQ) << 2;
    const size_t base = static_cast<size_t>(b) * C * HW + pix;
    const float4 al4 = __ldcs(reinterpret_cast<const float4*>(a.alpha + static_cast<size_t>(b) * a.alpha_stride + pix));
    float al[4];
    comps(al4, al);
    const bool m1[4] = {al[0] > 0.f, al[1] > 0.f, al[2] > 0.f, al[3] > 0.f};   // mask1            base_module.py:133
    float cabs = 0.f, csq = 0.f, oabs = 0.f, osq = 0.f;            // colour channels / the other channels, mask1 pixels
#pragma unroll
    for (int g0 = 0; g0 < C; g0 += GRP) {
      float4 p4[GRP], t4[GRP];
#pragma unroll
      for (int c = 0; c < GRP; ++c) {
        p4[c] = Elem<PT>::template ld4<false>(pred + base + static_cast<size_t>(g0 + c) * HW);
        t4[c] = Elem<float>::template ld4<false>(targ + base + static_cast<size_t>(g0 + c) * HW);
      }
      // front/back term (base_module.py:184-194): per pixel four scalars from channels 0-2 / 4-6 of this group
      float fi1[4], fr1[4], fi2[4], fr2[4];
      if (FB && g0 == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          auto at = [&](const float4 (&v)[GRP], int c) { return j == 0 ? v[c].x : (j == 1 ? v[c].y : (j == 2 ? v[c].z : v[c].w)); };
          const float pf[3] = {at(p4, 0), at(p4, 1), at(p4, 2)}, pb[3] = {at(p4, 4), at(p4, 5), at(p4, 6)};
          const float tf[3] = {at(t4, 0), at(t4, 1), at(t4, 2)}, tb[3] = {at(t4, 4), at(t4, 5), at(t4, 6)};
          const FbTerm f1 = fb_term(pf, tb), f2 = fb_term(pb, tf);
          const bool on = m1[j] && (flags & ST_FBDIV);
          fi1[j] = on ? k_fb / f1.mag : 0.f;
          fr1[j] = (on && f1.np > 0.f) ? k_fb * (f1.dot * f1.nt / (f1.mag * f1.mag * f1.np)) : 0.f;
          fi2[j] = on ? k_fb / f2.mag : 0.f;
          fr2[j] = (on && f2.np > 0.f) ? k_fb * (f2.dot * f2.nt / (f2.mag * f2.mag * f2.np)) : 0.f;
          if (on) v_fb += static_cast<double>(f1.f + f2.f);
        }
      }
#pragma unroll
      for (int cc = 0; cc < GRP; ++cc) {
        const int c = g0 + cc;
        float pv[4], tv[4], gv[4];
        comps(p4[cc], pv); comps(t4[cc], tv);
        float ov[4] = {0.f, 0.f, 0.f, 0.f};                         // the front/back partner channel of targ (c < 3: 4 + c, 4 <= c < 7: c - 4)
        if (FB && g0 == 0 && (cc < 3 || (cc >= 4 && cc < 7))) comps(t4[cc < 3 ? cc + 4 : cc - 4], ov);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p = pv[j], t = tv[j], diff = p - t;
          const float dm = m1[j] ? diff : 0.f;
          float gc;
          if (c < 8) {
            cabs += fabsf(dm); csq = fmaf(dm, dm, csq);
            gc = fmaf(b_col, dm, sgn_times(a_col, dm));
          } else {
            oabs += fabsf(dm); osq = fmaf(dm, dm, osq);
            gc = fmaf(b_all, dm, sgn_times(a_all, dm));
          }
          if (c >= d0 && c < d1) {                                 // (compile-time) depth channel
            const bool vd = t > 0.f;                               // maskD                      base_module.py:138
            if (dep_terms) {
              const float dd = vd ? diff : 0.f;
              gc += fmaf(b_dep, dd, sgn_times(a_dep, dd));
            }
            if (sil && vd && t > 1e-2f) gc += __fdividef(k_sil * ((logf(p) - logf(t)) - k_mean), p);
          }
          if (FB && g0 == 0 && cc < 3) gc += fmaf(fi1[j], ov[j], -fr1[j] * p);              // k_fb (tb_c / mag1 - r1 pf_c)
          if (FB && g0 == 0 && cc >= 4 && cc < 7) gc += fmaf(fi2[j], ov[j], -fr2[j] * p);   // k_fb (tf_c / mag2 - r2 pb_c)
          gv[j] = gc;
        }
        Elem<PT>::st4(grad + base + static_cast<size_t>(c) * HW, make_float4(gv[0], gv[1], gv[2], gv[3]));
      }
    }
    v_cabs += static_cast<double>(cabs); v_csq += static_cast<double>(csq);
    v_aabs += static_cast<double>(cabs + oabs); v_asq += static_cast<double>(csq + osq);
  }
  if (FB) {
    const double p2[2] = {v_fb, 0.0};
    const double t2 = block_sum<2>(p2, sm_d);
    if (threadIdx.x == 0 && t2 != 0.0) atomicAdd(&gacc[A_FB], t2);
  }
  {
    const double p0[2] = {v_cabs, v_csq}, p1[2] = {v_aabs, v_asq};
    const double t0 = block_sum<2>(p0, sm_d);
    if (threadIdx.x < 2 && t0 != 0.0) atomicAdd(&gacc[A_CABS + threadIdx.x], t0);
    const double t1 = block_sum<2>(p1, sm_d);
    if (threadIdx.x < 2 && t1 != 0.0) atomicAdd(&gacc[A_AABS + threadIdx.x], t1);
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ws.hdr->ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      float k0, k1;
      double n1, nD;
      terms(true, k0, k1, n1, nD);
      ws.hdr->ticket = 0u;
    }
  }
}

template <typename PT, int C>
int launch_stdepth(StdArgs& a, cudaStream_t st) {
  static const bool no_vec = [] { const char* e = getenv("MDE_STDEPTH_NO_VEC"); return e && atoi(e) != 0; }();
  const bool vec = !no_vec && a.grad != nullptr && (a.hw % 4 == 0) && aligned_to(a.pred, 4 * sizeof(PT)) &&
                   aligned_to(a.targ, 16) && aligned_to(a.alpha, 16) && (a.alpha_stride % 4 == 0) && aligned_to(a.grad, 4 * sizeof(PT));
  const void* fn = !vec ? reinterpret_cast<const void*>(&stdepth_loss_kernel<PT, C>)
                        : ((a.flags & ST_FBDIV) ? reinterpret_cast<const void*>(&stdepth_vec_kernel<PT, C, true>)
                                                : reinterpret_cast<const void*>(&stdepth_vec_kernel<PT, C, false>));
  const int64_t n = vec ? static_cast<int64_t>(a.n_img) * (a.hw / 4) : static_cast<int64_t>(a.n_img) * a.hw;
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
