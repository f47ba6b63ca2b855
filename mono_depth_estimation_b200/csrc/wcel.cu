// wcel.cu - the classification half of VNL's ModelLoss and the bin <-> depth maps
// either side of it.
//
//   WCEL_Loss.forward            reference criteria.py:839-863
//   VNLModule.depth_to_bins      reference modules/vnl.py:202-217
//   VNLModule.bins_to_depth      reference modules/vnl.py:219-230 (+ its backward)
//
// WCEL: loss = -sum_px sum_c W[bin_px][c] * log_softmax(z_px)[c] / #(gt > 0), W = the row-normalised
// weight matrix (criteria.py:846-848; exp(-0.2 (i-j)^2) rows in modules/vnl.py:162), a padding pixel
// (bin outside [0, C)) has an all-zero one-hot row and contributes nothing. The reference materialises
// log_softmax [B,C,H,W], a [BHW,C] one-hot, a [BHW,C] matmul and their product; here one thread owns
// one pixel and walks the C channel planes twice (the second time through L2): pass 1 is an online
// softmax (running max / rescaled sum, MUFU.EX2) fused with A = sum_c w_c z_c, pass 2 writes
//     d loss / d z_c = (softmax_c * sum(w) - w_c) / n_valid
// so HBM sees 4C B/px read + 4C B/px written (1208 B/px at C = 150 with bins and gt) instead of the
// reference's ~10 passes. Layout [n, C, hw]: a warp reads 32 consecutive pixels of one channel plane
// per load (full 128-byte lines also when hw is odd), 8 channels in flight per thread. The weight
// rows live in shared memory with an odd row stride (32 distinct bins -> 32 distinct banks); a table
// that does not fit (C > 230) is read through L1/L2 instead.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "metric_math.cuh"

namespace mde {
namespace {

constexpr int kWBlock = 512;
constexpr int kWWarps = kWBlock / 32;
constexpr int kWUnroll = 8;

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T>
__device__ __forceinline__ float ld_x(const T* p);
template <>
__device__ __forceinline__ float ld_x<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_x<__half>(const __half* p) { return Elem<__half>::ld1(p); }
template <>
__device__ __forceinline__ float ld_x<__nv_bfloat16>(const __nv_bfloat16* p) { return Elem<__nv_bfloat16>::ld1(p); }

// block sum of one double -> atomicAdd to *dst (thread 0)
__device__ __forceinline__ void publish_one_w(double v, double* dst, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += sm[w];
    if (t != 0.0) atomicAdd(dst, t);
  }
}

// pass 0: n_valid = #(gt > 0) (criteria.py:861) into tacc[10]
__global__ void __launch_bounds__(256) wcel_count_kernel(const float* __restrict__ gt, int64_t npx, void* ws_raw) {
  __shared__ double sm[8];
  double c = 0.0;
  for (int64_t px = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; px < npx;
       px += static_cast<int64_t>(gridDim.x) * 256)
    c += (__ldg(gt + px) > 0.f) ? 1.0 : 0.0;
  Ws ws = ws_view(ws_raw);
  publish_one_w(c, &ws.hdr->tacc[10], sm);
}

struct WcelArgs {
  const void* x;          // logits [n, C, hw]
  const int* bins;        // [n, hw] int32
  const float* weight;    // [C, C] row-normalised, row = gt bin
  const float* rowsum;    // [C] sum of each (fp32) row
  int64_t n, hw;
  int C;
  float grad_scale;
  void* ws;
  float* loss_out;
  void* grad;             // [n, C, hw], dtype of x; nullable
};

// U = channels in flight per thread, PER_SM = resident CTAs per SM.
// Measured at C4 size (8 x 150 x 385 x 385, profiles/r01_bench_all_final.jsonl): the pixels in flight between
// the two passes are 148 x 32 warps x 19 KB = 91 MB (plus as many gradient bytes), so the second read of the
// logits mostly misses the 126 MB L2 and the kernel moves ~12C B/px at ~5.2 TB/s (409 us; forward only: 8C/2 at
// 4.8 TB/s). Three variants were tried and were SLOWER, so they are not kept: L2 evict_last / evict_first cache
// hints (430 us), one CTA per SM with 32 channels in flight (45 MB in flight, 464 us), and a shared-memory
// staged kernel (cp.async tile of 32 px x C per warp, 7 warps per SM beside the 91 KB table: 576 us).
// Round 2 tried the obvious way to read the logits ONCE: a pixel's 150 logits held in registers between the passes
// (75 per thread, exponentials written over the logits, one MUFU.EX2 per element instead of three), the channels split
// over two lanes of a warp (64 contiguous bytes of two planes per access: 440 us) or over two warps behind a named
// barrier (128 bytes of one plane per access: 428 us). Both move 8C + 8 B/px and both are SLOWER than this kernel's
// 12C at 410 us: 170 registers per thread leave 12 warps per SM, each of them alternating between a burst of 75 loads,
// the arithmetic and a burst of 75 stores, so about half of the ~115 KB the register file can hold is in flight on
// average - less than the latency x bandwidth product of an SM (~66 KB at 1.5 us) needs with any margin.
template <typename XT, bool HAS_GRAD, bool SMEM_TABLE, int U, int PER_SM>
__global__ void __launch_bounds__(kWBlock, PER_SM) wcel_kernel(WcelArgs a) {
  extern __shared__ float sm_w[];   // SMEM_TABLE: [C][stride] weights, then [C] row sums
  __shared__ double sm_red[kWWarps];
  __shared__ bool sm_last;
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  XT* __restrict__ gx = static_cast<XT*>(a.grad);
  const int C = a.C;
  const int stride = SMEM_TABLE ? (C | 1) : C;
  if constexpr (SMEM_TABLE) {
    for (int i = threadIdx.x; i < C * C; i += kWBlock) sm_w[(i / C) * stride + (i % C)] = __ldg(a.weight + i);
    for (int i = threadIdx.x; i < C; i += kWBlock) sm_w[C * stride + i] = __ldg(a.rowsum + i);
    __syncthreads();
  }
  Ws ws = ws_view(a.ws);
  const double n_valid = __ldcg(&ws.hdr->tacc[10]);
  const float gcoef = a.grad_scale / static_cast<float>(n_valid);
  const int64_t npx = a.n * a.hw;
  const int64_t hw = a.hw;
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.69314718055994531f;
  double loss_acc = 0.0;
  for (int64_t px = static_cast<int64_t>(blockIdx.x) * kWBlock + threadIdx.x; px < npx;
       px += static_cast<int64_t>(gridDim.x) * kWBlock) {
    const int64_t img = px / hw;
    const int64_t base = img * static_cast<int64_t>(C) * hw + (px - img * hw);
    const int b = __ldg(a.bins + px);
    const bool valid = (b >= 0) && (b < C);   // bin C+1 marks padding (modules/vnl.py:213): all-zero one-hot row
    if (!valid) {
      if constexpr (HAS_GRAD) {
        for (int c = 0; c < C; ++c) Elem<XT>::st1(gx + base + static_cast<int64_t>(c) * hw, 0.f);
      }
      continue;
    }
    const float* wrow = SMEM_TABLE ? (sm_w + b * stride) : (a.weight + static_cast<int64_t>(b) * C);
    const float rs = SMEM_TABLE ? sm_w[C * stride + b] : __ldg(a.rowsum + b);
    // ---- pass 1: online softmax + weighted logit sum ----
    float m = -INFINITY, s = 0.f, A = 0.f;
    int c = 0;
    for (; c + U <= C; c += U) {
      float z[U];
#pragma unroll
      for (int k = 0; k < U; ++k) z[k] = ld_x<XT>(x + base + static_cast<int64_t>(c + k) * hw);
      float mc = z[0];
#pragma unroll
      for (int k = 1; k < U; ++k) mc = fmaxf(mc, z[k]);
      const float mn = fmaxf(m, mc);
      s *= ex2a((m - mn) * kLog2e);       // m = -inf on the first chunk: ex2(-inf) = 0
      const float mo = mn * kLog2e;
#pragma unroll
      for (int k = 0; k < U; ++k) {
        s += ex2a(fmaf(z[k], kLog2e, -mo));
        A = fmaf(SMEM_TABLE ? wrow[c + k] : __ldg(wrow + c + k), z[k], A);
      }
      m = mn;
    }
    for (; c < C; ++c) {
      const float z = ld_x<XT>(x + base + static_cast<int64_t>(c) * hw);
      const float mn = fmaxf(m, z);
      s = s * ex2a((m - mn) * kLog2e) + ex2a((z - mn) * kLog2e);
      A = fmaf(SMEM_TABLE ? wrow[c] : __ldg(wrow + c), z, A);
      m = mn;
    }
    const float lse2 = fmaf(m, kLog2e, mufu_lg2(s));   // log2 sum exp
    const float lse = lse2 * kLn2;
    loss_acc += static_cast<double>(fmaf(lse, rs, -A));  // -sum_c w_c (z_c - lse)
    // ---- pass 2: gradient (the logits of this pixel come back through L1 / L2) ----
    if constexpr (HAS_GRAD) {
      const float k_sm = rs * gcoef;
      c = 0;
      for (; c + U <= C; c += U) {
        float z[U];
#pragma unroll
        for (int k = 0; k < U; ++k) z[k] = ld_x<XT>(x + base + static_cast<int64_t>(c + k) * hw);
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const float sm = ex2a(fmaf(z[k], kLog2e, -lse2));
          const float w = SMEM_TABLE ? wrow[c + k] : __ldg(wrow + c + k);
          Elem<XT>::st1(gx + base + static_cast<int64_t>(c + k) * hw, fmaf(sm, k_sm, -w * gcoef));
        }
      }
      for (; c < C; ++c) {
        const float z = ld_x<XT>(x + base + static_cast<int64_t>(c) * hw);
        const float sm = ex2a(fmaf(z, kLog2e, -lse2));
        const float w = SMEM_TABLE ? wrow[c] : __ldg(wrow + c);
        Elem<XT>::st1(gx + base + static_cast<int64_t>(c) * hw, fmaf(sm, k_sm, -w * gcoef));
      }
    }
  }
  publish_one_w(loss_acc, &ws.hdr->tacc[0], sm_red);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sm_last = (atomicAdd(&ws.hdr->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (sm_last && threadIdx.x == 0) {
    __threadfence();
    const double sum = __ldcg(&ws.hdr->tacc[0]);
    *a.loss_out = static_cast<float>(sum / n_valid);   // n_valid == 0 -> NaN / inf as the reference's division gives
    ws.hdr->tacc[0] = 0.0;
    ws.hdr->tacc[10] = 0.0;
    ws.hdr->ticket = 0u;
  }
}

template <typename XT>
int launch_wcel(WcelArgs& a, cudaStream_t st) {
  const int64_t npx = a.n * a.hw;
  const int C = a.C;
  const size_t table = (static_cast<size_t>(C) * (C | 1) + C) * sizeof(float);
  const bool fits = table <= 100 * 1024;   // two CTAs per SM
  const bool g = a.grad != nullptr;
  int64_t grid = (npx + kWBlock - 1) / kWBlock;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 2;
  if (grid > cap) grid = cap;
#define MDE_WCEL_LAUNCH(G, S, UU, PS)                                                                      \
  do {                                                                                                     \
    auto fn = wcel_kernel<XT, G, S, UU, PS>;                                                               \
    const size_t smem = (S) ? table : 0;                                                                   \
    if (smem > 48 * 1024)                                                                                  \
      MDE_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    fn<<<static_cast<unsigned>(grid), kWBlock, smem, st>>>(a);                                             \
  } while (0)
  if (fits) {
    if (g) MDE_WCEL_LAUNCH(true, true, 8, 2); else MDE_WCEL_LAUNCH(false, true, 8, 2);
  } else {
    if (g) MDE_WCEL_LAUNCH(true, false, 8, 2); else MDE_WCEL_LAUNCH(false, false, 8, 2);
  }
#undef MDE_WCEL_LAUNCH
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

// ---- depth -> bins (modules/vnl.py:202-217), in place on `depth` exactly as the reference ------------------
__global__ void __launch_bounds__(256) depth_to_bins_kernel(float* __restrict__ depth, int64_t n, float depth_min,
                                                            float depth_max, float min_log, float interval, int C,
                                                            int* __restrict__ bins) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256) {
    float d = depth[i];
    const bool invalid = d < 0.f;                 // :208
    d = (d < depth_min) ? depth_min : d;          // :210 (NaN stays)
    d = (d > depth_max) ? depth_max : d;          // :211
    // :212  ((log10(depth) - min_log) / interval).to(int): fp32 ops, truncation toward zero
    const float q = __fdiv_rn(log10f(d) - min_log, interval);
    int b = static_cast<int>(q);
    if (invalid) b = C + 1;                       // :213
    if (b == C) b = C - 1;                        // :214
    bins[i] = b;
    depth[i] = invalid ? -1.0f : d;               // :215
  }
}

// ---- bins -> depth (modules/vnl.py:219-230): depth = 10 ^ sum_c p_c * border_c -------------------------------
template <typename XT>
__global__ void __launch_bounds__(256) bins_to_depth_kernel(const XT* __restrict__ p, const float* __restrict__ border,
                                                            int64_t n, int C, int64_t hw, float* __restrict__ depth) {
  extern __shared__ float sm_b[];
  for (int i = threadIdx.x; i < C; i += 256) sm_b[i] = __ldg(border + i);
  __syncthreads();
  const int64_t npx = n * hw;
  for (int64_t px = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; px < npx;
       px += static_cast<int64_t>(gridDim.x) * 256) {
    const int64_t img = px / hw;
    const int64_t base = img * static_cast<int64_t>(C) * hw + (px - img * hw);
    float acc = 0.f;
    int c = 0;
    for (; c + kWUnroll <= C; c += kWUnroll) {
      float z[kWUnroll];
#pragma unroll
      for (int k = 0; k < kWUnroll; ++k) z[k] = ld_x<XT>(p + base + static_cast<int64_t>(c + k) * hw);
#pragma unroll
      for (int k = 0; k < kWUnroll; ++k) acc = fmaf(z[k], sm_b[c + k], acc);
    }
    for (; c < C; ++c) acc = fmaf(ld_x<XT>(p + base + static_cast<int64_t>(c) * hw), sm_b[c], acc);
    depth[px] = exp10f(acc);
  }
}

// backward: grad_p[c] = g * ln(10) * depth * border_c. Write-only over C planes: a warp owns 128 consecutive
// pixels (4 per lane, 32 apart) so that every visit of a channel plane writes 512 contiguous bytes - with one
// pixel per lane (128 B per plane visit, 150 planes 593 KB apart) the stores ran at 3.0 TB/s, in this form at
// 5.0 TB/s. (The same lockstep applied to the READS of wcel_kernel was measured 10-30 % slower and is not used.)
template <typename XT>
__global__ void __launch_bounds__(256) bins_to_depth_bwd_kernel(const float* __restrict__ depth, const float* __restrict__ gdepth,
                                                                const float* __restrict__ border, int64_t n, int C,
                                                                int64_t hw, XT* __restrict__ gp) {
  extern __shared__ float sm_b[];
  for (int i = threadIdx.x; i < C; i += 256) sm_b[i] = __ldg(border + i);
  __syncthreads();
  const int64_t npx = n * hw;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * 256) >> 5;
  for (int64_t g0 = warp_global * 128; g0 < npx; g0 += n_warps * 128) {
    float k[4];
    int64_t base[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t px = g0 + j * 32 + lane;
      const bool ok = px < npx;
      const int64_t pc = ok ? px : 0;
      const int64_t img = pc / hw;
      base[j] = ok ? img * static_cast<int64_t>(C) * hw + (pc - img * hw) : -1;
      k[j] = ok ? __ldg(gdepth + pc) * (__ldg(depth + pc) * 2.302585092994046f) : 0.f;
    }
    for (int c = 0; c < C; ++c) {
      const float bc = sm_b[c];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (base[j] >= 0) Elem<XT>::st1(gp + base[j] + static_cast<int64_t>(c) * hw, k[j] * bc);
    }
  }
}

inline unsigned px_grid_w(int64_t npx, int block, int per_sm) {
  int64_t g = (npx + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace
}  // namespace mde

extern "C" int mde_wcel_loss(const void* logits, int x_dtype, const int* gt_bins, const float* gt_depth,
                             const float* weight, const float* rowsum, int64_t n, int64_t C, int64_t hw,
                             float grad_scale, void* ws, float* loss_out, void* grad_logits, void* stream) {
  using namespace mde;
  MDE_REQUIRE(logits && gt_bins && gt_depth && weight && rowsum && ws && loss_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n > 0 && C > 0 && hw > 0 && C < 32768, MDE_EINVAL, "bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  wcel_count_kernel<<<px_grid_w(n * hw, 256, 4), 256, 0, st>>>(gt_depth, n * hw, ws);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  WcelArgs a{};
  a.x = logits; a.bins = gt_bins; a.weight = weight; a.rowsum = rowsum; a.n = n; a.hw = hw; a.C = static_cast<int>(C);
  a.grad_scale = grad_scale; a.ws = ws; a.loss_out = loss_out; a.grad = grad_logits;
  switch (x_dtype) {
    case MDE_F32: return launch_wcel<float>(a, st);
    case MDE_F16: return launch_wcel<__half>(a, st);
    case MDE_BF16: return launch_wcel<__nv_bfloat16>(a, st);
    default: set_error("mde_wcel_loss: unknown x_dtype %d", x_dtype); return MDE_EINVAL;
  }
}

extern "C" int mde_depth_to_bins(float* depth_inout, int64_t n, float depth_min, float depth_max, float depth_min_log,
                                 float depth_bin_interval, int64_t C, int* bins_out, void* stream) {
  using namespace mde;
  MDE_REQUIRE(depth_inout && bins_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(C > 0 && C < 32768, MDE_EINVAL, "bad channel count");
  if (n <= 0) return MDE_OK;
  depth_to_bins_kernel<<<px_grid_w(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      depth_inout, n, depth_min, depth_max, depth_min_log, depth_bin_interval, static_cast<int>(C), bins_out);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

extern "C" int mde_bins_to_depth(const void* prob, int x_dtype, const float* border, int64_t n, int64_t C, int64_t hw,
                                 float* depth_out, void* stream) {
  using namespace mde;
  MDE_REQUIRE(prob && border && depth_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n > 0 && C > 0 && hw > 0 && C <= 8192, MDE_EINVAL, "bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = px_grid_w(n * hw, 256, 8);
  const size_t smem = static_cast<size_t>(C) * sizeof(float);
  const int c = static_cast<int>(C);
  switch (x_dtype) {
    case MDE_F32: bins_to_depth_kernel<float><<<grid, 256, smem, st>>>(static_cast<const float*>(prob), border, n, c, hw, depth_out); break;
    case MDE_F16: bins_to_depth_kernel<__half><<<grid, 256, smem, st>>>(static_cast<const __half*>(prob), border, n, c, hw, depth_out); break;
    case MDE_BF16: bins_to_depth_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(prob), border, n, c, hw, depth_out); break;
    default: set_error("mde_bins_to_depth: unknown x_dtype %d", x_dtype); return MDE_EINVAL;
  }
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

extern "C" int mde_bins_to_depth_bwd(const float* depth, const float* grad_depth, const float* border, int64_t n, int64_t C,
                                     int64_t hw, int x_dtype, void* grad_prob, void* stream) {
  using namespace mde;
  MDE_REQUIRE(depth && grad_depth && border && grad_prob, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n > 0 && C > 0 && hw > 0 && C <= 8192, MDE_EINVAL, "bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = px_grid_w(n * hw, 256, 8);
  const size_t smem = static_cast<size_t>(C) * sizeof(float);
  const int c = static_cast<int>(C);
  switch (x_dtype) {
    case MDE_F32: bins_to_depth_bwd_kernel<float><<<grid, 256, smem, st>>>(depth, grad_depth, border, n, c, hw, static_cast<float*>(grad_prob)); break;
    case MDE_F16: bins_to_depth_bwd_kernel<__half><<<grid, 256, smem, st>>>(depth, grad_depth, border, n, c, hw, static_cast<__half*>(grad_prob)); break;
    case MDE_BF16: bins_to_depth_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(depth, grad_depth, border, n, c, hw, static_cast<__nv_bfloat16*>(grad_prob)); break;
    default: set_error("mde_bins_to_depth_bwd: unknown x_dtype %d", x_dtype); return MDE_EINVAL;
  }
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
