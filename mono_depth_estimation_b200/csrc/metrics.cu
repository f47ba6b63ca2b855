// metrics.cu - fused masked error-metric kernel (replaces MetricComputation.compute,
// reference metrics.py:58-67, and the metric functions metrics.py:75-109,116-122).
//
// One pass over pred/target (8 B/px algorithmic): every CTA owns a contiguous, 128-byte aligned
// chunk of the batch, accumulates the 4 integer counts and up to 8 float sums PER IMAGE
// (fp32 over <= 16 pixels, fp64 across tiles), flushes with one warp-shuffle/shared-memory block
// reduction per image it touched and one fp64 atomic per quantity, and the last CTA to finish
// turns the per-image sums into pooled values, per-image values and the mean over images.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include <cstring>
#include "metric_math.cuh"

namespace mde {

namespace {

constexpr int kNQ = MDE_METRIC_NQ;
constexpr int kNM = MDE_METRIC_NM;

// LONG = false: a thread sees at most 128 pixels of an image, so its sums stay in fp32 registers
// until the flush (error <= ~4e-7 relative). LONG = true: every 16 iterations (128 pixels) the fp32
// tile sums are folded into fp64 running sums that live in SHARED memory (one column per thread), so
// the hot loop carries no fp64 registers and the software-pipelined loads fit without spilling.
// Fast mode (Ref = false) uses the LEAN per-pixel form of metric_math.cuh (round 2): float threshold / valid
// counters fed by FSET / FMUL.SAT + FADD, |p - t| = hi - lo, 11 ALU-pipe instructions per pixel instead of ~20.
// G may carry kGrpRsq (only the 'rmse' sum of the REL group); the reference-arithmetic paths widen it to kGrpRel.
template <unsigned G, bool Ref, bool LONG>
struct MetricThread {
  static constexpr unsigned kRefG = (G & kGrpRsq) ? ((G & 7u) | kGrpRel) : (G & 7u);
  MetricTile tile;     // Ref mode
  MetricCounts cnt;    // Ref mode; in fast mode: the folded integer counts (n, c1, c2, c3)
  MetricAcc acc;       // fast mode
  int lean_px;         // pixels that went through the lean form since the last fold
  double* srun;        // LONG: &sm_run[0][threadIdx.x], element q at srun[q * kBlock]
  int it;

  __device__ __forceinline__ void reset() {
    tile.zero();
    cnt.zero();
    acc.zero();
    lean_px = 0;
    it = 0;
    if constexpr (LONG) {
#pragma unroll
      for (int q = 0; q < 8; ++q) srun[q * kBlock] = 0.0;
    }
  }
  // fast mode: float counters -> exact integer counts (callers keep < 2^24 pixels between two calls)
  __device__ __forceinline__ void settle_counts() {
    if constexpr (!Ref) {
      const int nv = static_cast<int>(acc.nval), inv = lean_px - nv;   // invalid pixels are counted by all three float counters
      cnt.n += nv;
      cnt.c1 += static_cast<int>(acc.c1) - inv;
      cnt.c2 += static_cast<int>(acc.c2) - inv;
      cnt.c3 += static_cast<int>(acc.c3) - inv;
      acc.c1 = acc.c2 = acc.c3 = acc.nval = 0.f;
      lean_px = 0;
    }
  }
  // rare path (fast mode): one pixel in reference arithmetic; sums in the lean units, exact counts straight into cnt
  __device__ __forceinline__ void rare_px(float p, float t) {
    const MetricContrib r = metric_px_ref_contrib<kRefG>(p, t);
    acc.s_abs += r.s.s_abs; acc.s_sq += r.s.s_sq;
    acc.s_log10 += r.s.s_log10 * (1.0f / tile_scale<false>(2));
    acc.s_sle += r.s.s_sle * (1.0f / tile_scale<false>(3));
    acc.s_absrel += r.s.s_absrel; acc.s_sqrel += r.s.s_sqrel; acc.s_rsq += r.s.s_rsq;
    acc.s_lnsq += r.s.s_lnsq * (1.0f / tile_scale<false>(7));
    cnt.n += r.c.n; cnt.c1 += r.c.c1; cnt.c2 += r.c.c2; cnt.c3 += r.c.c3;
  }
  // scalar path: the rare-case test (valid subnormal target, see metric_quad_needs_ref) per pixel
  __device__ __forceinline__ void px(float p, float t) {
    if constexpr (Ref) {
      metric_px<kRefG, true>(p, t, tile, cnt);
    } else {
      if (t > 0.f && t < 1.17549435e-38f) {
        rare_px(p, t);
      } else {
        metric_px_lean<G, false>(p, t, acc);
        ++lean_px;
      }
    }
  }
  __device__ __forceinline__ void quad(const float4& p, const float4& t) {
    if constexpr (Ref) {
      metric_px<kRefG, true>(p.x, t.x, tile, cnt);
      metric_px<kRefG, true>(p.y, t.y, tile, cnt);
      metric_px<kRefG, true>(p.z, t.z, tile, cnt);
      metric_px<kRefG, true>(p.w, t.w, tile, cnt);
    } else if (metric_quad_needs_ref(t)) {
      rare_px(p.x, t.x);
      rare_px(p.y, t.y);
      rare_px(p.z, t.z);
      rare_px(p.w, t.w);
    } else {
      metric_px_lean<G, false>(p.x, t.x, acc);
      metric_px_lean<G, false>(p.y, t.y, acc);
      metric_px_lean<G, false>(p.z, t.z, acc);
      metric_px_lean<G, false>(p.w, t.w, acc);
      lean_px += 4;
    }
  }
  static __device__ __forceinline__ bool used(int q) {
    return (q < 2) || ((G & kGrpLog) && (q == 2 || q == 7)) || ((G & kGrpLog1p) && q == 3) ||
           ((G & kGrpRel) && (q >= 4 && q <= 6)) || ((G & kGrpRsq) && q == 6);
  }
  // tile sum q in the unit of its raw quantity (the fast forms carry logarithms in log2 units)
  __device__ __forceinline__ float tile_q(int q) const {
    if constexpr (Ref) {
      const float s = (q == 0) ? tile.s_abs : (q == 1) ? tile.s_sq : (q == 2) ? tile.s_log10 : (q == 3) ? tile.s_sle
                    : (q == 4) ? tile.s_absrel : (q == 5) ? tile.s_sqrel : (q == 6) ? tile.s_rsq : tile.s_lnsq;
      return s;
    } else {
      return acc.sum(q) * tile_scale<false>(q);
    }
  }
  // value of running sum q at flush time
  __device__ __forceinline__ float total(int q) const {
    if constexpr (LONG) return static_cast<float>(srun[q * kBlock] + static_cast<double>(tile_q(q)));
    return tile_q(q);
  }
  // called once per loop iteration (<= 8 pixels)
  __device__ __forceinline__ void fold() {
    if constexpr (!LONG) return;
    if ((++it & 15) != 0) return;
    fold_now();
  }
  // fp32 tile sums -> fp64 running sums (LONG); callers keep <= 128 pixels between two calls
  __device__ __forceinline__ void fold_now() {
    if constexpr (!LONG) return;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (used(q)) srun[q * kBlock] += static_cast<double>(tile_q(q));
    }
    if constexpr (Ref) {
      tile.zero();
      cnt.unpack();
    } else {
      acc.s_abs = acc.s_sq = acc.s_log10 = acc.s_sle = acc.s_absrel = acc.s_sqrel = acc.s_rsq = acc.s_lnsq = 0.f;
      settle_counts();
    }
  }
};

// raw-quantity index of run[i]
__constant__ int kRunToQ[8] = {MDE_Q_ABS, MDE_Q_SQ, MDE_Q_LOG10, MDE_Q_SLE,
                               MDE_Q_ABSREL, MDE_Q_SQREL, MDE_Q_RSQ, MDE_Q_LNSQ};

// Flush one image's partial sums of this CTA: warp shuffle -> shared memory -> 12 fp64 atomics into the image's
// accumulator row. Ends with the shared buffers consumed.
template <unsigned G, bool Ref, bool LONG>
__device__ __forceinline__ void flush_image(MetricThread<G, Ref, LONG>& th, int64_t img, double* iacc, double* sm_d, int* sm_i) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if constexpr (Ref) th.cnt.unpack();
  else th.settle_counts();
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
  // per-thread sums are fp64; the 32-lane tree is done in fp32 (5 levels, <= 4e-7 relative) and
  // widened again before the cross-warp and cross-CTA accumulation
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (!(MetricThread<G, Ref, LONG>::used(q))) continue;
    const float s = warp_sum(th.total(q));
    if (lane == 0) sm_d[q * kWarps + warp] = static_cast<double>(s);
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
      if (MetricThread<G, Ref, LONG>::used(q))
        for (int w = 0; w < kWarps; ++w) tot += sm_d[q * kWarps + w];
      qidx = kRunToQ[q];
    }
    if (tot != 0.0) atomicAdd(&iacc[img * kIacc + qidx], tot);
  }
  __syncthreads();
}

// Numerator quantity of metric value m (value = raw[num] / n, sqrt for the last two)
__constant__ int kValNum[kNM] = {MDE_Q_D1, MDE_Q_D2, MDE_Q_D3, MDE_Q_ABS, MDE_Q_SQ, MDE_Q_LOG10, MDE_Q_SLE,
                                 MDE_Q_ABSREL, MDE_Q_SQREL, MDE_Q_RSQ, MDE_Q_SQ, MDE_Q_LNSQ};

// ---- in-kernel exchange of the evaluation sums over NVLink peer memory -------------------------------
// A multi-GPU evaluation shards the images over the ranks and needs ONE sum of 25 doubles per rank (pooled raw sums,
// number of valid images, sum of per-image values). Instead of a collective launched behind the kernel, the finaliser
// of every rank's launch WRITES its 25 doubles straight into every peer's mailbox (peer-mapped device memory, stores
// over NVLink) and sums the world's rows of its own mailbox in rank order - the same order on every rank, so all ranks
// end with bit-identical totals. Every double travels as two self-validating words {seq:32 | half:32}: no flag, no
// fence, no ordering requirement between the stores. Rows are double-buffered by the parity of seq (a rank can be at
// most one call ahead of a peer that has not yet read its previous row). A peer that never shows up is given up on
// after timeout_ms (NaN results + the workspace error flag) instead of hanging the GPU.
struct PeerDesc {                                 // lives in DEVICE memory (mde_peer_comm_create): the kernels take a pointer
  unsigned long long* mailbox[MDE_MAX_PEERS];   // [r] = rank r's mailbox as mapped into THIS process
  int rank, world;
  unsigned timeout_ms;
  unsigned auto_seq;                            // seq == 0 at launch: the finaliser takes ++auto_seq (CUDA-graph friendly)
};
struct PeerXchg {                                 // what the finaliser works with (registers / constant bank)
  const PeerDesc* d;
  int rank, world;
  unsigned seq;
};
constexpr int kPeerRowWords = 64;                 // 25 doubles x 2 words, padded
constexpr int kPeerVals = 2 * kNM + 1;
static_assert(2 * kPeerVals <= kPeerRowWords, "mailbox row too small");
static_assert(static_cast<size_t>(2) * MDE_MAX_PEERS * kPeerRowWords * 8 == MDE_PEER_MAILBOX_BYTES, "mailbox size");

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_put(const PeerXchg& px, int idx, double v) {
  const unsigned long long tag = static_cast<unsigned long long>(px.seq) << 32;
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  const size_t off = (static_cast<size_t>(px.seq & 1u) * MDE_MAX_PEERS + px.rank) * kPeerRowWords + 2 * idx;
  for (int r = 0; r < px.world; ++r) {
    unsigned long long* box = px.d->mailbox[r];
    st_relaxed_sys_u64(box + off, tag | (b >> 32));
    st_relaxed_sys_u64(box + off + 1, tag | (b & 0xffffffffull));
  }
}
// sum over the ranks (in rank order) of value `idx`; false when a peer's row did not arrive in time
__device__ __forceinline__ bool peer_get(const PeerXchg& px, int idx, unsigned long long t_give_up, double& out) {
  const unsigned long long* mine = px.d->mailbox[px.rank] + static_cast<size_t>(px.seq & 1u) * MDE_MAX_PEERS * kPeerRowWords + 2 * idx;
  double acc = 0.0;
  for (int s = 0; s < px.world; ++s) {
    const unsigned long long* w = mine + static_cast<size_t>(s) * kPeerRowWords;
    unsigned long long hi = ld_relaxed_sys_u64(w), lo = ld_relaxed_sys_u64(w + 1);
    while ((hi >> 32) != px.seq || (lo >> 32) != px.seq) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now > t_give_up) return false;
      hi = ld_relaxed_sys_u64(w);
      lo = ld_relaxed_sys_u64(w + 1);
    }
    acc += __longlong_as_double(static_cast<long long>((hi << 32) | (lo & 0xffffffffull)));
  }
  out = acc;
  return true;
}

// Run by the LAST CTA only. One HALF-WARP per image: sub-lane q < 12 owns raw quantity q and metric value q, so an image
// costs one coalesced L2 read, two shuffles and one division per lane; the 32 half-warps of the CTA stride over the images,
// eight images per step with all their rows requested up front, then the halves of a warp and the warps are combined. Also
// re-zeroes the per-image accumulators so the workspace is clean for the next call.
// (Round 1 gave a whole warp one image at a time and divided / took roots with the IEEE fp64 sequences: a chain of L2 round
// trips and ~1400 cycles of dependent fp64 arithmetic per image, 37 of the 294 us of the C5 launch - measured by presenting
// the same 200.9 M pixels as 6 images instead of 654, tools/c5_split_probe.py. The per-image quotient is now num x (1 / n)
// with the reciprocal from the SFU and two Newton steps in fp64 (full double precision up to the last ulp or two), the
// root likewise from MUFU.RSQ; the POOLED values keep the IEEE operations.)
__device__ __forceinline__ double fast_quotient(double num, double n) {
  double inv = static_cast<double>(1.0f / static_cast<float>(n));   // n = a pixel count (exact in fp32); n == 0 -> NaN below, as 0 / 0
  inv = inv * (2.0 - n * inv);
  inv = inv * (2.0 - n * inv);
  return num * inv;
}
__device__ __forceinline__ double fast_root(double x) {
  if (!(x > 1e-30 && x < 1e30)) return sqrt(x);                   // 0, NaN, inf and whatever does not fit fp32: the exact path
  double y = static_cast<double>(rsqrtf(static_cast<float>(x)));
  y = y * (1.5 - 0.5 * x * y * y);
  y = y * (1.5 - 0.5 * x * y * y);
  return x * y;
}
__device__ __noinline__ void metrics_finalize(Ws ws, int64_t n_img, double* __restrict__ out_f64,
                                              float* __restrict__ out_f32, double* __restrict__ per_image_values,
                                              double* __restrict__ per_image_raw, double* sm_d, PeerDesc* pd, unsigned seq) {
  double* iacc = ws.iacc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 15, hbase = lane & 16;
  const bool own = sub < kNM;  // kNM == kNQ == 12
  const int num_idx = own ? kValNum[sub] : 0;
  double pooled = 0.0, vsum = 0.0, nvalid = 0.0;
  constexpr int kFinU = 8, kGroups = 2 * kWarps;
  const int grp = 2 * warp + (lane >> 4);
  // images per half-warp and step: up to kFinU, but never more than it takes to give every half-warp work (a C2 batch of
  // 16 images is one image for each of 16 half-warps, not eight images for each of two)
  const int per = static_cast<int>(n_img >= static_cast<int64_t>(kGroups) * kFinU ? kFinU : (n_img + kGroups - 1) / kGroups);
  const int U = per < 1 ? 1 : per;
  const int64_t steps = (n_img + static_cast<int64_t>(kGroups) * U - 1) / (static_cast<int64_t>(kGroups) * U);   // the same for every half-warp: the shuffles stay warp-wide
  for (int64_t k = 0; k < steps; ++k) {
    const int64_t b0 = (k * kGroups + grp) * U;
    double rawv[kFinU];
#pragma unroll
    for (int u = 0; u < kFinU; ++u) {
      const int64_t b = b0 + u;
      rawv[u] = (own && u < U && b < n_img) ? __ldcg(&iacc[b * kIacc + sub]) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kFinU; ++u) {
      if (u >= U) break;                                        // uniform
      const int64_t b = b0 + u;
      const bool live = b < n_img;
      const double raw = rawv[u];
      if (own && live) iacc[b * kIacc + sub] = 0.0;
      const double n = __shfl_sync(0xffffffffu, raw, hbase + MDE_Q_NVALID);
      const double num = __shfl_sync(0xffffffffu, raw, hbase + num_idx);
      double val = fast_quotient(num, n);
      if (sub >= MDE_M_RMSE_TRUE) val = fast_root(val);
      if (live) {
        pooled += raw;
        if (n > 0.0) {
          vsum += val;
          nvalid += 1.0;
        }
        if (own) {
          if (per_image_values) per_image_values[b * kNM + sub] = val;
          if (per_image_raw) per_image_raw[b * kNQ + sub] = raw;
        }
      }
    }
  }
  // the two half-warps of a warp
  pooled += __shfl_xor_sync(0xffffffffu, pooled, 16);
  vsum += __shfl_xor_sync(0xffffffffu, vsum, 16);
  nvalid += __shfl_xor_sync(0xffffffffu, nvalid, 16);
  // sm_d: [kWarps][12] pooled | [kWarps][12] vsum | [kWarps] nvalid
  if (lane < kNM) {
    sm_d[warp * kNM + lane] = pooled;
    sm_d[kWarps * kNM + warp * kNM + lane] = vsum;
  }
  if (lane == 0) sm_d[2 * kWarps * kNM + warp] = nvalid;
  __syncthreads();
  if (warp == 0) {
    const bool own0 = lane < kNM;   // (in this block lane q owns quantity q; `own` above is per half-warp)
    double P = 0.0, V = 0.0, N = 0.0;
    if (own0) {
      for (int w = 0; w < kWarps; ++w) {
        P += sm_d[w * kNM + lane];
        V += sm_d[kWarps * kNM + w * kNM + lane];
      }
    }
    for (int w = 0; w < kWarps; ++w) N += sm_d[2 * kWarps * kNM + w];
    if (pd != nullptr) {   // this rank's {P, N, V} -> every peer's mailbox; the world's rows of the own mailbox -> {P, N, V}
      PeerXchg px;
      if (seq == 0u) {   // the communicator's own counter: every rank's launches advance it in lockstep
        unsigned sq = 0u;
        if (lane == 0) {
          sq = pd->auto_seq + 1u;
          if (sq == 0u) sq = 1u;
          pd->auto_seq = sq;
        }
        seq = __shfl_sync(0xffffffffu, sq, 0);
      }
      px.d = pd; px.rank = pd->rank; px.world = pd->world; px.seq = seq;
      if (own0) {
        peer_put(px, lane, P);
        peer_put(px, kNM + 1 + lane, V);
      }
      if (lane == 0) peer_put(px, kNM, N);
      unsigned long long t_give_up;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_give_up));
      t_give_up += static_cast<unsigned long long>(pd->timeout_ms) * 1000000ull;
      bool ok = true;
      if (own0) ok = peer_get(px, lane, t_give_up, P) && peer_get(px, kNM + 1 + lane, t_give_up, V);
      double n_all = N;
      if (lane == 0) ok = peer_get(px, kNM, t_give_up, n_all) && ok;
      N = __shfl_sync(0xffffffffu, n_all, 0);
      if (!__all_sync(0xffffffffu, ok)) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        P = V = N = nan;
        if (lane == 0) ws.hdr->error = 1u;
      }
    }
    const double n = __shfl_sync(0xffffffffu, P, MDE_Q_NVALID);
    const double num = __shfl_sync(0xffffffffu, P, own0 ? num_idx : 0);
    double val = num / n;
    if (lane >= MDE_M_RMSE_TRUE) val = sqrt(val);
    if (own0) {
      const double mean_v = V / N;
      out_f64[lane] = val;
      out_f64[kNM + lane] = mean_v;
      out_f64[2 * kNM + lane] = P;
      out_f64[2 * kNM + kNQ + 1 + lane] = V;   // per-image value sums: {P, N, V} is the vector ranks all-reduce
      if (out_f32) {
        out_f32[lane] = static_cast<float>(val);
        out_f32[kNM + lane] = static_cast<float>(mean_v);
      }
    }
    if (lane == 0) {
      out_f64[2 * kNM + kNQ] = N;
      ws.hdr->ticket = 0;
    }
  }
}

template <typename PT, int VEC, unsigned G, bool Ref, bool LONG>
__global__ void __launch_bounds__(kBlock, kCtasPerSm)
metrics_kernel(const PT* __restrict__ pred, const float* __restrict__ gt, int64_t n_img, int64_t hw, Chunking chunk,
               void* ws_raw, double* __restrict__ out_f64, float* __restrict__ out_f32,
               double* __restrict__ per_image_values, double* __restrict__ per_image_raw, PeerDesc* __restrict__ pd,
               unsigned seq) {
  __shared__ double sm_d[2 * kNM * kWarps + kWarps];
  __shared__ int sm_i[4 * kWarps];
  __shared__ bool sm_last;

  pdl_wait();   // launched with launch_pdl: nothing a predecessor wrote may be read before this point
  Ws ws = ws_view(ws_raw);
  double* iacc = ws.iacc;  // parity set 0: [n_img][kIacc]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t units_per_img = hw / VEC;  // VEC == 4 requires hw % 4 == 0 (checked on the host)
  int64_t ub, ue;
  cta_chunk(chunk, blockIdx.x, ub, ue);

  __shared__ double sm_run[LONG ? 8 * kBlock : 1];
  MetricThread<G, Ref, LONG> th;
  th.srun = LONG ? &sm_run[threadIdx.x] : nullptr;
  int64_t u = ub;
  while (u < ue) {
    const int64_t img = u / units_per_img;
    int64_t seg_end = (img + 1) * units_per_img;
    if (seg_end > ue) seg_end = ue;
    th.reset();

    int64_t i = u + threadIdx.x;
    if (VEC == 4) {
      // Software pipeline over three register buffers (A, B, C) of one quad x {pred, target} each: while one
      // buffer is evaluated the other two are in flight (64 B per thread, 64 KB per SM). The loop is unrolled
      // over the buffer triple so that nothing is ever moved between registers (the rotating form spent 16
      // MOVs per 8 pixels and spilled), its counters are 32-bit offsets from the segment start, and the fp64
      // fold (LONG) runs after every 8 triples = 96 pixels.
      // (Feeding this loop from shared memory with cp.async.bulk + one __syncthreads per tile was measured:
      // 392 us instead of 325 us at C5 - the per-tile CTA barrier costs more than the plumbing it removes.)
      const PT* __restrict__ pbase = pred + 4 * u;
      const float* __restrict__ gbase = gt + 4 * u;
      const int nseg = static_cast<int>(seg_end - u);          // quads in this segment (< 2^31: one image at most)
      int j = threadIdx.x;                                       // quad offset of buffer A
      float4 pa, ta, pb, tb, pc, tc;
      auto load1 = [&](int jj, float4& p0, float4& t0) {
        if (jj < nseg) {
          p0 = Elem<PT>::template ld4<false>(pbase + 4 * static_cast<int64_t>(jj));
          t0 = Elem<float>::template ld4<false>(gbase + 4 * static_cast<int64_t>(jj));
        }
      };
      load1(j, pa, ta);
      load1(j + kBlock, pb, tb);
      while (j < nseg) {
#pragma unroll 1
        for (int r = 0; r < 8 && j < nseg; ++r) {
          load1(j + 2 * kBlock, pc, tc);
          th.quad(pa, ta);
          load1(j + 3 * kBlock, pa, ta);
          if (j + kBlock < nseg) th.quad(pb, tb);
          load1(j + 4 * kBlock, pb, tb);
          if (j + 2 * kBlock < nseg) th.quad(pc, tc);
          j += 3 * kBlock;
        }
        th.fold_now();
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

    flush_image<G, Ref, LONG>(th, img, iacc, sm_d, sm_i);
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

  metrics_finalize(ws, n_img, out_f64, out_f32, per_image_values, per_image_raw, sm_d, pd, seq);
}

// a rank without images (fewer images than ranks) still owes the world its (all-zero) row
__global__ void __launch_bounds__(kBlock) metrics_empty_shard_kernel(void* ws_raw, double* __restrict__ out_f64, float* __restrict__ out_f32,
                                                                      PeerDesc* __restrict__ pd, unsigned seq) {
  __shared__ double sm_d[2 * kNM * kWarps + kWarps];
  pdl_wait();
  metrics_finalize(ws_view(ws_raw), 0, out_f64, out_f32, nullptr, nullptr, sm_d, pd, seq);
}

// ---- the same exchange as a tiny stand-alone all-reduce (sum) of <= 31 doubles ---------------------------------------------
// What a training job needs once per logging interval: the ranks' pooled metric sums (12 doubles) added up. One warp:
// lane i puts value i into every peer's mailbox row and sums the world's rows of its own mailbox in rank order; with
// `zero_src` the source accumulator is cleared in the same launch (read - exchange - write - clear, nothing else on the
// stream). A collective kernel would hold an SM for tens of microseconds while the cooperative loss kernels, which need
// every SM, wait behind it.
__global__ void __launch_bounds__(32) peer_allreduce_kernel(double* __restrict__ src, double* __restrict__ dst, int n, int zero_src,
                                                            PeerDesc* __restrict__ pd, unsigned seq, void* ws_raw) {
  pdl_wait();
  const int lane = threadIdx.x;
  if (seq == 0u) {
    unsigned sq = 0u;
    if (lane == 0) {
      sq = pd->auto_seq + 1u;
      if (sq == 0u) sq = 1u;
      pd->auto_seq = sq;
    }
    seq = __shfl_sync(0xffffffffu, sq, 0);
  }
  PeerXchg px;
  px.d = pd; px.rank = pd->rank; px.world = pd->world; px.seq = seq;
  double v = 0.0;
  if (lane < n) {
    v = src[lane];
    if (zero_src) src[lane] = 0.0;
    peer_put(px, lane, v);
  }
  unsigned long long t_give_up;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_give_up));
  t_give_up += static_cast<unsigned long long>(pd->timeout_ms) * 1000000ull;
  bool ok = true;
  if (lane < n) ok = peer_get(px, lane, t_give_up, v);
  if (!__all_sync(0xffffffffu, ok)) {
    v = __longlong_as_double(0x7ff8000000000000LL);
    if (lane == 0) ws_view(ws_raw).hdr->error = 1u;
  }
  if (lane < n) dst[lane] = v;
}

// ---- metrics on bilinearly resized inputs -----------------------------------------------
// The test steps of the eigen / dorn / my modules resize BOTH the prediction and the target to 480 x 640 with
// F.interpolate(mode='bilinear') (align_corners=False) right before log_test (modules/eigen.py:49-51,
// modules/dorn.py:181-183, modules/my.py:64-66). Here every output pixel samples its two sources on the fly
// (ATen's rule: src = (dst + 0.5) * in/out - 0.5 clamped at 0, the 4-tap blend in ATen's CUDA op order) and goes
// straight into the metric accumulators: the two resized images are never written or re-read.
__device__ __forceinline__ float bilinear_tap(const float* __restrict__ img, int ih, int iw, float sy, float sx, int oy, int ox) {
  float fy = sy * (static_cast<float>(oy) + 0.5f) - 0.5f, fx = sx * (static_cast<float>(ox) + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = y0 + (y0 < ih - 1 ? 1 : 0), x1 = x0 + (x0 < iw - 1 ? 1 : 0);
  const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float v00 = __ldg(img + static_cast<int64_t>(y0) * iw + x0), v01 = __ldg(img + static_cast<int64_t>(y0) * iw + x1);
  const float v10 = __ldg(img + static_cast<int64_t>(y1) * iw + x0), v11 = __ldg(img + static_cast<int64_t>(y1) * iw + x1);
  return hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

template <unsigned G>
__global__ void __launch_bounds__(kBlock, kCtasPerSm)
metrics_resized_kernel(const float* __restrict__ pred, int ph, int pw, const float* __restrict__ gt, int gh, int gw,
                       int64_t n_img, int oh, int ow, int chunks_per_img, void* ws_raw, double* __restrict__ out_f64,
                       float* __restrict__ out_f32, double* __restrict__ per_image_values, double* __restrict__ per_image_raw) {
  __shared__ double sm_d[2 * kNM * kWarps + kWarps];
  __shared__ int sm_i[4 * kWarps];
  __shared__ bool sm_last;
  Ws ws = ws_view(ws_raw);
  const float spy = static_cast<float>(ph) / static_cast<float>(oh), spx = static_cast<float>(pw) / static_cast<float>(ow);
  const float sgy = static_cast<float>(gh) / static_cast<float>(oh), sgx = static_cast<float>(gw) / static_cast<float>(ow);
  const int ohw = oh * ow;
  const int per_chunk = ((ohw + chunks_per_img - 1) / chunks_per_img + kBlock - 1) / kBlock * kBlock;
  const int64_t n_work = n_img * chunks_per_img;
  MetricThread<G, false, false> th;
  th.srun = nullptr;
  for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
    const int64_t img = wk / chunks_per_img;
    const int c0 = static_cast<int>(wk - img * chunks_per_img) * per_chunk;
    const int c1 = (c0 + per_chunk < ohw) ? c0 + per_chunk : ohw;
    const float* pimg = pred + img * static_cast<int64_t>(ph) * pw;
    const float* gimg = gt + img * static_cast<int64_t>(gh) * gw;
    th.reset();
    for (int i = c0 + threadIdx.x; i < c1; i += kBlock) {
      const int oy = i / ow, ox = i - oy * ow;
      th.px(bilinear_tap(pimg, ph, pw, spy, spx, oy, ox), bilinear_tap(gimg, gh, gw, sgy, sgx, oy, ox));
    }
    flush_image<G, false, false>(th, img, ws.iacc, sm_d, sm_i);
  }
  __threadfence();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&ws.hdr->ticket, 1u);
    sm_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!sm_last) return;
  __threadfence();
  metrics_finalize(ws, n_img, out_f64, out_f32, per_image_values, per_image_raw, sm_d, nullptr, 0u);
}

template <typename PT, int VEC, unsigned G, bool Ref, bool LONG>
int launch_metrics_l(const void* pred, const float* gt, int64_t n_img, int64_t hw, void* ws, double* out_f64,
                   float* out_f32, double* piv, double* pir, cudaStream_t st, PeerDesc* pd, unsigned seq) {
  const int64_t units = n_img * (hw / VEC);
  const int64_t per_cta_min = static_cast<int64_t>(kBlock);  // at least one unit per thread
  int64_t grid = (units + per_cta_min - 1) / per_cta_min;
  const int64_t cap = static_cast<int64_t>(sm_count()) * kCtasPerSm;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const Chunking chunk = make_chunking(units, 32 / VEC, static_cast<int>(grid));
  const PT* pred_t = static_cast<const PT*>(pred);
  Chunking chunk_v = chunk;
  void* args[] = {&pred_t, &gt, &n_img, &hw, &chunk_v, &ws, &out_f64, &out_f32, &piv, &pir, &pd, &seq};
  MDE_CUDA_TRY(launch_pdl(reinterpret_cast<const void*>(&metrics_kernel<PT, VEC, G, Ref, LONG>), dim3(static_cast<unsigned>(grid)),
                          dim3(kBlock), args, 0, st, false));
  count_launch();
  return MDE_OK;
}

template <typename PT, int VEC, unsigned G, bool Ref>
int launch_metrics(const void* pred, const float* gt, int64_t n_img, int64_t hw, void* ws, double* out_f64,
                   float* out_f32, double* piv, double* pir, cudaStream_t st, PeerDesc* pd, unsigned seq) {
  // pixels one thread sees of one image: the whole batch is spread over <= 2 CTAs per SM
  const int64_t npx = n_img * hw;
  const int64_t threads = static_cast<int64_t>(sm_count()) * kCtasPerSm * kBlock;
  const bool is_long = (npx + threads - 1) / threads > 96;
  if (is_long) return launch_metrics_l<PT, VEC, G, Ref, true>(pred, gt, n_img, hw, ws, out_f64, out_f32, piv, pir, st, pd, seq);
  return launch_metrics_l<PT, VEC, G, Ref, false>(pred, gt, n_img, hw, ws, out_f64, out_f32, piv, pir, st, pd, seq);
}

template <typename PT>
int dispatch_metrics(const void* pred, const float* gt, int64_t n_img, int64_t hw, unsigned flags, void* ws,
                     double* out_f64, float* out_f32, double* piv, double* pir, cudaStream_t st, PeerDesc* pd, unsigned seq) {
  const bool ref = (flags & MDE_METRICS_REFERENCE_MATH) != 0;
  unsigned g = (flags >> 8) & kGrpMask;
  if ((g & kGrpRsq) && (g & kGrpRel)) g &= kGrpAll;                      // REL already covers the 'rmse' sum
  if ((g & kGrpRsq) && g != (kGrpLog | kGrpRsq)) g = (g & kGrpAll) | kGrpRel;   // one lean instantiation: {log, rsq}
  if (g == 0) g = kGrpAll;
  const bool vec = (hw % 4 == 0) && aligned_to(pred, 4 * sizeof(PT)) && aligned_to(gt, 16);
#define MDE_CASE(V, GG, R) return launch_metrics<PT, V, GG, R>(pred, gt, n_img, hw, ws, out_f64, out_f32, piv, pir, st, pd, seq)
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
    case 9: MDE_CASE(4, 9u, false);   // the reference's default train list: log10 + rmse (+ deltas, mse, mae)
    default: MDE_CASE(4, 7u, false);
  }
#undef MDE_CASE
}

}  // namespace
}  // namespace mde

extern "C" int mde_metrics_sharded(const void* pred, int pred_dtype, const float* target, int64_t n_img, int64_t hw,
                                   unsigned flags, void* ws, double* out_f64, float* out_f32, double* per_image_values,
                                   double* per_image_raw, void* comm, unsigned seq, void* stream) {
  using namespace mde;
  MDE_REQUIRE(ws && out_f64, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(aligned_to(out_f64, 8), MDE_EALIGN, "misaligned pointer");
  PeerDesc* pd = static_cast<PeerDesc*>(comm);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_img == 0) {   // an empty shard: legal only as a member of an exchange
    MDE_REQUIRE(pd != nullptr, MDE_EINVAL, "empty input");
    void* args[] = {&ws, &out_f64, &out_f32, &pd, &seq};
    MDE_CUDA_TRY(launch_pdl(reinterpret_cast<const void*>(&metrics_empty_shard_kernel), dim3(1), dim3(kBlock), args, 0, st, false));
    count_launch();
    return MDE_OK;
  }
  MDE_REQUIRE(pred && target, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(aligned_to(target, 4), MDE_EALIGN, "misaligned pointer");
  switch (pred_dtype) {
    case MDE_F32:
      return dispatch_metrics<float>(pred, target, n_img, hw, flags, ws, out_f64, out_f32, per_image_values,
                                     per_image_raw, st, pd, seq);
    case MDE_F16:
      return dispatch_metrics<__half>(pred, target, n_img, hw, flags, ws, out_f64, out_f32, per_image_values,
                                      per_image_raw, st, pd, seq);
    case MDE_BF16:
      return dispatch_metrics<__nv_bfloat16>(pred, target, n_img, hw, flags, ws, out_f64, out_f32,
                                             per_image_values, per_image_raw, st, pd, seq);
    default:
      set_error("mde_metrics: unknown pred_dtype %d", pred_dtype);
      return MDE_EINVAL;
  }
}

extern "C" int mde_metrics(const void* pred, int pred_dtype, const float* target, int64_t n_img, int64_t hw,
                           unsigned flags, void* ws, double* out_f64, float* out_f32, double* per_image_values,
                           double* per_image_raw, void* stream) {
  using namespace mde;
  MDE_REQUIRE(n_img > 0 && hw > 0, MDE_EINVAL, "empty input");
  return mde_metrics_sharded(pred, pred_dtype, target, n_img, hw, flags, ws, out_f64, out_f32, per_image_values,
                             per_image_raw, nullptr, 0u, stream);
}

// ---- peer-mapped mailboxes (one per rank; cudaMalloc'ed here so that the IPC handle covers exactly this block) ---------
extern "C" int mde_peer_alloc(size_t bytes, void** ptr_out) {
  using namespace mde;
  MDE_REQUIRE(ptr_out && bytes > 0, MDE_EINVAL, "null pointer");
  void* p = nullptr;
  MDE_CUDA_TRY(cudaMalloc(&p, bytes));
  MDE_CUDA_TRY(cudaMemset(p, 0, bytes));
  MDE_CUDA_TRY(cudaDeviceSynchronize());
  *ptr_out = p;
  return MDE_OK;
}
extern "C" int mde_peer_free(void* ptr) {
  using namespace mde;
  MDE_CUDA_TRY(cudaFree(ptr));
  return MDE_OK;
}
extern "C" int mde_peer_export(void* ptr, unsigned char* handle_out) {
  using namespace mde;
  static_assert(sizeof(cudaIpcMemHandle_t) == MDE_PEER_HANDLE_BYTES, "IPC handle size");
  MDE_REQUIRE(ptr && handle_out, MDE_EINVAL, "null pointer");
  cudaIpcMemHandle_t h;
  MDE_CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle_out, &h, sizeof(h));
  return MDE_OK;
}
extern "C" int mde_peer_open(const unsigned char* handle, void** ptr_out) {
  using namespace mde;
  MDE_REQUIRE(handle && ptr_out, MDE_EINVAL, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  MDE_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return MDE_OK;
}
extern "C" int mde_peer_allreduce_f64(double* src, double* dst, int n, int zero_src, void* comm, unsigned seq, void* ws, void* stream) {
  using namespace mde;
  MDE_REQUIRE(src && dst && comm && ws, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n >= 1 && n <= 31, MDE_EINVAL, "1 .. 31 doubles (one mailbox row)");
  MDE_REQUIRE(aligned_to(src, 8) && aligned_to(dst, 8), MDE_EALIGN, "misaligned pointer");
  PeerDesc* pd = static_cast<PeerDesc*>(comm);
  void* args[] = {&src, &dst, &n, &zero_src, &pd, &seq, &ws};
  MDE_CUDA_TRY(launch_pdl(reinterpret_cast<const void*>(&peer_allreduce_kernel), dim3(1), dim3(32), args, 0,
                          static_cast<cudaStream_t>(stream), false));
  count_launch();
  return MDE_OK;
}
extern "C" int mde_peer_comm_create(void* const* mailboxes, int rank, int world, unsigned timeout_ms, void** comm_out) {
  using namespace mde;
  MDE_REQUIRE(mailboxes && comm_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(world >= 1 && world <= MDE_MAX_PEERS && rank >= 0 && rank < world, MDE_EINVAL, "bad rank / world (one box: <= 8 peers)");
  PeerDesc h{};
  for (int r = 0; r < world; ++r) {
    MDE_REQUIRE(mailboxes[r] != nullptr && aligned_to(mailboxes[r], 8), MDE_EINVAL, "null / misaligned peer mailbox");
    h.mailbox[r] = static_cast<unsigned long long*>(mailboxes[r]);
  }
  h.rank = rank; h.world = world; h.timeout_ms = timeout_ms ? timeout_ms : 2000u; h.auto_seq = 0u;
  void* d = nullptr;
  MDE_CUDA_TRY(cudaMalloc(&d, sizeof(PeerDesc)));
  MDE_CUDA_TRY(cudaMemcpy(d, &h, sizeof(PeerDesc), cudaMemcpyHostToDevice));
  *comm_out = d;
  return MDE_OK;
}
extern "C" int mde_peer_comm_destroy(void* comm) {
  using namespace mde;
  MDE_CUDA_TRY(cudaFree(comm));
  return MDE_OK;
}
extern "C" int mde_peer_close(void* ptr) {
  using namespace mde;
  MDE_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return MDE_OK;
}

extern "C" void mde_metrics_finalize_host(const double* raw, double* values) { mde::metric_values(raw, values); }

extern "C" int mde_metrics_resized(const float* pred, int64_t pred_h, int64_t pred_w, const float* target, int64_t target_h,
                                   int64_t target_w, int64_t n_img, int64_t out_h, int64_t out_w, unsigned flags, void* ws,
                                   double* out_f64, float* out_f32, double* per_image_values, double* per_image_raw,
                                   void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && ws && out_f64, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && pred_h > 0 && pred_w > 0 && target_h > 0 && target_w > 0 && out_h > 0 && out_w > 0, MDE_EINVAL,
              "empty input");
  MDE_REQUIRE(out_h * out_w < (int64_t(1) << 30) && pred_h * pred_w < (int64_t(1) << 30) && target_h * target_w < (int64_t(1) << 30),
              MDE_ETOOBIG, "image too large");
  MDE_REQUIRE((flags & MDE_METRICS_REFERENCE_MATH) == 0, MDE_EINVAL, "the resized path uses the fast metric forms");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t cap = static_cast<int64_t>(sm_count()) * kCtasPerSm;
  int64_t cpi = cap / n_img;
  const int64_t max_cpi = (out_h * out_w + 4 * kBlock - 1) / (4 * kBlock);
  if (cpi > max_cpi) cpi = max_cpi;
  if (cpi < 1) cpi = 1;
  int64_t grid = n_img * cpi;
  if (grid > cap) grid = cap;
  metrics_resized_kernel<kGrpAll><<<static_cast<unsigned>(grid), kBlock, 0, st>>>(
      pred, static_cast<int>(pred_h), static_cast<int>(pred_w), target, static_cast<int>(target_h), static_cast<int>(target_w), n_img,
      static_cast<int>(out_h), static_cast<int>(out_w), static_cast<int>(cpi), ws, out_f64, out_f32, per_image_values, per_image_raw);
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
