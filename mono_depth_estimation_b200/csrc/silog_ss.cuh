// silog_ss.cuh - SILog (+ metric suite) forward+backward with the residuals parked in shared memory.
// Included by silog_ss.cu only (its own translation unit: the kernel is iterated on most).
#pragma once
#include "losses_kernel.cuh"

namespace mde {
namespace {

// ---- SILog (+ metric suite) with the residuals parked in SHARED memory between the phases ("SS") ----------
// fp32, 128-bit aligned, gradient requested, <= kSsSlots tiles per CTA (C1, C2). What the reduce loop went through
// (C2 fused SILog + 7 metrics, CUDA-event time per launch, measured on B200):
//   round 1   22.6 us  bulk copies by one producer thread, one mbarrier per slot, a CTA-wide __syncthreads per 2048-px
//                      tile, tiles claimed from a global counter: ~45 of ~205 instructions per quad were per-tile
//                      plumbing and only two tiles were in flight per CTA;
//   round 2a  20.0 us  thread-private cp.async pipeline, static interleaved tiles, lean per-pixel math (155 instructions
//                      per quad), all-reduce with an arrival counter - still 2 CTAs x 512 threads per SM: the warp
//                      scheduler serves the older CTA of an SM first, it finished its 8 tiles at 6.8 us while the
//                      younger one had done ~1.5 and then ran alone until 10-11 us;
//   round 2b  19.1 us  (this) ONE CTA of 1024 threads per SM: all 32 warps start together, 148 instead of 296 slots in
//                      the all-reduce (every thread spins on at most one), ordered first requests, p and d of the last
//                      kSsDepth tiles kept on chip for the gradient phase, shuffle-light warp reductions.
//   (Measured and dropped: per-warp chunks claimed from a shared-memory counter - the claims and spills cost 30 % more
//   instructions, 25.6 us; a rolled tile loop, 20.6 us; ring depth 2 / 3 / 5 instead of 4, within 0.5 us; an ordinary
//   instead of a cooperative launch, same time; default instead of streaming stores, +0.4 us.)
// Structure:
//   * THREAD-PRIVATE pipeline: every thread copies its own quad of pred and target with cp.async (LDGSTS,
//     16 B, no registers) kSsDepth tiles ahead and waits with cp.async.wait_group - no mbarrier, no CTA
//     barrier, no producer thread; warps drift freely;
//   * static interleaved tiles (tile = cta + k * grid): the slot index k is a compile-time constant of the
//     unrolled loop, so every shared-memory address is base + immediate and no tile list exists;
//   * the prediction quad lands directly in the slot that will hold its residuals (replaced in place),
//     the target quad in a ring of kSsDepth tiles (recycled by the same thread right after its LDS);
//   * per-pixel arithmetic in the lean form of metric_math.cuh (11 ALU-pipe instructions instead of ~20).
// Gradient phase: the residuals of a thread's LAST kSsDepth tiles went into the (by then free) ring slots instead of
// over the predictions, so p and d of those tiles are both still in shared memory; the predictions of the earlier
// tiles (<= 5 per thread) come back through L2 into registers, requested while the totals travel.
// Shared memory: (kSsSlots + kSsDepth) x 16 KB = 208 KB for the one CTA of an SM.
#ifndef MDE_SS_THREADS
#define MDE_SS_THREADS 1024
#endif
constexpr int kSsThreads = MDE_SS_THREADS;           // 1024: ONE CTA per SM (512: two)
constexpr int kSsWarps = kSsThreads / 32;
constexpr int kSsCtasPerSm = 1024 / kSsThreads;
constexpr int kSsSlots = 9;
constexpr int kSsDepth = 4;
constexpr size_t kSsBytes = static_cast<size_t>(kSsSlots + kSsDepth) * kSsThreads * sizeof(float4);

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one pixel of the rare path (exact reference arithmetic), booked into the lean accumulators; returns the stash
// value (log2 units). s0: sum d; s1x / nx: what the SILog mask (t > 0.01) adds to the suite's s_lnsq / n_valid.
template <unsigned MG>
static __device__ __noinline__ float4 ss_slow_px(float p, float t) {
  // .x = stash, .y = d (0 off the SILog mask), .z = d^2 - lnsq contribution (log2 units), .w = n(silog) - n(metric)
  const bool v2 = t > 0.01f;
  const float d2 = v2 ? log_ratio_slow(p, t) * 1.4426950408889634f : 0.f;
  float lnsq = 0.f, nm = 0.f;
  if constexpr (MG != 0) {
    const bool vm = t > 0.f;
    if (vm) {
      const float pp = (p < 1e-7f) ? 1e-7f : p;
      const float dl = logf(pp) - logf(t);
      lnsq = dl * dl * (1.0f / 0.48045301391820142f);
      nm = 1.f;
    }
  }
  float4 r;
  r.x = v2 ? d2 : __uint_as_float(kStashInvalid);
  r.y = d2;
  r.z = (MG != 0) ? d2 * d2 - lnsq : d2 * d2;
  r.w = (v2 ? 1.f : 0.f) - nm;
  return r;
}

// a whole rare quad, out of line: the unrolled reduce loop carries one CALL per tile instead of eight
// f[0..2] = sum d, d^2 minus the suite's lnsq contribution, n(silog) - n(metric); f[3..10] = the eight metric
// sums in lean units (MetricTile order); c = exact n / c1 / c2 / c3
template <unsigned MG, unsigned REFG>
static __device__ __noinline__ void ss_rare_quad(const float4 p4, const float4 t4, float4& d_out, float (&f)[3 + 8], int (&c)[4]) {
  const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, tv[4] = {t4.x, t4.y, t4.z, t4.w};
  float dv[4];
#pragma unroll
  for (int i = 0; i < 11; ++i) f[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if constexpr (MG != 0) {
      const MetricContrib r = metric_px_ref_contrib<REFG>(pv[j], tv[j]);
      f[3] += r.s.s_abs; f[4] += r.s.s_sq;
      f[5] += r.s.s_log10 * (1.0f / tile_scale<false>(2));
      f[6] += r.s.s_sle * (1.0f / tile_scale<false>(3));
      f[7] += r.s.s_absrel; f[8] += r.s.s_sqrel; f[9] += r.s.s_rsq;
      f[10] += r.s.s_lnsq * (1.0f / tile_scale<false>(7));
      c[0] += r.c.n; c[1] += r.c.c1; c[2] += r.c.c2; c[3] += r.c.c3;
    }
    const float4 r = ss_slow_px<MG>(pv[j], tv[j]);
    dv[j] = r.x;
    f[0] += r.y;
    f[1] += r.z;
    f[2] += r.w;
  }
  d_out = make_float4(dv[0], dv[1], dv[2], dv[3]);
}

// Phase trace. Product build: common.cuh's trace_point (thread 0 stamps global memory when the trace is armed).
// Instrumented twin (-DMDE_SS_TIMING, tools/libmde_dbg.so): the stamps are parked in shared memory and written out by
// the LAST warp of the CTA to leave, slot 5 being the latest exit over ALL warps, into the half of the buffer selected by
// the launch parity - two consecutive launches are read back together, which shows the launch-to-launch period and the
// gap between one grid's last exit and the next grid's first CTA (buffer: 2 x 296 x kTraceSlots words).
#ifdef MDE_SS_TIMING
#define SS_TP(k)                                                              \
  do {                                                                        \
    if ((k) != 5 && threadIdx.x == 0) {                                       \
      unsigned long long ns_;                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_));                 \
      sm_trace[k] = ns_;                                                      \
    }                                                                         \
  } while (0)
#else
#define SS_TP(k) trace_point(k)
#endif

template <unsigned MG>
__global__ void __launch_bounds__(kSsThreads, kSsCtasPerSm) silog_ss_kernel(LossArgs a) {
  __shared__ double sm_d[(MG ? 12 : 4) * kSsWarps];
  __shared__ double sm_own[4 * kSsWarps];
  __shared__ double sm_tot[4];
  __shared__ double sm_gather[kSsWarps * 4];
  __shared__ float sm_k[4];
  __shared__ unsigned sm_epoch;
#ifdef MDE_SS_TIMING
  __shared__ unsigned long long sm_trace[kTraceSlots];
  __shared__ unsigned sm_tcnt;
  if (threadIdx.x == 0) {
    sm_tcnt = 0u;
    sm_trace[5] = 0ull;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    sm_trace[7] = smid;
  }
#endif
  extern __shared__ float4 sm_ss[];   // [kSsSlots][kSsThreads] residual slots, then [kSsDepth][kSsThreads] target ring
  // metric groups evaluated in reference arithmetic on the rare path (kGrpRsq is a lean-form subset of kGrpRel)
  constexpr unsigned kRefG = (MG & kGrpRsq) ? ((MG & 7u) | kGrpRel) : (MG & 7u);

  const float* __restrict__ pred = static_cast<const float*>(a.pred);
  const float* __restrict__ gt = a.gt;
  float* grad = static_cast<float*>(a.grad);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = static_cast<int>(gridDim.x), cta = static_cast<int>(blockIdx.x);
  const int nq = static_cast<int>(a.n >> 2);
  float4* slots = sm_ss;
  float4* ring = sm_ss + kSsSlots * kSsThreads;
  const int q0 = cta * kSsThreads + tid;                         // this thread's quad in tile k: q0 + k * qs
  const int qs = G * kSsThreads;
  const int nst = (q0 < nq) ? (nq - 1 - q0) / qs + 1 : 0;    // quads of this thread (ns, or ns - 1 in a partial last tile)

  pdl_wait();   // launched with launch_pdl: nothing a predecessor wrote may be read before this point
  SS_TP(0);
  // The launch parity (workspace epoch) is needed only after the reduce loop. Its load is issued FIRST: the L1
  // returns loads in issue order, so behind the prefetch burst below it would come back after ~128 KB of copies
  // (measured: the loop of the second CTA of an SM started 3.7 us into the kernel while the prologue waited for it).
  Ws ws = ws_view(a.ws);
  unsigned epoch_reg = 0u;
  if (tid == 0) epoch_reg = __ldcg(&ws.hdr->epoch);
  auto issue = [&](int k) {
    const size_t q = static_cast<size_t>(q0) + static_cast<size_t>(k) * static_cast<size_t>(qs);
    cp_async16(&slots[k * kSsThreads + tid], pred + 4 * q);
    cp_async16(&ring[(k % kSsDepth) * kSsThreads + tid], gt + 4 * q);
  };
  // The warp scheduler serves the warps in a fixed priority order, so without the two barriers the favoured warp would
  // queue ALL its kSsDepth tiles before the last one has asked for its first: tile 0 of the whole CTA goes out first
  // (its data is back ~1 us earlier for the late warps), then tile 1, then the rest.
#pragma unroll
  for (int k = 0; k < kSsDepth; ++k) {
    if (k < nst) issue(k);
    cp_async_commit();
    if (k < 2) __syncthreads();
  }

  // ---------------- reduce phase ------------------------------------------------------------------------
  MetricAcc acc;
  acc.zero();
  float s0 = 0.f, s1 = 0.f;    // sum d; MG == 0: sum d^2, MG != 0: what the rare path adds to the suite's s_lnsq
  float nx = 0.f;              // MG == 0: valid count; MG != 0: rare-path difference n(silog) - n(metric)
  int lean_q = 0;              // quads evaluated in the lean form
  auto quad = [&](const float4& p4, const float4& t4) -> float4 {
    float4 d;
    bool rare;
    if constexpr (MG != 0) {
      // valid target <= 0.01 (SILog mask differs from the metric mask; includes subnormals) or prediction < 1e-7 / NaN
      const float ta = (t4.x > 0.f) ? t4.x : 1.0f, tb = (t4.y > 0.f) ? t4.y : 1.0f;
      const float tc = (t4.z > 0.f) ? t4.z : 1.0f, td = (t4.w > 0.f) ? t4.w : 1.0f;
      rare = !(fminf(fminf(ta, tb), fminf(tc, td)) > 0.01f) || !(fminf(fminf(p4.x, p4.y), fminf(p4.z, p4.w)) >= 1e-7f);
    } else {
      rare = !(fminf(fminf(p4.x, p4.y), fminf(p4.z, p4.w)) >= 1.17549435e-38f);   // MUFU.LG2 flushes subnormals
    }
    if (rare) {
      float f[11];
      int c[4];
      ss_rare_quad<MG, kRefG>(p4, t4, d, f, c);
      s0 += f[0]; s1 += f[1]; nx += f[2];
      if constexpr (MG != 0) {
        acc.s_abs += f[3]; acc.s_sq += f[4]; acc.s_log10 += f[5]; acc.s_sle += f[6];
        acc.s_absrel += f[7]; acc.s_sqrel += f[8]; acc.s_rsq += f[9]; acc.s_lnsq += f[10];
        acc.n_x += c[0]; acc.c1_x += c[1]; acc.c2_x += c[2]; acc.c3_x += c[3];
      }
    } else if constexpr (MG != 0) {
      // here the SILog mask equals the metric mask and the residual equals the suite's log2 p - log2 t:
      // the loss adds ONE accumulation (sum d); sum d^2 is the suite's s_lnsq, n its valid count
      const float dx = metric_px_lean<MG, true>(p4.x, t4.x, acc), dy = metric_px_lean<MG, true>(p4.y, t4.y, acc);
      const float dz = metric_px_lean<MG, true>(p4.z, t4.z, acc), dw = metric_px_lean<MG, true>(p4.w, t4.w, acc);
      s0 += (dx + dy) + (dz + dw);
      ++lean_q;
      const float inval = __uint_as_float(kStashInvalid);
      d = make_float4(t4.x > 0.f ? dx : inval, t4.y > 0.f ? dy : inval, t4.z > 0.f ? dz : inval, t4.w > 0.f ? dw : inval);
    } else {
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, tv[4] = {t4.x, t4.y, t4.z, t4.w};
      float dv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool v = tv[j] > 0.01f;                        // criteria.py:730
        const float dd = mufu_lg2(v ? pv[j] : 1.0f) - mufu_lg2(v ? tv[j] : 1.0f);
        s0 += dd;
        s1 = fmaf(dd, dd, s1);
        nx += v ? 1.f : 0.f;
        dv[j] = v ? dd : __uint_as_float(kStashInvalid);
      }
      d = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
    return d;
  };
#ifdef MDE_SS_TIMING
  long long tm_wait = 0, tm_comp = 0;   // cycles this warp spent waiting for its copies / evaluating its quads
#endif
#pragma unroll   // (a rolled loop with computed slot addresses was measured: 20.6 instead of 20.0 us)
  for (int k = 0; k < kSsSlots; ++k) {
    if (k < nst) {                      // uniform over the CTA except in a partial last tile
#ifdef MDE_SS_TIMING
      const long long tm0 = clock64();
#endif
      cp_async_wait<kSsDepth - 1>();    // this thread's copies of tile k have landed
#ifdef MDE_SS_TIMING
      const long long tm1 = clock64();
#endif
      const float4 p4 = slots[k * kSsThreads + tid];
      const float4 t4 = ring[(k % kSsDepth) * kSsThreads + tid];
      const bool refill = k + kSsDepth < nst;
      if (refill) issue(k + kSsDepth);   // the ring slot just read is this thread's to refill
      cp_async_commit();
      // residuals: over the prediction while the ring slot is needed again; for the LAST kSsDepth tiles into the ring
      // slot (free from here on), so that both p and d of those tiles are still in shared memory in the gradient phase
      float4* dst = refill ? &slots[k * kSsThreads + tid] : &ring[(k % kSsDepth) * kSsThreads + tid];
      *dst = quad(p4, t4);
#ifdef MDE_SS_TIMING
      const long long tm2 = clock64();
      tm_wait += tm1 - tm0;
      tm_comp += tm2 - tm1;
#endif
    }
  }
#ifdef MDE_SS_TIMING
  (void)tm_wait; (void)tm_comp;
#endif
  cp_async_wait<0>();
  if (cta == G - 1) {   // n % 4 tail: summed here, its gradient is recomputed below
    const int64_t i = (static_cast<int64_t>(nq) << 2) + tid;
    if (i < a.n) {
      const float p = __ldg(pred + i), t = __ldg(gt + i);
      if constexpr (MG != 0) metric_add_contrib(metric_px_ref_contrib<kRefG>(p, t), acc);
      const float4 r = ss_slow_px<MG>(p, t);
      s0 += r.y;
      s1 += r.z;
      nx += r.w;
    }
  }
  SS_TP(1);
  const int lean_px = 4 * lean_q;
  {
    // per-warp loss totals: one shuffle-light multi-sum in fp32 (a thread saw <= 36 pixels; the warp total of sum d and
    // sum d^2 carries ~1e-7 relative rounding, the count is exact), widened to fp64 across warps and CTAs
    float run[4];
    run[0] = s0;
    run[1] = s1 + ((MG != 0) ? acc.s_lnsq : 0.f);
    run[2] = nx + ((MG != 0) ? static_cast<float>(acc.n_valid(lean_px)) : 0.f);
    run[3] = 0.f;
    const float tot = warp_multi_sum<4>(run);              // quantity (lane >> 3) & 3
    if ((lane & 7) == 0) sm_own[(lane >> 3) * kSsWarps + warp] = static_cast<double>(tot);
    if (tid == 0) sm_epoch = epoch_reg;
    __syncthreads();
  }
  // launch parity: this launch uses workspace set `par`; CTA 0 cleans the OTHER set (used by the previous
  // cooperative launch, which has completed) for the next one - what coop_prologue does, moved behind the loop
  const unsigned epoch = sm_epoch;
  const int par = static_cast<int>(epoch & 1u);
  double* gacc = ws.gacc + par * kGacc;
  unsigned* ukey = ws.ukey + par * kUkey;
  if (cta == 0) {
    const int o = par ^ 1;
    for (int i = tid; i < kGacc; i += kSsThreads) ws.gacc[o * kGacc + i] = 0.0;
    for (int i = tid; i < kUkey; i += kSsThreads) ws.ukey[o * kUkey + i] = 0u;
    if (tid == 0) {
      const unsigned dirty = __ldcg(&ws.hdr->dirty[o]);
      if (dirty) {
        const unsigned cap = __ldcg(&ws.hdr->max_images);
        double* rows = ws.iacc + static_cast<size_t>(1 + o) * cap * kIacc;
        for (size_t i = 0; i < static_cast<size_t>(dirty) * kIacc; ++i) rows[i] = 0.0;
        ws.hdr->dirty[o] = 0u;
      }
    }
  }
  // pooled metric sums of this CTA -> 12 fp64 atomics (exact integer counts through REDUX, float sums through a
  // 32-lane fp32 tree, widened before crossing warps and CTAs); runs while the all-reduce slots travel
  auto flush_metrics = [&] {
    if constexpr (MG != 0) {
      const int r0 = __reduce_add_sync(0xffffffffu, acc.n_valid(lean_px)), r1 = __reduce_add_sync(0xffffffffu, acc.count(1, lean_px));
      const int r2 = __reduce_add_sync(0xffffffffu, acc.count(2, lean_px)), r3 = __reduce_add_sync(0xffffffffu, acc.count(3, lean_px));
      if (lane == 0) {
        sm_d[0 * kSsWarps + warp] = r0; sm_d[1 * kSsWarps + warp] = r1;
        sm_d[2 * kSsWarps + warp] = r2; sm_d[3 * kSsWarps + warp] = r3;
      }
      float v8[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v8[q] = acc.sum(q);
      const float sq = warp_multi_sum<8>(v8);               // quantity (lane >> 2) & 7: 9 shuffles instead of 40
      if ((lane & 3) == 0) {
        const int q = lane >> 2;
        const float sc = (q == 2) ? tile_scale<false>(2) : ((q == 3 || q == 7) ? tile_scale<false>(3) : 1.0f);
        sm_d[(4 + q) * kSsWarps + warp] = static_cast<double>(sq * sc);
      }
      __syncthreads();
      // (pooling by 12 WARPS with a shuffle tree each instead of this serial pass of 12 threads - what the small-input
      // kernel of resident_loss.cuh does - was measured on one box in both orders: 19.16 against 18.38 us)
      if (tid < 12) {
        double tot = 0.0;
        for (int w = 0; w < kSsWarps; ++w) tot += sm_d[tid * kSsWarps + w];
        const int qi = (tid < 4) ? tid : kTileToQ[tid - 4];
        if (tot != 0.0) atomicAdd(&gacc[kMetBase + qi], tot);
      }
    }
  };
  SS_TP(2);

  // ---------------- all-reduce of the totals (grid_sum4_counted, common.cuh); every CTA derives the coefficients ----
  constexpr int kRegTiles = kSsSlots - kSsDepth;   // gradient-phase predictions that are read again (through L2)
  // The early tiles' predictions were overwritten by their residuals: request them again (L2 hits) while the totals
  // travel - the gradient phase is bound by L2 bandwidth (20 MB of stores), so these reads are better out of its way.
  float4 preg[kRegTiles];
  auto prefetch_pred = [&] {
#pragma unroll
    for (int k = 0; k < kRegTiles; ++k) {
      const size_t q = static_cast<size_t>(q0) + static_cast<size_t>(k) * static_cast<size_t>(qs);
      if (k + kSsDepth < nst) preg[k] = __ldcs(reinterpret_cast<const float4*>(pred + 4 * q));
    }
  };
  grid_sum4_counted<kSsWarps>(ws.slots, ukey + 2, epoch * 4u + 2u, sm_own, sm_gather, sm_tot, [&] {
    if (grad != nullptr) prefetch_pred();
    flush_metrics();
#ifdef MDE_SS_TIMING
    SS_TP(4);   // (instrumented build) slot 4 = this CTA starts waiting for the other CTAs' totals
#endif
  });
#ifdef MDE_SS_TIMING
  SS_TP(6);     // (instrumented build) slot 6 = totals gathered and reduced
#endif
  if (tid < 32) __syncwarp();   // sm_tot was written by threads 0..3
  if (tid == 0) {
    // totals (log2 units) -> loss value and gradient coefficients. Every CTA waits for this chain, so the coefficients
    // come from the SFU (reciprocal and reciprocal square root, one Newton step each: ~1e-7 relative, the gradient
    // tolerance is 1e-5) instead of two IEEE divisions and a square root (-0.14 us); degenerate totals (no valid pixel,
    // zero or negative variance) take the exact operations, so that NaN / inf come out as they always did.
    const double S0 = sm_tot[0] * 0.69314718055994531, S1 = sm_tot[1] * 0.48045301391820142, N0 = sm_tot[2];
    double inv = static_cast<double>(mufu_rcp(static_cast<float>(N0)));   // + one Newton step in fp64
    inv = inv * (2.0 - N0 * inv);                                        // (relative error ~1e-14; N0 == 0 gives NaN as 1.0 / 0 * 0 does downstream)
    const double dm = S0 * inv, qm = S1 * inv;
    const double var = qm - static_cast<double>(a.vf) * dm * dm;   // the cancellation stays in fp64
    const float varf = static_cast<float>(var);
    float k1;                                                        // dL/dd_i = k1 * (d_i - k2), natural log
    if (varf > 1e-30f && varf < 1e30f) {
      float rs = mufu_rsq(varf);
      rs = rs * fmaf(-0.5f * varf, rs * rs, 1.5f);
      k1 = 10.0f * a.grad_scale * static_cast<float>(inv) * rs;
    } else {
      k1 = 10.0f * a.grad_scale * static_cast<float>(inv) / sqrtf(varf);
    }
    sm_k[0] = k1;
    sm_k[1] = a.vf * static_cast<float>(dm);
    sm_k[2] = k1 * 0.69314718055994531f;                             // same, for residuals held in log2 units
    sm_k[3] = a.vf * static_cast<float>(dm * 1.4426950408889634);
    if (cta == 0) {
      // (forming the loss value at the END of the kernel instead was measured: +0.75 us - a serial chain and a store
      // right before the exit delay the grid's completion)
      const double loss = 10.0 * static_cast<double>(sqrtf(varf));
      *a.loss_out = static_cast<float>(loss);
      if (a.totals_out) {
        a.totals_out[0] = S0; a.totals_out[1] = S1; a.totals_out[2] = N0; a.totals_out[3] = 0.0;
        a.totals_out[4] = 0.0; a.totals_out[5] = loss;
      }
      ws.hdr->epoch = epoch + 1u;
    }
  }
  __syncthreads();
  if constexpr (MG != 0) {
    // arrival for the metric finaliser (end of the kernel), off everybody's critical path: the CTA's metric
    // atomics (flush_metrics) were issued before the __syncthreads above, so this fence orders them
    if (tid == kSsThreads - 32) {
      __threadfence();
      atomicAdd(ukey + 6, 1u);
    }
  }
  SS_TP(3);
  pdl_trigger();   // a dependent launch may start filling the SMs this grid leaves
  // pooled metric values (one mean over all valid pixels of the call, metrics.py:58-67), formed by the LAST warp
  // of the LAST CTA at the very end of the kernel, when every CTA's arrival has long been counted
  auto finalize_metrics = [&] {
    if constexpr (MG != 0) {
      if (cta == G - 1 && tid >= kSsThreads - 32) {
        if (lane == 0) {   // the all-reduce above is no memory barrier: wait for every CTA's metric atomics
          while (*reinterpret_cast<volatile unsigned*>(ukey + 6) < gridDim.x) {}
          __threadfence();
        }
        __syncwarp();
        const bool own = lane < MDE_METRIC_NM;
        const double P = own ? __ldcg(&gacc[kMetBase + lane]) : 0.0;
        const double nn = __shfl_sync(0xffffffffu, P, MDE_Q_NVALID);
        const double num = __shfl_sync(0xffffffffu, P, own ? kValNumL[lane] : 0);
        double val = num / nn;
        if (lane >= MDE_M_RMSE_TRUE) val = sqrt(val);
        if (own) {
          const double im = (a.n_img == 1) ? val : __longlong_as_double(0x7ff8000000000000LL);
          a.met_f64[lane] = val;
          a.met_f64[MDE_METRIC_NM + lane] = im;   // per-image means are not formed by the fused path
          a.met_f64[2 * MDE_METRIC_NM + lane] = P;
          a.met_f64[2 * MDE_METRIC_NM + MDE_METRIC_NQ + 1 + lane] = im;   // per-image value sums (one image: the values)
          if (a.met_f32) {
            a.met_f32[lane] = static_cast<float>(val);
            a.met_f32[MDE_METRIC_NM + lane] = static_cast<float>(im);
          }
          if (a.met_accum) a.met_accum[lane] += static_cast<float>(val);   // MetricComputation's running sums
          if (a.met_raw_accum) a.met_raw_accum[lane] += P;
        }
        if (lane == 0) a.met_f64[2 * MDE_METRIC_NM + MDE_METRIC_NQ] = (a.n_img == 1 && nn > 0.0) ? 1.0 : __longlong_as_double(0x7ff8000000000000LL);
      }
    }
  };
  auto trace_exit = [&] {
#ifdef MDE_SS_TIMING
    unsigned long long* t = g_mde_trace;
    if (lane == 0 && t != nullptr) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      atomicMax(&sm_trace[5], ns);
      __threadfence_block();
      if (atomicAdd(&sm_tcnt, 1u) == static_cast<unsigned>(kSsWarps - 1)) {
        unsigned long long* row = t + (static_cast<size_t>(epoch & 1u) * 296 + static_cast<size_t>(cta)) * kTraceSlots;
        for (int k = 0; k < kTraceSlots; ++k) row[k] = *reinterpret_cast<volatile unsigned long long*>(&sm_trace[k]);
      }
    }
#endif
  };
  if (grad == nullptr) {
    finalize_metrics();
    trace_exit();
    return;
  }
#ifndef MDE_SS_TIMING
  SS_TP(4);
#endif

  // ---------------- gradient phase: g_i = k (d_i - c) / p_i on the mask, 0 elsewhere -------------------------
  const float ks1 = sm_k[2], ks12 = -sm_k[2] * sm_k[3];   // g = (ks1 d - ks1 ks2) / p; MUFU.RCP alone: ~1 ulp, tolerance 1e-5
  auto gquad = [&](const float4& p, const float4& d) -> float4 {
    float4 g;
    g.x = (__float_as_uint(d.x) != kStashInvalid) ? fmaf(d.x, ks1, ks12) * mufu_rcp(p.x) : 0.f;
    g.y = (__float_as_uint(d.y) != kStashInvalid) ? fmaf(d.y, ks1, ks12) * mufu_rcp(p.y) : 0.f;
    g.z = (__float_as_uint(d.z) != kStashInvalid) ? fmaf(d.z, ks1, ks12) * mufu_rcp(p.z) : 0.f;
    g.w = (__float_as_uint(d.w) != kStashInvalid) ? fmaf(d.w, ks1, ks12) * mufu_rcp(p.w) : 0.f;
    return g;
  };
  // the last kSsDepth tiles: p (slot) and d (ring) are both in shared memory
#pragma unroll
  for (int k = 0; k < kSsSlots; ++k) {
    if (k < nst && k + kSsDepth >= nst) {
      const size_t q = static_cast<size_t>(q0) + static_cast<size_t>(k) * static_cast<size_t>(qs);
      __stcs(reinterpret_cast<float4*>(grad + 4 * q), gquad(slots[k * kSsThreads + tid], ring[(k % kSsDepth) * kSsThreads + tid]));
    }
  }
  // the early tiles: d from the slot, p from the registers requested before the wait
#pragma unroll
  for (int k = 0; k < kRegTiles; ++k) {
    if (k + kSsDepth < nst) {
      const size_t q = static_cast<size_t>(q0) + static_cast<size_t>(k) * static_cast<size_t>(qs);
      __stcs(reinterpret_cast<float4*>(grad + 4 * q), gquad(preg[k], slots[k * kSsThreads + tid]));
    }
  }
  if (cta == G - 1) {   // n % 4 tail, recomputed in natural-log units
    const int64_t i = (static_cast<int64_t>(nq) << 2) + tid;
    if (i < a.n) {
      bool v;
      const float p = __ldg(pred + i);
      const float d = silog_resid(p, __ldg(gt + i), v);
      grad[i] = v ? sm_k[0] * (d - sm_k[1]) * rcp_nr(p) : 0.f;
    }
  }
  SS_TP(5);
  finalize_metrics();
  trace_exit();
}

// SILog / fp32 / 128-bit path with a gradient and few enough tiles: residuals parked in shared memory
template <unsigned MG>
int launch_loss_ss_mg(LossArgs& a, cudaStream_t st, bool& taken) {
  const void* fn = reinterpret_cast<const void*>(&silog_ss_kernel<MG>);
  const int cap = coop_grid(fn, kSsThreads, kSsBytes);
  if (cap <= 0) return MDE_OK;   // (e.g. the carve-out is not available) -> generic path
  const int64_t nq = a.n >> 2;
  const int64_t nt = (nq + kSsThreads - 1) / kSsThreads;
  int64_t grid = nt < cap ? nt : cap;
  if (grid < 1) grid = 1;
  if (nt > grid * kSsSlots) return MDE_OK;
  a.chunk = make_chunking(nq, 8, static_cast<int>(grid));
  void* args[] = {&a};
  // (an ordinary launch of the same grid was measured: 19.96 vs 20.01 us - the cooperative launch costs nothing)
  MDE_CUDA_TRY(launch_pdl(fn, dim3(static_cast<unsigned>(grid)), dim3(kSsThreads), args, kSsBytes, st, true));
  count_launch();
  taken = true;
  return MDE_OK;
}
template <unsigned MG>
int launch_loss_ss(LossArgs& a, cudaStream_t st, bool& taken) {
  taken = false;
  static const bool off = [] { const char* e = getenv("MDE_NO_SMEM_STASH"); return e && atoi(e) != 0; }();
  if (off) return MDE_OK;
  // the reference's default metric list needs only the 'rmse' sum of the REL group: leaner instantiation
  if constexpr (MG == (kGrpLog | kGrpRel)) {
    if (a.rsq_only) return launch_loss_ss_mg<(kGrpLog | kGrpRsq)>(a, st, taken);
  }
  return launch_loss_ss_mg<MG>(a, st, taken);
}

}  // namespace
}  // namespace mde
