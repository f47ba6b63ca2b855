// common.cuh - shared device/host helpers of libmde_b200 (sm_100a only).
//
// Design notes:
//  * every kernel here is HBM/L2-streaming fp32 work: 128-bit coalesced loads, per-thread fp32
//    accumulation over one tile, fp64 accumulation across tiles / warps / CTAs, warp-shuffle +
//    shared-memory block reduction, one fp64 atomic per CTA per quantity;
//  * grids are sized from the SM count (148 on B200) times the resident CTAs per SM, each CTA
//    owning one CONTIGUOUS chunk of the tensor (128-byte aligned) - persistent style, no tail wave;
//  * losses that need totals before the gradient run as ONE cooperative launch with a grid-wide
//    barrier between the reduce and gradient phases; the gradient phase walks the chunk backwards
//    so the most recently read lines are re-read first (L2 LRU friendly).
#pragma once

#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mde_b200.h"

namespace cg = cooperative_groups;

namespace mde {

constexpr int kBlock = 512;      // threads per CTA of the streaming kernels
constexpr int kWarps = kBlock / 32;
constexpr int kCtasPerSm = 2;    // 1024 threads/SM -> 64 registers/thread budget

// ---- host side: errors, launch accounting, device info -------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();
// max co-resident CTAs for a cooperative launch of `func` (cached per function)
int coop_grid(const void* func, int block, size_t smem);
// Launch with programmatic dependent launch allowed (on unless MDE_PDL=0): the grid may be scheduled while the previous
// kernel of the stream drains, so a kernel launched this way must execute griddepcontrol.wait (pdl_wait()) before its
// first global-memory access; ~1 us of launch latency per call hides behind the predecessor. `cooperative` adds the
// cooperative attribute. Falls back to the ordinary (cooperative) launch if the driver refuses the combination.
cudaError_t launch_pdl(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t st, bool cooperative);

#define MDE_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::mde::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                       \
      return MDE_ECUDA;                                                                 \
    }                                                                                   \
  } while (0)

#define MDE_REQUIRE(cond, code, msg)                        \
  do {                                                      \
    if (!(cond)) {                                          \
      ::mde::set_error("%s: %s", __func__, msg);            \
      return code;                                          \
    }                                                       \
  } while (0)

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- workspace layout -----------------------------------------------------------------------
// [0,256)      header: epoch (parity of cooperative launches), ticket (last-CTA detection),
//              dirty[2] (# per-image rows used in each cooperative parity set), capacity, error
//              flag, tacc[16] (accumulators of the single-pass ticket kernels, self-cleaned)
// [256, +1024) gacc[2][64] doubles : global accumulators of cooperative kernels, two parity sets
// [1280,+128)  ukey[2][16] unsigned: order-preserving float max keys / flags, two parity sets
// [1536, ...)  iacc[3][max_images][16] doubles : per-image accumulators; region 0 belongs to the
//              metrics kernel (self-cleaned), regions 1,2 are the cooperative parity sets
struct WsHeader {
  unsigned epoch;
  unsigned ticket;
  unsigned dirty[2];
  unsigned max_images;
  unsigned error;
  unsigned pad0[2];
  double tacc[16];
  unsigned long long bcast[8];  // grid_barrier_bcast words {seq:32 | payload:32}; never need zeroing
  unsigned pad1[8];
};
static_assert(sizeof(WsHeader) == 256, "workspace header must be 256 bytes");
constexpr int kGacc = 64;
constexpr int kUkey = 16;
constexpr int kIacc = 16;
// [1536, +24576) slots[384][4][2] words {seq:32 | half of a double:32}: per-CTA partial totals of
//              grid_sum4_bcast (self-validating, unique seq per launch and stage: never need zeroing)
constexpr int kSlotCtas = 384;                       // >= resident CTAs of any cooperative launch here (2 x 148)
constexpr size_t kSlotBytes = static_cast<size_t>(kSlotCtas) * 4 * 2 * 8;
constexpr size_t kWsHeadBytes = 256 + 2 * kGacc * 8 + 2 * kUkey * 4 + 128;  // 1536
// [26112, +49152) mslots[192][32] words {seq:32 | half of a double:32}: per-CTA pooled metric sums of the register-resident
//              small-input loss kernel (resident_loss.cuh), gathered by its finalising CTA (self-validating like `slots`)
constexpr int kMetSlotCtas = 192;
constexpr int kMetSlotWords = 32;
constexpr size_t kMetSlotBytes = static_cast<size_t>(kMetSlotCtas) * kMetSlotWords * 8;
constexpr size_t kWsFixedBytes = kWsHeadBytes + kSlotBytes + kMetSlotBytes;

struct Ws {
  WsHeader* hdr;
  double* gacc;     // [2][kGacc]
  unsigned* ukey;   // [2][kUkey]
  double* iacc;     // [3][max_images][kIacc]
  unsigned long long* slots;  // [kSlotCtas][4][2]
  unsigned long long* mslots; // [kMetSlotCtas][kMetSlotWords]
};

__host__ __device__ inline Ws ws_view(void* base) {
  Ws w;
  char* b = static_cast<char*>(base);
  w.hdr = reinterpret_cast<WsHeader*>(b);
  w.gacc = reinterpret_cast<double*>(b + 256);
  w.ukey = reinterpret_cast<unsigned*>(b + 256 + 2 * kGacc * 8);
  w.slots = reinterpret_cast<unsigned long long*>(b + kWsHeadBytes);
  w.mslots = reinterpret_cast<unsigned long long*>(b + kWsHeadBytes + kSlotBytes);
  w.iacc = reinterpret_cast<double*>(b + kWsFixedBytes);
  return w;
}

#ifdef __CUDACC__
// Prologue of every cooperative kernel: read the launch parity and let CTA 0 clean the OTHER
// parity set (used by the previous cooperative launch, which has completed) for the next one.
// Must run before the first grid barrier; the matching epilogue (epoch + 1) after the last one.
__device__ __forceinline__ int coop_prologue(const Ws& ws, unsigned& epoch_out) {
  // one thread per CTA reads the epoch (a hot line if every warp of the grid fetched it) and shares it
  __shared__ unsigned sm_epoch;
  if (threadIdx.x == 0) sm_epoch = __ldcg(&ws.hdr->epoch);
  __syncthreads();
  const unsigned epoch = sm_epoch;
  const int par = static_cast<int>(epoch & 1u);
  epoch_out = epoch;
  if (blockIdx.x == 0) {
    const int o = par ^ 1;
    for (int i = threadIdx.x; i < kGacc; i += blockDim.x) ws.gacc[o * kGacc + i] = 0.0;
    for (int i = threadIdx.x; i < kUkey; i += blockDim.x) ws.ukey[o * kUkey + i] = 0u;
    if (threadIdx.x == 0) {
      const unsigned dirty = __ldcg(&ws.hdr->dirty[o]);
      if (dirty) {
        const unsigned cap = __ldcg(&ws.hdr->max_images);
        double* rows = ws.iacc + static_cast<size_t>(1 + o) * cap * kIacc;
        for (size_t i = 0; i < static_cast<size_t>(dirty) * kIacc; ++i) rows[i] = 0.0;
        ws.hdr->dirty[o] = 0u;
      }
    }
  }
  return par;
}
#endif

#ifdef __CUDACC__
// programmatic dependent launch: nothing a predecessor in the stream wrote may be read before pdl_wait(); pdl_trigger()
// lets a dependent grid start filling the SMs this grid leaves (both are no-ops for ordinary launches)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---- grid barrier that also broadcasts what everybody needs next -------------------------------------
// A cooperative-groups grid.sync() followed by "every CTA reads the totals and derives its
// coefficients" costs three dependent L2 round trips after the last CTA arrives (barrier counter,
// totals, then the first data load) plus a hot-spot of ~4700 warps fetching one line. Here the LAST
// CTA to arrive (ticket == grid-1) reads the totals, computes up to four fp32 values once and stores
// them as four self-validating 8-byte words {seq | value}; the other CTAs spin on exactly those words,
// so leaving the barrier and receiving the coefficients is one and the same L2 read. `seq` is unique
// per launch and stage, so the words never need clearing. Release/acquire through the ticket and the
// words orders every CTA's earlier global writes before every CTA's later reads (read such data with
// ld.global.cg: the L1 of this SM is not invalidated). Needs a cooperative launch (co-residency).
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// last_fn(v, post): fills v[0..3]; anything it wants to write besides (loss value, ...) goes into the
// `post` callable it returns control to AFTER the words are out, i.e. off the other CTAs' critical path.
// mid_fn() runs on every thread after the CTA's ticket has been ISSUED and before anybody waits: loads
// started there overlap the barrier's round trips instead of delaying the release fence.
template <typename LastFn, typename PostFn, typename MidFn>
__device__ __forceinline__ void grid_barrier_bcast(unsigned* ticket, unsigned long long* words, unsigned seq,
                                                   float* out_smem4, LastFn&& last_fn, PostFn&& post_fn, MidFn&& mid_fn) {
  __syncthreads();
  unsigned t = 0u;
  if (threadIdx.x == 0) {
    __threadfence();                       // release: this CTA's global writes before its ticket
    t = atomicAdd(ticket, 1u);
  }
  mid_fn();
  if (threadIdx.x == 0) {
    const bool last = (t == gridDim.x - 1);
    if (last) {
      __threadfence();                     // acquire every CTA's writes; also orders them before the words
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      last_fn(v);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        st_relaxed_u64(words + i, (static_cast<unsigned long long>(seq) << 32) | __float_as_uint(v[i]));
      post_fn();
    }
    unsigned long long w0, w1, w2, w3;
    do {
      w0 = ld_relaxed_u64(words + 0);
      w1 = ld_relaxed_u64(words + 1);
      w2 = ld_relaxed_u64(words + 2);
      w3 = ld_relaxed_u64(words + 3);
    } while (static_cast<unsigned>(w0 >> 32) != seq || static_cast<unsigned>(w1 >> 32) != seq ||
             static_cast<unsigned>(w2 >> 32) != seq || static_cast<unsigned>(w3 >> 32) != seq);
    __threadfence();                       // acquire
    out_smem4[0] = __uint_as_float(static_cast<unsigned>(w0));
    out_smem4[1] = __uint_as_float(static_cast<unsigned>(w1));
    out_smem4[2] = __uint_as_float(static_cast<unsigned>(w2));
    out_smem4[3] = __uint_as_float(static_cast<unsigned>(w3));
  }
  __syncthreads();
}
template <typename LastFn, typename PostFn>
__device__ __forceinline__ void grid_barrier_bcast(unsigned* ticket, unsigned long long* words, unsigned seq,
                                                   float* out_smem4, LastFn&& last_fn, PostFn&& post_fn) {
  grid_barrier_bcast(ticket, words, seq, out_smem4, last_fn, post_fn, [] {});
}
#endif

// ---- bulk-copy engine (cp.async.bulk, SASS UBLKCP) + mbarrier -------------------------------------------
// One thread arms an mbarrier with the byte count and starts a contiguous global -> shared copy; the
// consumers wait on the barrier's phase parity. Copies cost no registers and run arbitrarily far ahead.
#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// contiguous global -> shared copy by the bulk-copy engine; completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif

#ifdef __CUDACC__
// Warp sum of N values per lane with N - 1 + log2(32 / N) shuffles instead of 5 N (the shuffle unit, not the issue
// slots, is what 32 warps reducing a dozen sums each run into): at every halving step a lane keeps one half of its
// values and sends the other half to the lane `mask` away; after log2(N) steps it holds ONE value summed over N lanes,
// the remaining bits are a plain butterfly. Every lane returns the total of quantity (lane >> log2(32 / N)) & (N - 1).
// N must be a power of two <= 32; v is clobbered.
template <int N, typename T>
__device__ __forceinline__ T warp_multi_sum(T (&v)[N]) {
  static_assert(N >= 1 && N <= 32 && (N & (N - 1)) == 0, "N must be a power of two");
  const int lane = threadIdx.x & 31;
  int mask = 16;
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1) {
    const bool up = (lane & mask) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const T send = up ? v[j] : v[j + half];
      const T keep = up ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
    mask >>= 1;
  }
  T r = v[0];
#pragma unroll
  for (int m = 16 / N; m >= 1; m >>= 1) r += __shfl_xor_sync(0xffffffffu, r, m);
  return r;
}

#endif

// ---- grid-wide sum of four doubles without a ticket ---------------------------------------------------
// The ticket barrier above costs four dependent L2 round trips after the last CTA is ready (its atomics,
// the fence, the ticket, the totals read, the broadcast). Here every CTA stores its four partial totals
// into its OWN slot as self-validating words {seq | 32 bits of the double} and then every CTA gathers
// all slots itself (<= 3 doubles per thread), spinning only on words whose seq is not there yet: after
// the last CTA's store the critical path is one store, one load and a block reduction. No fence is
// needed because the data travels inside the words that are waited for; consequently this is NOT a
// memory barrier for anything else (use the ticket barrier when one CTA must see another's buffers).
// The summation order is fixed by the grid size, so the totals are bit-reproducible run to run.
// `mine`: thread q < 4 passes this CTA's total q (as returned by block_sum<4>). Totals -> sm_tot[0..3].
#ifdef __CUDACC__
// own_fn(q): this CTA's total q again, from shared memory - the thread that would poll the CTA's own slot
// takes it from there (a load issued right behind another thread's store may still see the old word and
// would cost the last arriver one more round trip).
template <typename OwnFn, typename MidFn>
__device__ __forceinline__ void grid_sum4_bcast(unsigned long long* slots, unsigned seq, double mine, double* sm_tot,
                                                double* sm_scratch, OwnFn&& own_fn, MidFn&& mid_fn) {
  const unsigned long long tag = static_cast<unsigned long long>(seq) << 32;
  if (threadIdx.x < 4) {
    const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(mine));
    unsigned long long* w = slots + (static_cast<size_t>(blockIdx.x) * 4 + threadIdx.x) * 2;
    st_relaxed_u64(w, tag | (b >> 32));
    st_relaxed_u64(w + 1, tag | (b & 0xffffffffull));
  }
  mid_fn();
  // (Measured: looking at the slots BEFORE mid_fn's loads are issued makes the barrier slower, not faster.)
  const int n = static_cast<int>(gridDim.x) * 4;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += kBlock) {   // i & 3 == threadIdx.x & 3: one quantity per thread
    if ((i >> 2) == static_cast<int>(blockIdx.x)) {
      acc += own_fn(i & 3);
      continue;
    }
    const unsigned long long* w = slots + static_cast<size_t>(i) * 2;
    unsigned long long hi, lo;
    do {
      hi = ld_relaxed_u64(w);
      lo = ld_relaxed_u64(w + 1);
    } while ((hi >> 32) != seq || (lo >> 32) != seq);
    acc += __longlong_as_double(static_cast<long long>((hi << 32) | (lo & 0xffffffffull)));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 4; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane < 4) sm_scratch[warp * 4 + lane] = acc;
  __syncthreads();
  if (warp == 0) {
    double x = 0.0;
#pragma unroll
    for (int i = 0; i < (kWarps * 4) / 32; ++i) x += sm_scratch[lane + 32 * i];
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane < 4) sm_tot[lane] = x;
  }
  __syncthreads();
}
#endif

// Variant with an arrival counter (round 2). In grid_sum4_bcast every thread of every waiting CTA polls the slots:
// a CTA that finishes early hammers ~150 L2 lines with 74 warp-loads per round for microseconds, the lines are
// hot-spots for 296 x 16 warps at once, and the last arriver's stores queue behind that traffic (measured: 2.5 to
// 4.5 us between the last publication and the last CTA leaving). Here a CTA publishes its slot, bumps ONE counter
// (red.relaxed, no fence: the slot words validate themselves) and ONE thread per CTA polls that counter while the
// other warps sleep in bar.sync; only then are the slots gathered, normally in a single pass.
#ifdef __CUDACC__
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// sm_own: this CTA's totals, [4][NW] doubles (row q = per-warp partials of quantity q), complete before the call;
// sm_part: [NW][4] doubles of scratch. After the counter fills, ALL threads gather the slots in one go (every load of
// a thread in flight before the first is looked at: one L2 round trip), a shuffle tree and one shared-memory pass
// form the four totals in sm_tot. (Measured alternatives: three dependent rounds per thread = 1.9 us after the
// counter filled; warp 0 alone in five batches of 16 loads = 5 us.)
template <int NW, typename MidFn>
__device__ __forceinline__ void grid_sum4_counted(unsigned long long* slots, unsigned* arrivals, unsigned seq, const double* sm_own,
                                                  double* sm_part, double* sm_tot, MidFn&& mid_fn) {
  const unsigned long long tag = static_cast<unsigned long long>(seq) << 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NT = NW * 32;
  if (threadIdx.x < 32) {
    // this CTA's four totals from the per-warp partials: lane l takes warps l, l + 32, ...; one multi-sum over the warp
    double v4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double x = 0.0;
#pragma unroll
      for (int w = lane; w < NW; w += 32) x += sm_own[q * NW + w];
      v4[q] = x;
    }
    const double mine = warp_multi_sum<4>(v4);          // quantity (lane >> 3) & 3
    if ((lane & 7) == 0) {
      const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(mine));
      unsigned long long* w2 = slots + (static_cast<size_t>(blockIdx.x) * 4 + (lane >> 3)) * 2;
      st_relaxed_u64(w2, tag | (b >> 32));
      st_relaxed_u64(w2 + 1, tag | (b & 0xffffffffull));
    }
    __syncwarp();
    if (lane == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(arrivals) : "memory");
  }
  mid_fn();
  const int n = static_cast<int>(gridDim.x) * 4;
  if (n > NT) {   // more slots than threads (two CTAs per SM): wait for the counter, then gather in one pass
    if (threadIdx.x == 0) {
      while (ld_relaxed_u32(arrivals) < gridDim.x) {}
    }
    __syncthreads();
  }
  // (one CTA per SM: every thread owns at most one slot pair and simply spins on it below - the CTAs arrive within
  // ~1 us of each other, so the polling is short, and leaving the barrier costs one L2 round trip less)
  // slot words of CTA c, quantity q: slots[(c * 4 + q) * 2 + {0, 1}]; thread t handles pairs i = t, t + NT, ...
  // (i & 3 == t & 3: one quantity per thread)
  constexpr int kIt = (kSlotCtas * 4 + NT - 1) / NT;
  unsigned long long hi[kIt], lo[kIt];
#pragma unroll
  for (int u = 0; u < kIt; ++u) {
    const int i = static_cast<int>(threadIdx.x) + NT * u;
    if (i < n) {
      hi[u] = ld_relaxed_u64(slots + static_cast<size_t>(i) * 2);
      lo[u] = ld_relaxed_u64(slots + static_cast<size_t>(i) * 2 + 1);
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < kIt; ++u) {
    const int i = static_cast<int>(threadIdx.x) + NT * u;
    if (i < n) {
      // the counter can run ahead of a slot's words by a few hundred ns: re-read until both carry this launch's tag
      while ((hi[u] >> 32) != seq || (lo[u] >> 32) != seq) {
        hi[u] = ld_relaxed_u64(slots + static_cast<size_t>(i) * 2);
        lo[u] = ld_relaxed_u64(slots + static_cast<size_t>(i) * 2 + 1);
      }
      acc += __longlong_as_double(static_cast<long long>((hi[u] << 32) | (lo[u] & 0xffffffffull)));
    }
  }
#pragma unroll
  for (int o = 16; o >= 4; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane < 4) sm_part[warp * 4 + lane] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double x = 0.0;
#pragma unroll
      for (int w = lane; w < NW; w += 32) x += sm_part[w * 4 + q];
      v4[q] = x;
    }
    const double tot = warp_multi_sum<4>(v4);
    if ((lane & 7) == 0) sm_tot[lane >> 3] = tot;
  }
  // the caller's thread 0 reads sm_tot after a __syncwarp (the writers are in its warp)
}
#endif

// ---- optional per-CTA phase trace (debug / profiling aid) -------------------------------------------
// mde_debug_set_trace(buf) arms it: slot k of CTA b receives the GPU global timer (ns) when the CTA
// passes trace point k. One pointer per translation unit (no relocatable device code in this build).
#ifdef __CUDACC__
static __device__ unsigned long long* g_mde_trace = nullptr;
constexpr int kTraceSlots = 8;
__device__ __forceinline__ void trace_point(int k) {
  unsigned long long* t = g_mde_trace;
  if (t != nullptr && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    t[static_cast<size_t>(blockIdx.x) * kTraceSlots + k] = ns;
    if (k == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      t[static_cast<size_t>(blockIdx.x) * kTraceSlots + 7] = smid;
    }
  }
}
#define MDE_DEFINE_TRACE_SETTER(name)                                                        \
  int name(unsigned long long* buf) {                                                        \
    return cudaMemcpyToSymbol(g_mde_trace, &buf, sizeof(buf)) == cudaSuccess ? MDE_OK : MDE_ECUDA; \
  }
#endif

// ---- typed 4-element loads / stores ----------------------------------------------------------
// LDG.E.128 for fp32, LDG.E.64 for half/bf16 (4 elements either way). `Keep` = true leaves the
// line in L2 for the gradient phase; false streams it (evict-first).
template <typename T>
struct Elem;

template <>
struct Elem<float> {
  template <bool Keep>
  static __device__ __forceinline__ float4 ld4(const float* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    return Keep ? __ldg(q) : __ldcs(q);
  }
  static __device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void st4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
  static __device__ __forceinline__ void st1(float* p, float v) { __stcs(p, v); }
};

template <>
struct Elem<__half> {
  template <bool Keep>
  static __device__ __forceinline__ float4 ld4(const __half* p) {
    const uint2* q = reinterpret_cast<const uint2*>(p);
    uint2 r = Keep ? __ldg(q) : __ldcs(q);
    __half2 a = *reinterpret_cast<__half2*>(&r.x), b = *reinterpret_cast<__half2*>(&r.y);
    float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ float ld1(const __half* p) {
    return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(p))));
  }
  static __device__ __forceinline__ void st4(__half* p, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<unsigned*>(&a);
    r.y = *reinterpret_cast<unsigned*>(&b);
    __stcs(reinterpret_cast<uint2*>(p), r);
  }
  static __device__ __forceinline__ void st1(__half* p, float v) { *p = __float2half_rn(v); }
};

template <>
struct Elem<__nv_bfloat16> {
  template <bool Keep>
  static __device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    const uint2* q = reinterpret_cast<const uint2*>(p);
    uint2 r = Keep ? __ldg(q) : __ldcs(q);
    // bf16 -> fp32 is a 16-bit shift
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u),
                       __uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
  }
  static __device__ __forceinline__ float ld1(const __nv_bfloat16* p) {
    unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(p));
    return __uint_as_float(static_cast<unsigned>(u) << 16);
  }
  static __device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<unsigned*>(&a);
    r.y = *reinterpret_cast<unsigned*>(&b);
    __stcs(reinterpret_cast<uint2*>(p), r);
  }
  static __device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// ---- reductions --------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum N per-thread doubles over the CTA; the totals are returned to threads 0..N-1 (thread q
// holds total q). `sm` is N*kWarps doubles of shared memory. Ends with the data consumed, so
// the buffer can be reused after the next __syncthreads().
template <int N>
__device__ __forceinline__ double block_sum(const double (&v)[N], double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < N; ++q) {
    double s = warp_sum(v[q]);
    if (lane == 0) sm[q * kWarps + warp] = s;
  }
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x < N) {
#pragma unroll
    for (int w = 0; w < kWarps; ++w) tot += sm[threadIdx.x * kWarps + w];
  }
  __syncthreads();
  return tot;
}

// order-preserving map float -> unsigned (for atomicMax on floats of either sign); 0 is below
// every real float, so a zero-filled workspace is the identity.
__device__ __forceinline__ unsigned float_key(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
  unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

// Balanced contiguous partition of `total` units over the CTAs of a grid, every chunk boundary a
// multiple of `gran` units (128-byte alignment of every chunk start). The quotient / remainder are
// computed once on the host so that the kernels do no 64-bit division.
struct Chunking {
  int64_t total;     // units
  int64_t groups;    // ceil(total / gran)
  int64_t base;      // groups / n_cta
  int rem;           // groups % n_cta
  int gran;
};
inline Chunking make_chunking(int64_t total, int gran, int n_cta) {
  Chunking c;
  c.total = total;
  c.gran = gran;
  c.groups = (total + gran - 1) / gran;
  c.base = c.groups / n_cta;
  c.rem = static_cast<int>(c.groups % n_cta);
  return c;
}
__device__ __forceinline__ void cta_chunk(const Chunking& c, int cta, int64_t& begin, int64_t& end) {
  const int64_t g0 = c.base * cta + (cta < c.rem ? cta : c.rem);
  const int64_t g1 = g0 + c.base + (cta < c.rem ? 1 : 0);
  begin = g0 * c.gran;
  end = g1 * c.gran;
  if (end > c.total) end = c.total;
  if (begin > c.total) begin = c.total;
}

}  // namespace mde
