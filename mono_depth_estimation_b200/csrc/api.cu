// api.cu - host-side plumbing of libmde_b200: error text, launch accounting, device queries,
// workspace sizing / initialisation and the small utility kernels.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <map>
#include <mutex>

#include "common.cuh"

namespace mde {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

namespace {
std::mutex g_mu;
std::map<int, int> g_sm_count;                            // device -> SM count
std::map<std::pair<int, const void*>, int> g_coop_grid;   // (device, kernel) -> co-resident CTAs
}  // namespace

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_sm_count.find(dev);
  if (it != g_sm_count.end()) return it->second;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  g_sm_count[dev] = n;
  return n;
}

int coop_grid(const void* func, int block, size_t smem) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("cudaGetDevice failed");
    return -1;
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_coop_grid.find({dev, func});
    if (it != g_coop_grid.end()) return it->second;
  }
  int per_sm = 0;
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024) e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, block, smem);
  if (e != cudaSuccess || per_sm <= 0) {
    set_error("occupancy query failed: %s", cudaGetErrorString(e));
    return -1;
  }
  if (per_sm > kCtasPerSm) per_sm = kCtasPerSm;
  const int g = per_sm * sm_count();
  std::lock_guard<std::mutex> lk(g_mu);
  g_coop_grid[{dev, func}] = g;
  return g;
}

cudaError_t launch_pdl(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t st, bool cooperative) {
  static const bool enabled = [] { const char* e = getenv("MDE_PDL"); return !(e && atoi(e) == 0); }();
  static std::atomic<bool> refused{false};
  if (enabled && !refused.load(std::memory_order_relaxed)) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
    if (cooperative) {
      at[n].id = cudaLaunchAttributeCooperative;
      at[n].val.cooperative = 1;
      ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = static_cast<unsigned>(n);
    const cudaError_t e = cudaLaunchKernelExC(&cfg, func, args);
    if (e == cudaSuccess) return e;
    (void)cudaGetLastError();
    refused.store(true, std::memory_order_relaxed);
  }
  if (cooperative) return cudaLaunchCooperativeKernel(func, grid, block, args, smem, st);
  return cudaLaunchKernel(func, grid, block, args, smem, st);
}

namespace {

__global__ void ws_init_kernel(void* ws_raw, unsigned max_images) {
  Ws ws = ws_view(ws_raw);
  ws.hdr->max_images = max_images;
}

// x *= *s. One CTA of 256 threads per SM: with autograd's default grad_output (1.0) the kernel is pure launch cost, and
// a small grid retires sooner; otherwise every thread keeps four 128-bit (64-bit for half types) accesses in flight.
constexpr int kScaleBlock = 256;
template <typename T>
__global__ void __launch_bounds__(kScaleBlock) scale_kernel(T* __restrict__ x, int64_t n, const float* __restrict__ s) {
  pdl_trigger();   // lets a dependent launch that asked for it start early
  const float k = __ldg(s);
  if (k == 1.0f) return;  // autograd's default grad_output: the stashed gradient is already final
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * kScaleBlock;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kScaleBlock + threadIdx.x;
  int64_t done = 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) == 0u) {
    const int64_t nq = n >> 2;
    for (int64_t q0 = tid; q0 < nq; q0 += 4 * nthr) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t q = q0 + u * nthr;
        if (q < nq) v[u] = Elem<T>::template ld4<false>(x + 4 * q);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t q = q0 + u * nthr;
        if (q < nq) Elem<T>::st4(x + 4 * q, make_float4(v[u].x * k, v[u].y * k, v[u].z * k, v[u].w * k));
      }
    }
    done = nq << 2;
  }
  for (int64_t i = done + tid; i < n; i += nthr) x[i] = static_cast<T>(static_cast<float>(x[i]) * k);
}

}  // namespace
}  // namespace mde

extern "C" size_t mde_workspace_bytes(int64_t max_images) {
  if (max_images < 1) max_images = 1;
  return mde::kWsFixedBytes + static_cast<size_t>(3) * static_cast<size_t>(max_images) * mde::kIacc * sizeof(double);
}

extern "C" int mde_workspace_init(void* ws, int64_t max_images, void* stream) {
  using namespace mde;
  MDE_REQUIRE(ws != nullptr, MDE_EINVAL, "null workspace");
  MDE_REQUIRE(max_images >= 1 && max_images < (int64_t(1) << 31), MDE_EINVAL, "bad max_images");
  MDE_REQUIRE(aligned_to(ws, 16), MDE_EALIGN, "workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MDE_CUDA_TRY(cudaMemsetAsync(ws, 0, mde_workspace_bytes(max_images), st));
  ws_init_kernel<<<1, 1, 0, st>>>(ws, static_cast<unsigned>(max_images));
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

extern "C" int mde_scale_inplace(void* x, int dtype, int64_t n, const float* scale_dev, void* stream) {
  using namespace mde;
  MDE_REQUIRE(x && scale_dev, MDE_EINVAL, "null pointer");
  if (n <= 0) return MDE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t grid = (n + 4 * kScaleBlock - 1) / (4 * kScaleBlock);
  const int64_t cap = static_cast<int64_t>(sm_count());
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  switch (dtype) {
    case MDE_F32: scale_kernel<float><<<static_cast<unsigned>(grid), kScaleBlock, 0, st>>>(static_cast<float*>(x), n, scale_dev); break;
    case MDE_F16: scale_kernel<__half><<<static_cast<unsigned>(grid), kScaleBlock, 0, st>>>(static_cast<__half*>(x), n, scale_dev); break;
    case MDE_BF16: scale_kernel<__nv_bfloat16><<<static_cast<unsigned>(grid), kScaleBlock, 0, st>>>(static_cast<__nv_bfloat16*>(x), n, scale_dev); break;
    default: set_error("mde_scale_inplace: unknown dtype %d", dtype); return MDE_EINVAL;
  }
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

namespace mde {
int set_trace_losses(unsigned long long* buf);
int set_trace_losses_fused(unsigned long long* buf);
int set_trace_silog_ss(unsigned long long* buf);
int set_trace_vnl(unsigned long long* buf);
}  // namespace mde

extern "C" int mde_debug_set_trace(void* device_buf) {
  unsigned long long* b = static_cast<unsigned long long*>(device_buf);
  const int r1 = mde::set_trace_losses(b), r2 = mde::set_trace_losses_fused(b), r3 = mde::set_trace_silog_ss(b);
  const int r4 = mde::set_trace_vnl(b);
  return (r1 == MDE_OK && r2 == MDE_OK && r3 == MDE_OK && r4 == MDE_OK) ? MDE_OK : MDE_ECUDA;
}

extern "C" const char* mde_last_error(void) { return mde::g_err; }

extern "C" const char* mde_version(void) { return "mde_b200 0.1.0 (sm_100a)"; }

extern "C" uint64_t mde_launch_count(void) { return mde::g_launches.load(std::memory_order_relaxed); }

extern "C" int mde_device_info(int* sm_count_out, int* coop_ctas) {
  const int n = mde::sm_count();
  if (sm_count_out) *sm_count_out = n;
  if (coop_ctas) *coop_ctas = n * mde::kCtasPerSm;
  return MDE_OK;
}
