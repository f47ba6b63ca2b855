// pointcloud.cu - depth map -> camera-space point cloud (reference depth2pointcloud.py:12-31,
// point_cloud) with the camera->world transform of depth2pointcloud.py:103-108 fused.
//
//   factor = 2 tan(angle_x/2); valid = clip_start < d < clip_end; z = -d (NaN when invalid)
//   x = -(factor*z) * (c - cols/2) / max(rows, cols),  y = (factor*z) * (r - rows/2) / max(rows, cols)
//   invalid -> (0, 0, NaN)
// numpy promotion in the reference: factor*z is fp32 (python scalar times fp32 array), the pixel
// offsets are float64, so x and y are float64 and z is fp32 widened by dstack. The kernel follows
// that (fp32 product, fp64 scale) and writes fp64 [.,3] (reference layout) or fp32 [.,3].
// Streaming: 4 B/px read, 12 (fp32) or 24 (fp64) B/px written. Two kernels: a 128-bit tiled one with
// shared-memory staged, fully coalesced stores (w % 4 == 0) and a scalar one for any shape.
#include "common.cuh"

namespace mde {
namespace {

struct Mat4 {
  float m[16];
};

template <typename OT>
__global__ void __launch_bounds__(256) point_cloud_kernel(const float* __restrict__ depth, int64_t n_img, int h, int w,
                                                          float factor, float clip_start, float clip_end, int has_mat,
                                                          Mat4 M, OT* __restrict__ out) {
  const int64_t hw = static_cast<int64_t>(h) * w;
  const int64_t total = n_img * hw;
  const double half_c = static_cast<double>(w) / 2.0, half_r = static_cast<double>(h) / 2.0;
  const double ratio = static_cast<double>(h > w ? h : w);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * 256) {
    const int64_t rem = i % hw;
    const int r = static_cast<int>(rem / w), c = static_cast<int>(rem - static_cast<int64_t>(r) * w);
    const float d = __ldcs(depth + i);
    const bool valid = (d > clip_start) && (d < clip_end);
    double x = 0.0, y = 0.0, z = __longlong_as_double(0xfff8000000000000LL);   // -np.where(valid, depth, nan): the NaN carries the minus sign
    if (valid) {
      const float zf = -d;
      const float fz = factor * zf;  // fp32 product, as numpy evaluates `factor * z`
      x = -(static_cast<double>(fz) * (static_cast<double>(c) - half_c) / ratio);
      y = static_cast<double>(fz) * (static_cast<double>(r) - half_r) / ratio;
      z = static_cast<double>(zf);
    } else {
      x = -0.0;  // -np.where(valid, ., 0) gives -0.0
    }
    if (has_mat) {
      // mathutils holds fp32: Vector(p) rounds the point, Matrix @ Vector extends it with w = 1
      const double px = static_cast<double>(static_cast<float>(x)), py = static_cast<double>(static_cast<float>(y)),
                   pz = static_cast<double>(static_cast<float>(z));
      const double wx = M.m[0] * px + M.m[1] * py + M.m[2] * pz + M.m[3];
      const double wy = M.m[4] * px + M.m[5] * py + M.m[6] * pz + M.m[7];
      const double wz = M.m[8] * px + M.m[9] * py + M.m[10] * pz + M.m[11];
      x = wx; y = wy; z = wz;
    }
    out[3 * i + 0] = static_cast<OT>(x);
    out[3 * i + 1] = static_cast<OT>(y);
    out[3 * i + 2] = static_cast<OT>(z);
  }
}

// a / b in fp64, correctly rounded for the normal-range operands of this kernel, from a precomputed
// reciprocal: q = a * (1/b) followed by one residual correction (the tail of the IEEE divide routine).
// B200 issues fp64 at half the fp32 rate; a full divide per coordinate would make the kernel fp64-bound.
__device__ __forceinline__ double div_by(double a, double b, double rb) {
  const double q = a * rb;
  const double r = fma(-q, b, a);
  return (r == 0.0) ? q : fma(r, rb, q);   // exact quotients keep their sign of zero (x = -0.0 vs +0.0 on the centre column)
}

// 128-bit path (w % 4 == 0, 16-byte aligned buffers): a CTA turns 1024 consecutive pixels into 3072
// coordinates per tile. Each thread loads one quad of depths (one image row: a quad never straddles a
// row), evaluates its 12 coordinates and parks them in shared memory in output order; the tile is then
// written with fully coalesced 128-bit stores (a thread-per-point [.,3] store would touch every sector
// three times). Persistent grid over the tiles.
constexpr int kPcBlock = 256;
constexpr int kPcTilePx = 4 * kPcBlock;

template <typename OT>
__global__ void __launch_bounds__(kPcBlock) point_cloud_tile_kernel(const float* __restrict__ depth, int64_t n_img, int h,
                                                                   int w, float factor, float clip_start, float clip_end,
                                                                   int has_mat, Mat4 M, OT* __restrict__ out) {
  __shared__ __align__(16) OT so[3 * kPcTilePx];
  const int64_t hw = static_cast<int64_t>(h) * w;
  const int64_t total = n_img * hw;
  const int64_t n_tiles = (total + kPcTilePx - 1) / kPcTilePx;
  const double half_c = static_cast<double>(w) / 2.0, half_r = static_cast<double>(h) / 2.0;
  const double ratio = static_cast<double>(h > w ? h : w), inv_ratio = 1.0 / ratio;
  const double qnan = __longlong_as_double(0xfff8000000000000LL);   // -np.where(valid, depth, nan): the NaN carries the minus sign
  constexpr int kVec = 16 / sizeof(OT);                 // output elements per 128-bit store
  constexpr int kStores = 3 * kPcTilePx / kVec / kPcBlock;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t px0 = tile * kPcTilePx + 4 * threadIdx.x;
    if (px0 < total) {
      const float4 d4 = __ldcs(reinterpret_cast<const float4*>(depth + px0));
      const int64_t rem = px0 % hw;
      const int r = static_cast<int>(rem / w), c0 = static_cast<int>(rem - static_cast<int64_t>(r) * w);
      const double yr = static_cast<double>(r) - half_r;
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
      OT* dst = so + 12 * threadIdx.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d = dv[k];
        const bool valid = (d > clip_start) && (d < clip_end);
        double x = -0.0, y = 0.0, z = qnan;   // -np.where(valid, ., 0) gives -0.0
        if (valid) {
          const float zf = -d;
          const double fz = static_cast<double>(factor * zf);  // fp32 product, as numpy evaluates `factor * z`
          x = -div_by(fz * (static_cast<double>(c0 + k) - half_c), ratio, inv_ratio);
          y = div_by(fz * yr, ratio, inv_ratio);
          z = static_cast<double>(zf);
        }
        if (has_mat) {
          // mathutils holds fp32: Vector(p) rounds the point, Matrix @ Vector extends it with w = 1
          const double px = static_cast<double>(static_cast<float>(x)), py = static_cast<double>(static_cast<float>(y)),
                       pz = static_cast<double>(static_cast<float>(z));
          const double wx = M.m[0] * px + M.m[1] * py + M.m[2] * pz + M.m[3];
          const double wy = M.m[4] * px + M.m[5] * py + M.m[6] * pz + M.m[7];
          const double wz = M.m[8] * px + M.m[9] * py + M.m[10] * pz + M.m[11];
          x = wx; y = wy; z = wz;
        }
        dst[3 * k + 0] = static_cast<OT>(x);
        dst[3 * k + 1] = static_cast<OT>(y);
        dst[3 * k + 2] = static_cast<OT>(z);
      }
    }
    __syncthreads();
    const int64_t e0 = tile * (3 * kPcTilePx);              // first output element of the tile
    const int64_t e_end = 3 * total;
#pragma unroll
    for (int j = 0; j < kStores; ++j) {
      const int v = threadIdx.x + j * kPcBlock;             // 128-bit unit within the tile
      if (e0 + static_cast<int64_t>(v) * kVec < e_end)      // total % 4 == 0: units are all-or-nothing
        __stcs(reinterpret_cast<float4*>(out + e0) + v, reinterpret_cast<const float4*>(so)[v]);
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace mde

extern "C" int mde_point_cloud(const float* depth, int64_t n_img, int64_t h, int64_t w, float angle_x, float clip_start,
                               float clip_end, const float* matrix_world_host, int out_f64, void* out, void* stream) {
  using namespace mde;
  MDE_REQUIRE(depth && out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && h > 0 && w > 0 && h < (1 << 30) && w < (1 << 30), MDE_EINVAL, "bad shape");
  Mat4 M{};
  if (matrix_world_host)
    for (int i = 0; i < 16; ++i) M.m[i] = matrix_world_host[i];
  const float factor = static_cast<float>(2.0 * tan(static_cast<double>(angle_x) / 2.0));
  const int64_t total = n_img * h * w;
  int64_t grid = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (grid > cap) grid = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (w % 4 == 0 && aligned_to(depth, 16) && aligned_to(out, 16)) {
    int64_t tiles = (total + kPcTilePx - 1) / kPcTilePx;
    const int64_t tcap = static_cast<int64_t>(sm_count()) * (out_f64 ? 6 : 8);
    if (tiles > tcap) tiles = tcap;
    if (out_f64)
      point_cloud_tile_kernel<double><<<static_cast<unsigned>(tiles), kPcBlock, 0, st>>>(
          depth, n_img, static_cast<int>(h), static_cast<int>(w), factor, clip_start, clip_end,
          matrix_world_host != nullptr, M, static_cast<double*>(out));
    else
      point_cloud_tile_kernel<float><<<static_cast<unsigned>(tiles), kPcBlock, 0, st>>>(
          depth, n_img, static_cast<int>(h), static_cast<int>(w), factor, clip_start, clip_end,
          matrix_world_host != nullptr, M, static_cast<float*>(out));
    count_launch();
    MDE_CUDA_TRY(cudaGetLastError());
    return MDE_OK;
  }
  if (out_f64)
    point_cloud_kernel<double><<<static_cast<unsigned>(grid), 256, 0, st>>>(
        depth, n_img, static_cast<int>(h), static_cast<int>(w), factor, clip_start, clip_end,
        matrix_world_host != nullptr, M, static_cast<double*>(out));
  else
    point_cloud_kernel<float><<<static_cast<unsigned>(grid), 256, 0, st>>>(
        depth, n_img, static_cast<int>(h), static_cast<int>(w), factor, clip_start, clip_end,
        matrix_world_host != nullptr, M, static_cast<float*>(out));
  count_launch();
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
