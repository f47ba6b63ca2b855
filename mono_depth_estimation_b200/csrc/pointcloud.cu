// pointcloud.cu - depth map -> camera-space point cloud (reference depth2pointcloud.py:12-31,
// point_cloud) with the camera->world transform of depth2pointcloud.py:103-108 fused.
//
//   factor = 2 tan(angle_x/2); valid = clip_start < d < clip_end; z = -d (NaN when invalid)
//   x = -(factor*z) * (c - cols/2) / max(rows, cols),  y = (factor*z) * (r - rows/2) / max(rows, cols)
//   invalid -> (0, 0, NaN)
// numpy promotion in the reference: factor*z is fp32 (python scalar times fp32 array), the pixel
// offsets are float64, so x and y are float64 and z is fp32 widened by dstack. The kernel follows
// that (fp32 product, fp64 scale) and writes fp64 [.,3] (reference layout) or fp32 [.,3].
// Streaming: 4 B/px read, 12 (fp32) or 24 (fp64) B/px written.
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
    double x = 0.0, y = 0.0, z = __longlong_as_double(0x7ff8000000000000LL);
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
