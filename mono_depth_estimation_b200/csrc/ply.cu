// ply.cu - the on-disk format next to the back-projection kernel: the ASCII PLY file that
// reference depth2pointcloud.py:131-154 assembles with a Python loop over every pixel. Host code only.
//
//   for v in pixels:  front point, then back point, each skipped when its x is NaN (:134-138)
//       "%f %f %f %d %d %d 0\n" % (x, y, z, color[v,2], color[v,1], color[v,0])        # BGR (cv2) -> RGB
//   header with `element vertex <count>` (:141-152), the lines, and one more "\n" (the template ends "%s\n")
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

// "%f" of a finite double, appended to buf; returns the new length. snprintf is the specification here:
// Python's "%f" and C's give the same correctly rounded 6-decimal text.
inline size_t put_f(char* buf, size_t n, double v) { return n + static_cast<size_t>(snprintf(buf + n, 400, "%f", v)); }
inline size_t put_u8(char* buf, size_t n, unsigned v) {
  if (v >= 100) buf[n++] = static_cast<char>('0' + v / 100);
  if (v >= 10) buf[n++] = static_cast<char>('0' + (v / 10) % 10);
  buf[n++] = static_cast<char>('0' + v % 10);
  return n;
}

}  // namespace

extern "C" int64_t mde_write_ply(const char* path, const double* front_xyz, const double* back_xyz, const uint8_t* color_bgr,
                                 int64_t n_points) {
  using namespace mde;
  if (!path || !front_xyz || !color_bgr || n_points < 0) {
    set_error("mde_write_ply: null pointer or negative count");
    return MDE_EINVAL;
  }
  int64_t count = 0;
  for (int64_t v = 0; v < n_points; ++v) {
    if (!std::isnan(front_xyz[3 * v])) ++count;
    if (back_xyz && !std::isnan(back_xyz[3 * v])) ++count;
  }
  FILE* f = fopen(path, "wb");
  if (!f) {
    set_error("mde_write_ply: cannot open %s", path);
    return MDE_EINVAL;
  }
  fprintf(f,
          "ply\nformat ascii 1.0\nelement vertex %lld\nproperty float x\nproperty float y\nproperty float z\n"
          "property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\nend_header\n",
          static_cast<long long>(count));
  std::vector<char> buf(1 << 20);
  size_t n = 0;
  auto emit = [&](const double* p, const uint8_t* c) {
    if (n + 1400 > buf.size()) {
      fwrite(buf.data(), 1, n, f);
      n = 0;
    }
    n = put_f(buf.data(), n, p[0]); buf[n++] = ' ';
    n = put_f(buf.data(), n, p[1]); buf[n++] = ' ';
    n = put_f(buf.data(), n, p[2]); buf[n++] = ' ';
    n = put_u8(buf.data(), n, c[2]); buf[n++] = ' ';
    n = put_u8(buf.data(), n, c[1]); buf[n++] = ' ';
    n = put_u8(buf.data(), n, c[0]);
    memcpy(buf.data() + n, " 0\n", 3);
    n += 3;
  };
  for (int64_t v = 0; v < n_points; ++v) {
    if (!std::isnan(front_xyz[3 * v])) emit(front_xyz + 3 * v, color_bgr + 3 * v);
    if (back_xyz && !std::isnan(back_xyz[3 * v])) emit(back_xyz + 3 * v, color_bgr + 3 * v);
  }
  buf[n++] = '\n';
  fwrite(buf.data(), 1, n, f);
  const bool ok = (ferror(f) == 0);
  fclose(f);
  if (!ok) {
    set_error("mde_write_ply: write error on %s", path);
    return MDE_EINVAL;
  }
  return count;
}
