// losses.cu - masked scalar depth losses, forward and backward fused in ONE cooperative launch.
//
//   MDE_LOSS_L1           MaskedL1Loss      reference criteria.py:80-90
//   MDE_LOSS_MSE          MaskedMSELoss     reference criteria.py:67-77
//   MDE_LOSS_BERHU        berHuLoss         reference criteria.py:111-133
//   MDE_LOSS_LAINA_BERHU  LainaBerHuLoss    reference criteria.py:476-506
//   MDE_LOSS_SILOG        silog_loss        reference criteria.py:724-732
//   (MDE_LOSS_EIGEN lives in eigen.cu; the variants with the metric suite fused in are
//    instantiated in losses_fused.cu from the same kernel template, losses_kernel.cuh)
//
// Kernel shape: persistent grid (<= 148 SMs x 2 CTAs), each CTA owns one contiguous 128-byte
// aligned chunk. Phase A0 (berHu / Laina only) finds the global max, phase A1 reduces the masked
// sums and counts (fp32 per tile, fp64 across tiles, warp shuffle + shared memory per CTA, one
// fp64 atomic per CTA and quantity), a grid-wide barrier publishes the totals, and phase B
// writes dloss/dpred walking the SAME chunk backwards so that the lines touched last are re-read
// first (they are still in L2). HBM traffic is 8 B/px read + 4 B/px written when pred+target
// fit in the 126 MB L2.
//
// The per-pixel code is branch-free (invalid pixels are replaced by p = t = 1 and contribute exact
// zeros) and needs ONE logarithm per pixel and phase (log_ratio). SILog with an fp32 gradient
// buffer additionally STASHES d_i = ln p_i - ln t_i in that buffer during phase A, so phase B reads
// d_i and p_i only (no second logarithm, no second read of the target).
#include "losses_kernel.cuh"

namespace mde {
MDE_DEFINE_TRACE_SETTER(set_trace_losses)
}  // namespace mde

namespace mde {

int eigen_loss_launch(const void* pred, int pred_dtype, const float* target, int64_t n_img, int64_t h, int64_t w,
                      float grad_scale, void* ws, float* loss_out, double* totals_out, void* grad,
                      cudaStream_t st);  // eigen.cu

}  // namespace mde

extern "C" int mde_masked_loss(int kind, const void* pred, int pred_dtype, const float* target,
                               const uint8_t* mask_u8, int64_t n_img, int64_t h, int64_t w,
                               const mde_loss_params* params, float grad_scale, void* ws, float* loss_out,
                               double* totals_out, void* grad, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && ws && loss_out, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && h > 0 && w > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(aligned_to(target, 4), MDE_EALIGN, "misaligned target");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kind == MDE_LOSS_EIGEN)
    return eigen_loss_launch(pred, pred_dtype, target, n_img, h, w, grad_scale, ws, loss_out, totals_out, grad, st);
  LossArgs a = make_loss_args(pred, target, mask_u8, n_img, h, w, params, grad_scale, ws, loss_out, totals_out, grad);
  return launch_loss_kind<0u>(kind, a, pred_dtype, st);
}
