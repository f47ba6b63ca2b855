// losses_fused.cu - masked loss forward+backward WITH the pooled metric suite fused into the reduce
// phase: one read of pred/target serves the loss sums, the loss gradient and every metric of
// MetricComputation.compute (reference metrics.py:58-67) - 12 B/px for the whole training-step tail
// instead of 12 + 8. For SILog the metric suite's ln(max/min) doubles as the loss residual, so the
// fused step costs one IEEE divide and one logarithm per pixel.
#include "losses_kernel.cuh"

namespace mde {
MDE_DEFINE_TRACE_SETTER(set_trace_losses_fused)
}  // namespace mde

extern "C" int mde_masked_loss_metrics(int kind, const void* pred, int pred_dtype, const float* target,
                                       const uint8_t* mask_u8, int64_t n_img, int64_t h, int64_t w,
                                       const mde_loss_params* params, float grad_scale, unsigned metric_flags,
                                       void* ws, float* loss_out, double* totals_out, void* grad,
                                       double* metrics_f64, float* metrics_f32, void* stream) {
  using namespace mde;
  MDE_REQUIRE(pred && target && ws && loss_out && metrics_f64, MDE_EINVAL, "null pointer");
  MDE_REQUIRE(n_img > 0 && h > 0 && w > 0, MDE_EINVAL, "empty input");
  MDE_REQUIRE(kind >= 0 && kind < MDE_LOSS_COUNT && kind != MDE_LOSS_EIGEN, MDE_EINVAL,
              "fused metrics are available for L1, MSE, BERHU, LAINA_BERHU and SILOG");
  MDE_REQUIRE(aligned_to(target, 4) && aligned_to(metrics_f64, 8), MDE_EALIGN, "misaligned pointer");
  MDE_REQUIRE((metric_flags & MDE_METRICS_REFERENCE_MATH) == 0, MDE_EINVAL,
              "the fused path uses the fast metric forms; call mde_metrics for reference math");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LossArgs a = make_loss_args(pred, target, mask_u8, n_img, h, w, params, grad_scale, ws, loss_out, totals_out, grad);
  a.met_f64 = metrics_f64;
  a.met_f32 = metrics_f32;
  a.met_accum = params ? params->metrics_accum : nullptr;
  a.met_raw_accum = params ? params->metrics_raw_accum : nullptr;
  unsigned g = (metric_flags >> 8) & kGrpMask;
  // two instantiations: {log, rel} (the reference's default metric list) and everything; the SS SILog kernel has a
  // third, {log, rsq}, for lists that need only the 'rmse' sum of the REL group
  a.rsq_only = (g & kGrpRsq) != 0 && (g & kGrpRel) == 0;
  if (g != 0 && (g & kGrpLog1p) == 0) return launch_loss_kind<(kGrpLog | kGrpRel)>(kind, a, pred_dtype, st);
  return launch_loss_kind<kGrpAll>(kind, a, pred_dtype, st);
}
