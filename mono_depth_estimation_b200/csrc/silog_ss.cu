// silog_ss.cu - instantiations and the launcher of the shared-memory-stash SILog kernel (silog_ss.cuh):
// silog_loss forward+backward (reference criteria.py:724-732), optionally with the pooled metric suite of
// MetricComputation.compute (reference metrics.py:58-67) fused into the same pass.
#include "silog_ss.cuh"

namespace mde {
MDE_DEFINE_TRACE_SETTER(set_trace_silog_ss)

int launch_silog_ss(LossArgs& a, unsigned mg, cudaStream_t st, bool& taken) {
  switch (mg) {
    case 0u: return launch_loss_ss<0u>(a, st, taken);
    case (kGrpLog | kGrpRel): return launch_loss_ss<(kGrpLog | kGrpRel)>(a, st, taken);
    case kGrpAll: return launch_loss_ss<kGrpAll>(a, st, taken);
    default: taken = false; return MDE_OK;
  }
}
}  // namespace mde
