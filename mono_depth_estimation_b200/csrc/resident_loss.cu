// resident_loss.cu - instantiations and the launcher of the register-resident small-input loss kernel
// (resident_loss.cuh): MaskedL1Loss / MaskedMSELoss / berHuLoss / LainaBerHuLoss forward+backward (reference
// criteria.py:67-133, :476-506), optionally with the pooled metric suite of MetricComputation.compute (reference
// metrics.py:58-67) fused into the same pass.
#include "resident_loss.cuh"

namespace mde {

int launch_resident(LossArgs& a, int kind, unsigned mg, cudaStream_t st, bool& taken) {
  switch (mg) {
    case 0u: return launch_resident_kind<0u>(a, kind, st, taken);
    case (kGrpLog | kGrpRel): return launch_resident_kind<(kGrpLog | kGrpRel)>(a, kind, st, taken);
    case kGrpAll: return launch_resident_kind<kGrpAll>(a, kind, st, taken);
    default: taken = false; return MDE_OK;
  }
}
}  // namespace mde
