// explicit instantiation: Symphony kernels, distribution kind 0
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony<rb::kDistPowerLaw>(const BatchArgs &, bool, int, cudaStream_t);
template int stage_symphony_diag<rb::kDistPowerLaw>(const BatchArgs &, const DiagArgs &, int, cudaStream_t);
}
