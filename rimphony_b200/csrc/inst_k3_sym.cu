// explicit instantiation: Symphony kernels, distribution kind 3
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony<rb::kDistPitchyKappa>(const BatchArgs &, bool, int, cudaStream_t);
template int stage_symphony_diag<rb::kDistPitchyKappa>(const BatchArgs &, const DiagArgs &, int, cudaStream_t);
}
