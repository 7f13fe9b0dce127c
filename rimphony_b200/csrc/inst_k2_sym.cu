// explicit instantiation: Symphony kernels, distribution kind 2
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony<rb::kDistPitchyPL>(const BatchArgs &, bool, int, cudaStream_t);
template int stage_symphony_diag<rb::kDistPitchyPL>(const BatchArgs &, const DiagArgs &, int, cudaStream_t);
}
