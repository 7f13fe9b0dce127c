// explicit instantiation: Symphony kernels, distribution kind 3
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony<rb::kDistPitchyKappa>(const BatchArgs &, bool, int, cudaStream_t);
}
