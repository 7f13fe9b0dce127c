// explicit instantiation: product-path (fast) Symphony kernel, distribution kind 3
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony_fast<rb::kDistPitchyKappa>(const BatchArgs &, int, cudaStream_t);
}
