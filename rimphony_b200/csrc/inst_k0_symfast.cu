// explicit instantiation: product-path (fast) Symphony kernel, distribution kind 0
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony_fast<rb::kDistPowerLaw>(const BatchArgs &, int, cudaStream_t);
}
