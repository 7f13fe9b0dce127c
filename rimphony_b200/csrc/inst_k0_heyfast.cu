// explicit instantiation: product-path (fast) Heyvaerts kernel, distribution kind 0
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_heyvaerts_fast<rb::kDistPowerLaw>(const BatchArgs &, int, cudaStream_t);
}
