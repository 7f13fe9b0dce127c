// explicit instantiation: product-path (fast) Heyvaerts kernel, distribution kind 3
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_heyvaerts_fast<rb::kDistPitchyKappa>(const BatchArgs &, int, cudaStream_t);
}
