// explicit instantiation: product-path (fast) Heyvaerts kernel, distribution kind 2
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_heyvaerts_fast<rb::kDistPitchyPL>(const BatchArgs &, int, cudaStream_t);
}
