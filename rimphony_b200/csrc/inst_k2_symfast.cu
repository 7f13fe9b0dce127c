// explicit instantiation: product-path (fast) Symphony kernel, distribution kind 2
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony_fast<rb::kDistPitchyPL>(const BatchArgs &, int, cudaStream_t);
}
