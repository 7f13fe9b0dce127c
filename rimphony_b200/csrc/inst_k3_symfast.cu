// explicit instantiation: product-path (fast) Symphony kernel, distribution kind 3
#define RB_LEAN_MATH 1 // lean division, exp and log (rb_core.cuh)
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony_fast<rb::kDistPitchyKappa>(const BatchArgs &, int, cudaStream_t);
}
