// explicit instantiation: Symphony kernels, distribution kind 3
#define RB_LEAN_MATH 1 // lean division, exp, log and square root (rb_core.cuh): same rule sequence, -13 % time, GPU parity tests unchanged
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony<rb::kDistPitchyKappa>(const BatchArgs &, bool, int, cudaStream_t);
template int stage_symphony_diag<rb::kDistPitchyKappa>(const BatchArgs &, const DiagArgs &, int, cudaStream_t);
}
