// explicit instantiation: product-path (fast) Heyvaerts kernel, distribution kind 1
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_heyvaerts_fast<rb::kDistThermalJuettner>(const BatchArgs &, int, cudaStream_t);
}
