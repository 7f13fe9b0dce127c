// explicit instantiation: product-path (fast) Symphony kernel, distribution kind 1
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony_fast<rb::kDistThermalJuettner>(const BatchArgs &, int, cudaStream_t);
}
