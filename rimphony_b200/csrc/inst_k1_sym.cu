// explicit instantiation: Symphony kernels, distribution kind 1
#include "rb_kernels.cuh"
namespace rbhost {
template int stage_symphony<rb::kDistThermalJuettner>(const BatchArgs &, bool, int, cudaStream_t);
template int stage_symphony_diag<rb::kDistThermalJuettner>(const BatchArgs &, const DiagArgs &, int, cudaStream_t);
}
