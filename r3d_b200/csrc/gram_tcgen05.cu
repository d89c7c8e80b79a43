// tcgen05 / TMA Gram kernel -- placeholder until the tensor-core path lands; the
// SIMT Gram in erank_kernels.cu is used meanwhile (gram_tcgen05_supported == false).
#include "common.cuh"
namespace r3d {
bool gram_tcgen05_supported(int64_t, int64_t, int64_t, int) { return false; }
int gram_tcgen05_launch(const void*, int64_t, int64_t, int64_t, int, void*, float*, cudaStream_t) {
  set_error("tcgen05 Gram not built");
  return 1;
}
}  // namespace r3d
