// tcgen05 / TMA / TMEM nearest-code search -- placeholder until the tensor-core kernel lands.
#include "vq_kernels.h"

namespace vq {
bool tc_supported(int64_t, int, int) { return false; }
size_t tc_workspace_bytes(int64_t, int, int) { return 0; }
cudaError_t launch_dist_tc(const __half*, const float*, const float*, const CodebookView&, int64_t, int*, int*, int*,
                           int64_t*, void*, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace vq
