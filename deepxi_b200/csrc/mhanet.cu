// MHANetV3 forward (deepxi/network/attention.py:387-442).  Placeholder translation unit: replaced by the
// real kernels in a later step of this round; until then the entry points refuse loudly.
#include "net.cuh"

namespace dxi {
int64_t mhanet_workspace_bytes(const dxi_net&, int, int) { return 256; }
int mhanet_forward(const dxi_net&, const float*, int, int, float*, void*, size_t, cudaStream_t) {
  set_error("MHANetV3 forward is not built yet");
  return DXI_E_STATE;
}
}  // namespace dxi
