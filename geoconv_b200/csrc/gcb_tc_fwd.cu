// placeholder until the tcgen05 forward lands
#include "gcb_common.cuh"
namespace gcb {
bool tc_supported(int, int, int, int, int) { return false; }
size_t tc_fwd_workspace_bytes(int, int, int, int, int, int, int) { return 0; }
int tc_conv_fwd(const float*, const int32_t*, const float*, const float*, const float*, const float*,
                const RotTable&, int, int, int, int, int, int, int, int, int, float*, void*, size_t,
                cudaStream_t) {
  set_error("tcgen05 forward not built");
  return GCB_EUNSUPPORTED;
}
}  // namespace gcb
