// Fused read-once stage-1 path (placeholder until the streaming kernel lands).
#include "tile_scan_shared.cuh"

namespace hipac {
bool fused_available() { return false; }
size_t fused_workspace_bytes(int, int, int, int, int, int, int) { return 0; }
int fused_scan(const ScanParams&, const OutParams&, uint8_t*, int32_t*, int32_t*, uint8_t*, int32_t*, int, uint8_t*,
               cudaStream_t) {
  set_error("fused scan not built");
  return -4;
}
}  // namespace hipac
