// placeholder until the tcgen05 trunk lands
#include "common.cuh"
namespace dgdm {
size_t tc_trunk_workspace_bytes(int, int64_t, int) { return 256; }
int tc_trunk(const dgdm_dyn_weights*, const float*, const float*, const float*, int, int, int, const int32_t*, int,
             const dgdm_objective*, bool, float*, float*, float*, void*, size_t, int, cudaStream_t) {
  set_error("tensor-core trunk not built");
  return DGDM_EUNSUPPORTED;
}
}  // namespace dgdm
extern "C" size_t dgdm_dyn_tc_image_bytes(int32_t) { return 256; }
extern "C" int dgdm_dyn_pack_tc(const dgdm_dyn_weights*, void*, void*) {
  dgdm::set_error("tensor-core trunk not built");
  return DGDM_EUNSUPPORTED;
}
