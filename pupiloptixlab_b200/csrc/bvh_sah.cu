// Binned-SAH binary BVH builder (placeholder until the top-down builder lands): reports "unavailable" so
// build_bvh falls back to the LBVH path.
#include "scene.cuh"
namespace pb2 {
bool sah_builder_available() { return false; }
void build_binary_sah(cudaStream_t, uint32_t, const float4 *, const float4 *, const int *, int *, int *, int2 *, float4 *, float4 *, uint32_t *) {
    throw std::runtime_error("binned SAH builder not available");
}
}// namespace pb2
