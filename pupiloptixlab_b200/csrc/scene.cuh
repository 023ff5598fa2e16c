// Host-side scene object behind the pb2 C ABI.
#pragma once
#include "pb2_types.cuh"
#include "util.cuh"
#include <memory>
#include <string>
#include <vector>

namespace pb2 {

// bottom-level tree of a mesh that is placed through instance nodes (bvh_build.cu): object-space records and nodes inside the
// scene's arrays — GAS of the reference (framework/world/gas_manager.cpp)
struct Blas {
    bool valid = false;
    uint32_t node_offset = 0, n_nodes = 0, prim_offset = 0, n_prims = 0, depth = 0;
    float lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 }; // object-space bounds of the root
};
struct Mesh {
    DevBuf<float> pos, nrm, uv;
    DevBuf<uint32_t> idx;
    uint32_t n_verts = 0, n_tris = 0;
    Blas blas;
};

struct Wavefront; // wavefront.cu

struct Scene {
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::vector<std::unique_ptr<Mesh>> meshes;
    std::vector<DevInstance> h_inst;
    std::vector<int> h_inst_mesh; // mesh id of every instance (-1: analytic sphere)
    std::vector<DevMaterial> h_mat;
    std::vector<DevEmitter> h_areas;
    bool has_env = false;
    DevEmitter h_env{};
    Camera cam{};

    DevBuf<DevInstance> d_inst;
    DevBuf<DevMaterial> d_mat;
    DevBuf<DevEmitter> d_areas, d_env;
    DevBuf<float> d_area_cdf; // running sums of the area emitters' select probabilities (select_emitter)
    bool tables_dirty = true;

    DevBuf<Bvh8Node> d_nodes;
    DevBuf<PrimRec> d_prims;
    DevBuf<uint32_t> trace_work; // work counter of the persistent trace kernels
    uint32_t n_nodes = 0, n_prims = 0;
    bool bvh_valid = false;
    // two-level structure (bvh_build.cu): meshes with a bottom-level tree are reached through instance nodes of the top level
    int instancing = 1;          // 0 flatten everything, 1 bottom-level trees for meshes placed more than once, 2 for every mesh
    uint32_t n_blas = 0;         // bottom-level trees in the node array
    uint32_t top_node_offset = 0, root = 0; // the top level sits behind them; traversal starts at `root`
    bool blas_valid = false;     // the bottom-level trees match the meshes and instances: only the top level needs rebuilding
    int builder = 0; // 0 LBVH, 1 binned SAH sweep along the Morton order, 2 SAH-driven bottom-up clustering (bvh_ploc.cu)
    int ploc_radius = 8; // neighbours searched on either side along the Morton curve (builder 2)
    // binary tree -> BVH8: 1 = the cut that minimises the SAH cost of the wide tree (tables filled bottom-up by k_refit<true>; LBVH
    // and the top level), 0 = greedy expansion of the child with the largest surface area.  collapse_prim_cost_pct: cost of one
    // primitive test in per cent of one wide-node test
    int collapse = 1, collapse_prim_cost_pct = 30;
    int morton_bits = 0; // leading bits of the 63-bit Morton key that are sorted (and seen by the radix tree), 8 per sort pass; 0 = by primitive count
    pb2_build_stats build_stats{};

    // options
    bool profiling = false, counting = false;
    int sort_by_material = -1;  // 1 on, 0 off, -1 auto (on when the scene has more than one material type)
    int n_material_types = 0;   // distinct EMatType values among the instances (upload_tables)
    int only_material_type = -1; // that type when there is exactly one, else -1
    uint32_t material_type_mask = 0; // bit t: some instance has material type t
    uint64_t paths_in_flight = 0;
    int refill_threshold = 26;
    int shade_variant = 6;     // k_shade<MINB>: 4, 6, 7 or 8 resident CTAs per SM
    // two batches in flight on two streams, the second one kernel late, so that one batch's HBM-bound shade kernel meets the
    // other's issue-bound trace kernels.  Measured +1.9 % (Cornell), +5.4 % (material grid), +1.5 % (terrain): k_shade alone
    // fills the register file, so the kernels mostly time-share the SMs instead of co-residing.  Off by default: with it a
    // kernel's event duration includes the time it shares the GPU, which would blur the per-kernel numbers bench.py reports.
    bool two_lanes = false;
    int coop_prims = -1;       // warp-cooperative primitive tests in the trace kernels: 1 on, 0 off, -1 auto (by scene size)
    // auto: few lanes reach a leaf per step in deep trees (cooperation pays); in tiny scenes every lane does (it only costs)
    bool use_coop_prims() const { return coop_prims == 1 || (coop_prims < 0 && n_prims >= 4096u); } // persistent traversal: refill idle lanes when fewer than this many are busy

    // "L2-resident top levels": an access-policy window over the first bytes of the node array (the collapse writes the wide
    // tree level by level, so these are the top levels) marks them persisting in L2 for every kernel on the scene's stream.
    // 0 = off (default; measured in profiles/README.md).  l2_window_mb > l2_persist_mb spreads the carve-out over a larger
    // window with hit ratio persist / window.
    int l2_persist_mb = 0, l2_window_mb = 0;
    bool l2_dirty = false;
    void apply_l2_window(); // pb2_api.cu; called by the launch sites when l2_dirty

    Wavefront *wf = nullptr;
    pb2_render_stats render_stats{};
    // recorded by pb2_comm_reduce_frames on the communicator's stream while it reads the sum buffer; the next k_accumulate
    // (the only writer of that buffer) waits for it, everything before it overlaps the reduction (comm.cu)
    cudaEvent_t accumulate_gate = nullptr;
    bool gate_pending = false;

    Scene();
    ~Scene();
    void upload_tables();
    void check_emitter_ranges() const; // every emitting instance's [offset, offset + n_tris) lies inside the emitter table (throws)
    SceneView view() const;
};

// bvh_build.cu
void build_bvh(Scene &s);
uint32_t max_index_dev(const uint32_t *idx, uint64_t n, cudaStream_t st);
// trace.cu
void trace_closest_dev(Scene &s, const float4 *rays, uint64_t n, float4 *hit_tuvp, int32_t *hit_inst);
void trace_any_dev(Scene &s, const float4 *rays, uint64_t n, uint32_t *occluded);
// wavefront.cu
void render(Scene &s, const pb2_launch_params &p);
void finalize_sum(Scene &s, const float4 *sum, float4 *frame, uint64_t n, uint32_t spp);
void wavefront_destroy(Wavefront *wf);
// comm.cu
struct Comm;
void comm_unique_id(uint8_t id[PB2_COMM_ID_BYTES]);
Comm *comm_create(int n_ranks, int rank, const uint8_t *id);
void comm_destroy(Comm *c);
void comm_reduce_frames(Comm &c, Scene &s, const float4 *sum, float4 *frame, uint64_t n_pixels, uint32_t total_spp, int mode, int root);
void comm_synchronize(Comm &c);
void comm_last_reduction(Comm &c, float *ms, uint64_t *bytes);
int comm_nccl_version();
// test_hooks/kat.cu (libpb2_kat.so, not part of libpb2.so)
int run_kat(const char *what, const void *in0, const void *in1, const void *in2, uint64_t n, void *out);
}// namespace pb2
