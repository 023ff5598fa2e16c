/* pb2 — the C ABI of the B200-native path-tracing back end.
 *
 * This is the boundary the reference's kept C++ surface (Pupil::world::World, Pupil::pt::PTPass,
 * Pupil::BufferManager — re-implemented in pupiloptixlab_b200/host/) calls instead of OptiX 7.5:
 * plain pointers and sizes, `int` status codes (0 = ok, message via pb2_last_error()), no C++ or
 * torch types.  The caller owns every host array (copied during the call); the library owns all
 * device memory except buffers explicitly passed in as device pointers.  One host thread per scene
 * handle.  Citations name the reference interface each entry point replaces
 * (paths relative to the reference root).
 *
 * Enum values are the reference's own:
 *   material type  Pupil::EMatType        framework/render/material/predefine.h:15-22
 *   texture type   util::ETextureType     framework/util/texture.h:21-25
 *   emitter type   optix::EEmitterType    framework/render/emitter/types.h:7-15
 */
#ifndef PB2_H
#define PB2_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB2_OK 0
#define PB2_ERR_CUDA 1
#define PB2_ERR_ARG 2
#define PB2_ERR_STATE 3

enum { PB2_MAT_UNKNOWN = 0, PB2_MAT_DIFFUSE = 1, PB2_MAT_DIELECTRIC = 2, PB2_MAT_ROUGH_DIELECTRIC = 3, PB2_MAT_CONDUCTOR = 4,
       PB2_MAT_ROUGH_CONDUCTOR = 5, PB2_MAT_PLASTIC = 6, PB2_MAT_ROUGH_PLASTIC = 7 };
enum { PB2_TEX_RGB = 0, PB2_TEX_BITMAP = 1, PB2_TEX_CHECKERBOARD = 2 };
enum { PB2_EMIT_NONE = 0, PB2_EMIT_TRI = 1, PB2_EMIT_SPHERE = 2, PB2_EMIT_CONST_ENV = 3, PB2_EMIT_ENV_MAP = 4 };

/* instance flags */
#define PB2_INST_FLIP_NORMALS 1u /* TriMesh::flip_normals / Sphere::flip_normal, framework/render/geometry.h:19,26 */
#define PB2_INST_FLIP_TEX 2u     /* TriMesh::flip_tex_coords, geometry.h:20 */
#define PB2_MESH_SPHERE 0xFFFFFFFFu /* mesh id of the analytic unit sphere (centre 0, radius 1), world/render_object.cpp:30-35 */

/* cudaTextureAddressMode / cudaTextureFilterMode values, = util::ETextureAddressMode / ETextureFilterMode
 * (framework/util/texture.h:10-20) */
enum { PB2_ADDR_WRAP = 0, PB2_ADDR_CLAMP = 1, PB2_ADDR_MIRROR = 2, PB2_ADDR_BORDER = 3 };
enum { PB2_FILTER_POINT = 0, PB2_FILTER_LINEAR = 1 };

/* cuda::Texture, framework/cuda/texture.h:10-31 — RGB constant, checkerboard or bitmap; r0/r1 = rows 0 and 1 of
 * the to_uv transform (the only rows Sample() reads, :34-36).  `bitmap` is a handle from pb2_bitmap_create. */
typedef struct pb2_texture {
    int32_t type;
    float a[3]; /* rgb | patch1 */
    float b[3]; /*       patch2 */
    float r0[4], r1[4];
    uint32_t pad0;
    uint64_t bitmap; /* PB2_TEX_BITMAP: cudaTextureObject_t of a float4 array, normalised coordinates */
} pb2_texture;

/* optix::material::Material after LoadMaterial (framework/render/material/optix_material.h:10-33,
 * optix_material.cpp:41-130), flattened.  Texture slots by type — tex[0] is always the texture
 * LocalBsdf::GetAlbedo() returns (optix_material.h:93-111):
 *   diffuse          tex0 reflectance
 *   dielectric       tex0 specular_reflectance  tex1 specular_transmittance
 *   roughdielectric  tex0 specular_reflectance  tex1 specular_transmittance  tex2 alpha
 *   conductor        tex0 specular_reflectance  tex1 eta  tex2 k
 *   roughconductor   tex0 specular_reflectance  tex1 eta  tex2 k  tex3 alpha
 *   plastic          tex0 diffuse_reflectance   tex1 specular_reflectance
 *   roughplastic     tex0 diffuse_reflectance   tex1 specular_reflectance    tex2 alpha */
typedef struct pb2_material {
    int32_t type;
    int32_t twosided;
    float eta;                      /* int_ior / ext_ior */
    int32_t nonlinear;
    float int_fdr;                  /* m_int_fdr */
    float specular_sampling_weight; /* m_specular_sampling_weight */
    pb2_texture tex[4];
} pb2_material;

/* optix::Emitter, framework/render/emitter.h:13-23 with emitter/{area,sphere,env}.h payloads */
typedef struct pb2_emitter {
    int32_t type;
    float weight, select_probability;
    pb2_texture radiance; /* const env: radiance.a = color */
    float area;
    float pos[3][3], nrm[3][3], uv[3][2]; /* TriArea: world-space v0..v2 */
    float center[3], radius;              /* Sphere; EnvMap / ConstEnv: center = scene AABB centre */
    /* EnvMapEmitter, framework/render/emitter/env.h:6-22 (radiance = the bitmap texture) */
    float scale, normalization;
    uint32_t map_w, map_h;
    float to_world[9], to_local[9];       /* rows r0,r1,r2 */
    const float *env_tables;              /* DEVICE memory (pb2_malloc): row_cdf[map_h + 1], row_weight[map_h],
                                             col_cdf[(map_w + 1) * map_h] back to back (world/emitter.cpp:107-149) */
} pb2_emitter;

typedef struct pb2_hit {
    float t, u, v;
    int32_t inst, prim; /* inst = -1: miss */
} pb2_hit;

/* pt::OptixLaunchParams, example/path_tracer/type.h:9-33 (the camera, emitter group and AS handle live
 * in the scene handle).  Buffers are DEVICE pointers, row-major, pixel_index = y*width + x, row 0 = bottom. */
typedef struct pb2_launch_params {
    uint32_t max_depth;
    uint32_t accumulate;   /* 0: overwrite, 1: running mean (main.cu:190-194), 2: plain sums, w = frames summed (multi-GPU
                              shards); in modes 1 and 2 sample_cnt = 0 starts over without reading the buffer */
    uint32_t width, height;
    uint32_t random_seed;  /* seed of the first frame */
    uint32_t seed_stride;  /* frame i uses random_seed + i*seed_stride (0 is read as 1) */
    uint32_t sample_cnt;   /* frames already in accum_buffer */
    uint32_t n_frames;     /* consecutive PTPass::OnRun calls to execute (>= 1) */
    void *accum_buffer;    /* float4, required */
    void *frame_buffer;    /* float4, may be NULL */
    void *normal_buffer;   /* float3, may be NULL */
    void *albedo_buffer;   /* float3, may be NULL */
    void *test_buffer;     /* float,  may be NULL */
} pb2_launch_params;

typedef struct pb2_build_stats {
    uint64_t n_prims, n_triangles, n_spheres;
    uint64_t n_nodes;      /* BVH8 nodes (80 B each) */
    uint64_t bvh_bytes;    /* nodes + primitive records */
    float build_ms;        /* device time, CUDA events around the whole build */
    float sah_cost;        /* SAH cost of the wide tree (node cost 1, primitive cost 1), root area normalised */
    uint32_t max_depth;    /* depth of the wide tree (two-level: top level + 1 + deepest bottom-level tree) */
    uint32_t n_blas;       /* bottom-level trees (meshes reached through instance nodes); 0 = everything flattened to world space */
    uint64_t n_instance_leaves; /* placements of those meshes in the top level */
    float top_level_ms;    /* the top-level part of build_ms (all of it after pb2_scene_set_instance_transform) */
    uint32_t pad0;
} pb2_build_stats;

typedef struct pb2_render_stats {
    uint64_t closest_rays, shadow_rays; /* rays traced by the last pb2_render call */
    uint64_t kernel_launches;           /* kernels launched by the last pb2_render call */
    float total_ms;                     /* device time of the last pb2_render call (CUDA events on the scene's stream) */
    float generate_ms, extend_ms, shade_ms, shadow_ms, accumulate_ms; /* per stage, only when profiling is on */
    uint64_t nodes_visited, prims_tested;                             /* traversal counters (closest + shadow), only when counting is on */
    uint64_t nodes_shadow, prims_shadow;                              /* the shadow-ray share of the two counters above */
    uint32_t batches, rounds;                                         /* wavefront batches executed, bounce rounds per batch */
    uint32_t extend_launches, shade_launches, shadow_launches, other_launches; /* kernel launches by stage */
    uint64_t shaded_paths;                                            /* path-vertex shading invocations (= closest rays traced) */
    uint64_t shadow_unoccluded;                                       /* shadow rays that reached the light, only when counting is on */
    uint32_t sorted;                                                  /* 1: material-sorted shading (k_bin ran), 0: in-order shading */
    uint32_t pad0;
} pb2_render_stats;

/* ---- library / device ------------------------------------------------------------------------------- */
/* replaces cuda::Context::Init + optix::Context::Init (framework/cuda/context.cpp:15-44, optix/context.cpp:38-50) */
int pb2_init(int device);
const char *pb2_last_error(void);
int pb2_device_count(void);

/* ---- device memory (BufferManager / CudaMemcpyToDevice, framework/system/buffer.cpp:39-99, cuda/util.h:40-75) */
int pb2_malloc(void **dptr, uint64_t bytes); /* zero-initialised, like Buffer allocation (buffer.cpp:44-45) */
int pb2_free(void *dptr);
/* Device blocks freed by the library (pb2_free, scene destruction, BVH rebuilds) are cached per device and reused:
 * cudaMalloc/cudaFree of GB-sized arrays cost 0.3-0.8 s per scene reload on B200 (the reference frees and reallocates
 * through cudaMalloc on every SetScene, system.cpp:143-165).  pb2_trim returns the cache to the driver. */
int pb2_trim(void);
int pb2_upload(void *dptr, const void *host, uint64_t bytes);
int pb2_download(void *host, const void *dptr, uint64_t bytes);
int pb2_memset(void *dptr, int value, uint64_t bytes);

/* ---- bitmaps: CudaTextureManager::GetCudaTextureObject, framework/cuda/texture.cpp:60-102 ------------- */
/* rgba: HOST float4 texels, row 0 first (file order), copied into a cudaArray; the handle is a cudaTextureObject_t
 * with normalised coordinates, element read mode, the given address mode on both axes and filter mode. */
int pb2_bitmap_create(const float *rgba, uint32_t width, uint32_t height, int address_mode, int filter_mode, uint64_t *handle);
int pb2_bitmap_destroy(uint64_t handle);

/* ---- scene ------------------------------------------------------------------------------------------ */
typedef struct pb2_scene pb2_scene;
int pb2_scene_create(pb2_scene **scene);
int pb2_scene_destroy(pb2_scene *scene);
int pb2_scene_clear(pb2_scene *scene);
/* run the scene's kernels on an existing CUDA stream (cudaStream_t); NULL restores the scene's own stream.
 * replaces cuda::Stream (framework/cuda/stream.cpp:8) */
int pb2_scene_set_stream(pb2_scene *scene, void *cuda_stream);

/* object-space triangle mesh shared by instances: ShapeManager mesh data + CudaShapeDataManager upload
 * (framework/resource/shape.cpp:193-226,268-327).  nrm / uv may be NULL. */
int pb2_scene_add_mesh(pb2_scene *scene, const float *pos, const float *nrm, const float *uv, const uint32_t *idx, uint32_t n_vertices,
                       uint32_t n_triangles, uint32_t *mesh_id);
/* one RenderObject (framework/world/render_object.cpp:18-65) + its SBT hit record
 * (example/path_tracer/pt_pass.cpp:180-207): geometry, 3x4 object->world transform (row-major),
 * material, emitter_index_offset (-1: not an emitter).  mesh_id = PB2_MESH_SPHERE for spheres. */
int pb2_scene_add_instance(pb2_scene *scene, uint32_t mesh_id, const float xform[12], uint32_t flags, const pb2_material *material,
                           int32_t emitter_index_offset, uint32_t *instance_id);
/* EmitterHelper::GetEmitterGroup (framework/world/emitter.cpp:339-390); env may be NULL */
int pb2_scene_set_emitters(pb2_scene *scene, const pb2_emitter *areas, uint32_t n_areas, const pb2_emitter *env);
/* CameraHelper::GetCudaMemory (framework/world/camera.cpp:72-93): two row-major 4x4 */
int pb2_scene_set_camera(pb2_scene *scene, const float sample_to_camera[16], const float camera_to_world[16]);

/* replaces GAS::Create + IAS::Create / IAS::Update (framework/world/gas_manager.cpp:69-245, ias_manager.cpp:29-151): GPU build of
 * the compressed 8-wide BVH.  Two levels, as in the reference: a mesh that is placed more than once (option instancing = 1, the
 * default; 2 = every mesh of at least 64 triangles; 0 = none) gets ONE bottom-level tree in object space, shared by all its
 * placements (GASManager::RefGAS, gas_manager.cpp:10); the top level holds an instance node per placement next to the
 * world-space triangles of everything else and the analytic spheres.  After pb2_scene_set_instance_transform only the top
 * level is rebuilt (stats->top_level_ms).  stats may be NULL. */
int pb2_bvh_build(pb2_scene *scene, pb2_build_stats *stats);
/* RenderObject::UpdateTransform -> IASManager::UpdateInstance (framework/world/render_object.cpp:72-80, ias_manager.cpp:116-151):
 * new 3x4 object->world transform of one instance; the next pb2_bvh_build keeps every bottom-level tree */
int pb2_scene_set_instance_transform(pb2_scene *scene, uint32_t instance_id, const float xform[12]);
/* builder: 0 = LBVH over 63-bit Morton codes; 1 = binned-SAH sweep along the Morton order (16 equal-count bins per node, exact
 * sweep below 17 primitives; bvh_sah.cu — measured worse than LBVH: splits that do not fall on octree-cell boundaries of the
 * Z-curve produce overlapping boxes); 2 = SAH-driven bottom-up clustering (bvh_ploc.cu): mutual nearest neighbours by the
 * surface area of their union, searched `ploc_radius` clusters to either side along the Morton curve — the fast-trace build
 * the reference asks OptiX for (OPTIX_BUILD_FLAG_PREFER_FAST_TRACE, gas_manager.cpp:191).  Measured numbers: DESIGN.md 8. */
int pb2_scene_set_builder(pb2_scene *scene, int builder);

/* parity / benchmark hooks for the two optixTrace flavours (main.cu:80-85,161-166; emitter.h:91-100).
 * rays: n x 8 floats (ox oy oz tmin dx dy dz tmax).  Host-pointer versions copy in and out. */
int pb2_trace_closest(pb2_scene *scene, const float *rays, uint64_t n, pb2_hit *hits);
int pb2_trace_any(pb2_scene *scene, const float *rays, uint64_t n, uint8_t *occluded);
/* device-pointer versions: rays_dev float4[2n]; hits_tuvp float4[n] (t,u,v,bits(prim)); hits_inst int32[n];
 * asynchronous on the scene's stream */
int pb2_trace_closest_dev(pb2_scene *scene, const void *rays_dev, uint64_t n, void *hits_tuvp_dev, void *hits_inst_dev);
int pb2_trace_any_dev(pb2_scene *scene, const void *rays_dev, uint64_t n, void *occluded_u32_dev);
/* read back the wide BVH (for the oracle-side node/primitive counters): sizes first with NULL pointers */
int pb2_bvh_download(pb2_scene *scene, void *nodes, uint64_t *n_nodes, void *prims, uint64_t *n_prims);

/* ---- rendering -------------------------------------------------------------------------------------- */
/* replaces optix::Pass::Run + Synchronize (framework/optix/pass.h:71-93) as driven by PTPass::OnRun
 * (example/path_tracer/pt_pass.cpp:39-57): executes params->n_frames frames.  pb2_render is asynchronous on
 * the scene's stream; pb2_synchronize waits. */
int pb2_render(pb2_scene *scene, const pb2_launch_params *params);
int pb2_synchronize(pb2_scene *scene);
int pb2_render_stats_get(pb2_scene *scene, pb2_render_stats *stats); /* synchronises */
/* options: profiling (per-stage events, serialises stages), counting (traversal counters),
 * paths_in_flight (0 = default), sort_by_material (1 on, 0 off, -1 = auto: on when the scene has more than one
 * material type; default -1), refill_threshold (persistent traversal), shade_variant (4 | 6 | 7 | 8 resident CTAs per SM),
 * two_lanes (two batches in flight on two streams), coop_prims (warp-cooperative primitive tests: 1 on, 0 off, -1 = auto
 * by scene size), l2_persist_mb / l2_window_mb (persisting-L2 access-policy window over the top levels of the node array;
 * 0 = off, the default), ploc_radius (builder 2: neighbours searched on either side, 1..16, default 8), instancing (0 | 1 | 2, see pb2_bvh_build),
 * collapse (binary tree -> BVH8: 1 = SAH-optimal cut by dynamic programming, the default; 0 = greedy largest-area expansion) and
 * collapse_prim_cost_pct (cost of a primitive test in per cent of a wide-node test, default 30), morton_bits (leading bits of the
 * 63-bit Morton key that are sorted, 15..63; 0 = chosen by primitive count, the default).  Unknown names fail with PB2_ERR_ARG. */
int pb2_scene_set_option(pb2_scene *scene, const char *name, int64_t value);
/* frame[i] = (sum[i].xyz / total_spp, 1) on the scene's stream: the last step of a sharded render whose sums were combined
 * by the caller's own collective (pb2_comm_reduce_frames below does both).   */
int pb2_finalize_sum(pb2_scene *scene, const void *sum_buffer, void *frame_buffer, uint64_t n_pixels, uint32_t total_spp);

/* ---- multi-GPU: one process per GPU, sample-index shards, NCCL over NVLink --------------------------
 * The reference renders on one GPU (example/path_tracer/pt_pass.cpp:39-57); these entry points have no counterpart there.
 * Every rank loads the same scene, renders the seeds pb2_shard_plan gives it with accumulate = 2 and then calls
 * pb2_comm_reduce_frames, which adds the ranks' sum buffers and writes frame = sum / total_spp.  NCCL is loaded at run time
 * (libnccl.so.2); a communicator of one rank needs none. */
typedef struct pb2_comm pb2_comm;
#define PB2_COMM_ID_BYTES 128
#define PB2_REDUCE_ROOT 0 /* ncclReduce to `root`, which finalizes: only the root's frame buffer is written */
#define PB2_REDUCE_ALL 1  /* ncclReduceScatter, each rank finalizes 1/N of the pixels, ncclAllGather: every rank gets the frame
                             (falls back to PB2_REDUCE_ROOT when the ranks do not divide the pixel count) */
int pb2_comm_unique_id(uint8_t id[PB2_COMM_ID_BYTES]); /* rank 0; the caller hands the bytes to the other ranks */
/* collective over the ranks; binds the communicator to the CURRENT device (pb2_init) */
int pb2_comm_create(pb2_comm **comm, int n_ranks, int rank, const uint8_t id[PB2_COMM_ID_BYTES]);
int pb2_comm_destroy(pb2_comm *comm);
/* Asynchronous: ordered after everything queued on the scene's stream, runs on the communicator's stream.  The scene's next
 * pb2_render overlaps it up to its first accumulate kernel (the only writer of sum_buffer), which waits; read the frame
 * after pb2_comm_synchronize.  frame_buffer may be NULL on non-root ranks in PB2_REDUCE_ROOT mode. */
int pb2_comm_reduce_frames(pb2_comm *comm, pb2_scene *scene, const void *sum_buffer, void *frame_buffer, uint64_t n_pixels, uint32_t total_spp,
                           int mode, int root);
int pb2_comm_synchronize(pb2_comm *comm);
/* device time of the last pb2_comm_reduce_frames (collectives + finalize) and the bytes of one rank's sum buffer; synchronises */
int pb2_comm_last_reduction(pb2_comm *comm, float *ms, uint64_t *bytes);
int pb2_comm_nccl_version(int *version);
/* the seeds of `rank` in progressive step `step`: first_seed + k * seed_stride, k < spp_rank.  strong = 0: every rank renders
 * `spp` frames (the step holds spp * n_ranks); strong = 1: the step's `spp` frames are split over the ranks. */
int pb2_shard_plan(int rank, int n_ranks, uint32_t step, uint32_t spp, int strong, uint32_t *first_seed, uint32_t *seed_stride,
                   uint32_t *spp_rank, uint32_t *spp_total);

#ifdef __cplusplus
}
#endif
#endif
