/* pupil_host — C entry points of the host library (libpupil_host.so), for callers that cannot include the
 * C++ surface in pupiloptixlab_b200/host/ (Pupil::System, Pupil::pt::PTPass, Pupil::world::World ...):
 * the Python tests, bench.py and scripted use.  Everything here is a thin veneer over that C++ surface —
 * the same calls example/path_tracer/main.cpp makes (System::Init / AddPass / SetScene / Run) — which in
 * turn drives the CUDA back end only through include/pb2.h.
 *
 * All functions return 0 on success; pupil_last_error() describes the last failure.  Not thread-safe:
 * one caller thread, like the reference's render loop. */
#ifndef PUPIL_HOST_H
#define PUPIL_HOST_H
#include "pb2.h"

#ifdef __cplusplus
extern "C" {
#endif

/* System::Init(false) on a CUDA device + one PTPass registered with AddPass (example/path_tracer/main.cpp:5-12) */
int pupil_init(int device);
/* System::Destroy (main.cpp:20) */
int pupil_shutdown(void);
const char *pupil_last_error(void);
/* 0 silent, 1 warnings (default), 2 info */
int pupil_set_log_level(int level);

/* System::SetScene(path) (main.cpp:13-16; framework/system/system.cpp:136-173) */
int pupil_load_scene_xml(const char *path);
/* same dialect from memory; root_dir resolves relative file names (may be NULL) */
int pupil_load_scene_xml_string(const char *xml, const char *root_dir);
/* host-only variants (no CUDA device needed): parse + World::LoadScene precompute (camera matrices, material
 * precompute, emitter table), for inspection through the getters below; nothing is uploaded or rendered */
int pupil_parse_scene_xml(const char *path);
int pupil_parse_scene_xml_string(const char *xml, const char *root_dir);
/* make a triangle mesh available to scene files as <string name="filename" value="mem:KEY"/>
 * (ShapeManager::LoadMeshShape without the text round trip; arrays are copied; nrm / uv may be NULL) */
int pupil_register_mesh(const char *key, const float *pos, const float *nrm, const float *uv, const uint32_t *idx, uint32_t n_vertices,
                        uint32_t n_triangles);
/* the same without the copy: the caller keeps the four arrays alive (and unchanged) until pupil_unregister_mesh(key) or
 * pupil_clear_shapes.  Copying a 30 M-triangle mesh (840 MB) was 80 % of the host side of a scene reload; arrays in pinned
 * memory (cudaHostAlloc / cudaHostRegister) additionally make the upload a straight DMA. */
int pupil_register_mesh_borrowed(const char *key, const float *pos, const float *nrm, const float *uv, const uint32_t *idx, uint32_t n_vertices,
                                 uint32_t n_triangles);
/* forget one registered mesh; no loaded scene may still use it */
int pupil_unregister_mesh(const char *key);
/* forget every cached shape (ShapeManager::Clear) */
int pupil_clear_shapes(void);

/* PTPass knobs: max_depth <= 0 keeps the scene's integrator.max_depth; accumulate = the inspector checkbox
 * (pt_pass.cpp:256-268); frames_per_run = frames executed per PTPass::OnRun; first_seed / seed_stride = the
 * random_seed sequence (reference: 0, 1); sum_mode != 0 accumulates plain sums (multi-GPU shards). */
int pupil_pass_config(int max_depth, int accumulate, uint32_t frames_per_run, uint32_t first_seed, uint32_t seed_stride, int sum_mode);
/* Multi-GPU, one process per GPU (the reference is single-GPU).  rank 0 obtains the id and hands it to the other
 * ranks out of band; pupil_set_shard is collective (ncclCommInitRank) and binds this process, on the device of pupil_init, as
 * `rank` of `world`.  From then on one pass run renders this rank's seeds of the step (strong != 0: frames_per_run is the
 * step's TOTAL sample count, split over the ranks; else every rank renders frames_per_run), queues the NCCL reduction
 * (reduce_mode PB2_REDUCE_ROOT / PB2_REDUCE_ALL) and returns without waiting: "final result" holds sum / total spp after
 * pupil_synchronize (buffer downloads synchronise on their own).  world <= 0 switches sharding off. */
int pupil_comm_unique_id(uint8_t id[128]);
int pupil_set_shard(int rank, int world, const uint8_t *id, int strong, int reduce_mode);
int pupil_set_shard_plan(int strong); /* switch between the weak and the strong plan, keeping the communicator */
int pupil_synchronize(void);
int pupil_last_reduction(float *ms, uint64_t *bytes); /* pb2_comm_last_reduction of the pass's communicator */
/* System::Run() for n_pass_runs iterations of the pass list (headless: returns instead of looping forever) */
int pupil_run(uint64_t n_pass_runs);
/* frames (samples per pixel) accumulated so far and the next random_seed */
int pupil_pass_state(uint32_t *sample_cnt, uint32_t *random_seed);

/* BufferManager::GetBuffer(name): "final result", "pt accum buffer", "albedo", "normal", "test" */
/* PTPass::SaveCheckpoint / LoadCheckpoint: the progressive state (accum + frame buffers, sample_cnt, random_seed, pass settings;
 * pt_pass.cpp:55-56) to and from a file.  Load after the scene of the checkpoint was loaded; the run then continues with the
 * same seeds and running mean, bit-identical to an uninterrupted one.  The reference has no checkpoint. */
int pupil_checkpoint_save(const char *path);
int pupil_checkpoint_load(const char *path);
int pupil_buffer_info(const char *name, void **device_ptr, uint32_t *width, uint32_t *height, uint32_t *stride_in_byte);
int pupil_buffer_download(const char *name, void *host, uint64_t bytes);
int pupil_buffer_upload(const char *name, const void *host, uint64_t bytes);

/* introspection (parity tests): what World computed from the scene */
int pupil_get_film(uint32_t *width, uint32_t *height, uint32_t *max_depth);
int pupil_get_camera(float sample_to_camera[16], float camera_to_world[16], float *fov_y);
int pupil_num_instances(void);
int pupil_get_instance(uint32_t index, float xform[16], pb2_material *material, int32_t *emitter_offset, uint32_t *flags, uint32_t *n_prims,
                       int32_t *is_sphere);
/* bitmap textures / environment maps handed over in memory: key must start with "mem:" and is what the XML's
 * <string name="filename"> says; rgba = width*height float4 texels, row 0 = first row of the picture (copied) */
int pupil_register_image(const char *key, const float *rgba, uint32_t width, uint32_t height);
/* util::BitmapTexture::Load / Save (framework/util/texture.cpp:87-174, :13-85) on their own: hdr, exr, png, pfm in;
 * format 0 = hdr, 1 = exr, 2 = pfm, 3 = png as the reference's canvas shows the buffer (gamma 2.2; system/gui/output.hlsl:30-72,
 * gui.cpp:60-61), 4 = png with ACES tone mapping as well, out (rgba: row 0 = bottom of the picture, as in the frame buffers) */
int pupil_image_load(const char *path, uint32_t *width, uint32_t *height, float *rgba, uint64_t capacity_in_floats);
int pupil_image_save(const char *path, const float *rgba, uint32_t width, uint32_t height, int format);
/* saves a named device buffer of the current scene (float4 buffers only), e.g. "final result" */
int pupil_save_buffer(const char *name, const char *path, int format);
/* EmitterHelper's env-map tables (world/emitter.cpp:107-149): pass NULL arrays to query the sizes */
int pupil_get_env_tables(uint32_t *map_w, uint32_t *map_h, float *row_cdf, float *row_weight, float *col_cdf);
/* RenderObject::UpdateTransform (framework/world/render_object.cpp:72-80) on the index-th render object: row-major 4x4
 * object-to-world matrix.  Its area emitters are rebuilt (EmitterHelper::ResetAreaEmitter), the acceleration structure is
 * rebuilt on the next run and the pass restarts its accumulation, as after IAS::Update in the reference. */
int pupil_set_instance_transform(uint32_t index, const float xform[16]);
/* World::RemoveRenderObject (framework/world/world.cpp) on the index-th render object: the object, its entries of the emitter
 * table and its share of the selection probabilities go; later objects keep their order (their emitter offsets move down). */
int pupil_remove_instance(uint32_t index);
int pupil_num_area_emitters(void);
int pupil_get_emitters(pb2_emitter *areas, pb2_emitter *env, int32_t *has_env);
/* World::GetSceneHandle(): the pb2 scene (BVH built, camera and emitters uploaded) for pb2_trace_* etc. */
int pupil_scene_handle(pb2_scene **scene);
int pupil_set_bvh_builder(int builder);
/* World::SetInstancing: 0 flatten every shape, 1 bottom-level trees for shapes used by more than one render object (default), 2 for
 * every mesh (transform edits then never rebuild more than the top level) */
int pupil_set_instancing(int mode);
int pupil_build_stats(pb2_build_stats *stats);
int pupil_render_stats(pb2_render_stats *stats);
/* camera edits through CameraHelper (framework/world/camera.cpp): mark the pass dirty like the GUI does */
int pupil_camera_move(float dx, float dy, float dz);
int pupil_camera_rotate(float delta_x, float delta_y);
int pupil_camera_set_fov(float fov_y);

#ifdef __cplusplus
}
#endif
#endif
