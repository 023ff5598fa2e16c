// extern "C" veneer over the C++ host surface (include/pupil_host.h).  No logic of its own.
#include "../../include/pupil_host.h"
#include "image.h"
#include "pt_pass.h"

#include <cstring>
#include <memory>
#include <string>

using namespace Pupil;

namespace {
std::string g_error;
std::unique_ptr<pt::PTPass> g_pass;
pb2_comm *g_comm = nullptr; // the communicator of pupil_set_shard (owned here; the pass borrows it)
int g_rank = 0, g_world = 1, g_reduce_mode = PB2_REDUCE_ALL;
void SyncIfSharded() {
    if (g_pass && g_pass->IsSharded()) g_pass->Synchronize(); // a sharded OnRun returns before the reduction has finished
}
int Fail(const std::string &m) {
    g_error = m;
    return 1;
}
world::World *W() { return util::Singleton<world::World>::instance(); }
System *Sys() { return util::Singleton<System>::instance(); }
bool Ready() { return Sys()->IsInitialized() && g_pass; }
bool HasWorld() { return W()->scene && W()->camera && W()->emitters; }
}// namespace

extern "C" {
const char *pupil_last_error(void) { return g_error.c_str(); }
int pupil_set_log_level(int level) {
    Log::level = level;
    return 0;
}
int pupil_init(int device) {
    if (Sys()->IsInitialized()) pupil_shutdown();
    Sys()->device = device;
    Sys()->Init(false);
    if (!Sys()->IsInitialized()) return Fail(std::string("System::Init failed: ") + pb2_last_error());
    g_pass = std::make_unique<pt::PTPass>("Path Tracing");
    Sys()->AddPass(g_pass.get());
    return 0;
}
int pupil_shutdown(void) {
    if (g_pass) g_pass->SetShard(nullptr, 0, 1, false, PB2_REDUCE_ALL);
    if (g_comm) pb2_comm_destroy(g_comm), g_comm = nullptr;
    if (g_pass) Sys()->RemovePass(g_pass.get());
    g_pass.reset();
    if (Sys()->IsInitialized()) Sys()->Destroy();
    return 0;
}
int pupil_load_scene_xml(const char *path) {
    if (!Ready()) return Fail("pupil_init first");
    if (!path || !std::filesystem::exists(path)) return Fail(std::string("scene file does not exist: ") + (path ? path : "(null)"));
    if (!Sys()->SetScene(std::filesystem::path(path)) || !g_pass->GetLaunchParams().accum_buffer) return Fail("scene load failed");
    return 0;
}
int pupil_load_scene_xml_string(const char *xml, const char *root_dir) {
    if (!Ready()) return Fail("pupil_init first");
    if (!xml) return Fail("null xml");
    if (!W()->scene->LoadFromXMLString(xml, root_dir ? std::filesystem::path(root_dir) : std::filesystem::path())) {
        Sys()->SetScene(static_cast<resource::Scene *>(nullptr)); // fails, and clears what the previous scene left behind
        return Fail("XML parse failed");
    }
    if (!Sys()->SetScene(W()->scene.get()) || !g_pass->GetLaunchParams().accum_buffer) return Fail("scene load failed");
    return 0;
}
int pupil_parse_scene_xml(const char *path) {
    if (!HasWorld()) W()->Init();
    if (!path || !std::filesystem::exists(path)) return Fail(std::string("scene file does not exist: ") + (path ? path : "(null)"));
    if (!W()->scene->LoadFromXML(std::filesystem::path(path)) || !W()->LoadScene(W()->scene.get())) return Fail("scene load failed");
    return 0;
}
int pupil_parse_scene_xml_string(const char *xml, const char *root_dir) {
    if (!HasWorld()) W()->Init();
    if (!xml) return Fail("null xml");
    if (!W()->scene->LoadFromXMLString(xml, root_dir ? std::filesystem::path(root_dir) : std::filesystem::path()) || !W()->LoadScene(W()->scene.get()))
        return Fail("XML parse failed");
    return 0;
}
int pupil_register_mesh(const char *key, const float *pos, const float *nrm, const float *uv, const uint32_t *idx, uint32_t nv, uint32_t nf) {
    if (!key || std::strncmp(key, "mem:", 4) != 0) return Fail("mesh keys must start with \"mem:\"");
    if (!util::Singleton<resource::ShapeManager>::instance()->LoadMeshShape(key, pos, nrm, uv, idx, nv, nf)) return Fail("empty mesh");
    return 0;
}
int pupil_register_mesh_borrowed(const char *key, const float *pos, const float *nrm, const float *uv, const uint32_t *idx, uint32_t nv, uint32_t nf) {
    if (!key || std::strncmp(key, "mem:", 4) != 0) return Fail("mesh keys must start with \"mem:\"");
    if (!util::Singleton<resource::ShapeManager>::instance()->LoadMeshShape(key, pos, nrm, uv, idx, nv, nf, /*borrow=*/true)) return Fail("empty mesh");
    return 0;
}
int pupil_unregister_mesh(const char *key) {
    if (!key) return Fail("null key");
    util::Singleton<resource::ShapeManager>::instance()->DropMeshShape(key);
    return 0;
}
int pupil_register_image(const char *key, const float *rgba, uint32_t width, uint32_t height) {
    if (!key || std::strncmp(key, "mem:", 4) != 0) return Fail("image keys must start with \"mem:\"");
    optix::material::ClearDeviceBitmaps(); // a replaced image may reuse the address a cached texture object was made from
    if (!util::Singleton<resource::TextureManager>::instance()->RegisterImage(key, rgba, width, height)) return Fail("empty image");
    return 0;
}
int pupil_image_load(const char *path, uint32_t *width, uint32_t *height, float *rgba, uint64_t capacity_in_floats) {
    util::Image img;
    if (!path || !util::LoadImage(path, img)) return Fail(std::string("cannot load image: ") + (path ? path : "(null)"));
    if (width) *width = static_cast<uint32_t>(img.w);
    if (height) *height = static_cast<uint32_t>(img.h);
    if (rgba) {
        if (capacity_in_floats < img.rgba.size()) return Fail("pupil_image_load: buffer too small");
        std::memcpy(rgba, img.rgba.data(), img.rgba.size() * sizeof(float));
    }
    return 0;
}
int pupil_image_save(const char *path, const float *rgba, uint32_t width, uint32_t height, int format) {
    if (!path || format < 0 || format > 4) return Fail("pupil_image_save: bad arguments");
    const util::DisplayTransform display{ format == 4, true }; // 3: png as the canvas shows it by default, 4: with ACES tone mapping
    return util::SaveImage(rgba, width, height, path, static_cast<util::EImageFileFormat>(format > 3 ? 3 : format), display) ? 0 : Fail("image saving failed");
}
int pupil_save_buffer(const char *name, const char *path, int format) {
    if (!Ready() || !name || !path || format < 0 || format > 4) return Fail("pupil_save_buffer: bad arguments / no scene");
    Buffer *b = util::Singleton<BufferManager>::instance()->GetBuffer(name);
    if (!b || !b->cuda_ptr || b->desc.stride_in_byte != 16) return Fail("pupil_save_buffer: no such float4 buffer");
    SyncIfSharded();
    std::vector<float> host(static_cast<size_t>(b->desc.width) * b->desc.height * 4);
    if (pb2_download(host.data(), b->cuda_ptr, host.size() * sizeof(float)) != PB2_OK) return Fail(pb2_last_error());
    const util::DisplayTransform display{ format == 4, true };
    return util::SaveImage(host.data(), b->desc.width, b->desc.height, path, static_cast<util::EImageFileFormat>(format > 3 ? 3 : format), display) ? 0
                                                                                                                                                   : Fail("image saving failed");
}
int pupil_get_env_tables(uint32_t *map_w, uint32_t *map_h, float *row_cdf, float *row_weight, float *col_cdf) {
    if (!HasWorld()) return Fail("no scene");
    const pb2_emitter *e = W()->emitters->GetEnvEmitter();
    if (!e || e->type != PB2_EMIT_ENV_MAP) return Fail("the scene has no env-map emitter");
    if (map_w) *map_w = e->map_w;
    if (map_h) *map_h = e->map_h;
    auto copy = [](float *dst, const std::vector<float> &src) {
        if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(float));
    };
    copy(row_cdf, W()->emitters->GetEnvRowCdf()), copy(row_weight, W()->emitters->GetEnvRowWeight()), copy(col_cdf, W()->emitters->GetEnvColCdf());
    return 0;
}
int pupil_set_instance_transform(uint32_t index, const float xform[16]) {
    if (!HasWorld() || !xform) return Fail("no scene");
    world::RenderObject *ro = W()->GetRenderObject(static_cast<size_t>(index));
    if (!ro) return Fail("no such render object");
    util::Transform t;
    std::memcpy(t.matrix.e, xform, sizeof(float) * 16);
    ro->UpdateTransform(t); // EWorldEvent::RenderInstanceTransform -> emitters reset, acceleration structure dirty, pass restarts
    return 0;
}
int pupil_remove_instance(uint32_t index) {
    if (!HasWorld()) return Fail("no scene");
    if (!W()->GetRenderObject(static_cast<size_t>(index))) return Fail("no such render object");
    W()->RemoveRenderObject(static_cast<size_t>(index)); // World::RemoveRenderObject: its emitters leave the table with it
    EventDispatcher<EWorldEvent::RenderInstanceUpdate>(nullptr); // the pass restarts its accumulation
    return 0;
}
int pupil_clear_shapes(void) {
    util::Singleton<resource::ShapeManager>::instance()->Clear();
    return 0;
}
int pupil_pass_config(int max_depth, int accumulate, uint32_t frames_per_run, uint32_t first_seed, uint32_t seed_stride, int sum_mode) {
    if (!Ready()) return Fail("pupil_init first");
    if (max_depth > 0) g_pass->SetMaxDepth(max_depth);
    g_pass->SetAccumulate(accumulate != 0);
    g_pass->SetFramesPerRun(frames_per_run);
    g_pass->SetSumMode(sum_mode != 0);
    g_pass->Restart(first_seed, seed_stride);
    return 0;
}
int pupil_comm_unique_id(uint8_t id[PB2_COMM_ID_BYTES]) { return pb2_comm_unique_id(id) == PB2_OK ? 0 : Fail(pb2_last_error()); }
int pupil_set_shard(int rank, int world, const uint8_t *id, int strong, int reduce_mode) {
    if (!Ready()) return Fail("pupil_init first");
    g_pass->SetShard(nullptr, 0, 1, false, PB2_REDUCE_ALL);
    if (g_comm) pb2_comm_destroy(g_comm), g_comm = nullptr;
    if (world <= 0) return 0; // sharding off
    if (pb2_comm_create(&g_comm, world, rank, id) != PB2_OK) return Fail(pb2_last_error());
    g_rank = rank, g_world = world, g_reduce_mode = reduce_mode;
    g_pass->SetShard(g_comm, rank, world, strong != 0, reduce_mode);
    return 0;
}
int pupil_set_shard_plan(int strong) {
    if (!Ready() || !g_comm) return Fail("pupil_set_shard first");
    g_pass->SetShard(g_comm, g_rank, g_world, strong != 0, g_reduce_mode); // same communicator, other split of the step's seeds
    return 0;
}
int pupil_last_reduction(float *ms, uint64_t *bytes) {
    if (!Ready() || !g_comm) return Fail("pupil_set_shard first");
    return pb2_comm_last_reduction(g_comm, ms, bytes) == PB2_OK ? 0 : Fail(pb2_last_error());
}
int pupil_synchronize(void) {
    if (!Ready()) return Fail("pupil_init first");
    g_pass->Synchronize();
    return 0;
}
int pupil_run(uint64_t n) {
    if (!Ready()) return Fail("pupil_init first");
    if (!n) return 0;
    Sys()->max_frames = n;
    Sys()->Run();
    return 0;
}
int pupil_pass_state(uint32_t *sample_cnt, uint32_t *random_seed) {
    if (!Ready()) return Fail("pupil_init first");
    if (sample_cnt) *sample_cnt = g_pass->GetLaunchParams().sample_cnt;
    if (random_seed) *random_seed = g_pass->GetLaunchParams().random_seed;
    return 0;
}
int pupil_checkpoint_save(const char *path) {
    if (!Ready() || !path) return Fail("pupil_init first");
    return g_pass->SaveCheckpoint(path) ? 0 : Fail(std::string("cannot save checkpoint: ") + path);
}
int pupil_checkpoint_load(const char *path) {
    if (!Ready() || !path) return Fail("pupil_init first");
    return g_pass->LoadCheckpoint(path) ? 0 : Fail(std::string("cannot load checkpoint: ") + path);
}
int pupil_buffer_info(const char *name, void **dptr, uint32_t *w, uint32_t *h, uint32_t *stride) {
    Buffer *b = name ? util::Singleton<BufferManager>::instance()->GetBuffer(name) : nullptr;
    if (!b) return Fail(std::string("no buffer named ") + (name ? name : "(null)"));
    if (dptr) *dptr = b->cuda_ptr;
    if (w) *w = b->desc.width;
    if (h) *h = b->desc.height;
    if (stride) *stride = b->desc.stride_in_byte;
    return 0;
}
int pupil_buffer_download(const char *name, void *host, uint64_t bytes) {
    SyncIfSharded();
    Buffer *b = name ? util::Singleton<BufferManager>::instance()->GetBuffer(name) : nullptr;
    if (!b) return Fail(std::string("no buffer named ") + (name ? name : "(null)"));
    if (bytes > b->SizeInBytes()) return Fail("buffer is smaller than the request");
    if (pb2_download(host, b->cuda_ptr, bytes) != PB2_OK) return Fail(pb2_last_error());
    return 0;
}
int pupil_buffer_upload(const char *name, const void *host, uint64_t bytes) {
    Buffer *b = name ? util::Singleton<BufferManager>::instance()->GetBuffer(name) : nullptr;
    if (!b) return Fail(std::string("no buffer named ") + (name ? name : "(null)"));
    if (bytes > b->SizeInBytes()) return Fail("buffer is smaller than the request");
    if (pb2_upload(b->cuda_ptr, host, bytes) != PB2_OK) return Fail(pb2_last_error());
    return 0;
}
int pupil_get_film(uint32_t *w, uint32_t *h, uint32_t *max_depth) {
    if (!HasWorld()) return Fail("no scene");
    if (w) *w = W()->scene->sensor.film.w;
    if (h) *h = W()->scene->sensor.film.h;
    if (max_depth) *max_depth = W()->scene->integrator.max_depth;
    return 0;
}
int pupil_get_camera(float s2c[16], float c2w[16], float *fov_y) {
    if (!HasWorld()) return Fail("no scene");
    const util::Mat4 a = W()->camera->GetSampleToCameraMatrix(), b = W()->camera->GetToWorldMatrix();
    if (s2c) std::memcpy(s2c, a.e, 64);
    if (c2w) std::memcpy(c2w, b.e, 64);
    if (fov_y) *fov_y = W()->camera->GetDesc().fov_y;
    return 0;
}
int pupil_num_instances(void) { return HasWorld() ? (int)W()->GetRenderobjects().size() : 0; }
int pupil_get_instance(uint32_t index, float xform[16], pb2_material *material, int32_t *emitter_offset, uint32_t *flags, uint32_t *n_prims,
                       int32_t *is_sphere) {
    if (!HasWorld()) return Fail("no scene");
    auto ros = W()->GetRenderobjects();
    if (index >= ros.size()) return Fail("instance index out of range");
    const world::RenderObject *ro = ros[index];
    const int mine = W()->GetEmitterOffset(ro);
    if (xform) std::memcpy(xform, ro->transform.matrix.e, 64);
    if (material) *material = ro->mat;
    if (emitter_offset) *emitter_offset = mine;
    if (flags) *flags = (ro->flip_normals ? PB2_INST_FLIP_NORMALS : 0u) | (ro->flip_tex_coords ? PB2_INST_FLIP_TEX : 0u);
    if (n_prims) *n_prims = ro->sub_emitters_num;
    if (is_sphere) *is_sphere = ro->geo_type == world::RenderObject::EGeoType::Sphere;
    return 0;
}
int pupil_num_area_emitters(void) { return HasWorld() ? (int)W()->emitters->GetAreaEmitters().size() : 0; }
int pupil_get_emitters(pb2_emitter *areas, pb2_emitter *env, int32_t *has_env) {
    if (!HasWorld()) return Fail("no scene");
    auto &a = W()->emitters->GetAreaEmitters();
    if (areas && !a.empty()) std::memcpy(areas, a.data(), a.size() * sizeof(pb2_emitter));
    const pb2_emitter *e = W()->emitters->GetEnvEmitter();
    if (has_env) *has_env = e != nullptr;
    if (env && e) *env = *e;
    return 0;
}
int pupil_scene_handle(pb2_scene **scene) {
    if (!Ready() || !scene) return Fail("no scene");
    *scene = W()->GetSceneHandle();
    return *scene ? 0 : Fail("no device scene");
}
int pupil_set_bvh_builder(int builder) {
    if (!Ready()) return Fail("pupil_init first");
    W()->SetBvhBuilder(builder);
    return 0;
}
int pupil_set_instancing(int mode) {
    if (!Ready()) return Fail("pupil_init first");
    W()->SetInstancing(mode);
    return 0;
}
int pupil_build_stats(pb2_build_stats *stats) {
    if (!Ready() || !stats) return Fail("no scene");
    W()->GetSceneHandle();
    *stats = W()->GetBuildStats();
    return 0;
}
int pupil_render_stats(pb2_render_stats *stats) {
    if (!Ready() || !stats) return Fail("no scene");
    *stats = g_pass->GetRenderStats();
    return 0;
}
int pupil_camera_move(float dx, float dy, float dz) {
    if (!Ready() || !W()->camera) return Fail("no scene");
    W()->camera->Move(util::Float3{ dx, dy, dz });
    EventDispatcher<EWorldEvent::CameraChange>();
    return 0;
}
int pupil_camera_rotate(float delta_x, float delta_y) {
    if (!Ready() || !W()->camera) return Fail("no scene");
    W()->camera->Rotate(delta_x, delta_y);
    EventDispatcher<EWorldEvent::CameraChange>();
    return 0;
}
int pupil_camera_set_fov(float fov_y) {
    if (!Ready() || !W()->camera) return Fail("no scene");
    W()->camera->SetFov(fov_y);
    EventDispatcher<EWorldEvent::CameraChange>();
    return 0;
}
}
