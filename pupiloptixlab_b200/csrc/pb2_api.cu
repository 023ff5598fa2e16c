// extern "C" surface of libpb2.so (include/pb2.h): argument checking, exception -> status code mapping,
// host <-> device conversion of the POD records.  No compute lives here.
#include "scene.cuh"
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>

namespace pb2 {
void collect_render_stats(Scene &s); // wavefront.cu

DevTexture to_dev(const pb2_texture &t) {
    DevTexture d;
    float type_bits;
    memcpy(&type_bits, &t.type, 4);
    d.hdr = make_float4(type_bits, t.a[0], t.a[1], t.a[2]);
    d.b = make_float4(t.b[0], t.b[1], t.b[2], 0.f);
    if (t.type == PB2_TEX_BITMAP) { // hdr.y / hdr.z carry the 64-bit texture object
        const uint32_t lo = (uint32_t)(t.bitmap & 0xffffffffull), hi = (uint32_t)(t.bitmap >> 32);
        memcpy(&d.hdr.y, &lo, 4), memcpy(&d.hdr.z, &hi, 4);
        d.hdr.w = 0.f;
    }
    d.r0 = make_float4(t.r0[0], t.r0[1], t.r0[2], t.r0[3]);
    d.r1 = make_float4(t.r1[0], t.r1[1], t.r1[2], t.r1[3]);
    return d;
}
static DevMaterial to_dev(const pb2_material &m) {
    DevMaterial d{};
    d.type = m.type, d.twosided = m.twosided, d.eta = m.eta, d.nonlinear = m.nonlinear;
    d.int_fdr = m.int_fdr, d.specular_sampling_weight = m.specular_sampling_weight;
    for (int i = 0; i < 4; ++i) d.tex[i] = to_dev(m.tex[i]);
    return d;
}
DevEmitter to_dev(const pb2_emitter &e) {
    DevEmitter d{};
    d.type = e.type, d.weight = e.weight, d.select_probability = e.select_probability, d.area = e.area;
    d.radiance = to_dev(e.radiance);
    d.p0 = make_float4(e.pos[0][0], e.pos[0][1], e.pos[0][2], e.uv[0][0]);
    d.p1 = make_float4(e.pos[1][0], e.pos[1][1], e.pos[1][2], e.uv[0][1]);
    d.p2 = make_float4(e.pos[2][0], e.pos[2][1], e.pos[2][2], e.uv[1][0]);
    d.n0 = make_float4(e.nrm[0][0], e.nrm[0][1], e.nrm[0][2], e.uv[1][1]);
    d.n1 = make_float4(e.nrm[1][0], e.nrm[1][1], e.nrm[1][2], e.uv[2][0]);
    d.n2 = make_float4(e.nrm[2][0], e.nrm[2][1], e.nrm[2][2], e.uv[2][1]);
    d.center_r = make_float4(e.center[0], e.center[1], e.center[2], e.radius);
    if (e.type == PB2_EMIT_ENV_MAP) { // env.h:6-22: p* = to_world rows, n* = to_local rows, the spare w lanes carry the rest
        float wb, hb, plo, phi;
        const uint64_t tp = reinterpret_cast<uint64_t>(e.env_tables);
        const uint32_t lo = (uint32_t)(tp & 0xffffffffull), hi = (uint32_t)(tp >> 32);
        memcpy(&wb, &e.map_w, 4), memcpy(&hb, &e.map_h, 4), memcpy(&plo, &lo, 4), memcpy(&phi, &hi, 4);
        d.area = e.normalization;
        d.p0 = make_float4(e.to_world[0], e.to_world[1], e.to_world[2], e.scale);
        d.p1 = make_float4(e.to_world[3], e.to_world[4], e.to_world[5], wb);
        d.p2 = make_float4(e.to_world[6], e.to_world[7], e.to_world[8], hb);
        d.n0 = make_float4(e.to_local[0], e.to_local[1], e.to_local[2], plo);
        d.n1 = make_float4(e.to_local[3], e.to_local[4], e.to_local[5], phi);
        d.n2 = make_float4(e.to_local[6], e.to_local[7], e.to_local[8], 0.f);
    }
    return d;
}

// EmitterGroup::SelectOneEmiiter (framework/render/emitter.h:110-136) walks the table with `sum_p += select_probability`
// and stops at the first entry with p <= sum_p + select_probability.  The same fp32 running sums, computed once here in
// the same order, let the device find that entry by binary search: identical selection, O(log n) instead of O(n) per bounce.
std::vector<float> area_select_cdf(const DevEmitter *areas, size_t n) {
    std::vector<float> cdf(n);
    float sum_p = 0.f;
    for (size_t i = 0; i < n; ++i) {
        sum_p = sum_p + areas[i].select_probability;
        cdf[i] = sum_p;
    }
    return cdf;
}

Scene::Scene() {
    PB2_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    stream = own_stream;
}
Scene::~Scene() {
    if (wf) wavefront_destroy(wf);
    if (accumulate_gate) cudaEventDestroy(accumulate_gate);
    if (own_stream) cudaStreamDestroy(own_stream);
}
void Scene::apply_l2_window() {
    l2_dirty = false;
    cudaStreamAttrValue v{};
    int dev = 0, max_persist = 0, max_window = 0;
    PB2_CUDA(cudaGetDevice(&dev));
    PB2_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    PB2_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    const size_t persist = std::min((size_t)l2_persist_mb << 20, (size_t)std::max(0, max_persist));
    if (persist > 0 && bvh_valid && n_nodes > 0) {
        const size_t want = l2_window_mb > l2_persist_mb ? (size_t)l2_window_mb << 20 : persist;
        const size_t bytes = std::min(std::min((size_t)n_nodes * sizeof(Bvh8Node), (size_t)max_window), want);
        PB2_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
        v.accessPolicyWindow.base_ptr = d_nodes.ptr;
        v.accessPolicyWindow.num_bytes = bytes;
        v.accessPolicyWindow.hitRatio = bytes > persist ? (float)persist / (float)bytes : 1.f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } // else: a zero-byte window switches the policy off
    PB2_CUDA(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v));
    if (v.accessPolicyWindow.num_bytes == 0) cudaCtxResetPersistingL2Cache();
}
void Scene::upload_tables() {
    if (!tables_dirty) return;
    uint32_t seen = 0;
    for (const DevInstance &in : h_inst) seen |= 1u << (in.mat_type & 7);
    n_material_types = __builtin_popcount(seen);
    only_material_type = n_material_types == 1 ? __builtin_ctz(seen) : -1;
    material_type_mask = seen;
    // Texture coordinates are fetched and interpolated per shaded vertex only where they can change the result: an instance
    // whose material textures (and, for an emitter, radiance textures) are all constant colours is uploaded without
    // PB2_IF_HAS_UV (the host copy keeps the flag).
    std::vector<DevInstance> dev_inst = h_inst;
    auto constant = [](const DevTexture &t) { return __builtin_bit_cast(int32_t, t.hdr.x) == PB2_TEX_RGB; };
    for (size_t i = 0; i < dev_inst.size(); ++i) {
        DevInstance &in = dev_inst[i];
        if (!(in.flags & PB2_IF_HAS_UV) || i >= h_mat.size()) continue;
        bool uv_matters = false;
        for (const DevTexture &t : h_mat[i].tex) uv_matters |= !constant(t);
        if (in.emitter_offset >= 0)
            for (uint64_t e = (uint64_t)in.emitter_offset, end = e + in.n_tris; e < end; ++e) uv_matters |= e >= h_areas.size() || !constant(h_areas[e].radiance);
        if (!uv_matters) in.flags &= ~PB2_IF_HAS_UV;
    }
    d_inst.upload(dev_inst.data(), dev_inst.size(), stream);
    d_mat.upload(h_mat.data(), h_mat.size(), stream);
    d_areas.upload(h_areas.data(), h_areas.size(), stream);
    std::vector<float> cdf = area_select_cdf(h_areas.data(), h_areas.size());
    d_area_cdf.upload(cdf.data(), cdf.size(), stream);
    if (has_env) d_env.upload(&h_env, 1, stream);
    PB2_CUDA(cudaStreamSynchronize(stream)); // host vectors may change right after
    tables_dirty = false;
}
void Scene::check_emitter_ranges() const {
    for (size_t i = 0; i < h_inst.size(); ++i) {
        const DevInstance &in = h_inst[i];
        if (in.emitter_offset >= 0 && (uint64_t)in.emitter_offset + in.n_tris > h_areas.size())
            throw std::runtime_error("pb2_render: instance " + std::to_string(i) + " names emitter entries [" + std::to_string(in.emitter_offset) + ", " +
                                     std::to_string((uint64_t)in.emitter_offset + in.n_tris) + ") but the emitter table holds " + std::to_string(h_areas.size()) +
                                     " (pb2_scene_set_emitters)");
    }
}
SceneView Scene::view() const {
    SceneView v{};
    v.nodes = d_nodes.ptr, v.prims = d_prims.ptr, v.instances = d_inst.ptr, v.materials = d_mat.ptr;
    v.areas = d_areas.ptr, v.env = has_env ? d_env.ptr : nullptr, v.area_cdf = d_area_cdf.ptr;
    v.n_areas = (uint32_t)h_areas.size(), v.n_nodes = n_nodes, v.n_prims = n_prims, v.root = root;
    v.plane_bias = 0x47000000u;
    return v;
}
}// namespace pb2


// ---- device memory pool (util.cuh) ---------------------------------------------------------------------------
namespace pb2 {
namespace {
struct Pool {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;  // cached blocks by size
    std::unordered_map<void *, size_t> owned;   // every block handed out or cached -> its size
    size_t cached = 0;
    size_t cap = 48ull << 30;                   // cached bytes kept per device; beyond it blocks go back to the driver
};
Pool &pool_of_current_device() {
    static std::mutex mu;
    static std::map<int, std::unique_ptr<Pool>> pools;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    auto &p = pools[dev];
    if (!p) p = std::make_unique<Pool>();
    return *p;
}
size_t pool_round(size_t bytes) { return bytes <= (1u << 20) ? (bytes + 511) & ~size_t(511) : (bytes + (2u << 20) - 1) & ~size_t((2u << 20) - 1); }
void pool_trim_locked(Pool &p) {
    for (auto &kv : p.free_blocks) {
        cudaFree(kv.second);
        p.owned.erase(kv.second);
    }
    p.free_blocks.clear();
    p.cached = 0;
}
}// namespace
void *pool_alloc(size_t bytes) {
    if (!bytes) return nullptr;
    Pool &p = pool_of_current_device();
    const size_t want = pool_round(bytes);
    std::lock_guard<std::mutex> g(p.mu);
    auto it = p.free_blocks.lower_bound(want);
    if (it != p.free_blocks.end() && it->first <= want + want / 4 + (64u << 10)) { // at most 25 % (+64 KiB) of slack
        void *ptr = it->second;
        p.cached -= it->first;
        p.free_blocks.erase(it);
        return ptr;
    }
    void *ptr = nullptr;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e == cudaErrorMemoryAllocation) { // give the cache back and try once more
        cudaGetLastError();
        pool_trim_locked(p);
        e = cudaMalloc(&ptr, want);
    }
    PB2_CUDA(e);
    p.owned[ptr] = want;
    return ptr;
}
void pool_free(void *ptr) noexcept {
    if (!ptr) return;
    cudaDeviceSynchronize(); // what cudaFree guarantees: nothing in flight still touches the block when it changes owner
    Pool &p = pool_of_current_device();
    std::lock_guard<std::mutex> g(p.mu);
    auto it = p.owned.find(ptr);
    if (it == p.owned.end()) { // not ours (allocated before a device switch): hand it straight back
        cudaFree(ptr);
        return;
    }
    if (p.cached + it->second > p.cap) {
        cudaFree(ptr);
        p.owned.erase(it);
        return;
    }
    p.free_blocks.emplace(it->second, ptr);
    p.cached += it->second;
}
void pool_trim() noexcept {
    cudaDeviceSynchronize();
    Pool &p = pool_of_current_device();
    std::lock_guard<std::mutex> g(p.mu);
    pool_trim_locked(p);
}
}// namespace pb2

using namespace pb2;

static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
#define PB2_TRY try {
#define PB2_CATCH                                                  \
    }                                                              \
    catch (const CudaError &e) { return fail(PB2_ERR_CUDA, e.what()); } \
    catch (const std::exception &e) { return fail(PB2_ERR_STATE, e.what()); }

static Scene *S(pb2_scene *s) { return reinterpret_cast<Scene *>(s); }

// 3x4 affine inverse in fp64, rounded once (the reference lets OptiX / DirectXMath invert instance transforms)
static bool invert_affine(const float m[12], float out[12]) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0) return false;
    const double id = 1.0 / det;
    const double r[9] = { (e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id, (a * i - c * g) * id,
                          (c * d - a * f) * id, (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id };
    const double tx = m[3], ty = m[7], tz = m[11];
    for (int k = 0; k < 3; ++k) {
        out[k * 4 + 0] = (float)r[k * 3], out[k * 4 + 1] = (float)r[k * 3 + 1], out[k * 4 + 2] = (float)r[k * 3 + 2];
        out[k * 4 + 3] = (float)-(r[k * 3] * tx + r[k * 3 + 1] * ty + r[k * 3 + 2] * tz);
    }
    return true;
}

extern "C" {
const char *pb2_last_error(void) { return g_last_error.c_str(); }

int pb2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int pb2_init(int device) {
    PB2_TRY
    int n = 0;
    PB2_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(PB2_ERR_ARG, "pb2_init: no such CUDA device");
    PB2_CUDA(cudaSetDevice(device));
    PB2_CUDA(cudaFree(nullptr));
    return PB2_OK;
    PB2_CATCH
}
int pb2_malloc(void **dptr, uint64_t bytes) {
    PB2_TRY
    if (!dptr) return fail(PB2_ERR_ARG, "pb2_malloc: null");
    *dptr = nullptr;
    if (!bytes) return PB2_OK;
    *dptr = pool_alloc(bytes);
    PB2_CUDA(cudaMemset(*dptr, 0, bytes));
    return PB2_OK;
    PB2_CATCH
}
int pb2_free(void *dptr) {
    PB2_TRY
    pool_free(dptr);
    return PB2_OK;
    PB2_CATCH
}
int pb2_trim(void) {
    PB2_TRY
    pool_trim();
    return PB2_OK;
    PB2_CATCH
}
int pb2_upload(void *dptr, const void *host, uint64_t bytes) {
    PB2_TRY
    if (bytes) PB2_CUDA(cudaMemcpy(dptr, host, bytes, cudaMemcpyHostToDevice));
    return PB2_OK;
    PB2_CATCH
}
int pb2_download(void *host, const void *dptr, uint64_t bytes) {
    PB2_TRY
    if (bytes) PB2_CUDA(cudaMemcpy(host, dptr, bytes, cudaMemcpyDeviceToHost));
    return PB2_OK;
    PB2_CATCH
}
int pb2_memset(void *dptr, int value, uint64_t bytes) {
    PB2_TRY
    if (bytes) PB2_CUDA(cudaMemset(dptr, value, bytes));
    return PB2_OK;
    PB2_CATCH
}

// CudaTextureManager::GetCudaTextureObject, framework/cuda/texture.cpp:60-102
namespace {
std::mutex g_bitmap_mu;
std::unordered_map<uint64_t, cudaArray_t> g_bitmaps; // texture object -> its array
}
int pb2_bitmap_create(const float *rgba, uint32_t width, uint32_t height, int address_mode, int filter_mode, uint64_t *handle) {
    PB2_TRY
    if (!rgba || !width || !height || !handle) return fail(PB2_ERR_ARG, "pb2_bitmap_create: null / empty image");
    if (address_mode < 0 || address_mode > 3 || filter_mode < 0 || filter_mode > 1) return fail(PB2_ERR_ARG, "pb2_bitmap_create: bad address / filter mode");
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
    cudaArray_t arr = nullptr;
    PB2_CUDA(cudaMallocArray(&arr, &desc, width, height));
    const size_t pitch = (size_t)width * 4 * sizeof(float);
    PB2_CUDA(cudaMemcpy2DToArray(arr, 0, 0, rgba, pitch, pitch, height, cudaMemcpyHostToDevice));
    cudaResourceDesc res{};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = (cudaTextureAddressMode)address_mode;
    td.filterMode = (cudaTextureFilterMode)filter_mode;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    td.maxAnisotropy = 1;
    td.maxMipmapLevelClamp = 99, td.minMipmapLevelClamp = 0;
    td.mipmapFilterMode = cudaFilterModePoint;
    td.borderColor[0] = 1.0f;
    td.sRGB = 0;
    cudaTextureObject_t tex = 0;
    cudaError_t e = cudaCreateTextureObject(&tex, &res, &td, nullptr);
    if (e != cudaSuccess) {
        cudaFreeArray(arr);
        PB2_CUDA(e);
    }
    std::lock_guard<std::mutex> g(g_bitmap_mu);
    g_bitmaps[(uint64_t)tex] = arr;
    *handle = (uint64_t)tex;
    return PB2_OK;
    PB2_CATCH
}
int pb2_bitmap_destroy(uint64_t handle) {
    PB2_TRY
    std::lock_guard<std::mutex> g(g_bitmap_mu);
    auto it = g_bitmaps.find(handle);
    if (it == g_bitmaps.end()) return fail(PB2_ERR_ARG, "pb2_bitmap_destroy: unknown handle");
    cudaDeviceSynchronize();
    cudaDestroyTextureObject((cudaTextureObject_t)handle);
    cudaFreeArray(it->second);
    g_bitmaps.erase(it);
    return PB2_OK;
    PB2_CATCH
}

int pb2_scene_create(pb2_scene **scene) {
    PB2_TRY
    if (!scene) return fail(PB2_ERR_ARG, "pb2_scene_create: null");
    *scene = reinterpret_cast<pb2_scene *>(new Scene());
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_destroy(pb2_scene *scene) {
    PB2_TRY
    if (scene) {
        cudaStreamSynchronize(S(scene)->stream);
        delete S(scene);
    }
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_clear(pb2_scene *scene) {
    PB2_TRY
    Scene &s = *S(scene);
    PB2_CUDA(cudaStreamSynchronize(s.stream));
    s.meshes.clear(), s.h_inst.clear(), s.h_inst_mesh.clear(), s.h_mat.clear(), s.h_areas.clear();
    s.has_env = false, s.tables_dirty = true, s.bvh_valid = false, s.blas_valid = false, s.n_blas = 0, s.root = 0, s.n_nodes = s.n_prims = 0;
    s.d_nodes.release(), s.d_prims.release();
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_set_stream(pb2_scene *scene, void *cuda_stream) {
    PB2_TRY
    Scene &s = *S(scene);
    PB2_CUDA(cudaStreamSynchronize(s.stream));
    s.stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s.own_stream;
    if (s.l2_persist_mb > 0) s.l2_dirty = true;
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_add_mesh(pb2_scene *scene, const float *pos, const float *nrm, const float *uv, const uint32_t *idx, uint32_t n_vertices,
                       uint32_t n_triangles, uint32_t *mesh_id) {
    PB2_TRY
    if (!scene || !pos || !idx || !n_vertices || !n_triangles) return fail(PB2_ERR_ARG, "pb2_scene_add_mesh: empty mesh");
    Scene &s = *S(scene);
    auto m = std::make_unique<Mesh>();
    m->n_verts = n_vertices, m->n_tris = n_triangles;
    m->pos.upload(pos, (size_t)n_vertices * 3, s.stream);
    if (nrm) m->nrm.upload(nrm, (size_t)n_vertices * 3, s.stream);
    if (uv) m->uv.upload(uv, (size_t)n_vertices * 2, s.stream);
    m->idx.upload(idx, (size_t)n_triangles * 3, s.stream);
    // an index past the vertex arrays would be an out-of-bounds device read in every kernel that touches the mesh
    const uint32_t max_idx = max_index_dev(m->idx.ptr, (uint64_t)n_triangles * 3, s.stream); // synchronises: caller may free its arrays now
    if (max_idx >= n_vertices) return fail(PB2_ERR_ARG, "pb2_scene_add_mesh: vertex index " + std::to_string(max_idx) + " >= n_vertices " + std::to_string(n_vertices));
    if (mesh_id) *mesh_id = (uint32_t)s.meshes.size();
    s.meshes.push_back(std::move(m));
    s.blas_valid = false;
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_add_instance(pb2_scene *scene, uint32_t mesh_id, const float xform[12], uint32_t flags, const pb2_material *material,
                           int32_t emitter_index_offset, uint32_t *instance_id) {
    PB2_TRY
    if (!scene || !xform) return fail(PB2_ERR_ARG, "pb2_scene_add_instance: null");
    Scene &s = *S(scene);
    DevInstance in{};
    float inv[12];
    if (!invert_affine(xform, inv)) return fail(PB2_ERR_ARG, "pb2_scene_add_instance: singular transform");
    for (int r = 0; r < 3; ++r) {
        in.xf[r] = make_float4(xform[r * 4], xform[r * 4 + 1], xform[r * 4 + 2], xform[r * 4 + 3]);
        in.inv[r] = make_float4(inv[r * 4], inv[r * 4 + 1], inv[r * 4 + 2], inv[r * 4 + 3]);
    }
    in.flags = flags & (PB2_INST_FLIP_NORMALS | PB2_INST_FLIP_TEX);
    if (mesh_id == PB2_MESH_SPHERE) {
        in.flags |= PB2_IF_SPHERE;
        in.n_tris = 1;
    } else {
        if (mesh_id >= s.meshes.size()) return fail(PB2_ERR_ARG, "pb2_scene_add_instance: unknown mesh id");
        const Mesh &m = *s.meshes[mesh_id];
        in.pos = m.pos.ptr, in.nrm = m.nrm.ptr, in.uv = m.uv.ptr, in.idx = m.idx.ptr;
        if (m.nrm.ptr) in.flags |= PB2_IF_HAS_NRM;
        if (m.uv.ptr) in.flags |= PB2_IF_HAS_UV;
        in.n_tris = m.n_tris;
    }
    pb2_material none{};
    const pb2_material &mat = material ? *material : none;
    if (mat.twosided) in.flags |= PB2_IF_TWOSIDED;
    in.mat_type = mat.type;
    in.emitter_offset = emitter_index_offset;
    if (instance_id) *instance_id = (uint32_t)s.h_inst.size();
    s.h_inst.push_back(in);
    s.h_inst_mesh.push_back(mesh_id == PB2_MESH_SPHERE ? -1 : (int)mesh_id);
    s.h_mat.push_back(to_dev(mat));
    s.tables_dirty = true, s.bvh_valid = false, s.blas_valid = false; // how often a mesh is placed decides whether it gets a tree of its own
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_set_instance_transform(pb2_scene *scene, uint32_t instance_id, const float xform[12]) {
    PB2_TRY
    if (!scene || !xform) return fail(PB2_ERR_ARG, "pb2_scene_set_instance_transform: null");
    Scene &s = *S(scene);
    if (instance_id >= s.h_inst.size()) return fail(PB2_ERR_ARG, "pb2_scene_set_instance_transform: no such instance");
    float inv[12];
    if (!invert_affine(xform, inv)) return fail(PB2_ERR_ARG, "pb2_scene_set_instance_transform: singular transform");
    DevInstance &in = s.h_inst[instance_id];
    for (int r = 0; r < 3; ++r) {
        in.xf[r] = make_float4(xform[r * 4], xform[r * 4 + 1], xform[r * 4 + 2], xform[r * 4 + 3]);
        in.inv[r] = make_float4(inv[r * 4], inv[r * 4 + 1], inv[r * 4 + 2], inv[r * 4 + 3]);
    }
    s.tables_dirty = true, s.bvh_valid = false; // blas_valid stays: pb2_bvh_build rebuilds the top level only
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_set_emitters(pb2_scene *scene, const pb2_emitter *areas, uint32_t n_areas, const pb2_emitter *env) {
    PB2_TRY
    if (!scene || (n_areas && !areas)) return fail(PB2_ERR_ARG, "pb2_scene_set_emitters: null");
    Scene &s = *S(scene);
    s.h_areas.resize(n_areas);
    for (uint32_t i = 0; i < n_areas; ++i) s.h_areas[i] = to_dev(areas[i]);
    s.has_env = env != nullptr;
    if (env) s.h_env = to_dev(*env);
    s.tables_dirty = true;
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_set_camera(pb2_scene *scene, const float s2c[16], const float c2w[16]) {
    PB2_TRY
    if (!scene || !s2c || !c2w) return fail(PB2_ERR_ARG, "pb2_scene_set_camera: null");
    Scene &s = *S(scene);
    for (int r = 0; r < 4; ++r) {
        s.cam.s2c[r] = make_float4(s2c[r * 4], s2c[r * 4 + 1], s2c[r * 4 + 2], s2c[r * 4 + 3]);
        s.cam.c2w[r] = make_float4(c2w[r * 4], c2w[r * 4 + 1], c2w[r * 4 + 2], c2w[r * 4 + 3]);
    }
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_set_builder(pb2_scene *scene, int builder) {
    if (!scene || builder < 0 || builder > 2) return fail(PB2_ERR_ARG, "pb2_scene_set_builder: 0 = LBVH, 1 = binned SAH along the Morton order, 2 = SAH-driven clustering");
    S(scene)->builder = builder;
    S(scene)->bvh_valid = false;
    return PB2_OK;
}
int pb2_bvh_build(pb2_scene *scene, pb2_build_stats *stats) {
    PB2_TRY
    if (!scene) return fail(PB2_ERR_ARG, "pb2_bvh_build: null");
    build_bvh(*S(scene));
    if (stats) *stats = S(scene)->build_stats;
    return PB2_OK;
    PB2_CATCH
}
int pb2_trace_closest_dev(pb2_scene *scene, const void *rays_dev, uint64_t n, void *hits_tuvp_dev, void *hits_inst_dev) {
    PB2_TRY
    trace_closest_dev(*S(scene), static_cast<const float4 *>(rays_dev), n, static_cast<float4 *>(hits_tuvp_dev), static_cast<int32_t *>(hits_inst_dev));
    return PB2_OK;
    PB2_CATCH
}
int pb2_trace_any_dev(pb2_scene *scene, const void *rays_dev, uint64_t n, void *occluded_u32_dev) {
    PB2_TRY
    trace_any_dev(*S(scene), static_cast<const float4 *>(rays_dev), n, static_cast<uint32_t *>(occluded_u32_dev));
    return PB2_OK;
    PB2_CATCH
}
int pb2_trace_closest(pb2_scene *scene, const float *rays, uint64_t n, pb2_hit *hits) {
    PB2_TRY
    if (!scene || (n && (!rays || !hits))) return fail(PB2_ERR_ARG, "pb2_trace_closest: null");
    Scene &s = *S(scene);
    if (!n) return PB2_OK;
    DevBuf<float4> d_rays(n * 2), d_tuvp(n);
    DevBuf<int32_t> d_inst(n);
    PB2_CUDA(cudaMemcpyAsync(d_rays.ptr, rays, n * 32, cudaMemcpyHostToDevice, s.stream));
    trace_closest_dev(s, d_rays.ptr, n, d_tuvp.ptr, d_inst.ptr);
    std::vector<float4> tuvp(n);
    std::vector<int32_t> inst(n);
    PB2_CUDA(cudaMemcpyAsync(tuvp.data(), d_tuvp.ptr, n * 16, cudaMemcpyDeviceToHost, s.stream));
    PB2_CUDA(cudaMemcpyAsync(inst.data(), d_inst.ptr, n * 4, cudaMemcpyDeviceToHost, s.stream));
    PB2_CUDA(cudaStreamSynchronize(s.stream));
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t prim;
        memcpy(&prim, &tuvp[i].w, 4);
        hits[i] = pb2_hit{ tuvp[i].x, tuvp[i].y, tuvp[i].z, inst[i], inst[i] < 0 ? -1 : (int32_t)prim };
    }
    return PB2_OK;
    PB2_CATCH
}
int pb2_trace_any(pb2_scene *scene, const float *rays, uint64_t n, uint8_t *occluded) {
    PB2_TRY
    if (!scene || (n && (!rays || !occluded))) return fail(PB2_ERR_ARG, "pb2_trace_any: null");
    Scene &s = *S(scene);
    if (!n) return PB2_OK;
    DevBuf<float4> d_rays(n * 2);
    DevBuf<uint32_t> d_occ(n);
    PB2_CUDA(cudaMemcpyAsync(d_rays.ptr, rays, n * 32, cudaMemcpyHostToDevice, s.stream));
    trace_any_dev(s, d_rays.ptr, n, d_occ.ptr);
    std::vector<uint32_t> occ(n);
    PB2_CUDA(cudaMemcpyAsync(occ.data(), d_occ.ptr, n * 4, cudaMemcpyDeviceToHost, s.stream));
    PB2_CUDA(cudaStreamSynchronize(s.stream));
    for (uint64_t i = 0; i < n; ++i) occluded[i] = (uint8_t)occ[i];
    return PB2_OK;
    PB2_CATCH
}
int pb2_bvh_download(pb2_scene *scene, void *nodes, uint64_t *n_nodes, void *prims, uint64_t *n_prims) {
    PB2_TRY
    Scene &s = *S(scene);
    if (!s.bvh_valid) return fail(PB2_ERR_STATE, "pb2_bvh_download: no BVH");
    if (n_nodes) *n_nodes = s.n_nodes;
    if (n_prims) *n_prims = s.n_prims;
    PB2_CUDA(cudaStreamSynchronize(s.stream));
    if (nodes && s.n_nodes) PB2_CUDA(cudaMemcpy(nodes, s.d_nodes.ptr, (size_t)s.n_nodes * sizeof(Bvh8Node), cudaMemcpyDeviceToHost));
    if (prims && s.n_prims) PB2_CUDA(cudaMemcpy(prims, s.d_prims.ptr, (size_t)s.n_prims * sizeof(PrimRec), cudaMemcpyDeviceToHost));
    return PB2_OK;
    PB2_CATCH
}
int pb2_render(pb2_scene *scene, const pb2_launch_params *params) {
    PB2_TRY
    if (!scene || !params) return fail(PB2_ERR_ARG, "pb2_render: null");
    render(*S(scene), *params);
    return PB2_OK;
    PB2_CATCH
}
int pb2_synchronize(pb2_scene *scene) {
    PB2_TRY
    PB2_CUDA(cudaStreamSynchronize(S(scene)->stream));
    return PB2_OK;
    PB2_CATCH
}
int pb2_render_stats_get(pb2_scene *scene, pb2_render_stats *stats) {
    PB2_TRY
    if (!scene || !stats) return fail(PB2_ERR_ARG, "pb2_render_stats_get: null");
    collect_render_stats(*S(scene));
    *stats = S(scene)->render_stats;
    return PB2_OK;
    PB2_CATCH
}
int pb2_scene_set_option(pb2_scene *scene, const char *name, int64_t value) {
    if (!scene || !name) return fail(PB2_ERR_ARG, "pb2_scene_set_option: null");
    Scene &s = *S(scene);
    const std::string n = name;
    if (n == "profiling") s.profiling = value != 0;
    else if (n == "counting") s.counting = value != 0;
    else if (n == "sort_by_material") s.sort_by_material = value < 0 ? -1 : (value != 0);
    else if (n == "refill_threshold") s.refill_threshold = (int)std::min<int64_t>(33, std::max<int64_t>(0, value));
    else if (n == "shade_variant") s.shade_variant = (int)value;
    else if (n == "two_lanes") s.two_lanes = value != 0;
    else if (n == "coop_prims") s.coop_prims = (int)std::min<int64_t>(1, std::max<int64_t>(-1, value));
    else if (n == "paths_in_flight") s.paths_in_flight = (uint64_t)std::max<int64_t>(0, value);
    else if (n == "instancing") s.instancing = (int)std::min<int64_t>(2, std::max<int64_t>(0, value)), s.bvh_valid = false, s.blas_valid = false;
    else if (n == "ploc_radius") s.ploc_radius = (int)std::min<int64_t>(16, std::max<int64_t>(1, value)), s.bvh_valid = false;
    else if (n == "morton_bits") s.morton_bits = value <= 0 ? 0 : (int)std::min<int64_t>(63, std::max<int64_t>(15, value)), s.bvh_valid = false, s.blas_valid = false;
    else if (n == "collapse") s.collapse = value != 0, s.bvh_valid = false, s.blas_valid = false;
    else if (n == "collapse_prim_cost_pct") s.collapse_prim_cost_pct = (int)std::min<int64_t>(1000, std::max<int64_t>(1, value)), s.bvh_valid = false, s.blas_valid = false;
    else if (n == "l2_persist_mb") s.l2_persist_mb = (int)std::min<int64_t>(1024, std::max<int64_t>(0, value)), s.l2_dirty = true;
    else if (n == "l2_window_mb") s.l2_window_mb = (int)std::min<int64_t>(1 << 20, std::max<int64_t>(0, value)), s.l2_dirty = true;
    else return fail(PB2_ERR_ARG, "pb2_scene_set_option: unknown option " + n);
    return PB2_OK;
}
// ---- multi-GPU (comm.cu) ----
int pb2_comm_unique_id(uint8_t id[PB2_COMM_ID_BYTES]) {
    PB2_TRY
    if (!id) return fail(PB2_ERR_ARG, "pb2_comm_unique_id: null");
    comm_unique_id(id);
    return PB2_OK;
    PB2_CATCH
}
int pb2_comm_create(pb2_comm **comm, int n_ranks, int rank, const uint8_t id[PB2_COMM_ID_BYTES]) {
    PB2_TRY
    if (!comm) return fail(PB2_ERR_ARG, "pb2_comm_create: null");
    *comm = reinterpret_cast<pb2_comm *>(comm_create(n_ranks, rank, id));
    return PB2_OK;
    PB2_CATCH
}
int pb2_comm_destroy(pb2_comm *comm) {
    PB2_TRY
    comm_destroy(reinterpret_cast<Comm *>(comm));
    return PB2_OK;
    PB2_CATCH
}
int pb2_comm_reduce_frames(pb2_comm *comm, pb2_scene *scene, const void *sum_buffer, void *frame_buffer, uint64_t n_pixels, uint32_t total_spp, int mode, int root) {
    PB2_TRY
    if (!comm || !scene) return fail(PB2_ERR_ARG, "pb2_comm_reduce_frames: null");
    if (mode != PB2_REDUCE_ROOT && mode != PB2_REDUCE_ALL) return fail(PB2_ERR_ARG, "pb2_comm_reduce_frames: mode is PB2_REDUCE_ROOT or PB2_REDUCE_ALL");
    comm_reduce_frames(*reinterpret_cast<Comm *>(comm), *S(scene), static_cast<const float4 *>(sum_buffer), static_cast<float4 *>(frame_buffer), n_pixels, total_spp, mode, root);
    return PB2_OK;
    PB2_CATCH
}
int pb2_comm_synchronize(pb2_comm *comm) {
    PB2_TRY
    if (!comm) return fail(PB2_ERR_ARG, "pb2_comm_synchronize: null");
    comm_synchronize(*reinterpret_cast<Comm *>(comm));
    return PB2_OK;
    PB2_CATCH
}
int pb2_comm_last_reduction(pb2_comm *comm, float *ms, uint64_t *bytes) {
    PB2_TRY
    if (!comm) return fail(PB2_ERR_ARG, "pb2_comm_last_reduction: null");
    comm_last_reduction(*reinterpret_cast<Comm *>(comm), ms, bytes);
    return PB2_OK;
    PB2_CATCH
}
int pb2_comm_nccl_version(int *version) {
    PB2_TRY
    if (!version) return fail(PB2_ERR_ARG, "pb2_comm_nccl_version: null");
    *version = comm_nccl_version();
    return PB2_OK;
    PB2_CATCH
}
// rank r of N renders the seeds base + r + k N.  weak: every rank renders `spp` frames per step (the step
// covers spp * N seeds); strong: the step's `spp` frames are split, rank r takes ceil((spp - r) / N) of them.
int pb2_shard_plan(int rank, int n_ranks, uint32_t step, uint32_t spp, int strong, uint32_t *first_seed, uint32_t *seed_stride, uint32_t *spp_rank, uint32_t *spp_total) {
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || spp < 1) return fail(PB2_ERR_ARG, "pb2_shard_plan: bad rank / size / spp");
    const uint32_t n = (uint32_t)n_ranks, r = (uint32_t)rank;
    const uint32_t total = strong ? spp : spp * n;
    if (first_seed) *first_seed = step * total + r;
    if (seed_stride) *seed_stride = n;
    if (spp_rank) *spp_rank = strong ? (spp > r ? (spp - r + n - 1) / n : 0u) : spp;
    if (spp_total) *spp_total = total;
    return PB2_OK;
}
int pb2_finalize_sum(pb2_scene *scene, const void *sum_buffer, void *frame_buffer, uint64_t n_pixels, uint32_t total_spp) {
    PB2_TRY
    if (!scene || !sum_buffer || !frame_buffer || !total_spp) return fail(PB2_ERR_ARG, "pb2_finalize_sum: bad argument");
    finalize_sum(*S(scene), static_cast<const float4 *>(sum_buffer), static_cast<float4 *>(frame_buffer), n_pixels, total_spp);
    return PB2_OK;
    PB2_CATCH
}
}
