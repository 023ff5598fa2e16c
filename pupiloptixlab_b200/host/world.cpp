#include "world.h"

#include <cmath>
#include <cstring>
#include <map>
#include <tuple>

namespace Pupil {
namespace {
void Pb2Check(int code, const char *what) {
    if (code != PB2_OK) Log::Error("%s failed: %s", what, pb2_last_error());
}
float Luminance(util::Float3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; } // optix/util.h:161-163
float Max3(float r, float g, float b) { return r > g ? (r > b ? r : b) : (g > b ? g : b); }
util::Float3 PixelAverage(const util::Texture &t) { // optix_material.cpp:9-36
    if (t.type == util::ETextureType::Checkerboard)
        return util::Float3{ (t.patch1.x + t.patch2.x) * 0.5f, (t.patch1.y + t.patch2.y) * 0.5f, (t.patch1.z + t.patch2.z) * 0.5f };
    if (t.type == util::ETextureType::Bitmap) { // :19-30: running fp32 sums over all texels, then one division
        float r = 0.f, g = 0.f, b = 0.f;
        for (size_t i = 0, j = 0; i < t.bitmap.h; ++i)
            for (size_t k = 0; k < t.bitmap.w; ++k, j += 4) r += t.bitmap.data[j], g += t.bitmap.data[j + 1], b += t.bitmap.data[j + 2];
        const float inv = 1.0f / (1.f * t.bitmap.h * t.bitmap.w); // float3 / float multiplies by the reciprocal (cuda/vec_math.h:425-428)
        return util::Float3{ r * inv, g * inv, b * inv };
    }
    return t.rgb;
}
float SelectWeight(const util::Texture &t) { // world/emitter.cpp:77-101
    if (t.type == util::ETextureType::Checkerboard) return (Max3(t.patch1.x, t.patch1.y, t.patch1.z) + Max3(t.patch2.x, t.patch2.y, t.patch2.z)) * 0.5f;
    if (t.type == util::ETextureType::Bitmap) { // :89-99 with the reference's own (i * w + j) indexing; indices past the image are skipped
        float w = 0.f;
        for (size_t i = 0; i < t.bitmap.w; i++)
            for (size_t j = 0; j < t.bitmap.h; j++) {
                const size_t px = i * t.bitmap.w + j;
                if (px >= t.bitmap.w * t.bitmap.h) continue;
                w += Max3(t.bitmap.data[px * 4 + 0], t.bitmap.data[px * 4 + 1], t.bitmap.data[px * 4 + 2]);
            }
        return w / (1.f * t.bitmap.w * t.bitmap.h);
    }
    return Max3(t.rgb.x, t.rgb.y, t.rgb.z);
}
std::map<std::tuple<const float *, size_t, size_t, int, int>, uint64_t> g_device_bitmaps;
}// namespace

namespace optix::material {
float DiffuseFresnelReflectance(float eta) noexcept {
    if (eta < 1) return -1.4399f * (eta * eta) + 0.7099f * eta + 0.6681f + 0.0636f / eta; // Egan & Hilgeman 1973
    const float i1 = 1.0f / eta, i2 = i1 * i1, i3 = i2 * i1, i4 = i3 * i1, i5 = i4 * i1;  // d'Eon & Irving 2011
    return 0.919317f - 3.4793f * i1 + 6.75335f * i2 - 7.80989f * i3 + 4.98554f * i4 - 1.36881f * i5;
}
uint64_t GetDeviceBitmap(const util::BitmapTexture &b) noexcept {
    if (!b.data || !b.w || !b.h) return 0;
    const auto key = std::make_tuple(b.data, b.w, b.h, static_cast<int>(b.address_mode), static_cast<int>(b.filter_mode));
    auto it = g_device_bitmaps.find(key);
    if (it != g_device_bitmaps.end()) return it->second;
    uint64_t handle = 0;
    if (pb2_device_count() > 0)
        Pb2Check(pb2_bitmap_create(b.data, static_cast<uint32_t>(b.w), static_cast<uint32_t>(b.h), static_cast<int>(b.address_mode),
                                   static_cast<int>(b.filter_mode), &handle),
                 "pb2_bitmap_create");
    g_device_bitmaps[key] = handle;
    return handle;
}
void ClearDeviceBitmaps() noexcept {
    for (auto &kv : g_device_bitmaps)
        if (kv.second) pb2_bitmap_destroy(kv.second);
    g_device_bitmaps.clear();
}
pb2_texture ToDeviceTexture(const util::Texture &tex) noexcept {
    pb2_texture t{};
    t.type = static_cast<int32_t>(tex.type);
    if (tex.type == util::ETextureType::Bitmap) t.bitmap = GetDeviceBitmap(tex.bitmap);
    const util::Float3 a = tex.type == util::ETextureType::Checkerboard ? tex.patch1 : tex.rgb;
    for (int c = 0; c < 3; ++c) t.a[c] = a.e[c], t.b[c] = tex.patch2.e[c];
    for (int c = 0; c < 4; ++c) t.r0[c] = tex.transform.matrix.re[0][c], t.r1[c] = tex.transform.matrix.re[1][c];
    return t;
}
pb2_material LoadMaterial(const resource::Material &m) noexcept {
    pb2_material d{};
    d.type = static_cast<int32_t>(m.type), d.twosided = m.twosided, d.eta = 1.f, d.nonlinear = m.nonlinear;
    auto T = ToDeviceTexture;
    switch (m.type) {
        case EMatType::Diffuse: d.tex[0] = T(m.reflectance); break;
        case EMatType::Dielectric:
            d.eta = m.int_ior / m.ext_ior;
            d.tex[0] = T(m.specular_reflectance), d.tex[1] = T(m.specular_transmittance);
            break;
        case EMatType::RoughDielectric:
            d.eta = m.int_ior / m.ext_ior;
            d.tex[0] = T(m.specular_reflectance), d.tex[1] = T(m.specular_transmittance), d.tex[2] = T(m.alpha);
            break;
        case EMatType::Conductor: d.tex[0] = T(m.specular_reflectance), d.tex[1] = T(m.eta), d.tex[2] = T(m.k); break;
        case EMatType::RoughConductor: d.tex[0] = T(m.specular_reflectance), d.tex[1] = T(m.eta), d.tex[2] = T(m.k), d.tex[3] = T(m.alpha); break;
        case EMatType::Plastic:
        case EMatType::RoughPlastic: {
            d.eta = m.int_ior / m.ext_ior;
            const float yd = Luminance(PixelAverage(m.reflectance)), ys = Luminance(PixelAverage(m.specular_reflectance));
            d.specular_sampling_weight = ys / (ys + yd);
            d.int_fdr = DiffuseFresnelReflectance(1.f / d.eta);
            d.tex[0] = T(m.reflectance), d.tex[1] = T(m.specular_reflectance);
            if (m.type == EMatType::RoughPlastic) d.tex[2] = T(m.alpha);
        } break;
        default: break;
    }
    return d;
}
}// namespace optix::material

namespace world {
// ---- camera -----------------------------------------------------------------------------------------------
void CameraHelper::Reset(const util::CameraDesc &desc) noexcept {
    m_camera.SetProjectionFactor(desc.fov_y, desc.aspect_ratio, desc.near_clip, desc.far_clip);
    m_camera.SetWorldTransform(desc.to_world);
    m_desc = desc, m_dirty = true;
}
void CameraHelper::SetFov(float fov) noexcept {
    m_desc.fov_y = fov < 0.012f ? 0.012f : (fov > 180.f ? 180.f : fov);
    m_camera.SetFov(m_desc.fov_y);
    m_dirty = true;
}
void CameraHelper::SetFovDelta(float fov_delta) noexcept { SetFov(m_desc.fov_y + fov_delta); }
void CameraHelper::SetAspectRatio(float aspect_ratio) noexcept {
    m_desc.aspect_ratio = aspect_ratio;
    m_camera.SetProjectionFactor(m_desc.fov_y, m_desc.aspect_ratio, m_desc.near_clip, m_desc.far_clip);
    m_dirty = true;
}
void CameraHelper::SetNearClip(float near_clip) noexcept {
    m_desc.near_clip = near_clip;
    m_camera.SetProjectionFactor(m_desc.fov_y, m_desc.aspect_ratio, m_desc.near_clip, m_desc.far_clip);
    m_dirty = true;
}
void CameraHelper::SetFarClip(float far_clip) noexcept {
    m_desc.far_clip = far_clip;
    m_camera.SetProjectionFactor(m_desc.fov_y, m_desc.aspect_ratio, m_desc.near_clip, m_desc.far_clip);
    m_dirty = true;
}
void CameraHelper::SetWorldTransform(util::Transform to_world) noexcept {
    m_desc.to_world = to_world;
    m_camera.SetWorldTransform(to_world);
    m_dirty = true;
}
void CameraHelper::Rotate(float dx, float dy) noexcept { m_camera.Rotate(dx, dy), m_dirty = true; }
void CameraHelper::Move(util::Float3 t) noexcept { m_camera.Move(t), m_dirty = true; }
void CameraHelper::Upload(pb2_scene *scene) noexcept {
    if (!m_dirty || !scene) return;
    const util::Mat4 s2c = m_camera.GetSampleToCameraMatrix(), c2w = m_camera.GetToWorldMatrix();
    m_desc.to_world.matrix = c2w;
    Pb2Check(pb2_scene_set_camera(scene, s2c.e, c2w.e), "pb2_scene_set_camera");
    m_dirty = false;
}

// ---- emitters ---------------------------------------------------------------------------------------------
void EmitterHelper::Clear() noexcept {
    m_areas.clear();
    m_env = pb2_emitter{};
    m_row_cdf.clear(), m_row_weight.clear(), m_col_cdf.clear();
    FreeEnvTables();
    m_dirty = true;
}
void EmitterHelper::FreeEnvTables() noexcept {
    if (m_env_tables_device) pb2_free(m_env_tables_device);
    m_env_tables_device = nullptr;
}
// world/emitter.cpp:107-149, statement for statement (fp32 running sums in the same order)
void EmitterHelper::BuildEnvMapCdfTable(const resource::Emitter &emitter) noexcept {
    const size_t w = emitter.radiance.bitmap.w, h = emitter.radiance.bitmap.h;
    const float *data = emitter.radiance.bitmap.data;
    constexpr float kPi = 3.14159265358979323846f;
    m_col_cdf.resize((w + 1) * h);
    m_row_cdf.resize(h + 1);
    m_row_weight.resize(h);
    size_t col_index = 0, row_index = 0;
    float row_sum = 0.f;
    m_row_cdf[row_index++] = 0.f;
    for (auto y = 0u; y < h; ++y) {
        float col_sum = 0.f;
        m_col_cdf[col_index++] = 0.f;
        for (auto x = 0u; x < w; ++x) {
            const auto pixel_index = y * w + x;
            col_sum += Luminance(util::Float3{ data[pixel_index * 4 + 0], data[pixel_index * 4 + 1], data[pixel_index * 4 + 2] });
            m_col_cdf[col_index++] = col_sum;
        }
        for (auto x = 1u; x < w; ++x) m_col_cdf[col_index - x - 1] /= col_sum;
        m_col_cdf[col_index - 1] = 1.f;
        const float weight = std::sin((y + 0.5f) * kPi / h);
        m_row_weight[y] = weight;
        row_sum += col_sum * weight;
        m_row_cdf[row_index++] = row_sum;
    }
    for (auto y = 1u; y < h; ++y) m_row_cdf[row_index - y - 1] /= row_sum;
    m_row_cdf[row_index - 1] = 1.f;
    if (row_sum == 0) Log::Warn("The environment map is completely black.");
    m_env.normalization = 1.f / (row_sum * (2.f * kPi / w) * (kPi / h));
    m_env.map_w = static_cast<uint32_t>(w), m_env.map_h = static_cast<uint32_t>(h);
}
void EmitterHelper::SetMeshAreaEmitter(const resource::ShapeInstance &ins, size_t offset) noexcept {
    const util::Mat4 &xf = ins.transform.matrix;
    const util::Mat4 normal_transform = xf.GetInverse().GetTranspose();
    const resource::Mesh &mesh = ins.shape->mesh;
    const pb2_texture radiance = optix::material::ToDeviceTexture(ins.emitter.radiance);
    const float select_weight = SelectWeight(ins.emitter.radiance);
    for (uint32_t f = 0; f < mesh.face_num; ++f) {
        pb2_emitter e{};
        e.type = PB2_EMIT_TRI, e.radiance = radiance;
        util::Float3 p[3];
        for (int k = 0; k < 3; ++k) {
            const uint32_t vi = mesh.indices[f * 3 + k];
            p[k] = util::Transform::TransformPoint(util::Float3{ mesh.positions[vi * 3], mesh.positions[vi * 3 + 1], mesh.positions[vi * 3 + 2] }, xf);
            for (int c = 0; c < 3; ++c) e.pos[k][c] = p[k].e[c];
            if (mesh.texcoords) e.uv[k][0] = mesh.texcoords[vi * 2], e.uv[k][1] = mesh.texcoords[vi * 2 + 1];
        }
        const util::Float3 v1{ p[1].x - p[0].x, p[1].y - p[0].y, p[1].z - p[0].z }, v2{ p[2].x - p[0].x, p[2].y - p[0].y, p[2].z - p[0].z };
        const util::Float3 cr{ v1.y * v2.z - v1.z * v2.y, v1.z * v2.x - v1.x * v2.z, v1.x * v2.y - v1.y * v2.x };
        const float cr_len = std::sqrt(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
        for (int k = 0; k < 3; ++k) {
            const uint32_t vi = mesh.indices[f * 3 + k];
            util::Float3 n;
            if (mesh.normals) n = util::Transform::TransformNormal(util::Float3{ mesh.normals[vi * 3], mesh.normals[vi * 3 + 1], mesh.normals[vi * 3 + 2] }, normal_transform);
            else n = util::Float3{ cr.x * (1.0f / cr_len), cr.y * (1.0f / cr_len), cr.z * (1.0f / cr_len) }; // defined: the reference reads a null array here
            for (int c = 0; c < 3; ++c) e.nrm[k][c] = n.e[c];
        }
        e.area = cr_len * 0.5f;
        e.weight = select_weight * e.area;
        m_areas[offset + f] = e;
    }
}
void EmitterHelper::SetSphereAreaEmitter(const resource::ShapeInstance &ins, size_t offset) noexcept {
    pb2_emitter e{};
    e.type = PB2_EMIT_SPHERE;
    e.radiance = optix::material::ToDeviceTexture(ins.emitter.radiance);
    const util::Float3 c0 = ins.shape->sphere.center;
    const util::Float3 o = util::Transform::TransformPoint(c0, ins.transform.matrix);
    const util::Float3 p = util::Transform::TransformPoint(util::Float3{ c0.x + ins.shape->sphere.radius, c0.y, c0.z }, ins.transform.matrix);
    const util::Float3 d{ o.x - p.x, o.y - p.y, o.z - p.z };
    e.center[0] = o.x, e.center[1] = o.y, e.center[2] = o.z;
    e.radius = std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
    e.area = 4 * 3.14159265358979323846f * e.radius * e.radius;
    e.weight = SelectWeight(ins.emitter.radiance) * e.area;
    m_areas[offset] = e;
}
size_t EmitterHelper::AddAreaEmitter(const resource::ShapeInstance &ins) noexcept {
    const size_t offset = m_areas.size();
    if (ins.shape->type == resource::EShapeType::_sphere) {
        m_areas.resize(offset + 1);
        SetSphereAreaEmitter(ins, offset);
    } else if (ins.shape->type != resource::EShapeType::_hair && ins.shape->type != resource::EShapeType::_unknown) {
        m_areas.resize(offset + ins.shape->mesh.face_num);
        SetMeshAreaEmitter(ins, offset);
    }
    m_dirty = true;
    return m_areas.size();
}
void EmitterHelper::ResetAreaEmitter(const resource::ShapeInstance &ins, size_t offset) noexcept {
    if (ins.shape->type == resource::EShapeType::_sphere) SetSphereAreaEmitter(ins, offset);
    else SetMeshAreaEmitter(ins, offset);
    m_dirty = true;
}
void EmitterHelper::RemoveAreaEmitters(size_t offset, size_t count) noexcept {
    if (offset >= m_areas.size()) return;
    m_areas.erase(m_areas.begin() + offset, m_areas.begin() + std::min(m_areas.size(), offset + count));
    m_dirty = true;
}
void EmitterHelper::AddEmitter(const resource::Emitter &emitter) noexcept {
    if (emitter.type == resource::EEmitterType::ConstEnv) { // world/emitter.cpp:283-292
        m_env = pb2_emitter{};
        m_env.type = PB2_EMIT_CONST_ENV;
        m_env.radiance.type = PB2_TEX_RGB;
        for (int c = 0; c < 3; ++c) m_env.radiance.a[c] = emitter.color.e[c];
        m_env.weight = 1.f;
        const util::AABB aabb = util::Singleton<World>::instance()->GetAABB();
        for (int c = 0; c < 3; ++c) m_env.center[c] = (aabb.max.e[c] + aabb.min.e[c]) * 0.5f;
        m_dirty = true;
    } else if (emitter.type == resource::EEmitterType::EnvMap && emitter.radiance.type == util::ETextureType::Bitmap) { // :293-312
        m_env = pb2_emitter{};
        m_env.type = PB2_EMIT_ENV_MAP;
        m_env.radiance = optix::material::ToDeviceTexture(emitter.radiance);
        m_env.scale = emitter.scale;
        const util::AABB aabb = util::Singleton<World>::instance()->GetAABB();
        for (int c = 0; c < 3; ++c) m_env.center[c] = (aabb.max.e[c] + aabb.min.e[c]) * 0.5f;
        m_env.weight = 1.f;
        const util::Mat4 &m = emitter.transform.matrix;
        const util::Mat4 to_local = m.GetInverse();
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) m_env.to_world[r * 3 + c] = m.re[r][c], m_env.to_local[r * 3 + c] = to_local.re[r][c];
        BuildEnvMapCdfTable(emitter);
        FreeEnvTables();
        m_dirty = true;
    }
    // point emitters are parsed but never sampled by the reference
}
void EmitterHelper::ComputeProbability() noexcept { // world/emitter.cpp:321-337
    float area_weight_sum = 0.f;
    for (auto &e : m_areas) area_weight_sum += e.weight;
    if (!m_areas.empty())
        for (auto &e : m_areas) e.select_probability = e.weight / area_weight_sum * m_areas.size();
    const size_t emitter_num = (m_env.type == PB2_EMIT_NONE ? 0 : 1) + m_areas.size();
    for (auto &e : m_areas) e.select_probability = e.select_probability / emitter_num;
    m_env.select_probability = m_env.weight / emitter_num;
    m_dirty = true;
}
void EmitterHelper::Upload(pb2_scene *scene) noexcept {
    if (!m_dirty || !scene) return;
    if (m_env.type == PB2_EMIT_ENV_MAP && !m_env_tables_device) { // :362-376: one allocation, row_cdf | row_weight | col_cdf
        const size_t n_row = m_row_cdf.size(), n_w = m_row_weight.size(), n_col = m_col_cdf.size();
        Pb2Check(pb2_malloc(&m_env_tables_device, (n_row + n_w + n_col) * sizeof(float)), "pb2_malloc(env tables)");
        if (m_env_tables_device) {
            float *d = static_cast<float *>(m_env_tables_device);
            pb2_upload(d, m_row_cdf.data(), n_row * sizeof(float));
            pb2_upload(d + n_row, m_row_weight.data(), n_w * sizeof(float));
            pb2_upload(d + n_row + n_w, m_col_cdf.data(), n_col * sizeof(float));
        }
        m_env.env_tables = static_cast<const float *>(m_env_tables_device);
    }
    Pb2Check(pb2_scene_set_emitters(scene, m_areas.data(), (uint32_t)m_areas.size(), GetEnvEmitter()), "pb2_scene_set_emitters");
    m_dirty = false;
}

// ---- render objects ---------------------------------------------------------------------------------------
RenderObject::RenderObject(const resource::ShapeInstance &ins, unsigned int v_mask) noexcept
    : name(ins.name), visibility_mask(v_mask), transform(ins.transform) {
    Reset(ins.shape);
    is_emitter = ins.is_emitter;
    mat = optix::material::LoadMaterial(ins.mat);
}
void RenderObject::Reset(const resource::Shape *s) noexcept {
    shape = s, shape_id = s->id;
    aabb = s->aabb;
    aabb.Transform(transform.matrix);
    if (s->type == resource::EShapeType::_sphere) {
        geo_type = EGeoType::Sphere;
        flip_normals = s->sphere.flip_normals, flip_tex_coords = false;
        sub_emitters_num = 1;
    } else {
        geo_type = EGeoType::TriMesh;
        flip_normals = s->mesh.flip_normals, flip_tex_coords = s->mesh.flip_tex_coords;
        sub_emitters_num = s->mesh.face_num;
    }
    EventDispatcher<EWorldEvent::RenderInstanceUpdate>(static_cast<void *>(this));
}
void RenderObject::UpdateTransform(const util::Transform &t) noexcept {
    transform = t;
    EventDispatcher<EWorldEvent::RenderInstanceTransform>(static_cast<void *>(this));
}
void RenderObject::ApplyTransform(const util::Transform &t) noexcept {
    transform.matrix = t.matrix * transform.matrix;
    EventDispatcher<EWorldEvent::RenderInstanceTransform>(static_cast<void *>(this));
}

// ---- world ------------------------------------------------------------------------------------------------
void World::Init() noexcept {
    scene = std::make_unique<resource::Scene>();
    camera = std::make_unique<CameraHelper>();
    emitters = std::make_unique<EmitterHelper>();
    static bool bound = false;
    if (!bound) {
        bound = true;
        // world.cpp:15-43: a transformed instance invalidates the acceleration structure and its emitters
        EventBinder<EWorldEvent::RenderInstanceTransform>([this](void *p) {
            auto *ro = static_cast<RenderObject *>(p);
            if (ro == nullptr || !scene) return;
            auto idx = m_ro_in_scene_index.find(ro);
            if (idx != m_ro_in_scene_index.end()) {
                ro->aabb = ro->shape->aabb;
                ro->aabb.Transform(ro->transform.matrix);
                auto &ins = scene->shape_instances[idx->second];
                ins.transform = ro->transform;
                if (ro->is_emitter) {
                    emitters->ResetAreaEmitter(ins, m_ro_emitter_offset[ro]);
                    emitters->ComputeProbability();
                }
            }
            // IASManager::UpdateInstance + IAS::Update (ias_manager.cpp:116-151): the new transform goes to the device scene and
            // only the top level of the acceleration structure is rebuilt; bottom-level trees stay
            bool incremental = false;
            if (m_pb2 && !m_geometry_dirty) {
                for (size_t k = 0; k < m_ros.size(); ++k)
                    if (m_ros[k].get() == ro) {
                        incremental = pb2_scene_set_instance_transform(m_pb2, (uint32_t)k, ro->transform.matrix.e) == PB2_OK;
                        break;
                    }
            }
            if (incremental) m_transform_dirty = true;
            else m_geometry_dirty = true;
            EventDispatcher<EWorldEvent::RenderInstanceUpdate>(p);
        });
    }
}
void World::Destroy() noexcept {
    m_ros.clear(), m_ro_emitter_offset.clear(), m_ro_in_scene_index.clear();
    scene.reset(), camera.reset(), emitters.reset();
    if (m_pb2) pb2_scene_destroy(m_pb2), m_pb2 = nullptr;
    util::Singleton<resource::ShapeManager>::instance()->Clear();
    optix::material::ClearDeviceBitmaps();
    util::Singleton<resource::TextureManager>::instance()->Clear();
}
bool World::LoadScene(std::filesystem::path path) noexcept {
    if (!std::filesystem::exists(path)) {
        Log::Warn("scene file [%s] does not exist", path.string().c_str());
        return false;
    }
    Timer timer;
    timer.Start();
    m_ros.clear();
    if (!scene->LoadFromXML(path) || !LoadScene(scene.get())) {
        Log::Error("scene load failed: %s", path.string().c_str());
        Reset(); // the previous scene's geometry must not outlive a failed load
        return false;
    }
    timer.Stop();
    size_t tri_num = 0;
    for (auto &ro : m_ros)
        if (ro->geo_type == RenderObject::EGeoType::TriMesh) tri_num += ro->shape->mesh.face_num;
    Log::Info("scene triangles: %zu, loaded in %.3f s", tri_num, timer.ElapsedSeconds());
    EventDispatcher<EWorldEvent::CameraChange>();
    return true;
}
bool World::LoadScene(resource::Scene *s) noexcept {
    if (s == nullptr) return false;
    util::CameraDesc desc;
    desc.fov_y = s->sensor.fov;
    desc.aspect_ratio = static_cast<float>(s->sensor.film.w) / s->sensor.film.h;
    desc.near_clip = s->sensor.near_clip, desc.far_clip = s->sensor.far_clip;
    desc.to_world = s->sensor.transform;
    camera->Reset(desc);

    m_ros.clear(), m_ro_emitter_offset.clear(), m_ro_in_scene_index.clear();
    m_ros.reserve(s->shape_instances.size());
    emitters->Clear();
    size_t emitter_offset = 0;
    for (size_t index = 0; index < s->shape_instances.size(); ++index) {
        auto &ins = s->shape_instances[index];
        if (!ins.shape || ins.shape->type == resource::EShapeType::_unknown) continue;
        m_ros.push_back(std::make_unique<RenderObject>(ins));
        m_ro_in_scene_index[m_ros.back().get()] = index;
        if (ins.is_emitter) {
            m_ro_emitter_offset[m_ros.back().get()] = emitter_offset;
            emitter_offset = emitters->AddAreaEmitter(ins);
        }
    }
    for (auto &e : s->emitters) emitters->AddEmitter(e);
    emitters->ComputeProbability();
    m_geometry_dirty = true;
    return true;
}

void World::SetBvhBuilder(int builder) noexcept { m_builder = builder, m_geometry_dirty = true; }
void World::SetInstancing(int mode) noexcept { m_instancing = mode, m_geometry_dirty = true; }

// IASManager::SetInstance + GAS/IAS builds, replaced: every render object becomes one pb2 instance.  Meshes
// are uploaded once per shape; the emitter offset of an instance is the running sum of sub_emitters_num over
// the emitting objects before it (example/path_tracer/pt_pass.cpp:178-193).
void World::RebuildDeviceScene() noexcept {
    if (!m_pb2) Pb2Check(pb2_scene_create(&m_pb2), "pb2_scene_create");
    if (!m_pb2) return;
    Pb2Check(pb2_scene_clear(m_pb2), "pb2_scene_clear");
    emitters->Invalidate(); // pb2_scene_clear drops the emitter table with the geometry
    if (m_builder >= 0) Pb2Check(pb2_scene_set_builder(m_pb2, m_builder), "pb2_scene_set_builder");
    if (m_instancing >= 0) Pb2Check(pb2_scene_set_option(m_pb2, "instancing", m_instancing), "pb2_scene_set_option(instancing)");
    std::unordered_map<uint32_t, uint32_t> mesh_of_shape;
    for (auto &ro : m_ros) {
        uint32_t mesh_id = PB2_MESH_SPHERE;
        if (ro->geo_type == RenderObject::EGeoType::TriMesh) {
            auto it = mesh_of_shape.find(ro->shape_id);
            if (it == mesh_of_shape.end()) {
                const resource::Mesh &m = ro->shape->mesh;
                uint32_t id = 0;
                Pb2Check(pb2_scene_add_mesh(m_pb2, m.positions, m.normals, m.texcoords, m.indices, m.vertex_num, m.face_num, &id), "pb2_scene_add_mesh");
                it = mesh_of_shape.emplace(ro->shape_id, id).first;
            }
            mesh_id = it->second;
        }
        const uint32_t flags = (ro->flip_normals ? PB2_INST_FLIP_NORMALS : 0u) | (ro->flip_tex_coords ? PB2_INST_FLIP_TEX : 0u);
        // the offset EmitterHelper assigned when the object's emitters were added (kept current by RemoveRenderObject): the
        // device table is EmitterHelper's table, so no second running sum here
        int offset = -1;
        if (ro->is_emitter) {
            auto it = m_ro_emitter_offset.find(ro.get());
            if (it != m_ro_emitter_offset.end()) offset = (int)it->second;
        }
        Pb2Check(pb2_scene_add_instance(m_pb2, mesh_id, ro->transform.matrix.e, flags, &ro->mat, offset, nullptr), "pb2_scene_add_instance");
    }
    Pb2Check(pb2_bvh_build(m_pb2, &m_build_stats), "pb2_bvh_build");
    m_geometry_dirty = false, m_transform_dirty = false;
}
pb2_scene *World::GetSceneHandle() noexcept {
    if (m_geometry_dirty || !m_pb2) RebuildDeviceScene();
    else if (m_transform_dirty) {
        Pb2Check(pb2_bvh_build(m_pb2, &m_build_stats), "pb2_bvh_build (top level)");
        m_transform_dirty = false;
    }
    camera->Upload(m_pb2);
    emitters->Upload(m_pb2);
    return m_pb2;
}

RenderObject *World::GetRenderObject(std::string_view name) const noexcept {
    for (auto &ro : m_ros)
        if (ro->name == name) return ro.get();
    return nullptr;
}
RenderObject *World::GetRenderObject(size_t index) const noexcept { return index < m_ros.size() ? m_ros[index].get() : nullptr; }
void World::RemoveRenderObject(size_t index) noexcept {
    if (index >= m_ros.size()) return;
    RenderObject *ro = m_ros[index].get();
    EventDispatcher<EWorldEvent::RenderInstanceRemove>(static_cast<void *>(ro));
    // an emitting object takes its entries out of the emitter table with it: later objects move down by its entry count, and
    // the selection probabilities are recomputed without the removed light
    if (auto it = m_ro_emitter_offset.find(ro); it != m_ro_emitter_offset.end()) {
        const size_t offset = it->second, count = ro->sub_emitters_num;
        emitters->RemoveAreaEmitters(offset, count);
        for (auto &kv : m_ro_emitter_offset)
            if (kv.second > offset) kv.second -= count;
        emitters->ComputeProbability();
    }
    m_ro_emitter_offset.erase(ro), m_ro_in_scene_index.erase(ro);
    m_ros.erase(m_ros.begin() + index);
    m_geometry_dirty = true;
}
int World::GetEmitterOffset(const RenderObject *ro) const noexcept {
    auto it = m_ro_emitter_offset.find(ro);
    return it == m_ro_emitter_offset.end() ? -1 : (int)it->second;
}
void World::Reset() noexcept {
    m_ros.clear(), m_ro_emitter_offset.clear(), m_ro_in_scene_index.clear();
    if (emitters) emitters->Clear();
    if (m_pb2) Pb2Check(pb2_scene_clear(m_pb2), "pb2_scene_clear");
    m_build_stats = pb2_build_stats{};
    m_geometry_dirty = true;
}
void World::UpdateRenderObject(RenderObject *) noexcept { m_geometry_dirty = true; }
std::vector<RenderObject *> World::GetRenderobjects() noexcept {
    std::vector<RenderObject *> out;
    out.reserve(m_ros.size());
    for (auto &ro : m_ros) out.push_back(ro.get());
    return out;
}
util::AABB World::GetAABB() noexcept {
    util::AABB aabb;
    for (auto &ro : m_ros) aabb.Merge(ro->aabb);
    return aabb;
}
}// namespace world
}// namespace Pupil
