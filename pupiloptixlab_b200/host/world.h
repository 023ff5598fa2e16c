// Pupil::world — resource::Scene -> device-side tables, behind the pb2 C ABI instead of OptiX.
//
//   World              framework/world/world.{h,cpp}        (LoadScene, render objects, AABB; GetIASHandle -> GetSceneHandle)
//   RenderObject       framework/world/render_object.{h,cpp}
//   CameraHelper       framework/world/camera.{h,cpp}
//   EmitterHelper      framework/world/emitter.{h,cpp}      (area-emitter table, selection probabilities)
//   Material::LoadMaterial  framework/render/material/optix_material.cpp:41-130
//
// What is REPLACED: GASManager / IASManager (optixAccelBuild, per-shape GAS + instance AS) — World hands
// every render object to pb2_scene_add_instance and pb2_bvh_build builds one compressed 8-wide BVH on the GPU.
#pragma once
#include "../../include/pb2.h"
#include "resource.h"

#include <memory>
#include <unordered_map>

namespace Pupil {
enum class EWorldEvent { CameraChange, CameraMove, CameraFovChange, CameraViewChange, RenderInstanceTransform, RenderInstanceUpdate, RenderInstanceRemove };

namespace optix::material {
// host precompute of one material: eta = int/ext, plastic sampling weight and internal diffuse reflectance
pb2_material LoadMaterial(const resource::Material &mat) noexcept;
pb2_texture ToDeviceTexture(const util::Texture &tex) noexcept; // CudaTextureManager::GetCudaTexture, cuda/texture.cpp:104-131
// CudaTextureManager::GetCudaTextureObject / Clear (cuda/texture.cpp:60-102,133-143): one texture object per (image, modes)
uint64_t GetDeviceBitmap(const util::BitmapTexture &bitmap) noexcept; // 0 when no device is present (host-only parse)
void ClearDeviceBitmaps() noexcept;
float DiffuseFresnelReflectance(float eta) noexcept;           // fresnel::DiffuseReflectance, fresnel.h:58-84
}// namespace optix::material

namespace world {
class CameraHelper {
public:
    void Reset(const util::CameraDesc &desc) noexcept;
    util::CameraDesc GetDesc() const noexcept { return m_desc; }
    void SetFov(float fov) noexcept;
    void SetFovDelta(float fov_delta) noexcept;
    void SetAspectRatio(float aspect_ratio) noexcept;
    void SetNearClip(float near_clip) noexcept;
    void SetFarClip(float far_clip) noexcept;
    void SetWorldTransform(util::Transform to_world) noexcept;
    void Rotate(float delta_x, float delta_y) noexcept;
    void Move(util::Float3 translation) noexcept;
    util::Camera &GetUtilCamera() noexcept { return m_camera; }
    util::Mat4 GetSampleToCameraMatrix() noexcept { return m_camera.GetSampleToCameraMatrix(); }
    util::Mat4 GetProjectionMatrix() noexcept { return m_camera.GetProjectionMatrix(); }
    util::Mat4 GetToWorldMatrix() noexcept { return m_camera.GetToWorldMatrix(); }
    util::Mat4 GetViewMatrix() noexcept { return m_camera.GetViewMatrix(); }
    // replaces GetCudaMemory(): pushes the two matrices into the pb2 scene when they changed
    void Upload(pb2_scene *scene) noexcept;

private:
    util::CameraDesc m_desc;
    util::Camera m_camera;
    bool m_dirty = true;
};

class EmitterHelper {
public:
    ~EmitterHelper() { FreeEnvTables(); }
    void Clear() noexcept;
    size_t AddAreaEmitter(const resource::ShapeInstance &ins) noexcept; // returns the table size afterwards
    void ResetAreaEmitter(const resource::ShapeInstance &ins, size_t offset) noexcept;
    // drops the table entries [offset, offset + count) of a removed render object (the caller shifts the later objects'
    // offsets and calls ComputeProbability)
    void RemoveAreaEmitters(size_t offset, size_t count) noexcept;
    void AddEmitter(const resource::Emitter &emitter) noexcept;
    void ComputeProbability() noexcept;
    const std::vector<pb2_emitter> &GetAreaEmitters() const noexcept { return m_areas; }
    const pb2_emitter *GetEnvEmitter() const noexcept { return m_env.type == PB2_EMIT_NONE ? nullptr : &m_env; }
    // host copies of the env-map tables (BuildEnvMapCdfTable): row_cdf[h + 1], row_weight[h], col_cdf[(w + 1) * h]
    const std::vector<float> &GetEnvRowCdf() const noexcept { return m_row_cdf; }
    const std::vector<float> &GetEnvRowWeight() const noexcept { return m_row_weight; }
    const std::vector<float> &GetEnvColCdf() const noexcept { return m_col_cdf; }
    // replaces GetEmitterGroup(): uploads the table when it changed
    void Upload(pb2_scene *scene) noexcept;
    void Invalidate() noexcept { m_dirty = true; } // the device scene was cleared: upload again

private:
    void SetMeshAreaEmitter(const resource::ShapeInstance &ins, size_t offset) noexcept;
    void SetSphereAreaEmitter(const resource::ShapeInstance &ins, size_t offset) noexcept;
    void BuildEnvMapCdfTable(const resource::Emitter &emitter) noexcept;
    void FreeEnvTables() noexcept;
    std::vector<pb2_emitter> m_areas;
    pb2_emitter m_env{};
    std::vector<float> m_row_cdf, m_row_weight, m_col_cdf;
    void *m_env_tables_device = nullptr; // m_env_cdf_weight_cuda_memory of the reference (world/emitter.cpp:362-376)
    bool m_dirty = true;
};

struct RenderObject {
    enum class EGeoType { TriMesh, Sphere };
    std::string name;
    uint32_t shape_id = 0;
    const resource::Shape *shape = nullptr;
    unsigned int visibility_mask = 1;
    util::Transform transform;
    util::AABB aabb;
    EGeoType geo_type = EGeoType::TriMesh;
    bool flip_normals = false, flip_tex_coords = false; // copied from the (shared) shape at creation, render_object.cpp:32,55-56
    pb2_material mat{};
    bool is_emitter = false;
    unsigned int sub_emitters_num = 0;

    explicit RenderObject(const resource::ShapeInstance &ins, unsigned int v_mask = 1) noexcept;
    void Reset(const resource::Shape *shape) noexcept;
    void UpdateTransform(const util::Transform &new_transform) noexcept;
    void ApplyTransform(const util::Transform &new_transform) noexcept;
    RenderObject(const RenderObject &) = delete;
    RenderObject &operator=(const RenderObject &) = delete;
};

class World : public util::Singleton<World> {
public:
    std::unique_ptr<resource::Scene> scene;
    std::unique_ptr<CameraHelper> camera;
    std::unique_ptr<EmitterHelper> emitters;

    void Init() noexcept;
    void Destroy() noexcept;
    bool LoadScene(std::filesystem::path scene_file_path) noexcept;
    bool LoadScene(resource::Scene *scene) noexcept;

    // replaces GetIASHandle(gas_offset, allow_update): the pb2 scene with an up-to-date BVH, camera and
    // emitter table.  Rebuilds the BVH when instances changed.
    pb2_scene *GetSceneHandle() noexcept;
    const pb2_build_stats &GetBuildStats() const noexcept { return m_build_stats; }
    // 0 = LBVH, 1 = binned SAH along the Morton order, 2 = SAH-driven clustering (pb2_scene_set_builder); applies to the next build
    void SetBvhBuilder(int builder) noexcept;
    // which shapes get a bottom-level tree shared by their render objects, as GASManager::RefGAS does for every shape
    // (gas_manager.cpp:10): 0 none, 1 shapes used more than once (default), 2 every mesh; applies to the next build
    void SetInstancing(int mode) noexcept;

    RenderObject *GetRenderObject(std::string_view name) const noexcept;
    RenderObject *GetRenderObject(size_t index) const noexcept;
    void RemoveRenderObject(size_t index) noexcept;
    // first entry of the object's emitters in EmitterHelper's table (-1: not an emitter)
    int GetEmitterOffset(const RenderObject *ro) const noexcept;
    // a failed scene load leaves nothing behind: no render objects, no emitters, no device geometry
    void Reset() noexcept;
    void UpdateRenderObject(RenderObject *ro) noexcept;
    std::vector<RenderObject *> GetRenderobjects() noexcept;
    void SetDirty() noexcept { m_geometry_dirty = true; }
    bool IsDirty() const noexcept { return m_geometry_dirty; }
    util::Camera &GetUtilCamera() noexcept { return camera->GetUtilCamera(); }
    util::AABB GetAABB() noexcept;

private:
    void RebuildDeviceScene() noexcept;
    std::vector<std::unique_ptr<RenderObject>> m_ros;
    std::unordered_map<const RenderObject *, size_t> m_ro_emitter_offset, m_ro_in_scene_index;
    pb2_scene *m_pb2 = nullptr;
    pb2_build_stats m_build_stats{};
    bool m_geometry_dirty = true;
    bool m_transform_dirty = false; // only object transforms changed since the last build: the top level is rebuilt, nothing else
    int m_builder = -1, m_instancing = -1;
};
}// namespace world
}// namespace Pupil
