// Pupil::resource — what a scene file says, before any device-side precompute.
//
//   util::Texture (RGB / checkerboard / bitmap header)   framework/util/texture.h:21-60
//   resource::Material + per-type loaders                framework/resource/material.{h,cpp}
//   named IOR tables                                     framework/render/material/ior.h
//   resource::Emitter                                    framework/resource/emitter.h
//   Shape / ShapeInstance / ShapeManager                 framework/resource/shape.{h,cpp}
//   Scene::{LoadFromXML, LoadXmlObj}                     framework/resource/scene.{h,cpp}
//   typed property readers                               framework/resource/xml/util_loader.cpp
//
// Out of scope here: hair.  It is parsed,
// warned about and replaced by neutral defaults so a scene still loads.
#pragma once
#include "util.h"
#include "xml.h"

#include <filesystem>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace Pupil {
// framework/render/material/predefine.h:15-22 + decl/material_decl.inl — the order is the queue sort key
enum class EMatType : int { Unknown = 0, Diffuse = 1, Dielectric, RoughDielectric, Conductor, RoughConductor, Plastic, RoughPlastic, Twosided, Count };
inline constexpr const char *S_MAT_TYPE_NAME[] = { "diffuse", "dielectric", "roughdielectric", "conductor", "roughconductor", "plastic", "roughplastic", "twosided" };

namespace util {
enum class ETextureType : int { RGB = 0, Bitmap = 1, Checkerboard = 2 };
enum class ETextureAddressMode : int { Wrap = 0, Clamp = 1, Mirror = 2, Border = 3 }; // texture.h:10-15
enum class ETextureFilterMode : int { Point = 0, Linear = 1 };                        // texture.h:17-20
struct BitmapTexture { // texture.h:36-56: texels are owned by resource::TextureManager
    size_t w = 0, h = 0;
    const float *data = nullptr;
    ETextureAddressMode address_mode = ETextureAddressMode::Wrap;
    ETextureFilterMode filter_mode = ETextureFilterMode::Linear;
};
struct Texture {
    ETextureType type = ETextureType::RGB;
    Float3 rgb{ 0.f };              // RGB: color
    Float3 patch1{ 0.f }, patch2{ 0.f }; // checkerboard: xml color0 -> patch1, color1 -> patch2 (scene.cpp:170-172)
    BitmapTexture bitmap;
    Transform transform;            // to_uv
};
}// namespace util

namespace material {
float LoadDielectricIor(std::string_view value, float default_value) noexcept;                 // ior.h:178-194
bool LoadConductorIor(std::string_view name, util::Float3 &eta, util::Float3 &k) noexcept;     // ior.h:196-207
}// namespace material

namespace resource {
class Scene;

// one flat record instead of the reference's union: every type reads only its own slots
struct Material {
    EMatType type = EMatType::Unknown;
    bool twosided = false;
    float int_ior = 1.5046f, ext_ior = 1.000277f;
    bool nonlinear = false;
    util::Texture alpha, eta, k;
    util::Texture reflectance;          // diffuse.reflectance | (rough)plastic.diffuse_reflectance
    util::Texture specular_reflectance, specular_transmittance;
};
Material LoadMaterialFromXml(const xml::Object *obj, Scene *scene) noexcept;

// resource/texture.{h,cpp}: images are loaded once per path and shared; "mem:KEY" names registered images
class TextureManager : public util::Singleton<TextureManager> {
public:
    util::Texture GetTexture(std::string_view path) noexcept; // bitmap, or mid-grey RGB (with a warning) when the image cannot be read
    bool RegisterImage(std::string_view key, const float *rgba, size_t w, size_t h) noexcept;
    void Clear() noexcept;

private:
    struct ImageData {
        size_t w = 0, h = 0;
        std::vector<float> rgba;
    };
    std::unordered_map<std::string, std::unique_ptr<ImageData>> m_images;
};

enum class EEmitterType { Unknown, Area, Point, ConstEnv, EnvMap };
struct Emitter {
    EEmitterType type = EEmitterType::Unknown;
    util::Texture radiance;   // area | env map (bitmap)
    util::Float3 color{ 0.f }; // const env radiance | point intensity
    util::Float3 position{ 0.f };
    float scale = 1.f;         // env map
    util::Transform transform; // env map
};

enum class EShapeType : int { _unknown = 0, _obj, _sphere, _cube, _rectangle, _hair };

struct Mesh {
    bool face_normals = false, flip_normals = false, flip_tex_coords = false;
    uint32_t vertex_num = 0, face_num = 0;
    const float *positions = nullptr, *normals = nullptr, *texcoords = nullptr;
    const uint32_t *indices = nullptr;
};
struct Sphere {
    bool flip_normals = false;
    float radius = 1.f;
    util::Float3 center{ 0.f };
};
struct Shape {
    uint32_t id = 0;
    std::string file_path;
    EShapeType type = EShapeType::_unknown;
    Mesh mesh;
    Sphere sphere;
    util::AABB aabb;
};
struct ShapeInstance {
    std::string name;
    Shape *shape = nullptr;
    Material mat;
    bool is_emitter = false;
    Emitter emitter;
    util::Transform transform;
};
ShapeInstance LoadShapeInstanceFromXml(const xml::Object *obj, Scene *scene) noexcept;

// Shapes are process-wide and shared: cube / rectangle / sphere are singletons (so `flip_normals` is
// last-writer-wins across their instances, shape.cpp:91,105,124), file meshes are cached by path.
class ShapeManager : public util::Singleton<ShapeManager> {
public:
    Shape *LoadMeshShape(std::string_view file_path) noexcept; // wavefront .obj
    // programmatic route for meshes that should not round-trip through text:
    // the arrays are copied; `key` plays the role of the file path
    // borrow = true: the arrays are NOT copied — the caller keeps them alive until the key is dropped (DropMeshShape / Clear).
    // A 30 M-triangle mesh is 840 MB: copying it cost 0.5 s per scene load, 80 % of the host side of a reload.
    Shape *LoadMeshShape(std::string_view key, const float *pos, const float *nrm, const float *uv, const uint32_t *idx, uint32_t nv, uint32_t nf,
                         bool borrow = false) noexcept;
    void DropMeshShape(std::string_view key) noexcept; // forget one registered mesh (no render object may still use it)
    Shape *LoadSphere() noexcept;
    Shape *LoadCube() noexcept;
    Shape *LoadRectangle() noexcept;
    Shape *GetShape(uint32_t id) noexcept;
    void Clear() noexcept;

private:
    struct MeshData {
        std::vector<float> positions, normals, texcoords; // owned copies (empty when the arrays are borrowed)
        std::vector<uint32_t> indices;
        const float *pos = nullptr, *nrm = nullptr, *uv = nullptr; // what the shape points at: the vectors above or the caller's arrays
        const uint32_t *idx = nullptr;
        uint32_t nv = 0, nf = 0;
        util::AABB aabb;
    };
    Shape *Register(std::unique_ptr<Shape> shape);
    Shape *MakeMeshShape(std::string_view key, EShapeType type, const MeshData &data);
    uint32_t m_shape_id_cnt = 0;
    Shape *m_sphere = nullptr, *m_cube = nullptr, *m_rect = nullptr;
    std::unordered_map<uint32_t, std::unique_ptr<Shape>> m_id_shapes;
    std::unordered_map<std::string, std::unique_ptr<MeshData>> m_meshes;
    std::unordered_map<std::string, Shape *> m_mesh_shape;
};

struct Integrator {
    int max_depth = 1;
};
struct Film {
    int w = 768, h = 576;
};
struct Sensor {
    float fov = 90.f; // always fov_y after loading (scene.cpp:122-127)
    float near_clip = 0.01f, far_clip = 10000.f;
    util::Transform transform;
    Film film;
};

class Scene {
public:
    std::filesystem::path scene_root_path;
    Integrator integrator;
    Sensor sensor;
    std::vector<ShapeInstance> shape_instances;
    std::vector<Emitter> emitters; // non-area emitters

    void Reset() noexcept;
    bool LoadFromXML(std::filesystem::path file) noexcept;
    bool LoadFromXML(std::string_view file_name, std::string_view root) noexcept;
    bool LoadFromXMLString(std::string_view text, std::filesystem::path root = {}) noexcept;
    void LoadXmlObj(const xml::Object *xml_obj, void *dst) noexcept;

private:
    bool LoadFromRoot(const xml::Object *root) noexcept;
};

namespace xml {
bool LoadInt(const Object *obj, std::string_view name, int &param, int default_value = 0) noexcept;
bool LoadFloat(const Object *obj, std::string_view name, float &param, float default_value = 0.f) noexcept;
bool LoadFloat3(const Object *obj, std::string_view name, util::Float3 &param, util::Float3 default_value = {}) noexcept; // 3 or 1 values
bool Load3Float(const Object *obj, std::string_view name, util::Float3 &param, util::Float3 default_value = {}) noexcept; // exactly 3
bool LoadBool(const Object *obj, std::string_view name, bool &param, bool default_value = false) noexcept;
bool LoadTextureOrRGB(const Object *obj, Scene *scene, std::string_view name, util::Texture &param, util::Float3 default_value = {}) noexcept;
bool LoadTransform(const Object *obj, void *dst) noexcept;
}// namespace xml
}// namespace resource
}// namespace Pupil
