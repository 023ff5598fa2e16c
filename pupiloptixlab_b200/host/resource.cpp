#include "resource.h"
#include <cfloat>
#include <thread>
#include "image.h"

#include <charconv>
#include <cstring>
#include <fstream>
#include <map>

namespace Pupil {
// ---- named IORs (framework/render/material/ior.h) -------------------------------------------------------
// Measured constants (Hecht, Optics; ~589 nm) and the RGB eta/k of the metals people actually name in scenes.
namespace material {
namespace {
struct DielectricEntry {
    const char *name;
    float ior;
};
constexpr DielectricEntry kDielectrics[] = {
    { "vacuum", 1.0f }, { "helium", 1.000036f }, { "hydrogen", 1.000132f }, { "air", 1.000277f }, { "carbon dioxide", 1.00045f },
    { "water", 1.3330f }, { "acetone", 1.36f }, { "ethanol", 1.361f }, { "carbon tetrachloride", 1.461f }, { "glycerol", 1.4729f },
    { "benzene", 1.501f }, { "silicone oil", 1.52045f }, { "bromine", 1.661f }, { "water ice", 1.31f }, { "fused quartz", 1.458f },
    { "pyrex", 1.470f }, { "acrylic glass", 1.49f }, { "polypropylene", 1.49f }, { "bk7", 1.5046f }, { "sodium chloride", 1.544f },
    { "amber", 1.55f }, { "pet", 1.5750f }, { "diamond", 2.419f },
};
struct ConductorEntry {
    const char *name;
    float eta[3], k[3];
};
constexpr ConductorEntry kConductors[] = {
    { "a-C", { 2.93785f, 2.22242f, 1.96400f }, { 0.88555f, 0.79763f, 0.81356f } },
    { "Ag", { 0.15494f, 0.11648f, 0.13809f }, { 4.81810f, 3.11562f, 2.14240f } },
    { "Al", { 1.65394f, 0.87850f, 0.52012f }, { 9.20430f, 6.25621f, 4.82675f } },
    { "Au", { 0.14282f, 0.37414f, 1.43944f }, { 3.97472f, 2.38066f, 1.59981f } },
    { "Be", { 4.17618f, 3.17830f, 2.77819f }, { 3.82730f, 3.00374f, 2.86293f } },
    { "Cr", { 4.36041f, 2.91052f, 1.65119f }, { 5.19538f, 4.22239f, 3.74700f } },
    { "Cu", { 0.19999f, 0.92209f, 1.09988f }, { 3.90464f, 2.44763f, 2.13765f } },
    { "Fe", { 2.76404f, 1.95417f, 1.62766f }, { 3.83077f, 2.73841f, 2.31812f } },
    { "Hg", { 2.39384f, 1.43697f, 0.90762f }, { 6.31420f, 4.36266f, 3.41454f } },
    { "Ir", { 3.07986f, 2.07777f, 1.61446f }, { 5.58028f, 4.05855f, 3.26033f } },
    { "K", { 0.06391f, 0.04631f, 0.03810f }, { 2.09975f, 1.34607f, 0.91128f } },
    { "Li", { 0.26525f, 0.19519f, 0.22045f }, { 3.53305f, 2.30618f, 1.66505f } },
    { "Mo", { 4.47417f, 3.51799f, 2.77018f }, { 4.10240f, 3.41361f, 3.14393f } },
    { "Na", { 0.06014f, 0.05602f, 0.06186f }, { 3.17254f, 2.10800f, 1.57575f } },
    { "Nb", { 3.41288f, 2.78427f, 2.39051f }, { 3.43408f, 2.73183f, 2.57445f } },
    { "Ni", { 2.36225f, 1.65983f, 1.46395f }, { 4.48929f, 3.04369f, 2.34046f } },
    { "Rh", { 2.58031f, 1.85624f, 1.55114f }, { 6.76790f, 4.69297f, 3.96766f } },
    { "Ta", { 2.05820f, 2.38802f, 2.62250f }, { 2.40293f, 1.73767f, 1.94291f } },
    { "W", { 4.36142f, 3.29330f, 2.99191f }, { 3.49325f, 2.59934f, 2.26838f } },
    { "none", { 0.f, 0.f, 0.f }, { 1.f, 1.f, 1.f } }, // the perfect mirror
};
}// namespace

float LoadDielectricIor(std::string_view str, float default_value) noexcept {
    if (str.empty()) return default_value;
    float value = 0.f;
    auto [p, ec] = std::from_chars(str.data(), str.data() + str.size(), value);
    if (ec == std::errc() && p == str.data() + str.size()) return value;
    for (auto &e : kDielectrics)
        if (str == e.name) return e.ior;
    return default_value;
}
bool LoadConductorIor(std::string_view name, util::Float3 &eta, util::Float3 &k) noexcept {
    if (name.empty()) return false;
    for (auto &e : kConductors)
        if (name == e.name || (name.size() > 6 && name.substr(name.size() - 6) == "_palik" && name.substr(0, name.size() - 6) == e.name)) {
            eta = util::Float3{ e.eta[0], e.eta[1], e.eta[2] };
            k = util::Float3{ e.k[0], e.k[1], e.k[2] };
            return true;
        }
    return false;
}
}// namespace material

namespace resource {
// ---- typed property readers (xml/util_loader.cpp) ------------------------------------------------------------
namespace xml {
namespace {
bool ParseFloats(std::string_view value, std::string_view delims, std::vector<float> &out) {
    out.clear();
    for (auto &piece : util::Split(value, delims)) {
        try {
            out.push_back(std::stof(piece));
        } catch (...) {
            return false;
        }
    }
    return true;
}
}// namespace

bool LoadInt(const Object *obj, std::string_view name, int &param, int default_value) noexcept {
    const std::string value = obj->GetProperty(name);
    param = default_value;
    if (value.empty()) return false;
    try {
        param = std::stoi(value);
    } catch (...) {
        return false;
    }
    return true;
}
bool LoadFloat(const Object *obj, std::string_view name, float &param, float default_value) noexcept {
    const std::string value = obj->GetProperty(name);
    param = default_value;
    if (value.empty()) return false;
    try {
        param = std::stof(value);
    } catch (...) {
        return false;
    }
    return true;
}
static bool Float3From(std::string_view what, const std::string &value, util::Float3 &param, util::Float3 default_value, bool allow_scalar) noexcept {
    if (value.empty()) {
        param = default_value;
        return false;
    }
    std::vector<float> v;
    if (ParseFloats(value, ",", v) && v.size() == 3) {
        param = util::Float3{ v[0], v[1], v[2] };
        return true;
    }
    if (allow_scalar && v.size() == 1) {
        param = util::Float3{ v[0] };
        return true;
    }
    Log::Warn("[%.*s] has %zu components (%s)", (int)what.size(), what.data(), v.size(), allow_scalar ? "must be 3 or 1" : "must be 3");
    return false;
}
bool LoadFloat3(const Object *obj, std::string_view name, util::Float3 &param, util::Float3 default_value) noexcept {
    return Float3From(name, obj->GetProperty(name), param, default_value, true);
}
bool Load3Float(const Object *obj, std::string_view name, util::Float3 &param, util::Float3 default_value) noexcept {
    return Float3From(name, obj->GetProperty(name), param, default_value, false);
}
bool LoadBool(const Object *obj, std::string_view name, bool &param, bool default_value) noexcept {
    const std::string value = obj->GetProperty(name);
    param = default_value;
    if (value == "true") return param = true, true;
    if (value == "false") return param = false, true;
    return false;
}
bool LoadTextureOrRGB(const Object *obj, Scene *scene, std::string_view name, util::Texture &param, util::Float3 default_value) noexcept {
    auto [texture, rgb] = obj->GetParameter(name);
    param = util::Texture{};
    if (texture == nullptr) {
        util::Float3 color = default_value;
        const bool given = !rgb.empty();
        if (given) Float3From(name, rgb, color, default_value, true);
        param.type = util::ETextureType::RGB, param.rgb = color;
        return given;
    }
    scene->LoadXmlObj(texture, &param);
    return true;
}

// util_loader.cpp:128-191.  Precedence: matrix > lookat > scale/rotate/translate, and the s/r/t order is
// fixed (scale, then rotate, then translate) whatever the XML order was.
static bool LoadTransform3D(const Object *obj, util::Transform *transform) noexcept {
    const std::string value = obj->GetProperty("matrix");
    if (!value.empty()) {
        std::vector<float> m;
        ParseFloats(value, " \t\r\n,", m);
        if (m.size() == 16) {
            for (int i = 0; i < 16; ++i) transform->matrix.e[i] = m[i];
        } else if (m.size() == 9) { // 3x3 into the upper-left block
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) transform->matrix.re[r][c] = m[r * 3 + c];
        } else {
            Log::Warn("transform matrix has %zu values (must be 9 or 16)", m.size());
            for (size_t i = 0; i < m.size() && i < 16; ++i) transform->matrix.e[i] = m[i];
        }
        return true;
    }
    if (const Object *look_at = obj->GetUniqueSubObject("lookat")) {
        util::Float3 origin, target, up;
        Load3Float(look_at, "origin", origin, { 1.f, 0.f, 0.f });
        Load3Float(look_at, "target", target, { 0.f, 0.f, 0.f });
        Load3Float(look_at, "up", up, { 0.f, 1.f, 0.f });
        transform->LookAt(origin, target, up);
        // mitsuba (+X left, +Z view) -> Pupil (+X right, -Z view): negate the X and Z basis columns
        for (int r = 0; r < 3; ++r) transform->matrix.re[r][0] *= -1, transform->matrix.re[r][2] *= -1;
        if (!obj->GetProperty("scale").empty() || obj->GetUniqueSubObject("rotate") || !obj->GetProperty("translate").empty())
            Log::Warn("transform scale/rotate/translate ignored because a lookat exists");
        return true;
    }
    if (util::Float3 s; LoadFloat3(obj, "scale", s)) transform->Scale(s.x, s.y, s.z);
    if (const Object *rot = obj->GetUniqueSubObject("rotate")) {
        util::Float3 axis;
        float angle;
        if (Load3Float(rot, "axis", axis) && LoadFloat(rot, "angle", angle)) transform->Rotate(axis.x, axis.y, axis.z, angle);
    }
    if (util::Float3 t; Load3Float(obj, "translate", t)) transform->Translate(t.x, t.y, t.z);
    return true;
}
bool LoadTransform(const Object *obj, void *dst) noexcept {
    if (obj == nullptr || dst == nullptr) return false;
    auto *transform = static_cast<util::Transform *>(dst);
    if (obj->var_name == "to_world") return LoadTransform3D(obj, transform);
    if (obj->var_name == "to_uv") { // scale only (util_loader.cpp:198-205)
        if (util::Float3 s; LoadFloat3(obj, "scale", s)) transform->Scale(s.x, s.y, s.z);
        return true;
    }
    Log::Warn("transform [%s] unknown", obj->var_name.c_str());
    return false;
}
}// namespace xml

// ---- materials (resource/material.cpp:26-190) -----------------------------------------------------------------
Material LoadMaterialFromXml(const xml::Object *obj, Scene *scene) noexcept {
    Material mat;
    if (obj == nullptr || scene == nullptr) return mat;
    int index = -1;
    for (int i = 0; i < 8; ++i)
        if (obj->type == S_MAT_TYPE_NAME[i]) index = i;
    if (index < 0) {
        Log::Warn("unknown bsdf [%s]", obj->type.c_str());
        return mat;
    }
    const EMatType type = static_cast<EMatType>(index + 1);
    if (type == EMatType::Twosided) {
        mat = LoadMaterialFromXml(obj->GetUniqueSubObject("bsdf"), scene);
        mat.twosided = true;
        return mat;
    }
    mat.type = type;
    const bool plastic = type == EMatType::Plastic || type == EMatType::RoughPlastic;
    const bool conductor = type == EMatType::Conductor || type == EMatType::RoughConductor;
    if (type == EMatType::Diffuse) {
        xml::LoadTextureOrRGB(obj, scene, "reflectance", mat.reflectance, { 0.5f });
        return mat;
    }
    if (conductor) {
        util::Float3 eta, k;
        if (!material::LoadConductorIor(obj->GetProperty("material"), eta, k)) eta = { 0.f }, k = { 1.f };
        xml::LoadTextureOrRGB(obj, scene, "eta", mat.eta, eta);
        xml::LoadTextureOrRGB(obj, scene, "k", mat.k, k);
    } else {
        mat.int_ior = material::LoadDielectricIor(obj->GetProperty("int_ior"), plastic ? 1.49f : 1.5046f);
        mat.ext_ior = material::LoadDielectricIor(obj->GetProperty("ext_ior"), 1.000277f);
    }
    if (type == EMatType::RoughDielectric || type == EMatType::RoughConductor || type == EMatType::RoughPlastic)
        xml::LoadTextureOrRGB(obj, scene, "alpha", mat.alpha, { 0.1f });
    if (plastic) {
        mat.nonlinear = obj->GetProperty("nonlinear") == "true";
        xml::LoadTextureOrRGB(obj, scene, "diffuse_reflectance", mat.reflectance, { 0.5f });
    }
    xml::LoadTextureOrRGB(obj, scene, "specular_reflectance", mat.specular_reflectance, { 1.f });
    if (type == EMatType::Dielectric || type == EMatType::RoughDielectric)
        xml::LoadTextureOrRGB(obj, scene, "specular_transmittance", mat.specular_transmittance, { 1.f });
    return mat;
}

// ---- shapes (resource/shape.cpp) ------------------------------------------------------------------------------
namespace {
// built-in meshes; vertex and face order matter (primitive ids, emitter indices): shape.cpp:21-68
const float kRectPos[] = { -1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0 };
const float kRectNrm[] = { 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1 };
const float kRectUv[] = { 0, 0, 1, 0, 1, 1, 0, 1 };
const uint32_t kRectIdx[] = { 0, 1, 2, 0, 2, 3 };
struct CubeData {
    float pos[72], nrm[72], uv[48];
    uint32_t idx[36];
    CubeData() {
        // faces -X -Z +X +Z +Y -Y, four corners each
        const float p[72] = { -1, -1, -1, -1, -1, 1,  -1, 1,  1,  -1, 1,  -1, 1,  -1, -1, -1, -1, -1, -1, 1,  -1, 1,  1, -1,
                              1,  -1, 1,  1,  -1, -1, 1,  1,  -1, 1,  1,  1,  -1, -1, 1,  1,  -1, 1,  1,  1,  1,  -1, 1, 1,
                              -1, 1,  1,  1,  1,  1,  1,  1,  -1, -1, 1,  -1, -1, -1, -1, 1,  -1, -1, 1,  -1, 1,  -1, -1, 1 };
        std::memcpy(pos, p, sizeof p);
        const float fn[6][3] = { { -1, 0, 0 }, { 0, 0, -1 }, { 1, 0, 0 }, { 0, 0, 1 }, { 0, 1, 0 }, { 0, -1, 0 } };
        for (int f = 0; f < 6; ++f)
            for (int v = 0; v < 4; ++v) {
                for (int c = 0; c < 3; ++c) nrm[(f * 4 + v) * 3 + c] = fn[f][c];
                uv[(f * 4 + v) * 2] = (v == 1 || v == 2) ? 1.f : 0.f;
                uv[(f * 4 + v) * 2 + 1] = v >= 2 ? 1.f : 0.f;
            }
        for (uint32_t f = 0; f < 6; ++f) {
            const uint32_t quad[6] = { 0, 1, 2, 0, 2, 3 };
            for (int k = 0; k < 6; ++k) idx[f * 6 + k] = f * 4 + quad[k];
        }
    }
};
const CubeData &Cube() {
    static const CubeData c;
    return c;
}

EShapeType ShapeTypeOf(std::string_view name) {
    if (name == "obj") return EShapeType::_obj;
    if (name == "sphere") return EShapeType::_sphere;
    if (name == "cube") return EShapeType::_cube;
    if (name == "rectangle") return EShapeType::_rectangle;
    if (name == "hair") return EShapeType::_hair;
    return EShapeType::_unknown;
}

// Wavefront .obj reader: v / vt / vn / f with positive or negative indices, polygons triangulated as a
// fan in file order (what assimp's aiProcess_Triangulate does for convex faces); one mesh per file.
// Corners with identical (v, vt, vn) triples share a vertex.
bool ReadObj(const std::string &path, std::vector<float> &P, std::vector<float> &N, std::vector<float> &T, std::vector<uint32_t> &I) {
    std::ifstream f(path);
    if (!f) return false;
    std::vector<float> v, vn, vt;
    std::map<std::tuple<int, int, int>, uint32_t> remap;
    bool missing_n = false, missing_t = false;
    std::vector<std::tuple<int, int, int>> corners; // per output vertex
    std::string line;
    while (std::getline(f, line)) {
        const char *s = line.c_str();
        while (*s == ' ' || *s == '\t') ++s;
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            float x = 0, y = 0, z = 0;
            std::sscanf(s + 2, "%f %f %f", &x, &y, &z);
            v.insert(v.end(), { x, y, z });
        } else if (s[0] == 'v' && s[1] == 'n') {
            float x = 0, y = 0, z = 0;
            std::sscanf(s + 3, "%f %f %f", &x, &y, &z);
            vn.insert(vn.end(), { x, y, z });
        } else if (s[0] == 'v' && s[1] == 't') {
            float x = 0, y = 0;
            std::sscanf(s + 3, "%f %f", &x, &y);
            vt.insert(vt.end(), { x, y });
        } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
            std::vector<uint32_t> face;
            const char *p = s + 2;
            while (*p) {
                while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
                if (!*p) break;
                int idx[3] = { 0, 0, 0 }; // v, vt, vn (1-based, 0 = absent)
                for (int k = 0; k < 3; ++k) {
                    char *end = nullptr;
                    const long val = std::strtol(p, &end, 10);
                    if (end != p) idx[k] = (int)val;
                    p = end;
                    if (*p == '/') ++p;
                    else break;
                }
                while (*p && *p != ' ' && *p != '\t') ++p;
                const int nv = (int)v.size() / 3, nt = (int)vt.size() / 2, nn = (int)vn.size() / 3;
                const int iv = idx[0] < 0 ? nv + idx[0] : idx[0] - 1, it = idx[1] < 0 ? nt + idx[1] : idx[1] - 1, in = idx[2] < 0 ? nn + idx[2] : idx[2] - 1;
                if (iv < 0 || iv >= nv) return false;
                const auto key = std::make_tuple(iv, (it >= 0 && it < nt) ? it : -1, (in >= 0 && in < nn) ? in : -1);
                auto found = remap.find(key);
                if (found == remap.end()) {
                    found = remap.emplace(key, (uint32_t)corners.size()).first;
                    corners.push_back(key);
                    if (std::get<1>(key) < 0) missing_t = true;
                    if (std::get<2>(key) < 0) missing_n = true;
                }
                face.push_back(found->second);
            }
            for (size_t k = 2; k < face.size(); ++k) I.insert(I.end(), { face[0], face[k - 1], face[k] });
        }
    }
    if (I.empty()) return false;
    const bool use_n = !missing_n, use_t = !missing_t; // attributes are kept only when every corner has them
    for (auto &[iv, it, in] : corners) {
        P.insert(P.end(), { v[iv * 3], v[iv * 3 + 1], v[iv * 3 + 2] });
        if (use_n) N.insert(N.end(), { vn[in * 3], vn[in * 3 + 1], vn[in * 3 + 2] });
        if (use_t) T.insert(T.end(), { vt[it * 2], vt[it * 2 + 1] });
    }
    return true;
}
}// namespace

// bounding box of n packed float3 points, on up to eight threads (15 M vertices: 180 MB to read)
static util::AABB BoundsOf(const float *pos, uint32_t n) {
    const unsigned workers = n < (1u << 20) ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    std::vector<util::AABB> part(workers);
    auto run = [&](unsigned w) {
        const size_t b = (size_t)n * w / workers, e = (size_t)n * (w + 1) / workers;
        float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        for (size_t i = b; i < e; ++i)
            for (int k = 0; k < 3; ++k) lo[k] = std::min(lo[k], pos[i * 3 + k]), hi[k] = std::max(hi[k], pos[i * 3 + k]);
        if (e > b) part[w].Merge(util::Float3{ lo[0], lo[1], lo[2] }), part[w].Merge(util::Float3{ hi[0], hi[1], hi[2] });
    };
    std::vector<std::thread> threads;
    for (unsigned w = 1; w < workers; ++w) threads.emplace_back(run, w);
    run(0);
    for (auto &t : threads) t.join();
    util::AABB out;
    for (auto &p : part) out.Merge(p);
    return out;
}
Shape *ShapeManager::Register(std::unique_ptr<Shape> shape) {
    shape->id = m_shape_id_cnt++;
    Shape *p = shape.get();
    m_id_shapes[p->id] = std::move(shape);
    return p;
}
Shape *ShapeManager::MakeMeshShape(std::string_view key, EShapeType type, const MeshData &d) {
    auto shape = std::make_unique<Shape>();
    shape->file_path = key, shape->type = type;
    shape->mesh.vertex_num = d.nv, shape->mesh.face_num = d.nf;
    shape->mesh.positions = d.pos, shape->mesh.normals = d.nrm, shape->mesh.texcoords = d.uv, shape->mesh.indices = d.idx;
    shape->aabb = d.aabb;
    return Register(std::move(shape));
}
Shape *ShapeManager::LoadMeshShape(std::string_view file_path) noexcept {
    const std::string key(file_path);
    if (auto it = m_mesh_shape.find(key); it != m_mesh_shape.end()) return it->second;
    auto data = std::make_unique<MeshData>();
    if (!ReadObj(key, data->positions, data->normals, data->texcoords, data->indices)) {
        Log::Warn("mesh load from %s failed", key.c_str());
        return nullptr;
    }
    data->pos = data->positions.data(), data->nrm = data->normals.empty() ? nullptr : data->normals.data();
    data->uv = data->texcoords.empty() ? nullptr : data->texcoords.data(), data->idx = data->indices.data();
    data->nv = (uint32_t)(data->positions.size() / 3), data->nf = (uint32_t)(data->indices.size() / 3);
    data->aabb = BoundsOf(data->pos, data->nv);
    Shape *s = MakeMeshShape(key, EShapeType::_obj, *data);
    m_meshes[key] = std::move(data), m_mesh_shape[key] = s;
    return s;
}
Shape *ShapeManager::LoadMeshShape(std::string_view key_, const float *pos, const float *nrm, const float *uv, const uint32_t *idx, uint32_t nv, uint32_t nf,
                                   bool borrow) noexcept {
    const std::string key(key_);
    if (auto it = m_mesh_shape.find(key); it != m_mesh_shape.end()) return it->second;
    if (!pos || !idx || !nv || !nf) return nullptr;
    auto data = std::make_unique<MeshData>();
    if (borrow) {
        data->pos = pos, data->nrm = nrm, data->uv = uv, data->idx = idx;
    } else {
        data->positions.assign(pos, pos + (size_t)nv * 3);
        if (nrm) data->normals.assign(nrm, nrm + (size_t)nv * 3);
        if (uv) data->texcoords.assign(uv, uv + (size_t)nv * 2);
        data->indices.assign(idx, idx + (size_t)nf * 3);
        data->pos = data->positions.data(), data->nrm = nrm ? data->normals.data() : nullptr;
        data->uv = uv ? data->texcoords.data() : nullptr, data->idx = data->indices.data();
    }
    data->nv = nv, data->nf = nf;
    data->aabb = BoundsOf(data->pos, nv);
    Shape *s = MakeMeshShape(key, EShapeType::_obj, *data);
    m_meshes[key] = std::move(data), m_mesh_shape[key] = s;
    return s;
}
void ShapeManager::DropMeshShape(std::string_view key_) noexcept {
    const std::string key(key_);
    auto it = m_mesh_shape.find(key);
    if (it == m_mesh_shape.end()) return;
    m_id_shapes.erase(it->second->id);
    m_mesh_shape.erase(it);
    m_meshes.erase(key);
}
Shape *ShapeManager::LoadSphere() noexcept {
    if (m_sphere) return m_sphere;
    auto shape = std::make_unique<Shape>();
    shape->file_path = "sphere", shape->type = EShapeType::_sphere;
    shape->aabb.min = util::Float3{ -1.f }, shape->aabb.max = util::Float3{ 1.f };
    return m_sphere = Register(std::move(shape));
}
Shape *ShapeManager::LoadCube() noexcept {
    if (m_cube) return m_cube;
    auto shape = std::make_unique<Shape>();
    shape->file_path = "cube", shape->type = EShapeType::_cube;
    shape->mesh.vertex_num = 24, shape->mesh.face_num = 12;
    shape->mesh.positions = Cube().pos, shape->mesh.normals = Cube().nrm, shape->mesh.texcoords = Cube().uv, shape->mesh.indices = Cube().idx;
    shape->aabb.min = util::Float3{ -1.f }, shape->aabb.max = util::Float3{ 1.f };
    return m_cube = Register(std::move(shape));
}
Shape *ShapeManager::LoadRectangle() noexcept {
    if (m_rect) return m_rect;
    auto shape = std::make_unique<Shape>();
    shape->file_path = "rectangle", shape->type = EShapeType::_rectangle;
    shape->mesh.vertex_num = 4, shape->mesh.face_num = 2;
    shape->mesh.positions = kRectPos, shape->mesh.normals = kRectNrm, shape->mesh.texcoords = kRectUv, shape->mesh.indices = kRectIdx;
    shape->aabb.min = util::Float3{ -1.f, -1.f, 0.f }, shape->aabb.max = util::Float3{ 1.f, 1.f, 0.f };
    return m_rect = Register(std::move(shape));
}
Shape *ShapeManager::GetShape(uint32_t id) noexcept {
    auto it = m_id_shapes.find(id);
    return it == m_id_shapes.end() ? nullptr : it->second.get();
}
void ShapeManager::Clear() noexcept {
    m_id_shapes.clear(), m_meshes.clear(), m_mesh_shape.clear();
    m_sphere = m_cube = m_rect = nullptr;
}

ShapeInstance LoadShapeInstanceFromXml(const xml::Object *obj, Scene *scene) noexcept {
    ShapeInstance ins;
    if (obj == nullptr || scene == nullptr) return ins;
    auto *mngr = util::Singleton<ShapeManager>::instance();
    ins.name = obj->id;
    switch (ShapeTypeOf(obj->type)) {
        case EShapeType::_cube:
            ins.shape = mngr->LoadCube();
            xml::LoadBool(obj, "flip_normals", ins.shape->mesh.flip_normals, false); // writes into the shared shape
            ins.shape->mesh.face_normals = false, ins.shape->mesh.flip_tex_coords = false;
            break;
        case EShapeType::_rectangle:
            ins.shape = mngr->LoadRectangle();
            xml::LoadBool(obj, "flip_normals", ins.shape->mesh.flip_normals, false);
            ins.shape->mesh.face_normals = false, ins.shape->mesh.flip_tex_coords = false;
            break;
        case EShapeType::_sphere: { // unit sphere + T(center) S(radius) folded into the instance transform (shape.cpp:113-133)
            util::Float3 center;
            float radius;
            xml::Load3Float(obj, "center", center);
            xml::LoadFloat(obj, "radius", radius, 1.f);
            ins.shape = mngr->LoadSphere();
            xml::LoadBool(obj, "flip_normals", ins.shape->sphere.flip_normals, false);
            ins.transform.Scale(radius, radius, radius);
            ins.transform.Translate(center.x, center.y, center.z);
        } break;
        case EShapeType::_obj: {
            // "mem:<key>" names a mesh registered in memory (ShapeManager::LoadMeshShape(key, arrays...))
            const std::string filename = obj->GetProperty("filename");
            const std::string path = filename.rfind("mem:", 0) == 0 ? filename : (scene->scene_root_path / filename).lexically_normal().string();
            ins.shape = mngr->LoadMeshShape(path);
            if (!ins.shape) return ins;
            xml::LoadBool(obj, "face_normals", ins.shape->mesh.face_normals, false);
            xml::LoadBool(obj, "flip_tex_coords", ins.shape->mesh.flip_tex_coords, true);
            xml::LoadBool(obj, "flip_normals", ins.shape->mesh.flip_normals, false);
        } break;
        case EShapeType::_hair: Log::Warn("shape type [hair] is out of scope (OptiX built-in curves); skipped"); return ins;
        default: Log::Warn("unknown shape type [%s]", obj->type.c_str()); return ins;
    }
    scene->LoadXmlObj(obj->GetUniqueSubObject("bsdf"), &ins.mat);
    util::Transform transform;
    scene->LoadXmlObj(obj->GetUniqueSubObject("transform"), &transform);
    if (ins.shape->type == EShapeType::_sphere) ins.transform = util::Transform(transform.matrix * ins.transform.matrix);
    else ins.transform = transform;

    if (const xml::Object *em = obj->GetUniqueSubObject("emitter")) {
        scene->LoadXmlObj(em, &ins.emitter);
        if (ins.emitter.type != EEmitterType::Area) Log::Warn("only area emitters can be attached to a shape");
        else ins.is_emitter = true;
    }
    return ins;
}

// ---- scene (resource/scene.cpp) ---------------------------------------------------------------------------------
void Scene::Reset() noexcept {
    emitters.clear(), shape_instances.clear();
    integrator = Integrator{}, sensor = Sensor{};
}
bool Scene::LoadFromRoot(const xml::Object *root) noexcept {
    if (!root) return false;
    for (const xml::Object *o : root->sub_object) {
        switch (o->tag) {
            case xml::ETag::_integrator: LoadXmlObj(o, &integrator); break;
            case xml::ETag::_sensor: LoadXmlObj(o, &sensor); break;
            case xml::ETag::_shape: {
                auto ins = LoadShapeInstanceFromXml(o, this);
                if (ins.shape) shape_instances.push_back(std::move(ins));
            } break;
            case xml::ETag::_emitter: {
                Emitter e;
                LoadXmlObj(o, &e);
                if (e.type != EEmitterType::Area) emitters.push_back(e); // a free-standing area emitter has no geometry
            } break;
            default: break;
        }
    }
    return true;
}
bool Scene::LoadFromXML(std::filesystem::path file) noexcept {
    Reset();
    scene_root_path = file.parent_path();
    xml::Parser parser;
    return LoadFromRoot(parser.LoadFromFile(file.string()));
}
bool Scene::LoadFromXML(std::string_view file_name, std::string_view root) noexcept {
    const std::filesystem::path file = std::filesystem::path(root) / file_name;
    if (file.extension() != ".xml") {
        Log::Error("scene file format not supported");
        return false;
    }
    return LoadFromXML(file);
}
bool Scene::LoadFromXMLString(std::string_view text, std::filesystem::path root) noexcept {
    Reset();
    scene_root_path = std::move(root);
    xml::Parser parser;
    return LoadFromRoot(parser.LoadFromString(text));
}

namespace {
// "mem:KEY" names an image registered through TextureManager::RegisterImage; anything else is a path below the scene
std::string ResolveImagePath(const std::filesystem::path &root, const std::string &value) {
    if (value.rfind("mem:", 0) == 0) return value;
    return (root / value).make_preferred().string();
}
}// namespace

// ---- resource/texture.cpp: image cache ------------------------------------------------------------------------------
util::Texture TextureManager::GetTexture(std::string_view path) noexcept {
    const std::string key(path);
    auto it = m_images.find(key);
    if (it == m_images.end()) {
        util::Image image;
        if (key.rfind("mem:", 0) == 0 || !util::LoadImage(key, image)) {
            if (key.rfind("mem:", 0) == 0) Log::Warn("image [%s] was not registered", key.c_str());
            util::Texture grey; // the reference would hand a null bitmap to the device here
            grey.type = util::ETextureType::RGB, grey.rgb = util::Float3{ 0.5f };
            return grey;
        }
        auto data = std::make_unique<ImageData>();
        data->w = image.w, data->h = image.h, data->rgba = std::move(image.rgba);
        it = m_images.emplace(key, std::move(data)).first;
    }
    util::Texture t;
    t.type = util::ETextureType::Bitmap;
    t.bitmap.data = it->second->rgba.data(), t.bitmap.w = it->second->w, t.bitmap.h = it->second->h;
    return t;
}
bool TextureManager::RegisterImage(std::string_view key, const float *rgba, size_t w, size_t h) noexcept {
    if (!rgba || !w || !h) return false;
    auto data = std::make_unique<ImageData>();
    data->w = w, data->h = h, data->rgba.assign(rgba, rgba + w * h * 4);
    m_images[std::string(key)] = std::move(data);
    return true;
}
void TextureManager::Clear() noexcept { m_images.clear(); }

void Scene::LoadXmlObj(const xml::Object *o, void *dst) noexcept {
    if (o == nullptr || dst == nullptr) return;
    switch (o->tag) {
        case xml::ETag::_integrator: xml::LoadInt(o, "max_depth", static_cast<Integrator *>(dst)->max_depth, 1); break;
        case xml::ETag::_transform: xml::LoadTransform(o, dst); break;
        case xml::ETag::_film: {
            auto *film = static_cast<Film *>(dst);
            if (o->type != "hdrfilm") {
                Log::Warn("film only supports hdrfilm");
                return;
            }
            xml::LoadInt(o, "width", film->w, 768);
            xml::LoadInt(o, "height", film->h, 576);
            if (film->w <= 0 || film->h <= 0 || film->w > 65536 || film->h > 65536) { // uint32 casts and the aspect ratio follow
                Log::Warn("film size %dx%d is not renderable: using 768x576", (int)film->w, (int)film->h);
                film->w = 768, film->h = 576;
            }
        } break;
        case xml::ETag::_sensor: {
            auto *s = static_cast<Sensor *>(dst);
            if (o->type != "perspective") {
                Log::Warn("sensor only supports perspective");
                return;
            }
            xml::LoadFloat(o, "fov", s->fov, 90.f);
            xml::LoadFloat(o, "near_clip", s->near_clip, 0.01f);
            xml::LoadFloat(o, "far_clip", s->far_clip, 10000.f);
            LoadXmlObj(o->GetUniqueSubObject("film"), &s->film);
            const std::string axis = o->GetProperty("fov_axis");
            char fov_axis = 'x';
            if (axis == "y" || axis == "Y") fov_axis = 'y';
            else if (!axis.empty() && axis != "x" && axis != "X") Log::Warn("sensor fov_axis must be x or y");
            if (fov_axis == 'x') { // horizontal -> vertical field of view (scene.cpp:122-127)
                const float aspect = static_cast<float>(s->film.h) / static_cast<float>(s->film.w);
                const float radian = s->fov * 3.14159265358979323846f / 180.f * 0.5f;
                const float t = std::tan(radian) * aspect;
                s->fov = 2.f * std::atan(t) * 180.f / 3.14159265358979323846f;
            }
            LoadXmlObj(o->GetUniqueSubObject("transform"), &s->transform);
            // mitsuba -> Pupil handedness (a lookat transform has already been flipped once by the loader)
            for (int r = 0; r < 3; ++r) s->transform.matrix.re[r][0] *= -1, s->transform.matrix.re[r][2] *= -1;
        } break;
        case xml::ETag::_texture: {
            auto *tex = static_cast<util::Texture *>(dst);
            *tex = util::Texture{};
            if (o->type == "checkerboard") {
                tex->type = util::ETextureType::Checkerboard;
                xml::LoadFloat3(o, "color0", tex->patch1, { 0.4f });
                xml::LoadFloat3(o, "color1", tex->patch2, { 0.2f });
            } else if (o->type == "bitmap") { // scene.cpp:144-166
                const std::string value = o->GetProperty("filename");
                *tex = util::Singleton<TextureManager>::instance()->GetTexture(ResolveImagePath(scene_root_path, value));
                tex->bitmap.filter_mode = o->GetProperty("filter_type") == "bilinear" ? util::ETextureFilterMode::Linear : util::ETextureFilterMode::Point;
                const std::string wrap = o->GetProperty("wrap_mode");
                tex->bitmap.address_mode = wrap == "mirror" ? util::ETextureAddressMode::Mirror
                                           : wrap == "clamp" ? util::ETextureAddressMode::Clamp
                                                             : util::ETextureAddressMode::Wrap; // "repeat" and the default
            } else {
                Log::Warn("unknown texture type [%s]", o->type.c_str());
            }
            LoadXmlObj(o->GetUniqueSubObject("transform"), &tex->transform);
        } break;
        case xml::ETag::_bsdf: *static_cast<Material *>(dst) = LoadMaterialFromXml(o, this); break;
        case xml::ETag::_emitter: {
            auto *e = static_cast<Emitter *>(dst);
            if (o->type == "area") {
                e->type = EEmitterType::Area;
                xml::LoadTextureOrRGB(o, this, "radiance", e->radiance);
            } else if (o->type == "point") { // parsed, never used by the renderer
                e->type = EEmitterType::Point;
                xml::Load3Float(o, "position", e->position);
                xml::LoadFloat3(o, "intensity", e->color);
            } else if (o->type == "constant") {
                e->type = EEmitterType::ConstEnv;
                xml::LoadFloat3(o, "radiance", e->color);
            } else if (o->type == "envmap") { // scene.cpp:207-219
                e->type = EEmitterType::EnvMap;
                xml::LoadFloat(o, "scale", e->scale, 1.f);
                e->radiance = util::Singleton<TextureManager>::instance()->GetTexture(ResolveImagePath(scene_root_path, o->GetProperty("filename")));
                if (e->radiance.type != util::ETextureType::Bitmap) {
                    Log::Warn("envmap emitter: no image, emitter ignored");
                    e->type = EEmitterType::Unknown;
                }
                e->radiance.bitmap.filter_mode = util::ETextureFilterMode::Linear;
                e->radiance.bitmap.address_mode = util::ETextureAddressMode::Wrap;
                e->transform = util::Transform{};
                LoadXmlObj(o->GetUniqueSubObject("transform"), &e->transform);
            } else {
                Log::Warn("unknown emitter type [%s]", o->type.c_str());
            }
        } break;
        default: break;
    }
}
}// namespace resource
}// namespace Pupil
