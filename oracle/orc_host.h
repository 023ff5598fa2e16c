/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_types.h).
 *
 * Host-side precompute of the reference, restated: transforms, camera matrices, built-in
 * shapes, material precompute, emitter table + selection probabilities.  Backend independent
 * (pure fp32/fp64 host arithmetic).  Citations are relative to /root/reference/framework.
 *
 * PARITY UNPINNED for the pieces the reference delegates to DirectXMath (Windows SDK, not
 * vendored): XMMatrixPerspectiveFovRH, XMMatrixLookAtRH, XMMatrixInverse, XMMatrixMultiply.
 * Their published definitions are restated; inverses are taken in fp64 and rounded once.
 */
#ifndef ORC_HOST_H
#define ORC_HOST_H
#include "orc_types.h"
#include "orc_vec.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

// ---- 4x4 helpers (row-major, column vectors; util/type.h:73-111) -------------------------------
inline m44 mul44(const m44 &a, const m44 &b) {
    m44 r{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += a.e[i * 4 + k] * b.e[k * 4 + j];
            r.e[i * 4 + j] = s;
        }
    return r;
}
inline m44 transpose44(const m44 &a) {
    m44 r{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.e[i * 4 + j] = a.e[j * 4 + i];
    return r;
}
// general inverse, Gauss-Jordan with partial pivoting in fp64 (stands in for XMMatrixInverse)
inline m44 inverse44(const m44 &a) {
    double m[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            m[i][j] = a.e[i * 4 + j];
            m[i][4 + j] = i == j ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r)
            if (std::fabs(m[r][c]) > std::fabs(m[p][c])) p = r;
        if (p != c)
            for (int j = 0; j < 8; ++j) std::swap(m[p][j], m[c][j]);
        double d = m[c][c];
        if (d == 0.0) continue;
        for (int j = 0; j < 8; ++j) m[c][j] /= d;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            double f = m[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
        }
    }
    m44 r{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.e[i * 4 + j] = static_cast<float>(m[i][4 + j]);
    return r;
}
// util/transform.cpp:99-123
inline f3 transform_point(f3 p, const m44 &t) {
    const float *m = t.e;
    float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    return f3{ x / w, y / w, z / w };
}
inline f3 transform_normal(f3 n, const m44 &inv_t) {
    const float *m = inv_t.e;
    float x = m[0] * n.x + m[1] * n.y + m[2] * n.z;
    float y = m[4] * n.x + m[5] * n.y + m[6] * n.z;
    float z = m[8] * n.x + m[9] * n.y + m[10] * n.z;
    float len = sqrtf(x * x + y * y + z * z);
    return f3{ x / len, y / len, z / len };
}

// util/transform.cpp:7-84 — Rotate / Translate / Scale all PRE-multiply
inline void xf_rotate(m44 &mat, float ux, float uy, float uz, float angle) {
    float u_len = sqrtf(ux * ux + uy * uy + uz * uz);
    ux /= u_len, uy /= u_len, uz /= u_len;
    float theta = angle / 180.f * 3.14159265358979323846f;
    float a = cosf(0.5f * theta);
    float b = sinf(0.5f * theta) * ux;
    float c = sinf(0.5f * theta) * uy;
    float d = sinf(0.5f * theta) * uz;
    m44 r = identity44();
    r.e[0] = 1.f - 2.f * c * c - 2.f * d * d, r.e[1] = 2.f * b * c - 2.f * a * d, r.e[2] = 2.f * a * c + 2.f * b * d;
    r.e[4] = 2.f * b * c + 2.f * a * d, r.e[5] = 1.f - 2.f * b * b - 2.f * d * d, r.e[6] = 2.f * c * d - 2.f * a * b;
    r.e[8] = 2.f * b * d - 2.f * a * c, r.e[9] = 2.f * a * b + 2.f * c * d, r.e[10] = 1.f - 2.f * b * b - 2.f * c * c;
    mat = mul44(r, mat);
}
inline void xf_translate(m44 &mat, float x, float y, float z) {
    m44 t = identity44();
    t.e[3] = x, t.e[7] = y, t.e[11] = z;
    mat = mul44(t, mat);
}
inline void xf_scale(m44 &mat, float x, float y, float z) {
    m44 s = identity44();
    s.e[0] = x, s.e[5] = y, s.e[10] = z;
    mat = mul44(s, mat);
}
// util/transform.cpp:86-97: camera_to_world = transpose(inverse(XMMatrixLookAtRH(eye, focus, up))).
// XMMatrixLookAtRH(eye, focus, up) == XMMatrixLookToLH(eye, eye - focus, up):
//   R2 = normalize(eye - focus), R0 = normalize(cross(up, R2)), R1 = cross(R2, R0); the view
//   matrix has rows (R0, -R0.eye), (R1, -R1.eye), (R2, -R2.eye) in column-vector form.
inline m44 xf_lookat(f3 eye, f3 focus, f3 up) {
    f3 r2 = normalize(eye - focus);
    f3 r0 = normalize(cross(up, r2));
    f3 r1 = cross(r2, r0);
    m44 view = identity44();
    view.e[0] = r0.x, view.e[1] = r0.y, view.e[2] = r0.z, view.e[3] = -dot(r0, eye);
    view.e[4] = r1.x, view.e[5] = r1.y, view.e[6] = r1.z, view.e[7] = -dot(r1, eye);
    view.e[8] = r2.x, view.e[9] = r2.y, view.e[10] = r2.z, view.e[11] = -dot(r2, eye);
    return inverse44(view);
}
// Mitsuba (+X left, +Z view) <-> Pupil (+X right, -Z view): negate columns 0 and 2 of the 3x3
// (resource/scene.cpp:134-139 and resource/xml/util_loader.cpp:161-166)
inline void flip_handedness(m44 &m) {
    m.e[0] *= -1, m.e[4] *= -1, m.e[8] *= -1;
    m.e[2] *= -1, m.e[6] *= -1, m.e[10] *= -1;
}
// resource/xml/util_loader.cpp:128-191 (LoadTransform3D).  Order is always scale -> rotate ->
// translate whatever the XML order was; lookat wins over s/r/t; matrix wins over everything.
inline m44 resolve_transform(const orc_transform &t) {
    m44 m = identity44();
    switch (t.kind) {
        case ORC_XF_MATRIX16: std::memcpy(m.e, t.m, sizeof(float) * 16); break;
        case ORC_XF_MATRIX9:
            for (int i = 0, j = 0; j < 9;) { // 3x3 into the upper-left block, util_loader.cpp:136-141
                m.e[i] = t.m[j];
                ++i, ++j;
                if (j % 3 == 0) ++i;
            }
            break;
        case ORC_XF_LOOKAT:
            m = xf_lookat(f3{ t.origin[0], t.origin[1], t.origin[2] }, f3{ t.target[0], t.target[1], t.target[2] },
                          f3{ t.up[0], t.up[1], t.up[2] });
            flip_handedness(m);
            break;
        case ORC_XF_SRT:
            if (t.has_scale) xf_scale(m, t.scale[0], t.scale[1], t.scale[2]);
            if (t.has_rotate) xf_rotate(m, t.axis[0], t.axis[1], t.axis[2], t.angle);
            if (t.has_translate) xf_translate(m, t.translate[0], t.translate[1], t.translate[2]);
            break;
        default: break;
    }
    return m;
}

// ---- camera: resource/scene.cpp:96-139, world/world.cpp:111-120, util/camera.cpp:7-47,80-101 ----
struct Camera {
    m44 sample_to_camera, camera_to_world;
    float fov_y;
};
inline Camera make_camera(float fov, bool fov_axis_x, float near_clip, float far_clip, const orc_transform &to_world, int film_w, int film_h) {
    Camera cam{};
    if (fov_axis_x) { // scene.cpp:122-127
        float aspect = static_cast<float>(film_h) / static_cast<float>(film_w);
        float radian = fov * 3.14159265358979323846f / 180.f * 0.5f;
        float t = tanf(radian) * aspect;
        fov = 2.f * atanf(t) * 180.f / 3.14159265358979323846f;
    }
    cam.fov_y = fov;
    m44 c2w = resolve_transform(to_world);
    flip_handedness(c2w);        // scene.cpp:134-139 (a lookat transform is therefore flipped twice)
    cam.camera_to_world = c2w;   // Camera::SetWorldTransform keeps the matrix verbatim (camera.cpp:80-81)

    // XMMatrixPerspectiveFovRH(fov_y, aspect, zn, zf), row-vector convention
    float aspect_ratio = static_cast<float>(film_w) / film_h; // world.cpp:114
    float half = 0.5f * (fov / 180.f * 3.14159265358979323846f);
    float height = cosf(half) / sinf(half);
    float width = height / aspect_ratio;
    float range = far_clip / (near_clip - far_clip);
    m44 proj{};
    proj.e[0] = width, proj.e[5] = height, proj.e[10] = range, proj.e[11] = -1.f, proj.e[14] = range * near_clip;
    m44 tr = identity44(); // XMMatrixTranslation(1,1,0): row 3 = (1,1,0,1)
    tr.e[12] = 1.f, tr.e[13] = 1.f;
    m44 sc = identity44(); // XMMatrixScaling(.5,.5,1)
    sc.e[0] = 0.5f, sc.e[5] = 0.5f;
    cam.sample_to_camera = transpose44(inverse44(mul44(mul44(proj, tr), sc))); // camera.cpp:9-16
    return cam;
}

// ---- built-in shapes: resource/shape.cpp:21-68 ----------------------------------------------------
struct MeshData {
    std::vector<float> pos, nrm, uv; // 3,3,2 per vertex; nrm / uv may be empty (obj without them)
    std::vector<uint32_t> idx;       // 3 per face
};
inline MeshData rectangle_mesh() { // XY-range [-1,1]^2, +Z normal
    MeshData m;
    m.pos = { -1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0 };
    m.nrm = { 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1 };
    m.uv = { 0, 0, 1, 0, 1, 1, 0, 1 };
    m.idx = { 0, 1, 2, 0, 2, 3 };
    return m;
}
inline MeshData cube_mesh() { // [-1,1]^3, 6 faces x 4 vertices, face order -X -Z +X +Z +Y -Y
    MeshData m;
    m.pos = { -1, -1, -1, -1, -1, 1,  -1, 1,  1,  -1, 1,  -1, 1,  -1, -1, -1, -1, -1, -1, 1,  -1, 1,  1, -1,
              1,  -1, 1,  1,  -1, -1, 1,  1,  -1, 1,  1,  1,  -1, -1, 1,  1,  -1, 1,  1,  1,  1,  -1, 1, 1,
              -1, 1,  1,  1,  1,  1,  1,  1,  -1, -1, 1,  -1, -1, -1, -1, 1,  -1, -1, 1,  -1, 1,  -1, -1, 1 };
    const float fn[6][3] = { { -1, 0, 0 }, { 0, 0, -1 }, { 1, 0, 0 }, { 0, 0, 1 }, { 0, 1, 0 }, { 0, -1, 0 } };
    for (int f = 0; f < 6; ++f)
        for (int v = 0; v < 4; ++v) {
            m.nrm.insert(m.nrm.end(), { fn[f][0], fn[f][1], fn[f][2] });
        }
    for (int f = 0; f < 6; ++f) m.uv.insert(m.uv.end(), { 0, 0, 1, 0, 1, 1, 0, 1 });
    for (uint32_t f = 0; f < 6; ++f) m.idx.insert(m.idx.end(), { 4 * f, 4 * f + 1, 4 * f + 2, 4 * f, 4 * f + 2, 4 * f + 3 });
    return m;
}

// ---- textures / materials: render/material/optix_material.cpp:9-130 -----------------------------
inline f3 tex_pixel_average(const orc_texture &t) { // :9-36
    if (t.type == ORC_TEX_CHECKERBOARD) return f3{ t.a[0] + t.b[0], t.a[1] + t.b[1], t.a[2] + t.b[2] } * 0.5f;
    if (t.type == ORC_TEX_BITMAP) { // :19-30
        float r = 0.f, g = 0.f, b = 0.f;
        for (int i = 0, idx = 0; i < t.bitmap_h; ++i) {
            for (int j = 0; j < t.bitmap_w; ++j) {
                r += t.bitmap[idx++];
                g += t.bitmap[idx++];
                b += t.bitmap[idx++];
                idx++; // a
            }
        }
        const float inv = 1.0f / (1.f * t.bitmap_h * t.bitmap_w); // float3 / float, cuda/vec_math.h:425-428
        return f3{ r, g, b } * inv;
    }
    return f3{ t.a[0], t.a[1], t.a[2] };
}
inline float tex_max_weight(const orc_texture &t) { // world/emitter.cpp:73-101 (GetWeight)
    auto mx = [](float r, float g, float b) { return (r > g ? (r > b ? r : b) : (g > b ? g : b)); };
    if (t.type == ORC_TEX_CHECKERBOARD) return (mx(t.a[0], t.a[1], t.a[2]) + mx(t.b[0], t.b[1], t.b[2])) * 0.5f;
    if (t.type == ORC_TEX_BITMAP) { // :89-99 — the reference indexes (i * w + j) with i < w, j < h (its own transposition);
        float w = 0.f;             // DEFINED: indices past the image (w > h) are skipped instead of read
        const size_t bw = (size_t)t.bitmap_w, bh = (size_t)t.bitmap_h;
        for (size_t i = 0; i < bw; i++)
            for (size_t j = 0; j < bh; j++) {
                const size_t px = i * bw + j;
                if (px >= bw * bh) continue;
                w += mx(t.bitmap[px * 4 + 0], t.bitmap[px * 4 + 1], t.bitmap[px * 4 + 2]);
            }
        return w / (1.f * bw * bh);
    }
    return mx(t.a[0], t.a[1], t.a[2]);
}
inline float lum(f3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; } // optix/util.h:161-163

/* optix::material::Material after LoadMaterial: what the device sees.  Texture slots by type:
 *   t0 = the GetAlbedo() texture (optix_material.h:93-111): reflectance | specular_reflectance |
 *        diffuse_reflectance;  see mat_slots() */
struct DeviceMaterial {
    int type = ORC_MAT_UNKNOWN;
    bool twosided = false;
    float eta = 1.f;             // int_ior / ext_ior  (optix_material.cpp:61,90,106; dielectric.h:49)
    bool nonlinear = false;
    float int_fdr = 0.f;         // fresnel::DiffuseReflectance(1/eta) (:99,116)
    float specular_sampling_weight = 0.f; // Ys/(Ys+Yd) (:95-97,112-114)
    orc_texture alpha, eta_tex, k_tex, reflectance, specular_reflectance, specular_transmittance;
};
template<class B>
inline DeviceMaterial load_material(const orc_material &m) {
    DeviceMaterial d;
    d.type = m.type, d.twosided = m.twosided != 0;
    d.alpha = m.alpha, d.eta_tex = m.eta, d.k_tex = m.k, d.reflectance = m.reflectance;
    d.specular_reflectance = m.specular_reflectance, d.specular_transmittance = m.specular_transmittance;
    d.nonlinear = m.nonlinear != 0;
    switch (m.type) {
        case ORC_MAT_DIELECTRIC:
        case ORC_MAT_ROUGH_DIELECTRIC: d.eta = m.int_ior / m.ext_ior; break;
        case ORC_MAT_PLASTIC:
        case ORC_MAT_ROUGH_PLASTIC: {
            d.eta = m.int_ior / m.ext_ior;
            float diffuse_luminance = lum(tex_pixel_average(m.reflectance));
            float specular_luminance = lum(tex_pixel_average(m.specular_reflectance));
            d.specular_sampling_weight = specular_luminance / (specular_luminance + diffuse_luminance);
            d.int_fdr = B::fresnel_diffuse(1.f / d.eta);
        } break;
        default: break;
    }
    return d;
}
// Material::GetLocalBsdf, optix_material.h:117-130 + each <Bsdf>::GetLocal
template<class B>
inline orc_local_bsdf get_local_bsdf(const DeviceMaterial &m, f2 uv) {
    orc_local_bsdf b{};
    auto put = [](float *dst, f3 v) { dst[0] = v.x, dst[1] = v.y, dst[2] = v.z; };
    b.type = m.type;
    b.eta = m.eta, b.int_fdr = m.int_fdr, b.specular_sampling_weight = m.specular_sampling_weight, b.nonlinear = m.nonlinear;
    switch (m.type) {
        case ORC_MAT_DIFFUSE: put(b.reflectance, B::tex_sample(m.reflectance, uv)); break;
        case ORC_MAT_DIELECTRIC:
            put(b.specular_reflectance, B::tex_sample(m.specular_reflectance, uv));
            put(b.specular_transmittance, B::tex_sample(m.specular_transmittance, uv));
            break;
        case ORC_MAT_ROUGH_DIELECTRIC:
            b.alpha = B::tex_sample(m.alpha, uv).x;
            put(b.specular_reflectance, B::tex_sample(m.specular_reflectance, uv));
            put(b.specular_transmittance, B::tex_sample(m.specular_transmittance, uv));
            break;
        case ORC_MAT_CONDUCTOR:
            put(b.eta3, B::tex_sample(m.eta_tex, uv)), put(b.k3, B::tex_sample(m.k_tex, uv));
            put(b.specular_reflectance, B::tex_sample(m.specular_reflectance, uv));
            break;
        case ORC_MAT_ROUGH_CONDUCTOR:
            b.alpha = B::tex_sample(m.alpha, uv).x;
            put(b.eta3, B::tex_sample(m.eta_tex, uv)), put(b.k3, B::tex_sample(m.k_tex, uv));
            put(b.specular_reflectance, B::tex_sample(m.specular_reflectance, uv));
            break;
        case ORC_MAT_PLASTIC:
            put(b.reflectance, B::tex_sample(m.reflectance, uv));
            put(b.specular_reflectance, B::tex_sample(m.specular_reflectance, uv));
            break;
        case ORC_MAT_ROUGH_PLASTIC:
            b.alpha = B::tex_sample(m.alpha, uv).x;
            put(b.reflectance, B::tex_sample(m.reflectance, uv));
            put(b.specular_reflectance, B::tex_sample(m.specular_reflectance, uv));
            break;
        default: break;
    }
    return b;
}
// LocalBsdf::GetAlbedo, optix_material.h:93-111
inline f3 local_albedo(const orc_local_bsdf &b) {
    switch (b.type) {
        case ORC_MAT_DIFFUSE:
        case ORC_MAT_PLASTIC:
        case ORC_MAT_ROUGH_PLASTIC: return f3{ b.reflectance[0], b.reflectance[1], b.reflectance[2] };
        case ORC_MAT_DIELECTRIC:
        case ORC_MAT_ROUGH_DIELECTRIC:
        case ORC_MAT_CONDUCTOR:
        case ORC_MAT_ROUGH_CONDUCTOR: return f3{ b.specular_reflectance[0], b.specular_reflectance[1], b.specular_reflectance[2] };
        default: return f3{ 0.f, 0.f, 0.f };
    }
}

// ---- emitter table: world/emitter.cpp:169-337 ------------------------------------------------------
inline void add_mesh_area_emitters(std::vector<orc_emitter> &out, const MeshData &mesh, const m44 &xf, const orc_texture &radiance) { // :169-222
    m44 normal_transform = transpose44(inverse44(xf));
    float select_weight = tex_max_weight(radiance);
    size_t nf = mesh.idx.size() / 3;
    for (size_t i = 0; i < nf; ++i) {
        orc_emitter e{};
        e.type = ORC_EMIT_TRI;
        e.radiance = radiance;
        for (int k = 0; k < 3; ++k) {
            uint32_t vi = mesh.idx[i * 3 + k];
            f3 p = transform_point(f3{ mesh.pos[vi * 3], mesh.pos[vi * 3 + 1], mesh.pos[vi * 3 + 2] }, xf);
            e.pos[k][0] = p.x, e.pos[k][1] = p.y, e.pos[k][2] = p.z;
            if (!mesh.uv.empty()) e.uv[k][0] = mesh.uv[vi * 2], e.uv[k][1] = mesh.uv[vi * 2 + 1];
        }
        f3 p0{ e.pos[0][0], e.pos[0][1], e.pos[0][2] }, p1{ e.pos[1][0], e.pos[1][1], e.pos[1][2] }, p2{ e.pos[2][0], e.pos[2][1], e.pos[2][2] };
        for (int k = 0; k < 3; ++k) {
            uint32_t vi = mesh.idx[i * 3 + k];
            f3 n;
            if (!mesh.nrm.empty()) n = transform_normal(f3{ mesh.nrm[vi * 3], mesh.nrm[vi * 3 + 1], mesh.nrm[vi * 3 + 2] }, normal_transform);
            else n = normalize(cross(p1 - p0, p2 - p0)); // DEFINED: the reference dereferences a null normals array here
            e.nrm[k][0] = n.x, e.nrm[k][1] = n.y, e.nrm[k][2] = n.z;
        }
        e.area = length(cross(p1 - p0, p2 - p0)) * 0.5f;
        e.weight = select_weight * e.area;
        out.push_back(e);
    }
}
inline void add_sphere_area_emitter(std::vector<orc_emitter> &out, const m44 &xf, const orc_texture &radiance) { // :224-243
    orc_emitter e{};
    e.type = ORC_EMIT_SPHERE;
    e.radiance = radiance;
    f3 o = transform_point(f3{ 0.f, 0.f, 0.f }, xf);
    f3 p = transform_point(f3{ 1.f, 0.f, 0.f }, xf);
    e.center[0] = o.x, e.center[1] = o.y, e.center[2] = o.z;
    e.radius = length(o - p);
    e.area = 4 * 3.14159265358979323846f * e.radius * e.radius;
    e.weight = tex_max_weight(radiance) * e.area;
    out.push_back(e);
}
// BuildEnvMapCdfTable, world/emitter.cpp:107-149.  Output layout (padded, see orc_types.h): row_cdf[h + 2],
// row_weight[h + 1], col_cdf[(w + 1) * (h + 1)]; returns the normalization.
struct EnvTables {
    std::vector<float> row_cdf, row_weight, col_cdf;
    float normalization = 0.f;
};
inline EnvTables build_env_tables(const float *rgba, size_t w, size_t h) {
    EnvTables t;
    t.col_cdf.resize((w + 1) * h);
    t.row_cdf.resize(h + 1);
    t.row_weight.resize(h);
    size_t col_index = 0, row_index = 0;
    float row_sum = 0.f;
    t.row_cdf[row_index++] = 0.f;
    for (auto y = 0u; y < h; ++y) {
        float col_sum = 0.f;
        t.col_cdf[col_index++] = 0.f;
        for (auto x = 0u; x < w; ++x) {
            auto pixel_index = y * w + x;
            auto r = rgba[pixel_index * 4 + 0];
            auto g = rgba[pixel_index * 4 + 1];
            auto b = rgba[pixel_index * 4 + 2];
            col_sum += lum(f3{ r, g, b });
            t.col_cdf[col_index++] = col_sum;
        }
        for (auto x = 1u; x < w; ++x) t.col_cdf[col_index - x - 1] /= col_sum;
        t.col_cdf[col_index - 1] = 1.f;
        float weight = std::sin((y + 0.5f) * 3.14159265358979323846f / h);
        t.row_weight[y] = weight;
        row_sum += col_sum * weight;
        t.row_cdf[row_index++] = row_sum;
    }
    for (auto y = 1u; y < h; ++y) t.row_cdf[row_index - y - 1] /= row_sum;
    t.row_cdf[row_index - 1] = 1.f;
    t.normalization = 1.f / (row_sum * (2.f * 3.14159265358979323846f / w) * (3.14159265358979323846f / h));
    // pad: one more row so that row_index == h (reachable in the reference, which then reads out of bounds) is defined
    t.row_cdf.push_back(1.f);
    t.row_weight.push_back(t.row_weight[h - 1]);
    t.col_cdf.insert(t.col_cdf.end(), t.col_cdf.end() - (w + 1), t.col_cdf.end());
    return t;
}

inline void compute_select_probability(std::vector<orc_emitter> &areas, orc_emitter *env) { // :321-337
    float area_weight_sum = 0.f;
    for (auto &e : areas) area_weight_sum += e.weight;
    if (!areas.empty())
        for (auto &e : areas) e.select_probability = e.weight / area_weight_sum * areas.size();
    size_t emitter_num = (env ? 1 : 0) + areas.size();
    for (auto &e : areas) e.select_probability = e.select_probability / emitter_num;
    if (env) env->select_probability = env->weight / emitter_num;
}
}// namespace orc
#endif
