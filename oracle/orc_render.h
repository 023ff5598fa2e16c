/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_types.h).
 *
 * Scene container, CPU ray tracer and the restated pt-with-MIS integrator, templated on the
 * arithmetic backend (orc_backend_port.h | orc_backend_ref.h).
 *
 *   raygen loop        example/path_tracer/main.cu:38-197
 *   miss               example/path_tracer/main.cu:199-215
 *   closest hit        example/path_tracer/main.cu:220-234, framework/render/geometry.h:60-100,176-180
 *   shadow ray         framework/render/emitter.h:91-100
 *   world assembly     framework/world/world.cpp:108-146, world/render_object.cpp:18-65,
 *                      example/path_tracer/pt_pass.cpp:178-193 (emitter_index_offset)
 *
 * PARITY UNPINNED for ray/primitive intersection: the reference delegates it to OptiX 7.5
 * (closed source).  The contract restated here is geometric: nearest hit in (tmin, tmax) of the
 * exact primitive set, fp32 Moeller-Trumbore on world-space triangles and an analytic unit
 * sphere in object space; exact-t ties resolve to the lowest (instance, primitive).
 * The intersection arithmetic is written with EXPLICIT fused multiply-adds (ix_* below): fmaf is
 * exactly rounded on every IEEE machine, so the CUDA kernels, which spell out the same sequence with
 * __fmaf_rn / __fmul_rn, produce bit-identical t, u, v.
 */
#ifndef ORC_RENDER_H
#define ORC_RENDER_H
#include "orc_host.h"
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <memory>
#include <thread>

namespace orc {

struct Instance {
    int shape_type = 0; // ORC_SHAPE_*
    int mesh = -1;      // index into Scene::meshes (not for spheres)
    m44 xf, inv;        // object->world and its inverse
    DeviceMaterial mat;
    int emitter_offset = -1; // HitGroupData::emitter_index_offset (type.h:35-39)
    bool flip_tex = false;
};
struct WorldTri { // pre-transformed triangle used only for intersection
    f3 v0, e1, e2;
    int inst, prim;
};
struct WorldSphere {
    int inst;
    f3 bmin, bmax;
};
struct BvhNode {
    f3 bmin, bmax;
    int left, right; // internal: children; leaf: left = -1 - first, right = count
};

template<class B>
struct Scene {
    // ---- resource level (what the XML says) ----
    int max_depth = 1; // resource/scene.h:18
    Camera cam{};
    int film_w = 768, film_h = 576;
    std::vector<MeshData> meshes;
    std::vector<Instance> instances;
    // cube / rectangle / sphere are process-wide singletons in the reference, so flip_normals is
    // last-writer-wins across ALL instances of that built-in (resource/shape.cpp:91,105,124)
    bool flip_rect = false, flip_cube = false, flip_sphere = false;
    int rect_mesh_id = -1, cube_mesh_id = -1;
    std::vector<bool> mesh_flip; // per obj mesh (also last-writer-wins per file; one entry per add here)
    // ---- world level ----
    std::vector<orc_emitter> areas;
    bool has_env = false;
    orc_emitter env{};
    EnvTables env_tables;          // owns the tables env.row_cdf / row_weight / col_cdf point at
    std::vector<float> env_bitmap; // owns the texels env.radiance.bitmap points at
    std::vector<WorldTri> tris;
    std::vector<WorldSphere> spheres;
    std::vector<BvhNode> bvh;
    std::vector<int> bvh_prims; // >=0: tri index; <0: sphere index = -1-v
    bool finalized = false;

    bool mesh_flips(const Instance &in) const {
        if (in.shape_type == ORC_SHAPE_RECTANGLE) return flip_rect;
        if (in.shape_type == ORC_SHAPE_CUBE) return flip_cube;
        if (in.shape_type == ORC_SHAPE_SPHERE) return flip_sphere;
        return mesh_flip[in.mesh];
    }

    int add_shape(int shape_type, const orc_transform *to_world, const orc_material *mat, int is_emitter, const orc_texture *radiance,
                  int flip_normals, const float *center, float radius, int flip_tex_coords, uint32_t nv, uint32_t nf, const float *pos,
                  const float *nrm, const float *uv, const uint32_t *idx) {
        Instance in;
        in.shape_type = shape_type;
        m44 xf = to_world ? resolve_transform(*to_world) : identity44();
        switch (shape_type) {
            case ORC_SHAPE_RECTANGLE:
                if (rect_mesh_id < 0) rect_mesh_id = (int)meshes.size(), meshes.push_back(rectangle_mesh()), mesh_flip.push_back(false);
                in.mesh = rect_mesh_id, flip_rect = flip_normals != 0;
                break;
            case ORC_SHAPE_CUBE:
                if (cube_mesh_id < 0) cube_mesh_id = (int)meshes.size(), meshes.push_back(cube_mesh()), mesh_flip.push_back(false);
                in.mesh = cube_mesh_id, flip_cube = flip_normals != 0;
                break;
            case ORC_SHAPE_SPHERE: { // shape.cpp:113-133 + :245-246: to_world * (T(center) * S(radius))
                flip_sphere = flip_normals != 0;
                m44 local = identity44();
                xf_scale(local, radius, radius, radius);
                xf_translate(local, center ? center[0] : 0.f, center ? center[1] : 0.f, center ? center[2] : 0.f);
                xf = mul44(xf, local);
            } break;
            case ORC_SHAPE_OBJ: {
                MeshData m;
                m.pos.assign(pos, pos + 3 * (size_t)nv);
                if (nrm) m.nrm.assign(nrm, nrm + 3 * (size_t)nv);
                if (uv) m.uv.assign(uv, uv + 2 * (size_t)nv);
                m.idx.assign(idx, idx + 3 * (size_t)nf);
                in.mesh = (int)meshes.size();
                meshes.push_back(std::move(m)), mesh_flip.push_back(flip_normals != 0);
                in.flip_tex = flip_tex_coords != 0;
            } break;
            default: return -1;
        }
        in.xf = xf, in.inv = inverse44(xf);
        if (mat) in.mat = load_material<B>(*mat);
        if (is_emitter && radiance) { // world.cpp:133-136, emitter.cpp:245-263
            in.emitter_offset = (int)areas.size();
            if (shape_type == ORC_SHAPE_SPHERE) add_sphere_area_emitter(areas, xf, *radiance);
            else add_mesh_area_emitters(areas, meshes[in.mesh], xf, *radiance);
        }
        instances.push_back(in);
        finalized = false;
        return (int)instances.size() - 1;
    }

    void set_env_const(float r, float g, float b) { // emitter.cpp:283-292
        has_env = true;
        env = orc_emitter{};
        env.type = ORC_EMIT_CONST_ENV;
        env.radiance.a[0] = r, env.radiance.a[1] = g, env.radiance.a[2] = b;
        env.weight = 1.f;
        finalized = false;
    }

    void set_env_map(const float *rgba, uint32_t w, uint32_t h, float scale, const orc_transform &xf) { // emitter.cpp:293-312, scene.cpp:207-219
        has_env = true;
        env = orc_emitter{};
        env.type = ORC_EMIT_ENV_MAP;
        env_bitmap.assign(rgba, rgba + (size_t)w * h * 4);
        env.radiance.type = ORC_TEX_BITMAP;
        env.radiance.bitmap = env_bitmap.data(), env.radiance.bitmap_w = (int)w, env.radiance.bitmap_h = (int)h;
        env.radiance.address_mode = 0 /* Wrap */, env.radiance.filter_mode = 1 /* Linear */;
        for (int i = 0; i < 16; ++i) env.radiance.to_uv[i] = (i % 5 == 0) ? 1.f : 0.f;
        env.scale = scale;
        env.weight = 1.f;
        const m44 m = resolve_transform(xf), inv = inverse44(m);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) env.to_world[r * 3 + c] = m.e[r * 4 + c], env.to_local[r * 3 + c] = inv.e[r * 4 + c];
        env_tables = build_env_tables(env_bitmap.data(), w, h);
        env.normalization = env_tables.normalization;
        env.map_w = w, env.map_h = h;
        env.row_cdf = env_tables.row_cdf.data(), env.row_weight = env_tables.row_weight.data(), env.col_cdf = env_tables.col_cdf.data();
        finalized = false;
    }

    // ---------- world assembly: emitter probabilities, world-space primitives, CPU BVH ----------
    void finalize() {
        compute_select_probability(areas, has_env ? &env : nullptr);
        tris.clear(), spheres.clear();
        for (size_t ii = 0; ii < instances.size(); ++ii) {
            const Instance &in = instances[ii];
            if (in.shape_type == ORC_SHAPE_SPHERE) {
                WorldSphere s;
                s.inst = (int)ii;
                const float *m = in.xf.e; // exact bounds of an affinely transformed unit sphere
                f3 c{ m[3], m[7], m[11] };
                f3 r{ sqrtf(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]), sqrtf(m[4] * m[4] + m[5] * m[5] + m[6] * m[6]),
                      sqrtf(m[8] * m[8] + m[9] * m[9] + m[10] * m[10]) };
                s.bmin = c - r, s.bmax = c + r;
                spheres.push_back(s);
                continue;
            }
            const MeshData &md = meshes[in.mesh];
            size_t nf = md.idx.size() / 3;
            for (size_t f = 0; f < nf; ++f) {
                f3 p[3];
                for (int k = 0; k < 3; ++k) {
                    uint32_t vi = md.idx[f * 3 + k];
                    p[k] = ix_point(f3{ md.pos[vi * 3], md.pos[vi * 3 + 1], md.pos[vi * 3 + 2] }, in.xf.e);
                }
                tris.push_back(WorldTri{ p[0], p[1] - p[0], p[2] - p[0], (int)ii, (int)f });
            }
        }
        build_bvh();
        finalized = true;
    }
    static f3 obj_to_world_point(f3 p, const m44 &t) { // 3x4 affine (OptiX instance transform)
        const float *m = t.e;
        return f3{ m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
                   m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11] };
    }
    static f3 obj_to_world_vector(f3 p, const m44 &t) {
        const float *m = t.e;
        return f3{ m[0] * p.x + m[1] * p.y + m[2] * p.z, m[4] * p.x + m[5] * p.y + m[6] * p.z, m[8] * p.x + m[9] * p.y + m[10] * p.z };
    }
    static f3 normal_obj_to_world(f3 n, const m44 &inv) { // n_w = (M^-1)^T n
        const float *m = inv.e;
        return f3{ m[0] * n.x + m[4] * n.y + m[8] * n.z, m[1] * n.x + m[5] * n.y + m[9] * n.z, m[2] * n.x + m[6] * n.y + m[10] * n.z };
    }

    // ---------- intersection ----------
    // fixed-rounding building blocks (same sequence as csrc/traverse.cuh)
    static float ix_dot(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
    static f3 ix_cross(f3 a, f3 b) { return f3{ fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)) }; }
    static f3 ix_point(f3 p, const float *m) { // rows of a 3x4 affine matrix
        return f3{ fmaf(m[2], p.z, fmaf(m[1], p.y, fmaf(m[0], p.x, m[3]))), fmaf(m[6], p.z, fmaf(m[5], p.y, fmaf(m[4], p.x, m[7]))),
                   fmaf(m[10], p.z, fmaf(m[9], p.y, fmaf(m[8], p.x, m[11]))) };
    }
    static f3 ix_vector(f3 v, const float *m) {
        return f3{ fmaf(m[2], v.z, fmaf(m[1], v.y, m[0] * v.x)), fmaf(m[6], v.z, fmaf(m[5], v.y, m[4] * v.x)), fmaf(m[10], v.z, fmaf(m[9], v.y, m[8] * v.x)) };
    }
    static bool hit_tri(const WorldTri &t, f3 o, f3 d, float tmin, float tmax, float &th, float &uh, float &vh) {
        f3 pvec = ix_cross(d, t.e2);
        float det = ix_dot(t.e1, pvec);
        if (det == 0.f) return false;
        float inv = 1.f / det;
        f3 tvec = o - t.v0;
        float u = ix_dot(tvec, pvec) * inv;
        if (u < 0.f || u > 1.f) return false;
        f3 qvec = ix_cross(tvec, t.e1);
        float v = ix_dot(d, qvec) * inv;
        if (v < 0.f || u + v > 1.f) return false;
        float tt = ix_dot(t.e2, qvec) * inv;
        if (!(tt > tmin && tt < tmax)) return false;
        th = tt, uh = u, vh = v;
        return true;
    }
    // unit sphere at the object-space origin; the ray is moved to object space with the inverse
    // instance transform and NOT renormalised, so t is shared between the two spaces.
    bool hit_sphere(const WorldSphere &s, f3 o, f3 d, float tmin, float tmax, float &th) const {
        const Instance &in = instances[s.inst];
        f3 oo = ix_point(o, in.inv.e), dd = ix_vector(d, in.inv.e);
        float a = ix_dot(dd, dd), b = ix_dot(oo, dd), c = ix_dot(oo, oo) - 1.f;
        float disc = fmaf(b, b, -(a * c));
        if (!(disc >= 0.f) || a == 0.f) return false;
        float sq = sqrtf(disc);
        float t0 = (-b - sq) / a, t1 = (-b + sq) / a;
        if (t0 > tmin && t0 < tmax) {
            th = t0;
            return true;
        }
        if (t1 > tmin && t1 < tmax) {
            th = t1;
            return true;
        }
        return false;
    }
    static bool better(float t, int inst, int prim, const orc_hit &h) {
        return h.inst < 0 || t < h.t || (t == h.t && (inst < h.inst || (inst == h.inst && prim < h.prim)));
    }
    void test_prim(int ref, f3 o, f3 d, float tmin, float tmax, orc_hit &h, uint64_t *n_tests) const {
        if (n_tests) ++*n_tests;
        float t, u = 0.f, v = 0.f;
        if (ref >= 0) {
            const WorldTri &tr = tris[ref];
            if (hit_tri(tr, o, d, tmin, tmax, t, u, v) && better(t, tr.inst, tr.prim, h)) h = orc_hit{ t, u, v, tr.inst, tr.prim };
        } else {
            const WorldSphere &s = spheres[-1 - ref];
            if (hit_sphere(s, o, d, tmin, tmax, t) && better(t, s.inst, 0, h)) h = orc_hit{ t, 0.f, 0.f, s.inst, 0 };
        }
    }
    // ---------- meshes behind an instance (IAS -> GAS, framework/world/ias_manager.cpp:29-114, gas_manager.cpp:10) ----------
    // OptiX moves the ray into the instance's object space with the inverse of its 3x4 transform and intersects the shared,
    // untransformed triangles there; t keeps its meaning because the direction is not renormalised.  This restates exactly
    // that for the instances flagged in `objspace`, with the arithmetic the device library spells out: the inverse is the
    // fp64 adjugate rounded once to fp32 (pupiloptixlab_b200/csrc/pb2_api.cu, invert_affine), point and vector transforms and
    // the triangle test are the ix_* sequences above.  Everything else is tested in world space as in trace_brute.
    static void invert_affine_f64(const float *m, float *out) {
        const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
        const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
        const double id = 1.0 / det;
        const double r[9] = { (e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id, (a * i - c * g) * id,
                              (c * d - a * f) * id, (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id };
        const double tx = m[3], ty = m[7], tz = m[11];
        for (int k = 0; k < 3; ++k) {
            out[k * 4 + 0] = (float)r[k * 3], out[k * 4 + 1] = (float)r[k * 3 + 1], out[k * 4 + 2] = (float)r[k * 3 + 2];
            out[k * 4 + 3] = (float)-(r[k * 3] * tx + r[k * 3 + 1] * ty + r[k * 3 + 2] * tz);
        }
    }
    orc_hit trace_brute_objspace(f3 o, f3 d, float tmin, float tmax, const uint8_t *objspace) const {
        orc_hit h{ 0.f, 0.f, 0.f, -1, -1 };
        for (int i = 0; i < (int)tris.size(); ++i)
            if (!objspace[tris[i].inst]) test_prim(i, o, d, tmin, tmax, h, nullptr);
        for (int i = 0; i < (int)spheres.size(); ++i) test_prim(-1 - i, o, d, tmin, tmax, h, nullptr);
        for (size_t ii = 0; ii < instances.size(); ++ii) {
            const Instance &in = instances[ii];
            if (!objspace[ii] || in.shape_type == ORC_SHAPE_SPHERE) continue;
            float inv[12];
            invert_affine_f64(in.xf.e, inv);
            const f3 oo = ix_point(o, inv), dd = ix_vector(d, inv);
            const MeshData &md = meshes[in.mesh];
            const size_t nf = md.idx.size() / 3;
            for (size_t f = 0; f < nf; ++f) {
                f3 p[3];
                for (int k = 0; k < 3; ++k) {
                    const uint32_t vi = md.idx[f * 3 + k];
                    p[k] = f3{ md.pos[vi * 3], md.pos[vi * 3 + 1], md.pos[vi * 3 + 2] };
                }
                const WorldTri ot{ p[0], p[1] - p[0], p[2] - p[0], (int)ii, (int)f };
                float t, u, v;
                if (hit_tri(ot, oo, dd, tmin, tmax, t, u, v) && better(t, ot.inst, ot.prim, h)) h = orc_hit{ t, u, v, ot.inst, ot.prim };
            }
        }
        return h;
    }
    orc_hit trace_brute(f3 o, f3 d, float tmin, float tmax) const {
        orc_hit h{ 0.f, 0.f, 0.f, -1, -1 };
        for (int i = 0; i < (int)tris.size(); ++i) test_prim(i, o, d, tmin, tmax, h, nullptr);
        for (int i = 0; i < (int)spheres.size(); ++i) test_prim(-1 - i, o, d, tmin, tmax, h, nullptr);
        return h;
    }

    // ---------- a plain binary BVH (spatial-median split) so mid-size scenes stay fast ----------
    void prim_bounds(int ref, f3 &lo, f3 &hi) const {
        if (ref >= 0) {
            const WorldTri &t = tris[ref];
            f3 a = t.v0, b = t.v0 + t.e1, c = t.v0 + t.e2;
            lo = f3{ std::min({ a.x, b.x, c.x }), std::min({ a.y, b.y, c.y }), std::min({ a.z, b.z, c.z }) };
            hi = f3{ std::max({ a.x, b.x, c.x }), std::max({ a.y, b.y, c.y }), std::max({ a.z, b.z, c.z }) };
        } else {
            lo = spheres[-1 - ref].bmin, hi = spheres[-1 - ref].bmax;
        }
    }
    void build_bvh() {
        bvh.clear(), bvh_prims.clear();
        for (int i = 0; i < (int)tris.size(); ++i) bvh_prims.push_back(i);
        for (int i = 0; i < (int)spheres.size(); ++i) bvh_prims.push_back(-1 - i);
        if (bvh_prims.empty()) return;
        std::vector<f3> lo(bvh_prims.size()), hi(bvh_prims.size());
        struct Item {
            int ref;
            f3 lo, hi;
        };
        std::vector<Item> items(bvh_prims.size());
        for (size_t i = 0; i < items.size(); ++i) {
            items[i].ref = bvh_prims[i];
            prim_bounds(items[i].ref, items[i].lo, items[i].hi);
        }
        bvh.reserve(items.size() * 2);
        struct Task {
            int node, first, count;
        };
        std::vector<Task> stack;
        bvh.push_back(BvhNode{});
        stack.push_back({ 0, 0, (int)items.size() });
        while (!stack.empty()) {
            Task t = stack.back();
            stack.pop_back();
            f3 bl = mk3(INFINITY), bh = mk3(-INFINITY), cl = mk3(INFINITY), ch = mk3(-INFINITY);
            for (int i = t.first; i < t.first + t.count; ++i) {
                const Item &it = items[i];
                bl = f3{ std::min(bl.x, it.lo.x), std::min(bl.y, it.lo.y), std::min(bl.z, it.lo.z) };
                bh = f3{ std::max(bh.x, it.hi.x), std::max(bh.y, it.hi.y), std::max(bh.z, it.hi.z) };
                f3 c = (it.lo + it.hi) * 0.5f;
                cl = f3{ std::min(cl.x, c.x), std::min(cl.y, c.y), std::min(cl.z, c.z) };
                ch = f3{ std::max(ch.x, c.x), std::max(ch.y, c.y), std::max(ch.z, c.z) };
            }
            bvh[t.node].bmin = bl, bvh[t.node].bmax = bh;
            f3 ext = ch - cl;
            int axis = ext.x >= ext.y ? (ext.x >= ext.z ? 0 : 2) : (ext.y >= ext.z ? 1 : 2);
            float extent = axis == 0 ? ext.x : axis == 1 ? ext.y : ext.z;
            if (t.count <= 4 || !(extent > 0.f)) {
                bvh[t.node].left = -1 - t.first, bvh[t.node].right = t.count;
                continue;
            }
            float mid = 0.5f * ((axis == 0 ? cl.x : axis == 1 ? cl.y : cl.z) + (axis == 0 ? ch.x : axis == 1 ? ch.y : ch.z));
            auto key = [axis](const Item &it) {
                return axis == 0 ? 0.5f * (it.lo.x + it.hi.x) : axis == 1 ? 0.5f * (it.lo.y + it.hi.y) : 0.5f * (it.lo.z + it.hi.z);
            };
            int m = (int)(std::partition(items.begin() + t.first, items.begin() + t.first + t.count, [&](const Item &it) { return key(it) < mid; }) -
                          items.begin());
            if (m == t.first || m == t.first + t.count) m = t.first + t.count / 2;
            int l = (int)bvh.size();
            bvh.push_back(BvhNode{}), bvh.push_back(BvhNode{});
            bvh[t.node].left = l, bvh[t.node].right = l + 1;
            stack.push_back({ l, t.first, m - t.first });
            stack.push_back({ l + 1, m, t.first + t.count - m });
        }
        for (size_t i = 0; i < items.size(); ++i) bvh_prims[i] = items[i].ref;
    }
    static bool hit_box(const BvhNode &n, f3 o, f3 inv_d, float tmin, float tmax) {
        // NaNs (0 * inf when the origin lies on a slab plane of an axis-parallel ray) are ignored by
        // the min/max chains below, which keeps the test conservative.
        float lo = -INFINITY, hi = INFINITY;
        auto slab = [&](float bmin, float bmax, float oo, float id) {
            float t0 = (bmin - oo) * id, t1 = (bmax - oo) * id;
            float a = t0 < t1 ? t0 : t1, b = t0 < t1 ? t1 : t0;
            if (a > lo) lo = a;
            if (b < hi) hi = b;
        };
        slab(n.bmin.x, n.bmax.x, o.x, inv_d.x), slab(n.bmin.y, n.bmax.y, o.y, inv_d.y), slab(n.bmin.z, n.bmax.z, o.z, inv_d.z);
        // widen by a few ulps so box culling can never drop a hit the brute-force loop finds
        return lo <= hi * 1.0000004f + 1e-30f && hi * 1.0000004f >= tmin && lo <= tmax;
    }
    orc_hit trace_closest(f3 o, f3 d, float tmin, float tmax, uint64_t *n_tests = nullptr) const {
        orc_hit h{ 0.f, 0.f, 0.f, -1, -1 };
        if (bvh.empty()) return h;
        if (bvh_prims.size() <= 64) { // tiny scenes (Cornell box): brute force is faster and identical
            for (int ref : bvh_prims) test_prim(ref, o, d, tmin, tmax, h, n_tests);
            return h;
        }
        f3 inv_d{ 1.f / d.x, 1.f / d.y, 1.f / d.z };
        int stack[128], sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const BvhNode &n = bvh[stack[--sp]];
            float far = h.inst >= 0 ? h.t : tmax;
            if (!hit_box(n, o, inv_d, tmin, far)) continue;
            if (n.left < 0) {
                int first = -1 - n.left;
                for (int i = 0; i < n.right; ++i) test_prim(bvh_prims[first + i], o, d, tmin, tmax, h, n_tests);
            } else {
                stack[sp++] = n.left, stack[sp++] = n.right;
            }
        }
        return h;
    }
    bool trace_any(f3 o, f3 d, float tmin, float tmax) const {
        if (bvh.empty()) return false;
        orc_hit h{ 0.f, 0.f, 0.f, -1, -1 };
        if (bvh_prims.size() <= 64) {
            for (int ref : bvh_prims) {
                test_prim(ref, o, d, tmin, tmax, h, nullptr);
                if (h.inst >= 0) return true;
            }
            return false;
        }
        f3 inv_d{ 1.f / d.x, 1.f / d.y, 1.f / d.z };
        int stack[128], sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const BvhNode &n = bvh[stack[--sp]];
            if (!hit_box(n, o, inv_d, tmin, tmax)) continue;
            if (n.left < 0) {
                int first = -1 - n.left;
                for (int i = 0; i < n.right; ++i) {
                    test_prim(bvh_prims[first + i], o, d, tmin, tmax, h, nullptr);
                    if (h.inst >= 0) return true;
                }
            } else {
                stack[sp++] = n.left, stack[sp++] = n.right;
            }
        }
        return false;
    }

    // ---------- closest-hit program: LocalGeometry + emitter index ----------
    struct LocalGeometry { // render/geometry.h:33-37
        f3 position, normal;
        f2 texcoord;
    };
    // Geometry::GetHitLocalGeometry, render/geometry.h:60-100 (TriMesh, Sphere) and :176-180 (twosided)
    void hit_local_geometry(const orc_hit &h, f3 ray_o, f3 ray_d, LocalGeometry &g) const {
        const Instance &in = instances[h.inst];
        g.texcoord = f2{ 0.f, 0.f }; // DEFINED: the reference leaves texcoord untouched for meshes without uvs
        if (in.shape_type == ORC_SHAPE_SPHERE) {
            g.position = ray_o + h.t * ray_d;
            f3 local_pos = obj_to_world_point(g.position, in.inv); // world->object; centre 0, radius 1
            g.texcoord = B::sphere_texcoord(normalize(local_pos));
            g.normal = normalize(normal_obj_to_world(local_pos, in.inv));
            if (flip_sphere) g.normal *= -1.f;
        } else {
            const MeshData &md = meshes[in.mesh];
            uint32_t v0 = md.idx[h.prim * 3], v1 = md.idx[h.prim * 3 + 1], v2 = md.idx[h.prim * 3 + 2];
            auto P = [&](uint32_t v) { return f3{ md.pos[v * 3], md.pos[v * 3 + 1], md.pos[v * 3 + 2] }; };
            f3 p0 = P(v0), p1 = P(v1), p2 = P(v2);
            float bx = h.u, by = h.v;
            g.position = (1.f - bx - by) * p0 + bx * p1 + by * p2;
            g.position = obj_to_world_point(g.position, in.xf);
            f3 n;
            if (!md.nrm.empty()) {
                auto N = [&](uint32_t v) { return f3{ md.nrm[v * 3], md.nrm[v * 3 + 1], md.nrm[v * 3 + 2] }; };
                n = (1.f - bx - by) * N(v0) + bx * N(v1) + by * N(v2);
            } else {
                n = cross(p1 - p0, p2 - p0);
            }
            g.normal = normalize(normal_obj_to_world(n, in.inv));
            if (mesh_flips(in)) g.normal *= -1.f;
            if (!md.uv.empty()) {
                auto T = [&](uint32_t v) { return f2{ md.uv[v * 2], md.uv[v * 2 + 1] }; };
                g.texcoord = (1.f - bx - by) * T(v0) + bx * T(v1) + by * T(v2);
                if (in.flip_tex) g.texcoord.y = 1.f - g.texcoord.y;
            }
        }
        if (dot(-ray_d, g.normal) < 0.f && in.mat.twosided) g.normal = -g.normal;
    }

    // ---------- the integrator ----------
    struct PixelOut {
        f3 radiance, albedo, normal;
        float test;
        uint32_t closest_rays, shadow_rays;
    };
    const orc_emitter &emitter_at(int idx) const { return idx == (int)areas.size() ? env : areas[idx]; }

    PixelOut render_pixel(uint32_t x, uint32_t y, uint32_t random_seed, int depth_limit) const {
        PixelOut out{};
        const uint32_t w = film_w, h = film_h;
        const uint32_t pixel_index = y * w + x;
        f3 throughput = mk3(1.f), radiance = mk3(0.f), env_radiance = mk3(0.f);
        float env_pdf = 0.f;
        uint32_t rng = B::rng_init(4, pixel_index, random_seed); // main.cu:55

        float jx = B::rng_next(rng), jy = B::rng_next(rng); // :58 (x first: nvcc evaluates left to right)
        float sx = (static_cast<float>(x) + jx) / static_cast<float>(w);
        float sy = (static_cast<float>(y) + jy) / static_cast<float>(h);
        const float *s2c = cam.sample_to_camera.e, *c2w = cam.camera_to_world.e;
        f4 pf{ sx, sy, 0.f, 1.f };
        f4 d4{ dot(f4{ s2c[0], s2c[1], s2c[2], s2c[3] }, pf), dot(f4{ s2c[4], s2c[5], s2c[6], s2c[7] }, pf),
               dot(f4{ s2c[8], s2c[9], s2c[10], s2c[11] }, pf), dot(f4{ s2c[12], s2c[13], s2c[14], s2c[15] }, pf) }; // :67
        {
            float inv = 1.0f / d4.w; // :69  d /= d.w
            d4 = f4{ d4.x * inv, d4.y * inv, d4.z * inv, d4.w * inv };
            d4.w = 0.f;                                              // :70
            float inv_len = 1.0f / sqrtf(dot(d4, d4));                // :71
            d4 = f4{ d4.x * inv_len, d4.y * inv_len, d4.z * inv_len, d4.w * inv_len };
        }
        f3 ray_direction = normalize(f3{ dot(f4{ c2w[0], c2w[1], c2w[2], c2w[3] }, d4), dot(f4{ c2w[4], c2w[5], c2w[6], c2w[7] }, d4),
                                         dot(f4{ c2w[8], c2w[9], c2w[10], c2w[11] }, d4) }); // :73
        f3 ray_origin{ c2w[3], c2w[7], c2w[11] };                                            // :75-78

        bool done = false;
        LocalGeometry geo{};
        orc_local_bsdf bsdf{};
        int emitter_index = -1;
        auto trace = [&](f3 o, f3 d) { // optixTrace(closest) + __closesthit__default / __miss__default
            ++out.closest_rays;
            orc_hit hit = trace_closest(o, d, 0.001f, 1e16f);
            if (hit.inst < 0) { // main.cu:199-215
                if (has_env) {
                    f3 rd = normalize(d);
                    f3 rad;
                    B::emitter_eval(env, o + rd, mk3(0.f), f2{ 0.f, 0.f }, o, rad, env_pdf);
                    env_radiance = rad;
                }
                done = true;
                return;
            }
            const Instance &in = instances[hit.inst]; // main.cu:220-234
            hit_local_geometry(hit, o, d, geo);
            emitter_index = in.emitter_offset >= 0 ? in.emitter_offset + hit.prim : -1;
            bsdf = get_local_bsdf<B>(in.mat, geo.texcoord);
        };
        trace(ray_origin, ray_direction); // :80-85

        int depth = 0;
        if (!done) { // :90-102
            if (emitter_index >= 0) radiance += B::emitter_radiance(areas[emitter_index], geo.texcoord);
            out.albedo = local_albedo(bsdf);
            out.normal = geo.normal;
        }
        out.test = B::rng_next(rng); // :104 — consumes one draw

        while (!done) { // :106-187
            ++depth;
            if (depth >= depth_limit) break;
            float rr = depth > 2 ? 0.95 : 1.0;
            if (B::rng_next(rng) > rr) break;
            throughput /= rr;

            { // direct light sampling :117-144
                float sel = B::rng_next(rng);
                float e0 = B::rng_next(rng), e1 = B::rng_next(rng);
                int ei = B::select_emitter(areas.data(), (int)areas.size(), has_env, sel);
                if (ei >= 0) {
                    const orc_emitter &em = emitter_at(ei);
                    orc_emit_sample es;
                    B::emitter_sample_direct(em, geo.position, geo.normal, f2{ e0, e1 }, es);
                    // DEFINED: when pdf == 0 the reference traces a ray with an uninitialised tmax and
                    // then discards the result (IsZero(f*pdf)); nothing is traced here.
                    if (es.pdf != 0.f) {
                        f3 wi{ es.wi[0], es.wi[1], es.wi[2] };
                        ++out.shadow_rays;
                        bool occluded = trace_any(geo.position, wi, 0.0001f, es.distance - 0.0001f);
                        if (!occluded) {
                            f3 f;
                            float pdf;
                            B::bsdf_eval(bsdf, B::to_local(wi, geo.normal), B::to_local(-ray_direction, geo.normal), f, pdf);
                            if (!B::is_zero(f * es.pdf)) {
                                float NoL = dot(geo.normal, wi);
                                if (NoL > 0.f) {
                                    float mis = es.is_delta ? 1.f : B::mis_weight(es.pdf, pdf);
                                    float pdf_e = es.pdf * em.select_probability;
                                    radiance += throughput * f3{ es.radiance[0], es.radiance[1], es.radiance[2] } * f * NoL * mis / pdf_e;
                                }
                            }
                        }
                    }
                }
            }
            { // bsdf sampling :146-186
                orc_bsdf_result bs{};
                B::bsdf_sample(bsdf, B::to_local(-ray_direction, geo.normal), rng, bs);
                f3 bf{ bs.f[0], bs.f[1], bs.f[2] }, bwi{ bs.wi[0], bs.wi[1], bs.wi[2] };
                if (B::is_zero(bf * fabsf(bwi.z)) || B::is_zero(bs.pdf)) break;
                throughput *= bf * fabsf(bwi.z) / bs.pdf;
                ray_origin = geo.position;
                ray_direction = B::to_world(bwi, geo.normal);
                trace(ray_origin, ray_direction);
                if (done) {
                    float mis = B::mis_weight(bs.pdf, env_pdf);
                    env_radiance *= throughput * mis;
                    break;
                }
                if (emitter_index >= 0) {
                    const orc_emitter &em = areas[emitter_index];
                    f3 rad;
                    float pdf;
                    B::emitter_eval(em, geo.position, geo.normal, geo.texcoord, ray_origin, rad, pdf);
                    if (!B::is_zero(pdf)) {
                        float mis = (bs.sampled_type & kLobeDelta) ? 1.f : B::mis_weight(bs.pdf, pdf * em.select_probability);
                        radiance += throughput * rad * mis;
                    }
                }
            }
        }
        radiance += env_radiance; // :188
        out.radiance = radiance;
        return out;
    }

    // PTPass::OnRun repeated n_frames times (pt_pass.cpp:39-57) with main.cu:190-196 accumulation.
    void render(uint32_t first_seed, uint32_t n_frames, uint32_t sample_cnt0, int depth_limit, int accumulate, int threads, float *accum4,
                float *frame4, float *albedo3, float *normal3, float *test1, uint64_t *ray_counts) const {
        const uint32_t w = film_w, h = film_h;
        if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
        std::atomic<uint32_t> next_row{ 0 };
        std::atomic<uint64_t> n_closest{ 0 }, n_shadow{ 0 };
        auto worker = [&]() {
            uint64_t lc = 0, ls = 0;
            for (;;) {
                uint32_t y = next_row.fetch_add(1);
                if (y >= h) break;
                for (uint32_t x = 0; x < w; ++x) {
                    uint32_t pi = y * w + x;
                    for (uint32_t fr = 0; fr < n_frames; ++fr) {
                        PixelOut po = render_pixel(x, y, first_seed + fr, depth_limit);
                        lc += po.closest_rays, ls += po.shadow_rays;
                        uint32_t sample_cnt = accumulate ? sample_cnt0 + fr : 0; // pt_pass.cpp:55
                        f3 rad = po.radiance;
                        if (accumulate && sample_cnt > 0) { // main.cu:190-194
                            const float t = 1.f / (sample_cnt + 1.f);
                            f3 pre{ accum4[pi * 4], accum4[pi * 4 + 1], accum4[pi * 4 + 2] };
                            rad = lerp(pre, rad, t);
                        }
                        accum4[pi * 4] = rad.x, accum4[pi * 4 + 1] = rad.y, accum4[pi * 4 + 2] = rad.z, accum4[pi * 4 + 3] = 1.f;
                        if (frame4) frame4[pi * 4] = rad.x, frame4[pi * 4 + 1] = rad.y, frame4[pi * 4 + 2] = rad.z, frame4[pi * 4 + 3] = 1.f;
                        if (albedo3) albedo3[pi * 3] = po.albedo.x, albedo3[pi * 3 + 1] = po.albedo.y, albedo3[pi * 3 + 2] = po.albedo.z;
                        if (normal3) normal3[pi * 3] = po.normal.x, normal3[pi * 3 + 1] = po.normal.y, normal3[pi * 3 + 2] = po.normal.z;
                        if (test1) test1[pi] = po.test;
                    }
                }
            }
            n_closest += lc, n_shadow += ls;
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
        worker();
        for (auto &t : pool) t.join();
        if (ray_counts) ray_counts[0] = n_closest, ray_counts[1] = n_shadow;
    }
};
}// namespace orc
#endif
